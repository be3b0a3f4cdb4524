// K7 — conv2d (NHWC bf16, stride 1) as IMPLICIT GEMM on the tcgen05 mainloop, forward/dgrad and wgrad.
//
// Replaces the cuDNN convolution the reference reaches through the user's oblivious train_fn
// (SURVEY.md §2.4 row C10 "conv2d implicit-GEMM (ResNet-50)"; reference call site: user code under
// maggy/core/executors/trial_executor.py:126-140 — the ASHA ResNet-50 config of BASELINE.json).
//
// No im2col buffer exists anywhere: one M tile is a box of 128 output pixels (bw x bh x bn, chosen on the host so
// that it tiles the feature map with the least padding) and, for filter tap (r, s), the A operand of that tile is
// the SAME box shifted by (r - pad_h, s - pad_w) in the input tensor. A 4-D tiled TMA load of the shifted box
// lands in shared memory in exactly the 128B-swizzled K-major layout `tcgen05.mma` wants, and TMA's out-of-bounds
// zero fill (negative coordinates included) is the convolution's zero padding. The K loop runs over
// (tap, 64-channel chunk); the weight operand [Kout, R*S*C] is a plain 2-D K-major tile.
//
//   fprop : Y[n,p,q,k]  = sum_{r,s,c} X[n, p+r-pad_h, q+s-pad_w, c] * Wt[k,r,s,c]
//   dgrad : the same kernel on dY with the flipped / transposed filter (prepared by the caller)
//   wgrad : dW[k,r,s,c] = sum_{n,p,q} dY[n,p,q,k] * X[n, p+r-pad_h, q+s-pad_w, c]
//           GEMM with K = pixels: both operands are MN-major boxes of 128 pixels x 64 channels (dY unshifted,
//           X shifted by the tap), split over pixel ranges, fp32 `red.global.add.v4` into dW.
#include "../include/mg_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace {

using namespace mg;

constexpr int BLOCK_M = 128;
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 256;
constexpr int kEpiWarpBase = 4;

struct ConvParams {
  int N, H, W, C;      // input  [N, H, W, C]
  int Kout, R, S;      // filter [Kout, R, S, C]
  int P, Q;            // output [N, P, Q, Kout]
  int pad_h, pad_w;
  int bw, bh, bn;      // pixel box of one tile: bw * bh * bn == 128
  int tiles_w, tiles_h, tiles_n;
  int num_m_tiles, num_n_blocks, c_chunks;
  int relu;            // fprop: fused ReLU on the output
  int splits;          // wgrad: pixel-range splits
  int tiles_per_split;
  float* dw;           // wgrad output, fp32 [Kout, R, S, C]
};

__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                    pack_bf16x2(v[6], v[7]));
}

__device__ __forceinline__ void tile_origin(const ConvParams& p, int m_tile, int& q0, int& p0, int& n0) {
  const int tw = m_tile % p.tiles_w;
  const int t = m_tile / p.tiles_w;
  q0 = tw * p.bw;
  p0 = (t % p.tiles_h) * p.bh;
  n0 = (t / p.tiles_h) * p.bn;
}

// ------------------------------------------------------------------------------------------------
// fprop / dgrad
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES>
struct FpropSmem {
  static constexpr int kABytes = BLOCK_M * 128;
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutOffset = STAGES * kStageBytes;
  static constexpr int kOutBytes = 2 * BLOCK_M * 128;  // two [128 pixels x 64 channels] staging tiles
  static constexpr int kBarOffset = kOutOffset + kOutBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1)
conv2d_fprop_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                            const __grid_constant__ CUtensorMap tma_y, const ConvParams p) {
  pdl_launch_dependents();
  using L = FpropSmem<BLOCK_N, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_blocks;
  const int num_k_blocks = p.R * p.S * p.c_chunks;
  constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_y);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.num_n_blocks, n_blk = tile - m_tile * p.num_n_blocks;
        int q0, p0, n0;
        tile_origin(p, m_tile, q0, p0, n0);
        int tap = 0, cc = 0, r = 0, s = 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          // the tile's pixel box, shifted by the filter tap; halo outside the image is zero-filled by TMA
          tma_load_4d(sa, &tma_x, &full_bar[stage], cc * 64, q0 + s - p.pad_w, p0 + r - p.pad_h, n0);
          tma_load_2d(sb, &tma_w, &full_bar[stage], kb * 64, n_blk * BLOCK_N);
          if (++cc == p.c_chunks) {
            cc = 0;
            ++tap;
            if (++s == p.S) {
              s = 0;
              ++r;
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, 0, 0, BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < 64 / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
            umma_f16_ss(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full[as]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp_idx >= kEpiWarpBase) {
    // ============================ epilogue ================================
    // TMEM -> registers -> bf16 -> 128B-swizzled [128 pixels x 64 channels] staging tile -> ONE 4-D TMA store
    // of the pixel box (all four warps fill the tile, so they meet on named barrier 1).
    const int q = warp_idx - kEpiWarpBase;
    const int epi_tid = threadIdx.x - kEpiWarpBase * 32;
    uint8_t* out_stage = smem + L::kOutOffset;
    int obuf = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_blocks, n_blk = tile - m_tile * p.num_n_blocks;
      int q0, p0, n0;
      tile_origin(p, m_tile, q0, p0, n0);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      constexpr int kChunks = BLOCK_N / 64;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + c * 64, r0);
        tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
        tmem_ld_wait();
        if (c == kChunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        float v[64];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = __uint_as_float(r0[i]);
          v[32 + i] = __uint_as_float(r1[i]);
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (epi_tid == 0) tma_store_wait_read<1>();  // the store issued two chunks ago has drained this tile
        named_bar_sync(1, 128);
        const int row = q * 32 + lane;
        uint8_t* buf = out_stage + obuf * (BLOCK_M * 128) + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(buf + ((j ^ (row & 7)) << 4)) = pack8(&v[j * 8]);
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (epi_tid == 0) {
          tma_store_4d(&tma_y, out_stage + obuf * (BLOCK_M * 128), n_blk * BLOCK_N + c * 64, q0, p0, n0);
          tma_store_commit();
        }
        obuf ^= 1;
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (epi_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES>
struct WgradSmem {
  static constexpr int kChunkBytes = 128 * 128;  // 128 pixels (K) x 64 channels
  static constexpr int kABytes = 2 * kChunkBytes;
  static constexpr int kBBytes = (BLOCK_N / 64) * kChunkBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1)
conv2d_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_x,
                            const ConvParams p) {
  pdl_launch_dependents();
  using L = WgradSmem<BLOCK_N, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m_blocks = (p.Kout + BLOCK_M - 1) / BLOCK_M;
  const int taps = p.R * p.S;
  // unit = ((split * taps + tap) * num_m_blocks + m_blk) * num_n_blocks + n_blk
  const int num_units = p.splits * taps * num_m_blocks * p.num_n_blocks;
  constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);

  auto decode = [&](int unit, int& split, int& tap, int& m_blk, int& n_blk) {
    n_blk = unit % p.num_n_blocks;
    int t = unit / p.num_n_blocks;
    m_blk = t % num_m_blocks;
    t /= num_m_blocks;
    tap = t % taps;
    split = t / taps;
  };

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_dy);
    tma_prefetch_desc(&tma_x);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp_idx == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        int split, tap, m_blk, n_blk;
        decode(unit, split, tap, m_blk, n_blk);
        const int r = tap / p.S, s = tap - r * p.S;
        const int t0 = split * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.num_m_tiles);
        for (int t = t0; t < t1; ++t) {
          int q0, p0, n0;
          tile_origin(p, t, q0, p0, n0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          // A: dY box (K = 128 pixels) x 128 output channels, two 64-channel chunks (MN-major); channels past
          // Kout are zero-filled
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_4d(sa + c * L::kChunkBytes, &tma_dy, &full_bar[stage], m_blk * BLOCK_M + c * 64, q0, p0, n0);
          // B: the same pixel box of X shifted by the tap
#pragma unroll
          for (int c = 0; c < BLOCK_N / 64; ++c)
            tma_load_4d(sb + c * L::kChunkBytes, &tma_x, &full_bar[stage], n_blk * BLOCK_N + c * 64,
                        q0 + s - p.pad_w, p0 + r - p.pad_h, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, 1, 1, BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        int split, tap, m_blk, n_blk;
        decode(unit, split, tap, m_blk, n_blk);
        const int t0 = split * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.num_m_tiles);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < 128 / UMMA_K; ++k) {
            // MN-major: 64-channel chunks kChunkBytes apart (LBO), 8-pixel groups 1024 B apart (SBO), 16 pixels
            // (= 2048 B) per UMMA_K step
            const uint64_t da = make_smem_desc_sw128(sa + k * (UMMA_K * 128), L::kChunkBytes, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * (UMMA_K * 128), L::kChunkBytes, 1024);
            umma_f16_ss(tmem_d, da, db, idesc, (t > t0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (t == t1 - 1) umma_commit(&tmem_full[as]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp_idx >= kEpiWarpBase) {
    const int q = warp_idx - kEpiWarpBase;
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      int split, tap, m_blk, n_blk;
      decode(unit, split, tap, m_blk, n_blk);
      const int row = m_blk * BLOCK_M + q * 32 + lane;  // output channel
      const bool row_ok = row < p.Kout;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
      float* out = p.dw + (static_cast<long long>(row) * taps + tap) * p.C + n_blk * BLOCK_N;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32b_x32(taddr + c * 32, rr);
        tmem_ld_wait();
        if (row_ok) {
          if (p.splits > 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c * 32 + j * 4),
                           "f"(__uint_as_float(rr[j * 4])), "f"(__uint_as_float(rr[j * 4 + 1])),
                           "f"(__uint_as_float(rr[j * 4 + 2])), "f"(__uint_as_float(rr[j * 4 + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(out + c * 32 + j * 4) =
                  make_float4(__uint_as_float(rr[j * 4]), __uint_as_float(rr[j * 4 + 1]),
                              __uint_as_float(rr[j * 4 + 2]), __uint_as_float(rr[j * 4 + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// explicit path for strided / thin-channel convolutions: im2col gather, col2im gather-sum (own kernels; the GEMMs
// in between run on gemm_tcgen05.cu)
// ------------------------------------------------------------------------------------------------
// cols[m, (r*S + s)*C + c] = X[n, p*stride + r - pad, q*stride + s - pad, c];  m = (n*P + p)*Q + q; rows are Kp
// elements long (Kp >= R*S*C, multiple of 8; the tail is zero)
template <int VEC>
__global__ void im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ cols, int N,
                                   int H, int W, int C, int R, int S, int P, int Q, int stride, int pad_h, int pad_w,
                                   int Kp) {
  pdl_prologue();
  const long long vec_per_row = Kp / VEC;
  const long long total = static_cast<long long>(N) * P * Q * vec_per_row;
  const int RSC = R * S * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / vec_per_row;
    const int kk = static_cast<int>(i - m * vec_per_row) * VEC;
    const int q = static_cast<int>(m % Q);
    const long long t = m / Q;
    const int pp = static_cast<int>(t % P);
    const int n = static_cast<int>(t / P);
    if (VEC == 8) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (kk < RSC) {
        const int tap = kk / C, c = kk - tap * C;
        const int r = tap / S, s = tap - r * S;
        const int h = pp * stride + r - pad_h, w = q * stride + s - pad_w;
        if (h >= 0 && h < H && w >= 0 && w < W)
          v = ld_nc_v4(x + ((static_cast<long long>(n) * H + h) * W + w) * C + c);
      }
      *reinterpret_cast<uint4*>(cols + m * Kp + kk) = v;
    } else {
      __nv_bfloat16 v = __float2bfloat16(0.f);
      if (kk < RSC) {
        const int tap = kk / C, c = kk - tap * C;
        const int r = tap / S, s = tap - r * S;
        const int h = pp * stride + r - pad_h, w = q * stride + s - pad_w;
        if (h >= 0 && h < H && w >= 0 && w < W) v = x[((static_cast<long long>(n) * H + h) * W + w) * C + c];
      }
      cols[m * Kp + kk] = v;
    }
  }
}

// Thin-channel variant (C % 8 != 0: the 3-channel 7x7/2 stem). The scalar kernel above wrote the 244 MB column matrix
// of the ResNet-50 stem at 0.3 TB/s (0.81 ms, twice per step). Here a CTA stages the input patch of kTileQ output
// pixels of one output row in shared memory (R input rows, contiguous in NHWC), then every thread assembles whole
// 16-byte vectors of the column matrix from it, so the global writes are full-width and coalesced and each input
// element is fetched from global once per tile. off[k] maps a column index to its patch offset (-1: zero tail).
constexpr int kTileQ = 32;
__global__ void __launch_bounds__(256)
im2col_thin_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ cols, int N, int H, int W, int C,
                   int R, int S, int P, int Q, int stride, int pad_h, int pad_w, int Kp) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int span = ((kTileQ - 1) * stride + S) * C;  // patch row length in elements
  short* off = reinterpret_cast<short*>(smem_raw);                                        // [Kp]
  __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(smem_raw + ((Kp * 2 + 15) & ~15));  // [R][span]
  const int RSC = R * S * C, SC = S * C;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    const int r = k / SC;
    off[k] = k < RSC ? static_cast<short>(r * span + (k - r * SC)) : static_cast<short>(-1);
  }
  const int qtiles = (Q + kTileQ - 1) / kTileQ;
  const long long tiles = static_cast<long long>(N) * P * qtiles;
  const int vec_per_row = Kp >> 3;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int qt = static_cast<int>(tile % qtiles);
    const long long t = tile / qtiles;
    const int pp = static_cast<int>(t % P), n = static_cast<int>(t / P);
    const int q0 = qt * kTileQ;
    const int w0 = q0 * stride - pad_w;  // first input column of the patch (may be negative)
    __syncthreads();                     // the previous tile's readers are done (and off[] is written)
    for (int i = threadIdx.x; i < R * span; i += blockDim.x) {
      const int r = i / span, j = i - r * span;
      const int h = pp * stride + r - pad_h;
      const int e = w0 * C + j;  // element offset inside input row h
      __nv_bfloat16 v = __float2bfloat16(0.f);
      if (h >= 0 && h < H && e >= 0 && e < W * C) v = x[(static_cast<long long>(n) * H + h) * W * C + e];
      patch[i] = v;
    }
    __syncthreads();
    const int npix = min(kTileQ, Q - q0);
    const long long m0 = (static_cast<long long>(n) * P + pp) * Q + q0;
    uint4* dst = reinterpret_cast<uint4*>(cols + m0 * Kp);
    for (int v = threadIdx.x; v < npix * vec_per_row; v += blockDim.x) {
      const int pix = v / vec_per_row, kv = v - pix * vec_per_row;
      const int base = pix * stride * C;
      uint32_t w[4];
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        const int o0 = off[kv * 8 + 2 * e2], o1 = off[kv * 8 + 2 * e2 + 1];
        const uint32_t lo = o0 >= 0 ? __bfloat16_as_ushort(patch[o0 + base]) : 0u;
        const uint32_t hi = o1 >= 0 ? __bfloat16_as_ushort(patch[o1 + base]) : 0u;
        w[e2] = lo | (hi << 16);
      }
      dst[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// dX[n,h,w,c..c+VEC) = sum over the taps (r,s) that hit pixel (h,w): dcols[m(n,p,q), (r*S+s)*C + c]
template <int VEC>
__global__ void col2im_nhwc_kernel(const __nv_bfloat16* __restrict__ dcols, __nv_bfloat16* __restrict__ dx, int N,
                                   int H, int W, int C, int R, int S, int P, int Q, int stride, int pad_h, int pad_w,
                                   int Kp) {
  pdl_prologue();
  const int cv = C / VEC;
  const long long total = static_cast<long long>(N) * H * W * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv) * VEC;
    long long t = i / cv;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int h = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    for (int r = 0; r < R; ++r) {
      const int hp = h + pad_h - r;
      if (hp < 0 || hp % stride != 0) continue;
      const int pp = hp / stride;
      if (pp >= P) continue;
      for (int s = 0; s < S; ++s) {
        const int wq = w + pad_w - s;
        if (wq < 0 || wq % stride != 0) continue;
        const int q = wq / stride;
        if (q >= Q) continue;
        const long long m = (static_cast<long long>(n) * P + pp) * Q + q;
        const __nv_bfloat16* src = dcols + m * Kp + (r * S + s) * C + c;
        if (VEC == 8) {
          const uint4 u = ld_nc_v4(src);
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), d = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
          acc[0] += a.x; acc[1] += a.y; acc[2 % VEC] += b.x; acc[3 % VEC] += b.y;
          acc[4 % VEC] += d.x; acc[5 % VEC] += d.y; acc[6 % VEC] += e.x; acc[7 % VEC] += e.y;
        } else {
          acc[0] += __bfloat162float(*src);
        }
      }
    }
    __nv_bfloat16* dst = dx + ((static_cast<long long>(n) * H + h) * W + w) * C + c;
    if (VEC == 8) *reinterpret_cast<uint4*>(dst) = pack8(acc);
    else *dst = __float2bfloat16(acc[0]);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// NHWC bf16 tensor [n, h, w, c] viewed as a 4-D map {c, w, h, n}; box {64, bw, bh, bn}; 128B swizzle
int make_map_nhwc(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int bw, int bh, int bn) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2,
                           static_cast<cuuint64_t>(h) * w * c * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

int make_map_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint32_t box_inner,
                uint32_t box_outer) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// pixel box (power-of-two sides, product 128) that tiles [N, P, Q] with the least padding; ties go to the widest
// box (longer contiguous runs per TMA row group)
void pick_box(int N, int P, int Q, int& bw, int& bh, int& bn) {
  long long best = -1;
  for (int w = 1; w <= 128; w *= 2) {
    for (int h = 1; w * h <= 128; h *= 2) {
      const int n = 128 / (w * h);
      const long long vol = static_cast<long long>((Q + w - 1) / w) * w * ((P + h - 1) / h) * h * ((N + n - 1) / n) * n;
      const bool better = best < 0 || vol < best || (vol == best && (w > bw || (w == bw && h > bh)));
      if (better) {
        best = vol;
        bw = w;
        bh = h;
        bn = n;
      }
    }
  }
}

void fill_geometry(ConvParams& p, int N, int H, int W, int C, int Kout, int R, int S, int pad_h, int pad_w) {
  p.N = N; p.H = H; p.W = W; p.C = C; p.Kout = Kout; p.R = R; p.S = S;
  p.pad_h = pad_h; p.pad_w = pad_w;
  p.P = H + 2 * pad_h - R + 1;
  p.Q = W + 2 * pad_w - S + 1;
  pick_box(N, p.P, p.Q, p.bw, p.bh, p.bn);
  p.tiles_w = (p.Q + p.bw - 1) / p.bw;
  p.tiles_h = (p.P + p.bh - 1) / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.c_chunks = C / 64;
  p.relu = 0;
  p.splits = 1;
  p.tiles_per_split = p.num_m_tiles;
  p.dw = nullptr;
}

template <int BLOCK_N, int STAGES>
int launch_fprop(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const ConvParams& p,
                 int max_ctas, cudaStream_t stream) {
  using L = FpropSmem<BLOCK_N, STAGES>;
  auto kern = conv2d_fprop_tcgen05_kernel<BLOCK_N, STAGES>;
  constexpr int smem_bytes = L::kTotal + 1024;
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_blocks;
  const int grid = tiles < max_ctas ? tiles : max_ctas;
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kNumThreads), smem_bytes, stream, 1, mx, mw, my, p);
  if (e != cudaSuccess) fprintf(stderr, "maggy_b200 conv fprop: launch failed: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

template <int BLOCK_N, int STAGES>
int launch_wgrad(const CUtensorMap& mdy, const CUtensorMap& mx, const ConvParams& p, int max_ctas,
                 cudaStream_t stream) {
  using L = WgradSmem<BLOCK_N, STAGES>;
  auto kern = conv2d_wgrad_tcgen05_kernel<BLOCK_N, STAGES>;
  constexpr int smem_bytes = L::kTotal + 1024;
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  const int units = p.splits * p.R * p.S * ((p.Kout + BLOCK_M - 1) / BLOCK_M) * p.num_n_blocks;
  const int grid = units < max_ctas ? units : max_ctas;
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kNumThreads), smem_bytes, stream, 1, mdy, mx, p);
  if (e != cudaSuccess) fprintf(stderr, "maggy_b200 conv wgrad: launch failed: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // namespace

extern "C" {

// Y[N,P,Q,Kout] = conv(X[N,H,W,C], Wt[Kout,R,S,C]), stride 1, zero padding (pad_h, pad_w). Requires C % 64 == 0 and
// Kout % 64 == 0. relu != 0 fuses max(.,0) into the epilogue.
int mg_conv2d_fprop(const void* x, const void* w, void* y, int N, int H, int W, int C, int Kout, int R, int S,
                    int pad_h, int pad_w, int relu, int sm_limit, void* stream) {
  if (C % 64 != 0 || Kout % 64 != 0 || N <= 0) return -1;
  ConvParams p;
  fill_geometry(p, N, H, W, C, Kout, R, S, pad_h, pad_w);
  if (p.P <= 0 || p.Q <= 0) return -2;
  p.relu = relu;
  int max_ctas = num_sms();
  if (sm_limit > 0 && sm_limit < max_ctas) max_ctas = sm_limit;
  // Output-channel tile width by a wave model: cost = waves * (k-blocks * max(bn, 128) + fixed). A 64-wide tile costs as
  // much per k-block as a 128-wide one (the A operand's shared-memory reads bound it), and a small pixel count with
  // 256-wide tiles leaves SMs idle: the 7x7 layers of ResNet-50 (25 pixel tiles x 512 channels) fill 50 SMs with
  // bn = 256 and 100 with bn = 128 at half the time per tile.
  int block_n = 64;
  {
    const long long kblocks = static_cast<long long>(R) * S * (C / 64);
    double best = 0.0;
    for (int bn : {256, 128, 64}) {
      if (Kout % bn != 0) continue;
      const long long tiles = static_cast<long long>(p.num_m_tiles) * (Kout / bn);
      const long long waves = (tiles + max_ctas - 1) / max_ctas;
      const double cost = static_cast<double>(waves) * (static_cast<double>(kblocks) * (bn > 128 ? bn : 128) + 1500.0);
      if (best == 0.0 || cost < best * 0.97) {
        best = cost;
        block_n = bn;
      }
    }
  }
  p.num_n_blocks = Kout / block_n;
  CUtensorMap mx, mw, my;
  int rc = make_map_nhwc(&mx, x, N, H, W, C, p.bw, p.bh, p.bn);
  if (rc) return rc;
  rc = make_map_2d(&mw, w, static_cast<uint64_t>(R) * S * C, Kout, 64, block_n);
  if (rc) return rc;
  rc = make_map_nhwc(&my, y, N, p.P, p.Q, Kout, p.bw, p.bh, p.bn);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (block_n == 256) return launch_fprop<256, 4>(mx, mw, my, p, max_ctas, s);
  if (block_n == 128) return launch_fprop<128, 6>(mx, mw, my, p, max_ctas, s);
  return launch_fprop<64, 8>(mx, mw, my, p, max_ctas, s);
}

// dW[Kout,R,S,C] (fp32) = sum over pixels of dY[N,P,Q,Kout] (x) shifted X[N,H,W,C]. With splits > 1 the caller
// must have zeroed dw (partial sums are accumulated with fp32 reductions). splits == 0: chosen here.
int mg_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int Kout, int R, int S,
                    int pad_h, int pad_w, int splits, int sm_limit, void* stream) {
  if (C % 64 != 0 || Kout % 64 != 0 || N <= 0) return -1;
  ConvParams p;
  fill_geometry(p, N, H, W, C, Kout, R, S, pad_h, pad_w);
  if (p.P <= 0 || p.Q <= 0) return -2;
  const int block_n = (C % 256 == 0) ? 256 : (C % 128 == 0 ? 128 : 64);
  p.num_n_blocks = C / block_n;
  p.dw = dw;
  int max_ctas = num_sms();
  if (sm_limit > 0 && sm_limit < max_ctas) max_ctas = sm_limit;
  const int base_units = R * S * ((Kout + BLOCK_M - 1) / BLOCK_M) * p.num_n_blocks;
  if (splits <= 0) {
    splits = (2 * max_ctas + base_units - 1) / base_units;  // ~2 waves of units
    const int max_splits = (p.num_m_tiles + 7) / 8;          // at least 8 pixel tiles per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  if (splits > p.num_m_tiles) splits = p.num_m_tiles;
  p.tiles_per_split = (p.num_m_tiles + splits - 1) / splits;
  p.splits = (p.num_m_tiles + p.tiles_per_split - 1) / p.tiles_per_split;  // no empty splits
  CUtensorMap mdy, mx;
  int rc = make_map_nhwc(&mdy, dy, N, p.P, p.Q, Kout, p.bw, p.bh, p.bn);
  if (rc) return rc;
  rc = make_map_nhwc(&mx, x, N, H, W, C, p.bw, p.bh, p.bn);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.splits > 1) {
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * static_cast<size_t>(Kout) * R * S * C, s);
    if (e != cudaSuccess) return -500 - static_cast<int>(e);
  }
  if (block_n == 256) return launch_wgrad<256, 2>(mdy, mx, p, max_ctas, s);
  if (block_n == 128) return launch_wgrad<128, 3>(mdy, mx, p, max_ctas, s);
  return launch_wgrad<64, 4>(mdy, mx, p, max_ctas, s);
}

// cols[N*P*Q, Kp] from X[N,H,W,C] (any stride / padding / channel count; Kp % 8 == 0, Kp >= R*S*C)
int mg_im2col_nhwc(const void* x, void* cols, int N, int H, int W, int C, int R, int S, int stride, int pad_h,
                   int pad_w, int Kp, void* stream) {
  const int P = (H + 2 * pad_h - R) / stride + 1, Q = (W + 2 * pad_w - S) / stride + 1;
  if (P <= 0 || Q <= 0 || Kp % 8 != 0 || Kp < R * S * C) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (C % 8 == 0);
  const long long total = static_cast<long long>(N) * P * Q * (vec ? Kp / 8 : Kp);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (vec)
    launch_pdl(im2col_nhwc_kernel<8>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1, 
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(cols), N, H, W, C, R, S, P, Q, stride, pad_h,
        pad_w, Kp);
  else {
    const size_t smem = ((Kp * 2 + 15) & ~15) + static_cast<size_t>(R) * ((kTileQ - 1) * stride + S) * C * 2;
    const long long tiles = static_cast<long long>(N) * P * ((Q + kTileQ - 1) / kTileQ);
    static const bool scalar = getenv("MAGGY_B200_IM2COL_SCALAR") != nullptr;  // the first kernel, for A/B runs
    if (!scalar && smem <= 48 * 1024 && R * ((kTileQ - 1) * stride + S) * C < 32768 && Kp % 8 == 0) {
      const int grid = static_cast<int>(tiles < 148LL * 8 ? tiles : 148LL * 8);
      launch_pdl(im2col_thin_kernel, dim3(grid), dim3(256), smem, s, 1, static_cast<const __nv_bfloat16*>(x),
                                                 static_cast<__nv_bfloat16*>(cols), N, H, W, C, R, S, P, Q, stride,
                                                 pad_h, pad_w, Kp);
    } else {
      launch_pdl(im2col_nhwc_kernel<1>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1, 
          static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(cols), N, H, W, C, R, S, P, Q, stride,
          pad_h, pad_w, Kp);
    }
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// dX[N,H,W,C] from dcols[N*P*Q, Kp]
int mg_col2im_nhwc(const void* dcols, void* dx, int N, int H, int W, int C, int R, int S, int stride, int pad_h,
                   int pad_w, int Kp, void* stream) {
  const int P = (H + 2 * pad_h - R) / stride + 1, Q = (W + 2 * pad_w - S) / stride + 1;
  if (P <= 0 || Q <= 0) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (C % 8 == 0);
  const long long total = static_cast<long long>(N) * H * W * (vec ? C / 8 : C);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (vec)
    launch_pdl(col2im_nhwc_kernel<8>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1, 
        static_cast<const __nv_bfloat16*>(dcols), static_cast<__nv_bfloat16*>(dx), N, H, W, C, R, S, P, Q, stride,
        pad_h, pad_w, Kp);
  else
    launch_pdl(col2im_nhwc_kernel<1>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, 1, 
        static_cast<const __nv_bfloat16*>(dcols), static_cast<__nv_bfloat16*>(dx), N, H, W, C, R, S, P, Q, stride,
        pad_h, pad_w, Kp);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // extern "C"
