// K5 — persistent, warp-specialised tcgen05/TMEM GEMM for sm_100a (bf16 in, fp32 accumulate).
//
//   C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
//
// Replaces the cuBLAS GEMM the reference reaches through the user's oblivious train_fn
// (reference call site: user code under
// maggy/core/executors/torch_dist_executor.py:145-159).
//
// Roles (256 threads, 1 CTA / SM, persistent over output tiles):
//   warp 0      TMA producer     cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1      MMA issuer       one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into TMEM
//   warp 2      TMEM allocator   tcgen05.alloc / dealloc (2 accumulator stages)
//   warps 4..7  epilogue         tcgen05.ld TMEM->regs, fused bias / GELU / dGELU / residual, global store
//
// Both operands can be K-major or MN-major (set in the instruction descriptor + smem descriptor), so
// forward (x W^T), dgrad (dY W) and wgrad (dY^T x) all run on this one kernel without transposes.
#include "../include/mg_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace {

using namespace mg;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 256;
constexpr int kEpiWarpBase = 4;

enum Epi : int {
  EPI_NONE = 0,
  EPI_BIAS = 1,           // out = acc + bias[n]
  EPI_BIAS_GELU = 2,      // aux = acc + bias[n]; out = gelu(aux)
  EPI_DGELU = 3,          // out = acc * gelu'(aux[m,n]); a non-null `bias` is an fp32 [4 * m_blocks][N] buffer that
                          // receives the column sums of `out` per 32-row warp slice (the bias gradient, see below)
  EPI_BIAS_RESIDUAL = 4,  // out = acc + bias[n] + aux[m,n]
};

struct GemmParams {
  int M, N, K;
  void* C;
  long long ldc;
  const __nv_bfloat16* bias;  // [N] or null
  __nv_bfloat16* aux;         // [M, ld_aux] (written for BIAS_GELU, read for DGELU / RESIDUAL)
  long long ld_aux;
  int epi;
  int out_f32;     // 0: bf16 output, 1: fp32 output
  int accumulate;  // 1: C += result
  int tma_store;   // 1: epilogue stages bf16 tiles in smem and writes them with TMA (fast path)
  int splits;      // split-K factor (>1: fp32 partials go to `ws`, reduced by splitk_reduce_kernel)
  int kb_per_split;
  float* ws;       // [splits][M][N] fp32
  int num_m_blocks, num_n_blocks, num_k_blocks;
  float alpha;     // fp8 variant: output scale applied to the accumulator, times the optional device scalars:
  const float* scale_a;  // per-tensor dequantisation scale of A (device pointer or null)
  const float* scale_b;
};

constexpr int kStageOutBytes = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutOffset = STAGES * kStageBytes;  // 1024-aligned (stage sizes are multiples of 16 KB)
  static constexpr int kBarOffset = kOutOffset + kStageOutBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16;
};

// tile rasterisation: groups of 8 M-blocks sweep N so concurrently resident CTAs share B tiles in L2
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
  constexpr int GROUP_M = 8;
  int per_group = GROUP_M * num_n;
  int g = tile / per_group;
  int first_m = g * GROUP_M;
  int gsz = min(num_m - first_m, GROUP_M);
  int r = tile - g * per_group;
  m_blk = first_m + (r % gsz);
  n_blk = r / gsz;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                    pack_bf16x2(v[6], v[7]));
}

// MC = true: thread-block clusters of 2 CTAs that work on vertically adjacent output tiles (same n-block,
// m-blocks 2i and 2i+1). Each CTA fetches only HALF of the shared B tile and TMA-multicasts it into both
// CTAs' shared memory, which cuts the L2 -> smem operand traffic per CTA from 48 KB to 32 KB per k-block.
// MODE 2 (PAIR): the two CTAs of a cluster form a tcgen05 CTA pair: ONE `tcgen05.mma.cta_group::2` (M = 256,
// issued by the leader CTA) multiplies both CTAs' A tiles with a B tile whose N columns are split between the
// two shared memories, so each SM stages and reads only half of B — the shared-memory bandwidth (TMA fill +
// MMA operand reads) per SM drops from ~190 to ~128 B/clk, which is what caps the 1-CTA kernel near 67 %
// tensor-pipe utilisation (profiles/gemm_ncu_v2.md).
// FP8 = true: both operands are e4m3 bytes (K-major only), `tcgen05.mma.kind::f8f6f4` with K = 32 per instruction;
// a k-block is still one 128-byte swizzle row (128 elements), so the smem layout, descriptors and pipeline are
// byte-for-byte those of the bf16 kernel and the tensor pipe does twice the FLOPs per staged byte.
template <int BLOCK_N, int STAGES, int A_MAJOR, int B_MAJOR, int MODE, bool FP8 = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                         const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_c,
                         const __grid_constant__ CUtensorMap tma_aux, const GemmParams p) {
  constexpr bool MC = MODE == 1, PAIR = MODE == 2, CL = MODE != 0;
  pdl_launch_dependents();  // the next kernel may be scheduled behind this grid's last wave (see mg_ptx.cuh)
  constexpr int KB_ELEMS = FP8 ? 128 : BLOCK_K;  // elements per k-block (always 128 bytes)
  static_assert(!FP8 || (A_MAJOR == 0 && B_MAJOR == 0), "fp8 operands are K-major");
  using L = SmemLayout<(PAIR ? BLOCK_N / 2 : BLOCK_N), STAGES>;  // PAIR: each CTA stages half of the B tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16B aligned by the ABI: round up to the 1024B swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if constexpr (CL) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  // iteration domain: one work unit per CTA (MC = false) or per CTA pair (MC = true)
  const int m_groups = CL ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks;
  const int num_tiles = m_groups * p.num_n_blocks;
  const int num_units = num_tiles * p.splits;
  const int unit_begin = CL ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int unit_step = CL ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  auto coords = [&](int tile, int& m_blk, int& n_blk) {
    int mg;
    tile_coords(tile, m_groups, p.num_n_blocks, mg, n_blk);
    m_blk = CL ? mg * 2 + static_cast<int>(crank) : mg;
  };
  constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;  // allocation must be a power of two

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (p.tma_store) {
      tma_prefetch_desc(&tma_c);
      if (p.epi == EPI_BIAS_GELU) tma_prefetch_desc(&tma_aux);
    }
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);              // PAIR: only the leader's is used (it expects both CTAs' bytes)
      mbar_init(&empty_bar[s], MC ? 2 : 1);    // MC: the peer multicasts into this stage too
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], PAIR ? 8 : 4);  // PAIR: the peer's epilogue warps arrive here too
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    if constexpr (PAIR) {
      tmem_alloc_2cta(tmem_ptr_smem, kTmemCols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_ptr_smem, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL) cluster_sync();  // peer barriers must be initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // barriers, TMEM and descriptors were set up under the predecessor's tail; its output is needed from here on

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit_begin; unit < num_units; unit += unit_step) {
        const int tile = unit % num_tiles, split = unit / num_tiles;
        int m_blk, n_blk;
        coords(tile, m_blk, n_blk);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          const int k0 = kb * KB_ELEMS;
          if constexpr (PAIR) {
            // every byte of both CTAs' loads is accounted on the LEADER's barrier (only the leader issues MMAs)
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            // the peer's bytes may land before this expect_tx (transiently negative tx-count is legal); they
            // cannot land a phase early because the peer's slot is only released by the leader's MMA commit
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            if constexpr (A_MAJOR == 0) {
              tma_load_2d_2sm(sa, &tma_a, lead_bar, k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c)
                tma_load_2d_2sm(sa + c * (BLOCK_K * 128), &tma_a, lead_bar, m0 + c * 64, k0);
            }
            if constexpr (B_MAJOR == 0) {
              tma_load_2d_2sm(sb, &tma_b, lead_bar, k0, n0 + static_cast<int>(crank) * (BLOCK_N / 2));
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_N / 128; ++c)
                tma_load_2d_2sm(sb + c * (BLOCK_K * 128), &tma_b, lead_bar,
                                n0 + (static_cast<int>(crank) * (BLOCK_N / 128) + c) * 64, k0);
            }
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          if constexpr (A_MAJOR == 0) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d(sa + c * (BLOCK_K * 128), &tma_a, &full_bar[stage], m0 + c * 64, k0);
          }
          if constexpr (!MC) {
            if constexpr (B_MAJOR == 0) {
              tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_N / 64; ++c)
                tma_load_2d(sb + c * (BLOCK_K * 128), &tma_b, &full_bar[stage], n0 + c * 64, k0);
            }
          } else {
            // my half of the B tile, delivered to both CTAs of the cluster (same smem offset, same barrier)
            if constexpr (B_MAJOR == 0) {
              tma_load_2d_mc(sb + crank * (L::kBBytes / 2), &tma_b, &full_bar[stage], k0,
                             n0 + static_cast<int>(crank) * (BLOCK_N / 2), 0x3);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_N / 128; ++c) {
                const int cc = static_cast<int>(crank) * (BLOCK_N / 128) + c;
                tma_load_2d_mc(sb + cc * (BLOCK_K * 128), &tma_b, &full_bar[stage], n0 + cc * 64, k0, 0x3);
              }
            }
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
    if (lane == 0 && (!PAIR || crank == 0)) {
      constexpr uint32_t kFmt = FP8 ? kFmtE4M3 : kFmtBF16;
      constexpr uint32_t idesc = make_idesc(kFmt, kFmt, A_MAJOR, B_MAJOR, PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = unit_begin; unit < num_units; unit += unit_step) {
        const int split = unit / num_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: 8-row groups 1024B apart (SBO); advance 32B along the swizzled row per UMMA_K.
            // MN-major: 64-element MN chunks BLOCK_K*128B apart (LBO), 8-k groups 1024B apart (SBO);
            //           advance 16 k-rows = 2048B per UMMA_K.
            const uint64_t da = (A_MAJOR == 0)
                                    ? make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024)
                                    : make_smem_desc_sw128(sa + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
            const uint64_t db = (B_MAJOR == 0)
                                    ? make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024)
                                    : make_smem_desc_sw128(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
            const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
            if constexpr (FP8 && PAIR) umma_f8_ss_2cta(tmem_d, da, db, idesc, acc);
            else if constexpr (FP8) umma_f8_ss(tmem_d, da, db, idesc, acc);
            else if constexpr (PAIR) umma_f16_ss_2cta(tmem_d, da, db, idesc, acc);
            else umma_f16_ss(tmem_d, da, db, idesc, acc);
          }
          if constexpr (PAIR) umma_commit_2cta_mc(&empty_bar[stage], 0x3);
          else if constexpr (MC) umma_commit_mc(&empty_bar[stage], 0x3);  // release the slot in BOTH CTAs
          else umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) {
            if constexpr (PAIR) umma_commit_2cta_mc(&tmem_full[as], 0x3);  // both CTAs' epilogues may start
            else umma_commit(&tmem_full[as]);
          }
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
    const int q = warp_idx - kEpiWarpBase;  // == warp_idx % 4: TMEM lane quarter this warp may read
    auto arrive_tmem_empty = [&](uint64_t* bar) {
      if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));  // the leader's MMA warp waits there
      else mbar_arrive(bar);
    };
    uint8_t* out_stage = smem + L::kOutOffset + q * 8192;  // two 4 KB buffers for this warp
    int obuf = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = unit_begin; unit < num_units; unit += unit_step) {
      const int tile = unit % num_tiles, split = unit / num_tiles;
      int m_blk, n_blk;
      coords(tile, m_blk, n_blk);
      const int row0 = m_blk * BLOCK_M + q * 32;
      const int row = row0 + lane;
      const int n0 = n_blk * BLOCK_N;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;

      if (p.tma_store) {
        // ---------------- fast path: bf16 tile -> swizzled smem -> TMA store -----------------------
        const bool has_bias = p.bias != nullptr && (p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU ||
                                                    p.epi == EPI_BIAS_RESIDUAL);
        const bool reads_aux = p.epi == EPI_DGELU || p.epi == EPI_BIAS_RESIDUAL;
        // dGELU + bias gradient: the tile this warp stages for the TMA store is read back column-wise (32 rows x 2
        // columns per lane, conflict-free under the 128-byte swizzle) and the sums go to row (4 * m_blk + q) of the
        // partial buffer — the separate column-sum pass over d_pre (2 x 400 MB per GPT-2 layer) disappears.
        float* cs_row = (p.epi == EPI_DGELU && p.bias != nullptr)
                            ? reinterpret_cast<float*>(const_cast<__nv_bfloat16*>(p.bias)) +
                                  static_cast<long long>(m_blk * 4 + q) * p.N
                            : nullptr;
        const __nv_bfloat16* aux_row = p.aux + static_cast<long long>(row) * p.ld_aux;
        uint4 aux_cur[8], aux_nxt[8];
        if (reads_aux && row_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + j * 8 + 8 <= p.N) aux_cur[j] = *reinterpret_cast<const uint4*>(aux_row + n0 + j * 8);
        }
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        constexpr int kChunks = BLOCK_N / 64;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = n0 + c * 64;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr + c * 64, r0);
          tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
          // overlap the TMEM read with the bias / next-chunk aux loads
          uint4 bias_v[8];
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              bias_v[j] = (col0 + j * 8 + 8 <= p.N)
                              ? __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + j * 8))
                              : make_uint4(0, 0, 0, 0);
          }
          if (reads_aux && row_ok && c + 1 < kChunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (col0 + 64 + j * 8 + 8 <= p.N)
                aux_nxt[j] = *reinterpret_cast<const uint4*>(aux_row + col0 + 64 + j * 8);
          }
          tmem_ld_wait();
          if (c == kChunks - 1 || col0 + 64 >= p.N) {
            // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tmem_empty(&tmem_empty[as]);
          }
          float v[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = __uint_as_float(r0[i]);
            v[32 + i] = __uint_as_float(r1[i]);
          }
          if constexpr (FP8) {
            const float alpha = p.alpha * (p.scale_a ? __ldg(p.scale_a) : 1.f) * (p.scale_b ? __ldg(p.scale_b) : 1.f);
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] *= alpha;
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float b[8];
              unpack8(bias_v[j], b);
#pragma unroll
              for (int t = 0; t < 8; ++t) v[j * 8 + t] += b[t];
            }
          }
          auto stage_and_store = [&](const CUtensorMap* map) {
            if (lane == 0) tma_store_wait_read<1>();  // the buffer we are about to fill is free again
            __syncwarp();
            uint8_t* buf = out_stage + obuf * 4096 + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(buf + ((j ^ (lane & 7)) << 4)) = pack8(&v[j * 8]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(map, out_stage + obuf * 4096, col0, row0);
              tma_store_commit();
            }
            if (cs_row != nullptr && map == &tma_c) {
              const uint8_t* base = out_stage + obuf * 4096;
              const int rows_valid = p.M - row0;  // rows past M hold whatever gelu'(unloaded aux) produced: skip them
              float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                if (r < rows_valid) {
                  const uint32_t w = *reinterpret_cast<const uint32_t*>(
                      base + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
                  const float2 f = unpack_bf16x2(w);
                  s0 += f.x;
                  s1 += f.y;
                }
              }
              if (col0 + 2 * lane + 2 <= p.N) *reinterpret_cast<float2*>(cs_row + col0 + 2 * lane) = make_float2(s0, s1);
            }
            obuf ^= 1;
          };
          if (p.epi == EPI_BIAS_GELU) {
            stage_and_store(&tma_aux);  // pre-activation, needed by backward
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = gelu_tanh(v[i]);
          } else if (reads_aux) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a[8];
              unpack8(aux_cur[j], a);
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                if (p.epi == EPI_DGELU) v[j * 8 + t] *= dgelu_tanh(a[t]);
                else v[j * 8 + t] += a[t];
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) aux_cur[j] = aux_nxt[j];
          }
          stage_and_store(&tma_c);
        }
      } else {
        // ---------------- general path: fp32 / accumulate / split-K partials, direct stores ---------
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        float* ws_base = p.splits > 1 ? p.ws + static_cast<long long>(split) * p.M * p.N : nullptr;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if constexpr (FP8) {
            const float alpha = p.alpha * (p.scale_a ? __ldg(p.scale_a) : 1.f) * (p.scale_b ? __ldg(p.scale_b) : 1.f);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= alpha;
          }
          if (p.bias != nullptr &&
              (p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU || p.epi == EPI_BIAS_RESIDUAL)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (col0 + j * 8 + 8 <= p.N) {
                float b[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias + col0 + j * 8)), b);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[j * 8 + t] += b[t];
              }
            }
          }
          if (row_ok) {
            if (p.epi == EPI_BIAS_GELU) {
              __nv_bfloat16* auxp = p.aux + static_cast<long long>(row) * p.ld_aux + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + j * 8 + 8 <= p.N) *reinterpret_cast<uint4*>(auxp + j * 8) = pack8(&v[j * 8]);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
            } else if (p.epi == EPI_DGELU || p.epi == EPI_BIAS_RESIDUAL) {
              const __nv_bfloat16* auxp = p.aux + static_cast<long long>(row) * p.ld_aux + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (col0 + j * 8 + 8 <= p.N) {
                  float a[8];
                  unpack8(*reinterpret_cast<const uint4*>(auxp + j * 8), a);
#pragma unroll
                  for (int t = 0; t < 8; ++t) {
                    if (p.epi == EPI_DGELU) v[j * 8 + t] *= dgelu_tanh(a[t]);
                    else v[j * 8 + t] += a[t];
                  }
                }
              }
            }
            if (p.splits > 1 || p.out_f32) {
              float* cp = p.splits > 1
                              ? ws_base + static_cast<long long>(row) * p.N + col0
                              : reinterpret_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (col0 + j * 4 + 4 <= p.N) {
                  float4 o = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                  if (p.accumulate && p.splits == 1) {
                    float4 old = *reinterpret_cast<float4*>(cp + j * 4);
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  *reinterpret_cast<float4*>(cp + j * 4) = o;
                }
              }
            } else {
              __nv_bfloat16* cp =
                  reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (col0 + j * 8 + 8 <= p.N) {
                  if (p.accumulate) {
                    float o[8];
                    unpack8(*reinterpret_cast<const uint4*>(cp + j * 8), o);
#pragma unroll
                    for (int t = 0; t < 8; ++t) v[j * 8 + t] += o[t];
                  }
                  *reinterpret_cast<uint4*>(cp + j * 8) = pack8(&v[j * 8]);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tmem_empty(&tmem_empty[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (p.tma_store && lane == 0) tma_store_wait<0>();  // smem must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL) cluster_sync();  // nobody leaves while the peer may still multicast into / arrive on this CTA
  if (warp_idx == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2cta(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// out[i] (+)= sum_s ws[s][i], one thread per output vector walking the splits: the right shape for FEW splits over a
// LARGE output (GPT-2 weight gradients: 2-8 splits x 0.6-2.4 M floats; the split-lane kernel below took 2.5x longer there)
__global__ void __launch_bounds__(256)
splitk_reduce_flat_kernel(const float* __restrict__ ws, int splits, long long n, void* __restrict__ out,
                          int out_f32, int accumulate) {
  pdl_prologue();
  const long long nvec = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(ws)[i];
    for (int s = 1; s < splits; ++s) {
      const float4 t = reinterpret_cast<const float4*>(ws + static_cast<long long>(s) * n)[i];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (out_f32) {
      float4* o = reinterpret_cast<float4*>(out) + i;
      if (accumulate) {
        const float4 t = *o;
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      *o = acc;
    } else {
      uint2* o = reinterpret_cast<uint2*>(out) + i;
      if (accumulate) {
        const uint2 t = *o;
        const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
        acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
      }
      *o = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
    }
  }
}

// out[i] (+)= sum_s ws[s][i]   (split-K second pass; n multiple of 4). A CTA is 32 output vectors x 8 split lanes:
// lane y sums splits y, y+8, ... (up to four loads in flight), the eight partial sums meet in shared memory. With one
// split per SM (thin weight gradients: up to 148 partials) a single thread walking all the splits was a chain of
// L2 latencies (12.6 us per launch in the ResNet step).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long n, void* __restrict__ out,
                     int out_f32, int accumulate) {
  pdl_prologue();
  const long long nvec = n >> 2;
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  __shared__ float4 red[8][32];
  for (long long i0 = static_cast<long long>(blockIdx.x) * 32; i0 < nvec; i0 += static_cast<long long>(gridDim.x) * 32) {
    const long long i = i0 + x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nvec) {
      int s = y;
      for (; s + 24 < splits; s += 32) {
        float4 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = reinterpret_cast<const float4*>(ws + static_cast<long long>(s + 8 * k) * n)[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc.x += t[k].x; acc.y += t[k].y; acc.z += t[k].z; acc.w += t[k].w; }
      }
      for (; s < splits; s += 8) {
        const float4 t = reinterpret_cast<const float4*>(ws + static_cast<long long>(s) * n)[i];
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
    }
    red[y][x] = acc;
    __syncthreads();
    if (y == 0 && i < nvec) {
#pragma unroll
      for (int k = 1; k < 8; ++k) {
        const float4 t = red[k][x];
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      if (out_f32) {
        float4* o = reinterpret_cast<float4*>(out) + i;
        if (accumulate) {
          float4 t = *o;
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        *o = acc;
      } else {
        uint2* o = reinterpret_cast<uint2*>(out) + i;
        if (accumulate) {
          uint2 t = *o;
          float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
          acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
        }
        *o = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
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

// 2-D byte (e4m3) tensor map: `inner` contiguous bytes, `outer` rows of `ld` bytes; 128B swizzle.
int make_map_u8(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                uint32_t box_outer);

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of `ld` elements; 128B swizzle.
int make_map_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                  uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

int make_map_u8(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

int g_num_sms = 0;

struct Maps {
  CUtensorMap a, b, c, aux;
};

template <int BLOCK_N, int STAGES, int A_MAJOR, int B_MAJOR, int MODE, bool FP8 = false>
int launch(const Maps& m, const GemmParams& p, int max_ctas, cudaStream_t stream) {
  constexpr bool MC = MODE != 0;  // launched as 2-CTA clusters
  using L = SmemLayout<(MODE == 2 ? BLOCK_N / 2 : BLOCK_N), STAGES>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, STAGES, A_MAJOR, B_MAJOR, MODE, FP8>;
  constexpr int smem_bytes = L::kTotal + 1024;
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  const int m_groups = MC ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks;
  int units = m_groups * p.num_n_blocks * p.splits;
  int grid;
  if (MC) {
    int clusters = max_ctas / 2;
    grid = 2 * (units < clusters ? units : clusters);
  } else {
    grid = units < max_ctas ? units : max_ctas;
  }
  cudaError_t stale = cudaGetLastError();  // do not blame this launch for an older, unchecked error
  if (stale != cudaSuccess)
    fprintf(stderr, "maggy_b200 gemm: clearing stale CUDA error before launch: %s\n", cudaGetErrorString(stale));
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kNumThreads), smem_bytes, stream, MC ? 2 : 1, m.a, m.b, m.c, m.aux, p);
  if (e != cudaSuccess) {
    int dev = -1;
    cudaGetDevice(&dev);
    fprintf(stderr,
            "maggy_b200 gemm: launch failed: %s (BN=%d stages=%d A=%d B=%d MC=%d grid=%d smem=%d M=%d N=%d K=%d "
            "splits=%d dev=%d stream=%p)\n",
            cudaGetErrorString(e), BLOCK_N, STAGES, A_MAJOR, B_MAJOR, MODE, grid, smem_bytes, p.M, p.N, p.K, p.splits,
            dev, static_cast<void*>(stream));
  }
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

template <int BLOCK_N, int STAGES, int MODE = 0>
int dispatch_major(int a_major, int b_major, const Maps& m, const GemmParams& p, int max_ctas,
                   cudaStream_t s) {
  if (a_major == 0 && b_major == 0) return launch<BLOCK_N, STAGES, 0, 0, MODE>(m, p, max_ctas, s);
  if (a_major == 0 && b_major == 1) return launch<BLOCK_N, STAGES, 0, 1, MODE>(m, p, max_ctas, s);
  if (a_major == 1 && b_major == 0) return launch<BLOCK_N, STAGES, 1, 0, MODE>(m, p, max_ctas, s);
  return launch<BLOCK_N, STAGES, 1, 1, MODE>(m, p, max_ctas, s);
}

int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

int pick_block_n(int M, int N, int K = 0) {
  // Wide tiles have the best arithmetic intensity per byte of smem/L2 traffic, so 128-wide tiles are only
  // for narrow outputs. Between 256 and 192 pick the one that wastes less of the last wave of the
  // persistent grid (e.g. M=32768, N=768: 768 tiles of 256 = 5.2 waves -> 6; 1024 tiles of 192 = 6.9 -> 7,
  // i.e. 1344 vs 1536 column-units of work).
  if (N < 192) return 128;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const long long mb = (M + BLOCK_M - 1) / BLOCK_M;
  auto cost = [&](int bn) {
    long long tiles = mb * ((N + bn - 1) / bn);
    long long waves = (tiles + sms - 1) / sms;
    return waves * bn;
  };
  if (N < 256) return 192;
  // Measured on B200 (profiles/gemm_bench_*.log): 192-wide tiles lose more to their lower arithmetic
  // intensity (L2 -> smem traffic per FLOP) than they gain from the better last wave, even at N = 768, so
  // 256 stays the default and 192 is only taken when it saves more than a quarter of the work units.
  // Under-filled grids (few row blocks, e.g. the 7x7 layers of ResNet-50: 25 row blocks x 512 columns = 50 tiles on
  // 148 SMs): 128-wide tiles double the tile count at half the time per tile. Taken only when that saves more than a
  // quarter of the wave-time, so the large training GEMMs (hundreds of row blocks) keep their 256-wide tiles.
  // Long reductions are excluded: there split-K over 256-wide CTA-pair tiles fills the chip with better operand reuse
  // (the GPT-2 weight gradients lost 8 % end to end when this rule caught them).
  if (K > 0 && K < 4096 && cost(128) * 4 < cost(256) * 3 && cost(128) * 4 < cost(192) * 3) return 128;
  return cost(192) * 4 < cost(256) * 3 ? 192 : 256;
}

}  // namespace

// How many K-splits mg_gemm_bf16 will use for (M, N, K) when asked for `splits == 0` (auto): the caller
// needs it to size the fp32 workspace (splits * M * N floats). 1 means no split-K.
extern "C" int mg_gemm_auto_splits(int M, int N, int K, int block_n) {
  if (block_n == 0) block_n = pick_block_n(M, N, K);
  const long long m_blocks = (M + BLOCK_M - 1) / BLOCK_M, n_blocks = (N + block_n - 1) / block_n;
  long long tiles = m_blocks * n_blocks;
  int kb = (K + BLOCK_K - 1) / BLOCK_K;
  int sms = num_sms() > 0 ? num_sms() : 148;
  if (tiles * 2 > sms || kb < 32) return 1;
  // Work units are scheduled on CTA pairs for 256-wide tiles (one unit = two vertically adjacent tiles), on single
  // CTAs otherwise. Model, in SM cycles:  T(sp) = waves(sp) * ceil(kb / sp) * t_kb  +  reduce(sp)
  //   t_kb      one k-block of one tile on the tensor pipe (8192 dense bf16 FLOP/clk/SM): 2 * block_n cycles
  //   reduce    second pass over sp fp32 partials of the output (~3.4 KB/clk of HBM at 6.5 TB/s) + one launch
  // e.g. the GPT-2 qkv weight gradient (18 x 3 tiles, kb = 512..1024) gets 5 splits = 135 pair units = 1.82 waves
  // (91 % busy) instead of 2 splits = 54 units (73 %); the ResNet stem / layer-1 weight gradients (1-8 tiles, K = 200 k
  // ... 800 k pixels) get one split per SM — the 16-split cap of round 1 ran them on 16-32 SMs at 0.2x of the HBM
  // bound (profiles/resnet_r2/). At least 8 k-blocks (512 of K) per split.
  const bool pairs = block_n == 256 && m_blocks >= 2;
  const long long units1 = pairs ? ((m_blocks + 1) / 2) * n_blocks : tiles;
  const long long slots = pairs ? sms / 2 : sms;
  long long max_s = kb / 8;
  if (max_s > slots) max_s = slots;
  const double t_kb = 2.0 * block_n;
  const double out_bytes = 8.0 * static_cast<double>(M) * static_cast<double>(N);  // one fp32 partial: write + read
  int best = 1;
  double best_t = 0.0;
  for (int sp = 1; sp <= max_s; ++sp) {
    const long long units = units1 * sp;
    const long long waves = (units + slots - 1) / slots;
    const long long kbs = (kb + sp - 1) / sp;
    double t = static_cast<double>(waves) * static_cast<double>(kbs) * t_kb;
    if (sp > 1) t += 4000.0 + sp * out_bytes / 3400.0;
    if (sp == 1 || t < best_t * 0.97) {  // a further split has to buy at least 3 %
      best_t = t;
      best = sp;
    }
  }
  return best;
}

// C[M,N] = epi( sum_k A(m,k) B(n,k) ).
//   a_major == 0: A stored [M][K] (row stride lda);  a_major == 1: A stored [K][M] (row stride lda)
//   b_major == 0: B stored [N][K] (row stride ldb);  b_major == 1: B stored [K][N] (row stride ldb)
// Alignment: every pointer 16B aligned, lda/ldb/ldc/ld_aux multiples of 8 elements, N % 8 == 0.
// block_n: 0 = auto, else 128, 192 or 256.  sm_limit: 0 = all SMs, else cap on persistent CTAs.
// splits: 1 = none, >1 = split-K into `ws` (fp32, splits*M*N floats) + reduce pass (epi must be NONE).
extern "C" int mg_gemm_bf16(const void* A, const void* B, void* C, const void* bias, void* aux,
                            int M, int N, int K, long long lda, long long ldb, long long ldc,
                            long long ld_aux, int a_major, int b_major, int epi, int out_f32,
                            int accumulate, int block_n, int sm_limit, int splits, float* ws,
                            void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((N % 8) != 0 || (lda % 8) != 0 || (ldb % 8) != 0 || (ldc % (out_f32 ? 4 : 8)) != 0) return -1;
  if (block_n == 0) block_n = pick_block_n(M, N, K);
  if (splits < 1) splits = 1;
  if (splits > 1 && (epi != EPI_NONE || ws == nullptr || ldc != N)) return -3;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.aux = reinterpret_cast<__nv_bfloat16*>(aux);
  p.ld_aux = ld_aux;
  p.epi = epi; p.out_f32 = out_f32; p.accumulate = accumulate;
  p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (N + block_n - 1) / block_n;
  p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  if (splits > p.num_k_blocks) splits = p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + splits - 1) / splits;
  p.splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.ws = ws;
  p.alpha = 1.f;
  p.scale_a = nullptr;
  p.scale_b = nullptr;
  p.tma_store = (!out_f32 && !accumulate && p.splits == 1) ? 1 : 0;
  if (epi == EPI_DGELU && bias != nullptr && !p.tma_store) return -5;  // column sums exist on the staged path only

  Maps m;
  int rc;
  if (a_major == 0) rc = make_map_bf16(&m.a, A, K, M, lda, BLOCK_K, BLOCK_M);
  else              rc = make_map_bf16(&m.a, A, M, K, lda, 64, BLOCK_K);
  if (rc) return rc;
  // cluster-multicast variant: 128x256 tiles, at least two m-blocks, enabled by MAGGY_B200_GEMM_MC (default on)
  static const int mc_env = [] {
    const char* v = getenv("MAGGY_B200_GEMM_MC");
    return v == nullptr ? 1 : atoi(v);
  }();
  // Measured (profiles/gemm_bench_mc.md): multicast pays when the mainloop is long enough for operand traffic
  // to dominate (8192^3: 1327 -> 1469 TFLOP/s) and is neutral on the K = 768 shapes, which are bound by the
  // power-capped tensor pipe + epilogue; so it is only used from K = 2048 up (mc_env = 2 forces it everywhere).
  const bool use_mc = mc_env != 0 && block_n == 256 && p.num_m_blocks >= 2 && (p.num_k_blocks >= 32 || mc_env == 2);
  // CTA-pair (cta_group::2) variant — the default for 256-wide tiles (MAGGY_B200_GEMM_PAIR=0 disables it);
  // takes precedence over the multicast variant
  static const int pair_env = [] {
    const char* v = getenv("MAGGY_B200_GEMM_PAIR");
    return v == nullptr ? 1 : atoi(v);
  }();
  const bool use_pair = pair_env != 0 && block_n == 256 && p.num_m_blocks >= 2;
  if (b_major == 0) rc = make_map_bf16(&m.b, B, K, N, ldb, BLOCK_K, (use_mc || use_pair) ? block_n / 2 : block_n);
  else              rc = make_map_bf16(&m.b, B, N, K, ldb, 64, BLOCK_K);
  if (rc) return rc - 1000;
  m.c = m.a;
  m.aux = m.a;
  if (p.tma_store) {
    rc = make_map_bf16(&m.c, C, N, M, ldc, 64, 32);
    if (rc) return rc - 2000;
    if (epi == EPI_BIAS_GELU) {
      rc = make_map_bf16(&m.aux, aux, N, M, ld_aux, 64, 32);
      if (rc) return rc - 3000;
    }
  }

  int max_ctas = (sm_limit > 0 && sm_limit < num_sms()) ? sm_limit : num_sms();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (block_n == 256 && use_pair) rc = dispatch_major<256, 6, 2>(a_major, b_major, m, p, max_ctas, s);
  else if (block_n == 256 && use_mc) rc = dispatch_major<256, 4, 1>(a_major, b_major, m, p, max_ctas, s);
  else if (block_n == 256) rc = dispatch_major<256, 4>(a_major, b_major, m, p, max_ctas, s);
  else if (block_n == 192) rc = dispatch_major<192, 4>(a_major, b_major, m, p, max_ctas, s);
  else if (block_n == 128) rc = dispatch_major<128, 6>(a_major, b_major, m, p, max_ctas, s);
  else return -4;
  if (rc) return rc;
  if (p.splits > 1) {
    long long n = static_cast<long long>(M) * N;
    long long nvec = n >> 2;
    cudaError_t e;
    if (p.splits <= 8) {
      long long blocks = (nvec + 255) / 256;
      int grid = static_cast<int>(blocks < 148LL * 8 ? blocks : 148LL * 8);
      e = launch_pdl(splitk_reduce_flat_kernel, dim3(grid), dim3(256), 0, s, 1, ws, p.splits, n, C, out_f32, accumulate);
    } else {
      long long blocks = (nvec + 31) / 32;
      int grid = static_cast<int>(blocks < 148LL * 32 ? blocks : 148LL * 32);
      e = launch_pdl(splitk_reduce_kernel, dim3(grid), dim3(256), 0, s, 1, ws, p.splits, n, C, out_f32, accumulate);
    }
    if (e != cudaSuccess) return -500 - static_cast<int>(e);
  }
  return 0;
}

// fp8 (e4m3) GEMM: C[M,N] (bf16) = alpha * sum_k A(m,k) B(n,k) (+ bias[n]); A is [M][K] bytes (row stride lda),
// B is [N][K] bytes (row stride ldb); K % 16 == 0, lda / ldb multiples of 16. The accumulator is multiplied by
// alpha * (*scale_a) * (*scale_b): scale_a / scale_b are optional DEVICE scalars (the per-tensor dequantisation
// scales written by mg_quant_e4m3), so quantise -> GEMM needs no host round trip. epi: EPI_NONE or EPI_BIAS.
extern "C" int mg_gemm_fp8(const void* A, const void* B, void* C, const void* bias, int M, int N, int K,
                           long long lda, long long ldb, long long ldc, float alpha, const float* scale_a,
                           const float* scale_b, int epi, int sm_limit, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((N % 8) != 0 || (K % 16) != 0 || (lda % 16) != 0 || (ldb % 16) != 0 || (ldc % 8) != 0) return -1;
  if (epi != EPI_NONE && epi != EPI_BIAS) return -3;
  constexpr int block_n = 256;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.aux = nullptr;
  p.ld_aux = 0;
  p.epi = epi; p.out_f32 = 0; p.accumulate = 0;
  p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (N + block_n - 1) / block_n;
  p.num_k_blocks = (K + 127) / 128;
  p.kb_per_split = p.num_k_blocks;
  p.splits = 1;
  p.ws = nullptr;
  p.tma_store = 1;
  p.alpha = alpha;
  p.scale_a = scale_a;
  p.scale_b = scale_b;
  static const int pair_env = [] {
    const char* v = getenv("MAGGY_B200_GEMM_PAIR");
    return v == nullptr ? 1 : atoi(v);
  }();
  const bool use_pair = pair_env != 0 && p.num_m_blocks >= 2;
  Maps m;
  int rc = make_map_u8(&m.a, A, K, M, lda, 128, BLOCK_M);
  if (rc) return rc;
  rc = make_map_u8(&m.b, B, K, N, ldb, 128, use_pair ? block_n / 2 : block_n);
  if (rc) return rc - 1000;
  rc = make_map_bf16(&m.c, C, N, M, ldc, 64, 32);
  if (rc) return rc - 2000;
  m.aux = m.c;
  int max_ctas = (sm_limit > 0 && sm_limit < num_sms()) ? sm_limit : num_sms();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (use_pair) return launch<256, 6, 0, 0, 2, true>(m, p, max_ctas, s);
  return launch<256, 4, 0, 0, 0, true>(m, p, max_ctas, s);
}
