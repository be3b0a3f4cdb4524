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
  EPI_DGELU = 3,          // out = acc * gelu'(aux[m,n])
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
  int num_m_blocks, num_n_blocks, num_k_blocks;
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
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

template <int BLOCK_N, int STAGES, int A_MAJOR, int B_MAJOR>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                         const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  using L = SmemLayout<BLOCK_N, STAGES>;
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
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // 256 or 512: power of two

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
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

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          const int k0 = kb * BLOCK_K;
          if constexpr (A_MAJOR == 0) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d(sa + c * (BLOCK_K * 128), &tma_a, &full_bar[stage], m0 + c * 64, k0);
          }
          if constexpr (B_MAJOR == 0) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_N / 64; ++c)
              tma_load_2d(sb + c * (BLOCK_K * 128), &tma_b, &full_bar[stage], n0 + c * 64, k0);
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
      constexpr uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, A_MAJOR, B_MAJOR, BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
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
            umma_f16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[as]);
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
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const int n0 = n_blk * BLOCK_N;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
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
        const bool full_chunk = col0 + 32 <= p.N;
        if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU || p.epi == EPI_BIAS_RESIDUAL) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (col0 + j * 8 + 8 <= p.N) {
                uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + j * 8));
                float2 f0 = unpack_bf16x2(b.x), f1 = unpack_bf16x2(b.y), f2 = unpack_bf16x2(b.z),
                       f3 = unpack_bf16x2(b.w);
                v[j * 8 + 0] += f0.x; v[j * 8 + 1] += f0.y; v[j * 8 + 2] += f1.x; v[j * 8 + 3] += f1.y;
                v[j * 8 + 4] += f2.x; v[j * 8 + 5] += f2.y; v[j * 8 + 6] += f3.x; v[j * 8 + 7] += f3.y;
              }
            }
          }
        }
        if (row_ok) {
          if (p.epi == EPI_BIAS_GELU) {
            __nv_bfloat16* auxp = p.aux + static_cast<long long>(row) * p.ld_aux + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (col0 + j * 8 + 8 <= p.N) {
                uint4 o;
                o.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]);
                o.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
                o.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]);
                o.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
                *reinterpret_cast<uint4*>(auxp + j * 8) = o;
              }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
          } else if (p.epi == EPI_DGELU || p.epi == EPI_BIAS_RESIDUAL) {
            const __nv_bfloat16* auxp = p.aux + static_cast<long long>(row) * p.ld_aux + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (col0 + j * 8 + 8 <= p.N) {
                uint4 a = *reinterpret_cast<const uint4*>(auxp + j * 8);
                float2 f[4] = {unpack_bf16x2(a.x), unpack_bf16x2(a.y), unpack_bf16x2(a.z),
                               unpack_bf16x2(a.w)};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  if (p.epi == EPI_DGELU) {
                    v[j * 8 + 2 * t] *= dgelu_tanh(f[t].x);
                    v[j * 8 + 2 * t + 1] *= dgelu_tanh(f[t].y);
                  } else {
                    v[j * 8 + 2 * t] += f[t].x;
                    v[j * 8 + 2 * t + 1] += f[t].y;
                  }
                }
              }
            }
          }
          if (p.out_f32) {
            float* cp = reinterpret_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (full_chunk || col0 + j * 4 + 4 <= p.N) {
                float4 o = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                if (p.accumulate) {
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
              if (full_chunk || col0 + j * 8 + 8 <= p.N) {
                if (p.accumulate) {
                  uint4 old = *reinterpret_cast<const uint4*>(cp + j * 8);
                  float2 f0 = unpack_bf16x2(old.x), f1 = unpack_bf16x2(old.y),
                         f2 = unpack_bf16x2(old.z), f3 = unpack_bf16x2(old.w);
                  v[j * 8 + 0] += f0.x; v[j * 8 + 1] += f0.y; v[j * 8 + 2] += f1.x; v[j * 8 + 3] += f1.y;
                  v[j * 8 + 4] += f2.x; v[j * 8 + 5] += f2.y; v[j * 8 + 6] += f3.x; v[j * 8 + 7] += f3.y;
                }
                uint4 o;
                o.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]);
                o.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
                o.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]);
                o.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
                *reinterpret_cast<uint4*>(cp + j * 8) = o;
              }
            }
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

int g_num_sms = 0;

template <int BLOCK_N, int STAGES, int A_MAJOR, int B_MAJOR>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int max_ctas,
           cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, STAGES, A_MAJOR, B_MAJOR>;
  constexpr int smem_bytes = L::kTotal + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  int tiles = p.num_m_blocks * p.num_n_blocks;
  int grid = tiles < max_ctas ? tiles : max_ctas;
  kern<<<grid, kNumThreads, smem_bytes, stream>>>(ma, mb, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

template <int BLOCK_N, int STAGES>
int dispatch_major(int a_major, int b_major, const CUtensorMap& ma, const CUtensorMap& mb,
                   const GemmParams& p, int max_ctas, cudaStream_t s) {
  if (a_major == 0 && b_major == 0) return launch<BLOCK_N, STAGES, 0, 0>(ma, mb, p, max_ctas, s);
  if (a_major == 0 && b_major == 1) return launch<BLOCK_N, STAGES, 0, 1>(ma, mb, p, max_ctas, s);
  if (a_major == 1 && b_major == 0) return launch<BLOCK_N, STAGES, 1, 0>(ma, mb, p, max_ctas, s);
  return launch<BLOCK_N, STAGES, 1, 1>(ma, mb, p, max_ctas, s);
}

}  // namespace

// C[M,N] = epi( sum_k A(m,k) B(n,k) ).
//   a_major == 0: A stored [M][K] (row stride lda);  a_major == 1: A stored [K][M] (row stride lda)
//   b_major == 0: B stored [N][K] (row stride ldb);  b_major == 1: B stored [K][N] (row stride ldb)
// Alignment: every pointer 16B aligned, lda/ldb/ldc/ld_aux multiples of 8 elements, N % 8 == 0.
// block_n: 0 = auto, else 128 or 256.  sm_limit: 0 = all SMs, else cap on persistent CTAs.
extern "C" int mg_gemm_bf16(const void* A, const void* B, void* C, const void* bias, void* aux,
                            int M, int N, int K, long long lda, long long ldb, long long ldc,
                            long long ld_aux, int a_major, int b_major, int epi, int out_f32,
                            int accumulate, int block_n, int sm_limit, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((N % 8) != 0 || (lda % 8) != 0 || (ldb % 8) != 0 || (ldc % (out_f32 ? 4 : 8)) != 0) return -1;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (block_n == 0) {
    // prefer 128x256 tiles unless that leaves most SMs idle
    long long t256 = static_cast<long long>((M + 127) / 128) * ((N + 255) / 256);
    block_n = (N % 256 == 0 || t256 >= 2LL * g_num_sms) && N >= 256 ? 256 : 128;
    if (t256 < g_num_sms && N >= 128) block_n = 128;
  }
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

  CUtensorMap ma, mb;
  int rc;
  if (a_major == 0) rc = make_map_bf16(&ma, A, K, M, lda, BLOCK_K, BLOCK_M);
  else              rc = make_map_bf16(&ma, A, M, K, lda, 64, BLOCK_K);
  if (rc) return rc;
  if (b_major == 0) rc = make_map_bf16(&mb, B, K, N, ldb, BLOCK_K, block_n);
  else              rc = make_map_bf16(&mb, B, N, K, ldb, 64, BLOCK_K);
  if (rc) return rc - 1000;

  int max_ctas = (sm_limit > 0 && sm_limit < g_num_sms) ? sm_limit : g_num_sms;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (block_n == 256) return dispatch_major<256, 4>(a_major, b_major, ma, mb, p, max_ctas, s);
  return dispatch_major<128, 6>(a_major, b_major, ma, mb, p, max_ctas, s);
}
