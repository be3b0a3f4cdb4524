// K5 (block-scaled fp8 variant) — MXFP8 GEMM on tcgen05.mma.kind::mxf8f6f4.block_scale (sm_100a).
//
//   C[M, N] (bf16) = sum over 32-wide K blocks of  2^(ea-127) * 2^(eb-127) * (A_q . B_q^T)  (+ bias[n])
//
// A_q [M, K] and B_q [N, K] are e4m3 bytes (K-major), each 32-element block along K carries ONE UE8M0 scale byte
// (the OCP MX format). The scales never pass through registers: they are staged to shared memory next to the operand
// tiles, copied to TENSOR MEMORY with tcgen05.cp (UTCCP, 32x128b.warpx4) and applied by the tensor core itself; the
// instruction descriptor of each K = 32 MMA names which byte of the 4-byte scale word to use (a_sf_id / b_sf_id).
//
// The reference reaches fp8 nowhere (its stack predates it); BASELINE.json config 5 asks for it ("block-scaled fp8
// where the trial trains in fp8"). Quantisation (bf16 -> e4m3 + UE8M0, csrc below) writes the scales directly in the
// layout UTCCP wants: for every (128-row block, 128-wide K block) one 512-byte chunk with
//   offset(row r, k-block kb) = (r % 32) * 16 + ((r % 128) / 32) * 4 + (kb % 4)
// (CUTLASS calls this the Sm1xx block-scaled K-major atom) so that a chunk is one 1-D bulk copy and one UTCCP.
//
// Kernel structure: persistent, 256 threads = TMA producer warp / MMA warp / TMEM allocator warp / 4 epilogue warps,
// 128 x 128 x 128 tiles, 5-stage operand ring (A 16 KB + B 16 KB + 2 x 512 B of scales per stage), two accumulator
// stages in TMEM (columns [0,256)), scale columns at [256,264).
#include "../include/mg_ptx.cuh"

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <mutex>

namespace {
using namespace mg;

constexpr int BLOCK_M = 128, BLOCK_N = 128, BLOCK_K = 128, UMMA_K = 32, STAGES = 5;
constexpr int kNumThreads = 256, kEpiWarpBase = 4;
constexpr int kABytes = BLOCK_M * BLOCK_K, kBBytes = BLOCK_N * BLOCK_K, kSfBytes = 512;
constexpr int kStageBytes = kABytes + kBBytes + 2 * kSfBytes;   // 33792 = 33 * 1024
constexpr int kOutOffset = STAGES * kStageBytes;
constexpr int kBarOffset = kOutOffset + 4 * 2 * 4096;
constexpr int kSmemTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16;
constexpr uint32_t kTmemCols = 512, kTmemSfa = 256, kTmemSfb = 260;
static_assert(kStageBytes % 1024 == 0, "operand tiles must stay 1024-byte aligned (128B swizzle)");

struct MxParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  const uint8_t* sfa;   // [num_m_blocks][num_k_blocks][512]
  const uint8_t* sfb;   // [num_n_blocks][num_k_blocks][512]
  const __nv_bfloat16* bias;
};

// shared-memory descriptor without swizzle (UTCCP source: 8 x 16-byte core matrices, 128 bytes apart along rows)
MG_DEVICE uint64_t make_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}

// instruction descriptor of the block-scaled kinds (different from kind::f16 / f8f6f4: bits [4,6) and [29,31) select
// the scale bytes, bit 23 the scale format): e4m3 x e4m3, K-major operands, UE8M0 scales, dense
__host__ __device__ constexpr uint32_t make_idesc_mx(uint32_t M, uint32_t N, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (a_sf_id << 29);
}

MG_DEVICE void umma_mxf8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa,
                            uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

// shared memory -> tensor memory: 32 rows x 128 bits, broadcast to the four 32-lane quarters (the scale-factor copy)
MG_DEVICE void utccp_32x128b_warpx4(uint32_t tmem_dst, uint64_t smem_desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}

__global__ void __launch_bounds__(kNumThreads, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, const MxParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile / p.num_m_blocks, m_blk = tile - n_blk * p.num_m_blocks;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(sa + kABytes, &tma_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          bulk_load_1d(sa + kABytes + kBBytes,
                       p.sfa + (static_cast<size_t>(m_blk) * p.num_k_blocks + kb) * kSfBytes, kSfBytes, &full_bar[stage]);
          bulk_load_1d(sa + kABytes + kBBytes + kSfBytes,
                       p.sfb + (static_cast<size_t>(n_blk) * p.num_k_blocks + kb) * kSfBytes, kSfBytes, &full_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          // this k-block's scale words -> tensor memory; tcgen05.cp and the MMAs that follow execute in issue order
          utccp_32x128b_warpx4(tmem_base + kTmemSfa, make_smem_desc_nosw(sb + kBBytes, 0, 128));
          utccp_32x128b_warpx4(tmem_base + kTmemSfb, make_smem_desc_nosw(sb + kBBytes + kSfBytes, 0, 128));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_mxf8_ss(tmem_d, make_smem_desc_sw128(sa + k * UMMA_K, 16, 1024),
                         make_smem_desc_sw128(sb + k * UMMA_K, 16, 1024), make_idesc_mx(BLOCK_M, BLOCK_N, k, k),
                         tmem_base + kTmemSfa, tmem_base + kTmemSfb, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
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
    const int q = warp_idx - kEpiWarpBase;
    uint8_t* out_stage = smem + kOutOffset + q * 8192;
    int obuf = 0, as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile / p.num_m_blocks, m_blk = tile - n_blk * p.num_m_blocks;
      const int row0 = m_blk * BLOCK_M + q * 32, n0 = n_blk * BLOCK_N;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 64; ++c) {
        const int col0 = n0 + c * 64;
        if (col0 >= p.N) break;   // (c == 0 is always inside the matrix; its iteration released the accumulator)
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + c * 64, r0);
        tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
        uint4 bias_v[8];
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            bias_v[j] = (col0 + j * 8 + 8 <= p.N) ? __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + j * 8))
                                                  : make_uint4(0, 0, 0, 0);
        }
        tmem_ld_wait();
        if (c == BLOCK_N / 64 - 1 || col0 + 64 >= p.N) {
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
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float2 a = unpack_bf16x2(bias_v[j].x), b = unpack_bf16x2(bias_v[j].y),
                   cc = unpack_bf16x2(bias_v[j].z), d = unpack_bf16x2(bias_v[j].w);
            v[j * 8 + 0] += a.x; v[j * 8 + 1] += a.y; v[j * 8 + 2] += b.x; v[j * 8 + 3] += b.y;
            v[j * 8 + 4] += cc.x; v[j * 8 + 5] += cc.y; v[j * 8 + 6] += d.x; v[j * 8 + 7] += d.y;
          }
        }
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        uint8_t* buf = out_stage + obuf * 4096 + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(buf + ((j ^ (lane & 7)) << 4)) =
              make_uint4(pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                         pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tma_c, out_stage + obuf * 4096, col0, row0);
          tma_store_commit();
        }
        obuf ^= 1;
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// quantisation: bf16 [rows, K] -> e4m3 bytes [rows, K] + UE8M0 scale bytes in the UTCCP chunk layout
// ---------------------------------------------------------------------------------------------------------------
// Thread t of a row handles 8 consecutive elements; 4 threads share one 32-element scale block.
__global__ void quant_mxfp8_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, uint8_t* __restrict__ q,
                                   long long ldq, uint8_t* __restrict__ sf, int rows, int K, int num_k_blocks) {
  const long long vec_per_row = (K + 7) / 8;
  const long long total = static_cast<long long>(rows) * ((vec_per_row + 3) / 4 * 4);
  const long long vpr4 = (vec_per_row + 3) / 4 * 4;   // whole 32-element blocks per row (threads stay in quads)
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / vpr4);
    const int vc = static_cast<int>(i - static_cast<long long>(r) * vpr4);
    const int c0 = vc * 8;
    float v[8];
    if (c0 + 8 <= K) {
      const uint4 u = ld_nc_v4(x + static_cast<long long>(r) * ldx + c0);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = c0 + t < K ? __bfloat162float(x[static_cast<long long>(r) * ldx + c0 + t]) : 0.f;
    }
    float am = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) am = fmaxf(am, fabsf(v[t]));
    const unsigned quad_mask = __activemask();   // rows are padded to whole quads: a quad is active as a whole
    am = fmaxf(am, __shfl_xor_sync(quad_mask, am, 1));
    am = fmaxf(am, __shfl_xor_sync(quad_mask, am, 2));
    // smallest power of two s with amax / s <= 448 (e4m3 max): e = ceil(log2(amax / 448)), clamped to UE8M0's range
    int e = am > 0.f ? static_cast<int>(ceilf(log2f(am * (1.f / 448.f)))) : -127;
    e = e < -127 ? -127 : (e > 127 ? 127 : e);
    const float inv = exp2f(static_cast<float>(-e));
    if (c0 < K) {
      const uint32_t p0 = __nv_cvt_float2_to_fp8x2(make_float2(v[0] * inv, v[1] * inv), __NV_SATFINITE, __NV_E4M3);
      const uint32_t p1 = __nv_cvt_float2_to_fp8x2(make_float2(v[2] * inv, v[3] * inv), __NV_SATFINITE, __NV_E4M3);
      const uint32_t p2 = __nv_cvt_float2_to_fp8x2(make_float2(v[4] * inv, v[5] * inv), __NV_SATFINITE, __NV_E4M3);
      const uint32_t p3 = __nv_cvt_float2_to_fp8x2(make_float2(v[6] * inv, v[7] * inv), __NV_SATFINITE, __NV_E4M3);
      if (c0 + 8 <= K) {
        *reinterpret_cast<uint2*>(q + static_cast<long long>(r) * ldq + c0) = make_uint2(p0 | (p1 << 16), p2 | (p3 << 16));
      } else {
        const uint32_t w[2] = {p0 | (p1 << 16), p2 | (p3 << 16)};
        for (int t = 0; c0 + t < K; ++t) q[static_cast<long long>(r) * ldq + c0 + t] = (w[t >> 2] >> ((t & 3) * 8)) & 0xFF;
      }
    }
    if ((vc & 3) == 0) {
      const int kb32 = vc >> 2;   // 32-element block index along K
      const size_t chunk = static_cast<size_t>(r >> 7) * num_k_blocks + (kb32 >> 2);
      sf[chunk * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + (kb32 & 3)] = static_cast<uint8_t>(e + 127);
    }
  }
}

// bf16 values of a block-scaled operand (the backward pass of the sharded fp8 linear needs the gathered weight in bf16)
__global__ void dequant_mxfp8_kernel(const uint8_t* __restrict__ q, long long ldq, const uint8_t* __restrict__ sf,
                                     __nv_bfloat16* __restrict__ out, long long ldo, int rows, int K, int num_k_blocks) {
  const long long vec_per_row = K / 8;
  const long long total = static_cast<long long>(rows) * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / vec_per_row);
    const int c0 = static_cast<int>(i - static_cast<long long>(r) * vec_per_row) * 8;
    const int kb32 = c0 >> 5;
    const size_t chunk = static_cast<size_t>(r >> 7) * num_k_blocks + (kb32 >> 2);
    const int e = static_cast<int>(sf[chunk * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + (kb32 & 3)]) - 127;
    const float s = exp2f(static_cast<float>(e));
    const uint2 u = *reinterpret_cast<const uint2*>(q + static_cast<long long>(r) * ldq + c0);
    float v[8];
    const uint32_t w[2] = {u.x, u.y};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>((w[t >> 1] >> ((t & 1) * 16)) & 0xFFFF),
                                                       __NV_E4M3);
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
      v[2 * t] = f.x * s;
      v[2 * t + 1] = f.y * s;
    }
    *reinterpret_cast<uint4*>(out + static_cast<long long>(r) * ldo + c0) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

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

int make_map(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, uint64_t inner, uint64_t outer,
             uint64_t ld, uint32_t bi, uint32_t bo) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * esize};
  cuuint32_t box[2] = {bi, bo};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

}  // namespace

extern "C" {

// bytes of the scale buffer for a [rows, K] operand: ceil(rows / 128) * ceil(K / 128) chunks of 512 bytes
unsigned long long mg_mxfp8_sf_bytes(int rows, int K) {
  return static_cast<unsigned long long>((rows + 127) / 128) * ((K + 127) / 128) * 512ull;
}

// x bf16 [rows, K] (row stride ldx elements) -> q e4m3 [rows, K] (row stride ldq bytes, multiple of 16) and sf
// (mg_mxfp8_sf_bytes bytes; zeroed here so that padding rows / k-blocks carry scale byte 0)
int mg_quant_mxfp8(const void* x, long long ldx, void* q, long long ldq, void* sf, int rows, int K, void* stream) {
  if (rows <= 0 || K <= 0) return 0;
  if (ldq % 16 != 0 || ldx % 8 != 0) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(sf, 0, mg_mxfp8_sf_bytes(rows, K), s);
  if (e != cudaSuccess) return -500 - static_cast<int>(e);
  const long long threads = static_cast<long long>(rows) * (((K + 7) / 8 + 3) / 4 * 4);
  long long blocks = (threads + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  quant_mxfp8_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                             static_cast<uint8_t*>(q), ldq, static_cast<uint8_t*>(sf),
                                                             rows, K, (K + 127) / 128);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// out bf16 [rows, K] (row stride ldo) = q * 2^(sf - 127); K % 8 == 0
int mg_dequant_mxfp8(const void* q, long long ldq, const void* sf, void* out, long long ldo, int rows, int K,
                     void* stream) {
  if (rows <= 0 || K <= 0) return 0;
  if (K % 8 != 0 || ldq % 8 != 0 || ldo % 8 != 0) return -1;
  long long blocks = (static_cast<long long>(rows) * (K / 8) + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  dequant_mxfp8_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(q), ldq, static_cast<const uint8_t*>(sf), static_cast<__nv_bfloat16*>(out), ldo, rows,
      K, (K + 127) / 128);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// C [M, N] bf16 (row stride ldc) = blockscaled(A [M, K] e4m3, sfa) x blockscaled(B [N, K] e4m3, sfb)^T (+ bias)
int mg_gemm_mxfp8(const void* A, const void* sfa, const void* B, const void* sfb, void* C, const void* bias, int M,
                  int N, int K, long long lda, long long ldb, long long ldc, int sm_limit, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((N % 8) != 0 || (lda % 16) != 0 || (ldb % 16) != 0 || (ldc % 8) != 0) return -1;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  MxParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
  p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.sfa = static_cast<const uint8_t*>(sfa);
  p.sfb = static_cast<const uint8_t*>(sfb);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  CUtensorMap ma, mb, mc;
  int rc = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, A, K, M, lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, B, K, N, ldb, BLOCK_K, BLOCK_N);
  if (rc) return rc - 1000;
  rc = make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, N, M, ldc, 64, 32);
  if (rc) return rc - 2000;
  constexpr int smem_bytes = kSmemTotal + 1024;
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  int max_ctas = (sm_limit > 0 && sm_limit < sms) ? sm_limit : sms;
  const int grid = tiles < max_ctas ? tiles : max_ctas;
  gemm_mxfp8_kernel<<<grid, kNumThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(ma, mb, mc, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // extern "C"
