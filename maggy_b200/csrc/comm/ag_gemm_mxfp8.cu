// K3 (block-scaled fp8 variant) — parameter all-gather fused into the MXFP8 tcgen05 GEMM (sm_100a, NVLink 5).
//
//   y[M, N] (bf16) = x_mx[M, K] . W_mx^T (+ bias),  W = [N, K] row-sharded over `world` ranks as e4m3 bytes + UE8M0 scales
//
// BASELINE.json config 5 ("Llama-3 1B fp8, param all-gather + GEMM fused"; reference mechanism: fairscale FSDP's NCCL
// all-gather followed by the cuBLAS GEMM, maggy/core/patching/modules.py:75-95). Same structure as csrc/comm/ag_gemm.cu:
// a few gather CTAs stream the remote shards out of the owners' HBM over NVLink with bulk async copies — here HALF the
// bytes of the bf16 version (one byte per weight + 1/32 byte of scale) — while the other CTAs already multiply against
// the local shard; tiles are visited local-shard-first and the producer warp acquires `ready[n_blk]` before it issues
// TMA loads on a gathered block. The compute pipeline is the one of csrc/ops/gemm_mxfp8.cu (scales staged to tensor
// memory with tcgen05.cp, tcgen05.mma.kind::mxf8f6f4.block_scale); this file is generated from it by hand-editing the
// tile schedule, so fixes to the MMA / epilogue code belong in both.
#include "../include/mg_ptx.cuh"

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

constexpr int kMaxWorld = 8;
constexpr int kGatherChunk = 32768, kGatherStages = 5;  // ring of 32 KB bulk copies inside the (unused) operand stages
static_assert(kGatherStages * kGatherChunk <= kOutOffset, "gather ring must fit in the operand stage area");

struct MxParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  const uint8_t* sfa;   // [num_m_blocks][num_k_blocks][512]
  const __nv_bfloat16* bias;
  // gather: every rank owns N / world rows of the weight as e4m3 bytes + their scale chunks (peer-mapped)
  const uint8_t* q_peer[kMaxWorld];     // [N / world, K] bytes
  const uint8_t* sf_peer[kMaxWorld];    // [(N / world) / 128][num_k_blocks][512]
  uint8_t* gathered_q;                  // local [N, K] (remote rows only are written)
  uint8_t* gathered_sf;                 // local [N / 128][num_k_blocks][512]
  uint32_t* ready;                      // [num_n_blocks], zeroed before launch
  uint32_t* flag_peer[kMaxWorld];       // symmetric words: [0..8) shard-ready epochs, [8..16) done-reading
  uint32_t* local_sync;                 // [0] finished-CTA counter, [1] launches completed (epoch source)
  int rank, world, n_copy;
  int blocks_per_shard;                 // (N / world) / BLOCK_N
  uint32_t epoch;
};

MG_DEVICE uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
MG_DEVICE void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class Pred>
MG_DEVICE void spin_until(Pred ready, const char* what) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!ready()) {
    if ((++spins & 0x3FF) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ull) {
        printf("maggy_b200 ag_gemm_mxfp8: timed out waiting for %s (block %d)\n", what, static_cast<int>(blockIdx.x));
        __trap();
      }
    }
  }
}

// One thread drives a ring of 32 KB bulk copies: peer HBM -> shared memory (mbarrier complete_tx) -> local HBM.
MG_DEVICE void ring_copy(uint8_t* smem, uint64_t* gbar, uint32_t (&gphase)[kGatherStages], const uint8_t* sp, uint8_t* dp,
                         long long bytes) {
  const int nchunks = static_cast<int>((bytes + kGatherChunk - 1) / kGatherChunk);
  auto chunk_bytes = [&](int c) {
    long long left = bytes - static_cast<long long>(c) * kGatherChunk;
    return static_cast<uint32_t>(left < kGatherChunk ? left : kGatherChunk);
  };
  const int prefetch = nchunks < kGatherStages ? nchunks : kGatherStages;
  for (int c = 0; c < prefetch; ++c) {
    mbar_arrive_expect_tx(&gbar[c], chunk_bytes(c));
    bulk_load_1d(smem + c * kGatherChunk, sp + static_cast<long long>(c) * kGatherChunk, chunk_bytes(c), &gbar[c]);
  }
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % kGatherStages;
    mbar_wait(&gbar[s], gphase[s]);
    gphase[s] ^= 1;
    bulk_store_1d(dp + static_cast<long long>(c) * kGatherChunk, smem + s * kGatherChunk, chunk_bytes(c));
    tma_store_commit();
    if (c + kGatherStages < nchunks) {
      tma_store_wait_read<0>();
      const int cn = c + kGatherStages;
      mbar_arrive_expect_tx(&gbar[s], chunk_bytes(cn));
      bulk_load_1d(smem + s * kGatherChunk, sp + static_cast<long long>(cn) * kGatherChunk, chunk_bytes(cn), &gbar[s]);
    }
  }
  tma_store_wait<0>();
  // the ring is reused by the next call: every stage's barrier has completed an equal number of phases per use only
  // if the stages beyond nchunks are left untouched, which the phase array tracks per stage
}

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
ag_gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b_local,
                     const __grid_constant__ CUtensorMap tma_b_gathered, const __grid_constant__ CUtensorMap tma_c,
                     const MxParams p_in) {
  MxParams p = p_in;
  if (p.epoch == 0) p.epoch = *reinterpret_cast<volatile uint32_t*>(p.local_sync + 1) + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first_local_blk = p.rank * p.blocks_per_shard;

  if (static_cast<int>(blockIdx.x) < p.n_copy) {
    // =========================== gather CTAs: peer shards (bytes + scale chunks) -> local buffers ==================
    if (blockIdx.x == 0 && threadIdx.x < p.world) {
      __threadfence_system();
      st_release_sys_u32(p.flag_peer[threadIdx.x] + p.rank, p.epoch);   // my shard may be read for this epoch
    }
    const int remote_blocks = p.num_n_blocks - p.blocks_per_shard;
    const long long q_bytes = static_cast<long long>(BLOCK_N) * p.K;
    const long long sf_bytes = static_cast<long long>(p.num_k_blocks) * kSfBytes;
    uint64_t* gbar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
    uint32_t gphase[kGatherStages];
#pragma unroll
    for (int s = 0; s < kGatherStages; ++s) gphase[s] = 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < kGatherStages; ++s) mbar_init(&gbar[s], 1);
      fence_barrier_init();
    }
    __syncthreads();
    for (int i = blockIdx.x; i < remote_blocks; i += p.n_copy) {
      int n_blk = first_local_blk + p.blocks_per_shard + i;   // nearest shard first
      if (n_blk >= p.num_n_blocks) n_blk -= p.num_n_blocks;
      const int src = n_blk / p.blocks_per_shard;
      const int lb = n_blk - src * p.blocks_per_shard;
      if (threadIdx.x == 0) {
        spin_until([&] { return ld_acquire_sys_u32(p.flag_peer[p.rank] + src) >= p.epoch; }, "peer shard");
        ring_copy(smem, gbar, gphase, p.q_peer[src] + lb * q_bytes, p.gathered_q + n_blk * q_bytes, q_bytes);
        ring_copy(smem, gbar, gphase, p.sf_peer[src] + lb * sf_bytes, p.gathered_sf + n_blk * sf_bytes, sf_bytes);
        __threadfence();
        st_release_gpu(p.ready + n_blk, 1u);
      }
    }
    __syncthreads();
    __shared__ uint32_t is_last;
    if (threadIdx.x == 0) {
      __threadfence_system();
      is_last = (atomicAdd(p.local_sync, 1u) == static_cast<uint32_t>(p.n_copy) - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
      if (threadIdx.x == 0) {
        p.local_sync[0] = 0;
        p.local_sync[1] = p.epoch;
      }
      if (threadIdx.x < p.world) {
        const int peer = threadIdx.x;
        st_release_sys_u32(p.flag_peer[peer] + kMaxWorld + p.rank, p.epoch);
        spin_until([&] { return ld_acquire_sys_u32(p.flag_peer[p.rank] + kMaxWorld + peer) >= p.epoch; },
                   "peers to finish reading");
      }
    }
    return;
  }
  const int cta = blockIdx.x - p.n_copy, num_ctas = gridDim.x - p.n_copy;
  // tile -> (m_blk, n_blk): n-major, rotated so the local shard's column blocks come first
  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int n_idx = tile / p.num_m_blocks;
    m_blk = tile - n_idx * p.num_m_blocks;
    n_blk = first_local_blk + n_idx;
    if (n_blk >= p.num_n_blocks) n_blk -= p.num_n_blocks;
  };
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b_local);
    tma_prefetch_desc(&tma_b_gathered);
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
      int ready_blk = -1;
      for (int tile = cta; tile < num_tiles; tile += num_ctas) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const bool local = n_blk >= first_local_blk && n_blk < first_local_blk + p.blocks_per_shard;
        if (!local && ready_blk != n_blk) {
          spin_until([&] { return ld_acquire_gpu(p.ready + n_blk) != 0u; }, "gathered weight block");
          asm volatile("fence.proxy.async;" ::: "memory");
          ready_blk = n_blk;
        }
        const CUtensorMap* mb = local ? &tma_b_local : &tma_b_gathered;
        const int nb0 = local ? (n_blk - first_local_blk) * BLOCK_N : n_blk * BLOCK_N;
        const uint8_t* sfb = local ? p.sf_peer[p.rank] + static_cast<size_t>(n_blk - first_local_blk) * p.num_k_blocks * kSfBytes
                                   : p.gathered_sf + static_cast<size_t>(n_blk) * p.num_k_blocks * kSfBytes;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(sa + kABytes, mb, &full_bar[stage], kb * BLOCK_K, nb0);
          bulk_load_1d(sa + kABytes + kBBytes,
                       p.sfa + (static_cast<size_t>(m_blk) * p.num_k_blocks + kb) * kSfBytes, kSfBytes, &full_bar[stage]);
          bulk_load_1d(sa + kABytes + kBBytes + kSfBytes, sfb + static_cast<size_t>(kb) * kSfBytes, kSfBytes,
                       &full_bar[stage]);
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
      for (int tile = cta; tile < num_tiles; tile += num_ctas) {
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
    for (int tile = cta; tile < num_tiles; tile += num_ctas) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
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

// x_q [M, K] e4m3 (row stride ldx bytes) with scales sfa; y [M, N] bf16 (row stride ldy). q_ptrs / sf_ptrs: `world`
// peer-mapped pointers to the [N / world, K] byte shards and their scale chunks; gathered_q [N, K] and gathered_sf
// local scratch; ready: N / 128 uint32 (zeroed here); flag_ptrs: `world` symmetric buffers of >= 16 zeroed uint32;
// local_sync: 2 zeroed uint32. Requirements: (N / world) % 128 == 0, K % 16 == 0.
extern "C" int mg_ag_gemm_mxfp8(const void* x_q, const void* sfa, void* y, const void* bias,
                                const unsigned long long* q_ptrs, const unsigned long long* sf_ptrs, void* gathered_q,
                                void* gathered_sf, void* ready, const unsigned long long* flag_ptrs, void* local_sync,
                                int M, int N, int K, long long ldx, long long ldy, int rank, int world, int n_copy,
                                unsigned int epoch, void* stream) {
  if (world < 1 || world > kMaxWorld || N % world != 0 || (N / world) % BLOCK_N != 0 || K % 16 != 0) return -1;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  MxParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = N / BLOCK_N;
  p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.sfa = static_cast<const uint8_t*>(sfa);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  for (int r = 0; r < kMaxWorld; ++r) {
    p.q_peer[r] = r < world ? reinterpret_cast<const uint8_t*>(q_ptrs[r]) : nullptr;
    p.sf_peer[r] = r < world ? reinterpret_cast<const uint8_t*>(sf_ptrs[r]) : nullptr;
    p.flag_peer[r] = r < world ? reinterpret_cast<uint32_t*>(flag_ptrs[r]) : nullptr;
  }
  p.gathered_q = static_cast<uint8_t*>(gathered_q);
  p.gathered_sf = static_cast<uint8_t*>(gathered_sf);
  p.ready = static_cast<uint32_t*>(ready);
  p.local_sync = static_cast<uint32_t*>(local_sync);
  p.rank = rank; p.world = world;
  p.blocks_per_shard = (N / world) / BLOCK_N;
  p.epoch = epoch;
  if (n_copy <= 0) n_copy = world > 1 ? 8 : 1;
  p.n_copy = n_copy;
  CUtensorMap ma, mbl, mbg, mc;
  int rc = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, x_q, K, M, ldx, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&mbl, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, reinterpret_cast<const void*>(q_ptrs[rank]), K, N / world, K,
                BLOCK_K, BLOCK_N);
  if (rc) return rc - 1000;
  rc = make_map(&mbg, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, gathered_q, K, N, K, BLOCK_K, BLOCK_N);
  if (rc) return rc - 2000;
  rc = make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, N, M, ldy, 64, 32);
  if (rc) return rc - 3000;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  constexpr int smem_bytes = kSmemTotal + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ag_gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  cudaMemsetAsync(ready, 0, sizeof(uint32_t) * p.num_n_blocks, s);
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int compute = tiles < sms - n_copy ? tiles : sms - n_copy;
  ag_gemm_mxfp8_kernel<<<n_copy + compute, kNumThreads, smem_bytes, s>>>(ma, mbl, mbg, mc, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
