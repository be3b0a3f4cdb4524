// K10 — fused attention (flash-style, head_dim 64, bf16) on tcgen05 / TMEM, forward and backward.
//
// Replaces the cuDNN / SDPA attention the reference reaches through the user's oblivious train_fn
// (reference call site: user code under maggy/core/executors/torch_dist_executor.py:145-159).
//
// Layout: the kernels read Q, K and V straight out of the fused projection output `qkv` [B, T, Cq + 2*Ckv]
// (q head h at columns [h*64, h*64+64), k head g at Cq + g*64, v head g at Cq + Ckv + g*64; grouped-query attention
// when Ckv < Cq) through 3-D TMA maps {feature, token, batch}, and write O as [B, T, Cq] — the layout the output
// projection GEMM consumes. No transposes, no q/k/v splits, no concatenation in backward.
//
// Forward (one CTA = 128 query rows of one (batch, head); two CTAs co-reside per SM so one CTA's softmax overlaps
// the other's MMAs):
//   warp 0      TMA producer: Q once, K / V tiles through a 2-stage ring
//   warp 1      MMA issuer:   S = Q K^T (128x128x64) into TMEM; O_blk = P V (128x64x128) into TMEM
//   warps 2..5  softmax:      one thread per query row; S is read from TMEM twice (row max, then exp2) so a thread
//                             never holds more than 32 scores; P goes to shared memory as the bf16 A operand of the
//                             second MMA; the running output lives in registers (O = O*alpha + O_blk)
#include "../include/mg_ptx.cuh"

#include <cstdio>
#include <mutex>

namespace {

using namespace mg;

constexpr int kD = 64;        // head dim
constexpr int kBQ = 128;      // query rows per tile
constexpr int kBKV = 128;     // keys per tile
constexpr int kTile = 128 * 128;  // bytes of a [128 rows x 64 bf16] tile

struct AttnParams {
  int B, T, H, KVH;     // batch, sequence, query heads, key/value heads
  int Cq, Ckv;          // H*64, KVH*64
  int causal;
  float scale_log2;     // softmax scale * log2(e)
  float* lse;           // [B, H, T] log2-domain log-sum-exp of the scaled scores
  // backward only
  const float* delta;   // [B, H, T] rowsum(dO * O)
  float* dq_acc;        // [B, T, Cq] fp32
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                    pack_bf16x2(v[6], v[7]));
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct FwdSmem {
  static constexpr int kQ = 0;                    // also the O staging tile in the epilogue
  static constexpr int kK = kQ + kTile;           // 2 stages
  static constexpr int kV = kK + 2 * kTile;       // 2 stages
  static constexpr int kP = kV + 2 * kTile;       // [128 x 128] bf16 as two 64-wide K chunks
  static constexpr int kBar = kP + 2 * kTile;
  static constexpr int kTotal = kBar + 16 * 8 + 16;
};

__global__ void __launch_bounds__(192, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_o,
                        const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  if (smem + FwdSmem::kTotal > smem_raw + 113 * 1024) __trap();  // base was not 1024-aligned enough
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* v_full = bars + 3;        // 2
  uint64_t* kv_empty = bars + 5;      // 2
  uint64_t* s_full = bars + 7;        // 1
  uint64_t* p_ready = bars + 8;       // 1 (4 arrivals: one per softmax warp)
  uint64_t* o_full = bars + 9;        // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_q_tiles = (p.T + kBQ - 1) / kBQ;
  const int bh_count = p.B * p.H;
  // heaviest (last) query tiles first: with a causal mask tile i visits i + 1 key tiles
  const int qt = num_q_tiles - 1 - static_cast<int>(blockIdx.x) / bh_count;
  const int bh = static_cast<int>(blockIdx.x) % bh_count;
  const int b = bh / p.H, h = bh - b * p.H;
  const int g = h / (p.H / p.KVH);
  const int q0 = qt * kBQ;
  const int kv_end = p.causal ? min(p.T, q0 + kBQ) : p.T;
  const int num_kv = (kv_end + kBKV - 1) / kBKV;
  const int col_q = h * kD, col_k = p.Cq + g * kD, col_v = p.Cq + p.Ckv + g * kD;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_o);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base;          // 128 columns
  const uint32_t tmem_o = tmem_base + 128;    // 64 columns

  if (warp_idx == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTile);
      tma_load_3d(smem + FwdSmem::kQ, &tma_qkv, q_full, col_q, q0, b);
      for (int j = 0; j < num_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kTile);
        tma_load_3d(smem + FwdSmem::kK + s * kTile, &tma_qkv, &k_full[s], col_k, j * kBKV, b);
        mbar_arrive_expect_tx(&v_full[s], kTile);
        tma_load_3d(smem + FwdSmem::kV + s * kTile, &tma_qkv, &v_full[s], col_v, j * kBKV, b);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, kFmtBF16, 0, 0, 128, 128);
      constexpr uint32_t idesc_o = make_idesc(kFmtBF16, kFmtBF16, 0, 1, 128, 64);
      const uint32_t sq = smem_u32(smem + FwdSmem::kQ);
      const uint32_t sp = smem_u32(smem + FwdSmem::kP);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + FwdSmem::kK + s * kTile);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_s, make_smem_desc_sw128(sq + k * 32, 16, 1024), make_smem_desc_sw128(sk + k * 32, 16, 1024),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < num_kv; ++j) {
        const int s = j & 1;
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        mbar_wait(&v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sv = smem_u32(smem + FwdSmem::kV + s * kTile);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ss(tmem_o, make_smem_desc_sw128(sp + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                      make_smem_desc_sw128(sv + k * 2048, kTile, 1024), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(&kv_empty[s]);
        if (j == num_kv - 1) umma_commit(o_full);
        else issue_s(j + 1);
      }
    }
  } else {
    // ============================ softmax / epilogue ============================
    const int wq = warp_idx & 3;               // TMEM lane quarter this warp may touch
    const int row = wq * 32 + lane;            // query row inside the tile
    const int qi = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    uint8_t* sp_row = smem + FwdSmem::kP + row * 128;
    float m = -1e30f, l = 0.f, alpha_prev = 1.f;
    float o[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o[i] = 0.f;

    for (int j = 0; j < num_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (j > 0) {
        // O_blk of the previous key tile (its MMA retired before this tile's S did): fold it in
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tmem_o + lane_off, r0);
        tmem_ld_32x32b_x32(tmem_o + lane_off + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[i] = o[i] * alpha_prev + __uint_as_float(r0[i]);
          o[32 + i] = o[32 + i] * alpha_prev + __uint_as_float(r1[i]);
        }
      }
      const int kv0 = j * kBKV;
      const bool need_mask = (p.causal && kv0 + kBKV > q0) || (kv0 + kBKV > p.T);
      const int limit = p.causal ? min(qi, p.T - 1) : p.T - 1;  // last visible key index for this row
      // Both passes keep one TMEM load in flight while the previous 32 columns are processed.
      uint32_t ra[32], rb[32];
      // pass 1: row maximum
      float mx = -1e30f;
      auto max_chunk = [&](const uint32_t (&r)[32], int c) {
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kv0 + c * 32 + i <= limit) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      };
      tmem_ld_32x32b_x32(tmem_s + lane_off, ra);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem_s + lane_off + 32, rb);
      max_chunk(ra, 0);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem_s + lane_off + 64, ra);
      max_chunk(rb, 1);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem_s + lane_off + 96, rb);
      max_chunk(ra, 2);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem_s + lane_off, ra);  // first chunk of pass 2
      max_chunk(rb, 3);
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float alpha = fast_exp2(m - m_new);
      // pass 2: probabilities -> bf16 A operand of the P V product
      float rs = 0.f;
      auto exp_chunk = [&](const uint32_t (&r)[32], int c) {
        uint8_t* dst = sp_row + (c >> 1) * kTile;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = fast_exp2(fmaf(__uint_as_float(r[t * 8 + i]), p.scale_log2, -m_new));
            if (need_mask && kv0 + c * 32 + t * 8 + i > limit) e[i] = 0.f;
            rs += e[i];
          }
          *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + t) ^ (row & 7)) << 4)) = pack8(e);
        }
      };
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem_s + lane_off + 32, rb);
      exp_chunk(ra, 0);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem_s + lane_off + 64, ra);
      exp_chunk(rb, 1);
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tmem_s + lane_off + 96, rb);
      exp_chunk(ra, 2);
      tmem_ld_wait();
      exp_chunk(rb, 3);
      l = l * alpha + rs;
      m = m_new;
      alpha_prev = alpha;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // last key tile's O_blk
    mbar_wait(o_full, 0);
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tmem_o + lane_off, r0);
      tmem_ld_32x32b_x32(tmem_o + lane_off + 32, r1);
      tmem_ld_wait();
      const float inv = 1.f / l;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = (o[i] * alpha_prev + __uint_as_float(r0[i])) * inv;
        o[32 + i] = (o[32 + i] * alpha_prev + __uint_as_float(r1[i])) * inv;
      }
    }
    if (qi < p.T) p.lse[(static_cast<long long>(b) * p.H + h) * p.T + qi] = m + log2f(l);
    // O tile -> swizzled staging (the Q buffer: every MMA that read it has retired) -> one TMA store
    uint8_t* so = smem + FwdSmem::kQ + row * 128;
#pragma unroll
    for (int t = 0; t < 8; ++t) *reinterpret_cast<uint4*>(so + ((t ^ (row & 7)) << 4)) = pack8(&o[t * 8]);
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 64) {
      tma_store_3d(&tma_o, smem + FwdSmem::kQ, col_q, q0, b);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
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

// [B, T, C] bf16 (row stride `ld` elements) as a 3-D map {C, T, B}; box {64, 128, 1}; 128B swizzle
int make_map_btc(CUtensorMap* map, const void* ptr, int B, int T, int C, long long ld) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * T};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

}  // namespace

extern "C" {

// O[B,T,H*64] = softmax(scale * Q K^T (+ causal mask)) V with Q/K/V read from qkv[B,T,H*64 + 2*KVH*64].
// lse: fp32 [B,H,T] (log2 domain), needed by the backward pass.
int mg_attn_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int KVH, int causal, float scale,
                void* stream) {
  if (B <= 0 || T <= 0 || H <= 0 || KVH <= 0 || H % KVH != 0) return -1;
  AttnParams p;
  p.B = B; p.T = T; p.H = H; p.KVH = KVH;
  p.Cq = H * kD; p.Ckv = KVH * kD;
  p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.delta = nullptr;
  p.dq_acc = nullptr;
  CUtensorMap mqkv, mo;
  int rc = make_map_btc(&mqkv, qkv, B, T, p.Cq + 2 * p.Ckv, p.Cq + 2 * p.Ckv);
  if (rc) return rc;
  rc = make_map_btc(&mo, o, B, T, p.Cq, p.Cq);
  if (rc) return rc - 1000;
  // two CTAs per SM: 2 x (dynamic + 1 KB reserved) must fit the SM's 228 KB; the base is 1024-aligned in practice,
  // the kernel traps if rounding it up would overrun the allocation
  constexpr int smem_bytes = 113 * 1024;
  static_assert(FwdSmem::kTotal <= smem_bytes, "forward tiles do not fit two CTAs per SM");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  const int grid = ((T + kBQ - 1) / kBQ) * B * H;
  attn_fwd_tcgen05_kernel<<<grid, 192, smem_bytes, static_cast<cudaStream_t>(stream)>>>(mqkv, mo, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fprintf(stderr, "maggy_b200 attention fwd: launch failed: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // extern "C"
