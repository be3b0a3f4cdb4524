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
//                             never holds more than 64 scores; P goes to shared memory as the bf16 A operand of the
//                             second MMA. O accumulates IN TMEM across key tiles (MMA accumulate) relative to a lazily
//                             updated reference max: it is only rescaled in place (tcgen05.ld / .st) when a row's max
//                             grows by more than 2^8, which after the first tiles almost never happens
#include "../include/mg_attn.cuh"

#include <cstdio>
#include <cstdlib>
#include <type_traits>

// tuned generation (attention_v2.cu); this file's kernels stay as the reference implementation (MAGGY_B200_ATTN_GEN=1)
extern "C" int mg_attn_fwd_v2(const void* qkv, void* o, float* lse, int B, int T, int H, int KVH, int causal, float scale,
                              void* stream);
extern "C" int mg_attn_bwd_v2(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                              float* delta, float* dq_acc, int B, int T, int H, int KVH, int causal, float scale,
                              void* stream);

namespace {

using namespace mg;
using namespace mg::attn;

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// Key tiles are 64 wide so that TWO score tiles (2 x 64 TMEM columns) fit next to the output accumulator inside
// the 256 columns a CTA may own when two CTAs share an SM: the MMA warp runs two score tiles ahead of the
// softmax warps, which therefore never wait for the tensor pipe's issue-to-commit latency (the first version,
// one 128-wide score tile, spent 30 % of its softmax time in that wait — profiles/attention_fwd.md).
constexpr int kFBKV = 64;
constexpr int kKVTile = 64 * 128;        // bytes of a [64 keys x 64 bf16] tile
constexpr int kPTile = 128 * 128;        // bytes of a [128 rows x 64 bf16] probability tile
constexpr bool kSumViaMma = true;        // row sums come out of the P V product (a ones column appended to V)
constexpr int kKStages = 4, kVStages = 3;

struct FwdSmem {
  static constexpr int kQ = 0;                                // also the O staging tile in the epilogue
  static constexpr int kK = kQ + kTile;
  static constexpr int kV = kK + kKStages * kKVTile;
  static constexpr int kP = kV + kVStages * kKVTile;          // 2 buffers
  static constexpr int kOnes = kP + 2 * kPTile;               // [64 keys x 128 B]: column 0 is 1.0
  static constexpr int kBar = kOnes + kKVTile;
  static constexpr int kTotal = kBar + 26 * 8 + 16;
};
constexpr int kFwdSmemBytes = 113 * 1024;  // 2 x (dynamic + 1 KB reserved) must fit the SM's 228 KB
static_assert(FwdSmem::kTotal <= kFwdSmemBytes, "forward tiles do not fit two CTAs per SM");

__global__ void __launch_bounds__(192, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                        const __grid_constant__ CUtensorMap tma_o, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  if (smem + FwdSmem::kTotal > smem_raw + kFwdSmemBytes) __trap();  // base was not 1024-aligned enough
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = bars + 1;              // 4
  uint64_t* k_empty = bars + 5;             // 4
  uint64_t* v_full = bars + 9;              // 3
  uint64_t* v_empty = bars + 12;            // 3
  uint64_t* s_full = bars + 15;             // 2
  uint64_t* p_ready = bars + 17;            // 2 (4 arrivals each: one per softmax warp)
  uint64_t* p_free = bars + 19;             // 2
  uint64_t* o_full = bars + 21;             // 1
  uint64_t* s_free = bars + 22;             // 2 (4 arrivals each)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_q_tiles = (p.T + kBQ - 1) / kBQ;
  const int bh_count = p.B * p.H;
  // heaviest (last) query tiles first: with a causal mask tile i visits the most key tiles
  const int qt = num_q_tiles - 1 - static_cast<int>(blockIdx.x) / bh_count;
  const int bh = static_cast<int>(blockIdx.x) % bh_count;
  const int b = bh / p.H, h = bh - b * p.H;
  const int g = h / (p.H / p.KVH);
  const int q0 = qt * kBQ;
  const int kv_end = p.causal ? min(p.T, q0 + kBQ) : p.T;
  const int num_kv = (kv_end + kFBKV - 1) / kFBKV;
  const int col_q = h * kD, col_k = p.Cq + g * kD, col_v = p.Cq + p.Ckv + g * kD;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    tma_prefetch_desc(&tma_o);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < kVStages; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);
      mbar_init(&p_free[s], 1);
      mbar_init(&s_free[s], 4);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  if (warp_idx >= 2 && kSumViaMma) {
    // ones tile, MN-major 128B-swizzled: logical (key k, column 0) = 1.0 sits in 16-byte unit (k & 7) of row k
    const int t = threadIdx.x - 64;  // 0..127: two threads per row, 64 B each
    const int k = t >> 1;
    uint4* rowp = reinterpret_cast<uint4*>(smem + FwdSmem::kOnes + k * 128) + (t & 1) * 4;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int unit = (t & 1) * 4 + u;
      rowp[u] = make_uint4(unit == (k & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_o = tmem_base + 128;    // 64 output columns (+ column 64: row sum) inside a 96-column region

  if (warp_idx == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTile);
      tma_load_3d(smem + FwdSmem::kQ, &tma_q, q_full, col_q, q0, b);
      for (int j = 0; j < num_kv; ++j) {
        const int sk = j % kKStages, sv = j % kVStages;
        mbar_wait_parked(&k_empty[sk], ((j / kKStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[sk], kKVTile);
        tma_load_3d(smem + FwdSmem::kK + sk * kKVTile, &tma_kv, &k_full[sk], col_k, j * kFBKV, b);
        mbar_wait_parked(&v_empty[sv], ((j / kVStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[sv], kKVTile);
        tma_load_3d(smem + FwdSmem::kV + sv * kKVTile, &tma_kv, &v_full[sv], col_v, j * kFBKV, b);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, kFmtBF16, 0, 0, 128, kFBKV);
      constexpr uint32_t idesc_o = make_idesc(kFmtBF16, kFmtBF16, 0, 1, 128, kSumViaMma ? 80 : 64);
      const uint32_t sq = smem_u32(smem + FwdSmem::kQ);
      const uint32_t sones = smem_u32(smem + FwdSmem::kOnes);
      auto issue_s = [&](int j) {
        const int sk = j % kKStages;
        mbar_wait_parked(&k_full[sk], (j / kKStages) & 1);
        tc_fence_after();
        const uint32_t skk = smem_u32(smem + FwdSmem::kK + sk * kKVTile);
        const uint32_t tmem_s = tmem_base + (j & 1) * kFBKV;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_s, make_smem_desc_sw128(sq + k * 32, 16, 1024), make_smem_desc_sw128(skk + k * 32, 16, 1024),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&k_empty[sk]);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait_parked(q_full, 0);
      issue_s(0);
      if (num_kv > 1) issue_s(1);
      for (int j = 0; j < num_kv; ++j) {
        const int pb = j & 1, sv = j % kVStages;
        if (j + 2 < num_kv) {
          // S(j) sits in the softmax warps' registers: its TMEM buffer can take S(j + 2) right away, a whole
          // tile before the probabilities of tile j are ready
          mbar_wait_parked(&s_free[pb], (j >> 1) & 1);
          tc_fence_after();
          issue_s(j + 2);
        }
        mbar_wait_parked(&p_ready[pb], (j >> 1) & 1);
        tc_fence_after();
        mbar_wait_parked(&v_full[sv], (j / kVStages) & 1);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + FwdSmem::kP + pb * kPTile);
        const uint32_t svv = smem_u32(smem + FwdSmem::kV + sv * kKVTile);
        // B = V tile (64 columns) followed, one "chunk stride" later, by the ones tile (16 more columns)
        const uint32_t lbo = kSumViaMma ? sones - svv : static_cast<uint32_t>(kKVTile);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_o, make_smem_desc_sw128(sp + k * 32, 16, 1024),
                      make_smem_desc_sw128(svv + k * 2048, lbo, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&v_empty[sv]);
        umma_commit(&p_free[pb]);
        if (j == num_kv - 1) umma_commit(o_full);
      }
    }
  } else {
    // ============================ softmax / epilogue ============================
    const int wq = warp_idx & 3;               // TMEM lane quarter this warp may touch
    const int row = wq * 32 + lane;            // query row inside the tile
    const int qi = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int limit = p.causal ? min(qi, p.T - 1) : p.T - 1;  // last visible key index for this row
    float m = -1e30f;  // reference max of the exponents (log2 domain), updated lazily
    float l = 0.f;     // running row sum (only when !kSumViaMma)

    for (int j = 0; j < num_kv; ++j) {
      const int pb = j & 1;
      const int kv0 = j * kFBKV;
      const bool need_mask = (p.causal && kv0 + kFBKV > q0) || (kv0 + kFBKV > p.T);
      const uint32_t tmem_s = tmem_base + pb * kFBKV + lane_off;
      const uint32_t sp_row = smem_u32(smem + FwdSmem::kP + pb * kPTile + row * 128);
      uint32_t ra[32], rb[32];
      mbar_wait_parked(&s_full[pb], (j >> 1) & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32(tmem_s, ra);
      tmem_ld_32x32b_x32(tmem_s + 32, rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[pb]);
      mbar_wait_parked(&p_free[pb], ((j >> 1) & 1) ^ 1);  // the P V product that read this P buffer two tiles ago retired
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      // online softmax at 32-column granularity against the lazily updated reference `m`: a chunk whose max
      // exceeds m by more than 2^8 moves the reference; the accumulator in TMEM (and what this row already wrote of
      // the current P tile) is rescaled — rare after the first tiles, so its cost does not matter.
      auto chunk = [&](const uint32_t (&r)[32], int c, auto masked) {
        float mx4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (!decltype(masked)::value || kv0 + c * 32 + i <= limit)
            mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
        }
        const float cm = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;
        const bool grow = cm > m + 8.f;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m - cm) : 1.f;
          if (grow) m = cm;
          l *= alpha;
#pragma unroll
          for (int i = 0; i < 4; ++i) rs4[i] *= alpha;
          for (int t = 0; t < 4 * c; ++t) {  // chunk 0 of this tile, already in shared memory (own row)
            const uint32_t addr = sp_row + ((t ^ (row & 7)) << 4);
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                         : "r"(addr));
            float f[8];
            const float2 a = unpack_bf16x2(u.x), b2 = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z),
                         d = unpack_bf16x2(u.w);
            f[0] = a.x * alpha; f[1] = a.y * alpha; f[2] = b2.x * alpha; f[3] = b2.y * alpha;
            f[4] = c2.x * alpha; f[5] = c2.y * alpha; f[6] = d.x * alpha; f[7] = d.y * alpha;
            st_shared_v4(addr, pack8(f));
          }
          if (j > 0) {
            // the accumulator may only be touched while no P V product is in flight: the previous tile's product
            // must have retired (its p_free commit), the next one waits for this warp's p_ready arrival
            mbar_wait_parked(&p_free[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tc_fence_after();
            uint32_t t0[32];
#pragma unroll 1
            for (int oc = 0; oc < (kSumViaMma ? 3 : 2); ++oc) {
              tmem_ld_32x32b_x32(tmem_o + lane_off + oc * 32, t0);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) t0[i] = __float_as_uint(__uint_as_float(t0[i]) * alpha);
              tmem_st_32x32b_x32(tmem_o + lane_off + oc * 32, t0);
            }
            tmem_st_wait();
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = fast_exp2(fmaf(__uint_as_float(r[t * 8 + i]), p.scale_log2, -m));
            if (decltype(masked)::value && kv0 + c * 32 + t * 8 + i > limit) e[i] = 0.f;
            if (!kSumViaMma) rs4[i & 3] += e[i];
          }
          st_shared_v4(sp_row + (((c * 4 + t) ^ (row & 7)) << 4), pack8(e));
        }
      };
      if (need_mask) {
        chunk(ra, 0, std::true_type{});
        chunk(rb, 1, std::true_type{});
      } else {
        chunk(ra, 0, std::false_type{});
        chunk(rb, 1, std::false_type{});
      }
      if (!kSumViaMma) l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[pb]);
    }
    // all P V products retired: normalise and write out
    mbar_wait_parked(o_full, 0);
    tc_fence_after();
    float o[kD];
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tmem_o + lane_off, r0);
      tmem_ld_32x32b_x32(tmem_o + lane_off + 32, r1);
      tmem_ld_wait();
      if (kSumViaMma) {
        uint32_t r2[32];
        tmem_ld_32x32b_x32(tmem_o + lane_off + 64, r2);
        tmem_ld_wait();
        l = __uint_as_float(r2[0]);
      }
      const float inv = 1.f / l;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = __uint_as_float(r0[i]) * inv;
        o[32 + i] = __uint_as_float(r1[i]) * inv;
      }
    }
    if (qi < p.T) p.lse[(static_cast<long long>(b) * p.H + h) * p.T + qi] = m + log2f(l);
    // O tile -> swizzled staging (the Q buffer: every MMA that read it has retired) -> one TMA store
    const uint32_t so = smem_u32(smem + FwdSmem::kQ + row * 128);
#pragma unroll
    for (int t = 0; t < 8; ++t) st_shared_v4(so + ((t ^ (row & 7)) << 4), pack8(&o[t * 8]));
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
// backward
// ------------------------------------------------------------------------------------------------
// One CTA owns 128 keys of one (batch, kv head): K_j and V_j stay in shared memory, dK_j / dV_j accumulate in
// TMEM over every query tile (and every query head of the group) that sees those keys; dQ partials are added to a
// fp32 buffer with vector reductions.
//
//   warp 0      TMA producer: K_j, V_j once; (Q_i, dO_i) through a 2-stage ring
//   warp 1      MMA issuer:   S = Q K^T, dP = dO V^T                  (128 x 128 x 64 each, into TMEM)
//                             dV += P^T dO, dK += dS^T Q, dQ = dS K    (P / dS read back from shared memory)
//   warps 2..5  compute:      one thread per query row: p = exp2(s*c - lse), ds = p * (dp - delta), both written
//                             to shared memory as bf16 in the one layout that serves as the MN-major A operand
//                             of dV / dK and as the K-major A operand of dQ
//   warps 6..9  dQ writers:   TMEM -> registers -> red.global.add.v4.f32
// The next tile's S / dP are issued as soon as the compute warps have drained the current ones from TMEM, i.e.
// they overlap the tail of the compute phase and the three gradient MMAs.
struct BwdSmem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTile;
  static constexpr int kQ = kV + kTile;            // 2 stages
  static constexpr int kdO = kQ + 2 * kTile;       // 2 stages
  static constexpr int kP = kdO + 2 * kTile;       // [128 q x 128 kv] bf16, two 64-wide chunks
  static constexpr int kdS = kP + 2 * kTile;
  static constexpr int kBar = kdS + 2 * kTile;
  static constexpr int kTotal = kBar + 16 * 8 + 16;
};

__global__ void __launch_bounds__(320, 1)
attn_bwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                        const __grid_constant__ CUtensorMap tma_dqkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBar);
  uint64_t* kv_full = bars;            // 1
  uint64_t* qdo_full = bars + 1;       // 2
  uint64_t* qdo_empty = bars + 3;      // 2
  uint64_t* sdp_full = bars + 5;       // 1
  uint64_t* sdp_free = bars + 6;       // 1 (4 arrivals)
  uint64_t* pds_ready = bars + 7;      // 1 (4 arrivals)
  uint64_t* pds_free = bars + 8;       // 1
  uint64_t* dq_full = bars + 9;        // 1
  uint64_t* dq_free = bars + 10;       // 1 (4 arrivals)
  uint64_t* dkv_full = bars + 11;      // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kv_tiles = (p.T + 127) / 128;
  const int num_q_tiles = num_kv_tiles;
  const int bg_count = p.B * p.KVH;
  // heaviest key tiles first: with a causal mask tile j is seen by num_q_tiles - j query tiles
  const int jt = static_cast<int>(blockIdx.x) / bg_count;
  const int bg = static_cast<int>(blockIdx.x) % bg_count;
  const int b = bg / p.KVH, g = bg - b * p.KVH;
  const int group = p.H / p.KVH;
  const int kv0 = jt * 128;
  const int i_first = p.causal ? jt : 0;
  const int tiles_per_head = num_q_tiles - i_first;
  const int n_iter = group * tiles_per_head;  // it -> (head g*group + it / tiles_per_head, q tile i_first + it % ...)
  const int col_k = p.Cq + g * kD, col_v = p.Cq + p.Ckv + g * kD;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_dqkv);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 4);
    mbar_init(pds_ready, 4);
    mbar_init(pds_free, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_init(dkv_full, 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dv = tmem_base + 256,
                 tmem_dk = tmem_base + 320, tmem_dq = tmem_base + 384;

  if (warp_idx == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * kTile);
      tma_load_3d(smem + BwdSmem::kK, &tma_qkv, kv_full, col_k, kv0, b);
      tma_load_3d(smem + BwdSmem::kV, &tma_qkv, kv_full, col_v, kv0, b);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it & 1;
        const int hq = g * group + it / tiles_per_head;
        const int q0 = (i_first + it % tiles_per_head) * 128;
        mbar_wait_parked(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kTile);
        tma_load_3d(smem + BwdSmem::kQ + st * kTile, &tma_qkv, &qdo_full[st], hq * kD, q0, b);
        tma_load_3d(smem + BwdSmem::kdO + st * kTile, &tma_do, &qdo_full[st], hq * kD, q0, b);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, kFmtBF16, 0, 0, 128, 128);    // S, dP
      constexpr uint32_t idesc_g = make_idesc(kFmtBF16, kFmtBF16, 1, 1, 128, 64);     // dV, dK
      constexpr uint32_t idesc_q = make_idesc(kFmtBF16, kFmtBF16, 0, 1, 128, 64);     // dQ
      const uint32_t sk = smem_u32(smem + BwdSmem::kK), sv = smem_u32(smem + BwdSmem::kV);
      const uint32_t sp = smem_u32(smem + BwdSmem::kP), sds = smem_u32(smem + BwdSmem::kdS);
      auto issue_sdp = [&](int it) {
        const int st = it & 1;
        mbar_wait_parked(&qdo_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + BwdSmem::kQ + st * kTile);
        const uint32_t sdo = smem_u32(smem + BwdSmem::kdO + st * kTile);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_s, make_smem_desc_sw128(sq + k * 32, 16, 1024), make_smem_desc_sw128(sk + k * 32, 16, 1024),
                      idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_dp, make_smem_desc_sw128(sdo + k * 32, 16, 1024),
                      make_smem_desc_sw128(sv + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(sdp_full);
      };
      mbar_wait_parked(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it & 1;
        if (it + 1 < n_iter) {
          mbar_wait_parked(sdp_free, it & 1);  // the compute warps drained S(it) / dP(it) from TMEM
          tc_fence_after();
          issue_sdp(it + 1);
        }
        mbar_wait_parked(pds_ready, it & 1);
        tc_fence_after();
        mbar_wait_parked(dq_free, (it & 1) ^ 1);  // the writers drained dQ(it - 1)
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + BwdSmem::kQ + st * kTile);
        const uint32_t sdo = smem_u32(smem + BwdSmem::kdO + st * kTile);
        const uint32_t acc0 = it > 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dV[kv, d] += sum_q P[q, kv] dO[q, d]
          umma_f16_ss(tmem_dv, make_smem_desc_sw128(sp + k * 2048, kTile, 1024),
                      make_smem_desc_sw128(sdo + k * 2048, kTile, 1024), idesc_g, (k > 0) ? 1u : acc0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dK[kv, d] += sum_q dS[q, kv] Q[q, d]
          umma_f16_ss(tmem_dk, make_smem_desc_sw128(sds + k * 2048, kTile, 1024),
                      make_smem_desc_sw128(sq + k * 2048, kTile, 1024), idesc_g, (k > 0) ? 1u : acc0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dQ[q, d] = sum_kv dS[q, kv] K[kv, d]
          umma_f16_ss(tmem_dq, make_smem_desc_sw128(sds + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                      make_smem_desc_sw128(sk + k * 2048, kTile, 1024), idesc_q, k > 0 ? 1u : 0u);
        }
        umma_commit(pds_free);
        umma_commit(&qdo_empty[st]);
        umma_commit(dq_full);
      }
      umma_commit(dkv_full);
    }
  } else if (warp_idx < 6) {
    // ============================ compute warps ============================
    const int wq = warp_idx & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t sp_row = smem_u32(smem + BwdSmem::kP + row * 128);
    const uint32_t sds_row = smem_u32(smem + BwdSmem::kdS + row * 128);
    for (int it = 0; it < n_iter; ++it) {
      const int hq = g * group + it / tiles_per_head;
      const int q0 = (i_first + it % tiles_per_head) * 128;
      const int qi = q0 + row;
      const long long stat = (static_cast<long long>(b) * p.H + hq) * p.T + qi;
      const float lse_r = qi < p.T ? __ldg(p.lse + stat) : 0.f;
      const float delta_r = qi < p.T ? __ldg(p.delta + stat) : 0.f;
      const bool need_mask = (p.causal && kv0 + 128 > q0) || (kv0 + 128 > p.T);
      const int limit = p.causal ? min(qi, p.T - 1) : p.T - 1;
      mbar_wait_parked(sdp_full, it & 1);
      tc_fence_after();
      uint32_t rs[32], rd[32];
      tmem_ld_32x32b_x32(tmem_s + lane_off, rs);
      tmem_ld_32x32b_x32(tmem_dp + lane_off, rd);
      mbar_wait_parked(pds_free, (it & 1) ^ 1);  // the gradient MMAs of the previous tile no longer read P / dS
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        float pv[32], dsv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float pe = fast_exp2(fmaf(__uint_as_float(rs[i]), p.scale_log2, -lse_r));
          if (need_mask && kv0 + c * 32 + i > limit) pe = 0.f;
          pv[i] = pe;
          dsv[i] = pe * (__uint_as_float(rd[i]) - delta_r);
        }
        if (c < 3) {
          tmem_ld_32x32b_x32(tmem_s + lane_off + (c + 1) * 32, rs);
          tmem_ld_32x32b_x32(tmem_dp + lane_off + (c + 1) * 32, rd);
        } else {
          // S / dP fully drained: the MMA warp may overwrite them with the next tile's products
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_free);
        }
        const uint32_t off = (c >> 1) * kTile;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t sw = ((((c & 1) * 4 + t) ^ (row & 7)) << 4);
          st_shared_v4(sp_row + off + sw, pack8(&pv[t * 8]));
          st_shared_v4(sds_row + off + sw, pack8(&dsv[t * 8]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
    }
    // epilogue: dV, dK (x softmax scale) -> bf16 -> staging (P / dS buffers) -> TMA stores into dqkv
    mbar_wait_parked(dkv_full, 0);
    tc_fence_after();
    const float sc = p.scale_log2 * 0.6931471805599453f;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint32_t r0[32], r1[32];
      const uint32_t src = (which == 0 ? tmem_dv : tmem_dk) + lane_off;
      tmem_ld_32x32b_x32(src, r0);
      tmem_ld_32x32b_x32(src + 32, r1);
      tmem_ld_wait();
      float v[64];
      const float mul = which == 0 ? 1.f : sc;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(r0[i]) * mul;
        v[32 + i] = __uint_as_float(r1[i]) * mul;
      }
      const uint32_t dst = (which == 0 ? sp_row : sds_row);
#pragma unroll
      for (int t = 0; t < 8; ++t) st_shared_v4(dst + ((t ^ (row & 7)) << 4), pack8(&v[t * 8]));
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 64) {
      tma_store_3d(&tma_dqkv, smem + BwdSmem::kP, col_v, kv0, b);
      tma_store_3d(&tma_dqkv, smem + BwdSmem::kdS, col_k, kv0, b);
      tma_store_commit();
      tma_store_wait<0>();
    }
  } else {
    // ============================ dQ writers ============================
    const int wq = warp_idx & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    for (int it = 0; it < n_iter; ++it) {
      const int hq = g * group + it / tiles_per_head;
      const int qi = (i_first + it % tiles_per_head) * 128 + row;
      mbar_wait_parked(dq_full, it & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tmem_dq + lane_off, r0);
      tmem_ld_32x32b_x32(tmem_dq + lane_off + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      if (qi < p.T) {
        float* dst = p.dq_acc + (static_cast<long long>(b) * p.T + qi) * p.Cq + hq * kD;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + t * 4),
                       "f"(__uint_as_float(r0[t * 4])), "f"(__uint_as_float(r0[t * 4 + 1])),
                       "f"(__uint_as_float(r0[t * 4 + 2])), "f"(__uint_as_float(r0[t * 4 + 3]))
                       : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 32 + t * 4),
                       "f"(__uint_as_float(r1[t * 4])), "f"(__uint_as_float(r1[t * 4 + 1])),
                       "f"(__uint_as_float(r1[t * 4 + 2])), "f"(__uint_as_float(r1[t * 4 + 3]))
                       : "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b, h, t] = sum_d dO[b, t, h*64 + d] * O[b, t, h*64 + d]; one 8-lane group per (b, t, h)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  float* __restrict__ delta, int B, int T, int H) {
  const long long total = static_cast<long long>(B) * T * H;
  long long gidx = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool valid = gidx < total;  // keep every lane alive for the shuffles
  if (!valid) gidx = total - 1;
  const int h = static_cast<int>(gidx % H);
  const long long bt = gidx / H;
  const long long base = (bt * H + h) * kD + sub * 8;
  const uint4 a = ld_nc_v4(o + base), c = ld_nc_v4(d_o + base);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 c0 = unpack_bf16x2(c.x), c1 = unpack_bf16x2(c.y), c2 = unpack_bf16x2(c.z), c3 = unpack_bf16x2(c.w);
  float acc = a0.x * c0.x + a0.y * c0.y + a1.x * c1.x + a1.y * c1.y + a2.x * c2.x + a2.y * c2.y + a3.x * c3.x +
              a3.y * c3.y;
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0 && valid) {
    const int t = static_cast<int>(bt % T);
    const int bb = static_cast<int>(bt / T);
    delta[(static_cast<long long>(bb) * H + h) * T + t] = acc;
  }
}

// dqkv[b, t, 0:Cq] = bf16(scale * dq_acc[b, t, :])
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, long long rows,
                                       int Cq, int ld, float scale) {
  const int vec_per_row = Cq / 8;
  const long long total = rows * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row) * 8;
    const float4 x = *reinterpret_cast<const float4*>(acc + r * Cq + c);
    const float4 y = *reinterpret_cast<const float4*>(acc + r * Cq + c + 4);
    const float v[8] = {x.x * scale, x.y * scale, x.z * scale, x.w * scale,
                        y.x * scale, y.y * scale, y.z * scale, y.w * scale};
    *reinterpret_cast<uint4*>(dqkv + r * ld + c) = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
}  // namespace

namespace mg {
namespace attn {
void launch_delta(const void* o, const void* d_o, float* delta, int B, int T, int H, cudaStream_t s) {
  const long long groups = static_cast<long long>(B) * T * H;
  const long long blocks = (groups * 8 + 255) / 256;
  attn_delta_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(o),
                                                                  static_cast<const __nv_bfloat16*>(d_o), delta, B, T, H);
}
void launch_dq_convert(const float* dq_acc, void* dqkv, long long rows, int Cq, int ld, float scale, cudaStream_t s) {
  long long blocks = (rows * (Cq / 8) + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  attn_dq_convert_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(dq_acc, static_cast<__nv_bfloat16*>(dqkv), rows, Cq,
                                                                       ld, scale);
}
int generation() {
  static int gen = [] {
    const char* e = getenv("MAGGY_B200_ATTN_GEN");
    return e ? atoi(e) : 2;
  }();
  return gen;
}
}  // namespace attn
}  // namespace mg

extern "C" {

// O[B,T,H*64] = softmax(scale * Q K^T (+ causal mask)) V with Q/K/V read from qkv[B,T,H*64 + 2*KVH*64].
// lse: fp32 [B,H,T] (log2 domain), needed by the backward pass.
int mg_attn_fwd(const void* qkv, void* o, float* lse, int B, int T, int H, int KVH, int causal, float scale,
                void* stream) {
  if (B <= 0 || T <= 0 || H <= 0 || KVH <= 0 || H % KVH != 0) return -1;
  if (mg::attn::generation() >= 2) return mg_attn_fwd_v2(qkv, o, lse, B, T, H, KVH, causal, scale, stream);
  AttnParams p;
  p.B = B; p.T = T; p.H = H; p.KVH = KVH;
  p.Cq = H * kD; p.Ckv = KVH * kD;
  p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.delta = nullptr;
  p.dq_acc = nullptr;
  CUtensorMap mq, mkv, mo;
  int rc = make_map_btc(&mq, qkv, B, T, p.Cq + 2 * p.Ckv, p.Cq + 2 * p.Ckv, kBQ);
  if (rc) return rc;
  rc = make_map_btc(&mkv, qkv, B, T, p.Cq + 2 * p.Ckv, p.Cq + 2 * p.Ckv, kFBKV);
  if (rc) return rc - 500;
  rc = make_map_btc(&mo, o, B, T, p.Cq, p.Cq, kBQ);
  if (rc) return rc - 1000;
  constexpr int smem_bytes = kFwdSmemBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  const int grid = ((T + kBQ - 1) / kBQ) * B * H;
  attn_fwd_tcgen05_kernel<<<grid, 192, smem_bytes, static_cast<cudaStream_t>(stream)>>>(mq, mkv, mo, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fprintf(stderr, "maggy_b200 attention fwd: launch failed: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// dqkv[B,T,Cq+2Ckv] from dO[B,T,Cq], O, lse (forward outputs) and qkv. Workspaces: delta fp32 [B,H,T],
// dq_acc fp32 [B,T,Cq] (zeroed here).
int mg_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta,
                float* dq_acc, int B, int T, int H, int KVH, int causal, float scale, void* stream) {
  if (B <= 0 || T <= 0 || H <= 0 || KVH <= 0 || H % KVH != 0) return -1;
  if (mg::attn::generation() >= 2)
    return mg_attn_bwd_v2(qkv, o, d_o, lse, dqkv, delta, dq_acc, B, T, H, KVH, causal, scale, stream);
  AttnParams p;
  p.B = B; p.T = T; p.H = H; p.KVH = KVH;
  p.Cq = H * kD; p.Ckv = KVH * kD;
  p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dq_acc = dq_acc;
  const int W = p.Cq + 2 * p.Ckv;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap mqkv, mdo, mdqkv;
  int rc = make_map_btc(&mqkv, qkv, B, T, W, W, 128);
  if (rc) return rc;
  rc = make_map_btc(&mdo, d_o, B, T, p.Cq, p.Cq, 128);
  if (rc) return rc - 500;
  rc = make_map_btc(&mdqkv, dqkv, B, T, W, W, 128);
  if (rc) return rc - 1000;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, sizeof(float) * static_cast<size_t>(B) * T * p.Cq, s);
  if (e != cudaSuccess) return -500 - static_cast<int>(e);
  {
    const long long groups = static_cast<long long>(B) * T * H;
    const long long blocks = (groups * 8 + 255) / 256;
    attn_delta_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(o),
                                                                    static_cast<const __nv_bfloat16*>(d_o), delta, B, T,
                                                                    H);
  }
  constexpr int smem_bytes = BwdSmem::kTotal + 1024;
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(attn_bwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = true;
  }
  const int grid = ((T + 127) / 128) * B * KVH;
  attn_bwd_tcgen05_kernel<<<grid, 320, smem_bytes, s>>>(mqkv, mdo, mdqkv, p);
  {
    const long long rows = static_cast<long long>(B) * T;
    long long blocks = (rows * (p.Cq / 8) + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    attn_dq_convert_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(dq_acc, static_cast<__nv_bfloat16*>(dqkv), rows,
                                                                         p.Cq, W, scale);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) fprintf(stderr, "maggy_b200 attention bwd: launch failed: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // extern "C"
