// K10 (second generation) — fused attention forward + backward on tcgen05 / TMEM, head_dim 64, bf16.
//
// Same contract and layouts as csrc/ops/attention.cu (Q/K/V read straight out of the fused projection output through
// 3-D TMA maps, O / dqkv written in the projection layouts); what changed is how the non-tensor work is spread:
//
// Forward. For head_dim 64 the kernel is bound by the exponentials (16 MUFU lanes / clk / SM), not by the tensor pipe,
// and generation 1 kept only ONE softmax warp per SM sub-partition per CTA: every dependent FFMA -> MUFU -> F2FP -> STS
// chain was exposed (ncu: XU 41 %, issue 42 %, stall_wait dominant). Here a CTA runs TWO softmax warpgroups with NO
// communication between them: group A owns the even 64-key tiles, group B the odd ones, each with its own running
// maximum, row sum, probability buffer and its own output accumulator in TMEM (split-KV inside the CTA). The two
// partial results are merged once, in the epilogue: O = (O_a 2^(m_a-M) + O_b 2^(m_b-M)) / (l_a 2^(m_a-M) + l_b 2^(m_b-M)).
// With two CTAs per SM that is 16 softmax warps per SM (4 per sub-partition). Optionally a fraction of the
// exponentials is evaluated with a degree-3 polynomial on the FMA pipe (template parameter kPoly of every 4).
//
// Backward. One CTA per 128 keys of a (batch, kv head) as before, but (a) the P / dS computation is split by key
// column between two compute warpgroups (rows need no statistics in backward: lse and delta are inputs), (b) the
// dQ partial tiles leave through ONE bulk tensor reduce-add per 128 x 32 fp32 tile (cp.reduce.async.bulk.tensor:
// the addition happens in L2) instead of 2048 red.global.add.v4 per tile from the SM, and (c) the grid is ordered so
// that the CTAs adding into the same dQ rows run at the same time (their fp32 accumulator tile stays in L2).
#include "../include/mg_attn.cuh"

#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace {

using namespace mg;
using namespace mg::attn;

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int kFBKV = 64;
constexpr int kKVTile = 64 * 128;        // bytes of a [64 keys x 64 bf16] tile
// per softmax group: one K stage (a K tile only lives from its load to the end of one score MMA) and two V stages
struct F2Smem {
  static constexpr int kQ = 0;                                // also the O staging tile in the epilogue
  static constexpr int kK = kQ + kTile;                       // [2 groups]
  static constexpr int kV = kK + 2 * kKVTile;                 // [2 groups][2 stages]
  static constexpr int kStat = kV + 4 * kKVTile;              // [2 groups][128 rows] (m, l)
  static constexpr int kBar = kStat + 2 * 128 * 8;
  static constexpr int kTotal = kBar + 24 * 8 + 16;
};
constexpr int kFwd2SmemBytes = F2Smem::kTotal + 1024;
static_assert(kFwd2SmemBytes <= 113 * 1024, "forward tiles do not fit two CTAs per SM");

template <int kPoly>
__device__ __forceinline__ float exp2_mix(float x, int i) {
  // kPoly of every 4 consecutive elements go to the FMA pipe
  if (kPoly > 0 && (i & 3) < kPoly) return exp2_poly(x);
  return fast_exp2(x);
}

template <int kPoly, int kVar>  // kVar bit 0: driver threads spin instead of parking; bit 1: prefetch the next 16 score columns
__global__ void __launch_bounds__(320, 2)
attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                   const __grid_constant__ CUtensorMap tma_o, const AttnParams p, const int head_major) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F2Smem::kBar);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = bars + 1;              // [group]
  uint64_t* k_done = bars + 3;              // [group]: the score MMA that read the group's K stage retired
  uint64_t* v_full = bars + 5;              // [group][2]
  uint64_t* s_full = bars + 9;              // [group]
  uint64_t* p_ready = bars + 13;            // [group] (4 arrivals: one per warp of the group)
  uint64_t* p_free = bars + 15;             // [group]: P V retired (the S / P columns, O_g and the V stage are free)
  uint64_t* o_full = bars + 17;             // 2 arrivals: one per driver thread
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_q_tiles = (p.T + kBQ - 1) / kBQ;
  const int bh_count = p.B * p.H;
  int qt, bh;
  if (head_major) {
    // all query tiles of one (batch, head) are adjacent in the grid, i.e. co-resident: they read the SAME K / V, which
    // therefore comes out of L2 for all but the first of them. (Query-tile-major order spread the 8 readers of a head
    // over the whole launch: ncu showed a 21 % L2 hit rate and 950 MB of DRAM reads for 300 MB of inputs — the
    // kernel was pinned at ~0.28 ms by K / V re-fetch latency whatever the softmax code looked like.)
    bh = static_cast<int>(blockIdx.x) / num_q_tiles;
    qt = num_q_tiles - 1 - static_cast<int>(blockIdx.x) % num_q_tiles;   // heaviest first inside the head
  } else {
    qt = num_q_tiles - 1 - static_cast<int>(blockIdx.x) / bh_count;
    bh = static_cast<int>(blockIdx.x) % bh_count;
  }
  const int b = bh / p.H, h = bh - b * p.H;
  const int g = h / (p.H / p.KVH);
  const int q0 = qt * kBQ;
  const int kv_end = p.causal ? min(p.T, q0 + kBQ) : p.T;
  const int num_kv = (kv_end + kFBKV - 1) / kFBKV;
  const int col_q = h * kD, col_k = p.Cq + g * kD, col_v = p.Cq + p.Ckv + g * kD;

  if (warp_idx == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tma_q);
      tma_prefetch_desc(&tma_kv);
      tma_prefetch_desc(&tma_o);
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_done[s], 1);
        mbar_init(&v_full[2 * s], 1);
        mbar_init(&v_full[2 * s + 1], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&p_ready[s], 4);
        mbar_init(&p_free[s], 1);
      }
      mbar_init(o_full, 2);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // TMEM columns: [0,64) S of group A, [64,128) S of group B, [128,192) O_a, [192,256) O_b. The bf16 probabilities of a
  // tile overwrite the first 32 columns of its own score tile and feed the P V product as a TENSOR-MEMORY A operand:
  // they never touch shared memory (with 64-wide tiles the Q + K + P + V operand reads plus the P stores of the first
  // version added up to more shared-memory cycles per tile than the exponentials need MUFU cycles).

  if (warp_idx < 2) {
    // ============================ driver threads: one per softmax group ============================
    // Thread X_g (warp g, lane 0) loads the K / V tiles of ITS group's key tiles and issues ITS score and P V MMAs,
    // blocking only on events of its own group. (One producer + one MMA thread walking the tiles in order coupled
    // the groups: when one ran late the other's next score tile or P V product sat behind a wait that had nothing
    // to do with it — ncu showed the softmax warps parked on s_full / p_free for a third of their samples.)
    if (lane == 0) {
      auto drv_wait = [](uint64_t* bar, uint32_t parity) {
        if (kVar & 1) mbar_wait_spin(bar, parity); else mbar_wait_parked(bar, parity);
      };
      const int grp = warp_idx;
      const int n_g = (num_kv - grp + 1) / 2;           // this group's tiles: t_i = grp + 2 i
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, kFmtBF16, 0, 0, 128, kFBKV);
      constexpr uint32_t idesc_o = make_idesc(kFmtBF16, kFmtBF16, 0, 1, 128, kD);
      const uint32_t sq = smem_u32(smem + F2Smem::kQ);
      uint8_t* k_stage = smem + F2Smem::kK + grp * kKVTile;
      uint8_t* v_stage = smem + F2Smem::kV + grp * 2 * kKVTile;
      const uint32_t tmem_s = tmem_base + grp * kFBKV;
      const uint32_t tmem_o = tmem_base + 128 + grp * kD;
      auto load_k = [&](int i) {
        mbar_arrive_expect_tx(&k_full[grp], kKVTile);
        tma_load_3d(k_stage, &tma_kv, &k_full[grp], col_k, (grp + 2 * i) * kFBKV, b);
      };
      auto load_v = [&](int i) {
        mbar_arrive_expect_tx(&v_full[2 * grp + (i & 1)], kKVTile);
        tma_load_3d(v_stage + (i & 1) * kKVTile, &tma_kv, &v_full[2 * grp + (i & 1)], col_v, (grp + 2 * i) * kFBKV, b);
      };
      auto issue_s = [&](int i) {
        drv_wait(&k_full[grp], i & 1);
        tc_fence_after();
        const uint32_t skk = smem_u32(k_stage);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_s, make_smem_desc_sw128(sq + k * 32, 16, 1024), make_smem_desc_sw128(skk + k * 32, 16, 1024),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&k_done[grp]);
        umma_commit(&s_full[grp]);
      };
      if (grp == 0) {
        mbar_arrive_expect_tx(q_full, kTile);
        tma_load_3d(smem + F2Smem::kQ, &tma_q, q_full, col_q, q0, b);
      }
      if (n_g > 0) {
        load_k(0);
        load_v(0);
        if (n_g > 1) load_v(1);
        drv_wait(q_full, 0);
        issue_s(0);
        if (n_g > 1) {
          drv_wait(&k_done[grp], 0);
          load_k(1);
        }
        for (int i = 0; i < n_g; ++i) {
          drv_wait(&p_ready[grp], i & 1);
          tc_fence_after();
          drv_wait(&v_full[2 * grp + (i & 1)], (i >> 1) & 1);
          tc_fence_after();
          const uint32_t svv = smem_u32(v_stage + (i & 1) * kKVTile);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // A = P[128 x 16] at columns [8k, 8k + 8) of the group's score buffer
            umma_f16_ts(tmem_o, tmem_s + k * 8, make_smem_desc_sw128(svv + k * 2048, kKVTile, 1024), idesc_o,
                        (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&p_free[grp]);
          if (i + 1 < n_g) {
            // the next score tile lands on the columns the probabilities just occupied: P V must have retired
            drv_wait(&p_free[grp], i & 1);
            tc_fence_after();
            issue_s(i + 1);
            if (i + 2 < n_g) {
              load_v(i + 2);                              // into the V stage P V(t_i) just released
              drv_wait(&k_done[grp], (i + 1) & 1);
              load_k(i + 2);
            }
          }
        }
        umma_commit(o_full);
      } else {
        mbar_arrive(o_full);
      }
    }
  } else {
    // ============================ softmax groups / epilogue (warps 2..9) ============================
    const int grp = (warp_idx - 2) >> 2;       // 0: even key tiles, 1: odd key tiles
    const int wq = warp_idx & 3;               // TMEM lane quarter this warp may touch
    const int row = wq * 32 + lane;            // query row inside the tile
    const int qi = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int limit = p.causal ? min(qi, p.T - 1) : p.T - 1;  // last visible key index for this row
    const uint32_t tmem_s = tmem_base + grp * kFBKV + lane_off;
    const uint32_t tmem_og = tmem_base + 128 + grp * kD + lane_off;
    float m = -1e30f;  // this group's reference max of the exponents (log2 domain), updated lazily
    float l = 0.f;     // this group's running row sum

    int jj = 0;
    for (int j = grp; j < num_kv; j += 2, ++jj) {
      const int kv0 = j * kFBKV;
      const bool need_mask = (p.causal && kv0 + kFBKV > q0) || (kv0 + kFBKV > p.T);
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t r0[16], r1[16];
      uint32_t pk[32];  // the row's 64 probabilities as bf16 pairs: becomes columns [0,32) of the score buffer
      const int nvis = limit - kv0;  // columns [0, nvis] of this tile are visible to the row
      // online softmax at 16-column granularity against the lazily updated reference `m`: a chunk whose max exceeds
      // m by more than 2^8 moves the reference; this group's accumulator in TMEM (and what the row already packed of the
      // current tile) is rescaled — rare after the first tiles.
      auto chunk = [&](uint32_t (&r)[16], auto cc, auto masked) {
        constexpr int c = decltype(cc)::value;
        if (decltype(masked)::value) {
          // diagonal / tail tiles: hidden scores become -inf once, here; max and exp2 then need no mask code
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c * 16 + i > nvis) r[i] = 0xff800000u;
        }
        float mx4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
        for (int i = 0; i < 16; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
        const float cm = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;
        const bool grow = cm > m + 8.f;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m - cm) : 1.f;
          if (grow) m = cm;
          l *= alpha;
#pragma unroll
          for (int i = 0; i < 4; ++i) rs4[i] *= alpha;
#pragma unroll
          for (int t = 0; t < 8 * c; ++t) {  // the chunks of this tile that are already packed
            const float2 f = unpack_bf16x2(pk[t]);
            pk[t] = pack_bf16x2(f.x * alpha, f.y * alpha);
          }
          if (jj > 0) {
            // O_g is only written by this group's P V products; the previous one retired before this tile's scores
            // were even issued (they share TMEM columns with its probabilities)
            uint32_t t0[16];
#pragma unroll 1
            for (int oc = 0; oc < 4; ++oc) {
              tmem_ld_32x32b_x16(tmem_og + oc * 16, t0);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) t0[i] = __float_as_uint(__uint_as_float(t0[i]) * alpha);
              tmem_st_32x32b_x16(tmem_og + oc * 16, t0);
            }
            tmem_st_wait();
          }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float e0 = exp2_mix<kPoly>(fmaf(__uint_as_float(r[2 * t]), p.scale_log2, -m), 2 * t);
          const float e1 = exp2_mix<kPoly>(fmaf(__uint_as_float(r[2 * t + 1]), p.scale_log2, -m), 2 * t + 1);
          rs4[(2 * t) & 3] += e0;
          rs4[(2 * t + 1) & 3] += e1;
          pk[c * 8 + t] = pack_bf16x2(e0, e1);
        }
      };
      auto tile = [&](auto masked) {
        if (kVar & 2) {
          // the next 16 columns are in flight while the current ones are processed
          tmem_ld_32x32b_x16(tmem_s, r0);
          tmem_ld_wait();
          tmem_ld_32x32b_x16(tmem_s + 16, r1);
          chunk(r0, std::integral_constant<int, 0>{}, masked);
          tmem_ld_wait();
          tmem_ld_32x32b_x16(tmem_s + 32, r0);
          chunk(r1, std::integral_constant<int, 1>{}, masked);
          tmem_ld_wait();
          tmem_ld_32x32b_x16(tmem_s + 48, r1);
          chunk(r0, std::integral_constant<int, 2>{}, masked);
          tmem_ld_wait();
          chunk(r1, std::integral_constant<int, 3>{}, masked);
        } else {
          tmem_ld_32x32b_x16(tmem_s, r0);
          tmem_ld_wait();
          chunk(r0, std::integral_constant<int, 0>{}, masked);
          tmem_ld_32x32b_x16(tmem_s + 16, r0);
          tmem_ld_wait();
          chunk(r0, std::integral_constant<int, 1>{}, masked);
          tmem_ld_32x32b_x16(tmem_s + 32, r0);
          tmem_ld_wait();
          chunk(r0, std::integral_constant<int, 2>{}, masked);
          tmem_ld_32x32b_x16(tmem_s + 48, r0);
          tmem_ld_wait();
          chunk(r0, std::integral_constant<int, 3>{}, masked);
        }
      };
      mbar_wait_parked(&s_full[grp], jj & 1);
      tc_fence_after();
      if (need_mask) tile(std::true_type{}); else tile(std::false_type{});
      l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      // probabilities -> tensor memory, on top of the (now consumed) first half of this row's scores
      tmem_st_32x32b_x32(tmem_s, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[grp]);
    }

    // merge the two groups' partial softmaxes; group A finishes output columns [0,32), group B [32,64)
    float2* stat = reinterpret_cast<float2*>(smem + F2Smem::kStat);
    stat[grp * 128 + row] = make_float2(m, l);
    mbar_wait_parked(o_full, 0);
    tc_fence_after();
    named_bar_sync(1, 256);
    const float2 other = stat[(grp ^ 1) * 128 + row];
    const bool has_b = num_kv > 1;
    const float m_a = grp == 0 ? m : other.x, l_a = grp == 0 ? l : other.y;
    const float m_b = grp == 0 ? other.x : m, l_b = grp == 0 ? other.y : l;
    const float mm = has_b ? fmaxf(m_a, m_b) : m_a;
    const float wa = fast_exp2(m_a - mm), wb = has_b ? fast_exp2(m_b - mm) : 0.f;
    const float lt = l_a * wa + l_b * wb;
    const float inv = 1.f / lt;
    const float ka = wa * inv, kb = wb * inv;
    float o[32];
    {
      uint32_t ra[32];
      tmem_ld_32x32b_x32(tmem_base + 128 + grp * 32 + lane_off, ra);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(ra[i]) * ka;
      if (has_b) {
        tmem_ld_32x32b_x32(tmem_base + 128 + kD + grp * 32 + lane_off, ra);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(__uint_as_float(ra[i]), kb, o[i]);
      }
    }
    if (grp == 0 && qi < p.T) p.lse[(static_cast<long long>(b) * p.H + h) * p.T + qi] = mm + log2f(lt);
    // O tile -> swizzled staging (the Q buffer: every MMA that read it has retired) -> one TMA store
    const uint32_t so = smem_u32(smem + F2Smem::kQ + row * 128);
#pragma unroll
    for (int t = 0; t < 4; ++t) st_shared_v4(so + (((grp * 4 + t) ^ (row & 7)) << 4), pack8(&o[t * 8]));
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (threadIdx.x == 64) {
      tma_store_3d(&tma_o, smem + F2Smem::kQ, col_q, q0, b);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
//   warp 0        TMA producer: K_j, V_j once; (Q_i, dO_i) through a 2-stage ring
//   warp 1        MMA issuer:   S = Q K^T, dP = dO V^T (128 x 128 x 64 each, into TMEM);
//                               dV += P^T dO, dK += dS^T Q, dQ = dS K (P / dS read back from shared memory)
//   warps 2..5    dQ writers:   TMEM -> registers -> swizzled fp32 staging -> cp.reduce.async.bulk.tensor (add, in L2)
//   warps 6..21   compute:      four warpgroups, each owning 32 of the tile's 128 key columns. A thread pulls its
//                               32 S and 32 dP values into registers in one go and releases the TMEM buffers at once,
//                               so the NEXT tile's S / dP products run underneath this tile's exponentials; then
//                               p = exp2(s*c - lse), ds = p * (dp - delta) -> bf16 -> shared memory
constexpr int kQStages = 3;  // (Q_i, dO_i) ring: a stage is only released by the gradient MMAs of its tile, so with
                             // two stages the reload's L2 / HBM latency sat on the critical path of every second tile
struct B2Smem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTile;
  static constexpr int kQ = kV + kTile;                   // kQStages stages
  static constexpr int kdO = kQ + kQStages * kTile;       // kQStages stages
  static constexpr int kP = kdO + kQStages * kTile;       // [128 q x 128 kv] bf16, two 64-wide chunks
  static constexpr int kdS = kP + 2 * kTile;
  static constexpr int kDQ = kdS + 2 * kTile;      // fp32 [2 halves][128 rows x 32 floats], 128B-swizzled
  static constexpr int kBar = kDQ + 2 * kTile;
  static constexpr int kTotal = kBar + 20 * 8 + 16;
};
constexpr int kBwd2SmemBytes = B2Smem::kTotal + 1024;
static_assert(kBwd2SmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");

template <int kPoly>
__global__ void __launch_bounds__(704, 1)
attn_bwd_v2_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                   const __grid_constant__ CUtensorMap tma_dqkv, const __grid_constant__ CUtensorMap tma_dq,
                   const AttnParams p, const int group_major) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B2Smem::kBar);
  uint64_t* kv_full = bars;            // 1
  uint64_t* qdo_full = bars + 1;       // kQStages
  uint64_t* qdo_empty = bars + 5;      // kQStages
  uint64_t* sdp_full = bars + 9;       // 1
  uint64_t* sdp_free = bars + 10;      // 1 (16 arrivals)
  uint64_t* pds_ready = bars + 11;     // 1 (16 arrivals)
  uint64_t* pds_free = bars + 12;      // 1
  uint64_t* dq_full = bars + 13;       // 1
  uint64_t* dq_free = bars + 14;       // 1 (4 arrivals)
  uint64_t* dkv_full = bars + 15;      // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kv_tiles = (p.T + 127) / 128;
  const int num_q_tiles = num_kv_tiles;
  const int bg_count = p.B * p.KVH;
  int jt, bg;
  if (group_major) {
    // all key tiles of one (batch, kv head) are adjacent in the grid: the CTAs adding into the same dQ rows are
    // co-resident, so the fp32 accumulator tile they share stays in L2; heaviest key tile (0) first inside a group
    bg = static_cast<int>(blockIdx.x) / num_kv_tiles;
    jt = static_cast<int>(blockIdx.x) % num_kv_tiles;
  } else {
    jt = static_cast<int>(blockIdx.x) / bg_count;
    bg = static_cast<int>(blockIdx.x) % bg_count;
  }
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
    tma_prefetch_desc(&tma_dq);
    mbar_init(kv_full, 1);
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 16);
    mbar_init(pds_ready, 16);
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
      tma_load_3d(smem + B2Smem::kK, &tma_qkv, kv_full, col_k, kv0, b);
      tma_load_3d(smem + B2Smem::kV, &tma_qkv, kv_full, col_v, kv0, b);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % kQStages;
        const int hq = g * group + it / tiles_per_head;
        const int q0 = (i_first + it % tiles_per_head) * 128;
        mbar_wait_spin(&qdo_empty[st], ((it / kQStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kTile);
        tma_load_3d(smem + B2Smem::kQ + st * kTile, &tma_qkv, &qdo_full[st], hq * kD, q0, b);
        tma_load_3d(smem + B2Smem::kdO + st * kTile, &tma_do, &qdo_full[st], hq * kD, q0, b);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, kFmtBF16, 0, 0, 128, 128);    // S, dP
      constexpr uint32_t idesc_g = make_idesc(kFmtBF16, kFmtBF16, 1, 1, 128, 64);     // dV, dK
      constexpr uint32_t idesc_q = make_idesc(kFmtBF16, kFmtBF16, 0, 1, 128, 64);     // dQ
      const uint32_t sk = smem_u32(smem + B2Smem::kK), sv = smem_u32(smem + B2Smem::kV);
      const uint32_t sp = smem_u32(smem + B2Smem::kP), sds = smem_u32(smem + B2Smem::kdS);
      auto issue_sdp = [&](int it) {
        const int st = it % kQStages;
        mbar_wait_spin(&qdo_full[st], (it / kQStages) & 1);
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + B2Smem::kQ + st * kTile);
        const uint32_t sdo = smem_u32(smem + B2Smem::kdO + st * kTile);
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
      mbar_wait_spin(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % kQStages;
        if (it + 1 < n_iter) {
          mbar_wait_spin(sdp_free, it & 1);  // the compute warps drained S(it) / dP(it) from TMEM
          tc_fence_after();
          issue_sdp(it + 1);
        }
        mbar_wait_spin(pds_ready, it & 1);
        tc_fence_after();
        mbar_wait_spin(dq_free, (it & 1) ^ 1);  // the writers drained dQ(it - 1)
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + B2Smem::kQ + st * kTile);
        const uint32_t sdo = smem_u32(smem + B2Smem::kdO + st * kTile);
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
  } else if (warp_idx >= 6) {
    // ============================ compute warps ============================
    const int qtr = (warp_idx - 6) >> 2;       // which 32 key columns of the 128-wide tile
    const int wq = warp_idx & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t sp_row = smem_u32(smem + B2Smem::kP + (qtr >> 1) * kTile + row * 128);
    const uint32_t sds_row = smem_u32(smem + B2Smem::kdS + (qtr >> 1) * kTile + row * 128);
    const uint32_t ts = tmem_s + lane_off + qtr * 32, td = tmem_dp + lane_off + qtr * 32;
    const int kvq0 = kv0 + qtr * 32;
    // lse / delta of the NEXT tile are fetched while the current one is processed (their L2 latency used to sit
    // between the S / dP products and the first exponential of every tile)
    auto stats_of = [&](int it, float& lse_r, float& delta_r) {
      const int hq = g * group + it / tiles_per_head;
      const int qi = (i_first + it % tiles_per_head) * 128 + row;
      const long long stat = (static_cast<long long>(b) * p.H + hq) * p.T + qi;
      const bool ok = it < n_iter && qi < p.T;
      lse_r = ok ? __ldg(p.lse + stat) : 0.f;
      delta_r = ok ? __ldg(p.delta + stat) : 0.f;
    };
    float lse_n, delta_n;
    stats_of(0, lse_n, delta_n);
    for (int it = 0; it < n_iter; ++it) {
      const int q0 = (i_first + it % tiles_per_head) * 128;
      const int qi = q0 + row;
      const float lse_r = lse_n, delta_r = delta_n;
      stats_of(it + 1, lse_n, delta_n);
      const bool need_mask = (p.causal && kvq0 + 32 > q0) || (kvq0 + 32 > p.T);
      const int limit = p.causal ? min(qi, p.T - 1) : p.T - 1;
      uint32_t rs[32], rd[32];
      mbar_wait_parked(sdp_full, it & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32(ts, rs);
      tmem_ld_32x32b_x32(td, rd);
      tmem_ld_wait();
      // this warp's share of S / dP is in registers: the MMA warp may overwrite them with the next tile's products
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);
      uint4 pk[4], dk[4];
      auto math = [&](auto masked) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float pv[8], dsv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float pe = exp2_mix<kPoly>(fmaf(__uint_as_float(rs[t * 8 + i]), p.scale_log2, -lse_r), i);
            if (decltype(masked)::value && kvq0 + t * 8 + i > limit) pe = 0.f;
            pv[i] = pe;
            dsv[i] = pe * (__uint_as_float(rd[t * 8 + i]) - delta_r);
          }
          pk[t] = pack8(pv);
          dk[t] = pack8(dsv);
        }
      };
      if (need_mask) math(std::true_type{}); else math(std::false_type{});
      // only the stores have to wait until the gradient MMAs of the previous tile stopped reading P / dS
      mbar_wait_parked(pds_free, (it & 1) ^ 1);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t sw = ((((qtr & 1) * 4 + t) ^ (row & 7)) << 4);
        st_shared_v4(sp_row + sw, pk[t]);
        st_shared_v4(sds_row + sw, dk[t]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
    }
    // epilogue: quarters 0 / 1 store dV columns [0,32) / [32,64), quarters 2 / 3 the same of dK (x softmax scale):
    // bf16 -> staging (first 64-column chunk of the P resp. dS buffer) -> TMA stores into dqkv
    mbar_wait_parked(dkv_full, 0);
    tc_fence_after();
    {
      const bool is_dk = qtr >= 2;
      const int ch = qtr & 1;
      const float mul = is_dk ? p.scale_log2 * 0.6931471805599453f : 1.f;
      const uint32_t src = (is_dk ? tmem_dk : tmem_dv) + lane_off + ch * 32;
      const uint32_t dst = smem_u32(smem + (is_dk ? B2Smem::kdS : B2Smem::kP) + row * 128);
      uint32_t r0[32];
      float v[32];
      tmem_ld_32x32b_x32(src, r0);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) * mul;
#pragma unroll
      for (int t = 0; t < 4; ++t) st_shared_v4(dst + (((ch * 4 + t) ^ (row & 7)) << 4), pack8(&v[t * 8]));
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 512);
    if (threadIdx.x == 192) {
      tma_store_3d(&tma_dqkv, smem + B2Smem::kP, col_v, kv0, b);
      tma_store_3d(&tma_dqkv, smem + B2Smem::kdS, col_k, kv0, b);
      tma_store_commit();
      tma_store_wait<0>();
    }
  } else {
    // ============================ dQ writers (warps 2..5) ============================
    const int wq = warp_idx & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t stage_row = smem_u32(smem + B2Smem::kDQ + row * 128);
    const bool issuer = threadIdx.x == 64;
    for (int it = 0; it < n_iter; ++it) {
      const int hq = g * group + it / tiles_per_head;
      const int q0 = (i_first + it % tiles_per_head) * 128;
      mbar_wait_parked(dq_full, it & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tmem_dq + lane_off, r0);
      tmem_ld_32x32b_x32(tmem_dq + lane_off + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      // the previous tile's bulk reduce must have finished READING the staging buffer
      if (issuer) tma_store_wait_read<0>();
      named_bar_sync(2, 128);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const uint32_t sw = ((t ^ (row & 7)) << 4);
        st_shared_v4(stage_row + sw, make_uint4(r0[t * 4], r0[t * 4 + 1], r0[t * 4 + 2], r0[t * 4 + 3]));
        st_shared_v4(stage_row + kTile + sw, make_uint4(r1[t * 4], r1[t * 4 + 1], r1[t * 4 + 2], r1[t * 4 + 3]));
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (issuer) {
        tma_reduce_add_3d(&tma_dq, smem + B2Smem::kDQ, hq * kD, q0, b);
        tma_reduce_add_3d(&tma_dq, smem + B2Smem::kDQ + kTile, hq * kD + 32, q0, b);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename K>
int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? 0 : -300 - static_cast<int>(e);
}

}  // namespace

extern "C" {

// knobs (read once): MAGGY_B200_ATTN_POLY = 1: one of every 4 forward exponentials runs as a polynomial on the FMA pipe,
// MAGGY_B200_ATTN_POLY_BWD likewise for backward, MAGGY_B200_ATTN_BWD_ORDER = 1 group-major grid (default),
// MAGGY_B200_ATTN_FWD_ORDER = 1 head-major forward grid, MAGGY_B200_ATTN_VAR = forward scheduling variant (bit 0 / 1)
int mg_attn_fwd_v2(const void* qkv, void* o, float* lse, int B, int T, int H, int KVH, int causal, float scale,
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
  CUtensorMap mq, mkv, mo;
  int rc = make_map_btc(&mq, qkv, B, T, p.Cq + 2 * p.Ckv, p.Cq + 2 * p.Ckv, kBQ);
  if (rc) return rc;
  rc = make_map_btc(&mkv, qkv, B, T, p.Cq + 2 * p.Ckv, p.Cq + 2 * p.Ckv, kFBKV);
  if (rc) return rc - 500;
  rc = make_map_btc(&mo, o, B, T, p.Cq, p.Cq, kBQ);
  if (rc) return rc - 1000;
  static const int poly = env_int("MAGGY_B200_ATTN_POLY", 0);
  static const int fwd_order = env_int("MAGGY_B200_ATTN_FWD_ORDER", 0);
  static const int var = env_int("MAGGY_B200_ATTN_VAR", 0) & 3;
  const int grid = ((T + kBQ - 1) / kBQ) * B * H;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto launch = [&](auto kernel) -> int {
    static bool configured = false;  // one instance per kernel instantiation
    if (!configured) {
      int r = set_smem(kernel, kFwd2SmemBytes);
      if (r) return r;
      configured = true;
    }
    kernel<<<grid, 320, kFwd2SmemBytes, s>>>(mq, mkv, mo, p, fwd_order);
    return 0;
  };
  if (poly == 1) rc = launch(attn_fwd_v2_kernel<1, 0>);
  else if (var == 1) rc = launch(attn_fwd_v2_kernel<0, 1>);
  else if (var == 2) rc = launch(attn_fwd_v2_kernel<0, 2>);
  else if (var == 3) rc = launch(attn_fwd_v2_kernel<0, 3>);
  else rc = launch(attn_fwd_v2_kernel<0, 0>);
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fprintf(stderr, "maggy_b200 attention fwd v2: launch failed: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

int mg_attn_bwd_v2(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, float* delta,
                   float* dq_acc, int B, int T, int H, int KVH, int causal, float scale, void* stream) {
  if (B <= 0 || T <= 0 || H <= 0 || KVH <= 0 || H % KVH != 0) return -1;
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
  CUtensorMap mqkv, mdo, mdqkv, mdq;
  int rc = make_map_btc(&mqkv, qkv, B, T, W, W, 128);
  if (rc) return rc;
  rc = make_map_btc(&mdo, d_o, B, T, p.Cq, p.Cq, 128);
  if (rc) return rc - 500;
  rc = make_map_btc(&mdqkv, dqkv, B, T, W, W, 128);
  if (rc) return rc - 1000;
  rc = make_map_btc_f32(&mdq, dq_acc, B, T, p.Cq, 128);
  if (rc) return rc - 1500;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, sizeof(float) * static_cast<size_t>(B) * T * p.Cq, s);
  if (e != cudaSuccess) return -500 - static_cast<int>(e);
  launch_delta(o, d_o, delta, B, T, H, s);
  static const int poly = env_int("MAGGY_B200_ATTN_POLY_BWD", 0);
  static const int order = env_int("MAGGY_B200_ATTN_BWD_ORDER", 1);
  static int configured = -1;
  if (configured != poly) {
    rc = poly == 1 ? set_smem(attn_bwd_v2_kernel<1>, kBwd2SmemBytes) : set_smem(attn_bwd_v2_kernel<0>, kBwd2SmemBytes);
    if (rc) return rc;
    configured = poly;
  }
  const int grid = ((T + 127) / 128) * B * KVH;
  if (poly == 1) attn_bwd_v2_kernel<1><<<grid, 704, kBwd2SmemBytes, s>>>(mqkv, mdo, mdqkv, mdq, p, order);
  else attn_bwd_v2_kernel<0><<<grid, 704, kBwd2SmemBytes, s>>>(mqkv, mdo, mdqkv, mdq, p, order);
  launch_dq_convert(dq_acc, dqkv, static_cast<long long>(B) * T, p.Cq, W, scale, s);
  e = cudaGetLastError();
  if (e != cudaSuccess) fprintf(stderr, "maggy_b200 attention bwd v2: launch failed: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // extern "C"
