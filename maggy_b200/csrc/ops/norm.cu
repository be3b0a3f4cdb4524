// K6 — fused LayerNorm / RMSNorm forward + backward for sm_100a (bf16 I/O, fp32 math).
//
// Bandwidth-bound: one warp owns one row, the row stays in registers between the statistics pass
// and the normalise pass (one HBM read, one HBM write), 16-byte vector accesses, no shared memory in
// forward. Backward keeps per-lane dgamma/dbeta partials in registers across a persistent loop over
// rows and writes one partial row per CTA; a second tiny kernel folds the partials.
// Stands in for the cuDNN/ATen norm kernels the reference reaches through train_fn.
#include "../include/mg_ptx.cuh"

#include <cstdlib>

namespace {
using namespace mg;

constexpr int kMaxVec = 16;  // up to 16 * 32 lanes * 8 = 4096 columns held in registers

template <bool RMS, int NVEC>
__global__ void __launch_bounds__(128)
norm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int C,
                float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const int nvec = C >> 3;  // uint4 vectors per row
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * C);
  uint4 buf[NVEC];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    int v = lane + i * 32;
    if (v < nvec) {
      buf[i] = ld_nc_v4(xr + v);
      float2 a = unpack_bf16x2(buf[i].x), b = unpack_bf16x2(buf[i].y), c = unpack_bf16x2(buf[i].z),
             d = unpack_bf16x2(buf[i].w);
      s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
      ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x +
            d.y * d.y;
    }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float inv_c = 1.0f / static_cast<float>(C);
  float mean = RMS ? 0.f : s * inv_c;
  float var = RMS ? ss * inv_c : fmaxf(ss * inv_c - mean * mean, 0.f);
  float rstd = rsqrtf(var + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * C);
  const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
  const uint4* b4 = reinterpret_cast<const uint4*>(beta);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    int v = lane + i * 32;
    if (v < nvec) {
      uint4 g = __ldg(g4 + v);
      uint4 b = (RMS || beta == nullptr) ? make_uint4(0, 0, 0, 0) : __ldg(b4 + v);
      const uint32_t xi[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
      const uint32_t gi[4] = {g.x, g.y, g.z, g.w};
      const uint32_t bi[4] = {b.x, b.y, b.z, b.w};
      uint32_t o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 xv = unpack_bf16x2(xi[t]), gv = unpack_bf16x2(gi[t]), bv = unpack_bf16x2(bi[t]);
        o[t] = pack_bf16x2((xv.x - mean) * rstd * gv.x + bv.x, (xv.y - mean) * rstd * gv.y + bv.y);
      }
      st_na_v4(yr + v, make_uint4(o[0], o[1], o[2], o[3]));
    }
  }
}

// dx = rstd * (g*dy - mean_c(g*dy) - xhat * mean_c(g*dy*xhat))      (LayerNorm)
// dx = rstd * (g*dy - xhat * mean_c(g*dy*xhat))                     (RMSNorm)
// partial[blockIdx.x][0][c] = sum_rows dy*xhat ; partial[blockIdx.x][1][c] = sum_rows dy
template <bool RMS, int NVEC>
__global__ void __launch_bounds__(128)
norm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean_in,
                const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx,
                float* __restrict__ partial, int rows, int C,
                const __nv_bfloat16* __restrict__ res_grad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  float dg[NVEC][8], db[NVEC][8];
#pragma unroll
  for (int i = 0; i < NVEC; ++i)
#pragma unroll
    for (int t = 0; t < 8; ++t) dg[i][t] = db[i][t] = 0.f;
  const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
  const float inv_c = 1.0f / static_cast<float>(C);

  for (int row = blockIdx.x * 4 + warp; row < rows; row += gridDim.x * 4) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * C);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * C);
    uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * C);
    // gradient that reaches x through the residual stream (x also feeds "x + f(norm(x))"): added here so that
    // autograd does not launch a separate elementwise add over the whole activation
    const uint4* rgr = res_grad == nullptr
                           ? nullptr
                           : reinterpret_cast<const uint4*>(res_grad + static_cast<size_t>(row) * C);
    const float mean = RMS ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    float xh[NVEC][8], gy[NVEC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      int v = lane + i * 32;
      if (v < nvec) {
        uint4 xv = ld_nc_v4(xr + v), dv = ld_nc_v4(dyr + v), gv = __ldg(g4 + v);
        const uint32_t xi[4] = {xv.x, xv.y, xv.z, xv.w}, di[4] = {dv.x, dv.y, dv.z, dv.w},
                       gi[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 xf = unpack_bf16x2(xi[t]), df = unpack_bf16x2(di[t]), gf = unpack_bf16x2(gi[t]);
          float h0 = (xf.x - mean) * rstd, h1 = (xf.y - mean) * rstd;
          xh[i][2 * t] = h0; xh[i][2 * t + 1] = h1;
          gy[i][2 * t] = gf.x * df.x; gy[i][2 * t + 1] = gf.y * df.y;
          dg[i][2 * t] += df.x * h0; dg[i][2 * t + 1] += df.y * h1;
          db[i][2 * t] += df.x; db[i][2 * t + 1] += df.y;
          s1 += gy[i][2 * t] + gy[i][2 * t + 1];
          s2 += gy[i][2 * t] * h0 + gy[i][2 * t + 1] * h1;
        }
      }
    }
    s1 = RMS ? 0.f : warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      int v = lane + i * 32;
      if (v < nvec) {
        uint32_t o[4];
        uint32_t rg[4] = {0u, 0u, 0u, 0u};
        if (rgr != nullptr) {
          const uint4 rv = ld_nc_v4(rgr + v);
          rg[0] = rv.x; rg[1] = rv.y; rg[2] = rv.z; rg[3] = rv.w;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 r2 = unpack_bf16x2(rg[t]);
          float a = rstd * (gy[i][2 * t] - s1 - xh[i][2 * t] * s2) + r2.x;
          float b = rstd * (gy[i][2 * t + 1] - s1 - xh[i][2 * t + 1] * s2) + r2.y;
          o[t] = pack_bf16x2(a, b);
        }
        st_na_v4(dxr + v, make_uint4(o[0], o[1], o[2], o[3]));
      }
    }
  }

  // fold the 4 warps of this CTA through shared memory, then one partial row per CTA
  extern __shared__ float red[];  // [4][2][C]
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        red[(warp * 2 + 0) * C + v * 8 + t] = dg[i][t];
        red[(warp * 2 + 1) * C + v * 8 + t] = db[i][t];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    int which = c / C, col = c - which * C;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) acc += red[(w * 2 + which) * C + col];
    partial[(static_cast<size_t>(blockIdx.x) * 2 + which) * C + col] = acc;
  }
}

// Variant 2 (MAGGY_B200_NORM_BWD=2; the path for C > 1024; =1: the first kernel above; the register-capped 4 CTAs/SM
// instantiation of round 1 spilled 940 B/thread, was slower and is gone): the row is kept as the raw bf16 vectors it was loaded as (half the registers of
// the fp32 copies above) and the NEXT row's loads are issued before the current row is reduced and written, so every
// warp always has one row of loads in flight. The first variant spent two thirds of each row in the shuffle
// reductions / second pass with nothing outstanding (45 % of the copy roofline, profiles/bw_kernels_v1.log).
template <bool RMS, int NVEC, int MINB>
__global__ void __launch_bounds__(128, MINB)
norm_bwd_kernel_v2(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean_in,
                   const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx,
                   float* __restrict__ partial, int rows, int C,
                   const __nv_bfloat16* __restrict__ res_grad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  float dg[NVEC][8], db[NVEC][8];
  uint4 gv[NVEC];
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t) dg[i][t] = db[i][t] = 0.f;
    const int v = lane + i * 32;
    gv[i] = v < nvec ? __ldg(reinterpret_cast<const uint4*>(gamma) + v) : make_uint4(0, 0, 0, 0);
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  const int stride = gridDim.x * 4;
  int row = blockIdx.x * 4 + warp;
  uint4 xc[NVEC], dc[NVEC], xn[NVEC], dn[NVEC];
  auto load_row = [&](int r, uint4 (&xv)[NVEC], uint4 (&dv)[NVEC]) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * C);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(r) * C);
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        xv[i] = ld_nc_v4(xr + v);
        dv[i] = ld_nc_v4(dyr + v);
      } else {
        xv[i] = dv[i] = make_uint4(0, 0, 0, 0);
      }
    }
  };
  if (row < rows) load_row(row, xc, dc);
  for (; row < rows; row += stride) {
    const int next = row + stride;
    if (next < rows) load_row(next, xn, dn);  // in flight while this row is reduced and written
    const float mean = RMS ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const uint32_t xi[4] = {xc[i].x, xc[i].y, xc[i].z, xc[i].w}, di[4] = {dc[i].x, dc[i].y, dc[i].z, dc[i].w},
                     gi[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 xf = unpack_bf16x2(xi[t]), df = unpack_bf16x2(di[t]), gf = unpack_bf16x2(gi[t]);
        const float h0 = (xf.x - mean) * rstd, h1 = (xf.y - mean) * rstd;
        const float g0 = gf.x * df.x, g1 = gf.y * df.y;
        dg[i][2 * t] += df.x * h0; dg[i][2 * t + 1] += df.y * h1;
        db[i][2 * t] += df.x; db[i][2 * t + 1] += df.y;
        s1 += g0 + g1;
        s2 += g0 * h0 + g1 * h1;
      }
    }
    // padding lanes (v >= nvec) hold zeros in x: their h = -mean * rstd is harmless because dy = gamma = 0 there
    s1 = RMS ? 0.f : warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
    uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * C);
    const uint4* rgr = res_grad == nullptr
                           ? nullptr
                           : reinterpret_cast<const uint4*>(res_grad + static_cast<size_t>(row) * C);
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint32_t xi[4] = {xc[i].x, xc[i].y, xc[i].z, xc[i].w}, di[4] = {dc[i].x, dc[i].y, dc[i].z, dc[i].w},
                       gi[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
        uint32_t rg[4] = {0u, 0u, 0u, 0u};
        if (rgr != nullptr) {
          const uint4 rv = ld_nc_v4(rgr + v);
          rg[0] = rv.x; rg[1] = rv.y; rg[2] = rv.z; rg[3] = rv.w;
        }
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 xf = unpack_bf16x2(xi[t]), df = unpack_bf16x2(di[t]), gf = unpack_bf16x2(gi[t]),
                       r2 = unpack_bf16x2(rg[t]);
          const float h0 = (xf.x - mean) * rstd, h1 = (xf.y - mean) * rstd;
          const float a = rstd * (gf.x * df.x - s1 - h0 * s2) + r2.x;
          const float b = rstd * (gf.y * df.y - s1 - h1 * s2) + r2.y;
          o[t] = pack_bf16x2(a, b);
        }
        st_na_v4(dxr + v, make_uint4(o[0], o[1], o[2], o[3]));
      }
    }
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      xc[i] = xn[i];
      dc[i] = dn[i];
    }
  }

  extern __shared__ float red[];  // [4][2][C]
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        red[(warp * 2 + 0) * C + v * 8 + t] = dg[i][t];
        red[(warp * 2 + 1) * C + v * 8 + t] = db[i][t];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    int which = c / C, col = c - which * C;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) acc += red[(w * 2 + which) * C + col];
    partial[(static_cast<size_t>(blockIdx.x) * 2 + which) * C + col] = acc;
  }
}

// Variant 4 (default for C <= 1024): rows are STAGED IN SHARED MEMORY BY THE BULK-COPY ENGINE instead of in registers.
// Each warp owns a ring of kSlots row slots ([x | dy | residual-grad], 16-byte aligned) and lane 0 refills a slot with
// `cp.async.bulk` as soon as the warp has finished with it, so the bytes in flight per SM are set by the shared-memory
// budget (~150-200 KB) and not by the register file; the registers go to the column-sum accumulators. With kCol the
// kernel also accumulates the column sums of the dx it writes: dx of the norm that follows a `linear(..., residual)`
// IS that linear's output gradient, so its bias gradient needs no separate pass over the activation
// (the `colsum_partial/finalize` pair the producer would otherwise launch). One CTA per SM, partial[blockIdx][kAcc][C].
template <bool RMS, int NVEC, bool RES, bool COL>
__global__ void __launch_bounds__(384, 1)
norm_bwd_kernel_v4(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean_in,
                   const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx,
                   float* __restrict__ partial, int rows, int C, const __nv_bfloat16* __restrict__ res_grad,
                   int slots) {
  constexpr int kAcc = COL ? 3 : 2;
  constexpr int kArr = RES ? 3 : 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nvec = C >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(C) * 2u;
  const uint32_t slot_bytes = row_bytes * kArr;
  // [nwarps * slots mbarriers][pad to 128][nwarps][slots][slot_bytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned char* ring = smem_raw + ((static_cast<size_t>(nwarps) * slots * 8 + 127) & ~static_cast<size_t>(127));
  unsigned char* my_ring = ring + static_cast<size_t>(warp) * slots * slot_bytes;
  uint64_t* my_bars = bars + warp * slots;
  if (lane == 0) {
    for (int s = 0; s < slots; ++s) mbar_init(&my_bars[s], 1);
    fence_barrier_init();
  }
  __syncwarp();

  float dg[NVEC][8], db[NVEC][8], dc[COL ? NVEC : 1][8];
  uint4 gv[NVEC];
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      dg[i][t] = db[i][t] = 0.f;
      if (COL) dc[i][t] = 0.f;
    }
    const int v = lane + i * 32;
    gv[i] = v < nvec ? __ldg(reinterpret_cast<const uint4*>(gamma) + v) : make_uint4(0, 0, 0, 0);
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  const int stride = gridDim.x * nwarps;
  int row = blockIdx.x * nwarps + warp;

  auto issue = [&](int r, int s) {  // lane 0 only
    unsigned char* dst = my_ring + static_cast<size_t>(s) * slot_bytes;
    mbar_arrive_expect_tx(&my_bars[s], slot_bytes);
    bulk_load_1d(dst, x + static_cast<size_t>(r) * C, row_bytes, &my_bars[s]);
    bulk_load_1d(dst + row_bytes, dy + static_cast<size_t>(r) * C, row_bytes, &my_bars[s]);
    if (RES) bulk_load_1d(dst + 2 * row_bytes, res_grad + static_cast<size_t>(r) * C, row_bytes, &my_bars[s]);
  };
  if (lane == 0) {
    for (int s = 0; s < slots; ++s) {
      const int r = row + s * stride;
      if (r < rows) issue(r, s);
    }
  }
  float mean_n = 0.f, rstd_n = 0.f;
  if (row < rows) {
    mean_n = RMS ? 0.f : mean_in[row];
    rstd_n = rstd_in[row];
  }
  int slot = 0;
  uint32_t parity = 0;
  for (; row < rows; row += stride) {
    const float mean = mean_n, rstd = rstd_n;
    const int nrow = row + stride;
    if (nrow < rows) {  // the next row's statistics travel while this row is processed
      mean_n = RMS ? 0.f : mean_in[nrow];
      rstd_n = rstd_in[nrow];
    }
    mbar_wait(&my_bars[slot], parity);
    const uint32_t base = smem_u32(my_ring + static_cast<size_t>(slot) * slot_bytes);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint4 xv = ld_shared_v4(base + v * 16), dv = ld_shared_v4(base + row_bytes + v * 16);
        const uint32_t xi[4] = {xv.x, xv.y, xv.z, xv.w}, di[4] = {dv.x, dv.y, dv.z, dv.w},
                       gi[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 xf = unpack_bf16x2(xi[t]), df = unpack_bf16x2(di[t]), gf = unpack_bf16x2(gi[t]);
          const float h0 = (xf.x - mean) * rstd, h1 = (xf.y - mean) * rstd;
          const float g0 = gf.x * df.x, g1 = gf.y * df.y;
          dg[i][2 * t] += df.x * h0; dg[i][2 * t + 1] += df.y * h1;
          db[i][2 * t] += df.x; db[i][2 * t + 1] += df.y;
          s1 += g0 + g1;
          s2 += g0 * h0 + g1 * h1;
        }
      }
    }
    s1 = RMS ? 0.f : warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
    uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * C);
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint4 xv = ld_shared_v4(base + v * 16), dv = ld_shared_v4(base + row_bytes + v * 16);
        const uint4 rv = RES ? ld_shared_v4(base + 2 * row_bytes + v * 16) : make_uint4(0, 0, 0, 0);
        const uint32_t xi[4] = {xv.x, xv.y, xv.z, xv.w}, di[4] = {dv.x, dv.y, dv.z, dv.w},
                       gi[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w}, rg[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 xf = unpack_bf16x2(xi[t]), df = unpack_bf16x2(di[t]), gf = unpack_bf16x2(gi[t]),
                       r2 = unpack_bf16x2(rg[t]);
          const float h0 = (xf.x - mean) * rstd, h1 = (xf.y - mean) * rstd;
          const float a = rstd * (gf.x * df.x - s1 - h0 * s2) + r2.x;
          const float b = rstd * (gf.y * df.y - s1 - h1 * s2) + r2.y;
          o[t] = pack_bf16x2(a, b);
          if (COL) {  // sum what the consumer GEMMs will read: the bf16-rounded values
            const float2 q = unpack_bf16x2(o[t]);
            dc[i][2 * t] += q.x;
            dc[i][2 * t + 1] += q.y;
          }
        }
        st_na_v4(dxr + v, make_uint4(o[0], o[1], o[2], o[3]));
      }
    }
    __syncwarp();  // every lane has read the slot: lane 0 may hand it back to the copy engine
    if (lane == 0) {
      const int r = row + slots * stride;
      if (r < rows) issue(r, slot);
    }
    if (++slot == slots) {
      slot = 0;
      parity ^= 1u;
    }
  }

  // fold the per-warp column sums through shared memory (the ring is idle: every issued copy has been consumed)
  __syncthreads();
  float* red = reinterpret_cast<float*>(ring);  // [nwarps][kAcc][C]
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float* r0 = red + (static_cast<size_t>(warp) * kAcc) * C + v * 8;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        r0[t] = dg[i][t];
        r0[C + t] = db[i][t];
        if (COL) r0[2 * C + t] = dc[i][t];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kAcc * C; c += blockDim.x) {
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += red[static_cast<size_t>(w) * kAcc * C + c];
    partial[static_cast<size_t>(blockIdx.x) * kAcc * C + c] = acc;
  }
}

// kAcc-row variant of the finalize kernel below for the v4 partial layout [nparts][nacc][C];
// row 0 -> dgamma, row 1 -> dbeta (skipped when null), row 2 -> dcol (the producer's bias gradient)
__global__ void __launch_bounds__(1024)
norm_bwd_finalize_v4_kernel(const float* __restrict__ partial, int nparts, int C, int nacc,
                            __nv_bfloat16* __restrict__ dgamma, __nv_bfloat16* __restrict__ dbeta,
                            __nv_bfloat16* __restrict__ dcol, int accumulate, int accumulate_col) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;  // column in [0, nacc * C)
  float acc = 0.f;
  if (c < nacc * C) {
    // one partial row per SM (<= 160): every row this warp owns is requested before any is summed — one round trip
    float v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int pidx = warp + 32 * k;
      v[k] = pidx < nparts ? partial[static_cast<size_t>(pidx) * nacc * C + c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) acc += v[k];
    for (int pidx = warp + 160; pidx < nparts; pidx += 32) acc += partial[static_cast<size_t>(pidx) * nacc * C + c];
  }
  __shared__ float red[32][33];
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < nacc * C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += red[w][lane];
    const int which = c / C, col = c - which * C;
    __nv_bfloat16* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dcol);
    const int accf = which == 2 ? accumulate_col : accumulate;
    if (dst != nullptr) {
      if (accf) t += __bfloat162float(dst[col]);
      dst[col] = __float2bfloat16(t);
    }
  }
}

// one CTA per 32 columns: 8 warps stride over the partial rows, fold through shared memory
__global__ void __launch_bounds__(256)
norm_bwd_finalize_kernel(const float* __restrict__ partial, int nparts, int C,
                         __nv_bfloat16* __restrict__ dgamma, __nv_bfloat16* __restrict__ dbeta,
                         int accumulate) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;  // column in [0, 2C)
  float acc = 0.f;
  if (c < 2 * C) {
    const int which = c / C, col = c - which * C;
    for (int pidx = warp; pidx < nparts; pidx += 8)
      acc += partial[(static_cast<size_t>(pidx) * 2 + which) * C + col];
  }
  __shared__ float red[8][33];
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    const int which = c / C, col = c - which * C;
    __nv_bfloat16* dst = which == 0 ? dgamma : dbeta;
    if (dst != nullptr) {
      if (accumulate) t += __bfloat162float(dst[col]);
      dst[col] = __float2bfloat16(t);
    }
  }
}

template <bool RMS>
int launch_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
               const float* rstd, void* dx, float* partial, int rows, int C, int grid,
               cudaStream_t s, const void* res_grad) {
  const int nvec_lane = ((C >> 3) + 31) / 32;
  size_t smem = static_cast<size_t>(8) * C * sizeof(float);
  static const int variant = [] {
    const char* v = getenv("MAGGY_B200_NORM_BWD");
    return v == nullptr ? 2 : atoi(v);  // this path: 1 = first kernel, anything else = register-staged v2
  }();
#define MG_BWD_CASE(NV)                                                                            \
  {                                                                                                \
    auto k = (variant != 1 && NV <= 4) ? norm_bwd_kernel_v2<RMS, (NV <= 4 ? NV : 1), 1>            \
                                       : norm_bwd_kernel<RMS, NV>;                                 \
    if (smem > 48 * 1024)                                                                          \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
    k<<<grid, 128, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy),                           \
                              reinterpret_cast<const __nv_bfloat16*>(x),                            \
                              reinterpret_cast<const __nv_bfloat16*>(gamma), mean, rstd,            \
                              reinterpret_cast<__nv_bfloat16*>(dx), partial, rows, C,               \
                              reinterpret_cast<const __nv_bfloat16*>(res_grad));                    \
  }
  if (nvec_lane <= 1) MG_BWD_CASE(1)
  else if (nvec_lane <= 2) MG_BWD_CASE(2)
  else if (nvec_lane <= 3) MG_BWD_CASE(3)
  else if (nvec_lane <= 4) MG_BWD_CASE(4)
  else if (nvec_lane <= 8) MG_BWD_CASE(8)
  else return -1;
#undef MG_BWD_CASE
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

template <bool RMS, int NVEC>
void launch_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                int rows, int C, float eps, cudaStream_t s) {
  norm_fwd_kernel<RMS, NVEC><<<(rows + 3) / 4, 128, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(gamma),
      RMS ? nullptr : reinterpret_cast<const __nv_bfloat16*>(beta), reinterpret_cast<__nv_bfloat16*>(y),
      RMS ? nullptr : mean, rstd, rows, C, eps);
}

template <bool RMS>
int dispatch_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                 int rows, int C, float eps, cudaStream_t s) {
  // registers hold the row between the statistics pass and the normalise pass: size the buffer to the
  // row so small hidden sizes keep high occupancy (more bytes in flight per SM)
  const int nv = ((C >> 3) + 31) / 32;
  if (nv <= 1) launch_fwd<RMS, 1>(x, gamma, beta, y, mean, rstd, rows, C, eps, s);
  else if (nv <= 2) launch_fwd<RMS, 2>(x, gamma, beta, y, mean, rstd, rows, C, eps, s);
  else if (nv <= 3) launch_fwd<RMS, 3>(x, gamma, beta, y, mean, rstd, rows, C, eps, s);
  else if (nv <= 4) launch_fwd<RMS, 4>(x, gamma, beta, y, mean, rstd, rows, C, eps, s);
  else if (nv <= 8) launch_fwd<RMS, 8>(x, gamma, beta, y, mean, rstd, rows, C, eps, s);
  else launch_fwd<RMS, 16>(x, gamma, beta, y, mean, rstd, rows, C, eps, s);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // namespace

// rms: 0 = LayerNorm (mean/rstd saved), 1 = RMSNorm (rstd saved, beta ignored). C % 8 == 0, C <= 4096.
extern "C" int mg_norm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                           float* rstd, int rows, int C, float eps, int rms, void* stream) {
  if (C % 8 != 0 || C > kMaxVec * 256) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return rms ? dispatch_fwd<true>(x, gamma, beta, y, mean, rstd, rows, C, eps, s)
             : dispatch_fwd<false>(x, gamma, beta, y, mean, rstd, rows, C, eps, s);
}

// partial: fp32 scratch of at least mg_norm_bwd_grid(rows) * 2 * C floats. C % 8 == 0, C <= 2048.
extern "C" int mg_norm_bwd_grid(int rows) {
  int g = (rows + 3) / 4;
  return g < 592 ? g : 592;  // 148 SMs x 4 resident CTAs
}

extern "C" int mg_norm_bwd_res(const void* dy, const void* x, const void* gamma, const float* mean,
                               const float* rstd, void* dx, void* dgamma, void* dbeta, float* partial,
                               int rows, int C, int rms, int accumulate, const void* res_grad, void* stream);

extern "C" int mg_norm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                           const float* rstd, void* dx, void* dgamma, void* dbeta, float* partial,
                           int rows, int C, int rms, int accumulate, void* stream) {
  return mg_norm_bwd_res(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, partial, rows, C, rms, accumulate, nullptr,
                         stream);
}

namespace {
struct V4Plan {
  int warps, slots;
  size_t smem;
};
// shared-memory plan of the bulk-copy staged backward: at least 2 row slots per warp, ring <= ~200 KB
V4Plan plan_v4(int C, bool res, bool col) {
  const size_t slot = static_cast<size_t>(C) * 2 * (res ? 3 : 2);
  for (int warps : {12, 8, 4}) {
    int slots = static_cast<int>((200 * 1024) / (slot * warps));
    if (slots > 4) slots = 4;
    if (slots < 2) continue;
    const size_t bars = (static_cast<size_t>(warps) * slots * 8 + 127) & ~static_cast<size_t>(127);
    const size_t ring = slot * warps * slots, red = static_cast<size_t>(warps) * (col ? 3 : 2) * C * 4;
    return {warps, slots, bars + (ring > red ? ring : red)};
  }
  return {0, 0, 0};
}
int norm_bwd_variant() {
  static const int variant = [] {
    const char* v = getenv("MAGGY_B200_NORM_BWD");
    return v == nullptr ? 4 : atoi(v);
  }();
  return variant;
}
int sm_count() {
  static const int n = [] {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

template <bool RMS, bool RES, bool COL>
int launch_bwd_v4(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, void* dx,
                  float* partial, int rows, int C, const void* res_grad, int* grid_out, cudaStream_t s) {
  const V4Plan pl = plan_v4(C, RES, COL);
  if (pl.warps == 0) return -1;
  const int nvec_lane = ((C >> 3) + 31) / 32;
  int grid = (rows + pl.warps - 1) / pl.warps;
  if (grid > sm_count()) grid = sm_count();
  *grid_out = grid;
#define MG_V4_CASE(NV)                                                                                        \
  {                                                                                                           \
    auto k = norm_bwd_kernel_v4<RMS, NV, RES, COL>;                                                           \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));          \
    k<<<grid, pl.warps * 32, pl.smem, s>>>(                                                                   \
        reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x),                \
        reinterpret_cast<const __nv_bfloat16*>(gamma), mean, rstd, reinterpret_cast<__nv_bfloat16*>(dx),      \
        partial, rows, C, reinterpret_cast<const __nv_bfloat16*>(res_grad), pl.slots);                        \
  }
  if (nvec_lane <= 1) MG_V4_CASE(1)
  else if (nvec_lane <= 2) MG_V4_CASE(2)
  else if (nvec_lane <= 3) MG_V4_CASE(3)
  else if (nvec_lane <= 4) MG_V4_CASE(4)
  else return -1;
#undef MG_V4_CASE
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
}  // namespace

// 1 when mg_norm_bwd_col can produce the column sums of dx in the same pass for this width
extern "C" int mg_norm_bwd_supports_col(int C) {
  return norm_bwd_variant() == 4 && C % 8 == 0 && C <= 1024 && plan_v4(C, true, true).warps > 0 ? 1 : 0;
}

// dcol (optional, [C] bf16): column sums of the dx written — the bias gradient of the linear layer whose output this
// norm consumed (see norm_bwd_kernel_v4). Only valid when mg_norm_bwd_supports_col(C).
extern "C" int mg_norm_bwd_col(const void* dy, const void* x, const void* gamma, const float* mean,
                               const float* rstd, void* dx, void* dgamma, void* dbeta, void* dcol, float* partial,
                               int rows, int C, int rms, int accumulate, int accumulate_col, const void* res_grad,
                               void* stream) {
  if (C % 8 != 0 || C > 2048) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool v4 = norm_bwd_variant() == 4 && C <= 1024 && plan_v4(C, res_grad != nullptr, dcol != nullptr).warps > 0;
  if (!v4) {
    if (dcol != nullptr) return -2;
    return mg_norm_bwd_res(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, partial, rows, C, rms, accumulate, res_grad,
                           stream);
  }
  int grid = 0, rc;
  const bool res = res_grad != nullptr, col = dcol != nullptr;
#define MG_V4(R)                                                                                                     \
  (res ? (col ? launch_bwd_v4<R, true, true>(dy, x, gamma, mean, rstd, dx, partial, rows, C, res_grad, &grid, s)      \
              : launch_bwd_v4<R, true, false>(dy, x, gamma, mean, rstd, dx, partial, rows, C, res_grad, &grid, s))    \
       : (col ? launch_bwd_v4<R, false, true>(dy, x, gamma, mean, rstd, dx, partial, rows, C, res_grad, &grid, s)     \
              : launch_bwd_v4<R, false, false>(dy, x, gamma, mean, rstd, dx, partial, rows, C, res_grad, &grid, s)))
  rc = rms ? MG_V4(true) : MG_V4(false);
#undef MG_V4
  if (rc) return rc;
  const int nacc = col ? 3 : 2;
  norm_bwd_finalize_v4_kernel<<<(nacc * C + 31) / 32, 1024, 0, s>>>(
      partial, grid, C, nacc, reinterpret_cast<__nv_bfloat16*>(dgamma),
      rms ? nullptr : reinterpret_cast<__nv_bfloat16*>(dbeta), reinterpret_cast<__nv_bfloat16*>(dcol), accumulate,
      accumulate_col);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// res_grad (optional, [rows, C] bf16): added to dx — the gradient arriving through the residual connection
extern "C" int mg_norm_bwd_res(const void* dy, const void* x, const void* gamma, const float* mean,
                               const float* rstd, void* dx, void* dgamma, void* dbeta, float* partial,
                               int rows, int C, int rms, int accumulate, const void* res_grad, void* stream) {
  if (C % 8 != 0 || C > 2048) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int grid = mg_norm_bwd_grid(rows);
  int rc = rms ? launch_bwd<true>(dy, x, gamma, mean, rstd, dx, partial, rows, C, grid, s, res_grad)
               : launch_bwd<false>(dy, x, gamma, mean, rstd, dx, partial, rows, C, grid, s, res_grad);
  if (rc) return rc;
  norm_bwd_finalize_kernel<<<(2 * C + 31) / 32, 256, 0, s>>>(
      partial, grid, C, reinterpret_cast<__nv_bfloat16*>(dgamma),
      rms ? nullptr : reinterpret_cast<__nv_bfloat16*>(dbeta), accumulate);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
