// Fused softmax-cross-entropy (forward + backward in one pass over a shared-memory resident row),
// column-sum (bias gradients), GELU, and token/position embedding kernels for sm_100a.
// All are HBM-bound: 16-byte vector accesses, one read + one write of the big operand.
#include "../include/mg_ptx.cuh"

namespace {
using namespace mg;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* scratch) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? scratch[lane] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;
}

// logits [rows, ld] bf16 are overwritten with d(mean loss)/d(logits) * grad_scale.
// losses[row] = per-row NLL (0 for ignored rows). Columns >= v_real are padding.
__global__ void __launch_bounds__(1024)
ce_fwd_bwd_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ targets,
                  float* __restrict__ losses, int v_real, int v_pad, long long ld,
                  float grad_scale, const float* __restrict__ grad_scale_dev, int ignore_index) {
  extern __shared__ __align__(16) uint8_t smem_row[];
  __shared__ float scratch[32];
  uint4* srow = reinterpret_cast<uint4*>(smem_row);
  const int row = blockIdx.x;
  uint4* grow = reinterpret_cast<uint4*>(logits + static_cast<size_t>(row) * ld);
  const int nvec = v_pad >> 3;
  const long long tgt = targets[row];
  const bool ignored = (tgt == ignore_index) || tgt < 0 || tgt >= v_real;

  // The whole row (100 KB at V = 50304) is fetched with a handful of bulk async copies issued by one thread — the
  // per-thread 16-byte loads of the first version kept too few bytes in flight (48 % of the copy roofline,
  // profiles/bw_kernels.md) — and the gradient row leaves the same way.
  __shared__ uint64_t row_bar;
  const uint32_t row_bytes = static_cast<uint32_t>(nvec) * 16u;
  constexpr uint32_t kChunk = 32768;
  if (threadIdx.x == 0) {
    mbar_init(&row_bar, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(&row_bar, row_bytes);
    for (uint32_t off = 0; off < row_bytes; off += kChunk)
      bulk_load_1d(smem_row + off, reinterpret_cast<const uint8_t*>(grow) + off, min(kChunk, row_bytes - off),
                   &row_bar);
  }
  __syncthreads();  // barrier initialised before anyone polls it
  mbar_wait(&row_bar, 0);
  float mx = -INFINITY;
  const int full_vecs = v_real >> 3;  // vectors with no padding column: no per-element bound checks
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 u = srow[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    if (i < full_vecs) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = unpack_bf16x2(w[t]);
        mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = unpack_bf16x2(w[t]);
        int c = i * 8 + 2 * t;
        if (c < v_real) mx = fmaxf(mx, f.x);
        if (c + 1 < v_real) mx = fmaxf(mx, f.y);
      }
    }
  }
  mx = block_reduce(mx, true, scratch);
  const float kLog2e = 1.4426950408889634f;
  // target logit must be read before the row buffer is overwritten with exp() values below
  float x_tgt = 0.f;
  if (threadIdx.x == 0 && !ignored) x_tgt = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(smem_row)[tgt]);
  __syncthreads();
  // exp() is evaluated ONCE per element (MUFU.EX2 runs at 16/clk/SM: two passes of 50k exps per row would
  // cost more than the HBM traffic of this kernel); e = exp(x - max) is parked in the row buffer as bf16
  float sum = 0.f, sum_b = 0.f;
  const float mxl = mx * kLog2e;
  auto ex2 = [](float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
  };
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 u = srow[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t e[4];
    if (i < full_vecs) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = unpack_bf16x2(w[t]);
        const float e0 = ex2(fmaf(f.x, kLog2e, -mxl)), e1 = ex2(fmaf(f.y, kLog2e, -mxl));
        sum += e0;
        sum_b += e1;
        e[t] = pack_bf16x2(e0, e1);
      }
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = unpack_bf16x2(w[t]);
        int c = i * 8 + 2 * t;
        float e0 = (c < v_real) ? ex2(fmaf(f.x, kLog2e, -mxl)) : 0.f;
        float e1 = (c + 1 < v_real) ? ex2(fmaf(f.y, kLog2e, -mxl)) : 0.f;
        sum += e0 + e1;
        e[t] = pack_bf16x2(e0, e1);
      }
    }
    srow[i] = make_uint4(e[0], e[1], e[2], e[3]);
  }
  sum += sum_b;
  sum = block_reduce(sum, false, scratch);
  const float inv_sum = 1.0f / sum;
  if (threadIdx.x == 0) {
    float loss = 0.f;
    if (!ignored) loss = logf(sum) + mx - x_tgt;
    losses[row] = loss;
  }
  // grad_scale_dev (optional): a device scalar such as 1 / (number of non-ignored rows), known only on the device
  const float gs = ignored ? 0.f : (grad_scale_dev ? grad_scale * __ldg(grad_scale_dev) : grad_scale);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 u = srow[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
    const float k = inv_sum * gs;
    if ((tgt >> 3) != i) {  // f = exp(x - max) from the pass above (0 in the padding columns)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = unpack_bf16x2(w[t]);
        o[t] = pack_bf16x2(f.x * k, f.y * k);
      }
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = unpack_bf16x2(w[t]);
        int c = i * 8 + 2 * t;
        float p0 = f.x * inv_sum, p1 = f.y * inv_sum;
        if (c == tgt) p0 -= 1.0f;
        if (c + 1 == tgt) p1 -= 1.0f;
        o[t] = pack_bf16x2(p0 * gs, p1 * gs);
      }
    }
    srow[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t off = 0; off < row_bytes; off += kChunk)
      bulk_store_1d(reinterpret_cast<uint8_t*>(grow) + off, smem_row + off, min(kChunk, row_bytes - off));
    tma_store_commit();
    tma_store_wait_read<0>();  // shared memory must outlive the copy engine's reads
  }
}

// Vocabularies whose row does not fit in shared memory (e.g. Llama-3: 128256 columns = 250 KB): three streaming passes
// over the row (max, sum of exponentials, gradient); the re-reads hit L2 (a row is a few hundred KB).
__global__ void __launch_bounds__(1024)
ce_fwd_bwd_stream_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ targets,
                         float* __restrict__ losses, int v_real, int v_pad, long long ld, float grad_scale,
                         const float* __restrict__ grad_scale_dev, int ignore_index) {
  __shared__ float scratch[32];
  const int row = blockIdx.x;
  uint4* grow = reinterpret_cast<uint4*>(logits + static_cast<size_t>(row) * ld);
  const int nvec = v_pad >> 3;
  const long long tgt = targets[row];
  const bool ignored = (tgt == ignore_index) || tgt < 0 || tgt >= v_real;
  const float kLog2e = 1.4426950408889634f;
  auto ex2 = [](float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
  };
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 u = grow[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      const int c = i * 8 + 2 * t;
      if (c < v_real) mx = fmaxf(mx, f.x);
      if (c + 1 < v_real) mx = fmaxf(mx, f.y);
    }
  }
  mx = block_reduce(mx, true, scratch);
  const float mxl = mx * kLog2e;
  float sum = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 u = grow[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      const int c = i * 8 + 2 * t;
      if (c < v_real) sum += ex2(fmaf(f.x, kLog2e, -mxl));
      if (c + 1 < v_real) sum += ex2(fmaf(f.y, kLog2e, -mxl));
    }
  }
  sum = block_reduce(sum, false, scratch);
  const float inv_sum = 1.0f / sum;
  if (threadIdx.x == 0) {
    float loss = 0.f;
    if (!ignored) loss = logf(sum) + mx - __bfloat162float(logits[static_cast<size_t>(row) * ld + tgt]);
    losses[row] = loss;
  }
  __syncthreads();  // the target logit is read before any thread overwrites the row
  // grad_scale_dev (optional): a device scalar such as 1 / (number of non-ignored rows), known only on the device
  const float gs = ignored ? 0.f : (grad_scale_dev ? grad_scale * __ldg(grad_scale_dev) : grad_scale);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 u = grow[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      const int c = i * 8 + 2 * t;
      float p0 = (c < v_real) ? ex2(fmaf(f.x, kLog2e, -mxl)) * inv_sum : 0.f;
      float p1 = (c + 1 < v_real) ? ex2(fmaf(f.y, kLog2e, -mxl)) * inv_sum : 0.f;
      if (c == tgt) p0 -= 1.0f;
      if (c + 1 == tgt) p1 -= 1.0f;
      o[t] = pack_bf16x2(p0 * gs, p1 * gs);
    }
    st_na_v4(grow + i, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// partial[blockIdx.y][n] = sum over this slab's rows of x[row][n]
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, int M, int N, long long ld,
                      float* __restrict__ partial) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vcol = blockIdx.x * 32 + lane;  // vector column (8 elements)
  const bool ok = vcol * 8 < N;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int rows_per_slab = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per_slab;
  const int r1 = min(M, r0 + rows_per_slab);
  if (ok) {
    const uint4* base = reinterpret_cast<const uint4*>(x) + vcol;
    const size_t ldv = static_cast<size_t>(ld) >> 3;
    int r = r0 + warp;
    // four independent 16-byte loads in flight per thread (the one-load-per-iteration version ran at about half
    // of the copy bandwidth: profiles/train_step_breakdown.md)
    for (; r + 24 < r1; r += 32) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ld_nc_v4(base + static_cast<size_t>(r + 8 * k) * ldv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 a = unpack_bf16x2(u[k].x), b = unpack_bf16x2(u[k].y), c = unpack_bf16x2(u[k].z),
               d = unpack_bf16x2(u[k].w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      }
    }
    for (; r < r1; r += 8) {
      uint4 u = ld_nc_v4(base + static_cast<size_t>(r) * ldv);
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
             d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
  __shared__ float red[8][32][9];
#pragma unroll
  for (int t = 0; t < 8; ++t) red[warp][lane][t] = acc[t];
  __syncthreads();
  // thread (warp w, lane l) finalises element t = w of vector column l
  if (ok) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][lane][warp];
    partial[static_cast<size_t>(blockIdx.y) * N + vcol * 8 + warp] = s;
  }
}

// one CTA per 32 columns: warp w sums slabs w, w + 8, ... (coalesced 128-byte rows), then the 8 warps combine
__global__ void __launch_bounds__(256)
colsum_finalize_kernel(const float* __restrict__ partial, int nslabs, int N,
                       __nv_bfloat16* __restrict__ out, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N)
    for (int i = warp; i < nslabs; i += 8) s += partial[static_cast<size_t>(i) * N + n];
  __shared__ float red[8][33];
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    if (accumulate) t += __bfloat162float(out[n]);
    out[n] = __float2bfloat16(t);
  }
}

__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 u = ld_nc_v4(reinterpret_cast<const uint4*>(x) + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 f = unpack_bf16x2(w[t]);
      o[t] = pack_bf16x2(gelu_tanh(f.x), gelu_tanh(f.y));
    }
    st_na_v4(reinterpret_cast<uint4*>(y) + i, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                __nv_bfloat16* __restrict__ dx, long long n) {
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 u = ld_nc_v4(reinterpret_cast<const uint4*>(x) + i);
    uint4 g = ld_nc_v4(reinterpret_cast<const uint4*>(dy) + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 f = unpack_bf16x2(w[t]), d = unpack_bf16x2(gw[t]);
      o[t] = pack_bf16x2(d.x * dgelu_tanh(f.x), d.y * dgelu_tanh(f.y));
    }
    st_na_v4(reinterpret_cast<uint4*>(dx) + i, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// out[t, :] = wte[idx[t], :] + wpe[t % T, :]
__global__ void __launch_bounds__(128)
embed_fwd_kernel(const long long* __restrict__ idx, const __nv_bfloat16* __restrict__ wte,
                 const __nv_bfloat16* __restrict__ wpe, __nv_bfloat16* __restrict__ out, int ntok,
                 int T, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 4 + warp;
  if (tok >= ntok) return;
  const long long id = idx[tok];
  const uint4* te = reinterpret_cast<const uint4*>(wte + static_cast<size_t>(id) * C);
  const uint4* pe = reinterpret_cast<const uint4*>(wpe + static_cast<size_t>(tok % T) * C);
  uint4* o = reinterpret_cast<uint4*>(out + static_cast<size_t>(tok) * C);
  for (int v = lane; v < (C >> 3); v += 32) {
    if (wpe == nullptr) {  // token embedding only (rotary-position models)
      o[v] = __ldg(te + v);
      continue;
    }
    uint4 a = __ldg(te + v), b = __ldg(pe + v);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t r[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 fa = unpack_bf16x2(aw[t]), fb = unpack_bf16x2(bw[t]);
      r[t] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    }
    o[v] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

// dwte[idx[t], :] += dout[t, :]   (bf16x2 reductions; dwte already holds the tied LM-head wgrad)
// dwpe[p, :]      = sum_b dout[b*T + p, :]   (one warp per position, deterministic)
__global__ void __launch_bounds__(128)
embed_bwd_wte_kernel(const long long* __restrict__ idx, const __nv_bfloat16* __restrict__ dout,
                     __nv_bfloat16* __restrict__ dwte, int ntok, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 4 + warp;
  if (tok >= ntok) return;
  const long long id = idx[tok];
  const __nv_bfloat162* src = reinterpret_cast<const __nv_bfloat162*>(dout + static_cast<size_t>(tok) * C);
  __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(dwte + static_cast<size_t>(id) * C);
  for (int v = lane; v < (C >> 1); v += 32) atomicAdd(dst + v, src[v]);
}

__global__ void __launch_bounds__(128)
embed_bwd_wpe_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dwpe, int B,
                     int T, int C, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pos = blockIdx.x * 4 + warp;
  if (pos >= T) return;
  for (int v = lane; v < (C >> 3); v += 32) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) {
      uint4 u = ld_nc_v4(reinterpret_cast<const uint4*>(dout + (static_cast<size_t>(b) * T + pos) * C) + v);
      float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
             d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += bb.x; acc[3] += bb.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
    uint4* dst = reinterpret_cast<uint4*>(dwpe + static_cast<size_t>(pos) * C) + v;
    if (accumulate) {
      uint4 u = *dst;
      float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
             d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += bb.x; acc[3] += bb.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
    *dst = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                      pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

inline int grid_for(long long nvec) {
  long long g = (nvec + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}
inline int check() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // namespace

extern "C" int mg_ce_fwd_bwd(void* logits, const long long* targets, float* losses, int rows,
                             int v_real, int v_pad, long long ld, float grad_scale, const float* grad_scale_dev,
                             int ignore_index, void* stream) {
  if (v_pad % 8 != 0 || ld % 8 != 0) return -1;
  size_t smem = static_cast<size_t>(v_pad) * 2;
  if (smem > 110 * 1024) {
    // the row would not leave room for two CTAs per SM (or does not fit at all): streaming variant
    ce_fwd_bwd_stream_kernel<<<rows, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(logits), targets, losses, v_real, v_pad, ld, grad_scale, grad_scale_dev,
        ignore_index);
    return check();
  }
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(ce_fwd_bwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return -300 - static_cast<int>(e);
    configured = smem;
  }
  // one row per CTA, the row lives in shared memory: big vocabularies need many threads per CTA to
  // keep enough 16-byte loads in flight (2 CTAs x 1024 threads per SM at 100 KB/row)
  const int threads = v_pad >= 16384 ? 1024 : (v_pad >= 4096 ? 512 : 256);
  ce_fwd_bwd_kernel<<<rows, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(logits), targets, losses, v_real, v_pad, ld, grad_scale, grad_scale_dev,
      ignore_index);
  return check();
}

// many partial rows (the GEMM epilogue writes one per 32-row warp slice): 32 warps per CTA, eight rows requested per
// warp before any is summed, so 2048 rows are eight round trips instead of a 256-deep latency chain
__global__ void __launch_bounds__(1024)
colsum_finalize_wide_kernel(const float* __restrict__ partial, int nslabs, int N, __nv_bfloat16* __restrict__ out,
                            int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N) {
    int i = warp;
    for (; i + 7 * 32 < nslabs; i += 8 * 32) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = partial[static_cast<size_t>(i + 32 * k) * N + n];
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
    for (; i < nslabs; i += 32) s += partial[static_cast<size_t>(i) * N + n];
  }
  __shared__ float red[32][33];
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && n < N) {
    float t = 0.f;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) t += red[w][lane];
    if (accumulate) t += __bfloat162float(out[n]);
    out[n] = __float2bfloat16(t);
  }
}

// out[n] (+)= sum over `nslabs` rows of an fp32 partial buffer [nslabs][N] (written by the dGELU GEMM epilogue)
extern "C" int mg_colsum_finalize(const float* partial, int nslabs, int N, void* out, int accumulate, void* stream) {
  if (nslabs <= 0 || N <= 0) return -1;
  colsum_finalize_wide_kernel<<<(N + 31) / 32, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, nslabs, N, reinterpret_cast<__nv_bfloat16*>(out), accumulate);
  return check();
}

extern "C" int mg_colsum_slabs(int M) {
  int s = (M + 127) / 128;  // >= 128 rows per slab keeps the 4-deep load loop busy
  return s < 1 ? 1 : (s > 256 ? 256 : s);
}

// partial: fp32 scratch of mg_colsum_slabs(M) * N floats
extern "C" int mg_colsum_bf16(const void* x, int M, int N, long long ld, void* out, float* partial,
                              int accumulate, void* stream) {
  if (N % 8 != 0 || ld % 8 != 0) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int slabs = mg_colsum_slabs(M);
  dim3 grid((N / 8 + 31) / 32, slabs);
  colsum_partial_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), M, N, ld,
                                             partial);
  colsum_finalize_kernel<<<(N + 31) / 32, 256, 0, s>>>(partial, slabs, N,
                                                         reinterpret_cast<__nv_bfloat16*>(out),
                                                         accumulate);
  return check();
}

extern "C" int mg_gelu_fwd(const void* x, void* y, long long n, void* stream) {
  if (n % 8 != 0) return -1;
  gelu_fwd_kernel<<<grid_for(n >> 3), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), n);
  return check();
}

extern "C" int mg_gelu_bwd(const void* x, const void* dy, void* dx, long long n, void* stream) {
  if (n % 8 != 0) return -1;
  gelu_bwd_kernel<<<grid_for(n >> 3), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy),
      reinterpret_cast<__nv_bfloat16*>(dx), n);
  return check();
}

extern "C" int mg_embed_fwd(const long long* idx, const void* wte, const void* wpe, void* out,
                            int ntok, int T, int C, void* stream) {
  if (C % 8 != 0) return -1;
  embed_fwd_kernel<<<(ntok + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      idx, reinterpret_cast<const __nv_bfloat16*>(wte), reinterpret_cast<const __nv_bfloat16*>(wpe),
      reinterpret_cast<__nv_bfloat16*>(out), ntok, T, C);
  return check();
}

extern "C" int mg_embed_bwd(const long long* idx, const void* dout, void* dwte, void* dwpe, int B,
                            int T, int C, int accumulate_wpe, void* stream) {
  if (C % 8 != 0) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  embed_bwd_wte_kernel<<<(B * T + 3) / 4, 128, 0, s>>>(
      idx, reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<__nv_bfloat16*>(dwte),
      B * T, C);
  if (dwpe != nullptr)
    embed_bwd_wpe_kernel<<<(T + 3) / 4, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dout),
                                                     reinterpret_cast<__nv_bfloat16*>(dwpe), B, T, C,
                                                     accumulate_wpe);
  return check();
}
