// K7b — training BatchNorm for NHWC bf16 activations, fused with ReLU and the residual add (sm_100a).
//
// The ASHA ResNet-50 trial body (BASELINE.json configs 1-2; reference: maggy/core/executors/trial_executor.py:126-140
// runs the user's train_fn, whose BatchNorm layers hit ATen's channels-last kernels) spends 56 % of its step in those
// kernels (profiles/resnet_step_r2.md): they reach ~0.4 TB/s on [N*H*W, C] bf16 tensors. Everything here is a pure
// HBM stream, so the design goal is "every byte once, 16-byte accesses, enough loads in flight":
//
//   forward   stats   : per-channel sum / sum of squares            (1 read)
//             finalize: mean, 1/std, running statistics, scale/shift (tiny)
//             apply   : y = relu?(x * scale + shift (+ residual))    (1-2 reads, 1 write)
//   backward  reduce  : dz = dy * [y > 0]; sum dz, sum dz * xhat      (2-3 reads)
//             finalize: dgamma, dbeta, per-channel dx coefficients    (tiny)
//             elemt   : dx = a * dz + b * x + c (, dz for the skip)   (2-3 reads, 1-2 writes)
//
// Thread mapping: the tensor is a flat stream of 16-byte vectors (8 channels); the grid-wide thread count is a
// multiple of V = C / 8, so a thread meets the same 8 channels on every iteration: its per-channel constants and
// accumulators live in registers and consecutive threads read consecutive 16-byte vectors whatever C is (64 ... 2048).
#include "../include/mg_ptx.cuh"

namespace {
using namespace mg;

constexpr int kUnroll = 4;
constexpr int kPartCap = 296, kPartThreads = 512, kElemCap = 148 * 16;
constexpr int kRowsPerWarp = (kPartCap + 31) / 32;  // partial rows one finalize warp sums

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// fold per-thread accumulators acc[NACC][8] of the threads that share a vector column; partial[blockIdx][NACC][C].
// When V divides 32 the lanes of a warp that share a column are summed with shuffles first, so shared memory only
// carries one row per warp.
template <int NACC>
__device__ __forceinline__ void block_fold(float (&acc)[NACC][8], int V, int C, float* __restrict__ partial,
                                           float* red /* [blockDim][NACC*8+1] */) {
  constexpr int kW = NACC * 8 + 1;
  const int tid = threadIdx.x, T = blockDim.x;
  int rows, row;
  bool writer = true;
  if (V < 32 && (32 % V) == 0) {
    for (int o = V; o < 32; o <<= 1)
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[a][t] += __shfl_xor_sync(0xffffffffu, acc[a][t], o);
    rows = T >> 5;
    row = tid >> 5;
    writer = (tid & 31) < V;
  } else {
    rows = T / V;
    row = tid / V;
  }
  const int v_mine = tid % V;
  if (writer) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int t = 0; t < 8; ++t) red[(row * V + v_mine) * kW + a * 8 + t] = acc[a][t];
  }
  __syncthreads();
  for (int idx = tid; idx < V * NACC * 8; idx += T) {
    const int v = idx / (NACC * 8), j = idx - v * (NACC * 8);
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[(r * V + v) * kW + j];
    const int a = j >> 3, t = j & 7;
    partial[(static_cast<size_t>(blockIdx.x) * NACC + a) * C + v * 8 + t] = s;
  }
}

__global__ void __launch_bounds__(512)
bn_stats_kernel(const uint4* __restrict__ x, long long total, int V, int C, float* __restrict__ partial) {
  pdl_prologue();
  extern __shared__ float red[];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float acc[2][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = acc[1][t] = 0.f;
  for (; i + (kUnroll - 1) * stride < total; i += kUnroll * stride) {
    uint4 u[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) u[k] = ld_nc_v4(x + i + k * stride);
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      float f[8];
      unpack8(u[k], f);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        acc[0][t] += f[t];
        acc[1][t] = fmaf(f[t], f[t], acc[1][t]);
      }
    }
  }
  for (; i < total; i += stride) {
    float f[8];
    unpack8(ld_nc_v4(x + i), f);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      acc[0][t] += f[t];
      acc[1][t] = fmaf(f[t], f[t], acc[1][t]);
    }
  }
  block_fold<2>(acc, V, C, partial, red);
}

__device__ __forceinline__ float load_param(const void* p, int c, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[c]) : reinterpret_cast<const float*>(p)[c];
}
__device__ __forceinline__ void store_param(void* p, int c, int is_bf16, float v, int accumulate) {
  if (is_bf16) {
    __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(p);
    if (accumulate) v += __bfloat162float(q[c]);
    q[c] = __float2bfloat16(v);
  } else {
    float* q = reinterpret_cast<float*>(p);
    if (accumulate) v += q[c];
    q[c] = v;
  }
}

// 32 channels per CTA, 32 warps stride over the partial rows. One thread per channel finishes.
//   stats [0][C] = mean, [1][C] = 1/std (saved for backward); coef [0][C] = scale, [1][C] = shift
__global__ void __launch_bounds__(1024)
bn_fwd_finalize_kernel(const float* __restrict__ partial, int nparts, int C, double inv_m, double unbias,
                       const void* __restrict__ gamma, const void* __restrict__ beta, int param_bf16,
                       float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                       float* __restrict__ stats, float* __restrict__ coef, long long* __restrict__ num_batches) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f, q = 0.f;
  if (c < C) {
    // the kernel is one chain of memory latencies (9 us per launch in the ncu capture, 106 launches per ResNet step):
    // issue every row this warp owns (<= kPartCap / 32) before summing any, so the chain is ONE round trip
    float sv[kRowsPerWarp], qv[kRowsPerWarp];
#pragma unroll
    for (int k = 0; k < kRowsPerWarp; ++k) {
      const int p = warp + 32 * k;
      const bool ok = p < nparts;
      sv[k] = ok ? partial[(static_cast<size_t>(p) * 2) * C + c] : 0.f;
      qv[k] = ok ? partial[(static_cast<size_t>(p) * 2 + 1) * C + c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kRowsPerWarp; ++k) {
      s += sv[k];
      q += qv[k];
    }
  }
  __shared__ float rs[32][33], rq[32][33];
  rs[warp][lane] = s;
  rq[warp][lane] = q;
  __syncthreads();
  if (warp == 0 && c < C) {
    double ds = 0.0, dq = 0.0;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) {
      ds += rs[w][lane];
      dq += rq[w][lane];
    }
    const double mean = ds * inv_m;
    double var = dq * inv_m - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    stats[c] = static_cast<float>(mean);
    stats[C + c] = invstd;
    const float g = gamma ? load_param(gamma, c, param_bf16) : 1.f;
    const float b = beta ? load_param(beta, c, param_bf16) : 0.f;
    const float scale = g * invstd;
    coef[c] = scale;
    coef[C + c] = b - static_cast<float>(mean) * scale;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(var * unbias);
  }
  if (num_batches && blockIdx.x == 0 && threadIdx.x == 0) *num_batches += 1;
}

// inference / frozen statistics: scale and shift straight from the running statistics
__global__ void bn_eval_coef_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                    const void* __restrict__ gamma, const void* __restrict__ beta, int param_bf16,
                                    float eps, int C, float* __restrict__ coef) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(running_var[c] + eps);
  const float g = gamma ? load_param(gamma, c, param_bf16) : 1.f;
  const float b = beta ? load_param(beta, c, param_bf16) : 0.f;
  coef[c] = g * invstd;
  coef[C + c] = b - running_mean[c] * g * invstd;
}

template <bool RELU, bool RES>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, uint4* __restrict__ y,
                const float* __restrict__ coef, long long total, int V, int C) {
  pdl_prologue();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int v = static_cast<int>(i % V);
  float sc[8], sh[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    sc[t] = coef[v * 8 + t];
    sh[t] = coef[C + v * 8 + t];
  }
  auto one = [&](const uint4& ux, const uint4& ur) {
    float f[8], r[8];
    unpack8(ux, f);
    if (RES) unpack8(ur, r);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float o = fmaf(f[t], sc[t], sh[t]);
      if (RES) o += r[t];
      f[t] = RELU ? fmaxf(o, 0.f) : o;
    }
    return pack8f(f);
  };
  for (; i + (kUnroll - 1) * stride < total; i += kUnroll * stride) {
    uint4 u[kUnroll], r[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      u[k] = ld_nc_v4(x + i + k * stride);
      if (RES) r[k] = ld_nc_v4(res + i + k * stride);
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) y[i + k * stride] = one(u[k], RES ? r[k] : u[k]);
  }
  for (; i < total; i += stride) {
    const uint4 u = ld_nc_v4(x + i);
    y[i] = one(u, RES ? ld_nc_v4(res + i) : u);
  }
}

// partial[blockIdx][0][C] = sum dz, [1][C] = sum dz * xhat   (dz = dy masked by the ReLU that followed the norm)
template <bool RELU>
__global__ void __launch_bounds__(512)
bn_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, const uint4* __restrict__ x,
                     const float* __restrict__ stats, long long total, int V, int C, float* __restrict__ partial) {
  pdl_prologue();
  extern __shared__ float red[];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int v = static_cast<int>(i % V);
  float mean[8], invstd[8], acc[2][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    mean[t] = stats[v * 8 + t];
    invstd[t] = stats[C + v * 8 + t];
    acc[0][t] = acc[1][t] = 0.f;
  }
  auto one = [&](const uint4& ud, const uint4& uy, const uint4& ux) {
    float d[8], o[8], f[8];
    unpack8(ud, d);
    unpack8(ux, f);
    if (RELU) unpack8(uy, o);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float dz = (!RELU || o[t] > 0.f) ? d[t] : 0.f;
      acc[0][t] += dz;
      acc[1][t] = fmaf(dz, (f[t] - mean[t]) * invstd[t], acc[1][t]);
    }
  };
  constexpr int kU = 2;  // three streams: two rounds of loads keep the same bytes in flight as the 4-deep single stream
  for (; i + (kU - 1) * stride < total; i += kU * stride) {
    uint4 a[kU], b[kU], c[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      a[k] = ld_nc_v4(dy + i + k * stride);
      c[k] = ld_nc_v4(x + i + k * stride);
      if (RELU) b[k] = ld_nc_v4(y + i + k * stride);
    }
#pragma unroll
    for (int k = 0; k < kU; ++k) one(a[k], RELU ? b[k] : a[k], c[k]);
  }
  for (; i < total; i += stride) {
    const uint4 a = ld_nc_v4(dy + i);
    one(a, RELU ? ld_nc_v4(y + i) : a, ld_nc_v4(x + i));
  }
  block_fold<2>(acc, V, C, partial, red);
}

// dgamma / dbeta (written in the parameter dtype, optionally accumulated) and the dx coefficients:
//   dx = a * dz + b * x + c,  a = gamma / std,  b = -a * dgamma / (M std),  c = -a * dbeta / M - b * mean
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nparts, int C, float inv_m,
                       const void* __restrict__ gamma, int param_bf16, const float* __restrict__ stats,
                       void* __restrict__ dgamma, void* __restrict__ dbeta, int accumulate,
                       float* __restrict__ coef /* [3][C] */) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f, q = 0.f;
  if (c < C) {
    // the kernel is one chain of memory latencies (9 us per launch in the ncu capture, 106 launches per ResNet step):
    // issue every row this warp owns (<= kPartCap / 32) before summing any, so the chain is ONE round trip
    float sv[kRowsPerWarp], qv[kRowsPerWarp];
#pragma unroll
    for (int k = 0; k < kRowsPerWarp; ++k) {
      const int p = warp + 32 * k;
      const bool ok = p < nparts;
      sv[k] = ok ? partial[(static_cast<size_t>(p) * 2) * C + c] : 0.f;
      qv[k] = ok ? partial[(static_cast<size_t>(p) * 2 + 1) * C + c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kRowsPerWarp; ++k) {
      s += sv[k];
      q += qv[k];
    }
  }
  __shared__ float rs[32][33], rq[32][33];
  rs[warp][lane] = s;
  rq[warp][lane] = q;
  __syncthreads();
  if (warp == 0 && c < C) {
    float db = 0.f, dg = 0.f;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) {
      db += rs[w][lane];
      dg += rq[w][lane];
    }
    const float mean = stats[c], invstd = stats[C + c];
    const float g = gamma ? load_param(gamma, c, param_bf16) : 1.f;
    const float a = g * invstd;
    const float b = -a * invstd * dg * inv_m;
    coef[c] = a;
    coef[C + c] = b;
    coef[2 * C + c] = -a * db * inv_m - b * mean;
    if (dgamma) store_param(dgamma, c, param_bf16, dg, accumulate);
    if (dbeta) store_param(dbeta, c, param_bf16, db, accumulate);
  }
}

template <bool RELU, bool DZ>
__global__ void __launch_bounds__(256)
bn_bwd_elemt_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, const uint4* __restrict__ x,
                    const float* __restrict__ coef, uint4* __restrict__ dx, uint4* __restrict__ dz_out,
                    long long total, int V, int C) {
  pdl_prologue();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int v = static_cast<int>(i % V);
  float ca[8], cb[8], cc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    ca[t] = coef[v * 8 + t];
    cb[t] = coef[C + v * 8 + t];
    cc[t] = coef[2 * C + v * 8 + t];
  }
  auto one = [&](long long at, const uint4& ud, const uint4& uy, const uint4& ux) {
    float d[8], o[8], f[8];
    unpack8(ud, d);
    unpack8(ux, f);
    if (RELU) unpack8(uy, o);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float dz = (!RELU || o[t] > 0.f) ? d[t] : 0.f;
      d[t] = dz;
      f[t] = fmaf(ca[t], dz, fmaf(cb[t], f[t], cc[t]));
    }
    dx[at] = pack8f(f);
    if (DZ) dz_out[at] = pack8f(d);
  };
  constexpr int kU = 2;
  for (; i + (kU - 1) * stride < total; i += kU * stride) {
    uint4 a[kU], b[kU], c[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      a[k] = ld_nc_v4(dy + i + k * stride);
      c[k] = ld_nc_v4(x + i + k * stride);
      if (RELU) b[k] = ld_nc_v4(y + i + k * stride);
    }
#pragma unroll
    for (int k = 0; k < kU; ++k) one(i + k * stride, a[k], RELU ? b[k] : a[k], c[k]);
  }
  for (; i < total; i += stride) {
    const uint4 a = ld_nc_v4(dy + i);
    one(i, a, RELU ? ld_nc_v4(y + i) : a, ld_nc_v4(x + i));
  }
}

struct Geom {
  int threads, grid;
};
// threads per CTA: the largest multiple of V that fits max_threads; grid: every thread gets >= min_per_thread vectors,
// at most `cap` CTAs (the partial-sum kernels fold one row per CTA, so their grid also bounds the finalize work)
Geom geom(long long total, int V, int max_threads, int min_per_thread, int cap) {
  Geom g;
  g.threads = (max_threads / V) * V;
  long long blocks = total / (static_cast<long long>(g.threads) * min_per_thread);
  if (blocks < 1) blocks = 1;
  if (blocks > cap) blocks = cap;
  g.grid = static_cast<int>(blocks);
  return g;
}
int check() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
}  // namespace

// fp32 scratch the partial-sum kernels need: kPartCap * 2 * C floats
extern "C" long long mg_bn_partial_floats(int C) { return static_cast<long long>(kPartCap) * 2 * C; }

// x: [M, C] bf16 (NHWC flattened), C % 8 == 0, C <= 2048. stats [2][C], coef [2][C] fp32 outputs.
extern "C" int mg_bn_fwd_stats(const void* x, long long M, int C, const void* gamma, const void* beta, int param_bf16,
                               float* running_mean, float* running_var, float momentum, float eps, float* stats,
                               float* coef, long long* num_batches, float* partial, void* stream) {
  if (C % 8 != 0 || C > 2048 || M < 1) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int V = C / 8;
  const long long total = M * V;
  const Geom g = geom(total, V, kPartThreads, 8, kPartCap);
  launch_pdl(bn_stats_kernel, dim3(g.grid), dim3(g.threads), g.threads * 17 * sizeof(float), s, 1, reinterpret_cast<const uint4*>(x), total, V, C,
                                                                          partial);
  const double inv_m = 1.0 / static_cast<double>(M);
  const double unbias = M > 1 ? static_cast<double>(M) / static_cast<double>(M - 1) : 1.0;
  launch_pdl(bn_fwd_finalize_kernel, dim3((C + 31) / 32), dim3(1024), 0, s, 1, partial, g.grid, C, inv_m, unbias, gamma, beta, param_bf16,
                                                       running_mean, running_var, momentum, eps, stats, coef,
                                                       num_batches);
  return check();
}

extern "C" int mg_bn_eval_coef(const float* running_mean, const float* running_var, const void* gamma,
                               const void* beta, int param_bf16, float eps, int C, float* coef, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(bn_eval_coef_kernel, dim3((C + 127) / 128), dim3(128), 0, s, 1, running_mean, running_var, gamma, beta, param_bf16, eps, C, coef);
  return check();
}

// y = relu?(x * scale + shift (+ res))
extern "C" int mg_bn_apply(const void* x, const void* res, void* y, const float* coef, long long M, int C, int relu,
                           void* stream) {
  if (C % 8 != 0 || C > 2048) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int V = C / 8;
  const long long total = M * V;
  const Geom g = geom(total, V, 256, 8, kElemCap);
  const uint4 *xp = reinterpret_cast<const uint4*>(x), *rp = reinterpret_cast<const uint4*>(res);
  uint4* yp = reinterpret_cast<uint4*>(y);
  if (relu) {
    if (res) launch_pdl(bn_apply_kernel<true, true>, dim3(g.grid), dim3(g.threads), 0, s, 1, xp, rp, yp, coef, total, V, C);
    else launch_pdl(bn_apply_kernel<true, false>, dim3(g.grid), dim3(g.threads), 0, s, 1, xp, rp, yp, coef, total, V, C);
  } else {
    if (res) launch_pdl(bn_apply_kernel<false, true>, dim3(g.grid), dim3(g.threads), 0, s, 1, xp, rp, yp, coef, total, V, C);
    else launch_pdl(bn_apply_kernel<false, false>, dim3(g.grid), dim3(g.threads), 0, s, 1, xp, rp, yp, coef, total, V, C);
  }
  return check();
}

// dy, y (the forward output; read only when relu), x: [M, C] bf16. dx always written; dz (optional) receives the
// ReLU-masked dy — the gradient of the residual input. coef: [3][C] fp32 scratch.
extern "C" int mg_bn_bwd(const void* dy, const void* y, const void* x, const float* stats, const void* gamma,
                         int param_bf16, void* dgamma, void* dbeta, int accumulate, void* dx, void* dz, long long M,
                         int C, int relu, float* coef, float* partial, void* stream) {
  if (C % 8 != 0 || C > 2048 || M < 1) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int V = C / 8;
  const long long total = M * V;
  const uint4 *dyp = reinterpret_cast<const uint4*>(dy), *yp = reinterpret_cast<const uint4*>(y),
              *xp = reinterpret_cast<const uint4*>(x);
  const Geom g = geom(total, V, kPartThreads, 8, kPartCap);
  const size_t smem = g.threads * 17 * sizeof(float);
  if (relu) launch_pdl(bn_bwd_reduce_kernel<true>, dim3(g.grid), dim3(g.threads), smem, s, 1, dyp, yp, xp, stats, total, V, C, partial);
  else launch_pdl(bn_bwd_reduce_kernel<false>, dim3(g.grid), dim3(g.threads), smem, s, 1, dyp, yp, xp, stats, total, V, C, partial);
  launch_pdl(bn_bwd_finalize_kernel, dim3((C + 31) / 32), dim3(1024), 0, s, 1, partial, g.grid, C, 1.0f / static_cast<float>(M), gamma,
                                                       param_bf16, stats, dgamma, dbeta, accumulate, coef);
  const Geom e = geom(total, V, 256, 8, kElemCap);
  uint4 *dxp = reinterpret_cast<uint4*>(dx), *dzp = reinterpret_cast<uint4*>(dz);
  if (relu) {
    if (dz) launch_pdl(bn_bwd_elemt_kernel<true, true>, dim3(e.grid), dim3(e.threads), 0, s, 1, dyp, yp, xp, coef, dxp, dzp, total, V, C);
    else launch_pdl(bn_bwd_elemt_kernel<true, false>, dim3(e.grid), dim3(e.threads), 0, s, 1, dyp, yp, xp, coef, dxp, dzp, total, V, C);
  } else {
    if (dz) launch_pdl(bn_bwd_elemt_kernel<false, true>, dim3(e.grid), dim3(e.threads), 0, s, 1, dyp, yp, xp, coef, dxp, dzp, total, V, C);
    else launch_pdl(bn_bwd_elemt_kernel<false, false>, dim3(e.grid), dim3(e.threads), 0, s, 1, dyp, yp, xp, coef, dxp, dzp, total, V, C);
  }
  return check();
}
