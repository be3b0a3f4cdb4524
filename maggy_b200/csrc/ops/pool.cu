// K7c — max pooling for NHWC bf16 activations (the 3x3 / 2 pool after the ResNet stem), forward + backward.
//
// ATen's channels-last kernels take 0.21 ms (forward) + 0.44 ms (backward) on the [64, 112, 112, 64] stem output
// (profiles/resnet_step_r2.md); the tensors are 103 MB in, 26 MB out, so the copy roofline is ~25 us per direction.
// Forward: one thread per 8 output channels, 16-byte loads of the (overlapping, cache-resident) window, the position
// of the first maximum is kept as one byte per element. Backward is a GATHER over input pixels — each input vector
// looks at the <= ceil(k/stride)^2 windows that cover it and adds dy where the stored position is its own — so there
// are no atomics and dx is written exactly once, in full 16-byte vectors.
// Tie-breaking matches torch.nn.functional.max_pool2d (first maximum in row-major window order; NaN propagates).
#include "../include/mg_ptx.cuh"

namespace {
using namespace mg;

__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint2* __restrict__ idx, int N, int H, int W,
                   int V, int P, int Q, int k, int stride, int pad) {
  pdl_prologue();
  const long long total = static_cast<long long>(N) * P * Q * V;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % V);
    long long t = i / V;
    const int q = static_cast<int>(t % Q);
    t /= Q;
    const int p = static_cast<int>(t % P), n = static_cast<int>(t / P);
    float best[8];
    uint32_t pos[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      best[e] = -INFINITY;
      pos[e] = 0;
    }
    for (int r = 0; r < k; ++r) {
      const int h = p * stride + r - pad;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int w = q * stride + s - pad;
        if (w < 0 || w >= W) continue;
        const uint4 u = __ldg(x + ((static_cast<long long>(n) * H + h) * W + w) * V + v);
        const uint32_t ws[4] = {u.x, u.y, u.z, u.w};
        const uint32_t here = static_cast<uint32_t>(r * k + s);
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 f = unpack_bf16x2(ws[e2]);
          if (f.x > best[2 * e2] || f.x != f.x) { best[2 * e2] = f.x; pos[2 * e2] = here; }
          if (f.y > best[2 * e2 + 1] || f.y != f.y) { best[2 * e2 + 1] = f.y; pos[2 * e2 + 1] = here; }
        }
      }
    }
    y[i] = make_uint4(pack_bf16x2(best[0], best[1]), pack_bf16x2(best[2], best[3]), pack_bf16x2(best[4], best[5]),
                      pack_bf16x2(best[6], best[7]));
    idx[i] = make_uint2(pos[0] | (pos[1] << 8) | (pos[2] << 16) | (pos[3] << 24),
                        pos[4] | (pos[5] << 8) | (pos[6] << 16) | (pos[7] << 24));
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx, uint4* __restrict__ dx, int N, int H,
                   int W, int V, int P, int Q, int k, int stride, int pad) {
  pdl_prologue();
  const long long total = static_cast<long long>(N) * H * W * V;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % V);
    long long t = i / V;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int h = static_cast<int>(t % H), n = static_cast<int>(t / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // windows p with p*stride - pad <= h <= p*stride - pad + k - 1
    const int p_lo = max(0, (h + pad - k + stride) / stride), p_hi = min(P - 1, (h + pad) / stride);
    const int q_lo = max(0, (w + pad - k + stride) / stride), q_hi = min(Q - 1, (w + pad) / stride);
    for (int p = p_lo; p <= p_hi; ++p) {
      const int r = h - (p * stride - pad);
      for (int q = q_lo; q <= q_hi; ++q) {
        const int s = w - (q * stride - pad);
        const uint32_t here = static_cast<uint32_t>(r * k + s);
        const long long o = ((static_cast<long long>(n) * P + p) * Q + q) * V + v;
        const uint2 id = __ldg(idx + o);
        const uint4 g = __ldg(dy + o);
        const uint32_t gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 f = unpack_bf16x2(gs[e2]);
          const uint32_t word = e2 < 2 ? id.x : id.y;
          const uint32_t p0 = (word >> (16 * (e2 & 1))) & 0xffu, p1 = (word >> (16 * (e2 & 1) + 8)) & 0xffu;
          if (p0 == here) acc[2 * e2] += f.x;
          if (p1 == here) acc[2 * e2 + 1] += f.y;
        }
      }
    }
    dx[i] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                       pack_bf16x2(acc[6], acc[7]));
  }
}

int check() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
}  // namespace

// x: [N,H,W,C] bf16, y: [N,P,Q,C] bf16, idx: [N,P,Q,C] uint8 (window position of the maximum). C % 8 == 0, k*k <= 255.
extern "C" int mg_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, int k, int stride, int pad,
                              void* stream) {
  if (C % 8 != 0 || k * k > 255 || 2 * pad > k) return -1;
  const int P = (H + 2 * pad - k) / stride + 1, Q = (W + 2 * pad - k) / stride + 1;
  if (P <= 0 || Q <= 0) return -1;
  const long long total = static_cast<long long>(N) * P * Q * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  launch_pdl(maxpool_fwd_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), reinterpret_cast<uint2*>(idx), N, H, W, C / 8, P, Q,
      k, stride, pad);
  return check();
}

extern "C" int mg_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, int k, int stride,
                              int pad, void* stream) {
  if (C % 8 != 0 || k * k > 255 || 2 * pad > k) return -1;
  const int P = (H + 2 * pad - k) / stride + 1, Q = (W + 2 * pad - k) / stride + 1;
  if (P <= 0 || Q <= 0) return -1;
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  launch_pdl(maxpool_bwd_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
      reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(idx), reinterpret_cast<uint4*>(dx), N, H, W,
      C / 8, P, Q, k, stride, pad);
  return check();
}
