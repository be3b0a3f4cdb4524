// K4 — fused multi-tensor Adam/AdamW over one flat parameter arena (sm_100a).
//
// The reference forces DeepSpeed's fused Adam OFF (torch_adam=True,
// maggy/core/executors/torch_dist_executor.py:403-405) and steps torch.optim.* per tensor. Here the
// whole model is one flat arena [decay params | no-decay params]: fp32 master + m + v, bf16 gradient
// in, bf16 compute copy out, one launch, 16-byte vector accesses, grid sized to 148 SMs x 8.
// The same update body is reused by the fused reduce-scatter kernel (csrc/comm/rs_adam.cu).
#include "../include/mg_ptx.cuh"
#include "adam_math.cuh"

namespace {
using namespace mg;

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ v,
                 const __nv_bfloat16* __restrict__ grad, __nv_bfloat16* __restrict__ param_bf16,
                 long long n, AdamHyper h_in) {
  const AdamHyper h = adam_resolve(h_in);
  const float clip = adam_clip_coef(h);
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 g = ld_nc_v4(reinterpret_cast<const uint4*>(grad) + i);
    float gf[8];
    {
      float2 a = unpack_bf16x2(g.x), b = unpack_bf16x2(g.y), c = unpack_bf16x2(g.z),
             d = unpack_bf16x2(g.w);
      gf[0] = a.x; gf[1] = a.y; gf[2] = b.x; gf[3] = b.y;
      gf[4] = c.x; gf[5] = c.y; gf[6] = d.x; gf[7] = d.y;
    }
    const float wd = (i * 8 < h.decay_boundary) ? h.weight_decay : 0.f;
    uint4 out;
    adam_update8(master + i * 8, m + i * 8, v + i * 8, gf, h, wd, clip, out);
    st_na_v4(reinterpret_cast<uint4*>(param_bf16) + i, out);
  }
}

// fp32 arena (generic torch modules keep fp32 parameters): the parameter arena IS the master copy, gradients are fp32
__global__ void __launch_bounds__(256)
adam_flat_f32_kernel(float* __restrict__ param, float* __restrict__ m, float* __restrict__ v,
                     const float* __restrict__ grad, long long n, AdamHyper h_in) {
  const AdamHyper h = adam_resolve(h_in);
  const float clip = adam_clip_coef(h);
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 g0 = reinterpret_cast<const float4*>(grad)[2 * i], g1 = reinterpret_cast<const float4*>(grad)[2 * i + 1];
    const float gf[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float wd = (i * 8 < h.decay_boundary) ? h.weight_decay : 0.f;
    float p[8];
    adam_update8_f32(param + i * 8, m + i * 8, v + i * 8, gf, h, wd, clip, p);
  }
}

__global__ void __launch_bounds__(256)
sumsq_f32_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < (n >> 2);
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// sum of squares of a bf16 arena into *out (fp32, must be zeroed by the caller)
__global__ void __launch_bounds__(256)
sumsq_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long n, float* __restrict__ out) {
  float acc = 0.f;
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 g = ld_nc_v4(reinterpret_cast<const uint4*>(x) + i);
    float2 a = unpack_bf16x2(g.x), b = unpack_bf16x2(g.y), c = unpack_bf16x2(g.z),
           d = unpack_bf16x2(g.w);
    acc += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x +
           d.y * d.y;
  }
  acc = warp_sum(acc);
  __shared__ float wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = wsum[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256)
cast_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                        long long n) {
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(src)[2 * i];
    float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y),
                         pack_bf16x2(b.z, b.w));
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}

}  // namespace

// n must be a multiple of 8 (the arena is padded). gnorm_sq may be null (no clipping).
extern "C" int mg_adam_flat(float* master, float* m, float* v, const void* grad, void* param_bf16,
                            long long n, long long decay_boundary, float lr, float beta1,
                            float beta2, float eps, float weight_decay, int step, float grad_scale,
                            const float* gnorm_sq, float max_norm, const int* step_ptr,
                            const float* lr_ptr, void* stream) {
  if (n % 8 != 0) return -1;
  AdamHyper h = make_adam_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                                decay_boundary, gnorm_sq, max_norm);
  h.step_ptr = step_ptr;
  h.lr_ptr = lr_ptr;
  long long nvec = n >> 3;
  int grid = static_cast<int>((nvec + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  adam_flat_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      master, m, v, reinterpret_cast<const __nv_bfloat16*>(grad),
      reinterpret_cast<__nv_bfloat16*>(param_bf16), n, h);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

extern "C" int mg_adam_flat_f32(float* param, float* m, float* v, const float* grad, long long n,
                                long long decay_boundary, float lr, float beta1, float beta2, float eps,
                                float weight_decay, int step, float grad_scale, const float* gnorm_sq, float max_norm,
                                const int* step_ptr, const float* lr_ptr, void* stream) {
  if (n % 8 != 0) return -1;
  AdamHyper h = make_adam_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale, decay_boundary, gnorm_sq, max_norm);
  h.step_ptr = step_ptr;
  h.lr_ptr = lr_ptr;
  int grid = static_cast<int>(((n >> 3) + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  adam_flat_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(param, m, v, grad, n, h);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

extern "C" int mg_sumsq_f32(const float* x, long long n, float* out, void* stream) {
  if (n % 4 != 0) return -1;
  int grid = static_cast<int>(((n >> 2) + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  sumsq_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

extern "C" int mg_sumsq_bf16(const void* x, long long n, float* out, void* stream) {
  if (n % 8 != 0) return -1;
  long long nvec = n >> 3;
  int grid = static_cast<int>((nvec + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  sumsq_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), n, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

extern "C" int mg_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n % 8 != 0) return -1;
  long long nvec = n >> 3;
  int grid = static_cast<int>((nvec + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  cast_f32_to_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// Bind this library's (statically linked) CUDA runtime to device `dev` on the calling thread. PyTorch runs
// backward on worker threads that may not have a driver context bound yet; driver-API entry points used
// here (cuTensorMapEncodeTiled) need one, and on rank > 0 it must be the right device.
extern "C" int mg_set_device(int dev) {
  static thread_local int current = -1;
  if (current == dev) return 0;
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) return -400 - static_cast<int>(e);
  cudaFree(nullptr);  // force context creation / binding on this thread
  current = dev;
  return 0;
}
