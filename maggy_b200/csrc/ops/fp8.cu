// Per-tensor e4m3 quantisation for the fp8 tcgen05 GEMM (K5 fp8 variant; SURVEY.md §2.4 row C10, and the
// "Llama-3 1B fp8" ablation config — the reference reaches fp8 nowhere, its stack predates it).
//
//   amax  = max |x|                       (one pass, bf16 reads, atomicMax on the non-negative float bits)
//   q     = sat_e4m3(x * 448 / amax)      (second pass, 8 elements per thread: 16 B in, 8 B out)
//   scale = amax / 448                    (device scalar consumed by the GEMM epilogue; no host round trip)
#include "../include/mg_ptx.cuh"

#include <cuda_fp8.h>

namespace {

using namespace mg;

__global__ void amax_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long n, float* __restrict__ amax) {
  float m = 0.f;
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 u = ld_nc_v4(x + i * 8);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(d.x), fabsf(d.y))));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = nvec * 8; i < n; ++i) m = fmaxf(m, fabsf(__bfloat162float(x[i])));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

__global__ void quant_e4m3_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, long long n,
                                  const float* __restrict__ amax, float* __restrict__ scale_out) {
  const float am = fmaxf(__ldg(amax), 1e-12f);
  const float s = 448.f / am;
  if (blockIdx.x == 0 && threadIdx.x == 0) *scale_out = am / 448.f;
  const long long nvec = n >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 u = ld_nc_v4(x + i * 8);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    const uint32_t p0 = __nv_cvt_float2_to_fp8x2(make_float2(a.x * s, a.y * s), __NV_SATFINITE, __NV_E4M3);
    const uint32_t p1 = __nv_cvt_float2_to_fp8x2(make_float2(b.x * s, b.y * s), __NV_SATFINITE, __NV_E4M3);
    const uint32_t p2 = __nv_cvt_float2_to_fp8x2(make_float2(c.x * s, c.y * s), __NV_SATFINITE, __NV_E4M3);
    const uint32_t p3 = __nv_cvt_float2_to_fp8x2(make_float2(d.x * s, d.y * s), __NV_SATFINITE, __NV_E4M3);
    *reinterpret_cast<uint2*>(q + i * 8) = make_uint2(p0 | (p1 << 16), p2 | (p3 << 16));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = nvec * 8; i < n; ++i)
      q[i] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(__bfloat162float(x[i]) * s, __NV_SATFINITE, __NV_E4M3));
}

}  // namespace

// q (e4m3 bytes) and *scale (fp32 device scalar, dequantisation factor) from x (bf16, 16-byte aligned).
// `amax_ws` is a 1-float device workspace.
extern "C" int mg_quant_e4m3(const void* x, void* q, long long n, float* amax_ws, float* scale, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(amax_ws, 0, sizeof(float), s);
  if (e != cudaSuccess) return -500 - static_cast<int>(e);
  long long blocks = ((n >> 3) + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (blocks < 1) blocks = 1;
  amax_bf16_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), n, amax_ws);
  quant_e4m3_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x),
                                                             static_cast<uint8_t*>(q), n, amax_ws, scale);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
