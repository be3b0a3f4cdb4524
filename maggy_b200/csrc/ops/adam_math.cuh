// Shared Adam/AdamW update body: used by the single-GPU flat kernel (ops/optim.cu) and by the
// reduce-scatter-fused kernel (comm/rs_adam.cu) so both paths are bit-identical per element.
#pragma once
#include "../include/mg_ptx.cuh"

namespace mg {

struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;
  float bias_c1, bias_c2;  // 1 - beta1^t, 1 - beta2^t
  float grad_scale;        // applied to the incoming gradient (1/world, 1/loss_scale, ...)
  long long decay_boundary;  // elements [0, decay_boundary) get weight decay
  const float* gnorm_sq;   // optional: global sum of squares of the (unscaled) gradient
  float max_norm;          // clip threshold when gnorm_sq != null
  // Optional device-resident overrides so a captured CUDA graph can be replayed with a moving step
  // count / learning-rate schedule: when non-null they replace `bias_c*` / `lr` at kernel start.
  const int* step_ptr;
  const float* lr_ptr;
};

inline AdamHyper make_adam_hyper(float lr, float beta1, float beta2, float eps, float weight_decay,
                                 int step, float grad_scale, long long decay_boundary,
                                 const float* gnorm_sq, float max_norm) {
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  double b1 = 1.0, b2 = 1.0;
  for (int i = 0; i < step; ++i) { b1 *= beta1; b2 *= beta2; }
  h.bias_c1 = static_cast<float>(1.0 - b1);
  h.bias_c2 = static_cast<float>(1.0 - b2);
  h.grad_scale = grad_scale;
  h.decay_boundary = decay_boundary;
  h.gnorm_sq = gnorm_sq;
  h.max_norm = max_norm;
  h.step_ptr = nullptr;
  h.lr_ptr = nullptr;
  return h;
}

MG_DEVICE AdamHyper adam_resolve(AdamHyper h) {
  if (h.step_ptr != nullptr) {
    float t = static_cast<float>(*h.step_ptr);
    h.bias_c1 = 1.0f - powf(h.beta1, t);
    h.bias_c2 = 1.0f - powf(h.beta2, t);
  }
  if (h.lr_ptr != nullptr) h.lr = *h.lr_ptr;
  return h;
}

MG_DEVICE float adam_clip_coef(const AdamHyper& h) {
  if (h.gnorm_sq == nullptr) return 1.0f;
  float norm = sqrtf(*h.gnorm_sq) * h.grad_scale;
  return norm > h.max_norm ? h.max_norm / (norm + 1e-6f) : 1.0f;
}

// Update 8 consecutive elements. master/m/v are fp32 (read-modify-write), gf is the fp32 gradient,
// `p` receives the 8 updated parameters in fp32.
MG_DEVICE void adam_update8_f32(float* master, float* m, float* v, const float (&gf)[8],
                                const AdamHyper& h, float wd, float clip, float (&p)[8]) {
  float4 p0 = reinterpret_cast<float4*>(master)[0], p1 = reinterpret_cast<float4*>(master)[1];
  float4 m0 = reinterpret_cast<float4*>(m)[0], m1 = reinterpret_cast<float4*>(m)[1];
  float4 v0 = reinterpret_cast<float4*>(v)[0], v1 = reinterpret_cast<float4*>(v)[1];
  p[0] = p0.x; p[1] = p0.y; p[2] = p0.z; p[3] = p0.w; p[4] = p1.x; p[5] = p1.y; p[6] = p1.z; p[7] = p1.w;
  float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
  float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  const float gs = h.grad_scale * clip;
  const float inv_c1 = 1.0f / h.bias_c1;
  const float inv_sqrt_c2 = rsqrtf(h.bias_c2);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float g = gf[t] * gs;
    mm[t] = h.beta1 * mm[t] + (1.0f - h.beta1) * g;
    vv[t] = h.beta2 * vv[t] + (1.0f - h.beta2) * g * g;
    float mhat = mm[t] * inv_c1;
    float denom = sqrtf(vv[t]) * inv_sqrt_c2 + h.eps;
    p[t] = p[t] - h.lr * (mhat / denom + wd * p[t]);  // decoupled weight decay (AdamW)
  }
  reinterpret_cast<float4*>(master)[0] = make_float4(p[0], p[1], p[2], p[3]);
  reinterpret_cast<float4*>(master)[1] = make_float4(p[4], p[5], p[6], p[7]);
  reinterpret_cast<float4*>(m)[0] = make_float4(mm[0], mm[1], mm[2], mm[3]);
  reinterpret_cast<float4*>(m)[1] = make_float4(mm[4], mm[5], mm[6], mm[7]);
  reinterpret_cast<float4*>(v)[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  reinterpret_cast<float4*>(v)[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
}

// Same, `out` receives the 8 updated parameters packed as bf16 (the compute copy of a bf16 arena).
MG_DEVICE void adam_update8(float* master, float* m, float* v, const float (&gf)[8],
                            const AdamHyper& h, float wd, float clip, uint4& out) {
  float p[8];
  adam_update8_f32(master, m, v, gf, h, wd, clip, p);
  out = make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]),
                   pack_bf16x2(p[6], p[7]));
}

}  // namespace mg
