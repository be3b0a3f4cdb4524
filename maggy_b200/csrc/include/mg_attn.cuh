// maggy_b200 — pieces shared by the attention kernels (csrc/ops/attention.cu: first generation, kept as the
// reference implementation; csrc/ops/attention_v2.cu: the tuned generation).
#pragma once
#include "mg_ptx.cuh"

#include <mutex>

namespace mg {
namespace attn {

constexpr int kD = 64;            // head dim
constexpr int kBQ = 128;          // query rows per tile
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

MG_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (max relative error 7.7e-5, far below bf16's 2^-9), n added into the exponent field.
// Valid for x <= 127; inputs below -125 are clamped (the result is then ~2^-125, i.e. zero for softmax purposes).
MG_DEVICE float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;              // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = 0.05508868396282196f;
  p = fmaf(p, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}


MG_DEVICE uint4 pack8(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                    pack_bf16x2(v[6], v[7]));
}

// 32 lanes x 16 columns of fp32 (thread t of the warp receives row lane-base + t)
MG_DEVICE void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

MG_DEVICE void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// shared -> global element-wise ADD of a tile (fp32 map): the reduction happens in L2, no SM atomics
MG_DEVICE void tma_reduce_add_3d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled encode_fn() {
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

// [B, T, C] bf16 (row stride `ld` elements) as a 3-D map {C, T, B}; box {64, box_rows, 1}; 128B swizzle
inline int make_map_btc(CUtensorMap* map, const void* ptr, int B, int T, int C, long long ld, int box_rows = 128) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * T};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

// [B, T, C] fp32 as a 3-D map {C, T, B}; box {32 floats = 128 B, box_rows, 1}; 128B swizzle (reduce-add target)
inline int make_map_btc_f32(CUtensorMap* map, const void* ptr, int B, int T, int C, int box_rows = 128) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(C) * 4 * T};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -200 - static_cast<int>(r);
}

// defined in csrc/ops/attention.cu
void launch_delta(const void* o, const void* d_o, float* delta, int B, int T, int H, cudaStream_t s);
void launch_dq_convert(const float* dq_acc, void* dqkv, long long rows, int Cq, int ld, float scale, cudaStream_t s);
int generation();  // MAGGY_B200_ATTN_GEN (default 2)

}  // namespace attn
}  // namespace mg
