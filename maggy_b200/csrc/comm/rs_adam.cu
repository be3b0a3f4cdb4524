// K1 + K2 — gradient reduce-scatter fused with scale/cast + sharded Adam(W) + parameter all-gather,
// in ONE kernel over NVLink 5 / NVSwitch peer memory (sm_100a).
//
// This is the B200-native replacement for the three library steps the reference strings together:
//   DDP bucketed all-reduce        (maggy/core/patching/modules.py:44-63)
//   ZeroRedundancyOptimizer.step() (maggy/core/patching/optim.py:28-45; local shard step)
//   ... followed by its per-owner parameter broadcast.
//
// Every rank holds the full bf16 gradient arena and the full bf16 parameter arena in symmetric
// (peer-mapped) memory. Rank r owns elements [r*shard, (r+1)*shard) of both and keeps fp32
// master / m / v only for that shard (ZeRO-1/2 memory footprint). Per 16-byte vector of its shard
// a thread:
//   1. reduces the gradient across ranks — either `multimem.ld_reduce.add.acc::f32` on the NVSwitch
//      multicast address (in-switch reduction, NVLS) or `world` direct peer loads summed in fp32;
//   2. applies grad_scale (1/world, loss scale, clip) and the Adam update on the fp32 shard state;
//   3. stores the updated bf16 parameters to EVERY rank — `multimem.st` (one store, switch
//      replicates) or `world` peer stores.
// Cross-GPU ordering is two flag barriers on symmetric signal words (release/acquire at .sys scope):
// "all ranks' gradients are complete" before step 1 and "all ranks finished writing parameters"
// before the kernel retires, so plain stream order makes the next forward safe. No NCCL call, no
// separate collective launch, no host involvement.
#include "../include/mg_ptx.cuh"
#include "../ops/adam_math.cuh"

namespace {
using namespace mg;

constexpr int kMaxWorld = 8;  // one NVSwitch domain (single 8xB200 node)

struct RsAdamArgs {
  const __nv_bfloat16* grad_peer[kMaxWorld];  // each rank's gradient arena (peer-mapped)
  __nv_bfloat16* param_peer[kMaxWorld];       // each rank's parameter arena (peer-mapped)
  uint32_t* flag_peer[kMaxWorld];             // each rank's signal words: [0..world) begin, [8..8+world) end
  const __nv_bfloat16* grad_mc;               // multicast address of the gradient arena (or null)
  __nv_bfloat16* param_mc;                    // multicast address of the parameter arena (or null)
  float* master;                              // fp32 state of THIS rank's shard
  float* m;
  float* v;
  uint32_t* local_sync;  // [0] begin-epoch broadcast to all CTAs, [1] finished-CTA counter,
                         // [2] launches completed so far (device-side epoch source for CUDA graphs)
  long long shard_elems;  // multiple of 8
  int rank, world;
  uint32_t epoch;  // strictly increasing per launch (>= 1); 0 = take local_sync[2] + 1 on device
  AdamHyper h;
};

MG_DEVICE uint4 multimem_ld_reduce_bf16x8(const void* mc_addr) {
  uint4 r;
  asm volatile(
      "multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(mc_addr)
      : "memory");
  return r;
}

MG_DEVICE void multimem_st_v4(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// Bounded spin: a rank that never arrives (crashed peer, early-stopped gang member) must surface as a
// launch failure on the waiting ranks, not as a GPU that spins forever.
template <class Pred>
MG_DEVICE void spin_until(Pred ready, const char* what) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!ready()) {
    if ((++spins & 0x3FF) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) {
        printf("maggy_b200 rs_adam: timed out waiting for %s (block %d)\n", what, static_cast<int>(blockIdx.x));
        __trap();
      }
    }
  }
}

// All ranks rendezvous: block 0 signals every peer, waits for every peer, then releases the local CTAs.
MG_DEVICE void begin_barrier(const RsAdamArgs& a) {
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) {
      const int peer = threadIdx.x;
      __threadfence_system();
      st_release_sys_u32(a.flag_peer[peer] + a.rank, a.epoch);
      spin_until([&] { return ld_acquire_sys_u32(a.flag_peer[a.rank] + peer) >= a.epoch; }, "peer gradients");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.local_sync), "r"(a.epoch) : "memory");
    }
  } else {
    if (threadIdx.x == 0) {
      spin_until([&] {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.local_sync) : "memory");
        return v >= a.epoch;
      }, "local begin barrier");
    }
    __syncthreads();
  }
}

// The last CTA of this rank to finish tells every peer, then waits until every peer has finished too.
MG_DEVICE void end_barrier(const RsAdamArgs& a) {
  __syncthreads();
  __shared__ uint32_t is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    uint32_t prev = atomicAdd(a.local_sync + 1, 1u);
    is_last = (prev == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) {
      a.local_sync[1] = 0;  // re-arm for the next launch
      a.local_sync[2] = a.epoch;
    }
    if (threadIdx.x < a.world) {
      const int peer = threadIdx.x;
      __threadfence_system();
      st_release_sys_u32(a.flag_peer[peer] + kMaxWorld + a.rank, a.epoch);
      spin_until([&] { return ld_acquire_sys_u32(a.flag_peer[a.rank] + kMaxWorld + peer) >= a.epoch; },
                 "peers to finish writing parameters");
    }
    __syncthreads();
  }
}

MG_DEVICE float4 multimem_ld_reduce_f32x4(const void* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}

// fp32 arenas (generic torch modules: fp32 parameters, fp32 autograd gradients). The pointers in RsAdamArgs are
// reinterpreted as float arrays; `master` is this rank's shard of its OWN parameter arena (the arena is the master).
template <bool USE_MC>
__global__ void __launch_bounds__(256)
rs_adam_f32_kernel(const RsAdamArgs a_in) {
  RsAdamArgs a = a_in;
  if (a.epoch == 0) a.epoch = *reinterpret_cast<volatile uint32_t*>(a.local_sync + 2) + 1;
  a.h = adam_resolve(a.h);
  begin_barrier(a);
  const float clip = adam_clip_coef(a.h);
  const long long nvec = a.shard_elems >> 3;
  const long long base = static_cast<long long>(a.rank) * a.shard_elems;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = base + i * 8;
    float gf[8];
    if constexpr (USE_MC) {
      const float* mc = reinterpret_cast<const float*>(a.grad_mc);
      const float4 g0 = multimem_ld_reduce_f32x4(mc + e), g1 = multimem_ld_reduce_f32x4(mc + e + 4);
      gf[0] = g0.x; gf[1] = g0.y; gf[2] = g0.z; gf[3] = g0.w; gf[4] = g1.x; gf[5] = g1.y; gf[6] = g1.z; gf[7] = g1.w;
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) gf[t] = 0.f;
      uint4 g[kMaxWorld][2];
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < a.world) {
          int src = a.rank + r;
          if (src >= a.world) src -= a.world;
          const float* gp = reinterpret_cast<const float*>(a.grad_peer[src]) + e;
          g[r][0] = ld_relaxed_sys_v4(gp);
          g[r][1] = ld_relaxed_sys_v4(gp + 4);
        }
      }
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < a.world) {
          gf[0] += __uint_as_float(g[r][0].x); gf[1] += __uint_as_float(g[r][0].y);
          gf[2] += __uint_as_float(g[r][0].z); gf[3] += __uint_as_float(g[r][0].w);
          gf[4] += __uint_as_float(g[r][1].x); gf[5] += __uint_as_float(g[r][1].y);
          gf[6] += __uint_as_float(g[r][1].z); gf[7] += __uint_as_float(g[r][1].w);
        }
      }
    }
    const float wd = (e < a.h.decay_boundary) ? a.h.weight_decay : 0.f;
    float p[8];
    adam_update8_f32(a.master + i * 8, a.m + i * 8, a.v + i * 8, gf, a.h, wd, clip, p);
    const uint4 o0 = make_uint4(__float_as_uint(p[0]), __float_as_uint(p[1]), __float_as_uint(p[2]), __float_as_uint(p[3]));
    const uint4 o1 = make_uint4(__float_as_uint(p[4]), __float_as_uint(p[5]), __float_as_uint(p[6]), __float_as_uint(p[7]));
    if constexpr (USE_MC) {
      float* mc = reinterpret_cast<float*>(a.param_mc);
      multimem_st_v4(mc + e, o0);
      multimem_st_v4(mc + e + 4, o1);
    } else {
#pragma unroll
      for (int r = 1; r < kMaxWorld; ++r) {   // r = 0 is this rank: adam_update8_f32 already wrote its arena
        if (r < a.world) {
          int dst = a.rank + r;
          if (dst >= a.world) dst -= a.world;
          float* pp = reinterpret_cast<float*>(a.param_peer[dst]) + e;
          st_na_v4(pp, o0);
          st_na_v4(pp + 4, o1);
        }
      }
    }
  }
  end_barrier(a);
}

template <bool USE_MC>
__global__ void __launch_bounds__(256)
rs_adam_kernel(const RsAdamArgs a_in) {
  RsAdamArgs a = a_in;
  // every CTA reads the launch counter before the last CTA of this launch can bump it
  if (a.epoch == 0) a.epoch = *reinterpret_cast<volatile uint32_t*>(a.local_sync + 2) + 1;
  a.h = adam_resolve(a.h);
  begin_barrier(a);
  const float clip = adam_clip_coef(a.h);
  const long long nvec = a.shard_elems >> 3;
  const long long base = static_cast<long long>(a.rank) * a.shard_elems;  // element offset of my shard
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = base + i * 8;
    float gf[8];
    if constexpr (USE_MC) {
      uint4 g = multimem_ld_reduce_bf16x8(a.grad_mc + e);
      float2 f0 = unpack_bf16x2(g.x), f1 = unpack_bf16x2(g.y), f2 = unpack_bf16x2(g.z),
             f3 = unpack_bf16x2(g.w);
      gf[0] = f0.x; gf[1] = f0.y; gf[2] = f1.x; gf[3] = f1.y;
      gf[4] = f2.x; gf[5] = f2.y; gf[6] = f3.x; gf[7] = f3.y;
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) gf[t] = 0.f;
      uint4 g[kMaxWorld];
      // issue every peer load before consuming any (NVLink latency ~2 us: keep `world` loads in flight);
      // start at my own rank so the 8 ranks do not all hit the same peer at the same time
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < a.world) {
          int src = a.rank + r;
          if (src >= a.world) src -= a.world;
          g[r] = ld_relaxed_sys_v4(a.grad_peer[src] + e);
        }
      }
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < a.world) {
          float2 f0 = unpack_bf16x2(g[r].x), f1 = unpack_bf16x2(g[r].y), f2 = unpack_bf16x2(g[r].z),
                 f3 = unpack_bf16x2(g[r].w);
          gf[0] += f0.x; gf[1] += f0.y; gf[2] += f1.x; gf[3] += f1.y;
          gf[4] += f2.x; gf[5] += f2.y; gf[6] += f3.x; gf[7] += f3.y;
        }
      }
    }
    const float wd = (e < a.h.decay_boundary) ? a.h.weight_decay : 0.f;
    uint4 out;
    adam_update8(a.master + i * 8, a.m + i * 8, a.v + i * 8, gf, a.h, wd, clip, out);
    if constexpr (USE_MC) {
      multimem_st_v4(a.param_mc + e, out);
    } else {
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < a.world) {
          int dst = a.rank + r;
          if (dst >= a.world) dst -= a.world;
          st_na_v4(a.param_peer[dst] + e, out);
        }
      }
    }
  }
  end_barrier(a);
}

}  // namespace

// grad_ptrs / param_ptrs / flag_ptrs: host arrays of `world` device pointers (peer-mapped).
// flag buffers must hold >= 32 zero-initialised uint32 words; local_sync >= 2 zeroed uint32 words.
// epoch must increase by one per call on every rank (start at 1), or be 0 on every rank to use the
// device-side launch counter (CUDA-graph replay safe).
static int rs_adam_launch(bool fp32, const unsigned long long* grad_ptrs, const unsigned long long* param_ptrs,
                          const unsigned long long* flag_ptrs, unsigned long long grad_mc,
                          unsigned long long param_mc, float* master, float* m, float* v,
                          void* local_sync, long long shard_elems, int rank, int world,
                          unsigned int epoch, long long decay_boundary, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int step, float grad_scale,
                          const float* gnorm_sq, float max_norm, const int* step_ptr,
                          const float* lr_ptr, int num_ctas, void* stream);

// fp32 arenas: grad / param pointers address float arrays, `master` = this rank's shard of its own parameter arena
extern "C" int mg_rs_adam_f32(const unsigned long long* grad_ptrs, const unsigned long long* param_ptrs,
                              const unsigned long long* flag_ptrs, unsigned long long grad_mc,
                              unsigned long long param_mc, float* master, float* m, float* v,
                              void* local_sync, long long shard_elems, int rank, int world,
                              unsigned int epoch, long long decay_boundary, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, float grad_scale,
                              const float* gnorm_sq, float max_norm, const int* step_ptr,
                              const float* lr_ptr, int num_ctas, void* stream) {
  return rs_adam_launch(true, grad_ptrs, param_ptrs, flag_ptrs, grad_mc, param_mc, master, m, v, local_sync, shard_elems,
                        rank, world, epoch, decay_boundary, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                        gnorm_sq, max_norm, step_ptr, lr_ptr, num_ctas, stream);
}

extern "C" int mg_rs_adam(const unsigned long long* grad_ptrs, const unsigned long long* param_ptrs,
                          const unsigned long long* flag_ptrs, unsigned long long grad_mc,
                          unsigned long long param_mc, float* master, float* m, float* v,
                          void* local_sync, long long shard_elems, int rank, int world,
                          unsigned int epoch, long long decay_boundary, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int step, float grad_scale,
                          const float* gnorm_sq, float max_norm, const int* step_ptr,
                          const float* lr_ptr, int num_ctas, void* stream) {
  return rs_adam_launch(false, grad_ptrs, param_ptrs, flag_ptrs, grad_mc, param_mc, master, m, v, local_sync,
                        shard_elems, rank, world, epoch, decay_boundary, lr, beta1, beta2, eps, weight_decay, step,
                        grad_scale, gnorm_sq, max_norm, step_ptr, lr_ptr, num_ctas, stream);
}

static int rs_adam_launch(bool fp32, const unsigned long long* grad_ptrs, const unsigned long long* param_ptrs,
                          const unsigned long long* flag_ptrs, unsigned long long grad_mc,
                          unsigned long long param_mc, float* master, float* m, float* v,
                          void* local_sync, long long shard_elems, int rank, int world,
                          unsigned int epoch, long long decay_boundary, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int step, float grad_scale,
                          const float* gnorm_sq, float max_norm, const int* step_ptr,
                          const float* lr_ptr, int num_ctas, void* stream) {
  if (world < 1 || world > kMaxWorld || shard_elems % 8 != 0) return -1;
  RsAdamArgs a;
  for (int r = 0; r < kMaxWorld; ++r) {
    a.grad_peer[r] = r < world ? reinterpret_cast<const __nv_bfloat16*>(grad_ptrs[r]) : nullptr;
    a.param_peer[r] = r < world ? reinterpret_cast<__nv_bfloat16*>(param_ptrs[r]) : nullptr;
    a.flag_peer[r] = r < world ? reinterpret_cast<uint32_t*>(flag_ptrs[r]) : nullptr;
  }
  a.grad_mc = reinterpret_cast<const __nv_bfloat16*>(grad_mc);
  a.param_mc = reinterpret_cast<__nv_bfloat16*>(param_mc);
  a.master = master; a.m = m; a.v = v;
  a.local_sync = reinterpret_cast<uint32_t*>(local_sync);
  a.shard_elems = shard_elems;
  a.rank = rank; a.world = world; a.epoch = epoch;
  a.h = make_adam_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale, decay_boundary,
                        gnorm_sq, max_norm);
  a.h.step_ptr = step_ptr;
  a.h.lr_ptr = lr_ptr;
  long long nvec = shard_elems >> 3;
  int grid = static_cast<int>((nvec + 255) / 256);
  // every CTA must be co-resident (the begin barrier spins): cap at 148 SMs x 2
  int cap = num_ctas > 0 ? num_ctas : 148 * 2;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (fp32) {
    if (grad_mc != 0 && param_mc != 0) rs_adam_f32_kernel<true><<<grid, 256, 0, s>>>(a);
    else rs_adam_f32_kernel<false><<<grid, 256, 0, s>>>(a);
  } else if (grad_mc != 0 && param_mc != 0)
    rs_adam_kernel<true><<<grid, 256, 0, s>>>(a);
  else
    rs_adam_kernel<false><<<grid, 256, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
