// Reduce-scatter (sum) over NVLink peer memory — the backward half of the sharded (ZeRO-3) linear layer.
//
//   out[i] = scale * sum_r in_r[rank * shard + i],   i in [0, shard)
//
// Every rank holds its full-size contribution (e.g. the wgrad dW = dY^T X of a row-sharded weight) in a symmetric,
// peer-mapped buffer; rank r reduces ONLY the rows it owns, reading its slice of every peer's buffer directly
// (`multimem.ld_reduce` on the NVSwitch multicast address when one is bound, else `world` peer loads issued before
// first use). This replaces `dist.reduce_scatter_tensor` (NCCL) in parallel/fsdp.py; the reference reaches the same
// step through fairscale FSDP's reduce-scatter (maggy/core/patching/modules.py:75-95). Cross-GPU ordering is the
// two-flag protocol of csrc/comm/rs_adam.cu: "my buffer is complete" before anyone reads, "I finished reading"
// before the kernel retires, so plain stream order protects the buffer's next writer.
#include "../include/mg_ptx.cuh"

namespace {
using namespace mg;

constexpr int kMaxWorld = 8;

struct RsSumArgs {
  const __nv_bfloat16* in_peer[kMaxWorld];
  uint32_t* flag_peer[kMaxWorld];   // [0..8) begin epochs, [8..16) end epochs
  const __nv_bfloat16* in_mc;       // multicast address of the input buffer (or null)
  __nv_bfloat16* out;               // local [shard]
  uint32_t* local_sync;             // [0] begin broadcast, [1] finished-CTA counter, [2] launches completed
  long long shard_elems;            // multiple of 8
  int rank, world;
  uint32_t epoch;                   // 0 = local_sync[2] + 1
  float scale;
};

template <class Pred>
MG_DEVICE void spin_until(Pred ready, const char* what) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!ready()) {
    if ((++spins & 0x3FF) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) {
        printf("maggy_b200 rs_sum: timed out waiting for %s (block %d)\n", what, static_cast<int>(blockIdx.x));
        __trap();
      }
    }
  }
}

MG_DEVICE uint4 multimem_ld_reduce_bf16x8(const void* mc_addr) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}

template <bool USE_MC>
__global__ void __launch_bounds__(256)
rs_sum_kernel(const RsSumArgs a_in) {
  RsSumArgs a = a_in;
  if (a.epoch == 0) a.epoch = *reinterpret_cast<volatile uint32_t*>(a.local_sync + 2) + 1;
  // ---- begin: every rank's buffer is complete
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) {
      const int peer = threadIdx.x;
      __threadfence_system();
      st_release_sys_u32(a.flag_peer[peer] + a.rank, a.epoch);
      spin_until([&] { return ld_acquire_sys_u32(a.flag_peer[a.rank] + peer) >= a.epoch; }, "peer buffers");
    }
    __syncthreads();
    if (threadIdx.x == 0)
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.local_sync), "r"(a.epoch) : "memory");
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
  const long long nvec = a.shard_elems >> 3;
  const long long base = static_cast<long long>(a.rank) * a.shard_elems;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = base + i * 8;
    float gf[8];
    if constexpr (USE_MC) {
      const uint4 g = multimem_ld_reduce_bf16x8(a.in_mc + e);
      const float2 f0 = unpack_bf16x2(g.x), f1 = unpack_bf16x2(g.y), f2 = unpack_bf16x2(g.z), f3 = unpack_bf16x2(g.w);
      gf[0] = f0.x; gf[1] = f0.y; gf[2] = f1.x; gf[3] = f1.y; gf[4] = f2.x; gf[5] = f2.y; gf[6] = f3.x; gf[7] = f3.y;
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) gf[t] = 0.f;
      uint4 g[kMaxWorld];
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < a.world) {
          int src = a.rank + r;
          if (src >= a.world) src -= a.world;
          g[r] = ld_relaxed_sys_v4(a.in_peer[src] + e);
        }
      }
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < a.world) {
          const float2 f0 = unpack_bf16x2(g[r].x), f1 = unpack_bf16x2(g[r].y), f2 = unpack_bf16x2(g[r].z),
                       f3 = unpack_bf16x2(g[r].w);
          gf[0] += f0.x; gf[1] += f0.y; gf[2] += f1.x; gf[3] += f1.y;
          gf[4] += f2.x; gf[5] += f2.y; gf[6] += f3.x; gf[7] += f3.y;
        }
      }
    }
    const float s = a.scale;
    *reinterpret_cast<uint4*>(a.out + i * 8) =
        make_uint4(pack_bf16x2(gf[0] * s, gf[1] * s), pack_bf16x2(gf[2] * s, gf[3] * s),
                   pack_bf16x2(gf[4] * s, gf[5] * s), pack_bf16x2(gf[6] * s, gf[7] * s));
  }
  // ---- end: nobody reads my buffer any more once this kernel retires
  __syncthreads();
  __shared__ uint32_t is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    is_last = (atomicAdd(a.local_sync + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) {
      a.local_sync[1] = 0;
      a.local_sync[2] = a.epoch;
    }
    if (threadIdx.x < a.world) {
      const int peer = threadIdx.x;
      __threadfence_system();
      st_release_sys_u32(a.flag_peer[peer] + kMaxWorld + a.rank, a.epoch);
      spin_until([&] { return ld_acquire_sys_u32(a.flag_peer[a.rank] + kMaxWorld + peer) >= a.epoch; },
                 "peers to finish reading");
    }
    __syncthreads();
  }
}

// K2 — one-shot broadcast of a buffer from `root` to every rank over NVLink (DDP's constructor broadcast, SURVEY C3;
// reference maggy/core/patching/modules.py:62-63): the root's CTAs read their own copy once and store it to all peers
// — ONE multimem.st per 16 bytes when a multicast address is bound (the switch replicates), else `world - 1` peer
// stores; the other ranks' kernels only wait for the root's end flag. Flag words [16..24) / [24..32) of the same
// symmetric flag buffer are used so that broadcasts and reductions can share it.
struct BcastArgs {
  uint8_t* buf_peer[kMaxWorld];
  uint32_t* flag_peer[kMaxWorld];
  uint8_t* buf_mc;
  uint32_t* local_sync;   // [1] finished-CTA counter
  long long bytes;        // multiple of 16
  int rank, world, root;
  uint32_t epoch;
};

__global__ void __launch_bounds__(256)
peer_bcast_kernel(const BcastArgs a) {
  if (a.rank == a.root) {
    const long long nvec = a.bytes >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(a.buf_peer[a.root]);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const uint4 v = src[i];
      if (a.buf_mc != nullptr) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.buf_mc + i * 16), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
      } else {
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
          if (r < a.world && r != a.root) st_na_v4(a.buf_peer[r] + i * 16, v);
      }
    }
    __syncthreads();
    __shared__ uint32_t is_last;
    if (threadIdx.x == 0) {
      __threadfence_system();
      is_last = (atomicAdd(a.local_sync + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
      if (threadIdx.x == 0) a.local_sync[1] = 0;
      if (threadIdx.x < a.world) {
        __threadfence_system();
        st_release_sys_u32(a.flag_peer[threadIdx.x] + 16 + a.root, a.epoch);
      }
    }
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    spin_until([&] { return ld_acquire_sys_u32(a.flag_peer[a.rank] + 16 + a.root) >= a.epoch; }, "broadcast root");
  }
}

}  // namespace

// buf_ptrs / flag_ptrs: `world` peer-mapped pointers (symmetric buffers; flags >= 32 zeroed uint32); buf_mc: multicast
// address of the buffer or 0. `epoch` must increase with every call on every rank.
extern "C" int mg_peer_bcast(const unsigned long long* buf_ptrs, const unsigned long long* flag_ptrs,
                             unsigned long long buf_mc, void* local_sync, long long bytes, int rank, int world, int root,
                             unsigned int epoch, void* stream) {
  if (world < 1 || world > kMaxWorld || bytes % 16 != 0 || root < 0 || root >= world) return -1;
  BcastArgs a;
  for (int r = 0; r < kMaxWorld; ++r) {
    a.buf_peer[r] = r < world ? reinterpret_cast<uint8_t*>(buf_ptrs[r]) : nullptr;
    a.flag_peer[r] = r < world ? reinterpret_cast<uint32_t*>(flag_ptrs[r]) : nullptr;
  }
  a.buf_mc = reinterpret_cast<uint8_t*>(buf_mc);
  a.local_sync = static_cast<uint32_t*>(local_sync);
  a.bytes = bytes; a.rank = rank; a.world = world; a.root = root; a.epoch = epoch;
  int grid = 1;
  if (rank == root) {
    grid = static_cast<int>(((bytes >> 4) + 255) / 256);
    if (grid > 148 * 2) grid = 148 * 2;
    if (grid < 1) grid = 1;
  }
  peer_bcast_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// in_ptrs / flag_ptrs: host arrays of `world` peer-mapped device pointers (full buffers of world * shard_elems bf16,
// flag buffers of >= 32 zeroed uint32); local_sync: 4 zeroed uint32; in_mc: multicast address of the input or 0.
extern "C" int mg_rs_sum(const unsigned long long* in_ptrs, const unsigned long long* flag_ptrs,
                         unsigned long long in_mc, void* out, void* local_sync, long long shard_elems, int rank,
                         int world, unsigned int epoch, float scale, int num_ctas, void* stream) {
  if (world < 1 || world > kMaxWorld || shard_elems % 8 != 0) return -1;
  RsSumArgs a;
  for (int r = 0; r < kMaxWorld; ++r) {
    a.in_peer[r] = r < world ? reinterpret_cast<const __nv_bfloat16*>(in_ptrs[r]) : nullptr;
    a.flag_peer[r] = r < world ? reinterpret_cast<uint32_t*>(flag_ptrs[r]) : nullptr;
  }
  a.in_mc = reinterpret_cast<const __nv_bfloat16*>(in_mc);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.local_sync = static_cast<uint32_t*>(local_sync);
  a.shard_elems = shard_elems;
  a.rank = rank; a.world = world; a.epoch = epoch; a.scale = scale;
  const long long nvec = shard_elems >> 3;
  int grid = static_cast<int>((nvec + 255) / 256);
  const int cap = num_ctas > 0 ? num_ctas : 148 * 2;   // every CTA must be co-resident (the begin barrier spins)
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in_mc != 0) rs_sum_kernel<true><<<grid, 256, 0, s>>>(a);
  else rs_sum_kernel<false><<<grid, 256, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}
