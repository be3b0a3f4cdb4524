// K8 / K9 — device side of the control plane (sm_100a).
//
// K8 `report_store`: `reporter.broadcast(metric_tensor, step)` becomes ONE tiny kernel on the training
//   stream that reads the metric where it already lives (device memory), fills a 64-byte MgRecord in
//   the worker's *device ring* (shared segment, cudaHostRegister-mapped) and publishes it with
//   st.release.sys — no .item(), no host sync, no RPC. The reference ships the metric from a host
//   heartbeat thread once per second (maggy/core/rpc.py:716-737) after the user's loop has already
//   synchronised to read the loss.
//
// K9 `sched_kernel`: a persistent single-CTA kernel that polls every worker's rings, evaluates the
//   device-decidable early-stop policy (median stopping rule, reference
//   maggy/earlystop/medianrule.py:28-60 + the gating in optimization_driver.py:447-471) and writes
//   the per-worker STOP word itself. It only *observes* the rings (private cursors; the host pump
//   stays the consumer that advances `tail`), so host and device policy engines can both run — STOP is
//   idempotent — and the decision latency is one poll iteration instead of two 1 s heartbeats.
//
// Device mailbox (round 2): the scheduler's GPU owns a peer-visible allocation (MgMailboxCtl + one record ring per
// worker, mg_ring.h) that every worker maps through CUDA IPC. K8 then ALSO stores its record there — a 64-byte
// NVLink store — and reads the STOP word from it; K9 polls its own HBM instead of host memory over PCIe and writes
// STOP into both places. The host-mapped ring keeps receiving the same records: it is what the Python driver's
// bookkeeping (trial.json metric history) consumes.
//
// Clocks: device stamps are %globaltimer + a per-process offset measured against CLOCK_MONOTONIC at start-up
// (mg_gt_offset_ns), so every t_ns / stop_t_ns in the segment lives in ONE clock domain; K8 no longer touches the
// worker's heartbeat word (the host heartbeat thread owns it).
#include "../include/mg_ptx.cuh"
#include "../include/mg_ring.h"
#include "../include/mg_asha.h"

#include <string.h>
#include <time.h>

#include <cuda_fp16.h>

namespace {
using namespace mg;

struct DevProducer {  // lives in device memory, one per worker process (32 zeroed bytes)
  unsigned long long head;       // records produced so far on the host-mapped device ring
  unsigned long long dropped;
  unsigned int last_stop;        // 1 if the last published trial had its STOP word set at publish time
  unsigned int pad;
  unsigned long long mbox_head;  // records produced so far on the mailbox ring
};

__device__ __forceinline__ double load_metric(const void* p, int dtype) {
  switch (dtype) {
    case 0: return static_cast<double>(*reinterpret_cast<const float*>(p));
    case 1: return static_cast<double>(__bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p)));
    case 2: return static_cast<double>(__half2float(*reinterpret_cast<const __half*>(p)));
    default: return *reinterpret_cast<const double*>(p);
  }
}

__global__ void report_store_kernel(uint8_t* base, unsigned long long off_ring,
                                    unsigned long long off_ctl, unsigned int slots,
                                    const void* metric, int dtype, long long step,
                                    unsigned long long trial_id, unsigned int worker,
                                    DevProducer* st, uint8_t* mailbox, unsigned int mbox_workers,
                                    unsigned int mbox_slots, long long gt_offset_ns) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  MgRing* ring = reinterpret_cast<MgRing*>(base + off_ring);
  MgRecord* sl = reinterpret_cast<MgRecord*>(base + off_ring + sizeof(MgRing));
  MgWorkerCtl* ctl = reinterpret_cast<MgWorkerCtl*>(base + off_ctl);
  const double value = load_metric(metric, dtype);
  const unsigned long long now = globaltimer_ns() + static_cast<unsigned long long>(gt_offset_ns);
  bool stop = false;
  if (mailbox) {
    // peer-visible ring in the scheduler GPU's memory: never blocks (K9 keeps a private cursor and resynchronises
    // when lapped), written with the same "seq last, release" protocol
    MgMailboxCtl* mc = reinterpret_cast<MgMailboxCtl*>(mailbox) + worker;
    MgRecord* ms = reinterpret_cast<MgRecord*>(mailbox + static_cast<size_t>(mbox_workers) * sizeof(MgMailboxCtl) +
                                              static_cast<size_t>(worker) * mbox_slots * sizeof(MgRecord));
    const unsigned long long mh = st->mbox_head;
    MgRecord* s = &ms[mh & (mbox_slots - 1)];
    s->type = MG_REC_METRIC;
    s->worker = static_cast<uint16_t>(worker);
    s->aux = 1;
    s->trial_id = trial_id;
    s->step = step;
    s->value = value;
    s->t_ns = now;
    s->reserved[0] = 0;
    s->reserved[1] = 0;
    __threadfence_system();
    st_release_sys_u64(&s->seq, mh + 1);
    st->mbox_head = mh + 1;
    st_release_sys_u64(&mc->head, mh + 1);
    stop = ld_acquire_sys_u64(&mc->stop_trial_id) == trial_id;
  }
  const unsigned long long head = st->head;
  const unsigned long long tail = ld_acquire_sys_u64(reinterpret_cast<const uint64_t*>(&ring->tail));
  if (head - tail >= slots) {  // consumer is behind: metrics are droppable, never block the stream
    st->dropped += 1;
    st_release_sys_u64(reinterpret_cast<uint64_t*>(&ring->dropped), st->dropped);
  } else {
    MgRecord* s = &sl[head & (slots - 1)];
    s->type = MG_REC_METRIC;
    s->worker = static_cast<uint16_t>(worker);
    s->aux = 1;  // produced on device
    s->trial_id = trial_id;
    s->step = step;
    s->value = value;
    s->t_ns = now;
    s->reserved[0] = 0;
    s->reserved[1] = 0;
    __threadfence_system();
    st_release_sys_u64(&s->seq, head + 1);
    st->head = head + 1;
    st_release_sys_u64(reinterpret_cast<uint64_t*>(&ring->head), head + 1);
  }
  stop = stop || ld_acquire_sys_u64(reinterpret_cast<const uint64_t*>(&ctl->stop_trial_id)) == trial_id;
  st->last_stop = stop ? 1u : 0u;
}

__global__ void globaltimer_probe_kernel(unsigned long long* out) { *out = globaltimer_ns(); }

// ---------------------------------------------------------------------------------------------------
// K9: persistent scheduler / policy kernel
// ---------------------------------------------------------------------------------------------------
struct SchedArgs {
  uint8_t* base;
  unsigned long long off_ctl, off_host_ring, off_dev_ring, ring_stride;
  unsigned int num_workers, slots;
  int direction;
  unsigned int es_min, es_interval, max_steps, cap_final;
  // device scratch
  float* final_tab;          // [cap_final][max_steps] prefix averages of finalized trials
  unsigned int* final_len;   // [cap_final]
  unsigned int* num_final;   // [1]
  float* run_tab;            // [num_workers][max_steps] prefix averages of the running trial
  unsigned long long* cursors;  // [num_workers][3]: host ring, host-mapped device ring, mailbox ring
  volatile unsigned int* quit;  // host-mapped: != 0 terminates the kernel
  unsigned long long* stats;    // [0] polls, [1] metrics seen, [2] stops issued, [3] finals seen, [4] mailbox records,
                                // [5] next-trial (ASHA) decisions, [6] prefetched trials handed off
  uint8_t* mailbox;             // device mailbox (local HBM) or nullptr
  unsigned int mbox_slots;
  long long gt_offset_ns;
  MgAsha* asha;                 // rung tables (device memory) or nullptr: next-trial decisions are taken here
  unsigned int* asha_lock;      // [1] the worker threads of this CTA file finals one at a time
};

struct RunState {
  unsigned long long trial_id;
  long long last_step;
  unsigned int n;
  double sum, best;
  bool stopped;
};

__device__ float median_of_column(const SchedArgs& a, unsigned int nfin, unsigned int col,
                                  unsigned int* count_out) {
  // selection by counting: nfin is small (<= cap_final), runs in one thread
  unsigned int cnt = 0;
  for (unsigned int i = 0; i < nfin; ++i)
    if (a.final_len[i] > col) ++cnt;
  *count_out = cnt;
  if (cnt == 0) return 0.f;
  const unsigned int k_lo = (cnt - 1) / 2, k_hi = cnt / 2;
  float v_lo = 0.f, v_hi = 0.f;
  for (unsigned int i = 0; i < nfin; ++i) {
    if (a.final_len[i] <= col) continue;
    const float x = a.final_tab[static_cast<size_t>(i) * a.max_steps + col];
    unsigned int less = 0, eq = 0;
    for (unsigned int j = 0; j < nfin; ++j) {
      if (a.final_len[j] <= col) continue;
      const float y = a.final_tab[static_cast<size_t>(j) * a.max_steps + col];
      less += (y < x);
      eq += (y == x);
    }
    if (less <= k_lo && k_lo < less + eq) v_lo = x;
    if (less <= k_hi && k_hi < less + eq) v_hi = x;
  }
  return 0.5f * (v_lo + v_hi);
}

__global__ void __launch_bounds__(32)
sched_kernel(const SchedArgs a) {
  const unsigned int w = threadIdx.x;
  const bool active = w < a.num_workers;
  RunState rs;
  rs.trial_id = 0; rs.last_step = LLONG_MIN; rs.n = 0; rs.sum = 0; rs.best = 0; rs.stopped = false;
  MgWorkerCtl* ctl = active ? reinterpret_cast<MgWorkerCtl*>(a.base + a.off_ctl) + w : nullptr;
  unsigned long long polls = 0;
  while (*a.quit == 0) {
    if (active) {
      for (int which = 0; which < 3; ++which) {
        const unsigned long long off = (which ? a.off_dev_ring : a.off_host_ring) + w * a.ring_stride;
        MgRecord* sl = reinterpret_cast<MgRecord*>(a.base + off + sizeof(MgRing));
        unsigned int slots = a.slots;
        // A worker whose GPU mapped the mailbox stores every device record there (and K9 reads it from its own
        // HBM); one that could not (no P2P, or it attached before the handle was published) only fills the
        // host-mapped device ring. Exactly one of the two is consumed per worker.
        const bool uses_mailbox =
            a.mailbox && ld_acquire_sys_u64(&(reinterpret_cast<MgMailboxCtl*>(a.mailbox) + w)->head) > 0;
        if (which == 1 && uses_mailbox) continue;
        if (which == 2) {
          if (!uses_mailbox) continue;
          sl = reinterpret_cast<MgRecord*>(a.mailbox + static_cast<size_t>(a.num_workers) * sizeof(MgMailboxCtl) +
                                           static_cast<size_t>(w) * a.mbox_slots * sizeof(MgRecord));
          slots = a.mbox_slots;
        }
        unsigned long long cur = a.cursors[w * 3 + which];
        for (int burst = 0; burst < 64; ++burst) {
          MgRecord* s = &sl[cur & (slots - 1)];
          const unsigned long long seq = ld_acquire_sys_u64(&s->seq);
          if (seq < cur + 1) break;                        // nothing new
          if (seq > cur + 1) { cur = seq - 1; }            // we were lapped: resynchronise
          const unsigned int type = s->type;
          const unsigned long long tid = s->trial_id;
          const long long step = s->step;
          const double value = s->value;
          ++cur;
          if (which == 2) atomicAdd(&a.stats[4], 1ull);
          if (type == MG_REC_METRIC) {
            if (rs.trial_id != tid) {
              rs.trial_id = tid; rs.last_step = LLONG_MIN; rs.n = 0; rs.sum = 0; rs.stopped = false;
            }
            if (step > rs.last_step) {
              rs.last_step = step;
              rs.n += 1;
              rs.sum += value;
              rs.best = rs.n == 1 ? value
                                  : (a.direction >= 0 ? fmax(rs.best, value) : fmin(rs.best, value));
              if (rs.n <= a.max_steps)
                a.run_tab[static_cast<size_t>(w) * a.max_steps + rs.n - 1] =
                    static_cast<float>(rs.sum / rs.n);
              atomicAdd(&a.stats[1], 1ull);
              const unsigned int nfin = *reinterpret_cast<volatile unsigned int*>(a.num_final);
              if (!rs.stopped && nfin > a.es_min && rs.n <= a.max_steps &&
                  (a.es_interval <= 1 || (step % a.es_interval) == 0)) {
                unsigned int cnt = 0;
                const float med = median_of_column(a, nfin, rs.n - 1, &cnt);
                if (cnt > 0) {
                  const bool stop = a.direction >= 0 ? (rs.best < med) : (rs.best > med);
                  if (stop) {
                    rs.stopped = true;
                    const unsigned long long t_stop = globaltimer_ns() + static_cast<unsigned long long>(a.gt_offset_ns);
                    if (a.mailbox) {
                      MgMailboxCtl* mc = reinterpret_cast<MgMailboxCtl*>(a.mailbox) + w;
                      mc->stop_t_ns = t_stop;
                      st_release_sys_u64(&mc->stop_trial_id, tid);   // the worker's K8 reads this word over NVLink
                      red_release_sys_add_u32(&mc->stop_seq, 1u);
                    }
                    ctl->stop_t_ns = t_stop;
                    st_release_sys_u64(reinterpret_cast<uint64_t*>(&ctl->stop_trial_id), tid);
                    red_release_sys_add_u32(&ctl->stop_seq, 1u);
                    atomicAdd(&a.stats[2], 1ull);
                  }
                }
              }
            }
          } else if (type == MG_REC_FINAL) {
            if (rs.trial_id == tid && rs.n > 0) {
              const unsigned int row = atomicAdd(a.num_final + 1, 1u);  // reserve
              if (row < a.cap_final) {
                const unsigned int len = rs.n < a.max_steps ? rs.n : a.max_steps;
                for (unsigned int i = 0; i < len; ++i)
                  a.final_tab[static_cast<size_t>(row) * a.max_steps + i] =
                      a.run_tab[static_cast<size_t>(w) * a.max_steps + i];
                a.final_len[row] = len;
                __threadfence();
                atomicAdd(a.num_final, 1u);  // publish (rows may complete out of order; lens guard reads)
              }
            }
            rs.trial_id = 0; rs.n = 0; rs.sum = 0; rs.last_step = LLONG_MIN; rs.stopped = false;
            atomicAdd(&a.stats[3], 1ull);
            const unsigned long long tag = s->reserved[0];
            if (a.asha != nullptr && tag > 0) {
              // ASHA on the device (reference: maggy/optimizer/asha.py:95-147 runs on the driver after the FINAL
              // message): file the result in its rung, scan for a promotable configuration, publish the decision
              // for this worker. The driver only materialises it (samples / copies the hyperparameters).
              unsigned long long parent = 0;
              unsigned int rung = 0, kind = MG_DEC_NONE;
              bool done = false;
              while (!done) {
                if (atomicCAS(a.asha_lock, 0u, 1u) == 0u) {
                  __threadfence();
                  mg_asha_record(a.asha, static_cast<unsigned int>(tag - 1), tid, value);
                  kind = mg_asha_decide(a.asha, reinterpret_cast<uint64_t*>(&parent), &rung);
                  __threadfence();
                  atomicExch(a.asha_lock, 0u);
                  done = true;
                }
              }
              if (kind == MG_DEC_NEW) {
                // prefetched hand-off (reference: the worker polls GET with sleep(1) until the driver has an answer,
                // maggy/core/rpc.py:739-748): the driver parked this worker's next fresh configuration in the staged
                // area; publishing it is five words and one release store — the worker is training again before the
                // driver has even dequeued the FINAL message.
                const unsigned int seq = ld_acquire_sys_u32(&ctl->assign_seq);
                if (ld_acquire_sys_u32(&ctl->staged_seq) == seq + 1 && ld_acquire_sys_u32(&ctl->ack_seq) == seq) {
                  ctl->assign_len = ctl->staged_len;
                  ctl->assign_trial_id = ctl->staged_trial_id;
                  ctl->assign_tag = ctl->staged_tag;
                  ctl->assign_src = 1;
                  st_release_sys_u64(reinterpret_cast<uint64_t*>(&ctl->stop_trial_id), 0ull);
                  __threadfence_system();
                  st_release_sys_u32(&ctl->assign_seq, seq + 1);
                  kind = MG_DEC_NEW_PUBLISHED;
                  atomicAdd(&a.stats[6], 1ull);
                }
              }
              ctl->dec_kind = kind;
              ctl->dec_rung = rung;
              ctl->dec_parent = parent;
              ctl->dec_t_ns = globaltimer_ns() + static_cast<unsigned long long>(a.gt_offset_ns);
              __threadfence_system();
              red_release_sys_add_u32(&ctl->dec_seq, 1u);
              atomicAdd(&a.stats[5], 1ull);
            }
          }
        }
        a.cursors[w * 3 + which] = cur;
      }
    }
    ++polls;
    if (w == 0 && (polls & 1023) == 0) a.stats[0] = polls;
    __nanosleep(2000);
  }
  if (w == 0) a.stats[0] = polls;
}

}  // namespace

extern "C" {

// Pin + map the shared control segment into this process' CUDA address space. Returns the device
// pointer (0 on failure). Safe to call once per process per segment.
unsigned long long mg_ring_map(void* host_base, unsigned long long bytes) {
  cudaError_t e = cudaHostRegister(host_base, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    // A previous segment that lived at this virtual address was unmapped without being unregistered: the
    // stale registration still points at its old physical pages. Drop it and register the new mapping.
    cudaGetLastError();
    cudaHostUnregister(host_base);
    cudaGetLastError();
    e = cudaHostRegister(host_base, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  void* dptr = nullptr;
  if (cudaHostGetDevicePointer(&dptr, host_base, 0) != cudaSuccess) return 0;
  return reinterpret_cast<unsigned long long>(dptr);
}

int mg_ring_unmap(void* host_base) {
  return cudaHostUnregister(host_base) == cudaSuccess ? 0 : -1;
}

// CLOCK_MONOTONIC - %globaltimer of the current device, in ns (median of 5 probes): added to every device stamp
long long mg_gt_offset_ns() {
  unsigned long long* pinned = nullptr;
  if (cudaHostAlloc(&pinned, sizeof(*pinned), cudaHostAllocMapped) != cudaSuccess) return 0;
  long long best[5];
  for (int i = 0; i < 5; ++i) {
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    globaltimer_probe_kernel<<<1, 1>>>(pinned);
    cudaDeviceSynchronize();
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const long long m0 = t0.tv_sec * 1000000000ll + t0.tv_nsec, m1 = t1.tv_sec * 1000000000ll + t1.tv_nsec;
    best[i] = (m0 + m1) / 2 - static_cast<long long>(*pinned);
  }
  cudaFreeHost(pinned);
  for (int i = 0; i < 5; ++i)
    for (int j = i + 1; j < 5; ++j)
      if (best[j] < best[i]) { long long t = best[i]; best[i] = best[j]; best[j] = t; }
  return best[2];
}

// ---- device mailbox -------------------------------------------------------------------------------------------
// create (driver process, on the scheduler's GPU): zeroed allocation + its CUDA IPC handle (64 bytes)
unsigned long long mg_mailbox_create(unsigned int num_workers, unsigned int slots, unsigned char* handle_out) {
  void* p = nullptr;
  const size_t bytes = mg_mailbox_bytes(num_workers, slots);
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaMemset(p, 0, bytes);
  cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return 0; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  return reinterpret_cast<unsigned long long>(p);
}
int mg_mailbox_free(unsigned long long p) { return cudaFree(reinterpret_cast<void*>(p)) == cudaSuccess ? 0 : -1; }
// open (worker process): peer mapping of the scheduler GPU's allocation into this process; 0 if P2P is not possible
unsigned long long mg_mailbox_open(const unsigned char* handle) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return 0; }
  return reinterpret_cast<unsigned long long>(p);
}
int mg_mailbox_close(unsigned long long p) {
  return cudaIpcCloseMemHandle(reinterpret_cast<void*>(p)) == cudaSuccess ? 0 : -1;
}

// dtype: 0 fp32, 1 bf16, 2 fp16, 3 fp64. `producer_state` = 32 zeroed bytes of device memory. `mailbox` may be 0.
int mg_report_store(unsigned long long dev_base, unsigned long long off_ring,
                    unsigned long long off_ctl, unsigned int slots, const void* metric, int dtype,
                    long long step, unsigned long long trial_id, unsigned int worker,
                    void* producer_state, unsigned long long mailbox, unsigned int mbox_workers,
                    unsigned int mbox_slots, long long gt_offset_ns, void* stream) {
  report_store_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint8_t*>(dev_base), off_ring, off_ctl, slots, metric, dtype, step, trial_id,
      worker, reinterpret_cast<DevProducer*>(producer_state), reinterpret_cast<uint8_t*>(mailbox), mbox_workers,
      mbox_slots, gt_offset_ns);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

// scratch layout (device memory, zeroed by the caller), in this order:
//   float  final_tab[cap_final*max_steps]; uint final_len[cap_final]; uint num_final[2];
//   float  run_tab[num_workers*max_steps]; u64 cursors[num_workers*3]; u64 stats[8]
unsigned long long mg_sched_scratch_bytes(unsigned int num_workers, unsigned int max_steps,
                                          unsigned int cap_final) {
  unsigned long long b = 0;
  b += 4ull * cap_final * max_steps;
  b += 4ull * cap_final;
  b += 8;
  b += 4ull * num_workers * max_steps;
  b = (b + 7) / 8 * 8;
  b += 8ull * num_workers * 3;
  b += 8ull * 8;
  return b;
}

// bytes of the device buffer mg_sched_start's asha_table needs: header + rung tables + the lock word
unsigned long long mg_asha_table_bytes(unsigned int cap) { return mg_asha_bytes(cap) + 8; }

int mg_sched_start(unsigned long long dev_base, unsigned long long off_ctl,
                   unsigned long long off_host_ring, unsigned long long off_dev_ring,
                   unsigned long long ring_stride, unsigned int num_workers, unsigned int slots,
                   int direction, unsigned int es_min, unsigned int es_interval,
                   unsigned int max_steps, unsigned int cap_final, void* scratch,
                   unsigned long long quit_dev_ptr, unsigned long long mailbox, unsigned int mbox_slots,
                   long long gt_offset_ns, void* asha_table, void* stream) {
  if (num_workers > 32) return -1;
  SchedArgs a;
  a.base = reinterpret_cast<uint8_t*>(dev_base);
  a.off_ctl = off_ctl; a.off_host_ring = off_host_ring; a.off_dev_ring = off_dev_ring;
  a.ring_stride = ring_stride; a.num_workers = num_workers; a.slots = slots;
  a.direction = direction; a.es_min = es_min; a.es_interval = es_interval ? es_interval : 1;
  a.max_steps = max_steps; a.cap_final = cap_final;
  uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
  a.final_tab = reinterpret_cast<float*>(p); p += 4ull * cap_final * max_steps;
  a.final_len = reinterpret_cast<unsigned int*>(p); p += 4ull * cap_final;
  a.num_final = reinterpret_cast<unsigned int*>(p); p += 8;
  a.run_tab = reinterpret_cast<float*>(p); p += 4ull * num_workers * max_steps;
  p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 7) / 8 * 8);
  a.cursors = reinterpret_cast<unsigned long long*>(p); p += 8ull * num_workers * 3;
  a.stats = reinterpret_cast<unsigned long long*>(p);
  a.quit = reinterpret_cast<volatile unsigned int*>(quit_dev_ptr);
  a.mailbox = reinterpret_cast<uint8_t*>(mailbox);
  a.mbox_slots = mbox_slots;
  a.gt_offset_ns = gt_offset_ns;
  // asha_table: device buffer of mg_asha_table_bytes(cap) bytes whose MgAsha header (rf, max_rung, cap, direction) the
  // caller filled in and whose remainder is zero; null = next-trial decisions stay on the driver
  a.asha = reinterpret_cast<MgAsha*>(asha_table);
  a.asha_lock = nullptr;
  if (asha_table != nullptr) {
    unsigned int cap = 0;
    if (cudaMemcpy(&cap, reinterpret_cast<uint8_t*>(asha_table) + 8, 4, cudaMemcpyDeviceToHost) != cudaSuccess || cap == 0)
      return -2;
    a.asha_lock = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(asha_table) + mg_asha_bytes(cap));
  }
  sched_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -400 - static_cast<int>(e);
}

}  // extern "C"
