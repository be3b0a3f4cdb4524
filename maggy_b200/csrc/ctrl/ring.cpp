// Host side of the maggy_b200 control plane (see include/mg_ring.h for the layout and the mapping to
// the reference's RPC messages). Pure C++17 + POSIX: loads on CPU-only machines, so the whole
// driver <-> worker protocol (register, assign, report, early stop, final, idle, global stop,
// worker death) is testable without a GPU while using exactly the memory layout the device code uses.
//
// Threads: each worker process has ONE host producer per ring (calls are serialised by a per-handle
// mutex); the driver runs a native "pump" thread that drains every ring, applies the device-decidable
// early-stop policy (median rule, reference maggy/earlystop/medianrule.py:28-60) in native code —
// a STOP word is written microseconds after the offending METRIC record lands, independent of the
// Python GIL — and hands the records to Python through an in-process queue.
#include "../include/mg_ring.h"
#include "../include/mg_asha.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <linux/futex.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

namespace {

inline uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return static_cast<uint64_t>(ts.tv_sec) * 1000000000ull + ts.tv_nsec;
}

inline int futex_wait(uint32_t* addr, uint32_t expected, int64_t timeout_us) {
  timespec ts;
  timespec* tsp = nullptr;
  if (timeout_us >= 0) {
    ts.tv_sec = timeout_us / 1000000;
    ts.tv_nsec = (timeout_us % 1000000) * 1000;
    tsp = &ts;
  }
  return static_cast<int>(syscall(SYS_futex, addr, FUTEX_WAIT, expected, tsp, nullptr, 0));
}
inline void futex_wake_all(uint32_t* addr) {
  syscall(SYS_futex, addr, FUTEX_WAKE, INT32_MAX, nullptr, nullptr, 0);
}

template <class T>
inline T load_acq(const T* p) {
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
template <class T>
inline void store_rel(T* p, T v) {
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
}

struct PolicyRun {  // per-worker running trial state for the median rule
  uint64_t trial_id = 0;
  int64_t last_step = INT64_MIN;
  uint32_t n = 0;
  double sum = 0.0, best = 0.0;
  std::vector<float> prefix_avg;
  bool stopped = false;
};

struct Handle {
  std::string name;
  int fd = -1;
  uint8_t* base = nullptr;
  size_t bytes = 0;
  bool owner = false;
  std::mutex post_mu;  // serialises host producers in this process
  // driver-side pump
  std::thread pump;
  std::atomic<bool> pump_run{false};
  std::mutex q_mu;
  std::condition_variable q_cv;
  std::deque<MgRecord> queue;
  std::vector<PolicyRun> runs;
  std::mutex pol_mu;
  std::vector<std::vector<float>> final_rows;  // prefix averages of finalized trials
  uint64_t stops_issued = 0;

  MgHeader* hdr() const { return reinterpret_cast<MgHeader*>(base); }
  MgWorkerCtl* ctl(uint32_t w) const {
    return reinterpret_cast<MgWorkerCtl*>(base + hdr()->off_ctl) + w;
  }
  MgRing* ring(bool dev, uint32_t w) const {
    uint64_t off = dev ? hdr()->off_dev_ring : hdr()->off_host_ring;
    return reinterpret_cast<MgRing*>(base + off + static_cast<uint64_t>(w) * hdr()->ring_stride);
  }
  MgRecord* slots(bool dev, uint32_t w) const {
    return reinterpret_cast<MgRecord*>(reinterpret_cast<uint8_t*>(ring(dev, w)) + sizeof(MgRing));
  }
  uint8_t* assign_buf(uint32_t w) const {
    return base + hdr()->off_assign + static_cast<uint64_t>(w) * hdr()->payload_bytes;
  }
  uint8_t* staged_buf(uint32_t w) const {
    return base + hdr()->off_staged + static_cast<uint64_t>(w) * hdr()->payload_bytes;
  }
  uint8_t* blob_buf(uint32_t w) const {
    return base + hdr()->off_blob + static_cast<uint64_t>(w) * hdr()->payload_bytes;
  }
  MgPolicy* policy() const { return reinterpret_cast<MgPolicy*>(base + hdr()->off_policy); }
};

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// ---- median stopping rule, native (same decision as reference medianrule.py:28-60) -----------------
// returns true if the trial must stop
bool policy_on_metric(Handle* h, uint32_t w, uint64_t trial_id, int64_t step, double value) {
  MgPolicy* pol = h->policy();
  std::lock_guard<std::mutex> lk(h->pol_mu);
  PolicyRun& r = h->runs[w];
  if (r.trial_id != trial_id) {
    r = PolicyRun();
    r.trial_id = trial_id;
  }
  if (step <= r.last_step) return false;  // not a new unique step
  r.last_step = step;
  r.n += 1;
  r.sum += value;
  const int dir = pol->direction;
  if (r.n == 1) r.best = value;
  else r.best = dir >= 0 ? std::max(r.best, value) : std::min(r.best, value);
  r.prefix_avg.push_back(static_cast<float>(r.sum / r.n));
  if (!pol->enabled || r.stopped) return false;
  if (h->final_rows.size() <= pol->es_min) return false;
  if (pol->es_interval > 1 && (step % pol->es_interval) != 0) return false;
  std::vector<float> col;
  col.reserve(h->final_rows.size());
  for (const auto& row : h->final_rows)
    if (row.size() >= r.n) col.push_back(row[r.n - 1]);
  if (col.empty()) return false;
  std::sort(col.begin(), col.end());
  const size_t k = col.size();
  const double median = (k & 1) ? col[k / 2] : 0.5 * (static_cast<double>(col[k / 2 - 1]) + col[k / 2]);
  const bool stop = dir >= 0 ? (r.best < median) : (r.best > median);
  if (stop) r.stopped = true;
  return stop;
}

void policy_on_final(Handle* h, uint32_t w, uint64_t trial_id) {
  std::lock_guard<std::mutex> lk(h->pol_mu);
  PolicyRun& r = h->runs[w];
  if (r.trial_id == trial_id && !r.prefix_avg.empty()) h->final_rows.push_back(r.prefix_avg);
  h->policy()->num_final = static_cast<uint32_t>(h->final_rows.size());
  r = PolicyRun();
}

void issue_stop(Handle* h, uint32_t w, uint64_t trial_id) {
  MgWorkerCtl* c = h->ctl(w);
  c->stop_t_ns = now_ns();
  store_rel(&c->stop_trial_id, trial_id);
  __atomic_fetch_add(&c->stop_seq, 1u, __ATOMIC_RELEASE);
  h->stops_issued++;
}

// drain one ring; returns number of records moved into `out`
int drain_ring(Handle* h, bool dev, uint32_t w, std::vector<MgRecord>& out) {
  MgRing* rg = h->ring(dev, w);
  MgRecord* sl = h->slots(dev, w);
  const uint32_t mask = h->hdr()->ring_slots - 1;
  uint64_t tail = rg->tail;
  int n = 0;
  for (;;) {
    MgRecord* s = &sl[tail & mask];
    if (load_acq(&s->seq) != tail + 1) break;
    MgRecord rec;
    std::memcpy(&rec, s, sizeof(MgRecord));
    rec.worker = static_cast<uint16_t>(w);
    out.push_back(rec);
    ++tail;
    ++n;
  }
  if (n) store_rel(&rg->tail, tail);
  return n;
}

void pump_loop(Handle* h) {
  std::vector<MgRecord> batch;
  MgHeader* hd = h->hdr();
  while (h->pump_run.load(std::memory_order_acquire)) {
    uint32_t bell = load_acq(&hd->doorbell);
    batch.clear();
    for (uint32_t w = 0; w < hd->num_workers; ++w) {
      drain_ring(h, false, w, batch);
      drain_ring(h, true, w, batch);
    }
    if (batch.empty()) {
      // host producers ring the doorbell; device producers cannot, so bound the sleep
      futex_wait(&hd->doorbell, bell, 200);
      continue;
    }
    for (auto& rec : batch) {
      if (rec.type == MG_REC_METRIC) {
        if (policy_on_metric(h, rec.worker, rec.trial_id, rec.step, rec.value)) {
          issue_stop(h, rec.worker, rec.trial_id);
          rec.reserved[0] = 1;  // tell Python the native policy already stopped this trial
        }
      } else if (rec.type == MG_REC_FINAL) {
        policy_on_final(h, rec.worker, rec.trial_id);
      }
    }
    {
      std::lock_guard<std::mutex> lk(h->q_mu);
      for (auto& rec : batch) h->queue.push_back(rec);
    }
    h->q_cv.notify_all();
  }
}

int post_record(Handle* h, uint32_t w, uint16_t type, uint64_t trial_id, int64_t step, double value,
                uint32_t aux) {
  MgHeader* hd = h->hdr();
  if (w >= hd->num_workers) return -1;
  MgRing* rg = h->ring(false, w);
  MgRecord* sl = h->slots(false, w);
  const uint32_t slots = hd->ring_slots;
  std::lock_guard<std::mutex> lk(h->post_mu);
  uint64_t head = rg->head;
  // back-pressure: wait for the consumer (never drop host records: FINAL/REG must arrive)
  int spins = 0;
  while (head - load_acq(&rg->tail) >= slots) {
    if (++spins > 2000000) return -2;
    usleep(50);
  }
  MgRecord* s = &sl[head & (slots - 1)];
  s->type = type;
  s->worker = static_cast<uint16_t>(w);
  s->aux = aux;
  s->trial_id = trial_id;
  s->step = step;
  s->value = value;
  s->t_ns = now_ns();
  // FINAL echoes the tag the driver attached to the assignment (ASHA: rung + 1) so the device scheduler can file it
  s->reserved[0] = type == MG_REC_FINAL ? h->ctl(w)->assign_tag : 0;
  s->reserved[1] = 0;
  store_rel(&s->seq, head + 1);
  store_rel(&rg->head, head + 1);
  __atomic_fetch_add(&hd->doorbell, 1u, __ATOMIC_RELEASE);
  futex_wake_all(&hd->doorbell);
  store_rel(&h->ctl(w)->heartbeat_ns, now_ns());
  return 0;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------------
void* mgr_create(const char* name, uint32_t num_workers, uint32_t ring_slots, uint32_t payload_bytes) {
  if (num_workers == 0 || (ring_slots & (ring_slots - 1)) != 0 || ring_slots < 8) return nullptr;
  MgHeader tmp;
  std::memset(&tmp, 0, sizeof(tmp));
  uint64_t off = align_up(sizeof(MgHeader), 4096);
  tmp.off_ctl = off;
  off = align_up(off + sizeof(MgWorkerCtl) * num_workers, 4096);
  tmp.ring_stride = align_up(mg_ring_bytes(ring_slots), 4096);
  tmp.off_host_ring = off;
  off += tmp.ring_stride * num_workers;
  tmp.off_dev_ring = off;
  off += tmp.ring_stride * num_workers;
  payload_bytes = static_cast<uint32_t>(align_up(payload_bytes, 4096));
  tmp.off_assign = off;
  off += static_cast<uint64_t>(payload_bytes) * num_workers;
  tmp.off_blob = off;
  off += static_cast<uint64_t>(payload_bytes) * num_workers;
  tmp.off_policy = off;
  off = align_up(off + sizeof(MgPolicy), 4096);
  tmp.off_staged = off;
  off += static_cast<uint64_t>(payload_bytes) * num_workers;
  tmp.total_bytes = off;

  shm_unlink(name);
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (ftruncate(fd, static_cast<off_t>(tmp.total_bytes)) != 0) {
    close(fd);
    shm_unlink(name);
    return nullptr;
  }
  void* p = mmap(nullptr, tmp.total_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) {
    close(fd);
    shm_unlink(name);
    return nullptr;
  }
  std::memset(p, 0, tmp.total_bytes);
  Handle* h = new Handle();
  h->name = name;
  h->fd = fd;
  h->base = static_cast<uint8_t*>(p);
  h->bytes = tmp.total_bytes;
  h->owner = true;
  MgHeader* hd = h->hdr();
  *hd = tmp;
  hd->version = MG_RING_VERSION;
  hd->num_workers = num_workers;
  hd->ring_slots = ring_slots;
  hd->payload_bytes = payload_bytes;
  hd->driver_alive = 1;
  h->policy()->direction = 1;
  h->runs.resize(num_workers);
  store_rel(&hd->magic, static_cast<uint64_t>(MG_RING_MAGIC));
  return h;
}

void* mgr_attach(const char* name) {
  int fd = shm_open(name, O_RDWR, 0600);
  if (fd < 0) return nullptr;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < static_cast<off_t>(sizeof(MgHeader))) {
    close(fd);
    return nullptr;
  }
  void* p = mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) {
    close(fd);
    return nullptr;
  }
  MgHeader* hd = static_cast<MgHeader*>(p);
  if (load_acq(&hd->magic) != MG_RING_MAGIC || hd->version != MG_RING_VERSION) {
    munmap(p, static_cast<size_t>(st.st_size));
    close(fd);
    return nullptr;
  }
  Handle* h = new Handle();
  h->name = name;
  h->fd = fd;
  h->base = static_cast<uint8_t*>(p);
  h->bytes = static_cast<size_t>(st.st_size);
  h->runs.resize(hd->num_workers);
  return h;
}

void mgr_stop_pump(void* hv);

void mgr_close(void* hv, int unlink_it) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h) return;
  mgr_stop_pump(h);
  if (h->owner) h->hdr()->driver_alive = 0;
  munmap(h->base, h->bytes);
  close(h->fd);
  if (unlink_it) shm_unlink(h->name.c_str());
  delete h;
}

void* mgr_base(void* hv) { return static_cast<Handle*>(hv)->base; }
uint64_t mgr_bytes(void* hv) { return static_cast<Handle*>(hv)->bytes; }
uint32_t mgr_num_workers(void* hv) { return static_cast<Handle*>(hv)->hdr()->num_workers; }
// device mailbox handle (published by the driver once its scheduler kernel owns the allocation; read by the workers)
void mgr_set_mailbox(void* hv, const uint8_t* handle, uint32_t slots, uint64_t bytes) {
  MgHeader* hd = static_cast<Handle*>(hv)->hdr();
  memcpy(hd->mailbox_handle, handle, 64);
  hd->mailbox_slots = slots;
  hd->mailbox_bytes = bytes;
  __atomic_store_n(&hd->mailbox_valid, handle ? 1u : 0u, __ATOMIC_RELEASE);
}
void mgr_clear_mailbox(void* hv) {
  __atomic_store_n(&static_cast<Handle*>(hv)->hdr()->mailbox_valid, 0u, __ATOMIC_RELEASE);
}
uint32_t mgr_get_mailbox(void* hv, uint8_t* handle_out, uint32_t* slots_out) {
  MgHeader* hd = static_cast<Handle*>(hv)->hdr();
  if (__atomic_load_n(&hd->mailbox_valid, __ATOMIC_ACQUIRE) == 0) return 0;
  memcpy(handle_out, hd->mailbox_handle, 64);
  *slots_out = hd->mailbox_slots;
  return 1;
}
// size of one assign / blob slot: workers size their receive buffer from the segment, not from a constant
uint32_t mgr_payload_bytes(void* hv) { return static_cast<Handle*>(hv)->hdr()->payload_bytes; }
uint64_t mgr_now_ns() { return now_ns(); }

// byte offsets so the CUDA side (ring_device.cu) can address the same structures after
// cudaHostRegister(base, bytes, cudaHostRegisterMapped)
uint64_t mgr_off_dev_ring(void* hv, uint32_t w) {
  Handle* h = static_cast<Handle*>(hv);
  return h->hdr()->off_dev_ring + static_cast<uint64_t>(w) * h->hdr()->ring_stride;
}
uint64_t mgr_off_host_ring(void* hv, uint32_t w) {
  Handle* h = static_cast<Handle*>(hv);
  return h->hdr()->off_host_ring + static_cast<uint64_t>(w) * h->hdr()->ring_stride;
}
uint64_t mgr_ring_stride(void* hv) { return static_cast<Handle*>(hv)->hdr()->ring_stride; }
uint64_t mgr_off_ctl(void* hv, uint32_t w) {
  Handle* h = static_cast<Handle*>(hv);
  return h->hdr()->off_ctl + static_cast<uint64_t>(w) * sizeof(MgWorkerCtl);
}
uint32_t mgr_ring_slots(void* hv) { return static_cast<Handle*>(hv)->hdr()->ring_slots; }

// ---------------------------------------------------------------------------------------------------
// worker side
// ---------------------------------------------------------------------------------------------------
int mgr_post(void* hv, uint32_t w, uint32_t type, uint64_t trial_id, int64_t step, double value,
             uint32_t aux) {
  return post_record(static_cast<Handle*>(hv), w, static_cast<uint16_t>(type), trial_id, step, value,
                     aux);
}

// copy a blob (pickled final result / log text) into the worker's blob area, then post the record.
// Blocks until the driver has consumed the previous blob.
int mgr_post_blob(void* hv, uint32_t w, uint32_t type, uint64_t trial_id, int64_t step, double value,
                  const void* data, uint32_t len, int64_t timeout_us) {
  Handle* h = static_cast<Handle*>(hv);
  MgHeader* hd = h->hdr();
  if (w >= hd->num_workers || len > hd->payload_bytes) return -1;
  MgWorkerCtl* c = h->ctl(w);
  const uint64_t t0 = now_ns();
  while (load_acq(&c->blob_ack) != c->blob_seq) {
    if (timeout_us >= 0 && now_ns() - t0 > static_cast<uint64_t>(timeout_us) * 1000) return -3;
    usleep(20);
  }
  if (len) std::memcpy(h->blob_buf(w), data, len);
  c->blob_len = len;
  store_rel(&c->blob_seq, c->blob_seq + 1);
  return post_record(h, w, static_cast<uint16_t>(type), trial_id, step, value, len);
}

int mgr_register(void* hv, uint32_t w, uint32_t attempt, uint32_t pid) {
  Handle* h = static_cast<Handle*>(hv);
  if (w >= h->hdr()->num_workers) return -1;
  MgWorkerCtl* c = h->ctl(w);
  c->attempt = attempt;
  c->pid = pid;
  // a respawned worker must not see the assignment its predecessor already consumed
  c->ack_seq = load_acq(&c->assign_seq);
  // a predecessor may have died between writing a blob and the driver consuming it
  store_rel(&c->blob_seq, load_acq(&c->blob_ack));
  store_rel(&c->heartbeat_ns, now_ns());  // before the state flips: liveness checks key on the state
  store_rel(&c->state, static_cast<uint32_t>(MG_W_REGISTERED));
  return post_record(h, w, MG_REC_REG, 0, 0, 0.0, attempt);
}

void mgr_set_state(void* hv, uint32_t w, uint32_t state) {
  Handle* h = static_cast<Handle*>(hv);
  store_rel(&h->ctl(w)->state, state);
  store_rel(&h->ctl(w)->heartbeat_ns, now_ns());
}

void mgr_heartbeat(void* hv, uint32_t w) {
  store_rel(&static_cast<Handle*>(hv)->ctl(w)->heartbeat_ns, now_ns());
}

// Wait for the next assignment. Returns 1 and fills trial_id/len (payload copied into buf) when a new
// trial is available, 0 on timeout, 2 on global stop.
int mgr_wait_assign(void* hv, uint32_t w, int64_t timeout_us, uint64_t* trial_id, void* buf,
                    uint32_t cap, uint32_t* len) {
  Handle* h = static_cast<Handle*>(hv);
  MgHeader* hd = h->hdr();
  MgWorkerCtl* c = h->ctl(w);
  const uint64_t t0 = now_ns();
  for (;;) {
    uint32_t seq = load_acq(&c->assign_seq);
    if (seq != c->ack_seq) {
      uint32_t n = c->assign_len;
      if (n > cap) return -1;
      if (n) std::memcpy(buf, c->assign_src ? h->staged_buf(w) : h->assign_buf(w), n);
      *len = n;
      *trial_id = c->assign_trial_id;
      store_rel(&c->ack_seq, seq);
      store_rel(&c->heartbeat_ns, now_ns());
      return 1;
    }
    if (load_acq(&hd->gstop)) return 2;
    int64_t left = -1;
    if (timeout_us >= 0) {
      int64_t spent = static_cast<int64_t>((now_ns() - t0) / 1000);
      if (spent >= timeout_us) return 0;
      left = timeout_us - spent;
    }
    // The scheduler kernel publishes a prefetched trial with a plain store from the GPU (nobody futex-wakes us):
    // poll for the first 300 us after the call, nap briefly for the next 20 ms, then fall back to the bounded sleep
    // that notices a global stop / driver death even without a wake-up.
    const uint64_t waited = now_ns() - t0;
    if (waited < 300000) continue;
    int64_t cap_us = waited < 20000000 ? 200 : 50000;
    int64_t nap = (left < 0 || left > cap_us) ? cap_us : left;
    futex_wait(&c->assign_seq, seq, nap);
    store_rel(&c->heartbeat_ns, now_ns());
  }
}

// 1 if `trial_id` has been told to stop (the reporter raises EarlyStopException at its next broadcast)
int mgr_check_stop(void* hv, uint32_t w, uint64_t trial_id) {
  Handle* h = static_cast<Handle*>(hv);
  return load_acq(&h->ctl(w)->stop_trial_id) == trial_id && trial_id != 0 ? 1 : 0;
}

int mgr_gstop_set(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  return load_acq(&h->hdr()->gstop) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------
// driver side
// ---------------------------------------------------------------------------------------------------
int mgr_assign_tagged(void* hv, uint32_t w, uint64_t trial_id, const void* data, uint32_t len, uint32_t tag);
int mgr_assign(void* hv, uint32_t w, uint64_t trial_id, const void* data, uint32_t len) {
  return mgr_assign_tagged(hv, w, trial_id, data, len, 0);
}

// tag: opaque, echoed in the trial's FINAL record (ASHA: rung + 1; 0 = untagged, no device decision is taken)
int mgr_assign_tagged(void* hv, uint32_t w, uint64_t trial_id, const void* data, uint32_t len, uint32_t tag) {
  Handle* h = static_cast<Handle*>(hv);
  MgHeader* hd = h->hdr();
  if (w >= hd->num_workers || len > hd->payload_bytes) return -1;
  MgWorkerCtl* c = h->ctl(w);
  if (load_acq(&c->ack_seq) != c->assign_seq) return -2;  // previous assignment not picked up yet
  if (len) std::memcpy(h->assign_buf(w), data, len);
  c->assign_len = len;
  c->assign_trial_id = trial_id;
  c->assign_tag = tag;
  c->assign_src = 0;
  store_rel(&c->stop_trial_id, static_cast<uint64_t>(0));
  store_rel(&c->assign_seq, c->assign_seq + 1);
  futex_wake_all(&c->assign_seq);
  return 0;
}

// Park the worker's NEXT trial in its staged area. -2: one is already parked. The entry stays valid until somebody
// assigns something else to the worker (staged_seq == assign_seq + 1 no longer holds then).
int mgr_stage(void* hv, uint32_t w, uint64_t trial_id, const void* data, uint32_t len, uint32_t tag) {
  Handle* h = static_cast<Handle*>(hv);
  MgHeader* hd = h->hdr();
  if (w >= hd->num_workers || len > hd->payload_bytes) return -1;
  MgWorkerCtl* c = h->ctl(w);
  const uint32_t seq = load_acq(&c->assign_seq), next = seq + 1;
  if (load_acq(&c->staged_seq) == next) return -2;
  // the CURRENT assignment may itself sit in the staged area (it was published from there): the worker must have copied
  // it out (ack) before the area is overwritten
  if (c->assign_src == 1 && load_acq(&c->ack_seq) != seq) return -2;
  if (len) std::memcpy(h->staged_buf(w), data, len);
  c->staged_len = len;
  c->staged_tag = tag;
  c->staged_trial_id = trial_id;
  store_rel(&c->staged_seq, next);
  return 0;
}

// 1: a staged trial is parked and unpublished; 2: the worker's CURRENT assignment came out of the staged area;
// 0: neither. `trial_id` receives the staged / published trial id for 1 / 2.
int mgr_staged_state(void* hv, uint32_t w, uint64_t* trial_id) {
  MgWorkerCtl* c = static_cast<Handle*>(hv)->ctl(w);
  const uint32_t seq = load_acq(&c->assign_seq), st = load_acq(&c->staged_seq);
  *trial_id = c->staged_trial_id;
  if (st == seq + 1) return 1;
  if (st == seq && c->assign_src == 1) return 2;
  return 0;
}

// The driver-side publish of the parked trial (K9 does the same from the device: ring_device.cu publish_staged).
int mgr_publish_staged(void* hv, uint32_t w) {
  Handle* h = static_cast<Handle*>(hv);
  MgWorkerCtl* c = h->ctl(w);
  const uint32_t seq = load_acq(&c->assign_seq);
  if (load_acq(&c->staged_seq) != seq + 1 || load_acq(&c->ack_seq) != seq) return -2;
  c->assign_len = c->staged_len;
  c->assign_trial_id = c->staged_trial_id;
  c->assign_tag = c->staged_tag;
  c->assign_src = 1;
  store_rel(&c->stop_trial_id, static_cast<uint64_t>(0));
  store_rel(&c->assign_seq, seq + 1);
  futex_wake_all(&c->assign_seq);
  return 0;
}

void mgr_stop_trial(void* hv, uint32_t w, uint64_t trial_id) {
  issue_stop(static_cast<Handle*>(hv), w, trial_id);
}

void mgr_gstop(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  store_rel(&h->hdr()->gstop, 1u);
  for (uint32_t w = 0; w < h->hdr()->num_workers; ++w) {
    __atomic_fetch_add(&h->ctl(w)->assign_seq, 0u, __ATOMIC_RELEASE);
    futex_wake_all(&h->ctl(w)->assign_seq);
  }
}

// read the blob that accompanies a FINAL/LOG record of worker w; acks it so the worker may reuse the area
int mgr_read_blob(void* hv, uint32_t w, void* buf, uint32_t cap) {
  Handle* h = static_cast<Handle*>(hv);
  MgWorkerCtl* c = h->ctl(w);
  uint32_t seq = load_acq(&c->blob_seq);
  if (seq == c->blob_ack) return -1;
  uint32_t n = c->blob_len;
  if (n > cap) return -2;
  if (n) std::memcpy(buf, h->blob_buf(w), n);
  store_rel(&c->blob_ack, seq);
  return static_cast<int>(n);
}

uint64_t mgr_worker_heartbeat(void* hv, uint32_t w) {
  return load_acq(&static_cast<Handle*>(hv)->ctl(w)->heartbeat_ns);
}
uint32_t mgr_worker_state(void* hv, uint32_t w) {
  return load_acq(&static_cast<Handle*>(hv)->ctl(w)->state);
}
uint32_t mgr_worker_pid(void* hv, uint32_t w) { return static_cast<Handle*>(hv)->ctl(w)->pid; }
uint64_t mgr_dev_dropped(void* hv, uint32_t w) {
  return load_acq(&static_cast<Handle*>(hv)->ring(true, w)->dropped);
}
uint64_t mgr_stop_time_ns(void* hv, uint32_t w) { return static_cast<Handle*>(hv)->ctl(w)->stop_t_ns; }

// the device scheduler's n-th next-trial decision for worker w: returns dec_seq (acquire) and fills the fields
uint32_t mgr_decision(void* hv, uint32_t w, uint32_t* kind, uint32_t* rung, uint64_t* parent, uint64_t* t_ns) {
  MgWorkerCtl* c = static_cast<Handle*>(hv)->ctl(w);
  const uint32_t seq = load_acq(&c->dec_seq);
  *kind = c->dec_kind;
  *rung = c->dec_rung;
  *parent = c->dec_parent;
  *t_ns = c->dec_t_ns;
  return seq;
}

// ---- the rung tables of mg_asha.h on the host: the oracle the CPU tests compare maggy_b200.optimizer.Asha against
void* mgr_asha_create(uint32_t rf, uint32_t max_rung, uint32_t cap, int32_t direction) {
  void* mem = std::calloc(1, mg_asha_bytes(cap));
  if (!mem) return nullptr;
  MgAsha* a = static_cast<MgAsha*>(mem);
  a->rf = rf;
  a->max_rung = max_rung;
  a->cap = cap;
  a->direction = direction;
  return mem;
}
void mgr_asha_free(void* a) { std::free(a); }
uint32_t mgr_asha_final(void* av, uint32_t rung, uint64_t trial_id, double metric, uint64_t* parent, uint32_t* rung_out) {
  MgAsha* a = static_cast<MgAsha*>(av);
  mg_asha_record(a, rung, trial_id, metric);
  return mg_asha_decide(a, parent, rung_out);
}

void mgr_policy_config(void* hv, uint32_t enabled, int32_t direction, uint32_t es_min,
                       uint32_t es_interval) {
  Handle* h = static_cast<Handle*>(hv);
  std::lock_guard<std::mutex> lk(h->pol_mu);
  MgPolicy* p = h->policy();
  p->direction = direction;
  p->es_min = es_min;
  p->es_interval = es_interval ? es_interval : 1;
  store_rel(&p->enabled, enabled);
}
uint64_t mgr_policy_stops(void* hv) { return static_cast<Handle*>(hv)->stops_issued; }

void mgr_start_pump(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  if (h->pump_run.exchange(true)) return;
  h->pump = std::thread(pump_loop, h);
}

void mgr_stop_pump(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  if (!h->pump_run.exchange(false)) return;
  futex_wake_all(&h->hdr()->doorbell);
  if (h->pump.joinable()) h->pump.join();
}

// Blocking fetch of up to `max` records from the pump queue. Returns the count (0 on timeout).
int mgr_next(void* hv, MgRecord* out, int max, int64_t timeout_us) {
  Handle* h = static_cast<Handle*>(hv);
  std::unique_lock<std::mutex> lk(h->q_mu);
  if (h->queue.empty()) {
    if (timeout_us < 0) h->q_cv.wait(lk, [&] { return !h->queue.empty() || !h->pump_run.load(); });
    else h->q_cv.wait_for(lk, std::chrono::microseconds(timeout_us), [&] { return !h->queue.empty(); });
  }
  int n = 0;
  while (n < max && !h->queue.empty()) {
    out[n++] = h->queue.front();
    h->queue.pop_front();
  }
  return n;
}

// Synchronous drain without the pump thread (single-threaded tests / simple drivers)
int mgr_poll(void* hv, MgRecord* out, int max) {
  Handle* h = static_cast<Handle*>(hv);
  std::vector<MgRecord> batch;
  for (uint32_t w = 0; w < h->hdr()->num_workers; ++w) {
    drain_ring(h, false, w, batch);
    drain_ring(h, true, w, batch);
  }
  int n = 0;
  {
    std::lock_guard<std::mutex> lk(h->q_mu);  // records that did not fit into a previous call
    while (n < max && !h->queue.empty()) {
      out[n++] = h->queue.front();
      h->queue.pop_front();
    }
  }
  for (auto& rec : batch) {
    if (rec.type == MG_REC_METRIC) {
      if (policy_on_metric(h, rec.worker, rec.trial_id, rec.step, rec.value)) {
        issue_stop(h, rec.worker, rec.trial_id);
        rec.reserved[0] = 1;
      }
    } else if (rec.type == MG_REC_FINAL) {
      policy_on_final(h, rec.worker, rec.trial_id);
    }
    if (n < max) out[n++] = rec;
    else {
      std::lock_guard<std::mutex> lk(h->q_mu);
      h->queue.push_back(rec);
    }
  }
  return n;
}

}  // extern "C"
