// ThreadSanitizer stress of the shared-memory control rings ("TSAN build of the C++ host shim").
// Built by tests/test_ring_tsan.py as:  g++ -fsanitize=thread -O1 -g ring.cpp ring_stress.cpp -lpthread -lrt
// W worker threads post METRIC records (and FINAL blobs) through the worker-side API while the driver thread drains
// them through the pump, assigns trials and flips stop words. Every record must arrive exactly once and in per-worker
// order; TSAN must stay silent (a report makes the process exit with code 66).
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unistd.h>
#include <vector>

#include "mg_ring.h"

extern "C" {
void* mgr_create(const char*, uint32_t, uint32_t, uint32_t);
void* mgr_attach(const char*);
void mgr_close(void*, int);
int mgr_post(void*, uint32_t, uint32_t, uint64_t, int64_t, double, uint32_t);
int mgr_post_blob(void*, uint32_t, uint32_t, uint64_t, int64_t, double, const void*, uint32_t, int64_t);
int mgr_register(void*, uint32_t, uint32_t, uint32_t);
void mgr_heartbeat(void*, uint32_t);
int mgr_wait_assign(void*, uint32_t, int64_t, uint64_t*, void*, uint32_t, uint32_t*);
int mgr_check_stop(void*, uint32_t, uint64_t);
int mgr_assign(void*, uint32_t, uint64_t, const void*, uint32_t);
int mgr_stage(void*, uint32_t, uint64_t, const void*, uint32_t, uint32_t);
int mgr_publish_staged(void*, uint32_t);
void mgr_stop_trial(void*, uint32_t, uint64_t);
int mgr_read_blob(void*, uint32_t, void*, uint32_t);
void mgr_start_pump(void*);
int mgr_next(void*, MgRecord*, int, int64_t);
}

int main(int argc, char** argv) {
  const int W = 4, N = argc > 1 ? atoi(argv[1]) : 3000;
  char name[64];
  snprintf(name, sizeof name, "/mg_tsan_%d", static_cast<int>(getpid()));
  void* drv = mgr_create(name, W, 256, 4096);
  if (!drv) return 2;
  mgr_start_pump(drv);
  std::atomic<int> errors{0};
  std::vector<std::thread> workers;
  for (int w = 0; w < W; ++w) {
    workers.emplace_back([&, w] {
      void* h = mgr_attach(name);
      if (!h) { errors++; return; }
      mgr_register(h, w, 0, 1000 + w);
      uint64_t tid = 0;
      uint32_t len = 0;
      char buf[256];
      if (mgr_wait_assign(h, w, 5000000, &tid, buf, sizeof buf, &len) != 1 || tid != 0xA000u + w) errors++;
      for (int i = 0; i < N; ++i) {
        while (mgr_post(h, w, MG_REC_METRIC, tid, i, 0.5 * i, 0) != 0) usleep(10);  // ring full: back off
        if ((i & 63) == 0) mgr_heartbeat(h, w);
        (void)mgr_check_stop(h, w, tid);
      }
      const char blob[] = "final-result";
      if (mgr_post_blob(h, w, MG_REC_FINAL, tid, N, 1.0, blob, sizeof blob, 5'000'000) != 0) errors++;
      // second trial: parked by the driver thread, published by the "scheduler" thread (what K9 does from the GPU)
      if (mgr_wait_assign(h, w, 5000000, &tid, buf, sizeof buf, &len) != 1 || tid != 0xB000u + w ||
          std::strcmp(buf, "staged-params") != 0)
        errors++;
      for (int i = 0; i < 64; ++i)
        while (mgr_post(h, w, MG_REC_METRIC, tid, i, 0.25 * i, 0) != 0) usleep(10);
      if (mgr_post_blob(h, w, MG_REC_FINAL, tid, 64, 2.0, blob, sizeof blob, 5'000'000) != 0) errors++;
      mgr_close(h, 0);
    });
  }
  std::atomic<bool> sched_run{true};
  std::thread scheduler([&] {   // publishes whatever is parked, as soon as the worker has picked up its previous trial
    while (sched_run.load()) {
      for (int w = 0; w < W; ++w) (void)mgr_publish_staged(drv, w);
      usleep(20);
    }
  });
  std::vector<int64_t> next_step(W, 0);
  std::vector<int> finals(W, 0);
  int regs = 0, done = 0;
  MgRecord recs[128];
  char blob[4096];
  const uint64_t deadline_iters = 4'000'000;
  for (uint64_t it = 0; it < deadline_iters && done < W; ++it) {
    int n = mgr_next(drv, recs, 128, 2000);
    for (int i = 0; i < n; ++i) {
      const MgRecord& r = recs[i];
      const uint32_t w = r.worker;
      if (w >= static_cast<uint32_t>(W)) { errors++; continue; }
      if (r.type == MG_REC_REG) {
        ++regs;
        const char payload[] = "params";
        while (mgr_assign(drv, w, 0xA000u + w, payload, sizeof payload) == -2) usleep(10);
        const char staged[] = "staged-params";
        while (mgr_stage(drv, w, 0xB000u + w, staged, sizeof staged, 1) == -2) usleep(10);
      } else if (r.type == MG_REC_METRIC) {
        if (r.trial_id == 0xA000u + w) {
          if (r.step != next_step[w] || r.value != 0.5 * r.step) errors++;
          next_step[w] = r.step + 1;
          if (r.step == N / 2) mgr_stop_trial(drv, w, 0xA000u + w);
        }
      } else if (r.type == MG_REC_FINAL) {
        int k = mgr_read_blob(drv, w, blob, sizeof blob);
        if (k <= 0 || std::strcmp(blob, "final-result") != 0) errors++;
        finals[w]++;
        if (finals[w] == 2) ++done;
        // a third trial gets parked while the second one may still be copied out of the staged area: mgr_stage has to
        // refuse until the worker acknowledged it (-2), never overwrite under the reader
        if (finals[w] == 1) {
          const char late[] = "never-run";
          for (int tries = 0; tries < 200 && mgr_stage(drv, w, 0xC000u + w, late, sizeof late, 1) == -2; ++tries) usleep(10);
        }
      }
    }
  }
  for (auto& t : workers) t.join();
  sched_run.store(false);
  scheduler.join();
  for (int w = 0; w < W; ++w)
    if (next_step[w] != N || finals[w] != 2) errors++;
  mgr_close(drv, 1);
  if (regs != W) errors++;
  printf("RING_STRESS %s errors=%d\n", errors.load() == 0 ? "OK" : "FAILED", errors.load());
  return errors.load() == 0 ? 0 : 1;
}
