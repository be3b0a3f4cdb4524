// Asynchronous Successive Halving rung bookkeeping + promotion decision, shared by the persistent scheduler kernel (K9,
// csrc/ctrl/ring_device.cu) and the host library (csrc/ctrl/ring.cpp exposes it to the CPU tests).
//
// Reference behaviour: maggy/optimizer/asha.py:95-147 — scan the rungs top-down for a finished trial that ranks inside
// the top 1/reduction_factor of its rung and has not been promoted yet, else start a fresh configuration in rung 0.
// Here the finished trials of a rung are kept SORTED best-first (insertion on every FINAL), so the decision is a scan of
// the first `count / rf` entries. Ranking honours the experiment's direction, like maggy_b200.optimizer.Asha.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD static inline
#endif

#define MG_ASHA_MAX_RUNGS 8

enum MgDecision : uint32_t {
  MG_DEC_NONE = 0,
  MG_DEC_NEW = 1,      // no promotable configuration: start a fresh one in rung 0
  MG_DEC_PROMOTE = 2,  // promote `parent` from rung `rung` to rung `rung + 1`
  MG_DEC_NEW_PUBLISHED = 3,  // NEW, and the scheduler kernel already handed the worker its prefetched (staged) trial
};

struct MgAshaEntry {  // 32 bytes
  uint64_t trial_id;
  double metric;
  uint32_t promoted;
  uint32_t pad[3];
};

struct MgAsha {
  uint32_t rf, max_rung, cap;  // cap: entries per rung
  int32_t direction;           // +1 maximise, -1 minimise
  uint32_t count[MG_ASHA_MAX_RUNGS];     // finished trials recorded per rung
  uint32_t promoted[MG_ASHA_MAX_RUNGS];  // of which promoted
  uint32_t dropped;                      // finals that did not fit (cap exceeded)
  uint32_t decisions;
  uint32_t closed;                       // a configuration was promoted into the top rung: no further decisions
  uint32_t pad;
  // MgAshaEntry entries[MG_ASHA_MAX_RUNGS][cap] follows
};

static inline uint64_t mg_asha_bytes(uint32_t cap) {
  return sizeof(MgAsha) + static_cast<uint64_t>(MG_ASHA_MAX_RUNGS) * cap * sizeof(MgAshaEntry);
}

MG_HD MgAshaEntry* mg_asha_rung(MgAsha* a, uint32_t rung) {
  return reinterpret_cast<MgAshaEntry*>(reinterpret_cast<uint8_t*>(a) + sizeof(MgAsha)) +
         static_cast<uint64_t>(rung) * a->cap;
}

// a finished trial of `rung`: insert behind every entry that is at least as good (ties keep arrival order)
MG_HD void mg_asha_record(MgAsha* a, uint32_t rung, uint64_t trial_id, double metric) {
  if (rung >= MG_ASHA_MAX_RUNGS) return;
  const uint32_t n = a->count[rung];
  if (n >= a->cap) {
    a->dropped += 1;
    return;
  }
  MgAshaEntry* e = mg_asha_rung(a, rung);
  uint32_t pos = n;
  while (pos > 0) {
    const double other = e[pos - 1].metric;
    const bool better = a->direction >= 0 ? (metric > other) : (metric < other);
    if (!better) break;
    e[pos] = e[pos - 1];
    --pos;
  }
  e[pos].trial_id = trial_id;
  e[pos].metric = metric;
  e[pos].promoted = 0;
  a->count[rung] = n + 1;
}

// the promotion scan of asha.py:95-147. Returns the decision kind; for PROMOTE fills parent / rung and marks it.
MG_HD uint32_t mg_asha_decide(MgAsha* a, uint64_t* parent, uint32_t* rung_out) {
  a->decisions += 1;
  *parent = 0;
  *rung_out = 0;
  if (a->closed) return MG_DEC_NONE;  // asha.py:90-92: the experiment winds down once the top rung is populated
  for (int k = static_cast<int>(a->max_rung) - 1; k >= 0; --k) {
    const uint32_t quota = a->count[k] / a->rf;
    if (quota <= a->promoted[k]) continue;
    MgAshaEntry* e = mg_asha_rung(a, static_cast<uint32_t>(k));
    for (uint32_t i = 0; i < quota; ++i) {
      if (!e[i].promoted) {
        e[i].promoted = 1;
        a->promoted[k] += 1;
        *parent = e[i].trial_id;
        *rung_out = static_cast<uint32_t>(k);
        if (static_cast<uint32_t>(k) + 1 >= a->max_rung) a->closed = 1;
        return MG_DEC_PROMOTE;
      }
    }
  }
  return MG_DEC_NEW;
}
