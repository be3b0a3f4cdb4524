// maggy_b200 control plane: shared-memory layout of the driver <-> worker channel.
//
// Replaces the reference's hand-rolled TCP RPC (maggy/core/rpc.py: length-prefixed cloudpickle dicts,
// 1 s heartbeat, 1 s GET poll) with per-worker single-producer/single-consumer rings of 64-byte
// records in one shared segment. The SAME layout is written by
//   * host code in the worker process (csrc/ctrl/ring.cpp, C++ atomics), and
//   * device code (csrc/ctrl/ring_device.cu): the segment is cudaHostRegister'ed (mapped), so
//     `reporter.broadcast(loss_tensor)` is one st.release.sys from the GPU — no host sync, no RPC;
// and read by the Python driver / the native policy thread / the persistent scheduler kernel.
//
// Message mapping (reference protocol table):
//   REG   -> MG_REC_REG record          METRIC -> MG_REC_METRIC record (host or device ring)
//   FINAL -> MG_REC_FINAL (+blob)       LOG    -> MG_REC_LOG (+blob)
//   GET/TRIAL -> driver writes the assign slot (+payload) and bumps assign_seq; worker futex-waits
//   STOP  -> driver/policy writes ctl.stop_trial_id; GSTOP -> header.gstop
//   BLACK -> heartbeat timeout / re-REG with a trial still assigned (handled by the driver)
//   prefetched hand-off -> driver stages the next trial (mgr_stage); K9 publishes it on the worker's FINAL
//   next-trial decision (ASHA rung scan) -> K9 writes ctl.dec_* when it sees a tagged FINAL (mg_asha.h)
#pragma once
#include <stdint.h>

#define MG_RING_MAGIC 0x4d47523162323030ull /* "MGR1b200" */
#define MG_RING_VERSION 5u

enum MgRecType : uint16_t {
  MG_REC_NONE = 0,
  MG_REC_REG = 1,
  MG_REC_METRIC = 2,
  MG_REC_FINAL = 3,
  MG_REC_LOG = 4,
  MG_REC_IDLE = 5,
  MG_REC_ERROR = 6,
  MG_REC_HEARTBEAT = 7,
  MG_REC_EXIT = 8,
};

// 64-byte record. `seq` is written LAST with release semantics: slot i of lap k holds seq == k*slots+i+1.
struct alignas(64) MgRecord {
  uint64_t seq;
  uint16_t type;
  uint16_t worker;
  uint32_t aux;        // blob length for FINAL/LOG, attempt for REG, user tag otherwise
  uint64_t trial_id;   // first 64 bits of the trial hash (Trial.trial_id is 16 hex chars = 64 bits)
  int64_t step;
  double value;
  uint64_t t_ns;       // producer timestamp (CLOCK_MONOTONIC on host, %globaltimer on device)
  uint64_t reserved[2];
};

struct alignas(64) MgRing {
  // producer side
  alignas(64) uint64_t head;     // records produced
  uint64_t dropped;              // records dropped because the ring was full (device producer only)
  // consumer side
  alignas(64) uint64_t tail;     // records consumed
};

// per-worker control block (driver -> worker words first, then worker -> driver words)
struct alignas(64) MgWorkerCtl {
  // ---- written by the driver / policy engine
  alignas(64) uint32_t assign_seq;   // bumped (futex-woken) when a new trial sits in the assign slot
  uint32_t assign_len;               // payload bytes
  uint64_t assign_trial_id;
  uint64_t stop_trial_id;            // != 0: the trial with this id must stop at its next broadcast
  uint32_t stop_seq;                 // bumped with every stop request
  uint32_t assign_tag;               // opaque tag of the assigned trial (ASHA: rung + 1), echoed in its FINAL record
  uint64_t stop_t_ns;                // when the stop decision was taken (latency accounting)
  // prefetched next trial: the driver parks a trial in the worker's STAGED payload area; whoever sees the worker's
  // FINAL first (the scheduler kernel K9, else the driver) publishes it by copying the staged_* words over the
  // assign_* words and storing assign_seq = staged_seq. Valid iff staged_seq == assign_seq + 1.
  uint32_t assign_src;               // 0: payload sits in the assign area, 1: in the staged area
  uint32_t staged_seq;
  uint32_t staged_len;
  uint32_t staged_tag;
  uint64_t staged_trial_id;
  // ---- written by the worker
  alignas(64) uint64_t heartbeat_ns; // last sign of life
  uint32_t state;                    // MgWorkerState
  uint32_t attempt;
  uint32_t ack_seq;                  // last assign_seq picked up
  uint32_t blob_seq;                 // bumped when the worker wrote a blob; blob_ack follows it
  uint32_t blob_len;
  uint32_t pid;
  // ---- written by the driver
  alignas(64) uint32_t blob_ack;     // driver consumed blob #blob_ack
  // ---- written by the persistent scheduler kernel (K9): what the worker should run after its n-th tagged FINAL
  alignas(64) uint32_t dec_seq;      // decisions taken for this worker so far (release-stored last)
  uint32_t dec_kind;                 // MgDecision (mg_asha.h)
  uint32_t dec_rung;                 // PROMOTE: rung the parent finished in
  uint32_t dec_pad;
  uint64_t dec_parent;               // PROMOTE: trial id of the configuration to promote
  uint64_t dec_t_ns;                 // when K9 took the decision
};

enum MgWorkerState : uint32_t {
  MG_W_EMPTY = 0,
  MG_W_REGISTERED = 1,
  MG_W_IDLE = 2,
  MG_W_RUNNING = 3,
  MG_W_DONE = 4,
  MG_W_DEAD = 5,
};

// device-decidable early-stop policy state (median stopping rule), shared with the policy engine
struct alignas(64) MgPolicy {
  uint32_t enabled;        // 0 = off, 1 = median rule
  int32_t direction;       // +1 = maximise, -1 = minimise
  uint32_t es_min;         // need > es_min finalized trials before stopping anything
  uint32_t es_interval;    // check when step % es_interval == 0
  uint32_t max_steps;      // columns of the running-average table
  uint32_t num_final;      // finalized trials recorded so far
  uint32_t cap_final;      // rows of the table
  uint32_t pad;
};

struct alignas(64) MgHeader {
  uint64_t magic;
  uint32_t version;
  uint32_t num_workers;
  uint32_t ring_slots;       // power of two
  uint32_t payload_bytes;    // per worker, each direction
  uint64_t total_bytes;
  uint64_t off_ctl;          // MgWorkerCtl[num_workers]
  uint64_t off_host_ring;    // per worker: MgRing + MgRecord[ring_slots]
  uint64_t off_dev_ring;     // same, produced by the GPU
  uint64_t off_assign;       // per worker payload_bytes (driver -> worker)
  uint64_t off_blob;         // per worker payload_bytes (worker -> driver)
  uint64_t off_policy;       // MgPolicy + float table[cap_final][max_steps] + per-worker running state
  uint64_t off_staged;       // per worker payload_bytes: the prefetched next trial (driver -> worker)
  uint64_t ring_stride;      // bytes between consecutive workers' rings
  alignas(64) uint32_t doorbell;   // bumped + futex-woken by host producers
  uint32_t gstop;                  // global stop
  uint32_t driver_alive;
  // ---- device mailbox (GPU memory owned by the driver process, reached by the workers' GPUs over NVLink)
  alignas(64) uint32_t mailbox_valid;   // 1 once the driver published the handle below
  uint32_t mailbox_slots;               // records per worker ring (power of two)
  uint64_t mailbox_bytes;
  uint8_t mailbox_handle[64];           // cudaIpcMemHandle_t of the allocation
};

// Device mailbox: the peer-visible half of the control plane. It lives in the HBM of the GPU that runs the
// persistent scheduler kernel (K9); every worker maps it through CUDA IPC, so `reporter.broadcast(tensor)` is a
// 64-byte store over NVLink into this ring and K9 polls its own local memory instead of host memory across PCIe.
struct alignas(64) MgMailboxCtl {
  uint64_t head;            // records produced by the worker's K8 so far
  uint64_t stop_trial_id;   // written by K9: the trial with this id must stop (mirrors MgWorkerCtl.stop_trial_id)
  uint32_t stop_seq;
  uint32_t pad;
  uint64_t stop_t_ns;
  uint64_t reserved[3];
};

static inline uint64_t mg_mailbox_bytes(uint32_t num_workers, uint32_t slots) {
  return (uint64_t)num_workers * (sizeof(MgMailboxCtl) + (uint64_t)slots * sizeof(MgRecord));
}
// byte offset of worker w's ring inside the mailbox (the MgMailboxCtl array comes first)
static inline uint64_t mg_mailbox_ring_off(uint32_t num_workers, uint32_t slots, uint32_t w) {
  return (uint64_t)num_workers * sizeof(MgMailboxCtl) + (uint64_t)w * slots * sizeof(MgRecord);
}

static inline uint64_t mg_ring_bytes(uint32_t slots) {
  return sizeof(MgRing) + (uint64_t)slots * sizeof(MgRecord);
}
