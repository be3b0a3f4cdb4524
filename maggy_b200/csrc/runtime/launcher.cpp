// W1 — native per-GPU worker-pool launcher.
//
// Stands in for the Spark executor layer of the reference (JVM executors started by
// `sc.parallelize(range(n), n).foreachPartition(executor_fn)`, maggy/core/experiment_driver/
// spark_driver.py:136-145): one long-running OS process per GPU (or per GPU *gang* for multi-GPU
// trials), pinned with CUDA_VISIBLE_DEVICES and a CPU-affinity mask, supervised (exit status, signal,
// respawn count) by the driver. posix_spawn keeps the driver's (possibly large, multi-threaded)
// address space out of the child.
#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sched.h>
#include <spawn.h>
#include <sys/resource.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

extern char** environ;

namespace {

struct Slot {
  pid_t pid = -1;
  int exit_code = 0;     // valid when !alive
  int term_signal = 0;
  bool alive = false;
  int spawns = 0;
  std::string gpus;
};

struct Pool {
  std::vector<Slot> slots;
  std::mutex mu;
};

std::vector<std::string> split0(const char* blob) {
  // blob = "a\0b\0c\0\0"
  std::vector<std::string> out;
  const char* p = blob;
  while (p && *p) {
    out.emplace_back(p);
    p += out.back().size() + 1;
  }
  return out;
}

}  // namespace

extern "C" {

void* mgl_create(int num_slots) {
  Pool* p = new Pool();
  p->slots.resize(num_slots > 0 ? num_slots : 1);
  return p;
}

void mgl_destroy(void* pv) { delete static_cast<Pool*>(pv); }

// argv0 / envs are NUL-separated, double-NUL-terminated string lists. `gpus` becomes
// CUDA_VISIBLE_DEVICES ("" leaves it untouched, "-1" hides all GPUs). cpu_first/cpu_count select an
// affinity range (cpu_count <= 0: inherit). stdout/stderr are appended to log_path when non-empty.
int mgl_spawn(void* pv, int slot, const char* argv_blob, const char* env_blob, const char* gpus,
              int cpu_first, int cpu_count, const char* log_path) {
  Pool* p = static_cast<Pool*>(pv);
  if (slot < 0 || slot >= static_cast<int>(p->slots.size())) return -1;
  std::vector<std::string> args = split0(argv_blob);
  if (args.empty()) return -2;
  std::vector<std::string> extra = split0(env_blob);
  if (gpus && *gpus) extra.push_back(std::string("CUDA_VISIBLE_DEVICES=") + (std::strcmp(gpus, "-1") == 0 ? "" : gpus));

  std::vector<std::string> env;
  for (char** e = environ; *e; ++e) {
    std::string kv(*e);
    std::string key = kv.substr(0, kv.find('='));
    bool overridden = false;
    for (auto& x : extra)
      if (x.compare(0, key.size() + 1, key + "=") == 0) overridden = true;
    if (!overridden) env.push_back(kv);
  }
  for (auto& x : extra) env.push_back(x);

  std::vector<char*> argv, envp;
  for (auto& a : args) argv.push_back(const_cast<char*>(a.c_str()));
  argv.push_back(nullptr);
  for (auto& e : env) envp.push_back(const_cast<char*>(e.c_str()));
  envp.push_back(nullptr);

  posix_spawn_file_actions_t fa;
  posix_spawn_file_actions_init(&fa);
  if (log_path && *log_path) {
    posix_spawn_file_actions_addopen(&fa, 1, log_path, O_WRONLY | O_CREAT | O_APPEND, 0644);
    posix_spawn_file_actions_adddup2(&fa, 1, 2);
  }
  posix_spawnattr_t at;
  posix_spawnattr_init(&at);
  // own process group: the driver can kill a whole (multi-process) trial gang with one signal
  posix_spawnattr_setflags(&at, POSIX_SPAWN_SETPGROUP);
  posix_spawnattr_setpgroup(&at, 0);

  pid_t pid = -1;
  int rc = posix_spawnp(&pid, argv[0], &fa, &at, argv.data(), envp.data());
  posix_spawn_file_actions_destroy(&fa);
  posix_spawnattr_destroy(&at);
  if (rc != 0) return -100 - rc;

  if (cpu_count > 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    for (int c = 0; c < cpu_count; ++c) CPU_SET((cpu_first + c) % (ncpu > 0 ? ncpu : 1), &set);
    sched_setaffinity(pid, sizeof(set), &set);  // best effort; children inherit
  }
  std::lock_guard<std::mutex> lk(p->mu);
  Slot& s = p->slots[slot];
  s.pid = pid;
  s.alive = true;
  s.exit_code = 0;
  s.term_signal = 0;
  s.spawns += 1;
  s.gpus = gpus ? gpus : "";
  return static_cast<int>(pid);
}

// 1 alive, 0 exited (exit code / signal available), -1 never started
int mgl_poll(void* pv, int slot) {
  Pool* p = static_cast<Pool*>(pv);
  std::lock_guard<std::mutex> lk(p->mu);
  if (slot < 0 || slot >= static_cast<int>(p->slots.size())) return -1;
  Slot& s = p->slots[slot];
  if (s.pid <= 0) return -1;
  if (!s.alive) return 0;
  int status = 0;
  pid_t r = waitpid(s.pid, &status, WNOHANG);
  if (r == 0) return 1;
  s.alive = false;
  if (r < 0) {
    s.exit_code = -errno;
  } else if (WIFEXITED(status)) {
    s.exit_code = WEXITSTATUS(status);
  } else if (WIFSIGNALED(status)) {
    s.term_signal = WTERMSIG(status);
    s.exit_code = 128 + s.term_signal;
  }
  return 0;
}

int mgl_wait(void* pv, int slot, int timeout_ms) {
  int waited = 0;
  for (;;) {
    int st = mgl_poll(pv, slot);
    if (st != 1) return st;
    if (timeout_ms >= 0 && waited >= timeout_ms) return 1;
    usleep(2000);
    waited += 2;
  }
}

int mgl_pid(void* pv, int slot) { return static_cast<Pool*>(pv)->slots[slot].pid; }
int mgl_exit_code(void* pv, int slot) { return static_cast<Pool*>(pv)->slots[slot].exit_code; }
int mgl_term_signal(void* pv, int slot) { return static_cast<Pool*>(pv)->slots[slot].term_signal; }
int mgl_spawn_count(void* pv, int slot) { return static_cast<Pool*>(pv)->slots[slot].spawns; }

// signal the slot's whole process group (the worker and anything it started)
int mgl_kill(void* pv, int slot, int sig) {
  Pool* p = static_cast<Pool*>(pv);
  std::lock_guard<std::mutex> lk(p->mu);
  Slot& s = p->slots[slot];
  if (s.pid <= 0 || !s.alive) return 0;
  if (kill(-s.pid, sig) != 0) return kill(s.pid, sig);
  return 0;
}

}  // extern "C"
