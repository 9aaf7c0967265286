// group.hpp -- host-side plumbing of the multi-GPU data plane BEHIND the C ABI ("library owns ... NCCL
// communicators"; round-1 review, task 2).  No CUDA in this file: it is compiled by nvcc into the library and by g++ into
// tests/native/group_host.cpp (CPU tests of the partition, the worker pool and the barrier).
//
// The reference is ONE process whose wake loops call vind_onNwake_byRotor target by target (main.f90:814-841 ->
// libCommon.f90:114-171).  Targets are independent (libCommon.f90:132-139 is a parallel loop over them), so the library
// shards them over the GPUs of the box without the driver knowing:
//   * vlc_create_multi(n, devices): one LEADER context + n-1 member contexts, one persistent worker thread per member.
//     Every entry point that changes state is replicated on all members (each member holds the whole wake: the O(N)
//     mutators run redundantly, as SURVEY 8e prescribes); every sweep takes the member's contiguous slice of the target
//     list against all sources; the wake sweep of the resident path all-gathers the velocity slices once per predictor
//     and once per corrector stage (NCCL over NVLink; peer copies when NCCL cannot serve the device list).
//   * vlc_comm_init_rank(ctx, world, rank, id): the same sharding for one PROCESS per GPU (torchrun / MPI launches).
#pragma once

#include <dlfcn.h>

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace vlc {
namespace grp {

// Contiguous, equal-sized slices of a list of M items: slice r = [lo, hi), every slice `per` long except the last ones
// (possibly shorter or empty); per*world >= M is what an in-place all-gather of equal counts needs.
struct Shard {
  long long per = 0, lo = 0, hi = 0;
  long long count() const { return hi - lo; }
};
inline Shard shard_range(long long M, int world, int rank) {
  Shard s;
  if (world < 1) world = 1;
  if (M < 0) M = 0;
  s.per = (M + world - 1) / world;
  s.lo = (long long)rank * s.per;
  if (s.lo > M) s.lo = M;
  s.hi = s.lo + s.per;
  if (s.hi > M) s.hi = M;
  return s;
}

// Reusable barrier for the member threads of one group (C++17: no std::barrier).
class Barrier {
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    const unsigned long gen = gen_;
    if (++arrived_ == n_) {
      arrived_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen_ != gen; });
    }
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, arrived_ = 0;
  unsigned long gen_ = 0;
};

// n members: member 0 runs on the calling thread, members 1..n-1 on persistent worker threads (one per GPU, so that the
// launches of a replicated call are issued to all devices at once instead of one device after the other).
class Workers {
 public:
  explicit Workers(int n) : n_(n) {
    results.assign((size_t)n, 0);
    for (int k = 1; k < n_; ++k) threads_.emplace_back([this, k] { loop(k); });
  }
  ~Workers() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int size() const { return n_; }
  // fn(k) for k = 0..n-1 concurrently; returns the first non-zero result in member order (results[] has them all)
  int run(const std::function<int(int)>& fn) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &fn;
      pending_ = n_ - 1;
      ++gen_;
    }
    cv_.notify_all();
    results[0] = fn(0);
    {
      std::unique_lock<std::mutex> lk(mu_);
      done_.wait(lk, [&] { return pending_ == 0; });
      job_ = nullptr;
    }
    for (int k = 0; k < n_; ++k)
      if (results[k]) return results[k];
    return 0;
  }
  std::vector<int> results;

 private:
  void loop(int k) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<int(int)>* job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        job = job_;
      }
      const int r = (*job)(k);
      {
        std::lock_guard<std::mutex> lk(mu_);
        results[k] = r;
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  int n_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<int(int)>* job_ = nullptr;
  int pending_ = 0;
  unsigned long gen_ = 0;
  bool stop_ = false;
  std::vector<std::thread> threads_;
};

// ---- NCCL, bound at run time ----------------------------------------------------------------------------------------
// dlopen instead of -lnccl: a process that has imported torch already holds torch's libnccl.so.2 (2.28), a plain C / Fortran
// driver gets the system's (2.27); binding late takes whichever is loaded and keeps single-GPU users free of the dependency.
// Only entry points whose signatures have been stable since NCCL 2.0 are used.
struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
constexpr int kNcclFloat64 = 8;  // ncclDataType_t: ncclFloat64 / ncclDouble
constexpr int kNcclSum = 0;      // ncclRedOp_t: ncclSum

struct Nccl {
  bool ok = false;
  std::string why;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, void* /*cudaStream_t*/) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int /*ncclRedOp_t*/, NcclComm, void* /*cudaStream_t*/) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  static Nccl& get() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] { n.load(); });
    return n;
  }

 private:
  void load() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process (torch's)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      why = std::string("libnccl.so.2 not loadable: ") + (dlerror() ? dlerror() : "?");
      return;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(h, name);
      if (!p && why.empty()) why = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
    GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    ok = why.empty();
  }
};

}  // namespace grp
}  // namespace vlc
