// TEST INFRASTRUCTURE: host build of the product's multi-GPU plumbing (volcanor_b200/csrc/group.hpp: target partition,
// worker pool, barrier) driven WITHOUT a GPU: n "members" play the launch sequence of vlc_wake_sweep on plain host
// arrays -- each fills its slice, they exchange the slices the way allgather_slots' peer path does (barrier, pull the
// other slots, barrier), and every member must end up with the complete, identical list.  extern "C" for ctypes.
#include <cstring>

#include "../../volcanor_b200/csrc/group.hpp"

using vlc::grp::Barrier;
using vlc::grp::Shard;
using vlc::grp::shard_range;
using vlc::grp::Workers;

extern "C" {

// out[3*rank..] = per, lo, hi
void grp_shard(long long M, int world, long long* out) {
  for (int r = 0; r < world; ++r) {
    const Shard s = shard_range(M, world, r);
    out[3 * r] = s.per;
    out[3 * r + 1] = s.lo;
    out[3 * r + 2] = s.hi;
  }
}

// `rounds` emulated wake sweeps of M targets on n members.  Member k writes value f(round, target) into its slice of its
// own buffer, then the slices are exchanged.  Returns 0 when every member holds f(round, t) for all t after every round,
// and the results of run() are reported in member order (member `fail_member` returns `fail_code` in round 0).
int grp_emulate_wake_sweeps(int n, long long M, int rounds, int fail_member, int fail_code, int* first_error) {
  Workers w(n);
  Barrier bar(n);
  std::vector<std::vector<double>> buf(n);
  const long long per = shard_range(M, n, 0).per;
  for (auto& b : buf) b.assign((size_t)(per * n), -1.0);
  int bad = 0;
  std::mutex mu;
  *first_error = 0;
  for (int round = 0; round < rounds; ++round) {
    const int rc = w.run([&](int k) -> int {
      const Shard s = shard_range(M, n, k);
      for (long long t = s.lo; t < s.hi; ++t) buf[k][(size_t)t] = 1000.0 * round + (double)t;  // the member's sweep
      bar.wait();                                                                                 // all slots complete
      for (int o = 0; o < n; ++o)
        if (o != k) std::memcpy(&buf[k][(size_t)(o * per)], &buf[o][(size_t)(o * per)], sizeof(double) * (size_t)per);
      bar.wait();                                                                                 // all slots read
      for (long long t = 0; t < M; ++t)
        if (buf[k][(size_t)t] != 1000.0 * round + (double)t) {
          std::lock_guard<std::mutex> lk(mu);
          ++bad;
          break;
        }
      return (round == 0 && k == fail_member) ? fail_code : 0;
    });
    if (round == 0) *first_error = rc;
  }
  return bad;
}

// how many distinct threads served members 1..n-1 over `calls` runs (persistent workers: exactly n-1), and whether
// member 0 ran on the caller's thread
int grp_worker_threads(int n, int calls, int* member0_on_caller) {
  Workers w(n);
  std::vector<std::thread::id> ids((size_t)n);
  std::vector<std::vector<std::thread::id>> seen((size_t)n);
  const std::thread::id me = std::this_thread::get_id();
  *member0_on_caller = 1;
  for (int c = 0; c < calls; ++c) {
    w.run([&](int k) -> int {
      seen[(size_t)k].push_back(std::this_thread::get_id());
      return 0;
    });
  }
  int distinct = 0;
  for (int k = 0; k < n; ++k) {
    bool same = true;
    for (auto& id : seen[(size_t)k]) same = same && id == seen[(size_t)k][0];
    if (k == 0) {
      if (!same || seen[0][0] != me) *member0_on_caller = 0;
    } else if (same && seen[(size_t)k][0] != me) {
      ++distinct;
    }
  }
  return distinct;
}

// 1 if libnccl can be bound at run time (what vlc_create_multi / vlc_comm_init_rank use), with its version
int grp_nccl_available(int* version) {
  vlc::grp::Nccl& n = vlc::grp::Nccl::get();
  *version = 0;
  if (n.ok && n.GetVersion) n.GetVersion(version);
  return n.ok ? 1 : 0;
}
}
