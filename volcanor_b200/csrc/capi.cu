// capi.cu -- implementation of the C ABI in include/volcanor_b200.h.
//
// Owns: device buffers (packed source sets, reference-layout rotor copies, LU factors), the CUDA
// stream, the cuSOLVER handle.  The caller owns every host array.  No CPU compute path exists
// here: every entry point either launches sm_100a kernels or fails with an error.
#include "../../include/volcanor_b200.h"

#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "bs_sweep.cuh"
#include "bs_lattice.cuh"
#include "pack.cuh"
#include "wake_state.cuh"
#include "wake_records.cuh"
#include "cp_stage.cuh"
#include "group.hpp"
#include "plan.hpp"

namespace {

constexpr int kThreads = 128;  // threads per sweep CTA
#ifndef VLC_LAT_THREADS
#define VLC_LAT_THREADS 128
#endif
constexpr int kLatThreads = VLC_LAT_THREADS;  // threads per lattice-kernel CTA
constexpr int kTile = 128;     // filaments per shared-memory tile (12 KB)
constexpr int kStages = 3;     // TMA ring depth
// Lattice kernel shape (vlc_set_lattice_tuning): strip width W in 1..4, targets per thread T in 1..3; 0 = automatic:
// W minimises the measured cost per ring (profiles/r01h_wt_sweep.md: W=1,T=3 1.000; W=2,T=2 0.919; W=3,T=2 0.932;
// W=4,T=2 0.885, T=1 0.887) times the padding of the last strip, ceil(ns/W)*W/ns; T is the best measured one for that W
// (sweeps with at most one CTA of targets use T=1, sweep_shared).
constexpr double kLatCost[5] = {0.0, 1.000, 0.919, 0.932, 0.885};
constexpr int kLatBestT[5] = {0, 3, 2, 2, 2};

std::string g_create_error;

struct DevBuf {
  double* p = nullptr;
  size_t cap = 0;  // doubles
};

struct SourceSet {
  DevBuf rec;
  long long n = 0;      // logical filaments (reference enumeration)
  long long n_pad = 0;  // padded to kTile
  // shared-node form of the same sources (tier 3 lattices, bs_lattice.cuh): strip records + flat remainder
  DevBuf lat;
  long long n_lat = 0, n_lat_pad = 0;  // strip records, padded to lat_tile(lat_W)
  int lat_W = 1;                       // strip width the records were packed with
  // optional tail strips of another width (tier 2: ns = 4*floor(ns/4) + tail): same record kind, own buffer
  DevBuf lat2;
  long long n_lat2 = 0, n_lat2_pad = 0;
  int lat2_W = 0;
  DevBuf rem;
  long long n_rem = 0, n_rem_pad = 0;  // filaments no strip covers (last column, horseshoe, far chain)
  int* d_unmergeable = nullptr;        // device flag raised by the pack kernels
  bool has_shared = false;
  long long n_rings_main = 0;          // vortex rings covered by the strips of `lat` (4 reference filaments each)
  bool may_dual = false;               // the flag can take the value 2 (records classified by check_rings_kernel: tier 2)
};

// one dominant-kernel launch of a sweep, for vlc_sweep_stats
struct SweepStat {
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;  // before / after the dominant kernel, after the whole sweep
  int kind = 0;  // 0 = bs_lattice_kernel, 1 = bs_sweep_kernel
  double pairs = 0.0, instr = 0.0;            // of the dominant kernel
  double pairs_all = 0.0, instr_all = 0.0;    // of every kernel of the sweep (+ tail strips, flat remainder)
};

struct Rotor {
  bool defined = false;
  int nb = 0, nc = 0, ns = 0, nNwake = 0, nFwake = 0, surfaceType = 1;
  int rowNear = 1, rowFar = 1;
  // reference-layout copies (all blades, blade-major)
  DevBuf wiP, waN[2], waF[2], wapF[2];
  bool have_pf[2] = {false, false};
  DevBuf pfHelix[2], pfFits;  // prescribed far wake made on the device: (helixPitch, helixRadius) per blade and set; fit scratch
  // packed: [wing | wake] per set (C, P), and bound-vortex set
  SourceSet comb[2];
  long long wing_pad[2] = {0, 0};  // padded wing segment length inside comb[s]
  long long wing_n = 0;
  SourceSet bound;
  // packed set s needs (re)building: 0 = no; 1 = all of it (every `dirty[s] = true`); 2 = only its wing segments -- the
  // wing's records changed (moved wing, new circulation) while the wake records and rows did not (mark_wing_dirty)
  int dirty[2] = {1, 1};
  bool bound_dirty = true;
  SourceSet chord;  // chordwise-vortex set (classdef.f90:1398-1418), same shape as `bound`
  bool chord_dirty = true;
  // rows below these were not refreshed by the last put (only the active rows travel): [set][blade]
  std::vector<int> stale_near[2], stale_far[2];
  // tier 2b (device-resident stepping): rotor_class members the wake mutators read, and the velocity arrays of the
  // convection driver (classdef.f90:285-292): [0] vel, [1] vel1, [2] velPredicted, [3] velStep
  int nbConvect = 0, axisym = 0, duct = 0, suppressFwake = 0, rollupStart = 1, rollupEnd = 1, sgnPositive = 1;
  double apparentViscCoeff = 0.0, decayCoeff = 0.0, initWakeVel = 0.0;
  double shaftAxis[3] = {0.0, 0.0, 1.0}, hubCoords[3] = {0.0, 0.0, 0.0};
  DevBuf velN[4], velF[4];
  DevBuf velNx[2], velFx[2];  // [0] vel2, [1] vel3: histories of fdScheme 4 / 5 (classdef.f90:3733-3824), allocated on first use
  DevBuf waN_alt;  // second buffer of shiftwake (swapped with waN[0])
  DevBuf order2_tmp;
  vlc::AxiT* d_axi = nullptr;
  std::vector<vlc::AxiT> h_axi;
  // tier 2c (collocation-point stage on the device): right-hand side, circulation vector, section frames, loads
  DevBuf rhs, gamvec, sec, loads, loads_scr;  // loads_scr: per-panel scratch of the loads phases (cp_stage.cuh: kScr)
  bool have_rhs = false;
  std::vector<char> have_sec;  // per blade: vlc_rotor_put_sections seen
  // AIC
  int N = 0;
  DevBuf LU, Ainv;  // LU factors (cuSOLVER getrf) and the inverse the reference multiplies by every step (classdef.f90:4178)
  int* d_ipiv = nullptr;
  int* d_info = nullptr;
  bool factored = false;
};

}  // namespace

struct vlc_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  long long mem_bytes = 0;
  int tune_T = 0, tune_nsplit = 0;
  bool shared_nodes = true;  // lattice sources: use the shared-node kernel when the set allows it
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};  // last sweep: before / after the dominant kernel, after the reduce
  cudaStream_t aux = nullptr;                       // low-priority side stream: the flat remainder fills the lattice kernel's tail
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool ev_valid = false;
  int lat_W = 0, lat_T = 0;          // lattice kernel shape (vlc_set_lattice_tuning), 0 = automatic
  int occ_lat[5][4] = {};            // resident CTAs/SM of the lattice kernel [W][T]
  int occ_dual[5] = {};              // ... of its dual form [W] (T = 1)
  bool fast = false;  // rsqrt refinement: false = third order (~1e-16), true = second order (~4e-14)
  long long launches = 0;
  SourceSet sets[VLC_MAX_SETS];
  DevBuf part;     // source-split partials
  DevBuf stage_P;  // host-API staging
  DevBuf stage_V;
  DevBuf scratch;  // packing inputs for host-API set_sources
  DevBuf ws_P, ws_V, ws_acc;  // vlc_wake_sweep: targets of every convected blade, one source rotor's result, the sum
  SourceSet ws_comb[2];       // vlc_wake_sweep: the packed sets of ALL source rotors side by side (one launch per sweep)
  int* ws_flags = nullptr;    // device array of the addresses of the source rotors' mergeability flags (or_flags_kernel)
  std::vector<const int*> ws_flags_host[2];  // what ws_flags holds for set s (re-uploaded only when the list changes)
  DevBuf cp_P, cp_V;          // tier 2c: collocation points of one rotor, one source rotor's velocities there
  unsigned char* d_flag = nullptr;
  size_t flag_cap = 0;
  std::vector<Rotor> rotors;
  cusolverDnHandle_t solver = nullptr;
  vlc::grp::Workers* host_pool = nullptr;  // host_parallel
  // collocation-point sweeps of a group share their source splits from this many pair interactions on (cp_sweep);
  // VLC_RHS_SHARE_MIN_PAIRS overrides (0 = always, negative = never)
  double rhs_share_min = std::getenv("VLC_RHS_SHARE_MIN_PAIRS") ? std::atof(std::getenv("VLC_RHS_SHARE_MIN_PAIRS")) : 2e8;
  DevBuf solver_work;
  int occ[5] = {0, 0, 0, 0, 0};  // resident CTAs/SM of the sweep kernel for T = 1..4
  // multi-GPU data plane (group.hpp): this context's place in the target partition, its NCCL communicator (library-owned),
  // and -- for the members of an in-process group made by vlc_create_multi -- the group
  double* host_P = nullptr;  // pinned scratch for target lists gathered from the caller's records (vlc_vind_on?wake_byRotor)
  size_t host_P_cap = 0;
  static constexpr size_t kStageBytes = (size_t)1 << 20;
  void* stage_h[2] = {nullptr, nullptr};  // pinned staging of small uploads (upload())
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  int stage_next = 0;
  bool stats_on = false;         // vlc_sweep_stats: per-launch events of the dominant kernels
  std::vector<SweepStat> stats;
  size_t stats_n = 0;
  cudaEvent_t user_ev[8] = {};   // vlc_event_record slots
  unsigned char* d_flush = nullptr;  // vlc_l2_flush scratch (256 MiB, allocated on first use)
  int rank = 0, world = 1;
  vlc::grp::NcclComm comm = nullptr;
  struct vlc_group* group = nullptr;
  bool is_leader = false;
};

// One process, several GPUs: members[0] is the context the caller holds (the leader), members[1..] are replicas on the
// other devices, each driven by its own worker thread.
struct vlc_group {
  std::vector<vlc_ctx*> members;
  vlc::grp::Workers* workers = nullptr;
  vlc::grp::Barrier* barrier = nullptr;
  bool nccl = false;
};

namespace {

int fail(vlc_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

// vlc_sweep_stats: an event pair around a dominant-kernel launch (events are created once and reused)
SweepStat* stat_begin(vlc_ctx* c) {
  if (!c->stats_on) return nullptr;
  if (c->stats_n == c->stats.size()) {
    SweepStat st;
    if (cudaEventCreate(&st.e0) != cudaSuccess || cudaEventCreate(&st.e1) != cudaSuccess || cudaEventCreate(&st.e2) != cudaSuccess)
      return nullptr;
    c->stats.push_back(st);
  }
  SweepStat* st = &c->stats[c->stats_n];
  cudaEventRecord(st->e0, c->stream);
  return st;
}
void stat_end(vlc_ctx* c, SweepStat* st, int kind, double pairs, double instr) {
  if (!st) return;
  cudaEventRecord(st->e1, c->stream);
  st->kind = kind;
  st->pairs = st->pairs_all = pairs;
  st->instr = st->instr_all = instr;
}
// after the last kernel of the sweep (the reduce): closes the record
void stat_close(vlc_ctx* c, SweepStat* st, double pairs_extra, double instr_extra) {
  if (!st) return;
  cudaEventRecord(st->e2, c->stream);
  st->pairs_all += pairs_extra;
  st->instr_all += instr_extra;
  c->stats_n++;
}

// true while this thread executes ONE member's share of a replicated call (always on the worker threads): nested entry
// points then act on that member alone
thread_local bool t_in_member = false;

// Run fn on every member of the leader's group, each on its own thread and device.  First failure (in member order) is
// returned with that member's message.
int group_all(vlc_ctx* L, const std::function<int(vlc_ctx*)>& fn) {
  vlc_group* g = L->group;
  const int rc = g->workers->run([&](int k) -> int {
    const bool prev = t_in_member;
    t_in_member = true;
    vlc_ctx* m = g->members[k];
    int r = (cudaSetDevice(m->device) == cudaSuccess) ? fn(m) : fail(m, VLC_ERR_CUDA, "cudaSetDevice failed");
    t_in_member = prev;
    return r;
  });
  if (rc)
    for (size_t k = 1; k < g->members.size(); ++k)
      if (g->workers->results[k] && !g->workers->results[0]) {
        L->err = "[member " + std::to_string(k) + ", device " + std::to_string(g->members[k]->device) + "] " + g->members[k]->err;
        break;
      }
  return rc;
}
// Replicated entry point: on a group's leader, run CALL (an expression in `m`) on every member.
#define VLC_GROUP(c, CALL)                                  \
  if ((c)->group && (c)->is_leader && !t_in_member)         \
  return group_all((c), [&](vlc_ctx* m) -> int { return CALL; })
// Entry points that take DEVICE pointers address one device: not available on a group handle.
#define VLC_NO_GROUP(c)                                                                                           \
  if ((c)->group && !t_in_member)                                                                                 \
  return fail((c), VLC_ERR_STATE, std::string(__func__) + ": device-pointer entry points address ONE device; a context made by " \
                                  "vlc_create_multi spans several -- use the host-pointer / resident entry points")

#define CUDA_OK(c, expr)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail((c), VLC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
  } while (0)

#define CHECK_CTX(c) \
  if (!(c)) return VLC_ERR_ARG

int bind_device(vlc_ctx* c) {
  CUDA_OK(c, cudaSetDevice(c->device));
  return VLC_OK;
}

// Grow-only device buffer.  A first allocation takes what is asked (+1/8); a RE-allocation at least doubles, because
// cudaFree + cudaMalloc in the middle of a run is expensive and erratic (measured on the B200 box: 5 ms .. 2 s per
// event, profiles/r01h_resident_cases.md) -- a wake that grows by one row per time step must not pay it every few steps.
int reserve(vlc_ctx* c, DevBuf& b, size_t doubles) {
  if (doubles <= b.cap) return VLC_OK;
  size_t want = doubles + doubles / 8 + 1024;
  if (b.p) {
    want = std::max(want, 2 * b.cap);
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    CUDA_OK(c, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  CUDA_OK(c, cudaMalloc(&b.p, want * sizeof(double)));
  b.cap = want;
  return VLC_OK;
}

void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

inline long long pad_tile(long long n) { return (n + kTile - 1) / kTile * kTile; }
inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

constexpr size_t kSweepSmem = (size_t)kStages * kTile * vlc::kSrcBytes + kStages * sizeof(uint64_t);

template <int T, int MINB, bool FAST>
int launch_sweep_T(vlc_ctx* c, const double* src, long long n_pad, long long chunk, int nsplit, long long m,
                   const double* dP, double* out, const int* flag, int mask, int want) {
  auto kern = vlc::bs_sweep_kernel<T, kThreads, kTile, kStages, MINB, FAST>;
  dim3 grid(blocks_for(m, kThreads * T), (unsigned)nsplit, 1);
  kern<<<grid, kThreads, kSweepSmem, c->stream>>>(src, chunk, n_pad, dP, m, out, flag, mask, want);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  return VLC_OK;
}

template <int T, int MINB>
int query_occ(vlc_ctx* c, int* out) {
  auto kern = vlc::bs_sweep_kernel<T, kThreads, kTile, kStages, MINB, false>;
  CUDA_OK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem));
  CUDA_OK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kern, kThreads, kSweepSmem));
  return VLC_OK;
}

// Launch shape: T targets per thread and nsplit source splits.
constexpr int kFlatPerTile = 4;                   // a flat chunk is a multiple of a quarter tile (bs_sweep.cuh)
constexpr int kFlatGranule = kTile / kFlatPerTile;
// *unit_out = the records a chunk is a multiple of: the granule for a small sweep or a tuned split, else the tile
void choose_shape(const vlc_ctx* c, long long m, long long n_pad, int* T_out, int* nsplit_out, int* unit_out) {
  *unit_out = kFlatGranule;
  const long long src_tiles = n_pad / kTile;
  int T = c->tune_T;
  if (T < 1 || T > 4) {
    // enough target tiles to fill the machine at T=4?  otherwise fewer targets per thread
    const long long slots4 = (long long)c->sm_count * (c->occ[4] > 0 ? c->occ[4] : 3);
    const long long tiles4 = (m + kThreads * 4 - 1) / (kThreads * 4);
    T = (tiles4 * src_tiles >= 4 * slots4 || tiles4 >= slots4) ? 4 : 2;
    if (m <= kThreads) T = 1;
    if (T == 2) {  // still too small at two targets per thread: one per thread, twice the CTAs
      const long long slots2 = (long long)c->sm_count * (c->occ[2] > 0 ? c->occ[2] : 3);
      const long long tiles2 = (m + kThreads * 2 - 1) / (kThreads * 2);
      if (vlc::plan::small_split(tiles2, src_tiles, slots2, kFlatPerTile) > 0) T = 1;
    }
  }
  int nsplit = c->tune_nsplit;
  if (nsplit < 1) {
    const long long slots = (long long)c->sm_count * (c->occ[T] > 0 ? c->occ[T] : 3);
    const long long tiles = (m + (long long)kThreads * T - 1) / ((long long)kThreads * T);
    const long long cap_by_mem = (long long)((size_t)1 << 31) / (3 * (m > 0 ? m : 1) * 8) + 1;  // <= 2 GiB partials
    int best_s = vlc::plan::small_split(tiles, src_tiles, slots, kFlatPerTile);
    if (best_s == 0) {
      best_s = vlc::plan::wave_split(tiles, src_tiles, slots, cap_by_mem);
      *unit_out = kTile;
    }
    nsplit = best_s > 0 ? best_s : 1;
  }
  const long long units = n_pad / *unit_out;
  if (nsplit > units) nsplit = (int)(units > 0 ? units : 1);
  *T_out = T;
  *nsplit_out = nsplit;
}

struct FlatPlan {
  int T = 4, nsplit = 1;
  long long chunk = 0;
};

FlatPlan plan_flat(const vlc_ctx* c, long long m, long long n_pad) {
  FlatPlan p;
  int unit = kTile;
  choose_shape(c, m, n_pad, &p.T, &p.nsplit, &unit);
  const vlc::plan::Cut ct = vlc::plan::cut(n_pad, unit, p.nsplit);
  p.nsplit = ct.nsplit;
  p.chunk = ct.chunk;
  return p;
}

// out: [plan.nsplit][3 m] partial slots.  flag/mask/want: optional device-side dispatch (see sweep_shared).
int launch_flat(vlc_ctx* c, const double* src, long long n_pad, const FlatPlan& p, long long m, const double* dP,
                double* out, const int* flag, int mask, int want) {
  int rc;
#define VLC_SWEEP(TT, MB)                                                                                         \
  rc = c->fast ? launch_sweep_T<TT, MB, true>(c, src, n_pad, p.chunk, p.nsplit, m, dP, out, flag, mask, want)      \
               : launch_sweep_T<TT, MB, false>(c, src, n_pad, p.chunk, p.nsplit, m, dP, out, flag, mask, want)
  switch (p.T) {
    case 1: VLC_SWEEP(1, 5); break;
    case 2: VLC_SWEEP(2, 4); break;
    case 3: VLC_SWEEP(3, 3); break;
    default: VLC_SWEEP(4, 3); break;
  }
#undef VLC_SWEEP
  return rc;
}

int sweep(vlc_ctx* c, const double* src, long long n_pad, long long m, const double* dP, double* dV) {
  if (m <= 0) return VLC_OK;
  if (n_pad <= 0) {
    CUDA_OK(c, cudaMemsetAsync(dV, 0, sizeof(double) * 3 * (size_t)m, c->stream));
    return VLC_OK;
  }
  const FlatPlan p = plan_flat(c, m, n_pad);
  double* out = dV;
  if (p.nsplit > 1) {
    int rc = reserve(c, c->part, (size_t)p.nsplit * 3 * (size_t)m);
    if (rc) return rc;
    out = c->part.p;
  }
  cudaEventRecord(c->ev[0], c->stream);
  SweepStat* st = stat_begin(c);
  int rc = launch_flat(c, src, n_pad, p, m, dP, out, nullptr, 0, 0);
  if (rc) return rc;
  stat_end(c, st, 1, (double)m * (double)n_pad, (double)m * (double)n_pad * (c->fast ? 41.0 : 43.0));
  cudaEventRecord(c->ev[1], c->stream);
  if (p.nsplit > 1) {
    const long long len = 3 * m;
    vlc::bs_reduce_kernel<<<blocks_for(len, 256), 256, 0, c->stream>>>(c->part.p, p.nsplit, len, dV);
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
  }
  cudaEventRecord(c->ev[2], c->stream);
  stat_close(c, st, 0.0, 0.0);
  c->ev_valid = true;
  return VLC_OK;
}

inline int auto_strip_width(const vlc_ctx* c, int ns) {
  if (c->lat_W >= 1 && c->lat_W <= 4) return c->lat_W;
  int best = 1;
  double bc = 1e300;
  for (int W = 1; W <= 4; ++W) {
    const double cost = kLatCost[W] * (double)((ns + W - 1) / W * W) / (double)ns;
    if (cost < bc - 1e-12) {
      bc = cost;
      best = W;
    }
  }
  return best;
}
// Strip cover of a lattice with ns ring columns: one width for all strips (the last one padded), or -- when nothing was
// forced by vlc_set_lattice_tuning -- width-4 strips plus ONE tail strip of width ns mod 4 with no padding at all,
// whichever costs less per ring (kLatCost; the tail is a second, small launch: +0.5 % to prefer the single cover on ties).
struct StripPlan {
  int W = 1, tailW = 0, nmain = 0;  // nmain strips of width W, then (tailW > 0) one strip of width tailW
};
inline StripPlan plan_strips(const vlc_ctx* c, int ns, long long rings_max = -1) {
  StripPlan p;
  p.W = auto_strip_width(c, ns);
  p.nmain = (ns + p.W - 1) / p.W;
  if (c->lat_W >= 1 && c->lat_W <= 4) return p;
  const int t = ns % 4;
  // a tail strip is two more launches per sweep (merged + dual form): not worth it while the whole wake is small enough
  // for a sweep to be launch-bound (< ~2e4 rings: a few hundred microseconds)
  if (rings_max >= 0 && rings_max < 20000 && c->lat_W != 5) return p;  // lat_W == 5: tail strips whatever the size (tests)
  if (ns > 4 && t != 0) {
    const double single = kLatCost[p.W] * (double)(p.nmain * p.W) / (double)ns;
    const double mixed = (kLatCost[4] * (double)(ns - t) + kLatCost[t] * (double)t) / (double)ns + 0.005;
    if (mixed < single) {
      p.W = 4;
      p.nmain = ns / 4;
      p.tailW = t;
    }
  }
  return p;
}
inline int lat_tile_of(int W) { return vlc::lat_tile(W); }
inline int lat_rd_of(int W) { return vlc::lat_rec_doubles(W); }
inline size_t lat_smem_of(int W) { return (size_t)kStages * lat_tile_of(W) * lat_rd_of(W) * 8 + kStages * sizeof(uint64_t); }
inline long long pad_lat(long long n, int W) { return (n + lat_tile_of(W) - 1) / lat_tile_of(W) * lat_tile_of(W); }

// (W, T) instantiations of the lattice kernel and their __launch_bounds__ minimum resident CTAs
#ifndef VLC_LAT_MINB41
#define VLC_LAT_MINB41 2
#endif
#ifndef VLC_LAT_MINB42
#define VLC_LAT_MINB42 2
#endif
// the dual form (streamwise edges with two core radii, bs_lattice.cuh): one target per thread
#define VLC_DUAL_SHAPES(X) X(1, 5) X(2, 3) X(3, 3) X(4, 2)
#define VLC_LAT_SHAPES(X) X(1, 1, 6) X(1, 2, 4) X(1, 3, 2) X(2, 1, 4) X(2, 2, 2) X(2, 3, 2) X(3, 1, 3) X(3, 2, 2) X(4, 1, VLC_LAT_MINB41) X(4, 2, VLC_LAT_MINB42)

int query_occ_lat_all(vlc_ctx* c) {
#define X(WW, TT, MB)                                                                                           \
  {                                                                                                             \
    auto kern = vlc::bs_lattice_kernel<WW, TT, kLatThreads, kStages, MB>;                                          \
    CUDA_OK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lat_smem_of(WW)));  \
    CUDA_OK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_lat[WW][TT], kern, kLatThreads, lat_smem_of(WW))); \
  }
  VLC_LAT_SHAPES(X)
#undef X
#define X(WW, MB)                                                                                                \
  {                                                                                                              \
    auto kern = vlc::bs_lattice_kernel<WW, 1, kLatThreads, kStages, MB, true>;                                     \
    CUDA_OK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lat_smem_of(WW)));   \
    CUDA_OK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_dual[WW], kern, kLatThreads, lat_smem_of(WW))); \
  }
  VLC_DUAL_SHAPES(X)
#undef X
  return VLC_OK;
}

bool lat_shape_exists(int W, int T) {
#define X(WW, TT, MB) if (W == WW && T == TT) return true;
  VLC_LAT_SHAPES(X)
#undef X
  return false;
}

// Source splits of the lattice kernel: whole waves of (SMs x resident CTAs), chunks of >= 4 tiles when possible.
// *unit = the records a chunk is a multiple of: the granule (a quarter tile)
int plan_lattice_split(const vlc_ctx* c, int W, int T, long long m, long long n_lat_pad, bool dual, int* unit) {
  *unit = vlc::lat_granule(W);
  if (c->tune_nsplit > 0) return c->tune_nsplit;
  const long long tiles = n_lat_pad / lat_tile_of(W);
  const long long ttiles = (m + (long long)kLatThreads * T - 1) / ((long long)kLatThreads * T);
  const int occ = dual ? c->occ_dual[W] : c->occ_lat[W][T];
  const long long slots = (long long)c->sm_count * (occ > 0 ? occ : 2);
  const int per_tile = lat_tile_of(W) / vlc::lat_granule(W);
  const int small = vlc::plan::small_split(ttiles, tiles, slots, per_tile);
  if (small > 0) return small;
  const long long cap_by_mem = (long long)((size_t)1 << 31) / (3 * (m > 0 ? m : 1) * 8) + 1;  // <= 2 GiB partials
  return vlc::plan::wave_split(ttiles, tiles, slots, cap_by_mem, per_tile);
}

// Sweep over a set that also holds the shared-node form.  The launches are dispatched ON THE DEVICE by the set's flag so
// that no host synchronisation is needed: 0 = merged strips + flat remainder; 2 = DUAL strips (streamwise edges whose two
// copies carry different core radii, bs_lattice.cuh) + the same flat remainder; odd = the flat kernel over the reference
// enumeration.  bs_reduce_select_kernel sums the slots of whichever form ran, in fixed order.
//
// share_sources (collocation-point stage of a group / communicator, SURVEY 8e "RHS: shard sources and sum partials"): the
// targets are few and every member has them all, so the SOURCE SPLITS are shared out instead -- member k launches the
// k-th part of every slot range into a zeroed partial buffer, an all-reduce (sum) over NVLink completes the buffer on every
// member (each slot has exactly one non-zero contributor, so the sum is exact whatever NCCL's order), and the same
// fixed-order reduce follows: the result is bit-identical to one GPU's, for any number of members.
int sweep_shared(vlc_ctx* c, const SourceSet& s, long long m, const double* dP, double* dV, bool share_sources = false) {
  if (m <= 0) return VLC_OK;
  if (share_sources && !(c->world > 1 && c->comm)) share_sources = false;
  const int W_ = share_sources ? c->world : 1, k_ = share_sources ? c->rank : 0;
  auto my_lo = [&](int ns) { return (int)((long long)ns * k_ / W_); };
  auto my_hi = [&](int ns) { return (int)((long long)ns * (k_ + 1) / W_); };
  struct LatPlan {
    int W = 0, T = 1, ns = 0;
    long long chunk = 0;
  };
  auto plan = [&](int W, long long n_pad, bool dual) {
    LatPlan p;
    if (n_pad <= 0 || W < 1) return p;
    p.W = W;
    p.T = dual ? 1 : ((c->lat_T >= 1 && c->lat_T <= 3) ? c->lat_T : kLatBestT[W]);
    if (m <= kLatThreads) p.T = 1;
    while (p.T > 1 && !lat_shape_exists(W, p.T)) --p.T;
    const long long tiles = n_pad / lat_tile_of(W);
    if (p.T > 1 && !dual && c->lat_T < 1) {
      // a sweep too small to fill the machine at T targets per thread: one target per thread, T times the CTAs
      const long long tt = (m + (long long)kLatThreads * p.T - 1) / ((long long)kLatThreads * p.T);
      const long long slots = (long long)c->sm_count * (c->occ_lat[W][p.T] > 0 ? c->occ_lat[W][p.T] : 2);
      if (vlc::plan::small_split(tt, tiles, slots) > 0) p.T = 1;
    }
    int unit = 0;
    p.ns = plan_lattice_split(c, W, p.T, m, n_pad, dual, &unit);
    const vlc::plan::Cut ct = vlc::plan::cut(n_pad, unit, p.ns);
    p.ns = ct.nsplit;
    p.chunk = ct.chunk;
    return p;
  };
  const bool dual = s.may_dual;
  const LatPlan pm = plan(s.lat_W, s.n_lat_pad, false), pm2 = plan(s.lat2_W, s.n_lat2_pad, false);  // merged: main + tail strips
  const LatPlan pd = dual ? plan(s.lat_W, s.n_lat_pad, true) : LatPlan(), pd2 = dual ? plan(s.lat2_W, s.n_lat2_pad, true) : LatPlan();
  const FlatPlan pr = s.n_rem_pad > 0 ? plan_flat(c, m, s.n_rem_pad) : FlatPlan();
  const FlatPlan pf = plan_flat(c, m, s.n_pad);
  const int ns_r = s.n_rem_pad > 0 ? pr.nsplit : 0;
  const int na = pm.ns + pm2.ns, nd = pd.ns + pd2.ns;
  const size_t len = 3 * (size_t)m;
  int rc = reserve(c, c->part, (size_t)(na + ns_r + nd + pf.nsplit) * len);
  if (rc) return rc;
  double* part = c->part.p;
  double* part_rem = part + (size_t)na * len;
  double* part_dual = part_rem + (size_t)ns_r * len;
  double* part_flat = part_dual + (size_t)nd * len;
  if (share_sources) CUDA_OK(c, cudaMemsetAsync(part, 0, (size_t)(na + ns_r + nd + pf.nsplit) * len * sizeof(double), c->stream));
  auto launch_lat = [&](const LatPlan& p, bool dual_form, const double* rec, long long n_pad, double* out) -> int {
    if (p.ns <= 0) return VLC_OK;
    const int y0 = my_lo(p.ns), yn = my_hi(p.ns) - y0;  // this member's source splits (all of them without sharing)
    if (yn <= 0) return VLC_OK;
    dim3 grid(blocks_for(m, kLatThreads * p.T), (unsigned)yn, 1);
    if (!dual_form) {
#define X(WW, TT, MB)                                                                                            \
  if (p.W == WW && p.T == TT)                                                                                    \
    vlc::bs_lattice_kernel<WW, TT, kLatThreads, kStages, MB><<<grid, kLatThreads, lat_smem_of(WW), c->stream>>>( \
        rec, p.chunk, n_pad, dP, m, out, s.d_unmergeable, 3, 0, y0);
      VLC_LAT_SHAPES(X)
#undef X
    } else {
#define X(WW, MB)                                                                                                     \
  if (p.W == WW)                                                                                                      \
    vlc::bs_lattice_kernel<WW, 1, kLatThreads, kStages, MB, true><<<grid, kLatThreads, lat_smem_of(WW), c->stream>>>( \
        rec, p.chunk, n_pad, dP, m, out, s.d_unmergeable, 3, 2, y0);
      VLC_DUAL_SHAPES(X)
#undef X
    }
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
    return VLC_OK;
  };
  const bool side = (c->aux != nullptr);
  if (side) CUDA_OK(c, cudaEventRecord(c->ev_fork, c->stream));  // inputs (records, targets, partial buffer) are ready here
  cudaEventRecord(c->ev[0], c->stream);
  SweepStat* st = stat_begin(c);
  if ((rc = launch_lat(pm, false, s.lat.p, s.n_lat_pad, part))) return rc;
  if ((rc = launch_lat(pd, true, s.lat.p, s.n_lat_pad, part_dual))) return rc;  // exits at once unless the flag is 2
  // reference pairs = 4 filaments per ring; issued FP64 instructions = (11 (W+1) + 50 W) per (target, strip record) of
  // the merged form (the host does not know which form ran: a dual set issues 58 W)
  stat_end(c, st, 0, (double)m * 4.0 * (double)s.n_rings_main,
           (double)m * (double)s.n_lat_pad * (11.0 * (pm.W + 1) + 50.0 * pm.W));
  cudaEventRecord(c->ev[1], c->stream);
  // Tail strips, the flat remainder and the fallback (which exits at once unless the flag is odd) go to a low-priority
  // side stream launched AFTER the lattice kernel: their CTAs fill the SMs that the lattice kernel's last wave leaves idle.
  cudaStream_t main_stream = c->stream;
  if (side) {
    CUDA_OK(c, cudaStreamWaitEvent(c->aux, c->ev_fork, 0));  // ev_fork was recorded BEFORE the lattice kernel
    c->stream = c->aux;
  }
  rc = launch_lat(pm2, false, s.lat2.p, s.n_lat2_pad, part + (size_t)pm.ns * len);
  if (!rc) rc = launch_lat(pd2, true, s.lat2.p, s.n_lat2_pad, part_dual + (size_t)pd.ns * len);
  // the flat kernel needs no record before its chunk: a member's splits are a launch on the tail of the set
  auto launch_flat_part = [&](const double* src, long long n_pad, const FlatPlan& p, double* out, int want) -> int {
    const int y0 = my_lo(p.nsplit), yn = my_hi(p.nsplit) - y0;
    if (yn <= 0) return VLC_OK;
    FlatPlan q = p;
    q.nsplit = yn;
    return launch_flat(c, src + (size_t)y0 * p.chunk * vlc::kSrcDoubles, n_pad - (long long)y0 * p.chunk, q, m, dP,
                       out + (size_t)y0 * len, s.d_unmergeable, 1, want);
  };
  if (!rc && ns_r > 0) rc = launch_flat_part(s.rem.p, s.n_rem_pad, pr, part_rem, 0);
  if (!rc) rc = launch_flat_part(s.rec.p, s.n_pad, pf, part_flat, 1);
  c->stream = main_stream;  // restored before any early return below
  if (rc) return rc;
  if (side) {
    CUDA_OK(c, cudaEventRecord(c->ev_join, c->aux));
    CUDA_OK(c, cudaStreamWaitEvent(main_stream, c->ev_join, 0));
  }
  if (share_sources) {
    vlc::grp::Nccl& n = vlc::grp::Nccl::get();
    const int r = n.AllReduce(part, part, (size_t)(na + ns_r + nd + pf.nsplit) * len, vlc::grp::kNcclFloat64, vlc::grp::kNcclSum, c->comm,
                              (void*)c->stream);
    if (r != 0) return fail(c, VLC_ERR_CUDA, std::string("ncclAllReduce: ") + (n.GetErrorString ? n.GetErrorString(r) : "?"));
    c->launches++;
  }
  vlc::bs_reduce_select_kernel<<<blocks_for((long long)len, 256), 256, 0, c->stream>>>(part, s.d_unmergeable, na, ns_r, nd,
                                                                                      pf.nsplit, (long long)len, dV);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  cudaEventRecord(c->ev[2], c->stream);
  {  // the sweep's other kernels: tail strips (lattice form) and the flat remainder
    const double in2 = pm2.ns ? (double)m * (double)s.n_lat2_pad * (11.0 * (pm2.W + 1) + 50.0 * pm2.W) : 0.0;
    const double inr = (double)m * (double)s.n_rem_pad * (c->fast ? 41.0 : 43.0);
    stat_close(c, st, (double)m * (double)s.n_pad - (double)m * 4.0 * (double)s.n_rings_main, in2 + inr);
  }
  c->ev_valid = true;
  return VLC_OK;
}

// In-place all-gather of equal slots on the context's stream: slot k = doubles [k*cnt, (k+1)*cnt) of `buf`, this rank's
// slot is filled.  NCCL when the context has a communicator (vlc_comm_init_rank; vlc_create_multi on distinct devices).
// `peer_buf(member)` names the same buffer in another member of an in-process group for the fallback without NCCL: every
// member waits until all slots are complete, pulls the other slots with peer copies, and waits again before anybody may
// overwrite its slot.
int allgather_slots(vlc_ctx* c, double* buf, size_t cnt, const std::function<double*(vlc_ctx*)>& peer_buf) {
  if (c->world <= 1 || cnt == 0) return VLC_OK;
  if (c->comm) {
    vlc::grp::Nccl& n = vlc::grp::Nccl::get();
    const int r = n.AllGather(buf + (size_t)c->rank * cnt, buf, cnt, vlc::grp::kNcclFloat64, c->comm, (void*)c->stream);
    if (r != 0) return fail(c, VLC_ERR_CUDA, std::string("ncclAllGather: ") + (n.GetErrorString ? n.GetErrorString(r) : "?"));
    c->launches++;
    return VLC_OK;
  }
  if (c->group && t_in_member) {
    vlc_group* g = c->group;
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    g->barrier->wait();
    for (vlc_ctx* o : g->members) {
      if (o == c) continue;
      CUDA_OK(c, cudaMemcpyPeerAsync(buf + (size_t)o->rank * cnt, c->device, peer_buf(o) + (size_t)o->rank * cnt, o->device,
                                     cnt * sizeof(double), c->stream));
    }
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    g->barrier->wait();
    return VLC_OK;
  }
  return fail(c, VLC_ERR_STATE, "world > 1 without a communicator (vlc_comm_init_rank) or a group (vlc_create_multi)");
}

void host_parallel(vlc_ctx* c, long long n, const std::function<void(long long, long long)>& fn);

// VLC_TIMERS=1: host wall time of the phases of the host-buffer calls, printed at vlc_destroy (a debugging aid: where the
// time of a synchronous per-sweep hand-over goes besides the sweep itself)
struct HostTimers {
  bool on = std::getenv("VLC_TIMERS") != nullptr;
  double s[5] = {0, 0, 0, 0, 0};
  long long n[5] = {0, 0, 0, 0, 0};
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  ~HostTimers() {
    if (!on) return;
    static const char* name[5] = {"gather targets", "pack (enqueue)", "h2d + sweep (enqueue)", "d2h + wait", "put (upload)"};
    for (int k = 0; k < 5; ++k)
      if (n[k]) std::fprintf(stderr, "[vlc timers] %-22s %8lld calls %10.3f ms total %8.1f us each\n", name[k], n[k], 1e3 * s[k], 1e6 * s[k] / n[k]);
  }
};
HostTimers g_timers;
struct TimerScope {
  int k;
  double t0;
  explicit TimerScope(int k_) : k(k_), t0(g_timers.on ? HostTimers::now() : 0.0) {}
  ~TimerScope() {
    if (g_timers.on) {
      g_timers.s[k] += HostTimers::now() - t0;
      g_timers.n[k]++;
    }
  }
};

// host-buffer sweep: H2D targets, sweep, D2H velocities (synchronous).  With world > 1 the call is COLLECTIVE (every
// member / rank makes it with the same arguments) and each takes its contiguous slice of the targets against all
// sources (libCommon.f90:132-139 is a parallel loop over targets): the members of an in-process group write their
// slices of V straight into the caller's array; one process per GPU all-gathers the slices so that every rank returns
// the whole V.
int sweep_host(vlc_ctx* c, const double* src, long long n_pad, long long m, const double* P, double* V,
               const SourceSet* shared = nullptr) {
  if (m <= 0) return VLC_OK;
  if (!P || !V) return fail(c, VLC_ERR_ARG, "null target / output pointer");
  vlc::grp::Shard sh;
  sh.per = m;
  sh.hi = m;
  if (c->world > 1) sh = vlc::grp::shard_range(m, c->world, c->rank);
  const long long ml = sh.count();
  const bool gather = c->world > 1 && !c->group;
  int rc = reserve(c, c->stage_P, 3 * (size_t)std::max(ml, 1LL));
  if (rc) return rc;
  rc = reserve(c, c->stage_V, gather ? 3 * (size_t)sh.per * c->world : 3 * (size_t)std::max(ml, 1LL));
  if (rc) return rc;
  double* dV = c->stage_V.p + (gather ? 3 * (size_t)sh.per * c->rank : 0);
  if (ml > 0) {
    TimerScope ts(2);
    CUDA_OK(c, cudaMemcpyAsync(c->stage_P.p, P + 3 * sh.lo, sizeof(double) * 3 * (size_t)ml, cudaMemcpyHostToDevice, c->stream));
    rc = shared ? sweep_shared(c, *shared, ml, c->stage_P.p, dV) : sweep(c, src, n_pad, ml, c->stage_P.p, dV);
    if (rc) return rc;
  }
  TimerScope ts(3);
  if (gather) {
    if ((rc = allgather_slots(c, c->stage_V.p, 3 * (size_t)sh.per, nullptr))) return rc;
    CUDA_OK(c, cudaMemcpyAsync(V, c->stage_V.p, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  } else if (ml > 0) {
    // straight into the caller's (pageable) array: staging through pinned scratch + a threaded copy was measured and is
    // no faster at 0.7 MB per call (r02z)
    CUDA_OK(c, cudaMemcpyAsync(V + 3 * sh.lo, dV, sizeof(double) * 3 * (size_t)ml, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

// pinned host scratch of at least n doubles (the previous sweep_host has synchronised, so it is free to overwrite)
int host_targets(vlc_ctx* c, size_t n, double** out) {
  if (n > c->host_P_cap) {
    if (c->host_P) cudaFreeHost(c->host_P);
    c->host_P = nullptr;
    c->host_P_cap = 0;
    const size_t want = n + n / 4 + 1024;
    CUDA_OK(c, cudaMallocHost(&c->host_P, want * sizeof(double)));
    c->host_P_cap = want;
  }
  *out = c->host_P;
  return VLC_OK;
}

// Host loops over the caller's records (one 24-byte read per 400-byte record: a cache line per target) on a few
// persistent threads: at 3e4 targets per call the single-threaded gather was 0.27 ms of GPU idle time per call
// (bench.py e2e breakdown, r02w).  fn(lo, hi) over [0, n).
void host_parallel(vlc_ctx* c, long long n, const std::function<void(long long, long long)>& fn) {
  if (n < 8192) {
    fn(0, n);
    return;
  }
  if (!c->host_pool) {
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    const unsigned members = c->world > 1 && c->group ? (unsigned)c->world : 1u;  // members of a group gather concurrently
    int nt = (int)(hc / (2 * members));
    nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
    c->host_pool = new vlc::grp::Workers(nt);
  }
  const int nt = c->host_pool->size();
  if (nt <= 1) {
    fn(0, n);
    return;
  }
  c->host_pool->run([&](int k) {
    fn(n * k / nt, n * (k + 1) / nt);
    return 0;
  });
}

int check_set(vlc_ctx* c, int set) {
  if (set < 0 || set >= VLC_MAX_SETS) return fail(c, VLC_ERR_ARG, "source set index out of range");
  return VLC_OK;
}

Rotor* get_rotor(vlc_ctx* c, int ir) {
  if (ir < 0 || ir >= (int)c->rotors.size() || !c->rotors[ir].defined) {
    c->err = "rotor not defined";
    return nullptr;
  }
  return &c->rotors[ir];
}

// Builder of the segment table of pack_table_kernel (pack.cuh): every segment of a packed set in ONE launch.
struct PackBuilder {
  vlc::PackTable t;
  long long total = 0;
  PackBuilder() {
    t.n = 0;
    t.start[0] = 0;
  }
  vlc::PackSeg& add(int kind, long long count, int nb) {
    vlc::PackSeg& g = t.seg[t.n];
    std::memset(&g, 0, sizeof g);
    g.kind = kind;
    g.count = count;
    g.nb = nb;
    g.W = 4;
    total += count * nb;
    t.start[++t.n] = total;
    return g;
  }
  bool full() const { return t.n >= vlc::kPackSegs - 1; }
  // ring filaments: rings (i0 .. i0+ni-1, 0 .. nj-1) of `base`, the filaments of `mask`, per blade
  void rings(const double* base, int stride, int ld, int i0, int ni, int nj, int mask, int nfil, double sign, int wake, double* dst,
             int nb, long long src_blade, long long dst_blade) {
    if ((long long)ni * nj * nfil <= 0) return;
    vlc::PackSeg& g = add(0, (long long)ni * nj * nfil, nb);
    g.src = base;
    g.dst = dst;
    g.src_blade = src_blade;
    g.dst_blade = dst_blade;
    g.stride = stride;
    g.ld = ld;
    g.i0 = i0;
    g.ni = ni;
    g.mask = mask;
    g.nfil = nfil;
    g.sign = sign;
    g.wake = wake;
  }
  void fwake(const double* base, int i0, int ni, double* dst, int nb, long long src_blade, long long dst_blade) {
    if (ni <= 0) return;
    vlc::PackSeg& g = add(1, ni, nb);
    g.src = base;
    g.dst = dst;
    g.src_blade = src_blade;
    g.dst_blade = dst_blade;
    g.i0 = i0;
  }
  void nulls(double* dst, long long count) {
    if (count <= 0) return;
    vlc::PackSeg& g = add(2, count, 1);
    g.dst = dst;
  }
  void strips(int W, const double* base, int stride, int ld, int i0, int nrows, int ns, int col_base, int nstrips, double* dst, int* flag,
              int nb, long long src_blade) {
    vlc::PackSeg& g = add(3, (long long)nstrips * (nrows + 1), nb);
    g.W = W;
    g.src = base;
    g.dst = dst;
    g.src_blade = src_blade;
    g.stride = stride;
    g.ld = ld;
    g.i0 = i0;
    g.nrows = nrows;
    g.ns = ns;
    g.col_base = col_base;
    g.nstrips = nstrips;
    g.flag = flag;
  }
  void null_strips(int W, double* dst, long long count) {
    if (count <= 0) return;
    vlc::PackSeg& g = add(4, count, 1);
    g.W = W;
    g.dst = dst;
  }
  int launch(vlc_ctx* c) {
    if (total <= 0) return VLC_OK;
    vlc::pack_table_kernel<<<blocks_for(total, 128), 128, 0, c->stream>>>(t);
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
    return VLC_OK;
  }
};

// (Re)build the packed [wing | wake] set of a rotor in the reference's enumeration order -- and, for a near wake, its
// shared-node form: strip records per blade + the flat remainder [wing | last column, horseshoe corrections, far wakes].
// Three launches whatever the rotor looks like: clear the flag, classify the records (check_rings_kernel), pack every
// segment (pack_table_kernel).
void mark_wing_dirty(Rotor& r) {
  for (int s = 0; s < 2; ++s)
    if (r.dirty[s] == 0) r.dirty[s] = 2;
  r.bound_dirty = r.chord_dirty = true;
}

int pack_rotor(vlc_ctx* c, Rotor& r, int s) {
  if (!r.dirty[s]) return VLC_OK;
  for (int ib = 0; ib < r.nb; ++ib)
    if ((r.nNwake > 0 && r.rowNear < r.stale_near[s][ib]) || (r.nFwake > 0 && r.rowFar < r.stale_far[s][ib]))
      return fail(c, VLC_ERR_STATE, "wake rows between rowNear/rowFar and the last upload were never transferred: "
                                    "call vlc_rotor_set_rows before vlc_rotor_put_nwake / _put_fwake");
  const int nrows = r.nNwake > 0 ? (r.nNwake - r.rowNear + 1) : 0;
  const bool has_far = (r.nNwake > 0) && (r.rowFar <= r.nFwake);  // classdef.f90:1458
  const int nfar = has_far ? (r.nFwake - r.rowFar + 1) : 0;
  const bool lifting = (abs(r.surfaceType) == 1);  // classdef.f90:4432
  const long long wing_n = lifting ? 4LL * r.nc * r.ns * r.nb : 0;
  const long long wing_pad = pad_tile(wing_n);
  long long wake_per_blade = 4LL * nrows * r.ns;
  if (has_far) wake_per_blade += r.ns + nfar + (r.have_pf[s] ? VLC_NPFWAKE : 0);
  const long long wake_n = wake_per_blade * r.nb;
  const long long wake_pad = pad_tile(wake_n);
  const long long total_pad = wing_pad + wake_pad;
  int rc = reserve(c, r.comb[s].rec, (size_t)total_pad * vlc::kSrcDoubles);
  if (rc) return rc;
  double* rec = r.comb[s].rec.p;
  cudaStream_t st = c->stream;
  const long long wiP_blade = (long long)r.nc * r.ns * vlc::kWp, waN_blade = (long long)r.nNwake * r.ns * vlc::kVr;
  const long long waF_blade = (long long)r.nFwake * vlc::kFw, wapF_blade = (long long)VLC_NPFWAKE * vlc::kFw;
  if (r.dirty[s] == 2 && r.comb[s].n_pad == total_pad && r.wing_pad[s] == wing_pad) {
    // only the wing changed since the last pack (a solve, a moved wing): its rings again, into the flat enumeration and
    // into the flat remainder of the shared-node form; the wake's records, strips, flag and padding stand
    PackBuilder pw;
    if (wing_n > 0) {
      pw.rings(r.wiP.p, vlc::kWp, r.nc, 0, r.nc, r.ns, 0xF, 4, 1.0, 0, rec, r.nb, wiP_blade, 4LL * r.nc * r.ns);
      if (r.comb[s].has_shared) pw.rings(r.wiP.p, vlc::kWp, r.nc, 0, r.nc, r.ns, 0xF, 4, 1.0, 0, r.comb[s].rem.p, r.nb, wiP_blade, 4LL * r.nc * r.ns);
    }
    r.dirty[s] = 0;
    return pw.launch(c);
  }
  PackBuilder pb;
  // ---- the flat enumeration: [wing | padding | per blade: rings, horseshoe correction, far wake, prescribed wake | padding]
  auto flat_segments = [&](double* dst, bool rings_all) {
    if (wing_n > 0) pb.rings(r.wiP.p, vlc::kWp, r.nc, 0, r.nc, r.ns, 0xF, 4, 1.0, 0, dst, r.nb, wiP_blade, 4LL * r.nc * r.ns);
    pb.nulls(dst + (size_t)wing_n * vlc::kSrcDoubles, wing_pad - wing_n);
    double* w = dst + (size_t)wing_pad * vlc::kSrcDoubles;
    // rings_all: 4 filaments of every ring (reference enumeration); otherwise only what no strip covers: f3 of the last column
    const long long per_blade = rings_all ? wake_per_blade : (long long)nrows + (has_far ? r.ns + nfar + (r.have_pf[s] ? VLC_NPFWAKE : 0) : 0);
    long long off = 0;
    if (r.nNwake > 0 && nrows > 0) {
      if (rings_all) {
        pb.rings(r.waN[s].p, vlc::kVr, r.nNwake, r.rowNear - 1, nrows, r.ns, 0xF, 4, 1.0, 1, w, r.nb, waN_blade, per_blade);
        off += 4LL * nrows * r.ns;
      } else {  // last column: f3 of ring (i, ns-1), wake rule applies (classdef.f90:1452)
        pb.rings(r.waN[s].p + (size_t)vlc::kVr * r.nNwake * (r.ns - 1), vlc::kVr, r.nNwake, r.rowNear - 1, nrows, 1, 0x4, 1, 1.0, 1, w,
                 r.nb, waN_blade, per_blade);
        off += nrows;
      }
    }
    if (has_far) {
      // horseshoe correction: -vf(2) of the last near row, no gam rule (classdef.f90:1460-1463)
      pb.rings(r.waN[s].p, vlc::kVr, r.nNwake, r.nNwake - 1, 1, r.ns, 0x2, 1, -1.0, 0, w + (size_t)off * vlc::kSrcDoubles, r.nb, waN_blade,
               per_blade);
      off += r.ns;
      pb.fwake(r.waF[s].p, r.rowFar - 1, nfar, w + (size_t)off * vlc::kSrcDoubles, r.nb, waF_blade, per_blade);
      off += nfar;
      if (r.have_pf[s]) {
        pb.fwake(r.wapF[s].p, 0, VLC_NPFWAKE, w + (size_t)off * vlc::kSrcDoubles, r.nb, wapF_blade, per_blade);
        off += VLC_NPFWAKE;
      }
    }
    const long long n = per_blade * r.nb;
    pb.nulls(w + (size_t)n * vlc::kSrcDoubles, pad_tile(n) - n);
    return n;
  };
  flat_segments(rec, true);
  r.comb[s].n = wing_n + wake_n;
  r.comb[s].n_pad = total_pad;
  r.wing_pad[s] = wing_pad;
  r.wing_n = wing_n;
  r.dirty[s] = 0;

  // ---- shared-node form of the near wake (bs_lattice.cuh): strips per blade + [wing | remainder] flat ----
  SourceSet& cs = r.comb[s];
  cs.has_shared = false;
  cs.n_lat = cs.n_lat_pad = cs.n_rem = cs.n_rem_pad = cs.n_lat2 = cs.n_lat2_pad = 0;
  cs.lat2_W = 0;
  if (nrows > 0 && c->shared_nodes) {
    const StripPlan sp = plan_strips(c, r.ns, (long long)r.nb * r.nNwake * r.ns);
    const int LW = sp.W, RD = lat_rd_of(LW), nstrips = sp.nmain;
    const long long lat_n = (long long)r.nb * nstrips * (nrows + 1);
    const long long lat_pad = pad_lat(lat_n, LW);
    cs.lat_W = LW;
    const int TW = sp.tailW, RD2 = TW ? lat_rd_of(TW) : 0;  // one tail strip per blade
    const long long lat2_n = TW ? (long long)r.nb * (nrows + 1) : 0, lat2_pad = TW ? pad_lat(lat2_n, TW) : 0;
    if (TW && (rc = reserve(c, cs.lat2, (size_t)lat2_pad * RD2))) return rc;
    long long rem_per_blade = nrows;  // streamwise edges of the last column
    if (has_far) rem_per_blade += r.ns + nfar + (r.have_pf[s] ? VLC_NPFWAKE : 0);
    const long long rem_wake = rem_per_blade * r.nb;
    const long long rem_pad = wing_pad + pad_tile(rem_wake);
    if ((rc = reserve(c, cs.lat, (size_t)lat_pad * RD))) return rc;
    if ((rc = reserve(c, cs.rem, (size_t)rem_pad * vlc::kSrcDoubles))) return rc;
    if (!cs.d_unmergeable) CUDA_OK(c, cudaMalloc(&cs.d_unmergeable, sizeof(int)));
    CUDA_OK(c, cudaMemsetAsync(cs.d_unmergeable, 0, sizeof(int), st));
    {
      const long long nring = (long long)nrows * r.ns;
      vlc::check_rings_kernel<<<dim3(blocks_for(nring, 256), (unsigned)r.nb, 1), 256, 0, st>>>(r.waN[s].p, vlc::kVr, r.nNwake, r.rowNear - 1,
                                                                                           nrows, r.ns, cs.d_unmergeable, waN_blade);
      c->launches++;
    }
    flat_segments(cs.rem.p, false);
    pb.strips(LW, r.waN[s].p, vlc::kVr, r.nNwake, r.rowNear - 1, nrows, r.ns, 0, nstrips, cs.lat.p, cs.d_unmergeable, r.nb, waN_blade);
    pb.null_strips(LW, cs.lat.p + (size_t)lat_n * RD, lat_pad - lat_n);
    if (TW) {  // the tail strip: columns nstrips*LW .. ns-1
      pb.strips(TW, r.waN[s].p, vlc::kVr, r.nNwake, r.rowNear - 1, nrows, r.ns, nstrips * LW, 1, cs.lat2.p, cs.d_unmergeable, r.nb, waN_blade);
      pb.null_strips(TW, cs.lat2.p + (size_t)lat2_n * RD2, lat2_pad - lat2_n);
    }
    cs.n_lat = lat_n;
    cs.n_lat_pad = lat_pad;
    cs.n_rings_main = (long long)r.nb * nrows * std::min(r.ns, nstrips * LW);
    cs.may_dual = true;
    cs.n_lat2 = lat2_n;
    cs.n_lat2_pad = lat2_pad;
    cs.lat2_W = TW;
    cs.n_rem = wing_n + rem_wake;
    cs.n_rem_pad = rem_pad;
    cs.has_shared = true;
  }
  return pb.launch(c);
}

// View of a rotor's packed set without its wing segment (vind_bywake).
SourceSet wake_view(const Rotor& r, int s) {
  SourceSet v = r.comb[s];  // shallow: buffers stay owned by the rotor
  const size_t off = (size_t)r.wing_pad[s] * vlc::kSrcDoubles;
  v.rec.p = r.comb[s].rec.p + off;
  v.n_pad = r.comb[s].n_pad - r.wing_pad[s];
  if (v.has_shared) {
    v.rem.p = r.comb[s].rem.p + off;
    v.n_rem_pad = r.comb[s].n_rem_pad - r.wing_pad[s];
  }
  return v;
}

// bound-vortex set (classdef.f90:1376-1396): (vf2 + vf4)*gam of every ring, minus vf2*gam of row nc; chordwise-vortex set
// (classdef.f90:1398-1418): (vf1 + vf3)*gam of every ring, plus vf2*gam of row nc -- the same layout with the other two
// filaments and the opposite sign on the trailing-edge row.  One launch each (pack_table_kernel).
int pack_wing_subset(vlc_ctx* c, Rotor& r, SourceSet& set, bool& dirty, int mask, double te_sign) {
  if (!dirty) return VLC_OK;
  const long long per_blade = 2LL * r.nc * r.ns + r.ns;
  const long long n = per_blade * r.nb;
  const long long n_pad = pad_tile(n);
  int rc = reserve(c, set.rec, (size_t)n_pad * vlc::kSrcDoubles);
  if (rc) return rc;
  double* rec = set.rec.p;
  const long long cnt = 2LL * r.nc * r.ns, wiP_blade = (long long)r.nc * r.ns * vlc::kWp;
  PackBuilder pb;
  pb.rings(r.wiP.p, vlc::kWp, r.nc, 0, r.nc, r.ns, mask, 2, 1.0, 0, rec, r.nb, wiP_blade, per_blade);
  pb.rings(r.wiP.p, vlc::kWp, r.nc, r.nc - 1, 1, r.ns, 0x2, 1, te_sign, 0, rec + (size_t)cnt * vlc::kSrcDoubles, r.nb, wiP_blade, per_blade);
  pb.nulls(rec + (size_t)n * vlc::kSrcDoubles, n_pad - n);
  if ((rc = pb.launch(c))) return rc;
  set.n = n;
  set.n_pad = n_pad;
  dirty = false;
  return VLC_OK;
}
int pack_bound(vlc_ctx* c, Rotor& r) { return pack_wing_subset(c, r, r.bound, r.bound_dirty, 0xA, -1.0); }
int pack_chord(vlc_ctx* c, Rotor& r) { return pack_wing_subset(c, r, r.chord, r.chord_dirty, 0x5, 1.0); }

// Form of a set made of several rotors' records laid side by side: the rotors' common form when they agree, the flat
// enumeration (1) when one of them needs it or when they disagree (merged and dual records cannot share a launch).
__global__ void or_flags_kernel(int n, const int* const* __restrict__ flags, int* __restrict__ out) {
  int f = *flags[0];
  for (int k = 1; k < n; ++k)
    if (*flags[k] != f) f = 1;
  *out = (f & 1) ? 1 : f;
}

// The wake sweep sums over ALL source rotors (main.f90:817-826: vel = vel + vind_on?wake_byRotor(rotor(jr), ...)).  One
// sweep per source rotor costs a wave tail, a pipeline fill per CTA and four launches each (measured r02e: 3.3 % of the
// sweep at 1e6 filaments with 5 source rotors); instead the rotors' packed sets -- every one already padded with null
// records to whole tiles -- are laid side by side with device-to-device copies (O(N), < 0.1 % of a sweep) and swept as
// ONE set: strip records of all near wakes | flat remainders (wings, last columns, far wakes) | the flat enumeration
// for the fallback, with the OR of the rotors' mergeability flags.  Possible when every lattice uses the same strip
// width and none needs tail strips; *ok = false otherwise (the caller then sweeps rotor by rotor).  A strip record that
// starts a rotor's block starts a strip (streamwise strengths 0), so the nodes of the record before it are never used.
int build_ws_combined(vlc_ctx* c, int s, bool* ok) {
  *ok = false;
  std::vector<Rotor*> src;
  for (auto& r : c->rotors)
    if (r.defined) {
      int rc = pack_rotor(c, r, s);
      if (rc) return rc;
      if (r.comb[s].n_pad > 0) src.push_back(&r);
    }
  if (src.size() < 2 || src.size() > 64) return VLC_OK;  // one source rotor: its own set is the combined set
  int W = 0;
  bool any_shared = false;
  for (Rotor* r : src) {
    const SourceSet& v = r->comb[s];
    if (!(v.has_shared && c->shared_nodes && v.n_lat_pad > 0)) continue;
    if (v.n_lat2_pad > 0) return VLC_OK;
    if (W && v.lat_W != W) return VLC_OK;
    W = v.lat_W;
    any_shared = true;
  }
  SourceSet& w = c->ws_comb[s];
  long long n = 0, n_pad = 0, lat_n = 0, lat_pad = 0, rem_n = 0, rem_pad = 0, rings = 0;
  for (Rotor* r : src) {
    const SourceSet& v = r->comb[s];
    const bool sh = any_shared && v.has_shared && v.n_lat_pad > 0;
    n += v.n;
    n_pad += v.n_pad;
    if (sh) {
      lat_n += v.n_lat;
      lat_pad += v.n_lat_pad;
      rings += v.n_rings_main;
    }
    rem_n += sh ? v.n_rem : v.n;  // a rotor without a lattice (no wake row yet) goes to the remainder whole
    rem_pad += sh ? v.n_rem_pad : v.n_pad;
  }
  int rc;
  if ((rc = reserve(c, w.rec, (size_t)n_pad * vlc::kSrcDoubles))) return rc;
  const int RD = W ? lat_rd_of(W) : 0;
  if (any_shared) {
    if ((rc = reserve(c, w.lat, (size_t)lat_pad * RD)) || (rc = reserve(c, w.rem, (size_t)rem_pad * vlc::kSrcDoubles))) return rc;
    if (!w.d_unmergeable) CUDA_OK(c, cudaMalloc(&w.d_unmergeable, sizeof(int)));
    if (!c->ws_flags) CUDA_OK(c, cudaMalloc(&c->ws_flags, sizeof(int*) * 128));  // 64 per record set
  }
  cudaStream_t st = c->stream;
  auto d2d = [&](double* dst, const double* from, size_t doubles) -> int {
    if (doubles) CUDA_OK(c, cudaMemcpyAsync(dst, from, doubles * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return VLC_OK;
  };
  size_t o_rec = 0, o_lat = 0, o_rem = 0;
  std::vector<const int*> flags;
  for (Rotor* r : src) {
    const SourceSet& v = r->comb[s];
    const bool sh = any_shared && v.has_shared && v.n_lat_pad > 0;
    if ((rc = d2d(w.rec.p + o_rec, v.rec.p, (size_t)v.n_pad * vlc::kSrcDoubles))) return rc;
    o_rec += (size_t)v.n_pad * vlc::kSrcDoubles;
    if (!any_shared) continue;
    if (sh) {
      if ((rc = d2d(w.lat.p + o_lat, v.lat.p, (size_t)v.n_lat_pad * RD))) return rc;
      o_lat += (size_t)v.n_lat_pad * RD;
      if ((rc = d2d(w.rem.p + o_rem, v.rem.p, (size_t)v.n_rem_pad * vlc::kSrcDoubles))) return rc;
      o_rem += (size_t)v.n_rem_pad * vlc::kSrcDoubles;
      flags.push_back(v.d_unmergeable);
    } else {
      if ((rc = d2d(w.rem.p + o_rem, v.rec.p, (size_t)v.n_pad * vlc::kSrcDoubles))) return rc;
      o_rem += (size_t)v.n_pad * vlc::kSrcDoubles;
    }
  }
  if (any_shared) {
    const int* const* d_list = reinterpret_cast<const int* const*>(c->ws_flags) + 64 * s;
    if (c->ws_flags_host[s] != flags) {  // the addresses are stable: uploaded once per record set
      CUDA_OK(c, cudaStreamSynchronize(st));
      c->ws_flags_host[s] = flags;
      CUDA_OK(c, cudaMemcpy((void*)d_list, c->ws_flags_host[s].data(), sizeof(int*) * flags.size(), cudaMemcpyHostToDevice));
    }
    or_flags_kernel<<<1, 1, 0, st>>>((int)flags.size(), d_list, w.d_unmergeable);
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
  }
  w.n = n;
  w.n_pad = n_pad;
  w.has_shared = any_shared;
  w.lat_W = W ? W : 1;
  w.n_lat = lat_n;
  w.n_lat_pad = any_shared ? lat_pad : 0;
  w.n_lat2 = w.n_lat2_pad = 0;
  w.lat2_W = 0;
  w.n_rem = rem_n;
  w.n_rem_pad = any_shared ? rem_pad : 0;
  w.n_rings_main = rings;
  w.may_dual = any_shared;
  *ok = true;
  return VLC_OK;
}

// Host -> device copy of a caller's array.  The caller may reuse its buffer as soon as the call returns.  Small arrays
// (the moved wing of every time step, section frames: < 1 MiB) go through one of two pinned staging buffers and the copy
// is left in flight -- no synchronisation, the stream keeps running ahead (a synchronous hand-over costs ~20 us of idle
// device per call, several per time step of a small case); large arrays are copied from the caller's memory and waited for.
int upload(vlc_ctx* c, DevBuf& b, size_t total, size_t offset, const double* host, size_t count) {
  int rc = reserve(c, b, total);
  if (rc) return rc;
  const size_t bytes = count * sizeof(double);
  if (bytes > 0 && bytes <= vlc_ctx::kStageBytes) {
    const int k = c->stage_next;
    c->stage_next ^= 1;
    if (!c->stage_h[k]) {
      CUDA_OK(c, cudaMallocHost(&c->stage_h[k], vlc_ctx::kStageBytes));
      CUDA_OK(c, cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming));
    } else {
      CUDA_OK(c, cudaEventSynchronize(c->stage_ev[k]));  // the copy that last used this buffer (two uploads ago) is done
    }
    std::memcpy(c->stage_h[k], host, bytes);
    CUDA_OK(c, cudaMemcpyAsync(b.p + offset, c->stage_h[k], bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaEventRecord(c->stage_ev[k], c->stream));
    return VLC_OK;
  }
  CUDA_OK(c, cudaMemcpyAsync(b.p + offset, host, bytes, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

}  // namespace

// ============================================================================ context

extern "C" const char* vlc_version(void) { return "volcanor_b200 0.1 (sm_100a, FP64)"; }

extern "C" int vlc_create(int device, vlc_ctx** out) {
  if (!out) return VLC_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                     " (volcanor_b200 has no CPU fallback)";
    return VLC_ERR_NODEVICE;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return VLC_ERR_ARG;
  }
  vlc_ctx* c = new vlc_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    g_create_error = "cudaSetDevice / cudaGetDeviceProperties failed";
    delete c;
    return VLC_ERR_CUDA;
  }
  if (prop.major < 10) {
    g_create_error = std::string("device '") + prop.name + "' is not sm_100-class; this library ships sm_100a code only";
    delete c;
    return VLC_ERR_NODEVICE;
  }
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->mem_bytes = (long long)prop.totalGlobalMem;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "cudaStreamCreate failed";
    delete c;
    return VLC_ERR_CUDA;
  }
  c->stream = c->own_stream;
  for (auto& e : c->ev) cudaEventCreate(&e);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = numerically largest = least urgent
    if (cudaStreamCreateWithPriority(&c->aux, cudaStreamNonBlocking, lo) != cudaSuccess) c->aux = nullptr;
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  }
  int rc = 0;
  rc |= query_occ<1, 5>(c, &c->occ[1]);
  rc |= query_occ<2, 4>(c, &c->occ[2]);
  rc |= query_occ<3, 3>(c, &c->occ[3]);
  rc |= query_occ<4, 3>(c, &c->occ[4]);
  rc |= query_occ_lat_all(c);
  // Source-split partials: the planners aim at >= 8 waves of CTAs, which makes splits x targets ~ constant (tens of MB)
  // until the targets alone fill the machine; 128 MB up front covers every case-sized sweep without a re-allocation.
  if (!rc) rc = reserve(c, c->part, (size_t)16 << 20);
  if (rc) {
    g_create_error = "sweep kernel not loadable on this device: " + c->err;
    cudaStreamDestroy(c->own_stream);
    delete c;
    return VLC_ERR_CUDA;
  }
  *out = c;
  return VLC_OK;
}

extern "C" int vlc_destroy(vlc_ctx* c) {
  if (!c) return VLC_OK;
  if (c->group && c->is_leader) {  // a group handle: stop the workers, then destroy every member (the leader last)
    vlc_group* g = c->group;
    delete g->workers;  // joins the threads
    g->workers = nullptr;
    for (vlc_ctx* m : g->members) m->group = nullptr;
    for (size_t k = 1; k < g->members.size(); ++k) vlc_destroy(g->members[k]);
    delete g->barrier;
    delete g;
    c->is_leader = false;
  }
  cudaSetDevice(c->device);
  if (c->comm) {
    cudaStreamSynchronize(c->stream);
    vlc::grp::Nccl::get().CommDestroy(c->comm);
    c->comm = nullptr;
  }
  cudaStreamSynchronize(c->stream);
  for (auto& s : c->sets) {
    release(s.rec);
    release(s.lat);
    release(s.lat2);
    release(s.rem);
    if (s.d_unmergeable) cudaFree(s.d_unmergeable);
  }
  release(c->part);
  release(c->stage_P);
  release(c->stage_V);
  release(c->scratch);
  for (auto& w : c->ws_comb) {
    release(w.rec);
    release(w.lat);
    release(w.rem);
    if (w.d_unmergeable) cudaFree(w.d_unmergeable);
  }
  if (c->ws_flags) cudaFree(c->ws_flags);
  release(c->ws_P);
  release(c->ws_V);
  release(c->ws_acc);
  release(c->cp_P);
  release(c->cp_V);
  release(c->solver_work);
  if (c->d_flag) cudaFree(c->d_flag);
  for (auto& r : c->rotors) {
    release(r.wiP);
    for (int s = 0; s < 2; ++s) {
      release(r.waN[s]);
      release(r.waF[s]);
      release(r.wapF[s]);
      release(r.pfHelix[s]);
      release(r.comb[s].rec);
      release(r.comb[s].lat);
      release(r.comb[s].lat2);
      release(r.comb[s].rem);
      if (r.comb[s].d_unmergeable) cudaFree(r.comb[s].d_unmergeable);
    }
    for (int k = 0; k < 4; ++k) {
      release(r.velN[k]);
      release(r.velF[k]);
    }
    for (int k = 0; k < 2; ++k) {
      release(r.velNx[k]);
      release(r.velFx[k]);
    }
    release(r.waN_alt);
    release(r.order2_tmp);
    release(r.pfFits);
    if (r.d_axi) cudaFree(r.d_axi);
    release(r.bound.rec);
    release(r.chord.rec);
    release(r.rhs);
    release(r.gamvec);
    release(r.sec);
    release(r.loads);
    release(r.loads_scr);
    release(r.LU);
    release(r.Ainv);
    if (r.d_ipiv) cudaFree(r.d_ipiv);
    if (r.d_info) cudaFree(r.d_info);
  }
  if (c->solver) cusolverDnDestroy(c->solver);
  delete c->host_pool;
  c->host_pool = nullptr;
  if (c->host_P) cudaFreeHost(c->host_P);
  for (int k = 0; k < 2; ++k) {
    if (c->stage_h[k]) cudaFreeHost(c->stage_h[k]);
    if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]);
  }
  for (auto& e : c->user_ev)
    if (e) cudaEventDestroy(e);
  for (auto& st : c->stats) {
    if (st.e0) cudaEventDestroy(st.e0);
    if (st.e1) cudaEventDestroy(st.e1);
    if (st.e2) cudaEventDestroy(st.e2);
  }
  if (c->d_flush) cudaFree(c->d_flush);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return VLC_OK;
}

extern "C" int vlc_create_multi(int n_devices, const int* devices, vlc_ctx** out) {
  if (!out) return VLC_ERR_ARG;
  *out = nullptr;
  if (n_devices < 1 || n_devices > 64) {
    g_create_error = "vlc_create_multi: n_devices outside 1..64";
    return VLC_ERR_ARG;
  }
  std::vector<int> dev(n_devices);
  for (int k = 0; k < n_devices; ++k) dev[k] = devices ? devices[k] : k;
  vlc_group* g = new vlc_group();
  auto undo = [&](int rc) {
    for (vlc_ctx* m : g->members) {
      m->group = nullptr;
      m->is_leader = false;
      vlc_destroy(m);
    }
    delete g;
    return rc;
  };
  for (int k = 0; k < n_devices; ++k) {
    vlc_ctx* m = nullptr;
    const int rc = vlc_create(dev[k], &m);
    if (rc) return undo(rc);  // g_create_error is set by vlc_create
    m->rank = k;
    m->world = n_devices;
    g->members.push_back(m);
  }
  if (n_devices == 1) {  // a group of one is a plain context
    vlc_ctx* m = g->members[0];
    delete g;
    *out = m;
    return VLC_OK;
  }
  // NCCL serves a device list without repetitions (one rank per GPU).  A list that names a device twice -- the way the
  // group path is exercised on a single-GPU box -- and VLC_GROUP_NCCL=0 use peer copies instead.
  bool distinct = true;
  for (int a = 0; a < n_devices; ++a)
    for (int b = a + 1; b < n_devices; ++b) distinct = distinct && dev[a] != dev[b];
  const char* env = std::getenv("VLC_GROUP_NCCL");
  const bool want_nccl = distinct && !(env && env[0] == '0');
  if (want_nccl) {
    vlc::grp::Nccl& n = vlc::grp::Nccl::get();
    if (n.ok) {
      std::vector<vlc::grp::NcclComm> comms(n_devices, nullptr);
      const int r = n.CommInitAll(comms.data(), n_devices, dev.data());
      if (r == 0) {
        for (int k = 0; k < n_devices; ++k) g->members[k]->comm = comms[k];
        g->nccl = true;
      } else {
        g_create_error = std::string("ncclCommInitAll failed: ") + (n.GetErrorString ? n.GetErrorString(r) : "?");
        return undo(VLC_ERR_CUDA);
      }
    } else if (env && env[0] == '1') {
      g_create_error = "VLC_GROUP_NCCL=1 but " + n.why;
      return undo(VLC_ERR_CUDA);
    }
  }
  if (!g->nccl)  // peer copies between the members' buffers (cudaMemcpyPeerAsync stages through the host without it)
    for (int a = 0; a < n_devices; ++a)
      for (int b = 0; b < n_devices; ++b)
        if (dev[a] != dev[b]) {
          int can = 0;
          cudaDeviceCanAccessPeer(&can, dev[a], dev[b]);
          if (can) {
            cudaSetDevice(dev[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(dev[b], 0);
            if (e != cudaSuccess) cudaGetLastError();  // already enabled: fine
          }
        }
  cudaSetDevice(dev[0]);
  g->barrier = new vlc::grp::Barrier(n_devices);
  g->workers = new vlc::grp::Workers(n_devices);
  for (vlc_ctx* m : g->members) m->group = g;
  g->members[0]->is_leader = true;
  *out = g->members[0];
  return VLC_OK;
}

extern "C" int vlc_comm_unique_id(void* id128) {
  if (!id128) return VLC_ERR_ARG;
  vlc::grp::Nccl& n = vlc::grp::Nccl::get();
  if (!n.ok) {
    g_create_error = n.why;
    return VLC_ERR_STATE;
  }
  vlc::grp::NcclUniqueId id;
  const int r = n.GetUniqueId(&id);
  if (r != 0) {
    g_create_error = std::string("ncclGetUniqueId: ") + (n.GetErrorString ? n.GetErrorString(r) : "?");
    return VLC_ERR_CUDA;
  }
  std::memcpy(id128, &id, sizeof id);
  return VLC_OK;
}

extern "C" int vlc_comm_init_rank(vlc_ctx* c, int world, int rank, const void* id128) {
  CHECK_CTX(c);
  if (c->group) return fail(c, VLC_ERR_STATE, "vlc_comm_init_rank on a context made by vlc_create_multi");
  if (world < 1 || rank < 0 || rank >= world) return fail(c, VLC_ERR_ARG, "bad world / rank");
  int rc = bind_device(c);
  if (rc) return rc;
  if (c->comm) {
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    vlc::grp::Nccl::get().CommDestroy(c->comm);
    c->comm = nullptr;
  }
  c->world = 1;
  c->rank = 0;
  if (world == 1) return VLC_OK;
  if (!id128) return fail(c, VLC_ERR_ARG, "null unique id");
  vlc::grp::Nccl& n = vlc::grp::Nccl::get();
  if (!n.ok) return fail(c, VLC_ERR_STATE, n.why);
  vlc::grp::NcclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  const int r = n.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    c->comm = nullptr;
    return fail(c, VLC_ERR_CUDA, std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(r) : "?"));
  }
  c->world = world;
  c->rank = rank;
  return VLC_OK;
}

extern "C" int vlc_comm_info(const vlc_ctx* c, int* world, int* rank, int* transport) {
  if (!c) return VLC_ERR_ARG;
  if (world) *world = c->world;
  if (rank) *rank = c->rank;
  if (transport) *transport = c->world <= 1 ? 0 : (c->comm ? 1 : 2);  // 0 = single, 1 = NCCL, 2 = peer copies
  return VLC_OK;
}

extern "C" const char* vlc_last_error(const vlc_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

extern "C" int vlc_set_stream(vlc_ctx* c, void* s, int use_own) {
  CHECK_CTX(c);
  if (c->group && !use_own) return fail(c, VLC_ERR_STATE, "a caller's stream belongs to one device: a group runs on its members' own streams");
  VLC_GROUP(c, vlc_set_stream(m, nullptr, 1));
  c->stream = use_own ? c->own_stream : (cudaStream_t)s;  // s == NULL is the legacy default stream
  if (c->solver) cusolverDnSetStream(c->solver, c->stream);
  return VLC_OK;
}

extern "C" int vlc_sync(vlc_ctx* c) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_sync(m));
  int rc = bind_device(c);
  if (rc) return rc;
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_device_info(vlc_ctx* c, int* sm, int* maj, int* min, int64_t* mem) {
  CHECK_CTX(c);
  if (sm) *sm = c->sm_count;
  if (maj) *maj = c->cc_major;
  if (min) *min = c->cc_minor;
  if (mem) *mem = c->mem_bytes;
  return VLC_OK;
}

extern "C" int vlc_set_tuning(vlc_ctx* c, int T, int nsplit) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_set_tuning(m, T, nsplit));
  if (T < 0 || T > 4 || nsplit < 0) return fail(c, VLC_ERR_ARG, "targets_per_thread in 0..4, nsplit >= 0");
  c->tune_T = T;
  c->tune_nsplit = nsplit;
  return VLC_OK;
}

extern "C" int vlc_set_precision(vlc_ctx* c, int mode) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_set_precision(m, mode));
  if (mode != 0 && mode != 1) return fail(c, VLC_ERR_ARG, "precision mode must be 0 (full) or 1 (fast)");
  c->fast = (mode == 1);
  return VLC_OK;
}

extern "C" int64_t vlc_launch_count(const vlc_ctx* c) {
  if (!c) return 0;
  if (c->group && c->is_leader) {  // all members' kernels
    long long n = 0;
    for (const vlc_ctx* m : c->group->members) n += m->launches;
    return n;
  }
  return c->launches;
}

extern "C" int vlc_source_tile(void) { return kTile; }

// ============================================================================ tier 1

extern "C" int vlc_set_sources_dev(vlc_ctx* c, int set, int64_t n, const double* p1, const double* p2,
                                   const double* rvc, const double* gam, const uint8_t* flag) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = check_set(c, set))) return rc;
  if (n < 0) return fail(c, VLC_ERR_ARG, "n < 0");
  if (n > 0 && (!p1 || !p2 || !rvc || !gam)) return fail(c, VLC_ERR_ARG, "null source array");
  SourceSet& s = c->sets[set];
  const long long n_pad = pad_tile(n);
  if ((rc = reserve(c, s.rec, (size_t)n_pad * vlc::kSrcDoubles))) return rc;
  if (n_pad > 0) {
    vlc::pack_flat_kernel<<<blocks_for(n_pad, 256), 256, 0, c->stream>>>(n, n_pad, p1, p2, rvc, gam, flag, s.rec.p);
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
  }
  s.n = n;
  s.n_pad = n_pad;
  s.has_shared = false;
  s.n_lat = s.n_lat_pad = s.n_rem = s.n_rem_pad = s.n_lat2 = s.n_lat2_pad = 0;
  return VLC_OK;
}

extern "C" int vlc_set_sources(vlc_ctx* c, int set, int64_t n, const double* p1, const double* p2, const double* rvc,
                               const double* gam, const uint8_t* flag) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_set_sources(m, set, n, p1, p2, rvc, gam, flag));
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = check_set(c, set))) return rc;
  if (n < 0) return fail(c, VLC_ERR_ARG, "n < 0");
  if (n == 0) return vlc_set_sources_dev(c, set, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (!p1 || !p2 || !rvc || !gam) return fail(c, VLC_ERR_ARG, "null source array");
  if ((rc = reserve(c, c->scratch, 8 * (size_t)n))) return rc;
  double* d = c->scratch.p;
  cudaStream_t st = c->stream;
  CUDA_OK(c, cudaMemcpyAsync(d, p1, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
  CUDA_OK(c, cudaMemcpyAsync(d + 3 * n, p2, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
  CUDA_OK(c, cudaMemcpyAsync(d + 6 * n, rvc, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  CUDA_OK(c, cudaMemcpyAsync(d + 7 * n, gam, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  const uint8_t* dflag = nullptr;
  if (flag) {
    if (c->flag_cap < (size_t)n) {
      if (c->d_flag) {
        CUDA_OK(c, cudaStreamSynchronize(st));
        CUDA_OK(c, cudaFree(c->d_flag));
        c->d_flag = nullptr;
      }
      CUDA_OK(c, cudaMalloc(&c->d_flag, (size_t)n + 1024));
      c->flag_cap = (size_t)n + 1024;
    }
    CUDA_OK(c, cudaMemcpyAsync(c->d_flag, flag, (size_t)n, cudaMemcpyHostToDevice, st));
    dflag = c->d_flag;
  }
  rc = vlc_set_sources_dev(c, set, n, d, d + 3 * n, d + 6 * n, d + 7 * n, dflag);
  if (rc) return rc;
  CUDA_OK(c, cudaStreamSynchronize(st));
  return VLC_OK;
}

extern "C" int64_t vlc_num_sources(const vlc_ctx* c, int set) {
  if (!c || set < 0 || set >= VLC_MAX_SETS) return -1;
  return c->sets[set].n;
}

extern "C" int vlc_vind_dev(vlc_ctx* c, int set, int64_t m, const double* dP, double* dV) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = check_set(c, set))) return rc;
  if (m < 0) return fail(c, VLC_ERR_ARG, "m < 0");
  if (m > 0 && (!dP || !dV)) return fail(c, VLC_ERR_ARG, "null target / output pointer");
  const SourceSet& s = c->sets[set];
  if (s.has_shared && c->shared_nodes && s.n_lat_pad > 0) return sweep_shared(c, s, m, dP, dV);
  return sweep(c, s.rec.p, s.n_pad, m, dP, dV);
}

extern "C" int vlc_vind_range_dev(vlc_ctx* c, int set, int64_t first, int64_t count, int64_t m, const double* dP,
                                  double* dV) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = check_set(c, set))) return rc;
  const SourceSet& s = c->sets[set];
  if (first < 0 || count < 0 || first % kTile != 0 || first + count > s.n_pad)
    return fail(c, VLC_ERR_ARG, "source range must start on a tile boundary and lie inside the set");
  if (m > 0 && (!dP || !dV)) return fail(c, VLC_ERR_ARG, "null target / output pointer");
  long long cnt_pad = pad_tile(count);
  if (first + cnt_pad > s.n_pad) cnt_pad = s.n_pad - first;
  if (cnt_pad != count && first + count < s.n)
    return fail(c, VLC_ERR_ARG, "an interior source range must be a whole number of tiles");
  return sweep(c, s.rec.p + (size_t)first * vlc::kSrcDoubles, cnt_pad, m, dP, dV);
}

extern "C" int vlc_vind(vlc_ctx* c, int set, int64_t m, const double* P, double* V) {
  CHECK_CTX(c);
  if (c->group && c->is_leader && !t_in_member) return group_all(c, [&](vlc_ctx* g_) -> int { return vlc_vind(g_, set, m, P, V); });
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = check_set(c, set))) return rc;
  if (m < 0) return fail(c, VLC_ERR_ARG, "m < 0");
  const SourceSet& s = c->sets[set];
  const bool sh = s.has_shared && c->shared_nodes && s.n_lat_pad > 0;
  return sweep_host(c, s.rec.p, s.n_pad, m, P, V, sh ? &s : nullptr);
}

// ============================================================================ tier 2

extern "C" int vlc_rotor_define(vlc_ctx* c, int ir, int nb, int nc, int ns, int nNwake, int nFwake, int surfaceType) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_define(m, ir, nb, nc, ns, nNwake, nFwake, surfaceType));
  if (ir < 0 || ir > 1023) return fail(c, VLC_ERR_ARG, "rotor index out of range");
  if (nb < 1 || nc < 1 || ns < 1 || nNwake < 0 || nFwake < 0) return fail(c, VLC_ERR_ARG, "bad rotor sizes");
  if ((int)c->rotors.size() <= ir) c->rotors.resize(ir + 1);
  Rotor& r = c->rotors[ir];
  r.defined = true;
  r.nb = nb;
  r.nc = nc;
  r.ns = ns;
  r.nNwake = nNwake;
  r.nFwake = nFwake;
  r.surfaceType = surfaceType == 0 ? 1 : surfaceType;  // classdef.f90:3023
  r.rowNear = nNwake + 1;                              // main.f90:228-230
  r.rowFar = nFwake + 1;
  if (r.d_ipiv && r.N != nc * ns * nb) {  // re-definition with another size: the pivot array is sized by N
    cudaStreamSynchronize(c->stream);
    cudaFree(r.d_ipiv);
    r.d_ipiv = nullptr;
  }
  if (r.d_axi) {  // sized by nb
    cudaStreamSynchronize(c->stream);
    cudaFree(r.d_axi);
    r.d_axi = nullptr;
  }
  r.N = nc * ns * nb;
  r.dirty[0] = r.dirty[1] = r.bound_dirty = r.chord_dirty = true;
  r.factored = false;
  r.have_pf[0] = r.have_pf[1] = false;
  for (int s2 = 0; s2 < 2; ++s2) {
    r.stale_near[s2].assign(nb, 1);  // zero-filled below = gam 0 everywhere, like rotor_init (:3835-3836)
    r.stale_far[s2].assign(nb, 1);
  }
  r.nbConvect = nb;
  r.have_rhs = false;
  r.have_sec.assign(nb, 0);
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = reserve(c, r.rhs, (size_t)r.N + 1)) || (rc = reserve(c, r.gamvec, (size_t)r.N + 1))) return rc;
  if ((rc = reserve(c, r.sec, (size_t)nb * vlc::cp::sec_doubles(ns))) || (rc = reserve(c, r.loads, (size_t)nb * vlc::cp::loads_doubles(ns))))
    return rc;
  for (int k = 0; k < 2; ++k) {  // vel2 / vel3 of an earlier definition start from zero again
    if (r.velNx[k].p) CUDA_OK(c, cudaMemsetAsync(r.velNx[k].p, 0, r.velNx[k].cap * sizeof(double), c->stream));
    if (r.velFx[k].p) CUDA_OK(c, cudaMemsetAsync(r.velFx[k].p, 0, r.velFx[k].cap * sizeof(double), c->stream));
  }
  CUDA_OK(c, cudaMemsetAsync(r.gamvec.p, 0, r.gamvec.cap * sizeof(double), c->stream));
  CUDA_OK(c, cudaMemsetAsync(r.sec.p, 0, r.sec.cap * sizeof(double), c->stream));
  CUDA_OK(c, cudaMemsetAsync(r.loads.p, 0, r.loads.cap * sizeof(double), c->stream));
  for (int k = 0; k < 4; ++k) {  // velNwake etc. start at zero like rotor_init (classdef.f90:3795-3812)
    if ((rc = reserve(c, r.velN[k], (size_t)3 * nNwake * (ns + 1) * nb + 1))) return rc;
    if ((rc = reserve(c, r.velF[k], (size_t)3 * nFwake * nb + 1))) return rc;
    CUDA_OK(c, cudaMemsetAsync(r.velN[k].p, 0, r.velN[k].cap * sizeof(double), c->stream));
    CUDA_OK(c, cudaMemsetAsync(r.velF[k].p, 0, r.velF[k].cap * sizeof(double), c->stream));
  }
  {  // packed sets at their final size (all nNwake / nFwake rows active): no allocation while the wake grows
    const long long wing_pad = pad_tile(4LL * nc * ns * nb);
    const long long wake_n = (4LL * nNwake * ns + ns + nFwake + VLC_NPFWAKE) * nb;
    const long long rem_n = ((long long)nNwake + ns + nFwake + VLC_NPFWAKE) * nb;
    size_t lat_doubles = 0;
    for (int W = 1; W <= 4; ++W)
      lat_doubles = std::max(lat_doubles, (size_t)pad_lat((long long)nb * ((ns + W - 1) / W) * (nNwake + 1), W) * lat_rd_of(W));
    for (int s = 0; s < 2 && nNwake > 0; ++s) {
      if ((rc = reserve(c, r.comb[s].rec, (size_t)(wing_pad + pad_tile(wake_n)) * vlc::kSrcDoubles))) return rc;
      if ((rc = reserve(c, r.comb[s].lat, lat_doubles))) return rc;
      if ((rc = reserve(c, r.comb[s].lat2, (size_t)pad_lat((long long)nb * (nNwake + 1), 3) * lat_rd_of(3)))) return rc;
      if ((rc = reserve(c, r.comb[s].rem, (size_t)(wing_pad + pad_tile(rem_n)) * vlc::kSrcDoubles))) return rc;
    }
  }
  // zero-initialised device copies so that never-uploaded rows hold gam = 0 like rotor_init (:3835-3836)
  if ((rc = reserve(c, r.wiP, (size_t)nb * nc * ns * vlc::kWp))) return rc;
  CUDA_OK(c, cudaMemsetAsync(r.wiP.p, 0, r.wiP.cap * sizeof(double), c->stream));
  if ((rc = reserve(c, r.waN_alt, (size_t)nb * nNwake * ns * vlc::kVr + 1))) return rc;  // shiftwake's second buffer
  for (int s = 0; s < 2; ++s) {
    if ((rc = reserve(c, r.waN[s], (size_t)nb * nNwake * ns * vlc::kVr + 1))) return rc;
    if ((rc = reserve(c, r.waF[s], (size_t)nb * nFwake * vlc::kFw + 1))) return rc;
    if ((rc = reserve(c, r.wapF[s], (size_t)nb * VLC_NPFWAKE * vlc::kFw))) return rc;
    CUDA_OK(c, cudaMemsetAsync(r.waN[s].p, 0, r.waN[s].cap * sizeof(double), c->stream));
    CUDA_OK(c, cudaMemsetAsync(r.waF[s].p, 0, r.waF[s].cap * sizeof(double), c->stream));
    CUDA_OK(c, cudaMemsetAsync(r.wapF[s].p, 0, r.wapF[s].cap * sizeof(double), c->stream));
    if ((rc = reserve(c, r.pfHelix[s], (size_t)2 * nb))) return rc;  // helixPitch = helixRadius = 0 (classdef.f90:228-229)
    CUDA_OK(c, cudaMemsetAsync(r.pfHelix[s].p, 0, r.pfHelix[s].cap * sizeof(double), c->stream));
  }
  if ((rc = reserve(c, r.pfFits, (size_t)nb * (sizeof(vlc::pf::Fit) / sizeof(double))))) return rc;
  return VLC_OK;
}

extern "C" int vlc_rotors_clear(vlc_ctx* c) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotors_clear(m));
  for (auto& r : c->rotors) r.defined = false;  // buffers are kept for the next definition
  return VLC_OK;
}

extern "C" int vlc_rotor_set_rows(vlc_ctx* c, int ir, int rowNear, int rowFar) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_set_rows(m, ir, rowNear, rowFar));
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (rowNear < 1 || rowNear > r->nNwake + 1 || rowFar < 1 || rowFar > r->nFwake + 1)
    return fail(c, VLC_ERR_ARG, "rowNear / rowFar out of range");
  if (rowNear != r->rowNear || rowFar != r->rowFar) r->dirty[0] = r->dirty[1] = true;
  r->rowNear = rowNear;
  r->rowFar = rowFar;
  return VLC_OK;
}

extern "C" int vlc_rotor_put_wing(vlc_ctx* c, int ir, int ib, const double* wiP) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_wing(m, ir, ib, wiP));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !wiP) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const size_t per = (size_t)r->nc * r->ns * vlc::kWp;
  mark_wing_dirty(*r);
  // The LU factors stay valid: the reference computes AIC once, before the time loop, and keeps using it while the
  // wing moves rigidly (main.f90:65-81); only vlc_rotor_calcAIC replaces them.
  return upload(c, r->wiP, per * r->nb, per * ib, wiP, per);
}

extern "C" int vlc_rotor_put_wing_gam(vlc_ctx* c, int ir, int ib, const double* gam) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_wing_gam(m, ir, ib, gam));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !gam) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const size_t np = (size_t)r->nc * r->ns;
  CUDA_OK(c, cudaMemcpy2DAsync(r->wiP.p + (size_t)ib * np * vlc::kWp + vlc::kVrGam, vlc::kWp * sizeof(double), gam,
                               sizeof(double), sizeof(double), np, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  mark_wing_dirty(*r);
  return VLC_OK;
}

extern "C" int vlc_rotor_put_nwake(vlc_ctx* c, int ir, int ib, int predicted, const double* waN) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_nwake(m, ir, ib, predicted, waN));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !waN) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const int s = predicted ? 1 : 0;
  const size_t per = (size_t)r->nNwake * r->ns * vlc::kVr;
  r->dirty[s] = true;
  if (per == 0) return VLC_OK;
  // only the active rows rowNear..nNwake of every column travel (a sweep never reads the others, classdef.f90:1450)
  const int first = r->rowNear - 1, nact = r->nNwake - first;
  r->stale_near[s][ib] = r->rowNear;
  if (nact <= 0) return VLC_OK;
  if ((rc = reserve(c, r->waN[s], per * r->nb))) return rc;
  TimerScope ts(4);
  const size_t pitch = (size_t)r->nNwake * vlc::kVr * sizeof(double);
  CUDA_OK(c, cudaMemcpy2DAsync(r->waN[s].p + per * ib + (size_t)first * vlc::kVr, pitch, waN + (size_t)first * vlc::kVr, pitch,
                               (size_t)nact * vlc::kVr * sizeof(double), (size_t)r->ns, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));  // the caller may reuse its buffer right away
  return VLC_OK;
}

extern "C" int vlc_rotor_put_fwake(vlc_ctx* c, int ir, int ib, int predicted, const double* waF) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_fwake(m, ir, ib, predicted, waF));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb) return fail(c, VLC_ERR_ARG, "bad blade index");
  const int s = predicted ? 1 : 0;
  const size_t per = (size_t)r->nFwake * vlc::kFw;
  r->dirty[s] = true;
  if (per == 0) return VLC_OK;
  if (!waF) return fail(c, VLC_ERR_ARG, "null pointer");
  const int first = r->rowFar - 1, nact = r->nFwake - first;  // active rows rowFar..nFwake (classdef.f90:1465)
  r->stale_far[s][ib] = r->rowFar;
  if (nact <= 0) return VLC_OK;
  return upload(c, r->waF[s], per * r->nb, per * ib + (size_t)first * vlc::kFw, waF + (size_t)first * vlc::kFw,
                (size_t)nact * vlc::kFw);
}

extern "C" int vlc_rotor_put_pfwake(vlc_ctx* c, int ir, int ib, int predicted, const double* wapF) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_pfwake(m, ir, ib, predicted, wapF));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !wapF) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const int s = predicted ? 1 : 0;
  const size_t per = (size_t)VLC_NPFWAKE * vlc::kFw;
  r->dirty[s] = true;
  r->have_pf[s] = true;
  return upload(c, r->wapF[s], per * r->nb, per * ib, wapF, per);
}

extern "C" int vlc_rotor_vind_bywing(vlc_ctx* c, int ir, int64_t m, const double* P, double* V) {
  CHECK_CTX(c);
  if (c->group && c->is_leader && !t_in_member) return group_all(c, [&](vlc_ctx* g_) -> int { return vlc_rotor_vind_bywing(g_, ir, m, P, V); });
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if ((rc = pack_rotor(c, *r, 0))) return rc;
  return sweep_host(c, r->comb[0].rec.p, r->wing_pad[0], m, P, V);
}

extern "C" int vlc_rotor_vind_bywake(vlc_ctx* c, int ir, int predicted, int64_t m, const double* P, double* V) {
  CHECK_CTX(c);
  if (c->group && c->is_leader && !t_in_member) return group_all(c, [&](vlc_ctx* g_) -> int { return vlc_rotor_vind_bywake(g_, ir, predicted, m, P, V); });
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const int s = predicted ? 1 : 0;
  if ((rc = pack_rotor(c, *r, s))) return rc;
  const SourceSet v = wake_view(*r, s);
  return sweep_host(c, v.rec.p, v.n_pad, m, P, V, (v.has_shared && c->shared_nodes) ? &v : nullptr);
}

extern "C" int vlc_rotor_vind_bywing_boundVortices(vlc_ctx* c, int ir, int64_t m, const double* P, double* V) {
  CHECK_CTX(c);
  if (c->group && c->is_leader && !t_in_member) return group_all(c, [&](vlc_ctx* g_) -> int { return vlc_rotor_vind_bywing_boundVortices(g_, ir, m, P, V); });
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if ((rc = pack_bound(c, *r))) return rc;
  return sweep_host(c, r->bound.rec.p, r->bound.n_pad, m, P, V);
}

extern "C" int vlc_rotor_vind_bywing_chordwiseVortices(vlc_ctx* c, int ir, int64_t m, const double* P, double* V) {
  CHECK_CTX(c);
  if (c->group && c->is_leader && !t_in_member) return group_all(c, [&](vlc_ctx* g_) -> int { return vlc_rotor_vind_bywing_chordwiseVortices(g_, ir, m, P, V); });
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if ((rc = pack_chord(c, *r))) return rc;
  return sweep_host(c, r->chord.rec.p, r->chord.n_pad, m, P, V);
}

extern "C" int vlc_rotor_vind(vlc_ctx* c, int ir, int predicted, int64_t m, const double* P, double* V) {
  CHECK_CTX(c);
  if (c->group && c->is_leader && !t_in_member) return group_all(c, [&](vlc_ctx* g_) -> int { return vlc_rotor_vind(g_, ir, predicted, m, P, V); });
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const int s = predicted ? 1 : 0;
  {
    TimerScope ts(1);
    if ((rc = pack_rotor(c, *r, s))) return rc;
  }
  const SourceSet& v = r->comb[s];
  return sweep_host(c, v.rec.p, v.n_pad, m, P, V, (v.has_shared && c->shared_nodes) ? &v : nullptr);
}

extern "C" int vlc_vind_onNwake_byRotor(vlc_ctx* c, int ir, const double* Nwake, int rows, int cols, int ld,
                                        int predicted, double* vindArray) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_vind_onNwake_byRotor(m, ir, Nwake, rows, cols, ld, predicted, vindArray));
  if (rows < 0 || cols < 1 || ld < rows) return fail(c, VLC_ERR_ARG, "bad Nwake slice shape");
  if (rows == 0) return VLC_OK;
  if (!Nwake || !vindArray) return fail(c, VLC_ERR_ARG, "null pointer");
  // targets in the reference's order (libCommon.f90:133-145): corner 2 of ring (i,j) = vf(2)%fc(:,1),
  // then corner 3 of the last column = vf(3)%fc(:,1)
  double* P = nullptr;  // pinned, grow-only: no allocation, no page faults and an asynchronous H2D per call
  {
    int rc0 = bind_device(c);
    if (rc0) return rc0;
    if ((rc0 = host_targets(c, (size_t)3 * rows * (cols + 1), &P))) return rc0;
  }
  {
    TimerScope ts(0);
    // a member of a group / a rank of a communicator sweeps only its slice of the targets (sweep_host): gather that slice
    const long long mt = (long long)rows * (cols + 1);
    const vlc::grp::Shard sh = c->world > 1 ? vlc::grp::shard_range(mt, c->world, c->rank) : vlc::grp::shard_range(mt, 1, 0);
    host_parallel(c, sh.count(), [&](long long lo, long long hi) {
      for (long long q = sh.lo + lo; q < sh.lo + hi; ++q) {
        const long long j = q / rows, i = q - j * rows;
        const bool last = j == cols;  // corner 3 of the last column
        const double* rec = Nwake + (size_t)VLC_VR_DOUBLES * ((size_t)i + (size_t)ld * (last ? cols - 1 : j));
        __builtin_prefetch(rec + 16 * VLC_VR_DOUBLES + VLC_VF_DOUBLES);  // one cache line per record, 400 bytes apart: 16 rows ahead
        std::memcpy(&P[3 * (size_t)q], rec + VLC_VF_DOUBLES * (last ? 2 : 1), 3 * sizeof(double));
      }
    });
  }
  return vlc_rotor_vind(c, ir, predicted, (int64_t)rows * (cols + 1), P, vindArray);
}

extern "C" int vlc_vind_onFwake_byRotor(vlc_ctx* c, int ir, const double* Fwake, int rows, int predicted,
                                        double* vindArray) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_vind_onFwake_byRotor(m, ir, Fwake, rows, predicted, vindArray));
  if (rows < 0) return fail(c, VLC_ERR_ARG, "rows < 0");
  if (rows == 0) return VLC_OK;
  if (!Fwake || !vindArray) return fail(c, VLC_ERR_ARG, "null pointer");
  double* P = nullptr;
  {
    int rc0 = bind_device(c);
    if (rc0) return rc0;
    if ((rc0 = host_targets(c, (size_t)3 * rows, &P))) return rc0;
  }
  const vlc::grp::Shard shf = c->world > 1 ? vlc::grp::shard_range(rows, c->world, c->rank) : vlc::grp::shard_range(rows, 1, 0);
  host_parallel(c, shf.count(), [&](long long lo, long long hi) {
    for (long long i = shf.lo + lo; i < shf.lo + hi; ++i)
      std::memcpy(&P[3 * (size_t)i], Fwake + (size_t)VLC_FWAKE_DOUBLES * i, 3 * sizeof(double));
  });
  return vlc_rotor_vind(c, ir, predicted, rows, P, vindArray);
}

// ---------------------------------------------------------------------------- AIC

namespace {
int ensure_solver(vlc_ctx* c) {
  if (c->solver) return VLC_OK;
  cusolverStatus_t st = cusolverDnCreate(&c->solver);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(c, VLC_ERR_CUDA, "cusolverDnCreate failed: " + std::to_string((int)st));
  cusolverDnSetStream(c->solver, c->stream);
  return VLC_OK;
}
}  // namespace

extern "C" int vlc_rotor_calcAIC(vlc_ctx* c, int ir, double* AIC_out) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_calcAIC(m, ir, m->rank == 0 ? AIC_out : nullptr));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if ((rc = ensure_solver(c))) return rc;
  const int N = r->N;
  if ((rc = reserve(c, r->LU, (size_t)N * N))) return rc;
  if (!r->d_ipiv) CUDA_OK(c, cudaMalloc(&r->d_ipiv, sizeof(int) * (size_t)N));
  if (!r->d_info) CUDA_OK(c, cudaMalloc(&r->d_info, sizeof(int)));
  dim3 grid(blocks_for(N, 128), N, 1);
  vlc::aic_assemble_kernel<<<grid, 128, 0, c->stream>>>(r->wiP.p, N, r->LU.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  if (AIC_out) {
    CUDA_OK(c, cudaMemcpyAsync(AIC_out, r->LU.p, sizeof(double) * (size_t)N * N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  int lwork = 0;
  cusolverStatus_t st = cusolverDnDgetrf_bufferSize(c->solver, N, N, r->LU.p, N, &lwork);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(c, VLC_ERR_CUDA, "cusolverDnDgetrf_bufferSize failed");
  if ((rc = reserve(c, c->solver_work, (size_t)lwork + 1))) return rc;
  st = cusolverDnDgetrf(c->solver, N, N, r->LU.p, N, c->solver_work.p, r->d_ipiv, r->d_info);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(c, VLC_ERR_CUDA, "cusolverDnDgetrf failed: " + std::to_string((int)st));
  int info = 0;
  CUDA_OK(c, cudaMemcpyAsync(&info, r->d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  c->launches++;
  if (info != 0) return fail(c, VLC_ERR_SINGULAR, "Matrix is numerically singular!");  // libMath.f90:73
  // AIC_inv = inv2(AIC) once (classdef.f90:4178, libMath.f90:48-83: getrf + getri); here getrs on the identity
  const size_t NN = (size_t)N * N;
  if ((rc = reserve(c, r->Ainv, NN))) return rc;
  vlc::identity_kernel<<<blocks_for((long long)NN, 256), 256, 0, c->stream>>>(N, r->Ainv.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  st = cusolverDnDgetrs(c->solver, CUBLAS_OP_N, N, N, r->LU.p, N, r->d_ipiv, r->Ainv.p, N, r->d_info);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(c, VLC_ERR_CUDA, "cusolverDnDgetrs failed: " + std::to_string((int)st));
  c->launches++;
  r->factored = true;
  return VLC_OK;
}

namespace {
// gamVec = AIC_inv . RHS: the reference's per-step operation (main.f90:190, :596), one launch
int solve_dev(vlc_ctx* c, Rotor* r, const double* d_rhs, double* d_gam) {
  dim3 block(vlc::kGemvRows, vlc::kGemvSlices, 1);
  vlc::ainv_gemv_kernel<<<blocks_for(r->N, vlc::kGemvRows), block, 0, c->stream>>>(r->N, r->Ainv.p, d_rhs, d_gam);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  return VLC_OK;
}
}  // namespace

extern "C" int vlc_rotor_solve(vlc_ctx* c, int ir, const double* RHS, double* gamVec) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (!r->factored) return fail(c, VLC_ERR_STATE, "vlc_rotor_solve before vlc_rotor_calcAIC");
  if (!RHS || !gamVec) return fail(c, VLC_ERR_ARG, "null pointer");
  if ((rc = reserve(c, c->stage_V, 2 * (size_t)r->N))) return rc;
  CUDA_OK(c, cudaMemcpyAsync(c->stage_V.p, RHS, sizeof(double) * r->N, cudaMemcpyHostToDevice, c->stream));
  if ((rc = solve_dev(c, r, c->stage_V.p, c->stage_V.p + r->N))) return rc;
  CUDA_OK(c, cudaMemcpyAsync(gamVec, c->stage_V.p + r->N, sizeof(double) * r->N, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_rotor_get_AIC_inv(vlc_ctx* c, int ir, double* AIC_inv) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (!r->factored) return fail(c, VLC_ERR_STATE, "vlc_rotor_get_AIC_inv before vlc_rotor_calcAIC");
  if (!AIC_inv) return fail(c, VLC_ERR_ARG, "null pointer");
  const size_t NN = (size_t)r->N * r->N;
  CUDA_OK(c, cudaMemcpyAsync(AIC_inv, r->Ainv.p, sizeof(double) * NN, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

// ============================================================================ tier 3

#define LAUNCH1D(c, kern, n, ...)                                                      \
  do {                                                                                 \
    if ((n) > 0) {                                                                     \
      kern<<<blocks_for((n), 256), 256, 0, (c)->stream>>>(__VA_ARGS__);                \
      CUDA_OK((c), cudaGetLastError());                                                \
      (c)->launches++;                                                                 \
    }                                                                                  \
  } while (0)

// ============================================================================ tier 2b: device-resident stepping

namespace {

// getTransformAxis (libMath.f90:695-726): rotation by theta about axisVec, column-major 3x3
void transform_axis(double theta, const double axisVec[3], double T[9]) {
  const double n = std::sqrt(axisVec[0] * axisVec[0] + axisVec[1] * axisVec[1] + axisVec[2] * axisVec[2]);
  const double ax[3] = {axisVec[0] / n, axisVec[1] / n, axisVec[2] / n};
  const double ct = std::cos(theta), st = std::sin(theta), omct = 1.0 - ct;
  T[0] = ct + ax[0] * ax[0] * omct;
  T[1] = ax[2] * st + ax[1] * ax[0] * omct;
  T[2] = -ax[1] * st + ax[2] * ax[0] * omct;
  T[3] = -ax[2] * st + ax[0] * ax[1] * omct;
  T[4] = ct + ax[1] * ax[1] * omct;
  T[5] = ax[0] * st + ax[2] * ax[1] * omct;
  T[6] = ax[1] * st + ax[0] * ax[2] * omct;
  T[7] = -ax[0] * st + ax[1] * ax[2] * omct;
  T[8] = ct + ax[2] * ax[2] * omct;
}

inline long long wake_targets_of(const Rotor& r) {  // targets of one rotor's convected blades in a wake sweep
  if (r.nNwake <= 0) return 0;
  const long long nact = r.nNwake - r.rowNear + 1 > 0 ? r.nNwake - r.rowNear + 1 : 0;
  const long long nfar = r.nFwake - r.rowFar + 1 > 0 ? r.nFwake - r.rowFar + 1 : 0;
  return (nact * (r.ns + 1) + nfar) * r.nbConvect;
}

}  // namespace

extern "C" int vlc_rotor_set_wake_params(vlc_ctx* c, int ir, int nbConvect, int axisymmetrySwitch, int ductSwitch,
                                         int suppressFwakeSwitch, int rollupStart, int rollupEnd, double rollupSign,
                                         double apparentViscCoeff, double decayCoeff, double initWakeVel) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_set_wake_params(m, ir, nbConvect, axisymmetrySwitch, ductSwitch, suppressFwakeSwitch, rollupStart, rollupEnd, rollupSign, apparentViscCoeff, decayCoeff, initWakeVel));
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (nbConvect < 1 || nbConvect > r->nb) return fail(c, VLC_ERR_ARG, "nbConvect out of range");
  if (r->nNwake > 0 && (rollupStart < 1 || rollupEnd > r->ns)) return fail(c, VLC_ERR_ARG, "rollupStart / rollupEnd out of range");
  r->nbConvect = nbConvect;
  r->axisym = axisymmetrySwitch;
  r->duct = ductSwitch;
  r->suppressFwake = suppressFwakeSwitch;
  r->rollupStart = rollupStart;
  r->rollupEnd = rollupEnd;
  r->sgnPositive = std::copysign(1.0, rollupSign) > 2.220446049250313e-16 ? 1 : 0;  // classdef.f90:4541
  r->apparentViscCoeff = apparentViscCoeff;
  r->decayCoeff = decayCoeff;
  r->initWakeVel = initWakeVel;
  return VLC_OK;
}

extern "C" int vlc_rotor_set_frame(vlc_ctx* c, int ir, const double* shaftAxis, const double* hubCoords) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_set_frame(m, ir, shaftAxis, hubCoords));
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (!shaftAxis || !hubCoords) return fail(c, VLC_ERR_ARG, "null pointer");
  for (int k = 0; k < 3; ++k) {
    r->shaftAxis[k] = shaftAxis[k];
    r->hubCoords[k] = hubCoords[k];
  }
  return VLC_OK;
}

extern "C" int vlc_rotor_assignshed(vlc_ctx* c, int ir, int edge) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_assignshed(m, ir, edge));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (edge != 0 && edge != 1) return fail(c, VLC_ERR_ARG, "edge: 0 = 'LE', 1 = 'TE' (classdef.f90:4303, :4313)");
  if (r->nNwake <= 0) return VLC_OK;
  // 'LE' writes row rowNear; 'TE' writes row max(rowNear - 1, 1): also legal with rowNear = nNwake + 1, which is how the
  // reference pre-sheds the first row before the time loop (main.f90:227-234)
  if (r->rowNear < 1 || r->rowNear > r->nNwake + (edge == 1 ? 1 : 0))
    return fail(c, VLC_ERR_STATE, "assignshed: rowNear outside 1..nNwake");
  const int n = r->nb * r->ns;
  vlc::rec_assignshed_kernel<<<blocks_for(n, 128), 128, 0, c->stream>>>(edge, r->nb, r->nc, r->ns, r->nNwake, r->rowNear,
                                                                        r->wiP.p, r->waN[0].p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  r->dirty[0] = true;
  return VLC_OK;
}

extern "C" int vlc_rotor_age_wake(vlc_ctx* c, int ir, double dt, double omegaSlow) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_age_wake(m, ir, dt, omegaSlow));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (r->nNwake <= 0) return VLC_OK;
  const long long nact = std::max(0, r->nNwake - r->rowNear + 1), nfar = std::max(0, r->nFwake - r->rowFar + 1);
  const long long n = (long long)r->nb * (r->ns * nact + nfar);
  if (n <= 0) return VLC_OK;
  vlc::rec_age_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nb, r->ns, r->nNwake, r->nFwake, r->rowNear, r->rowFar, dt,
                                                                 dt * omegaSlow, r->waN[0].p, r->waF[0].p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  return VLC_OK;  // ages are not read by any sweep: the packed sets stay valid
}

extern "C" int vlc_rotor_dissipate_wake(vlc_ctx* c, int ir, double dt, double kinematicVisc) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_dissipate_wake(m, ir, dt, kinematicVisc));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (r->nNwake <= 0) return VLC_OK;
  const double oseenParameter = 1.2564;  // classdef.f90:4361
  const double growTerm = 4.0 * oseenParameter * r->apparentViscCoeff * kinematicVisc * dt;
  const double decayFactor = std::exp(-r->decayCoeff * dt);
  const long long nact = std::max(0, r->nNwake - r->rowNear + 1), nfar = std::max(0, r->nFwake - r->rowFar + 1);
  const long long n = (long long)r->nb * (r->ns * nact + nfar);
  if (n <= 0) return VLC_OK;
  vlc::rec_dissipate_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nb, r->ns, r->nNwake, r->nFwake, r->rowNear, r->rowFar,
                                                                       growTerm, decayFactor, r->waN[0].p, r->waF[0].p);
  c->launches++;
  const long long n4 = (long long)r->nb * r->ns * (nact - 1);
  if (n4 > 0) {
    vlc::rec_dissipate_vf4_kernel<<<blocks_for(n4, 256), 256, 0, c->stream>>>(r->nb, r->ns, r->nNwake, r->rowNear, r->waN[0].p);
    c->launches++;
  }
  CUDA_OK(c, cudaGetLastError());
  r->dirty[0] = true;
  return VLC_OK;
}

extern "C" int vlc_rotor_strain_wake(vlc_ctx* c, int ir) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_strain_wake(m, ir));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const int nfar = r->nFwake - r->rowFar + 1;
  if (r->nNwake <= 0 || nfar <= 0) return VLC_OK;
  vlc::rec_strain_kernel<<<blocks_for(r->nb * nfar, 128), 128, 0, c->stream>>>(r->nb, r->nFwake, r->rowFar, r->waF[0].p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  r->dirty[0] = true;
  return VLC_OK;
}

extern "C" int vlc_rotor_wake_to_predicted(vlc_ctx* c, int ir) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_wake_to_predicted(m, ir));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (r->nNwake <= 0) return VLC_OK;
  const int firstN = r->rowNear - 1, nactN = r->nNwake - firstN, firstF = r->rowFar - 1, nactF = r->nFwake - firstF;
  if (nactN > 0) {  // columns of all convected blades are contiguous: one strided copy
    const size_t pitch = (size_t)r->nNwake * vlc::kVr * sizeof(double);
    CUDA_OK(c, cudaMemcpy2DAsync(r->waN[1].p + (size_t)firstN * vlc::kVr, pitch, r->waN[0].p + (size_t)firstN * vlc::kVr, pitch,
                                 (size_t)nactN * vlc::kVr * sizeof(double), (size_t)r->ns * r->nbConvect,
                                 cudaMemcpyDeviceToDevice, c->stream));
  }
  if (nactF > 0) {
    const size_t pitch = (size_t)r->nFwake * vlc::kFw * sizeof(double);
    CUDA_OK(c, cudaMemcpy2DAsync(r->waF[1].p + (size_t)firstF * vlc::kFw, pitch, r->waF[0].p + (size_t)firstF * vlc::kFw, pitch,
                                 (size_t)nactF * vlc::kFw * sizeof(double), (size_t)r->nbConvect, cudaMemcpyDeviceToDevice,
                                 c->stream));
  }
  for (int ib = 0; ib < r->nbConvect; ++ib) {
    r->stale_near[1][ib] = std::min(r->stale_near[1][ib], std::max(r->rowNear, r->stale_near[0][ib]));
    r->stale_far[1][ib] = std::min(r->stale_far[1][ib], std::max(r->rowFar, r->stale_far[0][ib]));
  }
  r->dirty[1] = true;
  return VLC_OK;
}

namespace {
// Tmat(ib) = getTransformAxis(twoPi/nb*(ib-1), shaftAxis) of blades 2..nb (classdef.f90:4803-4805, :5204-5205) on the device
int upload_blade_rotations(vlc_ctx* c, Rotor& r) {
  const double twoPi = 2.0 * (std::atan(1.0) * 4.0);
  r.h_axi.resize(r.nb);
  for (int ib = 2; ib <= r.nb; ++ib) {
    const double bladeOffset = twoPi / r.nb * (ib - 1);
    vlc::AxiT& t = r.h_axi[ib - 1];
    t.rotate = std::fabs(bladeOffset) > 2.220446049250313e-16 ? 1 : 0;  // classdef.f90:1297
    transform_axis(bladeOffset, r.shaftAxis, t.T);
  }
  if (!r.d_axi) CUDA_OK(c, cudaMalloc(&r.d_axi, sizeof(vlc::AxiT) * r.nb));
  CUDA_OK(c, cudaMemcpyAsync(r.d_axi, r.h_axi.data(), sizeof(vlc::AxiT) * r.nb, cudaMemcpyHostToDevice, c->stream));
  return VLC_OK;
}
}  // namespace

extern "C" int vlc_rotor_convectwake(vlc_ctx* c, int ir, double dt, int predicted) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_convectwake(m, ir, dt, predicted));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (r->nNwake <= 0) return VLC_OK;
  const int s = predicted ? 1 : 0;
  const long long nfar = std::max(0, r->nFwake - r->rowFar + 1), nact = std::max(0, r->nNwake - r->rowNear + 1);
  {
    const long long n = (long long)r->nbConvect * ((long long)(r->ns + 1) * r->nNwake + nfar);
    vlc::rec_convect_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(s, r->nbConvect, r->ns, r->nNwake, r->nFwake, r->rowNear,
                                                                       r->rowFar, dt, r->velN[0].p, r->velF[0].p, r->waN[s].p,
                                                                       r->waF[s].p);
    c->launches++;
  }
  {
    const long long n = (long long)r->nbConvect * (r->ns * nact + std::max(0LL, nfar - 1));
    if (n > 0) {
      vlc::rec_continuity_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(0, r->nbConvect, r->ns, r->nNwake, r->nFwake,
                                                                            r->rowNear, r->rowFar, r->waN[s].p, r->waF[s].p);
      c->launches++;
      if (!predicted && r->duct == 1) {  // classdef.f90:1645-1656: only in the 'C' branch
        vlc::rec_continuity_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(1, r->nbConvect, r->ns, r->nNwake, r->nFwake,
                                                                              r->rowNear, r->rowFar, r->waN[s].p, r->waF[s].p);
        c->launches++;
      }
    }
  }
  if (r->axisym == 1 && r->nb > 1) {  // classdef.f90:4801-4823
    if ((rc = upload_blade_rotations(c, *r))) return rc;
    const long long n = (long long)(r->nb - 1) * (r->ns * nact + nfar);
    if (n > 0) {
      vlc::rec_axisym_kernel<<<blocks_for(n, 128), 128, 0, c->stream>>>(r->nb, r->ns, r->nNwake, r->nFwake, r->rowNear, r->rowFar,
                                                                        r->d_axi, r->hubCoords[0], r->hubCoords[1],
                                                                        r->hubCoords[2], r->waN[s].p, r->waF[s].p);
      c->launches++;
    }
  }
  CUDA_OK(c, cudaGetLastError());
  r->dirty[s] = true;
  return VLC_OK;
}

extern "C" int vlc_rotor_calc_skew(vlc_ctx* c, int ir) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_calc_skew(m, ir));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const long long n = (long long)r->nb * r->ns * std::max(0, r->nNwake - r->rowNear + 1);
  if (n <= 0) return VLC_OK;
  vlc::rec_skew_kernel<<<blocks_for(n, 128), 128, 0, c->stream>>>(r->nb, r->nbConvect, r->axisym, r->ns, r->nNwake, r->rowNear,
                                                                  r->waN[0].p);
  c->launches++;
  CUDA_OK(c, cudaGetLastError());
  return VLC_OK;  // the skew member is not a source quantity: the packed sets stay valid
}

extern "C" int vlc_rotor_burst_wake(vlc_ctx* c, int ir, double skewLimit, double largeCoreRadius) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_burst_wake(m, ir, skewLimit, largeCoreRadius));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const long long n = (long long)r->nb * std::max(0, r->nFwake - r->rowFar);
  if (n <= 0) return VLC_OK;  // classdef.f90:2323: fewer than two active far filaments
  vlc::rec_burst_kernel<<<blocks_for(n, 128), 128, 0, c->stream>>>(r->nb, r->nFwake, r->rowFar, skewLimit, largeCoreRadius,
                                                                   r->waF[0].p);
  c->launches++;
  CUDA_OK(c, cudaGetLastError());
  r->dirty[0] = true;
  return VLC_OK;
}

extern "C" int vlc_rotor_updatePrescribedWake(vlc_ctx* c, int ir, double deltaPsi, int prescWakeGenNt, int predicted) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_updatePrescribedWake(m, ir, deltaPsi, prescWakeGenNt, predicted));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (std::fabs(r->shaftAxis[0]) > 2.220446049250313e-16 || std::fabs(r->shaftAxis[1]) > 2.220446049250313e-16)
    return fail(c, VLC_ERR_ARG, "Prescribed far wake only implemented for shaft along Z-axis");  // classdef.f90:1010-1012
  if (r->nFwake <= 0 || prescWakeGenNt < 0) return fail(c, VLC_ERR_STATE, "prescribed far wake needs a far wake (nFwake > 0, prescWakeGenNt >= 0)");
  const int rowStart = prescWakeGenNt == 0 ? r->rowFar : r->nFwake - prescWakeGenNt;  // :5180-5184 (nFwakeEnd = nFwake)
  if (rowStart < 1 || rowStart > r->nFwake)
    return fail(c, VLC_ERR_STATE, "prescribed far wake: no far-wake row to fit (rowFar / prescWakeGenNt outside 1..nFwake)");
  const int s = predicted ? 1 : 0;
  const bool copies = r->axisym == 1 && r->nb > 1;
  if (copies && (rc = upload_blade_rotations(c, *r))) return rc;
  vlc::pf::Fit* fits = reinterpret_cast<vlc::pf::Fit*>(r->pfFits.p);
  vlc::pf_fit_kernel<<<blocks_for(r->nbConvect, 32), 32, 0, c->stream>>>(r->nbConvect, r->nFwake, rowStart, r->nFwake - rowStart + 1,
                                                                         deltaPsi, r->hubCoords[2], r->waF[s].p, r->pfHelix[s].p, fits);
  vlc::pf_helix_kernel<<<blocks_for((long long)r->nb * VLC_NPFWAKE, 128), 128, 0, c->stream>>>(
      r->nb, r->nbConvect, r->axisym, fits, copies ? r->d_axi : nullptr, r->hubCoords[0], r->hubCoords[1], r->hubCoords[2],
      r->wapF[s].p, r->pfHelix[s].p);
  c->launches += 2;
  CUDA_OK(c, cudaGetLastError());
  r->have_pf[s] = true;
  r->dirty[s] = true;
  return VLC_OK;
}

extern "C" int vlc_rotor_put_pfwake_helix(vlc_ctx* c, int ir, int ib, int predicted, const double* helix) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_pfwake_helix(m, ir, ib, predicted, helix));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !helix) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  return upload(c, r->pfHelix[predicted ? 1 : 0], (size_t)2 * r->nb, (size_t)2 * ib, helix, 2);
}

extern "C" int vlc_rotor_get_pfwake(vlc_ctx* c, int ir, int ib, int predicted, double* wapF, double* helix) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !wapF) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const int s = predicted ? 1 : 0;
  const size_t per = (size_t)VLC_NPFWAKE * vlc::kFw;
  CUDA_OK(c, cudaMemcpyAsync(wapF, r->wapF[s].p + per * ib, per * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (helix) CUDA_OK(c, cudaMemcpyAsync(helix, r->pfHelix[s].p + 2 * ib, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_rotor_rollup(vlc_ctx* c, int ir) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_rollup(m, ir));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (r->nNwake <= 0) return VLC_OK;
  if (r->rollupStart < 1 || r->rollupEnd > r->ns)
    return fail(c, VLC_ERR_STATE, "rollup: rollupStart / rollupEnd not set (vlc_rotor_set_wake_params)");
  if (r->nFwake > 0 && (r->rowFar < 1 || r->rowFar > r->nFwake + 1))
    return fail(c, VLC_ERR_STATE, "rollup: rowFar outside 1..nFwake+1");
  int rowFarNext = r->rowFar - 1;  // classdef.f90:4527
  if (r->nFwake > 0 && rowFarNext == 0) {  // :4570-4573 (shiftFwake moves every blade, once)
    vlc::rec_shiftFwake_kernel<<<blocks_for(r->nb * vlc::kFw, 64), 64, 0, c->stream>>>(r->nb, r->nFwake, r->waF[0].p);
    c->launches++;
    rowFarNext = 1;
  }
  vlc::rec_rollup_kernel<<<blocks_for(r->nb, 32), 32, 0, c->stream>>>(r->nb, r->ns, r->nNwake, r->nFwake, r->rollupStart,
                                                                      r->rollupEnd, r->sgnPositive, r->suppressFwake, rowFarNext,
                                                                      r->waN[0].p, r->waF[0].p);
  c->launches++;
  // shiftwake (:4603 -> :4481-4498) into the second buffer, then swap
  const size_t total = (size_t)r->nb * r->nNwake * r->ns * vlc::kVr;
  if ((rc = reserve(c, r->waN_alt, total + 1))) return rc;  // by the logical size: capacities differ after a swap
  vlc::rec_shiftwake_kernel<<<blocks_for((long long)total, 256), 256, 0, c->stream>>>((long long)r->nb * r->ns, r->nNwake,
                                                                                     r->waN[0].p, r->waN_alt.p);
  c->launches++;
  CUDA_OK(c, cudaGetLastError());
  std::swap(r->waN[0], r->waN_alt);
  r->dirty[0] = true;
  return VLC_OK;
}

// main.f90:800-838 ('C') / :1057-1081, :889-911 ('P'): for every rotor ir and convected blade, velNwake(:, rowNear:nNwake, :)
// and velFwake(:, rowFar:nFwake) <- sum over source rotors jr = 1..nr of vind_on{N,F}wake_byRotor(rotor(jr), ...) (+- the
// initial wake velocity while iter < initWakeVelNt).  All targets of all rotors go through ONE sweep per source rotor
// (the reference's per-(ir, ib, jr) calls evaluate the same sums target by target); nothing leaves the device.
extern "C" int vlc_wake_sweep_count(vlc_ctx* c, int64_t* M_out) {
  CHECK_CTX(c);
  if (!M_out) return fail(c, VLC_ERR_ARG, "null pointer");
  long long M = 0;
  for (auto& r : c->rotors)
    if (r.defined) M += wake_targets_of(r);
  *M_out = M;
  return VLC_OK;
}

// Targets [first, first + count) of the wake-sweep list against every rotor; velocities into d_vel (3, M) at the same
// positions.  One process per GPU calls this with its own slice and all-gathers d_vel (the targets are independent:
// libCommon.f90:132-139 is a parallel loop over them); count = M on one GPU.
extern "C" int vlc_wake_sweep_slice(vlc_ctx* c, int predicted, int64_t first, int64_t count, double* d_vel) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  const int s = predicted ? 1 : 0;
  long long M = 0, Mmax = 0;  // Mmax: every row active, sized once
  for (auto& r : c->rotors) {
    if (!r.defined) continue;
    M += wake_targets_of(r);
    if (r.nNwake > 0) Mmax += ((long long)r.nNwake * (r.ns + 1) + r.nFwake) * r.nbConvect;
  }
  if (first < 0 || count < 0 || first + count > M) return fail(c, VLC_ERR_ARG, "target slice outside [0, M)");
  if (count == 0) return VLC_OK;
  if (!d_vel) return fail(c, VLC_ERR_ARG, "null pointer");
  if ((rc = reserve(c, c->ws_P, 3 * (size_t)Mmax))) return rc;
  if ((rc = reserve(c, c->ws_V, 3 * (size_t)Mmax))) return rc;
  long long off = 0;
  for (auto& r : c->rotors) {
    const long long m = r.defined ? wake_targets_of(r) : 0;
    if (m <= 0) continue;
    vlc::rec_targets_kernel<<<blocks_for(m, 256), 256, 0, c->stream>>>(r.nbConvect, r.ns, r.nNwake, r.nFwake, r.rowNear, r.rowFar,
                                                                       r.waN[s].p, r.waF[s].p, c->ws_P.p + 3 * off);
    c->launches++;
    off += m;
  }
  CUDA_OK(c, cudaGetLastError());
  const double* P = c->ws_P.p + 3 * first;
  double* acc = d_vel + 3 * first;
  {  // all source rotors as ONE set (build_ws_combined); rotor by rotor below when their strip shapes differ
    bool ok = false;
    if ((rc = build_ws_combined(c, s, &ok))) return rc;
    if (ok) {
      const SourceSet& v = c->ws_comb[s];
      return (v.has_shared && c->shared_nodes) ? sweep_shared(c, v, count, P, acc) : sweep(c, v.rec.p, v.n_pad, count, P, acc);
    }
  }
  bool first_src = true;
  for (auto& src : c->rotors) {  // jr = 1..nr in order (main.f90:817)
    if (!src.defined) continue;
    if ((rc = pack_rotor(c, src, s))) return rc;
    const SourceSet& v = src.comb[s];
    if (v.n_pad <= 0) continue;
    rc = (v.has_shared && c->shared_nodes) ? sweep_shared(c, v, count, P, c->ws_V.p) : sweep(c, v.rec.p, v.n_pad, count, P, c->ws_V.p);
    if (rc) return rc;
    vlc::rec_accumulate_kernel<<<blocks_for(3 * count, 256), 256, 0, c->stream>>>(3 * count, first_src ? 1 : 0, c->ws_V.p, acc);
    c->launches++;
    first_src = false;
  }
  if (first_src) CUDA_OK(c, cudaMemsetAsync(acc, 0, sizeof(double) * 3 * (size_t)count, c->stream));
  CUDA_OK(c, cudaGetLastError());
  return VLC_OK;
}

extern "C" int vlc_wake_sweep_scatter(vlc_ctx* c, int predicted, int addInitWakeVel, const double* d_vel) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  const int s = predicted ? 1 : 0;
  long long off = 0;
  for (auto& r : c->rotors) {
    const long long m = r.defined ? wake_targets_of(r) : 0;
    if (m <= 0) continue;
    if (!d_vel) return fail(c, VLC_ERR_ARG, "null pointer");
    const double w[3] = {r.initWakeVel * r.shaftAxis[0], r.initWakeVel * r.shaftAxis[1], r.initWakeVel * r.shaftAxis[2]};
    vlc::rec_scatter_vel_kernel<<<blocks_for(m, 256), 256, 0, c->stream>>>(
        r.nbConvect, r.ns, r.nNwake, r.nFwake, r.rowNear, r.rowFar, s, addInitWakeVel ? 1 : 0, w[0], w[1], w[2], d_vel + 3 * off,
        r.velN[s ? 2 : 0].p, r.velF[s ? 2 : 0].p);
    c->launches++;
    off += m;
  }
  CUDA_OK(c, cudaGetLastError());
  return VLC_OK;
}

extern "C" int vlc_wake_sweep(vlc_ctx* c, int predicted, int addInitWakeVel) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_wake_sweep(m, predicted, addInitWakeVel));
  int64_t M = 0;
  int rc = vlc_wake_sweep_count(c, &M);
  if (rc || M <= 0) return rc;
  long long Mmax = 0;
  for (auto& r : c->rotors)
    if (r.defined && r.nNwake > 0) Mmax += ((long long)r.nNwake * (r.ns + 1) + r.nFwake) * r.nbConvect;
  if ((rc = bind_device(c))) return rc;
  // world > 1: this member sweeps its contiguous slice of the target list against ALL sources, then the
  // velocity slices are all-gathered (24 bytes per target, once per predictor and once per corrector stage) and every
  // member scatters the complete list into its own copy of the velocity arrays.
  vlc::grp::Shard sh;
  sh.per = M;
  sh.hi = M;
  if (c->world > 1) sh = vlc::grp::shard_range(M, c->world, c->rank);
  if ((rc = reserve(c, c->ws_acc, 3 * ((size_t)Mmax + (size_t)c->world)))) return rc;  // per*world <= M + world - 1
  if ((rc = vlc_wake_sweep_slice(c, predicted, sh.lo, sh.count(), c->ws_acc.p))) return rc;
  if (c->world > 1 &&
      (rc = allgather_slots(c, c->ws_acc.p, 3 * (size_t)sh.per, [](vlc_ctx* o) -> double* { return o->ws_acc.p; })))
    return rc;
  return vlc_wake_sweep_scatter(c, predicted, addInitWakeVel, c->ws_acc.p);
}

extern "C" int vlc_rotor_wakevel_op(vlc_ctx* c, int ir, int op) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_wakevel_op(m, ir, op));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (r->nNwake <= 0) return VLC_OK;
  const size_t nn = (size_t)3 * r->nNwake * (r->ns + 1) * r->nbConvect, nf = (size_t)3 * r->nFwake * r->nbConvect;
  auto copy = [&](DevBuf& dst, const DevBuf& src, size_t n) -> int {
    if (n == 0) return VLC_OK;
    CUDA_OK(c, cudaMemcpyAsync(dst.p, src.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return VLC_OK;
  };
  const vlc::VelArrays va{r->velN[0].p, r->velN[1].p, r->velN[2].p, r->velN[3].p, r->velF[0].p, r->velF[1].p, r->velF[2].p, r->velF[3].p};
  auto fused = [&](int fop, int dst, int src) -> int {  // near and far arrays in one launch (wake_state.cuh)
    const long long tot = (long long)(nn + nf);
    if (tot <= 0) return VLC_OK;
    vlc::wakevel_fused_kernel<<<blocks_for(tot, 256), 256, 0, c->stream>>>(fop, (long long)nn, (long long)nf, va, dst, src);
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
    return VLC_OK;
  };
  (void)copy;
  switch (op) {
    case VLC_VEL_FIRST_STEP:  // main.f90:1013-1020: vel1 = vel
      if ((rc = fused(2, 1, 0))) return rc;
      break;
    case VLC_VEL_AB2:  // main.f90:1031-1041: velStep = vel; vel = 0.5*(3*vel - vel1), whole arrays
      if ((rc = fused(0, 0, 0))) return rc;
      break;
    case VLC_VEL_AM2:  // main.f90:1094-1099: vel = (velPredicted + velStep)*0.5
      if ((rc = fused(1, 0, 0))) return rc;
      break;
    case VLC_VEL_SHIFT_HISTORY:  // main.f90:1103-1107: vel1 = velStep
      if ((rc = fused(2, 1, 3))) return rc;
      break;
    case VLC_VEL_COPY_TO_STEP:  // velStep = vel (fdScheme 2, main.f90:975-988 together with AB2 and FIRST_STEP)
      if ((rc = fused(2, 3, 0))) return rc;
      break;
    case VLC_VEL_ORDER2: {  // main.f90:927-940: vel(active) = vel_order2(vel(active), velPredicted(active))
      const int rowsN = r->nNwake - r->rowNear + 1, rowsF = r->nFwake - r->rowFar + 1;
      if ((rc = reserve(c, r->order2_tmp, std::max(nn, nf) + 1))) return rc;
      if (rowsN > 0) {
        const long long n = 3LL * rowsN * (r->ns + 1) * r->nbConvect;
        vlc::rec_vel_order2_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nbConvect, r->nNwake, r->ns + 1, r->rowNear, rowsN,
                                                                              r->velN[0].p, r->velN[2].p, r->order2_tmp.p);
        vlc::rec_copy_slice_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nbConvect, r->nNwake, r->ns + 1, r->rowNear, rowsN,
                                                                              r->order2_tmp.p, r->velN[0].p);
        c->launches += 2;
      }
      if (rowsF > 0) {
        const long long n = 3LL * rowsF * r->nbConvect;
        vlc::rec_vel_order2_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nbConvect, r->nFwake, 1, r->rowFar, rowsF,
                                                                              r->velF[0].p, r->velF[2].p, r->order2_tmp.p);
        vlc::rec_copy_slice_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nbConvect, r->nFwake, 1, r->rowFar, rowsF,
                                                                              r->order2_tmp.p, r->velF[0].p);
        c->launches += 2;
      }
      CUDA_OK(c, cudaGetLastError());
    } break;
    default: return fail(c, VLC_ERR_ARG, "unknown velocity operation");
  }
  return VLC_OK;
}

namespace {
// velocity array by id: 0 vel, 1 vel1, 2 velPredicted, 3 velStep, 4 vel2, 5 vel3 (VLC_VEL_ARRAY_*)
DevBuf& vel_array(Rotor& r, int id, bool far) {
  if (id < 4) return far ? r.velF[id] : r.velN[id];
  return far ? r.velFx[id - 4] : r.velNx[id - 4];
}
// vel2 / vel3 exist from their first use on, zero like the other arrays after vlc_rotor_define
int ensure_histories(vlc_ctx* c, Rotor& r) {
  int rc;
  for (int k = 0; k < 2; ++k) {
    const size_t nn = (size_t)3 * r.nNwake * (r.ns + 1) * r.nb + 1, nf = (size_t)3 * r.nFwake * r.nb + 1;
    if (r.velNx[k].cap < nn) {
      if ((rc = reserve(c, r.velNx[k], nn))) return rc;
      CUDA_OK(c, cudaMemsetAsync(r.velNx[k].p, 0, r.velNx[k].cap * sizeof(double), c->stream));
    }
    if (r.velFx[k].cap < nf) {
      if ((rc = reserve(c, r.velFx[k], nf))) return rc;
      CUDA_OK(c, cudaMemsetAsync(r.velFx[k].p, 0, r.velFx[k].cap * sizeof(double), c->stream));
    }
  }
  return VLC_OK;
}
}  // namespace

extern "C" int vlc_rotor_wakevel_copy(vlc_ctx* c, int ir, int dst, int src) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_wakevel_copy(m, ir, dst, src));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (dst < 0 || dst > 5 || src < 0 || src > 5) return fail(c, VLC_ERR_ARG, "velocity array id outside 0..5");
  if (r->nNwake <= 0 || dst == src) return VLC_OK;
  if ((dst > 3 || src > 3) && (rc = ensure_histories(c, *r))) return rc;
  const size_t nn = (size_t)3 * r->nNwake * (r->ns + 1) * r->nbConvect, nf = (size_t)3 * r->nFwake * r->nbConvect;
  if (nn) CUDA_OK(c, cudaMemcpyAsync(vel_array(*r, dst, false).p, vel_array(*r, src, false).p, nn * sizeof(double),
                                     cudaMemcpyDeviceToDevice, c->stream));
  if (nf) CUDA_OK(c, cudaMemcpyAsync(vel_array(*r, dst, true).p, vel_array(*r, src, true).p, nf * sizeof(double),
                                     cudaMemcpyDeviceToDevice, c->stream));
  return VLC_OK;
}

extern "C" int vlc_rotor_wakevel_lincomb(vlc_ctx* c, int ir, int dst, int nterms, const int* src, const double* coef,
                                         double divisor) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_wakevel_lincomb(m, ir, dst, nterms, src, coef, divisor));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (dst < 0 || dst > 5 || nterms < 1 || nterms > 4 || !src || !coef || divisor == 0.0)
    return fail(c, VLC_ERR_ARG, "bad linear combination of velocity arrays");
  bool hist = dst > 3;
  for (int k = 0; k < nterms; ++k) {
    if (src[k] < 0 || src[k] > 5) return fail(c, VLC_ERR_ARG, "velocity array id outside 0..5");
    hist = hist || src[k] > 3;
  }
  if (r->nNwake <= 0) return VLC_OK;
  if (hist && (rc = ensure_histories(c, *r))) return rc;
  const long long nn = 3LL * r->nNwake * (r->ns + 1) * r->nbConvect, nf = 3LL * r->nFwake * r->nbConvect;
  for (int far = 0; far < 2; ++far) {
    const long long n = far ? nf : nn;
    if (n <= 0) continue;
    const double* s[4];
    double cf[4];
    for (int k = 0; k < 4; ++k) {
      s[k] = vel_array(*r, src[k < nterms ? k : 0], far != 0).p;
      cf[k] = k < nterms ? coef[k] : 0.0;
    }
    vlc::rec_lincomb_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, nterms, s[0], s[1], s[2], s[3], cf[0], cf[1], cf[2], cf[3],
                                                                        divisor, vel_array(*r, dst, far != 0).p);
    c->launches++;
  }
  CUDA_OK(c, cudaGetLastError());
  return VLC_OK;
}

extern "C" int vlc_rotor_get_nwake(vlc_ctx* c, int ir, int ib, int predicted, double* waN) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !waN) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const size_t per = (size_t)r->nNwake * r->ns * vlc::kVr;
  if (per == 0) return VLC_OK;
  CUDA_OK(c, cudaMemcpyAsync(waN, r->waN[predicted ? 1 : 0].p + per * ib, per * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_rotor_get_fwake(vlc_ctx* c, int ir, int ib, int predicted, double* waF) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb) return fail(c, VLC_ERR_ARG, "bad blade index");
  const size_t per = (size_t)r->nFwake * vlc::kFw;
  if (per == 0) return VLC_OK;
  if (!waF) return fail(c, VLC_ERR_ARG, "null pointer");
  CUDA_OK(c, cudaMemcpyAsync(waF, r->waF[predicted ? 1 : 0].p + per * ib, per * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_rotor_put_wakevel(vlc_ctx* c, int ir, int ib, int which, const double* velN, const double* velF) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_wakevel(m, ir, ib, which, velN, velF));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || which < 0 || which > 5) return fail(c, VLC_ERR_ARG, "bad blade index / array selector");
  if (which > 3 && (rc = ensure_histories(c, *r))) return rc;
  const size_t pn = (size_t)3 * r->nNwake * (r->ns + 1), pf = (size_t)3 * r->nFwake;
  if (velN && pn) CUDA_OK(c, cudaMemcpyAsync(vel_array(*r, which, false).p + pn * ib, velN, pn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (velF && pf) CUDA_OK(c, cudaMemcpyAsync(vel_array(*r, which, true).p + pf * ib, velF, pf * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_rotor_get_wakevel(vlc_ctx* c, int ir, int ib, int which, double* velN, double* velF) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || which < 0 || which > 5) return fail(c, VLC_ERR_ARG, "bad blade index / array selector");
  if (which > 3 && (rc = ensure_histories(c, *r))) return rc;
  const size_t pn = (size_t)3 * r->nNwake * (r->ns + 1), pf = (size_t)3 * r->nFwake;
  if (velN && pn) CUDA_OK(c, cudaMemcpyAsync(velN, vel_array(*r, which, false).p + pn * ib, pn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (velF && pf) CUDA_OK(c, cudaMemcpyAsync(velF, vel_array(*r, which, true).p + pf * ib, pf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}


// ============================================================================ tier 2c: collocation-point stage

namespace {

// one source set swept at the collocation points held in c->cp_P, result in c->cp_V
// Collocation points against one packed set.  In a group / communicator a sweep of at least rhs_share_min pair interactions
// shares its source splits out between the members (sweep_shared: share_sources); below that an all-reduce costs more than
// the sweep (a caradonna-sized sweep is 70 us on one GPU).
int cp_sweep(vlc_ctx* c, const double* rec, long long n_pad, long long m, const SourceSet* shared) {
  if (!shared) return sweep(c, rec, n_pad, m, c->cp_P.p, c->cp_V.p);
  const bool share = c->world > 1 && c->comm && c->rhs_share_min >= 0.0 && (double)m * (double)shared->n_pad >= c->rhs_share_min;
  return sweep_shared(c, *shared, m, c->cp_P.p, c->cp_V.p, share);
}

int cp_accumulate(vlc_ctx* c, Rotor* r, long long m, int field, int sign) {
  vlc::cp_accumulate_kernel<<<blocks_for(3 * m, 256), 256, 0, c->stream>>>(m, field, sign, c->cp_V.p, r->wiP.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  return VLC_OK;
}

int cp_targets(vlc_ctx* c, Rotor* r, long long m) {
  int rc;
  if ((rc = reserve(c, c->cp_P, 3 * (size_t)m)) || (rc = reserve(c, c->cp_V, 3 * (size_t)m))) return rc;
  vlc::cp_targets_kernel<<<blocks_for(3 * m, 256), 256, 0, c->stream>>>(m, r->wiP.p, c->cp_P.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  return VLC_OK;
}

}  // namespace

extern "C" int vlc_rotor_calc_RHS(vlc_ctx* c, int ir, double* velCP_out, double* RHS_out) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_calc_RHS(m, ir, m->rank == 0 ? velCP_out : nullptr, m->rank == 0 ? RHS_out : nullptr));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const int npb = r->nc * r->ns;
  const long long m = (long long)r->nbConvect * npb;
  if ((rc = cp_targets(c, r, m))) return rc;
  for (int jr = 0; jr < (int)c->rotors.size(); ++jr) {  // main.f90:551-560: wake of every rotor, wing of every other
    Rotor& src = c->rotors[jr];
    if (!src.defined) continue;
    if ((rc = pack_rotor(c, src, 0))) return rc;
    {
      const SourceSet v = wake_view(src, 0);
      if ((rc = cp_sweep(c, v.rec.p, v.n_pad, m, (v.has_shared && c->shared_nodes) ? &v : nullptr))) return rc;
      if ((rc = cp_accumulate(c, r, m, vlc::cp::kVelCP, +1))) return rc;
    }
    if (jr != ir) {
      if ((rc = cp_sweep(c, src.comb[0].rec.p, src.wing_pad[0], m, nullptr))) return rc;
      if ((rc = cp_accumulate(c, r, m, vlc::cp::kVelCP, +1))) return rc;
    }
  }
  vlc::cp_rhs_kernel<<<blocks_for(r->N, 128), 128, 0, c->stream>>>(r->N, npb, r->nbConvect, r->axisym, r->wiP.p, r->rhs.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  r->have_rhs = true;
  if (velCP_out)
    CUDA_OK(c, cudaMemcpy2DAsync(velCP_out, 3 * sizeof(double), r->wiP.p + vlc::cp::kVelCP, vlc::kWp * sizeof(double),
                                 3 * sizeof(double), (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  if (RHS_out) CUDA_OK(c, cudaMemcpyAsync(RHS_out, r->rhs.p, sizeof(double) * r->N, cudaMemcpyDeviceToHost, c->stream));
  if (velCP_out || RHS_out) CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

// main.f90:528-547 at the head of EVERY sub-iteration of ntSubLoop (:522): velCP starts again from its kinematic part,
// which the driver keeps in velCPm (:545-546).  vlc_rotor_calc_RHS adds the induced velocities to velCP, so a second
// pass over the same wing records (ntSub > 0) calls this first.
extern "C" int vlc_rotor_reset_velCP(vlc_ctx* c, int ir) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_reset_velCP(m, ir));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const long long m = (long long)r->nbConvect * r->nc * r->ns;
  vlc::cp_copy_field_kernel<<<blocks_for(3 * m, 256), 256, 0, c->stream>>>(m, vlc::cp::kVelCPm, vlc::cp::kVelCP, r->wiP.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  return VLC_OK;
}

extern "C" int vlc_rotor_solve_map_gam(vlc_ctx* c, int ir, double* gamVec_out) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_solve_map_gam(m, ir, m->rank == 0 ? gamVec_out : nullptr));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (!r->factored) return fail(c, VLC_ERR_STATE, "vlc_rotor_solve_map_gam before vlc_rotor_calcAIC");
  if (!r->have_rhs) return fail(c, VLC_ERR_STATE, "vlc_rotor_solve_map_gam before vlc_rotor_calc_RHS");
  if ((rc = solve_dev(c, r, r->rhs.p, r->gamvec.p))) return rc;
  vlc::cp_map_gam_kernel<<<blocks_for(r->N, 128), 128, 0, c->stream>>>(r->nb, r->nc * r->ns, r->nbConvect, r->axisym, r->gamvec.p, r->wiP.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  r->have_rhs = false;  // one solve per right-hand side
  mark_wing_dirty(*r);  // the wing's circulation changed
  if (gamVec_out) {
    CUDA_OK(c, cudaMemcpyAsync(gamVec_out, r->gamvec.p, sizeof(double) * r->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  return VLC_OK;
}

extern "C" int vlc_rotor_put_sections(vlc_ctx* c, int ir, int ib, const double* sec) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_put_sections(m, ir, ib, sec));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !sec) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const size_t per = (size_t)vlc::cp::sec_doubles(r->ns);
  r->have_sec[ib] = 1;
  return upload(c, r->sec, per * r->nb, per * ib, sec, per);
}

extern "C" int vlc_rotor_calc_velCPTotal(vlc_ctx* c, int ir) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_calc_velCPTotal(m, ir));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  const int npb = r->nc * r->ns;
  const long long m = (long long)r->nbConvect * npb;
  if ((rc = cp_targets(c, r, m))) return rc;
  vlc::cp_copy_field_kernel<<<blocks_for(3 * m, 256), 256, 0, c->stream>>>(m, vlc::cp::kVelCP, vlc::cp::kVelCPTotal, r->wiP.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  for (auto& src : c->rotors) {  // main.f90:639-645: minus the bound vortices of every rotor
    if (!src.defined) continue;
    if ((rc = pack_bound(c, src))) return rc;
    if ((rc = cp_sweep(c, src.bound.rec.p, src.bound.n_pad, m, nullptr))) return rc;
    if ((rc = cp_accumulate(c, r, m, vlc::cp::kVelCPTotal, -1))) return rc;
  }
  if ((rc = pack_rotor(c, *r, 0))) return rc;  // main.f90:647-652: plus the whole wing of this rotor
  if ((rc = cp_sweep(c, r->comb[0].rec.p, r->wing_pad[0], m, nullptr))) return rc;
  if ((rc = cp_accumulate(c, r, m, vlc::cp::kVelCPTotal, +1))) return rc;
  if (r->axisym == 1 && r->nb > 1) {  // main.f90:658-663
    const long long n = (long long)(r->nb - 1) * npb * 3;
    vlc::cp_axisym_field_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nb, npb, vlc::cp::kVelCPTotal, 3, r->wiP.p);
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
  }
  return VLC_OK;
}

extern "C" int vlc_rotor_calc_force(vlc_ctx* c, int ir, double density, double dt, double Omega, int spanwiseLiftSwitch) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_rotor_calc_force(m, ir, density, dt, Omega, spanwiseLiftSwitch));
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (!(dt > 0.0)) return fail(c, VLC_ERR_ARG, "dt must be positive");
  for (int ib = 0; ib < r->nbConvect; ++ib)
    if (!r->have_sec[ib]) return fail(c, VLC_ERR_STATE, "vlc_rotor_calc_force before vlc_rotor_put_sections of every convected blade");
  const int npb = r->nc * r->ns, nld = vlc::cp::loads_doubles(r->ns);
  if ((rc = reserve(c, r->loads_scr, (size_t)r->nbConvect * npb * vlc::cp::kScr))) return rc;
  const int threads = npb >= 256 ? 256 : (npb + 31) / 32 * 32;  // a thread per panel of the blade (strided beyond 256)
  vlc::cp_loads_kernel<<<r->nbConvect, threads, 0, c->stream>>>(r->nc, r->ns, density, dt, Omega, spanwiseLiftSwitch, r->wiP.p,
                                                                r->sec.p, r->loads.p, r->loads_scr.p);
  CUDA_OK(c, cudaGetLastError());
  c->launches++;
  if (r->axisym == 1 && r->nb > 1) {  // classdef.f90:4623-4650: blades 2..nb take blade 1's pressures, forces and loads
    const int fields[3][2] = {{vlc::cp::kDelP, 2}, {vlc::cp::kGamPrev, 1}, {vlc::cp::kNormalForce, 6}};
    for (auto& f : fields) {
      const long long n = (long long)(r->nb - 1) * npb * f[1];
      vlc::cp_axisym_field_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nb, npb, f[0], f[1], r->wiP.p);
      c->launches++;
    }
    const long long n = (long long)(r->nb - 1) * nld;
    vlc::cp_axisym_loads_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(r->nb, r->ns, nld, r->loads.p);
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
  }
  return VLC_OK;
}

extern "C" int vlc_rotor_get_loads(vlc_ctx* c, int ir, int ib, double* loads) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !loads) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const size_t per = (size_t)vlc::cp::loads_doubles(r->ns);
  CUDA_OK(c, cudaMemcpyAsync(loads, r->loads.p + per * ib, per * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_rotor_get_wing(vlc_ctx* c, int ir, int ib, double* wiP) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (ib < 0 || ib >= r->nb || !wiP) return fail(c, VLC_ERR_ARG, "bad blade index / null pointer");
  const size_t per = (size_t)r->nc * r->ns * vlc::kWp;
  CUDA_OK(c, cudaMemcpyAsync(wiP, r->wiP.p + per * ib, per * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}

extern "C" int vlc_convect_dev(vlc_ctx* c, int64_t n, double* x, const double* v, double dt) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!x || !v))) return fail(c, VLC_ERR_ARG, "bad arguments");
  LAUNCH1D(c, vlc::convect_kernel, 3 * n, 3 * n, x, v, dt);
  return VLC_OK;
}

extern "C" int vlc_ab2_dev(vlc_ctx* c, int64_t n, const double* v, const double* v1, double* out) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!v || !v1 || !out))) return fail(c, VLC_ERR_ARG, "bad arguments");
  LAUNCH1D(c, vlc::ab2_kernel, 3 * n, 3 * n, v, v1, out);
  return VLC_OK;
}

extern "C" int vlc_am2_dev(vlc_ctx* c, int64_t n, const double* vp, const double* vs, double* out) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!vp || !vs || !out))) return fail(c, VLC_ERR_ARG, "bad arguments");
  LAUNCH1D(c, vlc::am2_kernel, 3 * n, 3 * n, vp, vs, out);
  return VLC_OK;
}

extern "C" int vlc_vel_order2_dev(vlc_ctx* c, int rows, int cols, const double* vn, const double* vnp1, double* out) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (rows < 0 || cols < 0 || (rows > 0 && cols > 0 && (!vn || !vnp1 || !out)) || out == vn || out == vnp1)
    return fail(c, VLC_ERR_ARG, "bad arguments (out must not alias the inputs)");
  const long long n = 3LL * rows * cols;
  LAUNCH1D(c, vlc::vel_order2_kernel, n, rows, cols, vn, vnp1, out);
  return VLC_OK;
}

extern "C" int vlc_dissipate_dev(vlc_ctx* c, int64_t n_rvc, double* rvc, int64_t n_gam, double* gam,
                                 double apparentViscCoeff, double nu, double decayCoeff, double dt) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (n_rvc < 0 || n_gam < 0) return fail(c, VLC_ERR_ARG, "bad arguments");
  if (rvc) LAUNCH1D(c, vlc::core_growth_kernel, n_rvc, n_rvc, rvc, apparentViscCoeff, nu, dt);
  if (gam) LAUNCH1D(c, vlc::decay_kernel, n_gam, n_gam, gam, decayCoeff, dt);
  return VLC_OK;
}

extern "C" int vlc_dissipate_lattice_dev(vlc_ctx* c, int nrows, int ns, double* rvc4, double* gam,
                                         double apparentViscCoeff, double nu, double decayCoeff, double dt) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (nrows < 0 || ns < 0 || !rvc4 || !gam) return fail(c, VLC_ERR_ARG, "bad arguments");
  const long long n = (long long)nrows * ns;
  LAUNCH1D(c, vlc::dissipate_lattice_kernel, n, nrows, ns, rvc4, gam, apparentViscCoeff, nu, decayCoeff, dt);
  LAUNCH1D(c, vlc::dissipate_lattice_vf4_kernel, n, nrows, ns, rvc4);
  return VLC_OK;
}

extern "C" int vlc_strain_dev(vlc_ctx* c, int64_t n, const double* p1, const double* p2, const double* l0,
                              const double* rvc0, double* rvc) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!p1 || !p2 || !l0 || !rvc0 || !rvc))) return fail(c, VLC_ERR_ARG, "bad arguments");
  LAUNCH1D(c, vlc::strain_kernel, n, n, p1, p2, l0, rvc0, rvc);
  return VLC_OK;
}

extern "C" int vlc_pack_lattice_dev(vlc_ctx* c, int set, int append, int nrows, int ns, const double* nodes,
                                    const double* gam, const double* rvc4, int nfar, const double* far_nodes,
                                    const double* gamF, const double* rvcF) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = check_set(c, set))) return rc;
  if (nrows < 0 || ns < 1 || nfar < 0) return fail(c, VLC_ERR_ARG, "bad lattice shape");
  if (nrows > 0 && (!nodes || !gam || !rvc4)) return fail(c, VLC_ERR_ARG, "null lattice array");
  if (nfar > 0 && (!far_nodes || !gamF || !rvcF || nrows == 0)) return fail(c, VLC_ERR_ARG, "bad far-wake arguments");
  SourceSet& s = c->sets[set];
  const long long base = append ? s.n : 0;
  const long long add = 4LL * nrows * ns + (nfar > 0 ? ns + nfar : 0);
  const long long n_new = base + add;
  const long long n_pad = pad_tile(n_new);
  if ((size_t)n_pad * vlc::kSrcDoubles > s.rec.cap) {
    // grow, preserving what is already packed
    DevBuf nb;
    if ((rc = reserve(c, nb, (size_t)n_pad * vlc::kSrcDoubles * 2))) return rc;
    if (base > 0)
      CUDA_OK(c, cudaMemcpyAsync(nb.p, s.rec.p, sizeof(double) * (size_t)base * vlc::kSrcDoubles,
                                 cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    release(s.rec);
    s.rec = nb;
  }
  double* rec = s.rec.p + (size_t)base * vlc::kSrcDoubles;
  long long off = 0;
  if (nrows > 0) {
    const long long cnt = 4LL * nrows * ns;
    LAUNCH1D(c, vlc::pack_lattice_kernel, cnt, nrows, ns, nodes, gam, rvc4, rec);
    off += cnt;
  }
  if (nfar > 0) {
    LAUNCH1D(c, vlc::pack_horseshoe_kernel, (long long)ns, nrows, ns, nodes, gam, rvc4,
             rec + (size_t)off * vlc::kSrcDoubles);
    off += ns;
    LAUNCH1D(c, vlc::pack_chain_kernel, (long long)nfar, nfar, far_nodes, gamF, rvcF,
             rec + (size_t)off * vlc::kSrcDoubles);
    off += nfar;
  }
  if (n_pad > n_new)
    LAUNCH1D(c, vlc::pack_null_kernel, n_pad - n_new, n_pad - n_new, s.rec.p + (size_t)n_new * vlc::kSrcDoubles);
  s.n = n_new;
  s.n_pad = n_pad;

  // ---- shared-node form of the same lattice: strip records + flat remainder (bs_lattice.cuh) ----
  if (!append) {
    s.n_rings_main = 0;
    s.may_dual = false;
    s.n_lat = s.n_rem = s.n_lat2 = s.n_lat2_pad = 0;
    s.has_shared = true;
    s.lat_W = auto_strip_width(c, ns);  // lattices appended later share the record width of the first one
    if (!s.d_unmergeable) CUDA_OK(c, cudaMalloc(&s.d_unmergeable, sizeof(int)));
    CUDA_OK(c, cudaMemsetAsync(s.d_unmergeable, 0, sizeof(int), c->stream));
  }
  if (!s.has_shared) return VLC_OK;  // appended to a set that was not started by a lattice: flat form only
  auto grow_keep = [&](DevBuf& b, size_t keep, size_t need) -> int {
    if (need <= b.cap) return VLC_OK;
    DevBuf nb;
    int r2 = reserve(c, nb, need * 2);
    if (r2) return r2;
    if (keep > 0) CUDA_OK(c, cudaMemcpyAsync(nb.p, b.p, sizeof(double) * keep, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    release(b);
    b = nb;
    return VLC_OK;
  };
  const int LW = s.lat_W, RD = lat_rd_of(LW);
  const long long lat_add = nrows > 0 ? (long long)((ns + LW - 1) / LW) * (nrows + 1) : 0;
  const long long lat_new = s.n_lat + lat_add;
  const long long lat_pad = pad_lat(lat_new, LW);
  const long long rem_add = (nrows > 0 ? nrows : 0) + (nfar > 0 ? ns + nfar : 0);
  const long long rem_new = s.n_rem + rem_add;
  const long long rem_pad = pad_tile(rem_new);
  if ((rc = grow_keep(s.lat, (size_t)s.n_lat * RD, (size_t)lat_pad * RD))) return rc;
  if ((rc = grow_keep(s.rem, (size_t)s.n_rem * vlc::kSrcDoubles, (size_t)rem_pad * vlc::kSrcDoubles))) return rc;
  if (nrows > 0) {
#define X(WW)                                                                                              \
  if (LW == WW)                                                                                            \
    vlc::pack_lattice_shared_kernel<WW><<<blocks_for(lat_add, 128), 128, 0, c->stream>>>(nrows, ns, nodes, gam, rvc4, \
                                                                                          s.lat.p + (size_t)s.n_lat * RD, s.d_unmergeable);
    X(1) X(2) X(3) X(4)
#undef X
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
    double* rrec = s.rem.p + (size_t)s.n_rem * vlc::kSrcDoubles;
    LAUNCH1D(c, vlc::pack_lastcol_kernel, (long long)nrows, nrows, ns, nodes, gam, rvc4, rrec);
    long long roff = nrows;
    if (nfar > 0) {
      LAUNCH1D(c, vlc::pack_horseshoe_kernel, (long long)ns, nrows, ns, nodes, gam, rvc4,
               rrec + (size_t)roff * vlc::kSrcDoubles);
      roff += ns;
      LAUNCH1D(c, vlc::pack_chain_kernel, (long long)nfar, nfar, far_nodes, gamF, rvcF,
               rrec + (size_t)roff * vlc::kSrcDoubles);
    }
  }
  if (lat_pad > lat_new) {
#define X(WW)                                                                                              \
  if (LW == WW)                                                                                            \
    vlc::pack_null_lat_kernel<WW><<<blocks_for(lat_pad - lat_new, 128), 128, 0, c->stream>>>(lat_pad - lat_new, \
                                                                                             s.lat.p + (size_t)lat_new * RD);
    X(1) X(2) X(3) X(4)
#undef X
    CUDA_OK(c, cudaGetLastError());
    c->launches++;
  }
  if (rem_pad > rem_new)
    LAUNCH1D(c, vlc::pack_null_kernel, rem_pad - rem_new, rem_pad - rem_new, s.rem.p + (size_t)rem_new * vlc::kSrcDoubles);
  s.n_lat = lat_new;
  s.n_lat_pad = lat_pad;
  s.n_rings_main += (long long)nrows * ns;
  s.n_rem = rem_new;
  s.n_rem_pad = rem_pad;
  return VLC_OK;
}

extern "C" int vlc_pack_lattice(vlc_ctx* c, int set, int append, int nrows, int ns, const double* nodes,
                                const double* gam, const double* rvc4, int nfar, const double* far_nodes,
                                const double* gamF, const double* rvcF) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_pack_lattice(m, set, append, nrows, ns, nodes, gam, rvc4, nfar, far_nodes, gamF, rvcF));
  int rc = bind_device(c);
  if (rc) return rc;
  if (nrows < 0 || ns < 1 || nfar < 0) return fail(c, VLC_ERR_ARG, "bad lattice shape");
  if (nrows > 0 && (!nodes || !gam || !rvc4)) return fail(c, VLC_ERR_ARG, "null lattice array");
  if (nfar > 0 && (!far_nodes || !gamF || !rvcF)) return fail(c, VLC_ERR_ARG, "null far-wake array");
  const size_t n_nodes = 3 * (size_t)(nrows + 1) * (ns + 1), n_g = (size_t)nrows * ns, n_r = 4 * n_g;
  const size_t n_fn = nfar > 0 ? 3 * (size_t)(nfar + 1) : 0, n_f = (size_t)nfar;
  if ((rc = reserve(c, c->scratch, n_nodes + n_g + n_r + n_fn + 2 * n_f + 8))) return rc;
  double* d = c->scratch.p;
  double *d_nodes = d, *d_gam = d_nodes + n_nodes, *d_rvc = d_gam + n_g, *d_fn = d_rvc + n_r, *d_gF = d_fn + n_fn,
         *d_rF = d_gF + n_f;
  cudaStream_t st = c->stream;
  if (nrows > 0) {
    CUDA_OK(c, cudaMemcpyAsync(d_nodes, nodes, sizeof(double) * n_nodes, cudaMemcpyHostToDevice, st));
    CUDA_OK(c, cudaMemcpyAsync(d_gam, gam, sizeof(double) * n_g, cudaMemcpyHostToDevice, st));
    CUDA_OK(c, cudaMemcpyAsync(d_rvc, rvc4, sizeof(double) * n_r, cudaMemcpyHostToDevice, st));
  }
  if (nfar > 0) {
    CUDA_OK(c, cudaMemcpyAsync(d_fn, far_nodes, sizeof(double) * n_fn, cudaMemcpyHostToDevice, st));
    CUDA_OK(c, cudaMemcpyAsync(d_gF, gamF, sizeof(double) * n_f, cudaMemcpyHostToDevice, st));
    CUDA_OK(c, cudaMemcpyAsync(d_rF, rvcF, sizeof(double) * n_f, cudaMemcpyHostToDevice, st));
  }
  rc = vlc_pack_lattice_dev(c, set, append, nrows, ns, d_nodes, d_gam, d_rvc, nfar, nfar > 0 ? d_fn : nullptr,
                            nfar > 0 ? d_gF : nullptr, nfar > 0 ? d_rF : nullptr);
  if (rc) return rc;
  CUDA_OK(c, cudaStreamSynchronize(st));  // the scratch buffer is reused by the next call
  return VLC_OK;
}

extern "C" int vlc_set_info(vlc_ctx* c, int set, int64_t* out) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if ((rc = check_set(c, set))) return rc;
  if (!out) return fail(c, VLC_ERR_ARG, "null pointer");
  const SourceSet& s = c->sets[set];
  out[0] = s.n;
  out[1] = s.has_shared ? s.n_lat + s.n_lat2 : 0;
  out[2] = s.has_shared ? s.n_rem : 0;
  out[3] = -1;
  out[4] = s.has_shared ? s.lat_W : 0;
  out[5] = s.has_shared ? s.lat2_W : 0;
  if (s.has_shared && s.d_unmergeable) {
    int f = 0;
    CUDA_OK(c, cudaMemcpyAsync(&f, s.d_unmergeable, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    out[3] = !c->shared_nodes ? 0 : (f == 0 ? 1 : (f == 2 && s.may_dual ? 2 : 0));
  }
  return VLC_OK;
}

extern "C" int vlc_rotor_info(vlc_ctx* c, int ir, int predicted, int64_t* out) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  Rotor* r = get_rotor(c, ir);
  if (!r) return VLC_ERR_STATE;
  if (!out) return fail(c, VLC_ERR_ARG, "null pointer");
  const int s = predicted ? 1 : 0;
  if ((rc = pack_rotor(c, *r, s))) return rc;
  const SourceSet& cs = r->comb[s];
  out[0] = cs.n;
  out[1] = cs.has_shared ? cs.n_lat + cs.n_lat2 : 0;
  out[2] = cs.has_shared ? cs.n_rem : 0;
  out[3] = -1;
  out[4] = cs.has_shared ? cs.lat_W : 0;
  out[5] = cs.has_shared ? cs.lat2_W : 0;
  if (cs.has_shared && cs.d_unmergeable) {
    int f = 0;
    CUDA_OK(c, cudaMemcpyAsync(&f, cs.d_unmergeable, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    out[3] = !c->shared_nodes ? 0 : (f == 0 ? 1 : (f == 2 ? 2 : 0));
  }
  return VLC_OK;
}

extern "C" int vlc_last_sweep_ms(vlc_ctx* c, double* ms_kernel, double* ms_total) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (!c->ev_valid) return fail(c, VLC_ERR_STATE, "no sweep has been launched yet");
  CUDA_OK(c, cudaEventSynchronize(c->ev[2]));
  float a = 0.f, b = 0.f;
  CUDA_OK(c, cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
  CUDA_OK(c, cudaEventElapsedTime(&b, c->ev[0], c->ev[2]));
  if (ms_kernel) *ms_kernel = a;
  if (ms_total) *ms_total = b;
  return VLC_OK;
}

extern "C" int vlc_set_lattice_tuning(vlc_ctx* c, int strip_width, int targets_per_thread) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_set_lattice_tuning(m, strip_width, targets_per_thread));
  const int W = strip_width, T = targets_per_thread;
  if (W < 0 || W > 5 || T < 0 || T > 3)
    return fail(c, VLC_ERR_ARG, "strip_width in 1..4 (0 = automatic, 5 = automatic with tail strips for small wakes too), targets_per_thread in 1..3");
  c->lat_W = W;
  c->lat_T = T;
  for (auto& r : c->rotors) r.dirty[0] = r.dirty[1] = true;  // tier-3 sets are re-packed by their owner
  return VLC_OK;
}

extern "C" int vlc_set_shared_nodes(vlc_ctx* c, int on) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_set_shared_nodes(m, on));
  c->shared_nodes = (on != 0);
  for (auto& r : c->rotors) r.dirty[0] = r.dirty[1] = true;
  return VLC_OK;
}

extern "C" int vlc_lattice_targets_dev(vlc_ctx* c, int nrows, int ns, const double* nodes, double* P) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (nrows < 0 || ns < 0 || !nodes || !P) return fail(c, VLC_ERR_ARG, "bad arguments");
  const long long n = 3LL * nrows * (ns + 1);
  LAUNCH1D(c, vlc::lattice_gather_kernel, n, nrows, ns, nodes, P);
  return VLC_OK;
}

extern "C" int vlc_lattice_scatter_dev(vlc_ctx* c, int nrows, int ns, double* nodes, const double* P) {
  CHECK_CTX(c);
  VLC_NO_GROUP(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (nrows < 0 || ns < 0 || !nodes || !P) return fail(c, VLC_ERR_ARG, "bad arguments");
  const long long n = 3LL * nrows * (ns + 1);
  LAUNCH1D(c, vlc::lattice_scatter_kernel, n, nrows, ns, nodes, P);
  return VLC_OK;
}

// ============================================================================ gridgen

// Cells [first, first + count) of the grid (x fastest, then y, then z): one process per GPU takes a contiguous slice
// of the cell list, like the wake sweeps (targets are independent, gridgen.f90:116-139 is a loop over cells).
extern "C" int vlc_gridgen_slice(vlc_ctx* c, int nx, int ny, int nz, const double* xyzMin, const double* xyzMax,
                                 const double* vel, int64_t nVrWing, const double* vrWing, int64_t nVrNwake,
                                 const double* vrNwake, int64_t nVfNwakeTE, const double* vfNwakeTE, const double* gamNwakeTE,
                                 int64_t nVfFwake, const double* vfFwake, const double* gamFwake, int64_t first, int64_t count,
                                 double* gridCentre, double* velCentre) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (nx < 2 || ny < 2 || nz < 2 || !xyzMin || !xyzMax || !vel || (count > 0 && !velCentre))
    return fail(c, VLC_ERR_ARG, "bad grid arguments");
  if (xyzMin[0] > xyzMax[0] || xyzMin[1] > xyzMax[1] || xyzMin[2] > xyzMax[2])
    return fail(c, VLC_ERR_ARG, "ERROR: All XYZmin values should be greater than XYZmax values");  // gridgen.f90:43-45
  if (nVrWing < 0 || nVrNwake < 0 || nVfNwakeTE < 0 || nVfFwake < 0 || (nVrWing > 0 && !vrWing) ||
      (nVrNwake > 0 && !vrNwake) || (nVfNwakeTE > 0 && (!vfNwakeTE || !gamNwakeTE)) ||
      (nVfFwake > 0 && (!vfFwake || !gamFwake)))
    return fail(c, VLC_ERR_ARG, "bad filament arguments");
  cudaStream_t st = c->stream;
  // sources in the file's order; the wing loop of gridgen.f90:121-123 assigns, so only its last ring counts (C12)
  const long long nw = nVrWing > 0 ? 1 : 0;
  const long long n = 4 * nw + 4 * (long long)nVrNwake + nVfNwakeTE + nVfFwake;
  const long long n_pad = pad_tile(n);
  const size_t up = (size_t)(nw + nVrNwake) * vlc::kVr + (size_t)(nVfNwakeTE + nVfFwake) * (vlc::kVf + 1);
  if ((rc = reserve(c, c->scratch, up + 8))) return rc;
  SourceSet& s = c->sets[VLC_MAX_SETS - 1];  // the last set is the grid tool's scratch set
  if ((rc = reserve(c, s.rec, (size_t)n_pad * vlc::kSrcDoubles + 8))) return rc;
  double* d = c->scratch.p;
  double* rec = s.rec.p;
  long long off = 0;
  auto h2d = [&](double* dst, const double* src, size_t cnt) -> int {
    CUDA_OK(c, cudaMemcpyAsync(dst, src, sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
    return VLC_OK;
  };
  if (nw) {
    if ((rc = h2d(d, vrWing + (size_t)(nVrWing - 1) * vlc::kVr, vlc::kVr))) return rc;
    vlc::pack_rings_kernel<<<1, 32, 0, st>>>(d, vlc::kVr, 1, 0, 1, 1, 0xF, 4, 1.0, 0, rec);
    c->launches++;
    d += vlc::kVr;
    off += 4;
  }
  if (nVrNwake > 0) {
    if ((rc = h2d(d, vrNwake, (size_t)nVrNwake * vlc::kVr))) return rc;
    if (nVrNwake > 0x7fffffff / 8) return fail(c, VLC_ERR_ARG, "too many near-wake rings for one file");
    vlc::pack_rings_kernel<<<blocks_for(4 * nVrNwake, 256), 256, 0, st>>>(d, vlc::kVr, (int)nVrNwake, 0, (int)nVrNwake, 1,
                                                                          0xF, 4, 1.0, 0, rec + (size_t)off * vlc::kSrcDoubles);
    c->launches++;
    d += (size_t)nVrNwake * vlc::kVr;
    off += 4 * nVrNwake;
  }
  const double* vfs[2] = {vfNwakeTE, vfFwake};
  const double* gms[2] = {gamNwakeTE, gamFwake};
  const long long cnt[2] = {nVfNwakeTE, nVfFwake};
  for (int k = 0; k < 2; ++k)
    if (cnt[k] > 0) {
      if ((rc = h2d(d, vfs[k], (size_t)cnt[k] * vlc::kVf))) return rc;
      if ((rc = h2d(d + (size_t)cnt[k] * vlc::kVf, gms[k], (size_t)cnt[k]))) return rc;
      vlc::pack_vf_gam_kernel<<<blocks_for(cnt[k], 256), 256, 0, st>>>(cnt[k], d, d + (size_t)cnt[k] * vlc::kVf,
                                                                       rec + (size_t)off * vlc::kSrcDoubles);
      c->launches++;
      d += (size_t)cnt[k] * (vlc::kVf + 1);
      off += cnt[k];
    }
  if (n_pad > n) {
    vlc::pack_null_kernel<<<blocks_for(n_pad - n, 256), 256, 0, st>>>(n_pad - n, rec + (size_t)n * vlc::kSrcDoubles);
    c->launches++;
  }
  CUDA_OK(c, cudaGetLastError());
  s.n = n;
  s.n_pad = n_pad;
  s.has_shared = false;
  s.n_lat = s.n_lat_pad = s.n_rem = s.n_rem_pad = s.n_lat2 = s.n_lat2_pad = 0;
  // targets = cell centres, computed on the device with the file's arithmetic
  const long long m = (long long)(nx - 1) * (ny - 1) * (nz - 1);
  if (first < 0 || count < 0 || first + count > m) return fail(c, VLC_ERR_ARG, "cell slice outside [0, (nx-1)(ny-1)(nz-1))");
  if (count == 0) return VLC_OK;
  if ((rc = reserve(c, c->stage_P, 3 * (size_t)m))) return rc;
  if ((rc = reserve(c, c->stage_V, 3 * (size_t)m))) return rc;
  const double dx = (xyzMax[0] - xyzMin[0]) / (nx - 1), dy = (xyzMax[1] - xyzMin[1]) / (ny - 1),
               dz = (xyzMax[2] - xyzMin[2]) / (nz - 1);
  vlc::grid_centres_kernel<<<blocks_for(m, 256), 256, 0, st>>>(nx, ny, nz, xyzMin[0], xyzMin[1], xyzMin[2], dx, dy, dz,
                                                                c->stage_P.p);
  c->launches++;
  const double* dP = c->stage_P.p + 3 * (size_t)first;
  double* dV = c->stage_V.p + 3 * (size_t)first;
  if ((rc = sweep(c, s.rec.p, s.n_pad, count, dP, dV))) return rc;
  vlc::add_freestream_kernel<<<blocks_for(count, 256), 256, 0, st>>>(count, vel[0], vel[1], vel[2], dV);
  c->launches++;
  CUDA_OK(c, cudaGetLastError());
  if (gridCentre) CUDA_OK(c, cudaMemcpyAsync(gridCentre, dP, sizeof(double) * 3 * (size_t)count, cudaMemcpyDeviceToHost, st));
  CUDA_OK(c, cudaMemcpyAsync(velCentre, dV, sizeof(double) * 3 * (size_t)count, cudaMemcpyDeviceToHost, st));
  CUDA_OK(c, cudaStreamSynchronize(st));
  return VLC_OK;
}

extern "C" int vlc_gridgen(vlc_ctx* c, int nx, int ny, int nz, const double* xyzMin, const double* xyzMax,
                           const double* vel, int64_t nVrWing, const double* vrWing, int64_t nVrNwake,
                           const double* vrNwake, int64_t nVfNwakeTE, const double* vfNwakeTE, const double* gamNwakeTE,
                           int64_t nVfFwake, const double* vfFwake, const double* gamFwake, double* gridCentre,
                           double* velCentre) {
  CHECK_CTX(c);
  if (nx < 2 || ny < 2 || nz < 2) return fail(c, VLC_ERR_ARG, "bad grid arguments");
  if (c->group && c->is_leader && !t_in_member) {  // every member takes a contiguous slice of the cell list (gridgen.f90:116-139)
    const long long cells = (long long)(nx - 1) * (ny - 1) * (nz - 1);
    return group_all(c, [&](vlc_ctx* g_) -> int {
      const vlc::grp::Shard sh = vlc::grp::shard_range(cells, g_->world, g_->rank);
      return vlc_gridgen_slice(g_, nx, ny, nz, xyzMin, xyzMax, vel, nVrWing, vrWing, nVrNwake, vrNwake, nVfNwakeTE, vfNwakeTE,
                               gamNwakeTE, nVfFwake, vfFwake, gamFwake, sh.lo, sh.count(),
                               gridCentre ? gridCentre + 3 * sh.lo : nullptr, velCentre ? velCentre + 3 * sh.lo : nullptr);
    });
  }
  return vlc_gridgen_slice(c, nx, ny, nz, xyzMin, xyzMax, vel, nVrWing, vrWing, nVrNwake, vrNwake, nVfNwakeTE, vfNwakeTE,
                           gamNwakeTE, nVfFwake, vfFwake, gamFwake, 0, (int64_t)(nx - 1) * (ny - 1) * (nz - 1), gridCentre,
                           velCentre);
}

// ============================================================================ measurement

// Device-side timing of whatever the caller brackets: the library runs on its own stream(s) -- several devices for a
// group handle -- which no event of the caller's can see.
extern "C" int vlc_event_record(vlc_ctx* c, int slot) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_event_record(m, slot));
  if (slot < 0 || slot >= 8) return fail(c, VLC_ERR_ARG, "event slot outside 0..7");
  int rc = bind_device(c);
  if (rc) return rc;
  if (!c->user_ev[slot]) CUDA_OK(c, cudaEventCreate(&c->user_ev[slot]));
  CUDA_OK(c, cudaEventRecord(c->user_ev[slot], c->stream));
  return VLC_OK;
}

extern "C" int vlc_event_elapsed_ms(vlc_ctx* c, int slot_a, int slot_b, double* ms) {
  CHECK_CTX(c);
  if (!ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8) return fail(c, VLC_ERR_ARG, "bad event slots / null pointer");
  if (c->group && c->is_leader && !t_in_member) {  // the slowest member
    std::vector<double> each(c->group->members.size(), 0.0);
    const int rc = group_all(c, [&](vlc_ctx* g_) -> int { return vlc_event_elapsed_ms(g_, slot_a, slot_b, &each[g_->rank]); });
    if (rc) return rc;
    *ms = *std::max_element(each.begin(), each.end());
    return VLC_OK;
  }
  int rc = bind_device(c);
  if (rc) return rc;
  if (!c->user_ev[slot_a] || !c->user_ev[slot_b]) return fail(c, VLC_ERR_STATE, "event slot was never recorded");
  CUDA_OK(c, cudaEventSynchronize(c->user_ev[slot_b]));
  float f = 0.f;
  CUDA_OK(c, cudaEventElapsedTime(&f, c->user_ev[slot_a], c->user_ev[slot_b]));
  *ms = f;
  return VLC_OK;
}

// Per-launch device times of the dominant kernels since the last reset, summed by kernel: [0] bs_lattice_kernel,
// [1] bs_sweep_kernel (the launches that cover a set's main part; remainder / tail-strip launches are not counted).
extern "C" int vlc_sweep_stats(vlc_ctx* c, int reset, int64_t* launches, double* ms, double* pairs, double* fp64_instr) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (launches || ms || pairs || fp64_instr) {
    int64_t n[2] = {0, 0};
    double t[4] = {0.0, 0.0, 0.0, 0.0}, pr[4] = {0.0, 0.0, 0.0, 0.0}, in[4] = {0.0, 0.0, 0.0, 0.0};
    for (size_t k = 0; k < c->stats_n; ++k) {
      SweepStat& st = c->stats[k];
      CUDA_OK(c, cudaEventSynchronize(st.e2));
      float f = 0.f, g = 0.f;
      CUDA_OK(c, cudaEventElapsedTime(&f, st.e0, st.e1));
      CUDA_OK(c, cudaEventElapsedTime(&g, st.e0, st.e2));
      n[st.kind]++;
      t[st.kind] += f;
      pr[st.kind] += st.pairs;
      in[st.kind] += st.instr;
      t[2 + st.kind] += g;
      pr[2 + st.kind] += st.pairs_all;
      in[2 + st.kind] += st.instr_all;
    }
    for (int k = 0; k < 2; ++k)
      if (launches) launches[k] = n[k];
    for (int k = 0; k < 4; ++k) {
      if (ms) ms[k] = t[k];
      if (pairs) pairs[k] = pr[k];
      if (fp64_instr) fp64_instr[k] = in[k];
    }
  }
  if (reset) {
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->stats_n = 0;
    c->stats_on = reset > 0;
  }
  return VLC_OK;
}

// Writes a 256 MiB scratch buffer (twice the 126 MB L2 of a B200) on the context's stream: bench.py calls it between
// timed steps so that no step starts with the previous step's sources in L2.
extern "C" int vlc_l2_flush(vlc_ctx* c) {
  CHECK_CTX(c);
  VLC_GROUP(c, vlc_l2_flush(m));
  int rc = bind_device(c);
  if (rc) return rc;
  constexpr size_t kBytes = (size_t)256 << 20;
  if (!c->d_flush) CUDA_OK(c, cudaMalloc(&c->d_flush, kBytes));
  CUDA_OK(c, cudaMemsetAsync(c->d_flush, 0, kBytes, c->stream));
  return VLC_OK;
}

static int measure_fp64(vlc_ctx* c, int iters, int pattern, double* flops_per_s, double* ms_out);
extern "C" int vlc_measure_fp64_peak(vlc_ctx* c, int iters, double* flops_per_s, double* ms_out) {
  return measure_fp64(c, iters, 0, flops_per_s, ms_out);
}
extern "C" int vlc_measure_fp64_rate(vlc_ctx* c, int iters, int pattern, double* flops_per_s, double* ms_out) {
  if (pattern != 0 && pattern != 1) return fail(c, VLC_ERR_ARG, "pattern: 0 = one register operand per DFMA, 1 = three");
  return measure_fp64(c, iters, pattern, flops_per_s, ms_out);
}
static int measure_fp64(vlc_ctx* c, int iters, int pattern, double* flops_per_s, double* ms_out) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (iters < 1) iters = 1;
  const int blocks = c->sm_count * 8, threads = 256;
  if ((rc = reserve(c, c->scratch, (size_t)blocks * threads))) return rc;
  cudaEvent_t e0, e1;
  CUDA_OK(c, cudaEventCreate(&e0));
  CUDA_OK(c, cudaEventCreate(&e1));
  auto kern = pattern ? vlc::dfma_peak3_kernel : vlc::dfma_peak_kernel;  // same DFMA count per iteration
  kern<<<blocks, threads, 0, c->stream>>>(iters / 4 + 1, c->scratch.p);  // warm-up
  CUDA_OK(c, cudaEventRecord(e0, c->stream));
  kern<<<blocks, threads, 0, c->stream>>>(iters, c->scratch.p);
  CUDA_OK(c, cudaEventRecord(e1, c->stream));
  CUDA_OK(c, cudaEventSynchronize(e1));
  CUDA_OK(c, cudaGetLastError());
  c->launches += 2;
  float ms = 0.f;
  CUDA_OK(c, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = (double)blocks * threads * (double)iters * vlc::kPeakChains * vlc::kPeakUnroll * 2.0;
  if (flops_per_s) *flops_per_s = flops / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  return VLC_OK;
}

extern "C" int vlc_probe_rsqrt(vlc_ctx* c, int64_t n, const double* x, double* seed, double* full, double* fast) {
  CHECK_CTX(c);
  int rc = bind_device(c);
  if (rc) return rc;
  if (n <= 0 || !x || !seed || !full || !fast) return fail(c, VLC_ERR_ARG, "bad arguments");
  if ((rc = reserve(c, c->scratch, 4 * (size_t)n))) return rc;
  double* d = c->scratch.p;
  CUDA_OK(c, cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  LAUNCH1D(c, vlc::rsqrt_probe_kernel, n, n, d, d + n, d + 2 * n, d + 3 * n);
  CUDA_OK(c, cudaMemcpyAsync(seed, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaMemcpyAsync(full, d + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaMemcpyAsync(fast, d + 3 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VLC_OK;
}
