// plan.hpp -- how a sweep is cut into CTAs: source splits (grid y) for a given number of target tiles (grid x) and CTA
// slots of the device.  Host code without CUDA: compiled by nvcc into the library (capi.cu) and by g++ into
// tests/native/plan_host.cpp (CPU tests of the planner).
//
// The reference has no counterpart: its sweep is an OpenMP loop over targets (libCommon.f90:132-139).  Here a target is
// summed over source chunks by different CTAs and the partial sums are added in chunk order (bs_reduce_kernel /
// bs_reduce_select_kernel), so the split decides both the machine fill and the summation order of a target; a split fixed by
// vlc_set_tuning makes the order independent of the launch (bit-identical results for any sharding of the targets).
#pragma once

#include <algorithm>
#include <cmath>

namespace vlc {
namespace plan {

constexpr double kSmallC0 = 0.3;  // fixed cost of a CTA of a small sweep, in source tiles (plan_small_split)

// Source split of a SMALL sweep: one that cannot fill the machine for two waves with chunks of >= 4 tiles.  Parallelism
// then matters more than the per-CTA prologue: the split is the one with the least estimated time
// ceil(waves) * (chunk + c0), chunk in UNITS of `per_tile` to a tile (1 for the flat kernel; 4 for the lattice kernel, whose
// chunks are multiples of a quarter tile), c0 ~ prologue + epilogue of a CTA in tiles (ncu launch list of K&P, r02s: the
// wake sweeps of a 4 000-node wake ran 7 splits x 15 target tiles = 105 CTAs on 296 slots for 106 us; 5e7 pairs are 53 us
// of the whole machine; scans of fixed splits in profiles/r02v_small_cases.md).  Returns 0 when the sweep is not small
// (the search for whole waves applies).
inline int small_split(long long target_tiles, long long src_tiles, long long slots, int per_tile = 1) {
  const long long by4 = std::max(1LL, src_tiles / 4);
  if (target_tiles * by4 >= 2 * slots || src_tiles * per_tile <= 1) return 0;
  const double c0 = kSmallC0 * per_tile;
  const long long units = src_tiles * per_tile;
  double best = 1e300;
  int best_s = 1;
  const long long max_split = std::min(units, 256LL);
  for (long long s = 1; s <= max_split; ++s) {
    const long long chunk = (units + s - 1) / s, real = (units + chunk - 1) / chunk;
    if (real != s) continue;
    const double waves = (double)target_tiles * (double)real / (double)slots;
    const double cost = std::ceil(waves - 1e-9) * ((double)chunk + c0);
    if (cost < best * (1.0 - 1e-9)) {
      best = cost;
      best_s = (int)s;
    }
  }
  return best_s;
}

// Source split of a sweep that fills the machine: chunks of >= 4 tiles (cut in units of `per_tile` to a tile, like
// small_split), at most `max_split` splits (the caller's cap on the partial-sum buffer), the split whose CTA count is closest
// below a whole number of waves -- equal-work CTAs leave a (1 - eff) tail idle -- with a mild preference for fewer splits.
// Finer units reach more CTA counts: 95 target tiles x 175 source tiles (a late caradonna step) fit 296 slots as 3 splits
// (0.963 of a wave) in tiles, as 28 splits (8.99 waves) in quarter tiles.
inline int wave_split(long long target_tiles, long long src_tiles, long long slots, long long max_split = 256, int per_tile = 1) {
  long long cap = src_tiles / 4;
  cap = std::max(1LL, std::min(cap, std::min(max_split, 256LL)));
  const long long units = src_tiles * per_tile;
  double best = -1.0;
  int best_s = 1;
  for (long long s = 1; s <= cap; ++s) {
    const long long chunk = (units + s - 1) / s;
    const long long real_s = (units + chunk - 1) / chunk;
    if (real_s != s) continue;
    const double waves = (double)target_tiles * (double)real_s / (double)slots;
    const double eff = waves / (double)(long long)(waves + 0.999999);
    const double score = (waves >= 1.0) ? eff - 1e-4 * (double)s : eff;
    if (score > best + 1e-9) {
      best = score;
      best_s = (int)s;
    }
    if (waves >= 8.0 && eff > 0.995) break;
  }
  return best_s;
}

// `n_pad` records cut into chunks that are multiples of `unit` records: the number of chunks actually needed for the
// requested split and the chunk length (the last chunk may be shorter).
struct Cut {
  int nsplit = 1;
  long long chunk = 0;
};
inline Cut cut(long long n_pad, long long unit, int nsplit) {
  Cut c;
  if (unit < 1) unit = 1;
  const long long units = std::max(1LL, n_pad / unit);
  if (nsplit < 1) nsplit = 1;
  const long long per = (units + nsplit - 1) / nsplit;
  c.nsplit = (int)((units + per - 1) / per);
  c.chunk = per * unit;
  return c;
}

}  // namespace plan
}  // namespace vlc
