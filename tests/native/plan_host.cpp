// TEST INFRASTRUCTURE: the product's launch planner (volcanor_b200/csrc/plan.hpp, host code without CUDA) behind a C
// interface for tests/test_plan_host.py.
#include "../../volcanor_b200/csrc/plan.hpp"

extern "C" {
int plan_small_split(long long target_tiles, long long src_tiles, long long slots, int per_tile) {
  return vlc::plan::small_split(target_tiles, src_tiles, slots, per_tile);
}
int plan_wave_split(long long target_tiles, long long src_tiles, long long slots, long long max_split, int per_tile) {
  return vlc::plan::wave_split(target_tiles, src_tiles, slots, max_split, per_tile);
}
void plan_cut(long long n_pad, long long unit, int nsplit, int* nsplit_out, long long* chunk_out) {
  const vlc::plan::Cut c = vlc::plan::cut(n_pad, unit, nsplit);
  *nsplit_out = c.nsplit;
  *chunk_out = c.chunk;
}
}
