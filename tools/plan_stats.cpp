// tools/plan_stats.cpp -- offline look at a run plan (csrc/p3cs_runplan.h): g++ -O2 -o /tmp/plan_stats tools/plan_stats.cpp
// usage: plan_stats <rows.u32> <blend_offsets.i32> <blend_transform.i32> <tile_rows> <max_item_rows> [trim] [T: two palette copies]
#include "../panda3d_b200/csrc/p3cs_runplan.h"
#include <chrono>
#include <cstdlib>
template <class T> static std::vector<T> slurp(const char *path) {
  std::vector<T> v;
  FILE *f = fopen(path, "rb");
  if (!f) return v;
  T x;
  while (fread(&x, sizeof(T), 1, f) == 1) v.push_back(x);
  fclose(f);
  return v;
}
int main(int argc, char **argv) {
  if (argc < 6) return 1;
  std::vector<uint32_t> idx = slurp<uint32_t>(argv[1]);
  std::vector<int> off = slurp<int>(argv[2]), xf = slurp<int>(argv[3]);
  p3cs::RunPlanInput in;
  in.num_rows = (int)idx.size();
  in.row_blend = idx.data();
  in.blend_offsets = off.data();
  in.blend_transform = xf.data();
  if (argc > 7 && atoi(argv[7]) > 0) {
    const int T = atoi(argv[7]);
    in.copy_b_base = (T + 7) & ~7;
    long long runs = 1;
    for (size_t r = 1; r < idx.size(); ++r) runs += idx[r] != idx[r - 1];
    in.match_every_candidate = (long long)idx.size() >= 3 * runs;  // as p3cs_mesh.cu does
  }
  auto t0 = std::chrono::steady_clock::now();
  p3cs::RunPlanHost p = p3cs::build_run_plan(in, atoi(argv[4]), atoi(argv[5]), argc > 6 ? atoi(argv[6]) : 0);
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  p3cs::print_run_plan_stats(p, in.num_rows, stdout);
  const long long issue = p.num_slots * 330 + p.warp_steps * 70, lsu = p.gather_wavefronts + p.warp_steps * 12;
  printf("  planned in %.1f ms; per 32 rows: issue %.1f, lsu wavefronts %.1f, cycles (issue / 4 + lsu) %.1f\n", ms, issue * 32.0 / in.num_rows, lsu * 32.0 / in.num_rows,
         (issue / 4.0 + lsu) * 32.0 / in.num_rows);
  return 0;
}
