// tools/plan_stats.cpp -- offline look at a run plan (csrc/p3cs_runplan.h): g++ -O2 -o /tmp/plan_stats tools/plan_stats.cpp
// usage: plan_stats <rows.u32 file> <tile_rows> <max_item_rows>
#include "../panda3d_b200/csrc/p3cs_runplan.h"
#include <cstdlib>
int main(int argc, char **argv) {
  if (argc < 4) return 1;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<uint32_t> idx;
  uint32_t v;
  while (fread(&v, 4, 1, f) == 1) idx.push_back(v);
  fclose(f);
  p3cs::RunPlanInput in;
  in.num_rows = (int)idx.size();
  in.row_blend = idx.data();
  p3cs::RunPlanHost p = p3cs::build_run_plan(in, atoi(argv[2]), atoi(argv[3]));
  p3cs::print_run_plan_stats(p, in.num_rows, stdout);
  return 0;
}
