// tests/run_plan_check.cpp -- CPU check of the run kernel's host planner (panda3d_b200/csrc/p3cs_runplan.h), built and run by
// tests/test_run_plan.py with g++ (no CUDA needed).  Prints "ok ..." lines; any failure prints "FAIL ..." and exits 1.
#include "../panda3d_b200/csrc/p3cs_runplan.h"

#include <cstdlib>
#include <random>
#include <set>

static int fails = 0;
#define CHECK(cond, ...)                     \
  do {                                       \
    if (!(cond)) {                           \
      printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                   \
      printf("\n");                          \
      if (++fails > 20) exit(1);             \
    }                                        \
  } while (0)

// exhaustive optimum of the two-copy gather: every distinct joint takes copy A or B, cost = fullest bank group
static int brute_force(const int *joints, int n, int base) {
  int best = 99;
  for (int mask = 0; mask < (1 << n); ++mask) {
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 0;
    for (int i = 0; i < n; ++i) {
      const int idx = ((mask >> i) & 1) && base > 0 ? p3cs::palette_copy_b_index(joints[i], base) : joints[i];
      worst = std::max(worst, ++load[idx & 7]);
    }
    best = std::min(best, worst);
  }
  return best;
}

static void check_matching(std::mt19937 &rng) {
  const int Ts[] = {8, 16, 26, 40, 64, 128, 200, 1000};
  long long cases = 0;
  for (int T : Ts)
    for (int two = 0; two < 2; ++two) {
      const int base = two ? (T + 7) & ~7 : 0;
      // copy B is a permutation of the joints inside [base, 2 * base), and palette_area_joint undoes it
      if (two) {
        std::set<int> seen;
        for (int j = 0; j < T; ++j) {
          const int idx = p3cs::palette_copy_b_index(j, base);
          CHECK(idx >= base && idx < 2 * base && seen.insert(idx).second, "copy B index of joint %d (T=%d): %d", j, T, idx);
          CHECK(p3cs::palette_area_joint(idx, base) == j && p3cs::palette_area_joint(j, base) == j, "palette_area_joint(%d) != %d", idx, j);
        }
      }
      for (int rep = 0; rep < 400; ++rep) {
        int n = 1 + (int)(rng() % 8), joints[8], choice[8];
        std::set<int> distinct;
        while ((int)distinct.size() < std::min(n, T)) distinct.insert((int)(rng() % T));
        n = 0;
        for (int j : distinct) joints[n++] = j;
        const int got = p3cs::two_copy_gather(joints, n, base, choice);
        const int want = brute_force(joints, n, base);
        CHECK(got == want, "two_copy_gather T=%d two=%d n=%d: %d, optimum %d", T, two, n, got, want);
        int load[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 0;
        for (int i = 0; i < n; ++i) {
          CHECK(two || choice[i] == 0, "copy B chosen without a copy B");
          const int idx = choice[i] ? p3cs::palette_copy_b_index(joints[i], base) : joints[i];
          worst = std::max(worst, ++load[idx & 7]);
        }
        CHECK(worst == got, "the returned choice costs %d, not %d", worst, got);
        ++cases;
      }
    }
  printf("ok matching: %lld cases equal the exhaustive optimum\n", cases);
}

static void check_plan(std::mt19937 &rng, int N, int T, int B, int order, bool sparse, int tile_rows, int K, bool two, int trim) {
  std::vector<uint32_t> row_blend(N), mask((N + 31) / 32, 0xFFFFFFFFu);
  std::vector<int> off(B + 1, 0), xf;
  for (int b = 0; b < B; ++b) {
    const int k = 1 + (int)(rng() % 6);  // blends of 1..6 entries
    for (int e = 0; e < k; ++e) xf.push_back((int)(rng() % T));
    off[b + 1] = (int)xf.size();
  }
  for (int r = 0; r < N;) {
    const int len = order == 0 ? 1 : order == 1 ? 1 + (int)(rng() % 14) : 40 + (int)(rng() % 60);
    const uint32_t b = rng() % B;
    for (int e = std::min(N, r + len); r < e; ++r) row_blend[r] = b;
  }
  if (sparse)
    for (int r = 0; r < N; ++r)
      if ((r / 37) % 5 == 2) mask[r >> 5] &= ~(1u << (r & 31));
  p3cs::RunPlanInput in;
  in.num_rows = N;
  in.row_blend = row_blend.data();
  in.live_mask = sparse ? mask.data() : nullptr;
  in.blend_offsets = off.data();
  in.blend_transform = xf.data();
  if (two) in.copy_b_base = (T + 7) & ~7;
  const p3cs::RunPlanHost p = p3cs::build_run_plan(in, tile_rows, K, trim);
  const int tiles = (N + tile_rows - 1) / tile_rows;
  CHECK(p.num_tiles == tiles && (int)p.tile_slot_offsets.size() == tiles + 1, "tile count");
  CHECK(p.items.size() == p.item_blend.size() && p.items.size() == p.item_copy.size() && p.items.size() % 32 == 0, "array sizes");
  CHECK(p.tile_slot_offsets.back() * 32 == (int)p.items.size(), "last slot offset");
  std::vector<int> covered(N, 0);
  long long gather = 0;
  for (int t = 0; t < tiles; ++t) {
    CHECK(p.tile_slot_offsets[t] <= p.tile_slot_offsets[t + 1], "slot offsets not monotone");
    const int rows_here = std::min(tile_rows, N - t * tile_rows);
    for (int s = p.tile_slot_offsets[t]; s < p.tile_slot_offsets[t + 1]; ++s) {
      for (int half = 0; half < 2; ++half) {
        // the conflict-free walk: at every step the active lanes of a half-warp are on rows that differ modulo 16
        for (int step = 0; step < 16 + K; ++step) {
          bool used[16] = {false};
          for (int l = half * 16; l < half * 16 + 16; ++l) {
            const uint32_t it = p.items[(size_t)s * 32 + l];
            const int cnt = (int)((it >> p3cs::kItemStartBits) & (uint32_t)p3cs::kItemMaxRows);
            const int start = (int)(it & ((1u << p3cs::kItemStartBits) - 1u)), delay = (int)((it >> p3cs::kItemDelayShift) & 15u);
            const int k = step - delay;
            if (cnt == 0 || k < 0 || k >= cnt) continue;
            CHECK(!used[(start + k) & 15], "tile %d slot %d: two lanes of a half-warp on the same row residue at step %d", t, s, step);
            used[(start + k) & 15] = true;
          }
        }
      }
      for (int l = 0; l < 32; ++l) {
        const size_t at = (size_t)s * 32 + l;
        const uint32_t it = p.items[at];
        const int cnt = (int)((it >> p3cs::kItemStartBits) & (uint32_t)p3cs::kItemMaxRows);
        if (cnt == 0) {
          CHECK(it == 0 && p.item_copy[at] == 0, "idle lane carries data");
          continue;
        }
        const int start = (int)(it & ((1u << p3cs::kItemStartBits) - 1u));
        const bool skinned = (it & p3cs::kItemSkinned) != 0;
        const int blend = p.item_blend[at];
        CHECK(cnt <= K && start + cnt <= rows_here, "item outside its tile: start %d rows %d", start, cnt);
        CHECK(skinned == (blend >= 0), "skinned flag");
        for (int k = 0; k < cnt; ++k) {
          const int r = t * tile_rows + start + k;
          ++covered[r];
          const bool live = !sparse || ((mask[r >> 5] >> (r & 31)) & 1u);
          CHECK(live == skinned, "row %d: live %d, item skinned %d", r, (int)live, (int)skinned);
          if (skinned) CHECK((int)row_blend[r] == blend, "row %d selects blend %u, its item %d", r, row_blend[r], blend);
        }
        const int entries = blend >= 0 ? std::min(off[blend + 1] - off[blend], 4) : 0;
        CHECK((p.item_copy[at] >> entries) == 0, "copy bit beyond the blend's entries");
        CHECK(two || p.item_copy[at] == 0, "copy bits without a second copy");
      }
      // gather wavefronts of the slot, recomputed from the emitted items and copy bits
      for (int q = 0; q < 4; ++q)
        for (int k = 0; k < 4; ++k) {
          std::set<int> group[8];
          for (int l = q * 8; l < q * 8 + 8; ++l) {
            const size_t at = (size_t)s * 32 + l;
            const int blend = p.item_blend[at];
            if (((p.items[at] >> p3cs::kItemStartBits) & (uint32_t)p3cs::kItemMaxRows) == 0 || blend < 0 || off[blend + 1] - off[blend] <= k) continue;
            const int j = xf[off[blend] + k];
            const int idx = ((p.item_copy[at] >> k) & 1) ? p3cs::palette_copy_b_index(j, in.copy_b_base) : j;
            group[idx & 7].insert(idx);
          }
          size_t worst = 0;
          for (auto &g : group) worst = std::max(worst, g.size());
          gather += 3 * (long long)worst;
        }
    }
  }
  for (int r = 0; r < N; ++r) CHECK(covered[r] == 1, "row %d covered %d times", r, covered[r]);
  CHECK(gather == p.gather_wavefronts, "gather wavefronts: emitted items give %lld, the planner reports %lld", gather, p.gather_wavefronts);
  printf("ok plan: N=%d T=%d B=%d order=%d sparse=%d tile=%d K=%d copies=%d: %lld items, %lld slots, lockstep %.1f %%, gather %lld (conflict-free %lld)\n", N, T, B,
         order, (int)sparse, tile_rows, K, two ? 2 : 1, p.num_items, p.num_slots, p.warp_steps ? 100.0 * p.lane_steps / (32.0 * p.warp_steps) : 0.0,
         p.gather_wavefronts, p.gather_ideal);
}

// every (instance, tile) of a launch belongs to exactly one CTA, with and without the tail split
static void check_units(int count, int num_tiles, int tiles_per_cta, int whole, int parts, int tail_tpc) {
  std::vector<int> covered((size_t)count * num_tiles, 0);
  long long ctas = 0;
  auto take = [&](int bx, int by) {
    int inst, t0, t1;
    p3cs::run_unit_decode(bx, by, num_tiles, tiles_per_cta, whole, parts, tail_tpc, &inst, &t0, &t1);
    CHECK(inst >= 0 && inst < count, "unit (%d,%d): instance %d of %d", bx, by, inst, count);
    for (int t = t0; t < t1; ++t) ++covered[(size_t)inst * num_tiles + t];
    ++ctas;
  };
  if (whole >= 0) {
    for (int u = 0; u < whole + (count - whole) * parts; ++u) take(u, 0);
  } else {
    const int cpi = (num_tiles + tiles_per_cta - 1) / tiles_per_cta;
    for (int y = 0; y < count; ++y)
      for (int x = 0; x < cpi; ++x) take(x, y);
  }
  for (size_t i = 0; i < covered.size(); ++i) CHECK(covered[i] == 1, "instance %zu tile %zu covered %d times", i / num_tiles, i % num_tiles, covered[i]);
  printf("ok units: %d instances x %d tiles, whole %d, parts %d x %d tiles: %lld CTAs\n", count, num_tiles, whole, parts, tail_tpc, ctas);
}

int main() {
  std::mt19937 rng(20261020);
  check_matching(rng);
  check_units(1250, 17, 17, 806, 2, 9);
  check_units(10000, 17, 17, 9556, 2, 9);
  check_units(900, 5, 5, 456, 3, 2);
  check_units(444, 4, 4, 0, 4, 1);
  check_units(37, 17, 6, -1, 0, 0);
  check_units(3, 391, 1, -1, 0, 0);
  for (int two = 0; two < 2; ++two) {
    check_plan(rng, 20000, 128, 4096, 1, false, 1248, 12, two, 4);
    check_plan(rng, 20000, 128, 4096, 0, false, 1120, 12, two, 4);
    check_plan(rng, 4571, 40, 700, 1, true, 256, 12, two, 4);
    check_plan(rng, 4571, 40, 700, 2, true, 608, 12, two, 0);
    check_plan(rng, 1338, 26, 197, 1, false, 256, 12, two, 4);
    check_plan(rng, 33, 9, 5, 0, true, 16, 3, two, 4);
    check_plan(rng, 100000, 64, 16384, 0, false, 2032, 31, two, 4);
  }
  if (fails) return 1;
  printf("ok all\n");
  return 0;
}
