// p3cs_runplan.h -- host-side planner of the run kernel (p3cs_run.cuh).  Plain C++ (no CUDA types), so tools/ and
// tests can compile it with g++ and look at a plan without a GPU.
//
// The reference skins a column in RUNS: maximal stretches of consecutive rows that select the same blend get one
// matrix fetch and one table_xform_* call (geomVertexData.cxx:1574-1661).  The run kernel keeps that shape: a
// work ITEM is a piece of such a run -- up to `max_item_rows` consecutive rows of one tile that select the same
// blend -- and belongs to ONE thread, which blends the item's matrix in registers (TransformBlend::update_blend)
// and walks the item's rows.  No per-row record read, no record in shared memory at all.
//
// What is planned here, once per mesh (the row -> blend map is static):
//   * tiles of `tile_rows` consecutive rows (the unit one TMA bulk copy stages);
//   * per tile its items, dealt to SLOTS of 32 lanes.  Lane l of a half-warp takes an item whose first row is
//     congruent to l modulo 16: at every step k of the row walk the sixteen lanes then touch rows that are pairwise
//     different modulo 16, which makes the 64-bit shared-memory accesses of 24- and 56-byte rows, the 128-bit
//     accesses of 48-byte rows and the alternating 128-bit pairs of 32-byte rows bank-conflict free (see
//     RowLayout in p3cs_strip.cuh);
//   * items of equal length share a slot as far as possible (the lanes of a warp walk in lockstep), and slots come
//     in descending length so the warps of a CTA can be dealt slots longest-first.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

namespace p3cs {

constexpr int kItemStartBits = 13;  // first row of the item inside its tile (tile_rows <= 8192)
constexpr int kItemCountBits = 5;   // rows of the item (1..31); 0 = idle lane
constexpr int kItemMaxRows = (1 << kItemCountBits) - 1;
constexpr int kItemEntriesShift = kItemStartBits + kItemCountBits + 1;  // entries of the blend: 0..4, 5 = more than four
constexpr uint32_t kItemSkinned = 1u << (kItemStartBits + kItemCountBits);  // rows inside TransformBlendTable::get_rows(); clear: rows
                                                                              // only sliders (and the bounds) touch

struct RunPlanHost {
  int tile_rows = 0;
  int max_item_rows = 0;
  int num_tiles = 0;
  std::vector<int> tile_slot_offsets;  // [num_tiles + 1]
  std::vector<uint32_t> items;         // [slots * 32]: start | count << 13 | skinned << 18
  std::vector<int32_t> item_blend;     // [slots * 32]: blend id of the item (-1: rows outside get_rows())
  // statistics (tools/plan_stats, P3CS_DEBUG_PLAN)
  long long num_items = 0, num_slots = 0, lane_steps = 0, warp_steps = 0, row_conflict_steps = 0, idle_lanes = 0;
};

struct RunPlanInput {
  int num_rows = 0;
  const uint32_t *row_blend = nullptr;    // transform_blend value of every row
  const uint32_t *live_mask = nullptr;    // bit per row inside get_rows(); nullptr = every row
  bool has_blend_table = true;            // false: no transform_blend column (morph-only meshes)
};

// Splits a run of `len` rows into ceil(len / K) pieces of nearly equal length.
inline void split_run(int len, int K, std::vector<int> &pieces) {
  pieces.clear();
  const int n = (len + K - 1) / K;
  const int base = len / n, extra = len % n;
  for (int i = 0; i < n; ++i) pieces.push_back(base + (i < extra ? 1 : 0));
}

inline RunPlanHost build_run_plan(const RunPlanInput &in, int tile_rows, int max_item_rows) {
  RunPlanHost p;
  p.tile_rows = tile_rows;
  p.max_item_rows = std::min(max_item_rows, kItemMaxRows);
  const int K = p.max_item_rows;
  const int N = in.num_rows;
  p.num_tiles = (N + tile_rows - 1) / tile_rows;
  p.tile_slot_offsets.assign(1, 0);
  auto live = [&](int r) { return in.has_blend_table && (in.live_mask == nullptr || ((in.live_mask[r >> 5] >> (r & 31)) & 1u)); };

  struct Item {
    int start, count, blend;
  };
  std::vector<Item> bucket[16];
  std::vector<int> pieces;
  for (int t = 0; t < p.num_tiles; ++t) {
    const int r0 = t * tile_rows, r1 = std::min(N, r0 + tile_rows);
    for (auto &b : bucket) b.clear();
    // runs of the tile -> items
    int r = r0;
    while (r < r1) {  // every row belongs to exactly one item: skinned runs, and stretches outside get_rows() (blend -1)
      const int blend = live(r) ? (int)in.row_blend[r] : -1;
      int e = r + 1;
      if (blend >= 0) while (e < r1 && live(e) && (int)in.row_blend[e] == blend) ++e;
      else while (e < r1 && !live(e)) ++e;
      split_run(e - r, K, pieces);
      int s = r;
      for (int len : pieces) {
        bucket[(s - r0) & 15].push_back(Item{s - r0, len, blend});
        s += len;
        ++p.num_items;
      }
      r = e;
    }
    // longest first inside every residue class, then half-warp h takes the h-th item of each class
    size_t halves = 0;
    for (auto &b : bucket) {
      std::stable_sort(b.begin(), b.end(), [](const Item &x, const Item &y) { return x.count > y.count; });
      halves = std::max(halves, b.size());
    }
    const size_t slots = (halves + 1) / 2;
    const size_t base = p.items.size();
    p.items.resize(base + slots * 32, 0u);
    p.item_blend.resize(base + slots * 32, -1);
    for (size_t h = 0; h < slots * 2; ++h) {
      for (int l = 0; l < 16; ++l) {
        const size_t at = base + h * 16 + l;
        if (h < bucket[l].size()) {
          const Item &it = bucket[l][h];
          p.items[at] = (uint32_t)it.start | ((uint32_t)it.count << kItemStartBits) | (it.blend >= 0 ? kItemSkinned : 0u);
          p.item_blend[at] = it.blend;
          p.lane_steps += it.count;
        } else {
          ++p.idle_lanes;
        }
      }
    }
    for (size_t s = 0; s < slots; ++s) {
      int longest = 0;
      for (int l = 0; l < 32; ++l) longest = std::max(longest, (int)((p.items[base + s * 32 + l] >> kItemStartBits) & kItemMaxRows));
      p.warp_steps += longest;
    }
    p.num_slots += (long long)slots;
    p.tile_slot_offsets.push_back((int)(p.items.size() / 32));
  }
  return p;
}

inline void print_run_plan_stats(const RunPlanHost &p, int num_rows, FILE *f) {
  fprintf(f,
          "run plan: %d rows, tiles of %d rows (%d tiles), items of <= %d rows: %lld items (%.2f rows each), %lld slots (%.2f per tile), "
          "idle lanes %.1f %%, lockstep efficiency %.1f %% (%lld lane steps / %lld warp steps x 32)\n",
          num_rows, p.tile_rows, p.num_tiles, p.max_item_rows, p.num_items, p.num_items ? (double)p.lane_steps / p.num_items : 0.0, p.num_slots,
          p.num_tiles ? (double)p.num_slots / p.num_tiles : 0.0, p.num_slots ? 100.0 * p.idle_lanes / (p.num_slots * 32.0) : 0.0,
          p.warp_steps ? 100.0 * p.lane_steps / (p.warp_steps * 32.0) : 0.0, p.lane_steps, p.warp_steps);
}

}  // namespace p3cs
