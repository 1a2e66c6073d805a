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
//   * per tile its items, dealt to SLOTS of 32 lanes that walk their items in lockstep.  Lane l of a half-warp
//     touches, at step j of the walk, a row congruent to l + j modulo 16: its item starts d = (first row - l) mod 16
//     steps late (`delay`).  The sixteen lanes of a half-warp are then on pairwise different rows modulo 16 at
//     every step, which makes the 64-bit shared-memory accesses of 24- and 56-byte rows, the 128-bit accesses of
//     48-byte rows and the alternating 128-bit pairs of 32-byte rows bank-conflict free (RowLayout, p3cs_strip.cuh);
//   * a slot lasts as long as its longest delay + item, so items are placed longest first, each into the lane that
//     keeps its slot shortest; among equally good lanes the one whose palette indices collide least, modulo 8, with
//     the items already in the same quarter-warp (the palette gather is a 128-bit load per lane, eight lanes per
//     wavefront, 48-byte palette rows: two different joints collide when they agree modulo 8);
//   * TWO COPIES of the palette (RunPlanInput::copy_b_base > 0): the kernel stages the instance's palette twice, the
//     second time permuted inside every octet of joints -- joint j = 8a + c sits at entry j of copy A (bank group c)
//     and at entry copy_b_base + 8a + ((a + 3c) mod 8) of copy B (bank group (a + 3c) mod 8): two joints that share a
//     group in one copy never share it in the other (unless they are a multiple of 64 apart).  Which copy a lane reads an entry's joint from is free, and static: per quarter-warp and entry the
//     planner picks the assignment with the fewest wavefronts (two choices per joint, a small matching problem), and
//     the item simply carries the chosen palette INDEX.  Random joints: 2.1 wavefronts per gather load -> ~1.2.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#ifdef __CUDACC__
#define P3CS_PLAN_HD __host__ __device__
#else
#define P3CS_PLAN_HD
#endif

namespace p3cs {

constexpr int kItemStartBits = 13;  // first row of the item inside its tile (tile_rows <= 8192)
constexpr int kItemCountBits = 5;   // rows of the item (1..31); 0 = idle lane
constexpr int kItemMaxRows = (1 << kItemCountBits) - 1;
constexpr int kItemEntriesShift = kItemStartBits + kItemCountBits + 1;  // entries of the blend: 0..4, 5 = more than four
constexpr int kItemDelayShift = kItemEntriesShift + 3;                  // steps the lane waits before its first row (0..15)
constexpr uint32_t kItemSkinned = 1u << (kItemStartBits + kItemCountBits);  // rows inside TransformBlendTable::get_rows(); clear: rows
                                                                              // only sliders (and the bounds) touch

struct RunPlanHost {
  int tile_rows = 0;
  int max_item_rows = 0;
  int num_tiles = 0;
  std::vector<int> tile_slot_offsets;  // [num_tiles + 1]
  std::vector<uint32_t> items;         // [slots * 32]: start | count << 13 | skinned << 18 | delay << 22 (entries << 19 added at upload)
  std::vector<int32_t> item_blend;     // [slots * 32]: blend id of the item (-1: rows outside get_rows())
  std::vector<uint8_t> item_copy;       // [slots * 32]: bit k set = entry k of the item's blend is read from palette copy B
  // statistics (tools/plan_stats, P3CS_DEBUG_PLAN)
  long long num_items = 0, num_slots = 0, lane_steps = 0, warp_steps = 0, idle_lanes = 0, gather_wavefronts = 0, gather_ideal = 0;
};

struct RunPlanInput {
  int num_rows = 0;
  const uint32_t *row_blend = nullptr;    // transform_blend value of every row
  const uint32_t *live_mask = nullptr;    // bit per row inside get_rows(); nullptr = every row
  bool has_blend_table = true;            // false: no transform_blend column (morph-only meshes)
  // the TransformBlendTable (CSR), for the palette-gather term of the placement; may be null
  const int *blend_offsets = nullptr;
  const int *blend_transform = nullptr;
  // two palette copies (see the header comment): first entry of copy B (a multiple of 8; 0 = one copy)
  int copy_b_base = 0;
  // Placement with two copies: ask the matching for every candidate item (true), or guide the placement by the copy-A
  // count alone and match once per finished slot (false).  The first is worth 1-2 % fewer slots on meshes with runs; on
  // meshes of one-row items (a random blend per vertex: every item is a candidate for every lane) it costs seven times
  // the planning time -- 1.3 s for 100 000 rows -- for 1 % fewer gather wavefronts.
  bool match_every_candidate = true;
};

// Which instance and which tiles a CTA of a run-kernel launch takes (p3cs_run.cuh, RunGeom).  Two-dimensional grid
// (whole_instances < 0): CTA (x, y) takes tiles [x * tiles_per_cta, ...) of instance y.  Tail split (one-dimensional
// grid): CTA u < whole_instances takes all of instance u; the CTAs behind take the remaining instances in `tail_parts`
// pieces of `tail_tiles_per_cta` tiles each, piece after piece, instance after instance.
P3CS_PLAN_HD inline void run_unit_decode(int block_x, int block_y, int num_tiles, int tiles_per_cta, int whole_instances, int tail_parts,
                                         int tail_tiles_per_cta, int *instance, int *tile_begin, int *tile_end) {
  int part = block_x, inst = block_y, tiles = tiles_per_cta;
  if (whole_instances >= 0) {
    if (block_x < whole_instances) {
      inst = block_x;
      part = 0;
      tiles = num_tiles;
    } else {
      const int v = block_x - whole_instances;
      inst = whole_instances + v / tail_parts;
      part = v % tail_parts;
      tiles = tail_tiles_per_cta;
    }
  }
  *instance = inst;
  *tile_begin = part * tiles;
  *tile_end = *tile_begin + tiles < num_tiles ? *tile_begin + tiles : num_tiles;
}

// Entry of joint j in palette copy B.
P3CS_PLAN_HD inline int palette_copy_b_index(int j, int copy_b_base) { return copy_b_base + (j & ~7) + (((j >> 3) + 3 * (j & 7)) & 7); }
// ... and back: the joint behind an index of the palette area (3 * 3 = 1 modulo 8).
P3CS_PLAN_HD inline int palette_area_joint(int index, int copy_b_base) {
  if (copy_b_base <= 0 || index < copy_b_base) return index;
  const int i = index - copy_b_base;
  return (i & ~7) + ((3 * ((i & 7) - (i >> 3))) & 7);
}

// The fewest wavefronts of one gather load of a quarter-warp: `n` distinct joints (n <= 8), each read from copy A or
// copy B; a wavefront serves one entry per bank group (index mod 8).  Smallest L such that every joint can take one of
// its two groups with no group taken more than L times -- augmenting paths, capacities L.  `choice[i]`: 1 = copy B.
inline int two_copy_gather(const int *joints, int n, int copy_b_base, int *choice) {
  if (n <= 0) return 0;
  int ga[8], gb[8];
  for (int i = 0; i < n; ++i) {
    ga[i] = joints[i] & 7;
    gb[i] = copy_b_base > 0 ? palette_copy_b_index(joints[i], copy_b_base) & 7 : ga[i];
  }
  for (int L = 1; L <= 8; ++L) {
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0}, at[8];
    for (int i = 0; i < n; ++i) at[i] = -1;
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
      // breadth-first search for an augmenting path from joint i to a group with spare capacity
      int parent_joint[8], via_group[8], queue[8], qh = 0, qt = 0;
      bool seen_group[8] = {false, false, false, false, false, false, false, false};
      bool seen_joint[8] = {false, false, false, false, false, false, false, false};
      queue[qt++] = i;
      seen_joint[i] = true;
      parent_joint[i] = -1;
      via_group[i] = -1;
      int end_joint = -1, end_group = -1;
      while (qh < qt && end_joint < 0) {
        const int u = queue[qh++];
        const int cand[2] = {ga[u], gb[u]};
        for (int c = 0; c < 2 && end_joint < 0; ++c) {
          const int g = cand[c];
          if (seen_group[g]) continue;
          seen_group[g] = true;
          if (load[g] < L) {
            end_joint = u;
            end_group = g;
            break;
          }
          for (int v = 0; v < n; ++v)  // every joint sitting in g may move to its other group
            if (at[v] == g && !seen_joint[v]) {
              seen_joint[v] = true;
              parent_joint[v] = u;
              via_group[v] = g;
              queue[qt++] = v;
            }
        }
      }
      if (end_joint < 0) {
        ok = false;
        break;
      }
      // shift along the path: end_joint takes end_group, its predecessor takes the group end_joint leaves, ...
      int u = end_joint, g = end_group;
      ++load[g];
      while (u >= 0) {
        const int left = at[u];
        at[u] = g;
        const int pu = parent_joint[u];
        if (pu < 0) break;
        g = via_group[u];  // == left: the group u sat in, now taken over by its parent
        (void)left;
        u = pu;
      }
    }
    if (ok) {
      for (int i = 0; i < n; ++i) choice[i] = (at[i] == ga[i]) ? 0 : 1;
      return L;
    }
  }
  for (int i = 0; i < n; ++i) choice[i] = 0;
  return n;
}

// Splits a run of `len` rows into ceil(len / K) pieces of nearly equal length.
inline void split_run(int len, int K, std::vector<int> &pieces) {
  pieces.clear();
  const int n = (len + K - 1) / K;
  const int base = len / n, extra = len % n;
  for (int i = 0; i < n; ++i) pieces.push_back(base + (i < extra ? 1 : 0));
}

// Relative costs the placement weighs (issue slots + load/store wavefronts of the run kernel, measured with ncu):
constexpr int kPlanSlotCost = 380;   // opening a slot: blending a matrix per lane, 12 gather loads, bookkeeping
constexpr int kPlanStepCost = 82;    // one more step of a slot: a row per lane
constexpr int kPlanGatherCost = 3;   // one more wavefront of a palette gather (three 128-bit loads per entry)

inline RunPlanHost build_run_plan(const RunPlanInput &in, int tile_rows, int max_item_rows, int trim_below = 0) {
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
    int joint[4];  // first four palette indices, -1 beyond the blend's entries
  };
  struct Slot {
    int steps = 0;         // max over its lanes of delay + count
    int used = 0;          // occupied lanes
    int lane_item[32];     // index into `items`, -1 free
    int lane_delay[32];
    // per quarter-warp (4) x entry (4) x residue (8): the distinct joints seen (up to 8 each) -- for the gather term
    int seen[4][4][8][8];
    int nseen[4][4][8];
  };
  std::vector<Item> items;
  std::vector<int> pieces;
  std::vector<int> bucket[16];
  std::vector<Slot> slots;
  auto gather_extra = [&](const Slot &sl, int quarter, const Item &it) {  // extra gather wavefronts if `it` joins this quarter-warp
    int extra = 0;
    for (int k = 0; k < 4; ++k) {
      const int j = it.joint[k];
      if (j < 0) break;
      const int res = j & 7;
      bool known = false;
      for (int q = 0; q < sl.nseen[quarter][k][res]; ++q) known = known || sl.seen[quarter][k][res][q] == j;
      if (known) continue;  // same joint: a broadcast
      if (in.copy_b_base > 0 && in.match_every_candidate) {  // two palette copies: the best assignment with and without the newcomer
        int distinct[9], nd = 0, choice[8];
        for (int r8 = 0; r8 < 8; ++r8)
          for (int q = 0; q < sl.nseen[quarter][k][r8]; ++q) distinct[nd++] = sl.seen[quarter][k][r8][q];
        if (nd >= 8) continue;
        const int before = two_copy_gather(distinct, nd, in.copy_b_base, choice);
        distinct[nd] = j;
        extra += two_copy_gather(distinct, nd + 1, in.copy_b_base, choice) - std::max(before, 1);
        continue;
      }
      int worst = 0;
      for (int r8 = 0; r8 < 8; ++r8) worst = std::max(worst, sl.nseen[quarter][k][r8]);
      if (sl.nseen[quarter][k][res] + 1 > std::max(worst, 1)) ++extra;
    }
    return extra;
  };
  auto place = [&](Slot &sl, int lane, int idx, int delay) {
    const Item &it = items[idx];
    sl.lane_item[lane] = idx;
    sl.lane_delay[lane] = delay;
    sl.steps = std::max(sl.steps, delay + it.count);
    ++sl.used;
    const int quarter = lane >> 3;
    for (int k = 0; k < 4 && it.joint[k] >= 0; ++k) {
      const int j = it.joint[k], res = j & 7;
      bool known = false;
      for (int q = 0; q < sl.nseen[quarter][k][res]; ++q) known = known || sl.seen[quarter][k][res][q] == j;
      if (!known && sl.nseen[quarter][k][res] < 8) sl.seen[quarter][k][res][sl.nseen[quarter][k][res]++] = j;
    }
  };

  for (int t = 0; t < p.num_tiles; ++t) {
    const int r0 = t * tile_rows, r1 = std::min(N, r0 + tile_rows);
    items.clear();
    // runs of the tile -> items: every row belongs to exactly one item (skinned runs; stretches outside get_rows(): blend -1)
    int r = r0;
    while (r < r1) {
      const int blend = live(r) ? (int)in.row_blend[r] : -1;
      int e = r + 1;
      if (blend >= 0) while (e < r1 && live(e) && (int)in.row_blend[e] == blend) ++e;
      else while (e < r1 && !live(e)) ++e;
      split_run(e - r, K, pieces);
      int s = r;
      for (int len : pieces) {
        Item it{s - r0, len, blend, {-1, -1, -1, -1}};
        if (blend >= 0 && in.blend_offsets != nullptr) {
          const int beg = in.blend_offsets[blend], cnt = in.blend_offsets[blend + 1] - beg;
          for (int k = 0; k < std::min(cnt, 4); ++k) it.joint[k] = in.blend_transform[beg + k];
        }
        items.push_back(it);
        s += len;
      }
      r = e;
    }
    p.num_items += (long long)items.size();
    // Half-warp by half-warp, longest items first.  A half-warp lasts c steps, c = the longest item still unplaced.
    // Its sixteen lanes are filled best fit first: for waste w = 0, 1, 2, ... every free lane l looks for an unplaced
    // item of exactly c - w rows whose first row is congruent to l + d for some delay d <= w (so the item ends within
    // the c steps); among several, the one whose joints collide least with the quarter-warp's, then the least delay.
    for (auto &bk : bucket) bk.clear();
    for (size_t i = 0; i < items.size(); ++i) bucket[items[i].start & 15].push_back((int)i);
    for (auto &bk : bucket) std::stable_sort(bk.begin(), bk.end(), [&](int x, int y) { return items[x].count > items[y].count; });
    size_t remaining = items.size();
    slots.clear();
    int half = 0;  // half-warps filled so far
    while (remaining > 0) {
      if ((half & 1) == 0) {
        slots.emplace_back();
        Slot &ns = slots.back();
        for (int l = 0; l < 32; ++l) ns.lane_item[l] = -1, ns.lane_delay[l] = 0;
        for (auto &a : ns.nseen) for (auto &b : a) for (int &c : b) c = 0;
      }
      Slot &sl = slots.back();
      const int lane0 = (half & 1) * 16;
      int c = sl.steps;  // the second half-warp of a slot lasts at least as long as the first
      if ((half & 1) == 0 && trim_below > 0) {
        // A new slot lasts as long as the longest item left.  If only a few items are that long, the slot's other lanes
        // idle for their sake: cut their last row off (it becomes an item of its own, which finds an idle lane in one of
        // the short slots at the end of the tile) until the longest class is worth a slot.
        for (;;) {
          int longest = 0, n_longest = 0;
          for (auto &bk : bucket)
            if (!bk.empty()) {
              const int top = items[bk.front()].count;
              if (top > longest) longest = top, n_longest = 0;
              if (top == longest)
                for (size_t q = 0; q < bk.size() && items[bk[q]].count == longest; ++q) ++n_longest;
            }
          if (longest <= 2 || n_longest >= trim_below) break;
          for (int r16 = 0; r16 < 16; ++r16) {
            std::vector<int> &bk = bucket[r16];
            size_t q = 0;
            while (q < bk.size() && items[bk[q]].count == longest) {
              Item &it = items[bk[q]];
              --it.count;
              Item tail = it;
              tail.start = it.start + it.count;
              tail.count = 1;
              items.push_back(tail);
              const int ti = (int)items.size() - 1;
              std::vector<int> &tb = bucket[tail.start & 15];
              tb.push_back(ti);  // one-row items sort last
              ++remaining;
              ++p.num_items;
              ++q;
            }
            // the trimmed items now tie with the next class: the bucket stays sorted (descending, stable)
          }
        }
      }
      for (auto &bk : bucket) if (!bk.empty()) c = std::max(c, items[bk.front()].count);
      int free_lanes = 16;
      for (int w = 0; w < c && free_lanes > 0 && remaining > 0; ++w) {
        const int want = c - w;
        for (int l = 0; l < 16 && remaining > 0; ++l) {
          if (sl.lane_item[lane0 + l] >= 0) continue;
          int best_pos = -1, best_bucket = -1, best_d = 0, best_extra = 1 << 30;
          for (int d = 0; d <= std::min(w, 15); ++d) {
            std::vector<int> &bk = bucket[(l + d) & 15];
            // bucket sorted by count, descending: the stretch of items with exactly `want` rows
            size_t lo = 0;
            while (lo < bk.size() && items[bk[lo]].count > want) ++lo;
            for (size_t q = lo; q < bk.size() && items[bk[q]].count == want && q < lo + 48; ++q) {
              const int extra = gather_extra(sl, (lane0 + l) >> 3, items[bk[q]]);
              if (extra < best_extra) {
                best_extra = extra;
                best_pos = (int)q;
                best_bucket = (l + d) & 15;
                best_d = d;
              }
            }
            if (best_extra == 0) break;
          }
          if (best_pos < 0) continue;
          std::vector<int> &bk = bucket[best_bucket];
          place(sl, lane0 + l, bk[best_pos], best_d);
          bk.erase(bk.begin() + best_pos);
          --remaining;
          --free_lanes;
        }
      }
      // Lanes still free: no unplaced item ends within c steps from them.  Letting the slot grow by g steps admits
      // items with delay + rows = c + g; every step costs the whole slot kPlanStepCost, every item placed saves its share
      // of some later slot.  Tentatively fill for g = 1, 2, ..., keep the prefix with the best balance.
      if (free_lanes > 0 && remaining > 0) {
        struct Tentative {
          int lane, bucket, item, d, g;
        };
        std::vector<Tentative> tent;
        std::vector<char> taken_lane(16, 0);
        std::vector<std::vector<char>> taken(16);
        for (int r16 = 0; r16 < 16; ++r16) taken[r16].assign(bucket[r16].size(), 0);
        // rows per cost of the half-warp (half a slot's opening cost + its steps): keep the growth that maximises it
        long long rows = 0;
        for (int l = 0; l < 16; ++l)
          if (sl.lane_item[lane0 + l] >= 0) rows += items[sl.lane_item[lane0 + l]].count;
        auto cost_of = [&](int steps) { return (double)(kPlanSlotCost / 2 + steps * kPlanStepCost / 2); };
        double best_ratio = rows > 0 ? (double)rows / cost_of(c) : 0.0;
        size_t best_prefix = 0;
        for (int g = 1; g <= 15; ++g) {
          for (int l = 0; l < 16; ++l) {
            if (sl.lane_item[lane0 + l] >= 0 || taken_lane[l]) continue;
            // the longest unplaced item with delay + rows == c + g
            int bi = -1, bb = -1, bd = 0;
            for (int d = 0; d <= 15; ++d) {
              const int want = c + g - d;
              if (want < 1) break;
              const std::vector<int> &bk = bucket[(l + d) & 15];
              for (size_t q = 0; q < bk.size(); ++q) {
                if (items[bk[q]].count > want) continue;
                if (items[bk[q]].count < want) break;
                if (!taken[(l + d) & 15][q]) {
                  bi = (int)q, bb = (l + d) & 15, bd = d;
                  break;
                }
              }
              if (bi >= 0) break;
            }
            if (bi < 0) continue;
            taken[bb][bi] = 1;
            taken_lane[l] = 1;
            tent.push_back(Tentative{l, bb, bucket[bb][bi], bd, g});
            rows += items[bucket[bb][bi]].count;
          }
          const double ratio = (double)rows / cost_of(c + g);
          if (ratio > best_ratio * 1.0001) {
            best_ratio = ratio;
            best_prefix = tent.size();
          }
        }
        for (size_t q = 0; q < best_prefix; ++q) {
          place(sl, lane0 + tent[q].lane, tent[q].item, tent[q].d);
          std::vector<int> &bk = bucket[tent[q].bucket];
          bk.erase(std::find(bk.begin(), bk.end(), tent[q].item));
          --remaining;
        }
      }
      ++half;
    }
    // Clean-up: the greedy fill leaves the tile's last slot(s) nearly empty (two or three one-row items, say), and a slot
    // costs its opening (a blend per lane) whatever it holds.  Try to dissolve the emptiest slot: every item of it moves
    // into a free lane of another slot -- any free lane will do, the delay (first row - lane) mod 16 keeps the residue rule,
    // at the price of the steps the host slot grows by -- and the move is kept when the growth costs less than the slot saved.
    while (slots.size() > 1) {
      size_t victim = 0;
      for (size_t i = 1; i < slots.size(); ++i)
        if (slots[i].used < slots[victim].used || (slots[i].used == slots[victim].used && slots[i].steps <= slots[victim].steps)) victim = i;
      const Slot &vs = slots[victim];
      if (vs.used > 24) break;
      struct Move {
        int item, slot, lane, delay;
      };
      std::vector<Move> moves;
      std::vector<int> new_steps(slots.size());
      std::vector<std::vector<char>> taken(slots.size(), std::vector<char>(32, 0));
      for (size_t i = 0; i < slots.size(); ++i) new_steps[i] = slots[i].steps;
      bool ok = true;
      std::vector<int> order;  // longest items first: they have the fewest places to go
      for (int l = 0; l < 32; ++l)
        if (vs.lane_item[l] >= 0) order.push_back(vs.lane_item[l]);
      std::sort(order.begin(), order.end(), [&](int x, int y) { return items[x].count > items[y].count; });
      for (int idx : order) {
        const Item &it = items[idx];
        int best_slot = -1, best_lane = -1, best_d = 0, best_growth = 1 << 30, best_free = -1;
        for (size_t i = 0; i < slots.size(); ++i) {
          if (i == victim) continue;
          for (int l = 0; l < 32; ++l) {
            if (slots[i].lane_item[l] >= 0 || taken[i][l]) continue;
            const int d = (it.start - l) & 15;
            const int growth = std::max(0, d + it.count - new_steps[i]);
            const int free_here = 32 - slots[i].used;
            if (growth < best_growth || (growth == best_growth && free_here > best_free)) {
              best_growth = growth, best_slot = (int)i, best_lane = l, best_d = d, best_free = free_here;
            }
          }
        }
        if (best_slot < 0) {
          ok = false;
          break;
        }
        taken[best_slot][best_lane] = 1;
        new_steps[best_slot] = std::max(new_steps[best_slot], best_d + it.count);
        moves.push_back(Move{idx, best_slot, best_lane, best_d});
      }
      if (!ok) break;
      long long growth_cost = 0;
      for (size_t i = 0; i < slots.size(); ++i) growth_cost += (long long)(new_steps[i] - slots[i].steps) * kPlanStepCost;
      if (growth_cost >= kPlanSlotCost + (long long)vs.steps * kPlanStepCost) break;
      for (const Move &mv : moves) place(slots[mv.slot], mv.lane, mv.item, mv.delay);
      slots.erase(slots.begin() + victim);
    }
    // wavefronts of the quarter-warp's 4 x 3 gather loads, in units of 3; with two palette copies every distinct joint
    // takes the copy that serves the quarter-warp best (two_copy_gather); `copy_bits`: per lane, bit k = entry k from copy B
    auto quarter_cost = [&](const Slot &sl, int quarter, uint8_t *copy_bits = nullptr) {
      int total = 0;
      for (int k = 0; k < 4; ++k) {
        int distinct[8], nd = 0;
        for (int l = quarter * 8; l < quarter * 8 + 8; ++l) {
          if (sl.lane_item[l] < 0) continue;
          const int j = items[sl.lane_item[l]].joint[k];
          if (j < 0) continue;
          bool known = false;
          for (int q = 0; q < nd; ++q) known = known || distinct[q] == j;
          if (!known) distinct[nd++] = j;
        }
        int choice[8];
        total += two_copy_gather(distinct, nd, in.copy_b_base, choice);
        if (copy_bits != nullptr)
          for (int l = quarter * 8; l < quarter * 8 + 8; ++l) {
            if (sl.lane_item[l] < 0) continue;
            const int j = items[sl.lane_item[l]].joint[k];
            for (int q = 0; q < nd; ++q)
              if (distinct[q] == j && choice[q]) copy_bits[l] |= (uint8_t)(1u << k);
          }
      }
      return total;
    };
    // Palette-gather conflicts, second pass.  Lanes l and l + 16 of a slot serve the same row residue, so their items may
    // trade places at no cost to the row walk; that changes which quarter-warp (eight lanes = one 128-bit load wavefront)
    // an item gathers with.  Hill-climb per slot on the number of gather wavefronts.
    {
      for (Slot &sl : slots) {
        for (int pass = 0; pass < 4; ++pass) {
          bool improved = false;
          for (int l = 0; l < 16; ++l) {
            if (sl.lane_item[l] < 0 && sl.lane_item[l + 16] < 0) continue;
            const int qa = l >> 3, qb = (l + 16) >> 3;
            const int before = quarter_cost(sl, qa) + quarter_cost(sl, qb);
            std::swap(sl.lane_item[l], sl.lane_item[l + 16]);
            std::swap(sl.lane_delay[l], sl.lane_delay[l + 16]);
            if (quarter_cost(sl, qa) + quarter_cost(sl, qb) < before) {
              improved = true;
            } else {
              std::swap(sl.lane_item[l], sl.lane_item[l + 16]);
              std::swap(sl.lane_delay[l], sl.lane_delay[l + 16]);
            }
          }
          if (!improved) break;
        }
        // the statistics below read the slot's `nseen` tables: rebuild them for the final arrangement
        for (auto &a : sl.nseen) for (auto &b : a) for (int &c : b) c = 0;
        for (int l = 0; l < 32; ++l) {
          if (sl.lane_item[l] < 0) continue;
          const Item &it = items[sl.lane_item[l]];
          const int quarter = l >> 3;
          for (int k = 0; k < 4 && it.joint[k] >= 0; ++k) {
            const int j = it.joint[k], res = j & 7;
            bool known = false;
            for (int q = 0; q < sl.nseen[quarter][k][res]; ++q) known = known || sl.seen[quarter][k][res][q] == j;
            if (!known && sl.nseen[quarter][k][res] < 8) sl.seen[quarter][k][res][sl.nseen[quarter][k][res]++] = j;
          }
        }
      }
    }
    // longest slot first (the kernel deals a tile's slots to its warps in order)
    std::stable_sort(slots.begin(), slots.end(), [](const Slot &x, const Slot &y) { return x.steps > y.steps; });
    const size_t base = p.items.size();
    p.items.resize(base + slots.size() * 32, 0u);
    p.item_blend.resize(base + slots.size() * 32, -1);
    p.item_copy.resize(base + slots.size() * 32, 0);
    for (size_t si = 0; si < slots.size(); ++si) {
      const Slot &sl = slots[si];
      for (int q = 0; q < 4; ++q) p.gather_wavefronts += 3 * quarter_cost(sl, q, p.item_copy.data() + base + si * 32);
      for (int l = 0; l < 32; ++l) {
        const size_t at = base + si * 32 + l;
        if (sl.lane_item[l] < 0) {
          ++p.idle_lanes;
          continue;
        }
        const Item &it = items[sl.lane_item[l]];
        p.items[at] = (uint32_t)it.start | ((uint32_t)it.count << kItemStartBits) | (it.blend >= 0 ? kItemSkinned : 0u) |
                      ((uint32_t)sl.lane_delay[l] << kItemDelayShift);
        p.item_blend[at] = it.blend;
        p.lane_steps += it.count;
      }
      p.warp_steps += sl.steps;
      for (int q = 0; q < 4; ++q)
        for (int k = 0; k < 4; ++k) {
          int worst = 0, any = 0;
          for (int r8 = 0; r8 < 8; ++r8) worst = std::max(worst, sl.nseen[q][k][r8]), any += sl.nseen[q][k][r8];
          (void)worst;
          p.gather_ideal += any ? 3 : 0;
        }
    }
    p.num_slots += (long long)slots.size();
    p.tile_slot_offsets.push_back((int)(p.items.size() / 32));
  }
  return p;
}

inline void print_run_plan_stats(const RunPlanHost &p, int num_rows, FILE *f) {
  fprintf(f,
          "run plan: %d rows, tiles of %d rows (%d tiles), items of <= %d rows: %lld items (%.2f rows each), %lld slots (%.2f per tile), "
          "idle lanes %.1f %%, lockstep efficiency %.1f %% (%lld lane steps / %lld warp steps x 32), gather wavefronts %lld (conflict-free: %lld)\n",
          num_rows, p.tile_rows, p.num_tiles, p.max_item_rows, p.num_items, p.num_items ? (double)p.lane_steps / p.num_items : 0.0, p.num_slots,
          p.num_tiles ? (double)p.num_slots / p.num_tiles : 0.0, p.num_slots ? 100.0 * p.idle_lanes / (p.num_slots * 32.0) : 0.0,
          p.warp_steps ? 100.0 * p.lane_steps / (p.warp_steps * 32.0) : 0.0, p.lane_steps, p.warp_steps, p.gather_wavefronts, p.gather_ideal);
}

}  // namespace p3cs
