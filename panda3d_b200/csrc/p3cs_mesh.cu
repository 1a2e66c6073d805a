// p3cs_mesh.cu -- p3cs_mesh_create: the one-time flattening / upload of a GeomVertexData's static half
// (replaces the per-frame copy_from, geomVertexData.cxx:1460) and the tables the strip kernel runs on:
// dynamic columns, the shared-memory plan, record groups with their conflict-aware record order, the
// per-row morph CSR.  Host C++.
#include "p3cs_internal.h"
#include "p3cs_packer.cuh"
#include "p3cs_run.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

using namespace p3cs;

// ------------------------------------------------------------------ mesh
static int check_column(const p3cs_mesh_desc *d, const p3cs_column_desc &c, const char *what) {
  if (c.array < 0 || c.array >= d->num_arrays) return p3cs_fail(P3CS_ERR_INVALID, "%s: array %d out of range", what, c.array);
  const int stride = d->arrays[c.array].stride;
  if (c.start < 0 || c.num_values < 1 || c.start + column_bytes(c.numeric_type, c.num_values) > stride)
    return p3cs_fail(P3CS_ERR_INVALID, "%s: column [%d,+%d bytes) of %d values exceeds stride %d", what, c.start,
                column_bytes(c.numeric_type, c.num_values), c.num_values, stride);
  return P3CS_OK;
}

// Is this a column the reference itself can animate, and one the kernels address?  float32 / float64 columns are
// handled as 32-bit words (4-byte aligned starts); the integer and packed ones byte by byte, anywhere.
static int check_dynamic_column(const p3cs_column_desc &c, const char *what, int index) {
  if (!column_supported(c.contents, c.numeric_type, c.num_values))
    return p3cs_fail(P3CS_ERR_UNSUPPORTED, "%s %d: numeric type %d x %d values with contents %d is a combination the reference's "
                     "GeomVertexColumn packers do not read or write either", what, index, c.numeric_type, c.num_values, c.contents);
  const bool is_float = c.numeric_type == P3CS_NT_FLOAT32 || c.numeric_type == P3CS_NT_FLOAT64;
  if (is_float && c.start % 4 != 0)
    return p3cs_fail(P3CS_ERR_UNSUPPORTED, "%s %d: float column at byte %d is not 4-byte aligned", what, index, c.start);
  return P3CS_OK;
}
static uint8_t dynamic_column_code(const p3cs_column_desc &c) {
  if (c.numeric_type == P3CS_NT_FLOAT32 || c.numeric_type == P3CS_NT_FLOAT64) return 0;
  return (uint8_t)typed_code(c.numeric_type, packer_kind(c.contents, c.numeric_type, c.num_values));
}

// GeomVertexReader::get_data4 on a morph delta column of any type, on the host (once, at upload): the values in the
// packer's order as floats, padded with 0; colour types scaled (p3cs_packer.cuh has the device twin and the citations).
static void host_get4(const p3cs_column_desc &c, const uint8_t *p, float v[4]) {
  const int n = c.num_values;
  auto ufloat = [](uint32_t field, int mbits) {
    const uint32_t exponent = field >> mbits;
    if (exponent == 0) return ldexpf((float)(field & ((1u << mbits) - 1u)) / (float)(1u << mbits), -14);
    uint32_t bits = field << (23 - mbits);
    bits = exponent == 31 ? (bits | 0x7f800000u) : bits + 0x38000000u;
    float f;
    memcpy(&f, &bits, 4);
    return f;
  };
  v[0] = v[1] = v[2] = v[3] = 0.0f;
  uint32_t w = 0;
  switch (c.numeric_type) {
  case P3CS_NT_UINT8: for (int k = 0; k < n; ++k) v[k] = (float)p[k]; break;
  case P3CS_NT_INT8: for (int k = 0; k < n; ++k) v[k] = (float)(int8_t)p[k]; break;
  case P3CS_NT_UINT16: for (int k = 0; k < n; ++k) { uint16_t x; memcpy(&x, p + 2 * k, 2); v[k] = (float)x; } break;
  case P3CS_NT_INT16: for (int k = 0; k < n; ++k) { int16_t x; memcpy(&x, p + 2 * k, 2); v[k] = (float)x; } break;
  case P3CS_NT_UINT32: for (int k = 0; k < n; ++k) { uint32_t x; memcpy(&x, p + 4 * k, 4); v[k] = (float)x; } break;
  case P3CS_NT_INT32: for (int k = 0; k < n; ++k) { int32_t x; memcpy(&x, p + 4 * k, 4); v[k] = (float)x; } break;
  case P3CS_NT_FLOAT32: memcpy(v, p, (size_t)n * 4); break;
  case P3CS_NT_FLOAT64: for (int k = 0; k < n; ++k) { double x; memcpy(&x, p + 8 * k, 8); v[k] = (float)x; } break;
  case P3CS_NT_PACKED_DCBA: v[0] = (float)p[0]; v[1] = (float)p[1]; v[2] = (float)p[2]; v[3] = (float)p[3]; break;
  case P3CS_NT_PACKED_DABC: v[0] = (float)p[2]; v[1] = (float)p[1]; v[2] = (float)p[0]; v[3] = (float)p[3]; break;
  case P3CS_NT_PACKED_UFLOAT:
    memcpy(&w, p, 4);
    v[0] = ufloat(w & 0x7ffu, 6); v[1] = ufloat((w >> 11) & 0x7ffu, 6); v[2] = ufloat(w >> 22, 5);
    break;
  default: break;
  }
  if (packer_kind(c.contents, c.numeric_type, c.num_values) != kPackPlain) {
    const float scale = color_scale(c.numeric_type);
    if (scale != 0.0f) {
      const float recip = 1.0f / scale;
      for (int k = 0; k < n; ++k) v[k] = v[k] * recip;
    }
    if (n == 3) v[3] = 1.0f;
  }
}

static int check_ranges(const int32_t *ranges, int count, int num_rows, const char *what) {
  int prev_end = 0;
  for (int i = 0; i < count; ++i) {
    const int b = ranges[2 * i], e = ranges[2 * i + 1];
    if (b < prev_end || e <= b || e > num_rows)
      return p3cs_fail(P3CS_ERR_INVALID, "%s: subrange %d = [%d,%d) is not a sorted, disjoint, non-empty range within %d rows", what, i, b, e, num_rows);
    prev_end = e;
  }
  return P3CS_OK;
}

// ------------------------------------------------------------------ record groups of the strip kernel
// Order of the records inside a group.  Phase A runs one thread per record and gathers the record's
// first four palette matrices with 128-bit shared-memory loads, eight lanes per wavefront; two lanes
// collide when their joints differ but agree modulo 8 (48-byte palette rows).  The order of a group's
// records is free (rows address them through their slot), so they are packed first-fit into
// eight-lane cells whose k-th joints are pairwise distinct modulo 8 (or equal: a broadcast).
// `windows`: for every half-warp of rows (16 consecutive rows of the group) the positions in `blends` of the records
// those rows select -- the records one phase-B record read touches together.
static void pack_group(const p3cs_mesh_desc *d, std::vector<int> &blends, long long *budget, const std::vector<std::vector<int>> &windows) {
  const int n = (int)blends.size();
  const int cells = (n + 7) / 8;
  std::vector<int> occ((size_t)cells * 32, -1), fill(cells, 0), cell_of(n, -1);
  std::vector<int> leftover;
  int first_open = 0;
  for (int i = 0; i < n; ++i) {
    const int b = blends[i];
    const int e0 = d->blend_offsets[b], cnt = std::min(4, d->blend_offsets[b + 1] - e0);
    int placed = -1;
    for (int c = first_open; c < cells && placed < 0; ++c) {
      if (fill[c] >= 8) continue;
      bool ok = true;
      for (int k = 0; k < cnt && ok; ++k) {
        const int j = d->blend_transform[e0 + k];
        const int o = occ[(size_t)c * 32 + k * 8 + (j & 7)];
        ok = o < 0 || o == j;
      }
      if (ok) placed = c;
    }
    if (placed < 0) {
      leftover.push_back(i);
      continue;
    }
    for (int k = 0; k < cnt; ++k) {
      const int j = d->blend_transform[e0 + k];
      occ[(size_t)placed * 32 + k * 8 + (j & 7)] = j;
    }
    cell_of[i] = placed;
    ++fill[placed];
    while (first_open < cells && fill[first_open] >= 8) ++first_open;
  }
  int c = 0;
  for (int i : leftover) {  // whatever did not fit conflict-free fills the remaining lanes
    while (fill[c] >= 8) ++c;
    cell_of[i] = c;
    ++fill[c];
  }
  // dense positions: cells in order; only the last cell may be short, so move its surplus forward
  std::vector<int> count_in(cells, 0);
  for (int i = 0; i < n; ++i) ++count_in[cell_of[i]];
  for (int cc = 0, donor = cells - 1; cc < donor; ++cc)
    while (count_in[cc] < 8 && cc < donor) {
      if (count_in[donor] == 0) { --donor; continue; }
      for (int i = n - 1; i >= 0; --i)
        if (cell_of[i] == donor) { cell_of[i] = cc; break; }
      --count_in[donor];
      ++count_in[cc];
    }
  // Refinement: records of cells that still collide are swapped with records of other cells whenever that lowers
  // the two cells' wavefront counts (a cell costs, per entry k, the largest number of distinct joints sharing
  // one residue).  Bounded by `*budget` swap evaluations per mesh, so upload time stays in check.
  if (budget != nullptr && *budget > 0) {
    auto joint = [&](int rec, int k) {
      const int b = blends[rec], e0 = d->blend_offsets[b];
      return k < std::min(4, d->blend_offsets[b + 1] - e0) ? d->blend_transform[e0 + k] : -1;
    };
    std::vector<std::vector<int>> members(cells);
    for (int i = 0; i < n; ++i) members[cell_of[i]].push_back(i);
    // cost of a cell's member list with member `skip` left out and record `extra` added (-1: none)
    auto cell_cost = [&](const std::vector<int> &mem, int skip, int extra) {
      int total = 0;
      for (int k = 0; k < 4; ++k) {
        int distinct[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seen[9], ns = 0, worst = 0;
        auto add = [&](int rec) {
          const int j = joint(rec, k);
          if (j < 0) return;
          for (int q = 0; q < ns; ++q)
            if (seen[q] == j) return;
          seen[ns++] = j;
          worst = std::max(worst, ++distinct[j & 7]);
        };
        for (int rec : mem)
          if (rec != skip) add(rec);
        if (extra >= 0) add(extra);
        total += worst;
      }
      return total;
    };
    std::vector<int> cost(cells);
    for (int c = 0; c < cells; ++c) cost[c] = cell_cost(members[c], -1, -1);
    for (int round = 0; round < 3 && *budget > 0; ++round) {
      bool improved = false;
      for (int ca = 0; ca < cells && *budget > 0; ++ca) {
        int floor_a = 0;  // a conflict-free cell costs one wavefront per entry index in use
        for (int k = 0; k < 4; ++k)
          for (int rec : members[ca])
            if (joint(rec, k) >= 0) { ++floor_a; break; }
        if (cost[ca] <= floor_a) continue;
        for (size_t ia = 0; ia < members[ca].size() && *budget > 0; ++ia) {
          int best_gain = 0, best_c = -1;
          size_t best_i = 0;
          for (int cb = 0; cb < cells; ++cb) {
            if (cb == ca) continue;
            for (size_t ib = 0; ib < members[cb].size(); ++ib) {
              const int ra = members[ca][ia], rb = members[cb][ib];
              const int gain = cost[ca] + cost[cb] - cell_cost(members[ca], ra, rb) - cell_cost(members[cb], rb, ra);
              --*budget;
              if (gain > best_gain) {
                best_gain = gain;
                best_c = cb;
                best_i = ib;
              }
            }
          }
          if (best_c >= 0) {
            std::swap(members[ca][ia], members[best_c][best_i]);
            cost[ca] = cell_cost(members[ca], -1, -1);
            cost[best_c] = cell_cost(members[best_c], -1, -1);
            improved = true;
          }
        }
      }
      if (!improved) break;
    }
    for (int c = 0; c < cells; ++c)
      for (int rec : members[c]) cell_of[rec] = c;
  }
  // Lane inside the cell.  Phase B reads a row's record with 64-bit loads, sixteen rows per wavefront; two records read
  // by the same half-warp collide there when their slots agree modulo 16 (record stride = an odd number of 8-byte
  // units).  The records that meet in a half-warp are known (`windows`), so this is a colouring of that graph with the
  // 16 residues, greedy in first-seen order: each record takes the free lane of its cell whose slot collides with the
  // fewest already placed neighbours.
  std::vector<std::vector<int>> nbr(n);
  for (const std::vector<int> &w : windows)
    for (int a : w)
      for (int b : w)
        if (a != b) nbr[a].push_back(b);
  std::vector<int> slot_of(n, -1);
  std::vector<uint8_t> used((size_t)cells, 0);
  for (int i = 0; i < n; ++i) {
    const int c = cell_of[i];
    int best = -1, best_cost = 1 << 30;
    for (int p = 0; p < count_in[c]; ++p) {
      if ((used[c] >> p) & 1) continue;
      const int sl = 8 * c + p;
      int cost = 0;
      for (int o : nbr[i])
        if (slot_of[o] >= 0 && ((slot_of[o] ^ sl) & 15) == 0) ++cost;
      if (cost < best_cost) {
        best_cost = cost;
        best = p;
        if (cost == 0) break;
      }
    }
    used[c] |= (uint8_t)(1u << best);
    slot_of[i] = 8 * c + best;
  }
  // ... then improved by swapping the lanes of two records of one cell while that lowers the number of collisions
  {
    auto cost_at = [&](int i, int sl, int ignore) {
      int cost = 0;
      for (int o : nbr[i])
        if (o != ignore && ((slot_of[o] ^ sl) & 15) == 0) ++cost;
      return cost;
    };
    std::vector<std::vector<int>> members(cells);
    for (int i = 0; i < n; ++i) members[cell_of[i]].push_back(i);
    for (int pass = 0; pass < 4; ++pass) {
      bool improved = false;
      for (int i = 0; i < n; ++i) {
        const int ci = cost_at(i, slot_of[i], -1);
        if (ci == 0) continue;
        int best_k = -1, best_delta = 0;
        for (int k : members[cell_of[i]]) {
          if (k == i) continue;
          const int before = ci + cost_at(k, slot_of[k], -1);
          const int after = cost_at(i, slot_of[k], k) + cost_at(k, slot_of[i], i);
          if (after - before < best_delta) {
            best_delta = after - before;
            best_k = k;
          }
        }
        if (best_k >= 0) {
          std::swap(slot_of[i], slot_of[best_k]);
          improved = true;
        }
      }
      if (!improved) break;
    }
  }
  if (getenv("P3CS_DEBUG_PACK") != nullptr) {  // predicted extra wavefronts of the record reads, by window
    long long extra = 0;
    for (const std::vector<int> &w : windows) {
      int cnt[16] = {0};
      int worst = 0;
      for (int a : w) worst = std::max(worst, ++cnt[slot_of[a] & 15]);
      extra += worst - 1;
    }
    fprintf(stderr, "pack_group: %d records, %zu windows, predicted extra wavefronts per record load %lld\n", n, windows.size(), extra);
  }
  std::vector<int> ordered(n);
  for (int i = 0; i < n; ++i) ordered[slot_of[i]] = blends[i];
  blends.swap(ordered);
}

// ---- strip-kernel tables: per record group the distinct blends, per row its slot in that list.
// Groups are grown greedily, tile by tile, while their distinct blends fit one phase-A pass
// (one thread per record) and the shared-memory record budget.
struct StripTables {
  std::vector<uint16_t> local;                             // per row: slot in its group's record list (0xFFFF: not skinned)
  std::vector<int> rec_off, rec_blend, ent_off, tile_off;  // group -> records, record -> blend / entries, group -> tiles
  std::vector<uint2> ent;                                  // (palette index, weight bits) of every record, stored order
  int num_groups = 0, max_records = 0;
};

// `row_index`: the transform_blend value of every row; `mask`: bit per row inside TransformBlendTable::get_rows()
// (empty = all rows); `cap`: most records a group may hold.
static StripTables build_record_groups(const p3cs_mesh_desc *d, const std::vector<uint32_t> &row_index, const std::vector<uint32_t> &mask,
                                       int kTileRows, int cap) {
  const int N = d->num_rows;
  const int nt = (N + kTileRows - 1) / kTileRows;
  StripTables tb;
  std::vector<uint16_t> &local = tb.local;
  std::vector<int> &rec_off = tb.rec_off, &rec_blend = tb.rec_blend, &ent_off = tb.ent_off, &tile_off = tb.tile_off;
  std::vector<uint2> &ent = tb.ent;
  local.assign(std::max(N, 1), 0xFFFFu);
  rec_off.assign(1, 0);
  ent_off.assign(1, 0);
  tile_off.assign(1, 0);
  std::vector<int> in_group(std::max(d->num_blends, 1), -1), seen_in_tile(std::max(d->num_blends, 1), -1),
      slot(std::max(d->num_blends, 1), 0);
  int &ng = tb.num_groups, &max_rec = tb.max_records;
  std::vector<int> open;   // distinct blends of the open group, first-seen order
  std::vector<int> fresh;  // blends first seen in the tile being added
  long long pack_budget = 1500000;  // swap evaluations of pack_group's refinement, for the whole mesh (~0.4 s at most)
  auto row_live = [&](int r) { return mask.empty() || ((mask[r >> 5] >> (r & 31)) & 1u); };
  auto row_blend = [&](int r) { return (int)row_index[r]; };

  std::vector<int> pos_in_open(std::max(d->num_blends, 1), -1);
  auto close_group = [&](int t_end) {
    // the records each half-warp of rows reads together (distinct, as positions in `open`)
    std::vector<std::vector<int>> windows;
    {
      for (size_t p = 0; p < open.size(); ++p) pos_in_open[open[p]] = (int)p;
      const int w0 = tile_off.back() * kTileRows, w1 = std::min(N, t_end * kTileRows);
      for (int base = w0; base < w1; base += 16) {
        std::vector<int> w;
        for (int r = base; r < std::min(w1, base + 16); ++r) {
          if (!row_live(r)) continue;
          const int q = pos_in_open[row_blend(r)];
          if (std::find(w.begin(), w.end(), q) == w.end()) w.push_back(q);
        }
        if (w.size() > 1) windows.push_back(std::move(w));
      }
    }
    pack_group(d, open, &pack_budget, windows);
    const int n = (int)open.size();
    for (int p = 0; p < n; ++p) {
      const int b = open[p];
      slot[b] = p;
      rec_blend.push_back(b);
      for (int e = d->blend_offsets[b]; e < d->blend_offsets[b + 1]; ++e) {
        uint32_t wbits;
        memcpy(&wbits, &d->blend_weight[e], 4);
        ent.push_back(make_uint2((uint32_t)d->blend_transform[e], wbits));
      }
      ent_off.push_back((int)ent.size());
    }
    const int r0 = tile_off.back() * kTileRows, r1 = std::min(N, t_end * kTileRows);
    for (int r = r0; r < r1; ++r)
      if (row_live(r)) local[r] = (uint16_t)slot[row_blend(r)];
    rec_off.push_back(rec_off.back() + n);
    max_rec = std::max(max_rec, n);
    tile_off.push_back(t_end);
    ++ng;
    open.clear();
  };

  for (int t = 0; t < nt; ++t) {
    const int r0 = t * kTileRows, r1 = std::min(N, r0 + kTileRows);
    fresh.clear();  // distinct blends this tile would add to the open group
    for (int r = r0; r < r1; ++r) {
      if (!row_live(r)) continue;
      const int b = row_blend(r);
      if (in_group[b] != ng && seen_in_tile[b] != t) {
        seen_in_tile[b] = t;
        fresh.push_back(b);
      }
    }
    if (tile_off.back() != t && (int)open.size() + (int)fresh.size() > cap) {  // close the group before this tile
      close_group(t);
      for (int b : fresh) seen_in_tile[b] = -1;
      --t;  // redo this tile in the new group
      continue;
    }
    for (int b : fresh) {
      in_group[b] = ng;
      open.push_back(b);
    }
  }
  if (nt > 0) close_group(nt);
  return tb;
}

// ---- run-kernel plans: items of <= K consecutive rows selecting one blend, dealt to 32-lane slots (p3cs_runplan.h),
// plus per item its first four (palette index, weight) entries.  Two plans per mesh: [0] large tiles for crowds
// (as many rows as two tile buffers of a CTA hold with four CTAs per SM), [1] 256-row tiles for launches of a few
// instances, where many small CTAs fill the machine better.
static int upload_run_plan(p3cs_mesh mesh, const p3cs_mesh_desc *d, const RunPlanHost &hp, int tile_bufs, int copy_b_base, RunPlanDev *out) {
  int rc;
  const size_t used = hp.items.size();
  const size_t n = used + 8 * 32;  // padded by a CTA's worth of slots: the kernel prefetches the warp's next slot unconditionally
  std::vector<uint4> a(n, make_uint4(0, 0, 0, 0)), b(n, make_uint4(0, 0, 0, 0xFFFFFFFFu));
  for (size_t i = 0; i < used; ++i) {
    const int blend = hp.item_blend[i];
    uint32_t item = hp.items[i], joints[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
    if (blend >= 0) {
      const int beg = d->blend_offsets[blend], cnt = d->blend_offsets[blend + 1] - beg;
      for (int k = 0; k < std::min(cnt, 4); ++k) {
        const int j = d->blend_transform[beg + k];
        // the entry's palette-area index: copy A = the joint itself, copy B where the planner chose it
        joints[k] = (uint32_t)(copy_b_base > 0 && ((hp.item_copy[i] >> k) & 1u) ? palette_copy_b_index(j, copy_b_base) : j);
        memcpy(&w[k], &d->blend_weight[beg + k], 4);
      }
      item |= (uint32_t)std::min(cnt, 5) << kItemEntriesShift;
    }
    a[i] = make_uint4(item, joints[0] | (joints[1] << 16), joints[2] | (joints[3] << 16), w[0]);
    b[i] = make_uint4(w[1], w[2], w[3], (uint32_t)blend);
  }
  out->tile_rows = hp.tile_rows;
  out->tile_bufs = tile_bufs;
  out->num_tiles = hp.num_tiles;
  out->max_item_rows = hp.max_item_rows;
  out->copy_b_base = copy_b_base;
  out->pal_entries = copy_b_base > 0 ? 2 * copy_b_base : mesh->dev.num_transforms;
  if ((rc = upload<int>(mesh, hp.tile_slot_offsets.data(), hp.tile_slot_offsets.size(), &out->tile_slot_offsets)) != P3CS_OK) return rc;
  if ((rc = upload<uint4>(mesh, a.data(), n, &out->item_a)) != P3CS_OK) return rc;
  if ((rc = upload<uint4>(mesh, b.data(), n, &out->item_b)) != P3CS_OK) return rc;
  // the uploads read host vectors that die with this frame
  if (cudaStreamSynchronize(mesh->ctx->stream) != cudaSuccess) return p3cs_fail(P3CS_ERR_CUDA, "run plan upload failed");
  return P3CS_OK;
}

static int build_run_plans(p3cs_mesh mesh, const p3cs_mesh_desc *d, const std::vector<uint32_t> &row_index, const std::vector<uint32_t> &mask) {
  MeshDev &m = mesh->dev;
  m.run_mode = 0;
  if (m.num_rows == 0 || m.index_bytes == 0 || m.num_transforms >= 65536) return P3CS_OK;
  const int mode = strip_pick_mode(m);
  const int row_bytes = strip_mode_row_bytes(mode);
  if (row_bytes == 0) return P3CS_OK;
  if (d->num_sliders > 65536) return P3CS_OK;
  int threads, ctas;
  run_shape(mode, &threads, &ctas);
  // Two palette copies in shared memory (p3cs_runplan.h) where the palette gather weighs enough to pay for the tile rows
  // the second copy displaces.  Measured on B200, crowds of the C3 character (A/B on one box): 24-byte rows 1.251 ->
  // 1.206 ms, 32-byte rows 1.590 -> 1.569 ms, a random blend per vertex 3.31 -> 3.14 ms; 48-byte rows with runs lose
  // (2.01 -> 2.04 ms: their tiles are small already -- 672 rows -- and a row carries more bytes per gathered matrix).
  // So: 24- and 32-byte rows, or runs shorter than three rows on average; not for tiny palettes nor where both copies
  // exceed 16 KB.  P3CS_RUN_COPIES=1 / 2 forces (A/B runs).
  int copy_b_base = (m.num_transforms + 7) & ~7;  // copy B behind copy A, both whole octets of joints
  long long runs = 1;
  for (int r = 1; r < m.num_rows; ++r) runs += row_index[r] != row_index[r - 1];
  bool two_copies = m.num_transforms >= 16 && 2 * copy_b_base * 48 <= 16 * 1024 &&
                    (mode == kRow24 || mode == kRow32 || (long long)m.num_rows < 3 * runs);
  if (const char *e = getenv("P3CS_RUN_COPIES")) two_copies = atoi(e) == 2 && m.num_transforms >= 8 && 2 * copy_b_base < 65536;
  if (!two_copies) copy_b_base = 0;
  const int pal_entries = two_copies ? 2 * copy_b_base : m.num_transforms;
  const int fixed = run_layout(pal_entries, m.num_transforms, 16, row_bytes, 1, threads).tiles_off;
  // `ctas` CTAs per SM: 228 KB per SM, 1 KB reserved per CTA
  int budget = (228 * 1024 / ctas - 1024) - fixed;
  if (budget < 2 * 128 * row_bytes) budget = 113 * 1024 - fixed;   // very large palettes: two CTAs per SM
  if (budget < 2 * 64 * row_bytes) return P3CS_OK;                 // does not fit: the strip kernel's own fallbacks apply
  int tile_rows = std::min(4096, (budget / 2 / row_bytes) & ~15), bufs = 2, K = 12;
  if (const char *e = getenv("P3CS_RUN_PLAN")) {  // tuning aid: "tile_rows,tile_bufs,max_item_rows"
    int a = 0, b = 0, c = 0;
    if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a >= 16 && a <= 8176 && a % 16 == 0 && b >= 2 && b <= 4 && c >= 1 && c <= kItemMaxRows &&
        fixed + (size_t)b * (((size_t)a * row_bytes + 127) & ~(size_t)127) <= 200 * 1024) {
      tile_rows = a;
      bufs = b;
      K = c;
    }
  }
  const int trim = getenv("P3CS_RUN_TRIM") != nullptr ? atoi(getenv("P3CS_RUN_TRIM")) : 4;  // see build_run_plan
  RunPlanInput in;
  in.num_rows = m.num_rows;
  in.row_blend = row_index.data();
  in.live_mask = mask.empty() ? nullptr : mask.data();
  in.blend_offsets = d->blend_offsets;
  in.blend_transform = d->blend_transform;
  in.copy_b_base = copy_b_base;
  in.match_every_candidate = (long long)m.num_rows >= 3 * runs;  // see RunPlanInput
  int rc;
  {
    const RunPlanHost hp = build_run_plan(in, tile_rows, K, trim);
    if (getenv("P3CS_DEBUG_PLAN") != nullptr) print_run_plan_stats(hp, m.num_rows, stderr);
    m.run_lockstep_pct = hp.warp_steps > 0 ? (int)(100 * hp.lane_steps / (32 * hp.warp_steps)) : 0;
    if ((rc = upload_run_plan(mesh, d, hp, bufs, copy_b_base, &m.run_plan[0])) != P3CS_OK) return rc;
  }
  if (tile_rows > 256) {
    const int small_bufs = fixed + 3 * 256 * row_bytes <= 228 * 1024 / ctas - 1024 ? 3 : 2;
    const RunPlanHost hp = build_run_plan(in, 256, K, trim);
    if ((rc = upload_run_plan(mesh, d, hp, small_bufs, copy_b_base, &m.run_plan[1])) != P3CS_OK) return rc;
  }
  m.run_mode = mode;
  return P3CS_OK;
}

extern "C" int p3cs_mesh_create(p3cs_context ctx, const p3cs_mesh_desc *d, p3cs_mesh *out) {
  if (out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_mesh_create: out is NULL");
  *out = nullptr;
  if (ctx == nullptr || d == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_mesh_create: NULL argument");
  if (d->struct_size != sizeof(p3cs_mesh_desc))
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_mesh_desc::struct_size %u != %zu (header mismatch)", d->struct_size, sizeof(p3cs_mesh_desc));
  if (d->num_rows < 0 || d->num_arrays < 0 || (d->num_arrays > 0 && d->arrays == nullptr))
    return p3cs_fail(P3CS_ERR_INVALID, "bad num_rows / arrays");
  int cs = d->coordinate_system == 0 ? P3CS_CS_ZUP_RIGHT : d->coordinate_system;
  if (cs < P3CS_CS_ZUP_RIGHT || cs > P3CS_CS_YUP_LEFT) return p3cs_fail(P3CS_ERR_INVALID, "bad coordinate_system %d", cs);
  const int N = d->num_rows;
  DeviceGuard g(ctx->device);

  p3cs_mesh mesh = new (std::nothrow) p3cs_mesh_s;
  if (mesh == nullptr) return p3cs_fail(P3CS_ERR_NOMEM, "out of host memory");
  mesh->ctx = ctx;
  MeshDev &m = mesh->dev;
  m.num_rows = N;
  m.coordinate_system = cs;
  m.num_transforms = d->num_transforms;
  m.num_blends = d->num_blends;
  m.num_sliders = d->num_sliders;
  int rc = P3CS_OK;
  auto bail = [&](int code) {
    p3cs_mesh_destroy(mesh);
    return code;
  };

  // ---- output arrays: static image uploaded once (replaces the per-frame copy_from)
  std::vector<int> out_of_array(d->num_arrays, -1);
  for (int a = 0; a < d->num_arrays; ++a) {
    const p3cs_array_desc &ad = d->arrays[a];
    if (ad.stride <= 0 || (N > 0 && ad.data == nullptr)) return bail(p3cs_fail(P3CS_ERR_INVALID, "array %d: bad stride/data", a));
    if (!ad.is_output) continue;
    if (m.num_outputs >= kMaxOutputs) return bail(p3cs_fail(P3CS_ERR_UNSUPPORTED, "more than %d output arrays", kMaxOutputs));
    if (ad.stride % 4 != 0) return bail(p3cs_fail(P3CS_ERR_UNSUPPORTED, "output array %d: stride %d is not a multiple of 4", a, ad.stride));
    out_of_array[a] = m.num_outputs;
    m.stride[m.num_outputs] = ad.stride;
    if ((rc = upload<uint8_t>(mesh, static_cast<const uint8_t *>(ad.data), (size_t)N * ad.stride, &m.src[m.num_outputs])) != P3CS_OK) return bail(rc);
    ++m.num_outputs;
  }

  // ---- dynamic columns: skinned columns, then morph bases
  auto find_dyn = [&](int output, int start) {
    for (int c = 0; c < m.num_dyn; ++c)
      if (m.dyn[c].output == output && m.dyn[c].start == start) return c;
    return -1;
  };
  for (int i = 0; i < d->num_skin_columns; ++i) {
    const p3cs_column_desc &c = d->skin_columns[i];
    if ((rc = check_column(d, c, "skin column")) != P3CS_OK) return bail(rc);
    if (out_of_array[c.array] < 0) return bail(p3cs_fail(P3CS_ERR_INVALID, "skin column %d lives in a non-output array", i));
    int skin = c.contents == P3CS_C_POINT ? 1 : c.contents == P3CS_C_VECTOR ? 2 : c.contents == P3CS_C_NORMAL ? 3 : 0;
    if (skin == 0) return bail(p3cs_fail(P3CS_ERR_INVALID, "skin column %d: contents %d is not point/vector/normal", i, c.contents));
    if ((rc = check_dynamic_column(c, "skin column", i)) != P3CS_OK) return bail(rc);
    if (m.num_dyn >= kMaxDynColumns) return bail(p3cs_fail(P3CS_ERR_UNSUPPORTED, "more than %d animated columns", kMaxDynColumns));
    DynColumn &dc = m.dyn[m.num_dyn++];
    dc.output = (int8_t)out_of_array[c.array];
    dc.start = (int16_t)c.start;
    dc.num_values = (int8_t)c.num_values;
    dc.skin = (int8_t)skin;
    dc.homogeneous = 0;
    dc.f64 = c.numeric_type == P3CS_NT_FLOAT64 ? 1 : 0;
    dc.typed = dynamic_column_code(c);
    if (c.num_values == 4) m.full_records = 1;
  }

  // column order is semantically irrelevant (columns are independent); a canonical order
  // point(s), normal(s), vector(s) lets the strip kernel pick a specialised variant
  // ... and, inside a kind, by position (output array, byte offset): the row-layout modes (kRow48 / kRow56) address
  // "the lower vector column" and "the upper one" by their dyn index, whatever order the caller listed them in
  std::stable_sort(m.dyn, m.dyn + m.num_dyn, [](const DynColumn &a, const DynColumn &b) {
    auto key = [](int skin) { return skin == 1 ? 0 : skin == 3 ? 1 : 2; };
    if (key(a.skin) != key(b.skin)) return key(a.skin) < key(b.skin);
    if (a.output != b.output) return a.output < b.output;
    return a.start < b.start;
  });

  // ---- transform_blend index column -> dense u16/u32
  std::vector<uint32_t> plan_row_index, plan_mask;  // kept for the run-kernel plans built after the morph tables
  if (d->blend_column.array >= 0) {
    const p3cs_column_desc &bc = d->blend_column;
    if ((rc = check_column(d, bc, "transform_blend column")) != P3CS_OK) return bail(rc);
    if (bc.numeric_type != P3CS_NT_UINT8 && bc.numeric_type != P3CS_NT_UINT16 && bc.numeric_type != P3CS_NT_UINT32)
      return bail(p3cs_fail(P3CS_ERR_UNSUPPORTED, "transform_blend column of numeric type %d", bc.numeric_type));
    if (d->num_blends < 0 || d->num_transforms < 0 || d->blend_offsets == nullptr)
      return bail(p3cs_fail(P3CS_ERR_INVALID, "blend table missing"));
    if ((rc = check_ranges(d->row_ranges, d->num_row_ranges, N, "TransformBlendTable rows")) != P3CS_OK) return bail(rc);
    const int nnz = d->blend_offsets[d->num_blends];
    if (d->blend_offsets[0] != 0) return bail(p3cs_fail(P3CS_ERR_INVALID, "blend_offsets[0] != 0"));
    for (int b = 0; b < d->num_blends; ++b)
      if (d->blend_offsets[b + 1] < d->blend_offsets[b]) return bail(p3cs_fail(P3CS_ERR_INVALID, "blend_offsets not monotonic at %d", b));
    for (int e = 0; e < nnz; ++e)
      if (d->blend_transform[e] < 0 || d->blend_transform[e] >= d->num_transforms)
        return bail(p3cs_fail(P3CS_ERR_INVALID, "blend entry %d refers to transform %d outside the palette (%d)", e, d->blend_transform[e], d->num_transforms));

    const uint8_t *base = static_cast<const uint8_t *>(d->arrays[bc.array].data) + bc.start;
    const int istride = d->arrays[bc.array].stride;
    const bool wide = bc.numeric_type == P3CS_NT_UINT32;
    std::vector<uint16_t> idx16;
    std::vector<uint32_t> idx32;
    if (wide) idx32.resize(std::max(N, 1)); else idx16.resize(std::max(N, 1));
    for (int r = 0; r < N; ++r) {
      const uint8_t *p = base + (size_t)r * istride;
      uint32_t v;
      if (bc.numeric_type == P3CS_NT_UINT8) v = *p;
      else if (bc.numeric_type == P3CS_NT_UINT16) { uint16_t t; memcpy(&t, p, 2); v = t; }
      else memcpy(&v, p, 4);
      if (wide) idx32[r] = v; else idx16[r] = (uint16_t)v;
    }
    // every row the reference would transform must select an existing blend
    bool full = d->num_row_ranges == 1 && d->row_ranges[0] == 0 && d->row_ranges[1] == N;
    std::vector<uint32_t> mask;
    if (!full) mask.assign(((size_t)N + 31) / 32 + 1, 0u);
    for (int i = 0; i < d->num_row_ranges; ++i)
      for (int r = d->row_ranges[2 * i]; r < d->row_ranges[2 * i + 1]; ++r) {
        uint32_t v = wide ? idx32[r] : idx16[r];
        if (v >= (uint32_t)d->num_blends)
          return bail(p3cs_fail(P3CS_ERR_INVALID, "row %d selects blend %u but the table has %d blends", r, v, d->num_blends));
        if (!full) mask[r >> 5] |= 1u << (r & 31);
      }
    // ---- vertex-animation-align-16 layout (gobj/config_gobj.cxx:286-300, geomVertexArrayFormat.cxx:349-373): the
    // strip kernel's kAligned modes
    {
      const DynColumn *c = m.dyn;
      bool ok = m.full_records && m.num_outputs == 1 && m.stride[0] % 16 == 0 &&
                ((m.num_dyn == 2 && c[0].skin == 1 && c[1].skin == 3) ||
                 (m.num_dyn == 4 && c[0].skin == 1 && c[1].skin == 3 && c[2].skin == 2 && c[3].skin == 2));
      for (int i = 0; ok && i < m.num_dyn; ++i) ok = c[i].num_values == 4 && !c[i].f64 && !c[i].typed && c[i].start % 16 == 0;
      for (int mi = 0; ok && mi < d->num_morphs; ++mi) {  // morphs only on those columns (no morph-only column)
        const p3cs_column_desc &b = d->morphs[mi].base;
        ok = b.array >= 0 && b.array < d->num_arrays && out_of_array[b.array] == 0 && b.num_values == 4 &&
             b.numeric_type == P3CS_NT_FLOAT32 && find_dyn(0, b.start) >= 0;
      }
      m.aligned = ok ? 1 : 0;
    }
    m.strip_rec_floats = m.aligned ? kRecAligned : m.full_records ? kRecFull : kRecCompact;
    // ---- shared-memory plan of the strip kernel.  Four CTAs per SM (the register-file limit) need <= ~55 KB each:
    // palette + records + staged tiles.  Preference: two rows per thread (512-row tiles halve the per-tile
    // bookkeeping), if need be with a record group of at least 160 blends (measured on 32-byte rows: 1.90 -> 1.87 ms);
    // else 256-row tiles with three, then two buffers, then a smaller record group.
    {
      size_t slice = 0;
      for (int o = 0; o < m.num_outputs; ++o) slice += round_up((size_t)32 * m.stride[o], 128);
      const size_t rec_bytes = (size_t)m.strip_rec_floats * 4;
      const size_t fixed = 32 + (size_t)d->num_transforms * 64 + 128;
      const size_t budget = 55 * 1024;
      const int full_cap = std::min(kStripThreads, (int)((32 * 1024) / rec_bytes));  // one phase-A pass: a thread per record
      auto fits = [&](int r, int bufs, int cap) { return fixed + (size_t)cap * rec_bytes + (size_t)bufs * 8 * r * slice <= budget; };
      m.rows_per_thread = 1;
      m.tile_bufs = 3;
      m.group_record_cap = full_cap;
      if (!m.full_records || m.aligned) {
        const char *force_r = getenv("P3CS_ROWS_PER_THREAD");  // tuning aid: "1" keeps 256-row tiles
        if (fits(2, 2, full_cap) && !(force_r != nullptr && atoi(force_r) == 1)) {
          m.rows_per_thread = 2;
          m.tile_bufs = 2;
        } else if ([&] {  // 512-row tiles with a smaller record group, down to 160 records (stride 32: 168)
                     int cap = full_cap;
                     while (cap >= 160 && !fits(2, 2, cap)) cap -= 8;
                     if (cap < 160 || (force_r != nullptr && atoi(force_r) == 1)) return false;
                     m.rows_per_thread = 2;
                     m.tile_bufs = 2;
                     m.group_record_cap = cap;
                     return true;
                   }()) {
        } else if (fits(1, 3, full_cap)) {
        } else if (fits(1, 2, full_cap)) {
          m.tile_bufs = 2;
        } else {
          int cap = full_cap;
          while (cap > 128 && !fits(1, 2, cap)) cap -= 8;
          if (fits(1, 2, cap)) {
            m.tile_bufs = 2;
            m.group_record_cap = cap;
          }
        }
      }
    }
    if (const char *plan = getenv("P3CS_STRIP_PLAN")) {  // tuning aid: "rows_per_thread,tile_bufs,record_cap"
      int r = 0, b = 0, c = 0;
      if (sscanf(plan, "%d,%d,%d", &r, &b, &c) == 3 && (r == 1 || r == 2) && (b == 2 || b == 3) && c >= 8 && c <= kStripThreads &&
          (!m.full_records || m.aligned)) {
        m.rows_per_thread = r;
        m.tile_bufs = r == 2 ? 2 : b;
        m.group_record_cap = c;
      }
    }
    const int kTileRows = kStripThreads * m.rows_per_thread;
    // ---- strip-kernel tables (build_record_groups above)
    {
      std::vector<uint32_t> row_index(std::max(N, 1));
      for (int r = 0; r < N; ++r) row_index[r] = wide ? idx32[r] : idx16[r];
      const StripTables tb = build_record_groups(d, row_index, full ? std::vector<uint32_t>() : mask, kTileRows, m.group_record_cap);
      plan_row_index = row_index;
      if (!full) plan_mask = mask;
      const std::vector<uint16_t> &local = tb.local;
      const std::vector<int> &rec_off = tb.rec_off, &rec_blend = tb.rec_blend, &ent_off = tb.ent_off, &tile_off = tb.tile_off;
      const std::vector<uint2> &ent = tb.ent;
      const int ng = tb.num_groups, max_rec = tb.max_records;
      if ((rc = upload<int>(mesh, tile_off.data(), tile_off.size(), &m.group_tile_offsets)) != P3CS_OK) return bail(rc);
      m.num_groups = ng;
      m.max_group_records = max_rec;
      if ((rc = upload<uint16_t>(mesh, local.data(), local.size(), &m.row_local)) != P3CS_OK) return bail(rc);
      if (m.rows_per_thread == 2) {  // thread t of a 512-row tile animates rows t and t + 256: both slots in one word
        const int ntiles = (N + kTileRows - 1) / kTileRows;
        std::vector<uint32_t> pairs((size_t)std::max(ntiles, 1) * kStripThreads, 0xFFFFFFFFu);
        for (int t = 0; t < ntiles; ++t)
          for (int k = 0; k < kStripThreads; ++k) {
            const int r0 = t * kTileRows + k, r1 = r0 + kStripThreads;
            const uint32_t lo = r0 < N ? local[r0] : 0xFFFFu, hi = r1 < N ? local[r1] : 0xFFFFu;
            pairs[(size_t)t * kStripThreads + k] = lo | (hi << 16);
          }
        if ((rc = upload<uint32_t>(mesh, pairs.data(), pairs.size(), &m.row_local2)) != P3CS_OK) return bail(rc);
      }
      if ((rc = upload<int>(mesh, rec_off.data(), rec_off.size(), &m.group_rec_offsets)) != P3CS_OK) return bail(rc);
      if ((rc = upload<int>(mesh, rec_blend.data(), rec_blend.size(), &m.group_rec_blend)) != P3CS_OK) return bail(rc);
      if ((rc = upload<int>(mesh, ent_off.data(), ent_off.size(), &m.grec_entry_offsets)) != P3CS_OK) return bail(rc);
      if ((rc = upload<uint2>(mesh, ent.data(), ent.size(), &m.grec_entries)) != P3CS_OK) return bail(rc);
      // fixed-width blocks: the first four entries of every group record, count in bits 16..31
      const size_t nrec_total = rec_blend.size();
      // padded by one CTA's worth of records: the kernel prefetches block (group start + tid)
      const size_t half = nrec_total + kStripThreads;  // first halves of all records, then the second halves
      std::vector<uint4> blocks(half * 2, make_uint4(0, 0, 0, 0));
      for (size_t r = 0; r < nrec_total; ++r) {
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int beg = ent_off[r], cnt = ent_off[r + 1] - ent_off[r];
        for (int k = 0; k < std::min(cnt, 4); ++k) {
          w[2 * k] = ent[beg + k].x;
          w[2 * k + 1] = ent[beg + k].y;
        }
        w[0] |= (uint32_t)std::min(cnt, 0xFFFF) << 16;
        blocks[r] = make_uint4(w[0], w[1], w[2], w[3]);
        blocks[half + r] = make_uint4(w[4], w[5], w[6], w[7]);
      }
      if ((rc = upload<uint4>(mesh, blocks.data(), blocks.size(), &m.grec_blocks)) != P3CS_OK) return bail(rc);
      m.grec_blocks_hi = m.grec_blocks + half;
    }
    m.index_bytes = wide ? 4 : 2;
    const void *dev_index = nullptr;
    if (wide) { const uint32_t *p; rc = upload<uint32_t>(mesh, idx32.data(), idx32.size(), &p); dev_index = p; }
    else { const uint16_t *p; rc = upload<uint16_t>(mesh, idx16.data(), idx16.size(), &p); dev_index = p; }
    if (rc != P3CS_OK) return bail(rc);
    m.index = dev_index;
    if (!full && (rc = upload<uint32_t>(mesh, mask.data(), mask.size(), &m.rows_mask)) != P3CS_OK) return bail(rc);
    if ((rc = upload<int>(mesh, d->blend_offsets, (size_t)d->num_blends + 1, &m.blend_offsets)) != P3CS_OK) return bail(rc);
    if ((rc = upload<int>(mesh, d->blend_transform, (size_t)nnz, &m.blend_transform)) != P3CS_OK) return bail(rc);
    if ((rc = upload<float>(mesh, d->blend_weight, (size_t)nnz, &m.blend_weight)) != P3CS_OK) return bail(rc);
    
  } else {
    m.num_blends = 0;
  }

  // ---- morphs: dense delta columns + slider rows -> per-row CSR, in the reference's
  // iteration order so rows hit by several morphs add up in the same order
  if (d->num_morphs > 0) {
    if (d->morphs == nullptr || d->num_sliders <= 0) return bail(p3cs_fail(P3CS_ERR_INVALID, "morphs without sliders"));
    if (d->num_sliders > 65536) return bail(p3cs_fail(P3CS_ERR_UNSUPPORTED, "more than 65536 sliders"));
    std::vector<int> counts((size_t)N + 1, 0);
    std::vector<int> morph_dyn(d->num_morphs, -1);
    for (int mi = 0; mi < d->num_morphs; ++mi) {
      const p3cs_morph_desc &mo = d->morphs[mi];
      if (mo.slider < 0 || mo.slider >= d->num_sliders) return bail(p3cs_fail(P3CS_ERR_INVALID, "morph %d: slider %d out of range", mi, mo.slider));
      if ((rc = check_column(d, mo.base, "morph base column")) != P3CS_OK) return bail(rc);
      if ((rc = check_column(d, mo.delta, "morph delta column")) != P3CS_OK) return bail(rc);
      if ((rc = check_ranges(mo.row_ranges, mo.num_row_ranges, N, "slider rows")) != P3CS_OK) return bail(rc);
      if (out_of_array[mo.base.array] < 0) return bail(p3cs_fail(P3CS_ERR_INVALID, "morph %d: base column in a non-output array", mi));
      if ((rc = check_dynamic_column(mo.base, "morph base column of morph", mi)) != P3CS_OK) return bail(rc);
      if (!column_supported(mo.delta.contents, mo.delta.numeric_type, mo.delta.num_values))
        return bail(p3cs_fail(P3CS_ERR_UNSUPPORTED, "morph %d: delta column of numeric type %d x %d values, contents %d", mi,
                              mo.delta.numeric_type, mo.delta.num_values, mo.delta.contents));
      int c = find_dyn(out_of_array[mo.base.array], mo.base.start);
      if (c < 0) {
        if (m.num_dyn >= kMaxDynColumns) return bail(p3cs_fail(P3CS_ERR_UNSUPPORTED, "more than %d animated columns", kMaxDynColumns));
        c = m.num_dyn++;
        m.dyn[c].output = (int8_t)out_of_array[mo.base.array];
        m.dyn[c].start = (int16_t)mo.base.start;
        m.dyn[c].num_values = (int8_t)mo.base.num_values;
        m.dyn[c].skin = 0;
        m.dyn[c].f64 = mo.base.numeric_type == P3CS_NT_FLOAT64 ? 1 : 0;
        m.dyn[c].typed = dynamic_column_code(mo.base);
      }
      m.dyn[c].homogeneous = (mo.base.num_values == 4 && (mo.base.contents == P3CS_C_POINT || mo.base.contents == P3CS_C_TEXCOORD)) ? 1 : 0;
      morph_dyn[mi] = c;
      for (int i = 0; i < mo.num_row_ranges; ++i)
        for (int r = mo.row_ranges[2 * i]; r < mo.row_ranges[2 * i + 1]; ++r) counts[r + 1]++;
    }
    for (int r = 0; r < N; ++r) counts[r + 1] += counts[r];
    const size_t total = counts[N];
    std::vector<uint32_t> keys(std::max<size_t>(total, 1));
    std::vector<float4> deltas(std::max<size_t>(total, 1));
    std::vector<int> cursor(counts.begin(), counts.end() - 1);
    for (int mi = 0; mi < d->num_morphs; ++mi) {
      const p3cs_morph_desc &mo = d->morphs[mi];
      const uint8_t *dbase = static_cast<const uint8_t *>(d->arrays[mo.delta.array].data) + mo.delta.start;
      const int dstride = d->arrays[mo.delta.array].stride;
      // the delta is read with get_data4 when the base holds four values without a homogeneous coordinate, else get_data3
      const bool want4 = mo.base.num_values == 4 && !m.dyn[morph_dyn[mi]].homogeneous;
      for (int i = 0; i < mo.num_row_ranges; ++i)
        for (int r = mo.row_ranges[2 * i]; r < mo.row_ranges[2 * i + 1]; ++r) {
          float v[4];  // Packer::get_data3f/4f pad with 0 (geomVertexColumn.cxx:718-737)
          host_get4(mo.delta, dbase + (size_t)r * dstride, v);
          if (!want4) v[3] = 0.0f;
          const int e = cursor[r]++;
          keys[e] = ((uint32_t)morph_dyn[mi] << 16) | (uint32_t)mo.slider;
          // a delta whose 4th component is never read carries its key there (one load per entry
          // in the specialised strip modes)
          if (!(mo.base.num_values == 4 && !m.dyn[morph_dyn[mi]].homogeneous)) memcpy(&v[3], &keys[e], 4);
          deltas[e] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    m.has_morphs = 1;
    if ((rc = upload<int>(mesh, counts.data(), counts.size(), &m.morph_row_offsets)) != P3CS_OK) return bail(rc);
    if ((rc = upload<uint32_t>(mesh, keys.data(), keys.size(), &m.morph_keys)) != P3CS_OK) return bail(rc);
    if ((rc = upload<float4>(mesh, deltas.data(), deltas.size(), &m.morph_deltas)) != P3CS_OK) return bail(rc);
  }
  // ---- run-kernel plans (p3cs_run.cuh) for the row-layout modes
  if ((rc = build_run_plans(mesh, d, plan_row_index, plan_mask)) != P3CS_OK) return bail(rc);
  // ---- slice geometry of the strip kernel (a slice = the 32 rows one warp stages at a time)
  if (m.rows_per_thread == 0) m.rows_per_thread = 1;
  if (m.tile_bufs == 0) m.tile_bufs = 3;
  if (m.strip_rec_floats == 0) m.strip_rec_floats = m.full_records ? kRecFull : kRecCompact;
  m.num_tiles = (N + kStripThreads * m.rows_per_thread - 1) / (kStripThreads * m.rows_per_thread);
  m.slice_bytes = 0;
  for (int o = 0; o < m.num_outputs; ++o) {
    m.slice_out_offset[o] = m.slice_bytes;
    m.slice_bytes += (int)round_up((size_t)32 * m.stride[o], 128);
  }
  if (m.group_rec_offsets == nullptr) {  // no blend table: one tile per group, zero records each
    m.num_groups = m.num_tiles;
    std::vector<int> zeros((size_t)m.num_groups + 1, 0), iota((size_t)m.num_groups + 1);
    for (size_t i = 0; i < iota.size(); ++i) iota[i] = (int)i;
    if ((rc = upload<int>(mesh, zeros.data(), zeros.size(), &m.group_rec_offsets)) != P3CS_OK) return bail(rc);
    if ((rc = upload<int>(mesh, iota.data(), iota.size(), &m.group_tile_offsets)) != P3CS_OK) return bail(rc);
    std::vector<uint4> blocks((size_t)kStripThreads * 2, make_uint4(0, 0, 0, 0));  // the kernel prefetches block [tid]
    if ((rc = upload<uint4>(mesh, blocks.data(), blocks.size(), &m.grec_blocks)) != P3CS_OK) return bail(rc);
    m.grec_blocks_hi = m.grec_blocks + kStripThreads;
  }
  {
    const char *path = getenv("P3CS_PATH");
    mesh->use_strip = strip_smem_bytes(m) <= 200 * 1024 && m.num_transforms < 65536 &&
                      !(path != nullptr && strcmp(path, "wide") == 0);
  }
  // uploads read from host vectors that die at scope exit: finish them now
  cudaError_t err = cudaStreamSynchronize(ctx->stream);
  if (err != cudaSuccess) return bail(p3cs_fail(P3CS_ERR_CUDA, "mesh upload failed: %s", cudaGetErrorString(err)));
  mesh->out_of_array = out_of_array;
  *out = mesh;
  return P3CS_OK;
}

extern "C" int p3cs_mesh_destroy(p3cs_mesh mesh) {
  if (mesh == nullptr) return P3CS_OK;
  DeviceGuard g(mesh->ctx->device);
  for (void *p : mesh->allocations) cudaFree(p);
  delete mesh;
  return P3CS_OK;
}

extern "C" int p3cs_mesh_num_outputs(p3cs_mesh mesh) { return mesh ? mesh->dev.num_outputs : 0; }
extern "C" int p3cs_mesh_output_stride(p3cs_mesh mesh, int o) {
  return (mesh && o >= 0 && o < mesh->dev.num_outputs) ? mesh->dev.stride[o] : 0;
}
extern "C" int p3cs_mesh_num_rows(p3cs_mesh mesh) { return mesh ? mesh->dev.num_rows : 0; }
