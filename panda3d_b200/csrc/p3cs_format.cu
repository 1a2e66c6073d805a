// p3cs_format.cu -- "next" row f3: the AT_hardware upload format, host side only.
//
// GeomVertexData::copy_from (panda/src/gobj/geomVertexData.cxx:574-648, indexed branch) flattens a
// TransformBlendTable into per-vertex transform_weight / transform_index columns plus a TransformTable in
// first-use order; p3cs_hardware_from_blends restates that conversion on a p3cs_mesh_desc (pinned against the
// reference's convert_to() by tests/golden/hardware/), p3cs_blends_from_hardware goes the other way so a mesh
// stored in the hardware format (a pre-flattened .bam, another engine's export) can feed p3cs_mesh_create.
#include "../../include/p3cudaskin.h"

#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {

struct Entry {
  int32_t transform;
  float weight;
};

int32_t read_blend_index(const p3cs_mesh_desc *d, int row) {
  const p3cs_column_desc &bc = d->blend_column;
  const uint8_t *p = static_cast<const uint8_t *>(d->arrays[bc.array].data) + (size_t)row * d->arrays[bc.array].stride + bc.start;
  if (bc.numeric_type == P3CS_NT_UINT8) return *p;
  if (bc.numeric_type == P3CS_NT_UINT16) { uint16_t v; memcpy(&v, p, 2); return v; }
  uint32_t v;
  memcpy(&v, p, 4);
  return (int32_t)v;
}

}  // namespace

extern "C" int p3cs_hardware_from_blends(const p3cs_mesh_desc *d, float *weights, int32_t *indices, int32_t *table, int32_t *table_size) {
  if (d == nullptr || weights == nullptr || indices == nullptr || table == nullptr || table_size == nullptr) return P3CS_ERR_INVALID;
  if (d->blend_column.array < 0 || d->blend_column.array >= d->num_arrays || d->blend_offsets == nullptr) return P3CS_ERR_INVALID;
  const int nt = d->blend_column.numeric_type;
  if (nt != P3CS_NT_UINT8 && nt != P3CS_NT_UINT16 && nt != P3CS_NT_UINT32) return P3CS_ERR_UNSUPPORTED;
  std::vector<int32_t> slot_of(d->num_transforms > 0 ? d->num_transforms : 1, -1);  // add_transform()'s already_added map
  int32_t used = 0;
  auto add_transform = [&](int32_t t) {
    if (slot_of[t] < 0) {
      slot_of[t] = used;
      table[used++] = t;
    }
    return slot_of[t];
  };
  std::vector<Entry> e;
  for (int r = 0; r < d->num_rows; ++r) {  // `while (!from.is_at_end())`: every row, whatever get_rows() says
    const int32_t b = read_blend_index(d, r);
    if (b < 0 || b >= d->num_blends) return P3CS_ERR_INVALID;
    float *w = weights + (size_t)r * 4;
    int32_t *ix = indices + (size_t)r * 4;
    w[0] = w[1] = w[2] = w[3] = 0.0f;
    ix[0] = ix[1] = ix[2] = ix[3] = 0;
    e.clear();
    for (int k = d->blend_offsets[b]; k < d->blend_offsets[b + 1]; ++k) {
      if (d->blend_transform[k] < 0 || d->blend_transform[k] >= d->num_transforms) return P3CS_ERR_INVALID;
      e.push_back(Entry{d->blend_transform[k], d->blend_weight[k]});
    }
    if (e.size() > 4) {
      // TransformBlend::limit_transforms(4): drop the (first) least weight until four remain
      // (transformBlend.cxx:95-118), then normalize_weights() (:125-140)
      while (e.size() > 4) {
        size_t least = 0;
        for (size_t i = 1; i < e.size(); ++i)
          if (e[i].weight < e[least].weight) least = i;
        e.erase(e.begin() + (long)least);
      }
      float net = 0.0f;
      for (const Entry &x : e) net += x.weight;
      if (net != 0.0f)
        for (Entry &x : e) x.weight /= net;
    }
    for (size_t i = 0; i < e.size(); ++i) {
      w[i] = e[i].weight;
      ix[i] = add_transform(e[i].transform);
    }
  }
  *table_size = used;
  return P3CS_OK;
}

extern "C" int p3cs_blends_from_hardware(int32_t num_rows, const float *weights, const int32_t *indices, const int32_t *table,
                                         int32_t table_size, int32_t *row_blend, int32_t *blend_offsets, int32_t *blend_transform,
                                         float *blend_weight, int32_t *num_blends) {
  if (num_rows < 0 || weights == nullptr || indices == nullptr || table == nullptr || row_blend == nullptr || blend_offsets == nullptr ||
      blend_transform == nullptr || blend_weight == nullptr || num_blends == nullptr)
    return P3CS_ERR_INVALID;
  // a blend = the row's non-zero slots in slot order (an entry of a TransformBlend never has weight 0:
  // TransformBlend::add_transform ignores those, transformBlend.cxx:52-53); identical rows share one blend
  typedef std::vector<uint32_t> Key;
  std::map<Key, int32_t> seen;
  int32_t nb = 0, nnz = 0;
  blend_offsets[0] = 0;
  for (int r = 0; r < num_rows; ++r) {
    Key key;
    for (int k = 0; k < 4; ++k) {
      const float w = weights[(size_t)r * 4 + k];
      if (w == 0.0f) continue;
      const int32_t slot = indices[(size_t)r * 4 + k];
      if (slot < 0 || slot >= table_size) return P3CS_ERR_INVALID;
      uint32_t bits;
      memcpy(&bits, &w, 4);
      key.push_back((uint32_t)table[slot]);
      key.push_back(bits);
    }
    auto it = seen.find(key);
    if (it == seen.end()) {
      for (size_t i = 0; i < key.size(); i += 2) {
        blend_transform[nnz] = (int32_t)key[i];
        memcpy(&blend_weight[nnz], &key[i + 1], 4);
        ++nnz;
      }
      blend_offsets[nb + 1] = nnz;
      it = seen.emplace(key, nb++).first;
    }
    row_blend[r] = it->second;
  }
  *num_blends = nb;
  return P3CS_OK;
}
