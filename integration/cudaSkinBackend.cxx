/**
 * cudaSkinBackend.cxx -- see cudaSkinBackend.h.  Host C++ against the real Panda3D API; all the
 * arithmetic happens behind the C-ABI of libp3cudaskin.so.
 */
#include "cudaSkinBackend.h"

#include "config_gobj.h"
#include "config_putil.h"
#include "configVariableInt.h"
#include "configVariableBool.h"
#include "configVariableString.h"
#include "coordinateSystem.h"
#include "geomVertexArrayData.h"
#include "geomVertexFormat.h"
#include "internalName.h"
#include "lightMutexHolder.h"
#include "load_dso.h"
#include "sliderTable.h"
#include "sparseArray.h"
#include "transformBlendTable.h"
#include "vertexSlider.h"
#include "vertexTransform.h"

#include <algorithm>

// The new PRC switch.  `hardware-animated-vertices` stays #f so the
// munger leaves the data AT_panda (display/standardMunger.cxx:125-162); this variable then
// decides who computes update_animated_vertices.
static ConfigVariableString vertex_animation_backend
("vertex-animation-backend", "cpu",
 PRC_DESC("Selects who computes CPU-style (AT_panda) vertex animation: 'cpu' runs "
          "GeomVertexData::update_animated_vertices as always; 'cuda' hands skinning and morph "
          "blending to the p3cudaskin plugin (B200 / sm_100a).  There is no silent fallback: if "
          "'cuda' is selected and the plugin or a CUDA device is missing, an error is logged "
          "and vertices are returned un-animated."));
static ConfigVariableString cuda_skinning_library
("cuda-skinning-library", "p3cudaskin",
 PRC_DESC("Name of the CUDA skinning plugin, loaded as lib<name>.so from plugin-path."));
static ConfigVariableInt cuda_skinning_device
("cuda-skinning-device", 0, PRC_DESC("CUDA device ordinal that animates vertices (the rendering GPU)."));
static ConfigVariableBool cuda_skinning_register_outputs
("cuda-skinning-register-outputs", true,
 PRC_DESC("Page-lock the animated GeomVertexArrayData buffers (cudaHostRegister) so the kernel stores "
          "the animated rows straight into them instead of copying device->host afterwards."));

CudaSkinBackend::CudaSkinBackend(void *dso, const p3cs_api *api, p3cs_context ctx) :
  _dso(dso), _api(api), _ctx(ctx), _lock("CudaSkinBackend") {
}

CudaSkinBackend *CudaSkinBackend::
get_global_ptr() {
  static CudaSkinBackend *global_ptr = nullptr;
  static bool tried = false;
  if (tried) {
    return global_ptr;
  }
  tried = true;
  if (vertex_animation_backend.get_value() != "cuda") {
    return nullptr;
  }
  // Same loading convention as AudioManager (audio/audioManager.cxx:63-104).
  Filename dl_name = Filename::dso_filename("lib" + cuda_skinning_library.get_value() + ".so");
  dl_name.to_os_specific();
  void *handle = load_dso(get_plugin_path().get_value(), dl_name);
  if (handle == nullptr) {
    gobj_cat.error()
      << "vertex-animation-backend cuda: load_dso(" << dl_name << ") failed: " << load_dso_error() << "\n";
    return nullptr;
  }
  void *symbol = get_dso_symbol(handle, "p3cs_get_api");
  if (symbol == nullptr) {
    gobj_cat.error() << dl_name << " does not export p3cs_get_api\n";
    unload_dso(handle);
    return nullptr;
  }
  typedef const p3cs_api *FuncType();
  const p3cs_api *api = (*(FuncType *)symbol)();
  if (api == nullptr || api->version != P3CS_API_VERSION) {
    gobj_cat.error() << dl_name << ": p3cs api version mismatch\n";
    unload_dso(handle);
    return nullptr;
  }
  p3cs_context ctx = nullptr;
  if (api->context_create(cuda_skinning_device, nullptr, &ctx) != P3CS_OK) {
    gobj_cat.error() << "p3cudaskin: " << api->last_error() << "\n";
    unload_dso(handle);
    return nullptr;
  }
  global_ptr = new CudaSkinBackend(handle, api, ctx);
  return global_ptr;
}

void CudaSkinBackend::
free_entry(Entry &entry) {
  for (auto &reg : entry._registered) {
    if (reg.first != nullptr) {
      _api->host_unregister(_ctx, reg.first);
    }
  }
  entry._registered.clear();
  if (entry._group != nullptr) {
    _api->group_destroy(entry._group);
  }
  if (entry._mesh != nullptr) {
    _api->mesh_destroy(entry._mesh);
  }
  entry._group = nullptr;
  entry._mesh = nullptr;
}

void CudaSkinBackend::
release(const GeomVertexData *source) {
  LightMutexHolder holder(_lock);
  auto it = _entries.find(source);
  if (it != _entries.end()) {
    free_entry(it->second);
    _entries.erase(it);
  }
}

static void
append_ranges(const SparseArray &rows, pvector<int32_t> &out) {
  nassertv(!rows.is_inverse());
  for (size_t i = 0; i < rows.get_num_subranges(); ++i) {
    out.push_back(rows.get_subrange_begin(i));
    out.push_back(rows.get_subrange_end(i));
  }
}

static p3cs_column_desc
make_column(const GeomVertexFormat *format, const InternalName *name) {
  const GeomVertexColumn *column = format->get_column(name);
  p3cs_column_desc desc;
  desc.array = format->get_array_with(name);
  desc.start = column->get_start();
  desc.num_values = column->get_num_values();
  desc.numeric_type = (int32_t)column->get_numeric_type();   // GeomEnums values cross the ABI as is
  desc.contents = (int32_t)column->get_contents();
  return desc;
}

/**
 * Flattens the static half of a GeomVertexData into p3cs_mesh_desc and uploads it.
 */
bool CudaSkinBackend::
upload(const GeomVertexData *source, Entry &entry, Thread *current_thread) {
  free_entry(entry);
  const GeomVertexFormat *format = source->get_format();
  CPT(GeomVertexFormat) post = format->get_post_animated_format();
  CPT(TransformBlendTable) table = source->get_transform_blend_table();
  CPT(SliderTable) sliders = source->get_slider_table();
  int num_rows = source->get_num_rows();

  // arrays: the ones that survive get_post_animated_format() are outputs
  pvector<p3cs_array_desc> arrays;
  pvector<CPT(GeomVertexArrayDataHandle)> handles;
  entry._out_arrays.clear();
  for (size_t a = 0; a < format->get_num_arrays(); ++a) {
    const GeomVertexArrayFormat *array_format = format->get_array(a);
    int kept = 0;
    for (int c = 0; c < array_format->get_num_columns(); ++c) {
      const GeomVertexColumn *column = array_format->get_column(c);
      if (column->get_name() != InternalName::get_transform_blend() &&
          column->get_contents() != GeomEnums::C_morph_delta) {
        ++kept;
      }
    }
    CPT(GeomVertexArrayDataHandle) handle = source->get_array(a)->get_handle(current_thread);
    p3cs_array_desc desc;
    desc.data = handle->get_read_pointer(true);
    desc.stride = array_format->get_stride();
    desc.is_output = kept > 0;
    arrays.push_back(desc);
    handles.push_back(handle);
    if (kept > 0) {
      entry._out_arrays.push_back((int)a);
    }
  }
  nassertr(entry._out_arrays.size() == post->get_num_arrays(), false);

  p3cs_mesh_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.struct_size = sizeof(desc);
  desc.num_rows = num_rows;
  desc.coordinate_system = (int32_t)get_default_coordinate_system();  // CS_zup_right = 1 ... CS_yup_left = 4
  desc.num_arrays = (int32_t)arrays.size();
  desc.arrays = arrays.data();

  // points then vectors of the post-animated format (geomVertexData.cxx:1582, 1622)
  pvector<p3cs_column_desc> skin;
  pvector<int32_t> offsets(1, 0), transform_index, row_ranges;
  pvector<float> weights;
  entry._palette.clear();
  desc.blend_column.array = -1;
  if (table != nullptr) {
    if (format->get_array_with(InternalName::get_transform_blend()) < 0) {
      gobj_cat.warning()
        << "Vertex data " << source->get_name()
        << " has a transform_blend_table, but no transform_blend data.\n";   // geomVertexData.cxx:1565-1570
    } else {
      for (size_t i = 0; i < post->get_num_points(); ++i) {
        skin.push_back(make_column(format, post->get_point(i)));
      }
      for (size_t i = 0; i < post->get_num_vectors(); ++i) {
        skin.push_back(make_column(format, post->get_vector(i)));
      }
      desc.blend_column = make_column(format, InternalName::get_transform_blend());
      for (size_t b = 0; b < table->get_num_blends(); ++b) {
        const TransformBlend &blend = table->get_blend(b);
        for (size_t k = 0; k < blend.get_num_transforms(); ++k) {
          entry._palette.push_back(blend.get_transform(k));
        }
      }
      std::sort(entry._palette.begin(), entry._palette.end());
      entry._palette.erase(std::unique(entry._palette.begin(), entry._palette.end()), entry._palette.end());
      for (size_t b = 0; b < table->get_num_blends(); ++b) {
        const TransformBlend &blend = table->get_blend(b);
        for (size_t k = 0; k < blend.get_num_transforms(); ++k) {   // stored order == accumulation order
          const VertexTransform *t = blend.get_transform(k);
          transform_index.push_back((int32_t)(std::lower_bound(entry._palette.begin(), entry._palette.end(), t) - entry._palette.begin()));
          weights.push_back((float)blend.get_weight(k));
        }
        offsets.push_back((int32_t)transform_index.size());
      }
      append_ranges(table->get_rows(), row_ranges);
      desc.num_blends = (int32_t)table->get_num_blends();
    }
  }
  desc.num_skin_columns = (int32_t)skin.size();
  desc.skin_columns = skin.data();
  desc.num_transforms = (int32_t)entry._palette.size();
  desc.blend_offsets = offsets.data();
  desc.blend_transform = transform_index.data();
  desc.blend_weight = weights.data();
  desc.num_row_ranges = (int32_t)row_ranges.size() / 2;
  desc.row_ranges = row_ranges.data();

  // morph applications in the reference's own iteration order (geomVertexData.cxx:1467-1480)
  pvector<p3cs_morph_desc> morphs;
  pvector<pvector<int32_t> > morph_rows;
  if (sliders != nullptr) {
    desc.num_sliders = (int32_t)sliders->get_num_sliders();
    for (size_t mi = 0; mi < format->get_num_morphs(); ++mi) {
      const SparseArray &found = sliders->find_sliders(format->get_morph_slider(mi));
      if (found.is_zero()) {
        continue;
      }
      for (size_t sni = 0; sni < found.get_num_subranges(); ++sni) {
        for (int sn = found.get_subrange_begin(sni); sn < found.get_subrange_end(sni); ++sn) {
          p3cs_morph_desc morph;
          morph.slider = sn;
          morph.base = make_column(format, format->get_morph_base(mi));
          morph.delta = make_column(format, format->get_morph_delta(mi));
          morph_rows.push_back(pvector<int32_t>());
          append_ranges(sliders->get_slider_rows(sn), morph_rows.back());
          morph.num_row_ranges = (int32_t)morph_rows.back().size() / 2;
          morph.row_ranges = nullptr;
          morphs.push_back(morph);
        }
      }
    }
    for (size_t i = 0; i < morphs.size(); ++i) {
      morphs[i].row_ranges = morph_rows[i].data();
    }
  }
  desc.num_morphs = (int32_t)morphs.size();
  desc.morphs = morphs.data();

  if (_api->mesh_create(_ctx, &desc, &entry._mesh) != P3CS_OK ||
      _api->group_create(entry._mesh, 1, nullptr, &entry._group) != P3CS_OK) {
    gobj_cat.error()
      << "p3cudaskin cannot animate " << source->get_name() << ": " << _api->last_error() << "\n";
    free_entry(entry);
    return false;
  }
  entry._static_modified = source->get_modified(current_thread);
  entry._table = table;
  entry._sliders = sliders;
  entry._palette_values.resize(entry._palette.size() * 16);
  entry._slider_values.resize(sliders != nullptr ? sliders->get_num_sliders() : 0);
  return true;
}

CPT(GeomVertexData) CudaSkinBackend::
animate_vertices(const GeomVertexData *source, bool force, Thread *current_thread) {
  // The gate of GeomVertexData::animate_vertices (geomVertexData.cxx:919-960), unchanged.
  if (source->get_format()->get_animation().get_animation_type() != GeomEnums::AT_panda) {
    return source;
  }
  CPT(TransformBlendTable) table = source->get_transform_blend_table();
  CPT(SliderTable) sliders = source->get_slider_table();
  UpdateSeq modified;
  if (table != nullptr) {
    modified = table->get_modified(current_thread);
    if (sliders != nullptr) {
      modified = std::max(modified, sliders->get_modified(current_thread));
    }
  } else if (sliders != nullptr) {
    modified = sliders->get_modified(current_thread);
  } else {
    return source;
  }

  LightMutexHolder holder(_lock);
  Entry &entry = _entries[source];
  if (entry._animated != nullptr && entry._animated_modified == modified) {
    return entry._animated;   // no changes
  }
  if (!force && !source->request_resident()) {
    return entry._animated != nullptr ? (const GeomVertexData *)entry._animated : source;
  }
  if (entry._mesh == nullptr || entry._static_modified != source->get_modified(current_thread) ||
      entry._table != table || entry._sliders != sliders) {
    if (!upload(source, entry, current_thread)) {
      return source;   // logged; deliberately NOT computed on the CPU
    }
  }
  if (entry._animated == nullptr) {   // geomVertexData.cxx:1448-1453
    entry._animated = new GeomVertexData(source->get_name(), source->get_format()->get_post_animated_format(),
                                         std::min(source->get_usage_hint(), GeomEnums::UH_dynamic));
  }
  PT(GeomVertexData) result = entry._animated;
  result->unclean_set_num_rows(source->get_num_rows());

  // "Character joint-matrix export": the palette in ascending-address order
  for (size_t j = 0; j < entry._palette.size(); ++j) {
    LMatrix4 mat;
    entry._palette[j]->get_matrix(mat);
    LMatrix4f matf = LCAST(float, mat);
    memcpy(&entry._palette_values[j * 16], matf.get_data(), 64);
  }
  for (size_t s = 0; s < entry._slider_values.size(); ++s) {
    entry._slider_values[s] = (float)sliders->get_slider(s)->get_slider();
  }
  // the kernel writes straight into the result's arrays (get_write_pointer bumps their
  // modified stamp, so the GSG re-uploads them: geomVertexArrayData.cxx:601-606)
  pvector<PT(GeomVertexArrayDataHandle)> out_handles;
  pvector<void *> out_pointers;
  for (size_t o = 0; o < entry._out_arrays.size(); ++o) {
    PT(GeomVertexArrayDataHandle) handle = result->modify_array_handle(o);
    void *pointer = handle->get_write_pointer();
    out_pointers.push_back(pointer);
    out_handles.push_back(handle);
    // Page-lock the result's buffer once (again only if Panda reallocated it): the kernel then stores the
    // animated rows straight into it, no device->host copy.  A failed registration just keeps the copy.
    if (cuda_skinning_register_outputs) {
      if (entry._registered.size() <= o) {
        entry._registered.resize(o + 1, std::make_pair((void *)nullptr, (size_t)0));
      }
      size_t bytes = (size_t)handle->get_data_size_bytes();
      if (entry._registered[o].first != pointer || entry._registered[o].second != bytes) {
        if (entry._registered[o].first != nullptr) {
          _api->host_unregister(_ctx, entry._registered[o].first);
        }
        entry._registered[o] = std::make_pair((void *)nullptr, (size_t)0);
        if (bytes > 0 && _api->host_register(_ctx, pointer, bytes) == P3CS_OK) {
          entry._registered[o] = std::make_pair(pointer, bytes);
        }
      }
    }
  }
  if (_api->animate_vertices(entry._group, 0, entry._palette_values.data(), entry._slider_values.data(),
                             out_pointers.data()) != P3CS_OK) {
    gobj_cat.error() << "p3cudaskin: " << _api->last_error() << "\n";
    return source;
  }
  entry._animated_modified = modified;
  return result;
}
