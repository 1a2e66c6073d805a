/**
 * cudaSkinBackend.cxx -- see cudaSkinBackend.h.  Host C++ against the real Panda3D API; all the
 * arithmetic happens behind the C-ABI of libp3cudaskin.so.
 */
#include "cudaSkinBackend.h"

#include "character.h"
#include "config_gobj.h"
#include "config_putil.h"
#include "configVariableInt.h"
#include "configVariableBool.h"
#include "configVariableString.h"
#include "coordinateSystem.h"
#include "geomNode.h"
#include "geomVertexArrayData.h"
#include "geomVertexFormat.h"
#include "internalName.h"
#include "lightMutexHolder.h"
#include "load_dso.h"
#include "pandaNode.h"
#include "pStatCollector.h"
#include "pStatTimer.h"
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
          "the animated rows straight into them instead of copying device->host afterwards.  The "
          "frame-batch path (one launch per shared mesh) needs it."));

// PStats: the hook runs inside update_animated_vertices' own PStatTimer ("*:Animation:<character>",
// geomVertexData.cxx:1435), so a character's GPU time shows up where its CPU time did.  The phases of the backend sit
// beside the reference's "Skinning" / "Morphs" / "Calc blends" children (geomVertexData.cxx:37, 46-49) under the
// same root; the frame batch, which runs outside any vertex data's timer, gets its own.
static PStatCollector _cuda_upload_pcollector("*:Animation:CUDA:Upload mesh");
static PStatCollector _cuda_gather_pcollector("*:Animation:CUDA:Gather joints");
static PStatCollector _cuda_animate_pcollector("*:Animation:CUDA:Animate");
static PStatCollector _cuda_frame_pcollector("*:Animation:CUDA:Frame batch");

CudaSkinBackend::CudaSkinBackend(void *dso, const p3cs_api *api, p3cs_context ctx) :
  _dso(dso), _api(api), _ctx(ctx), _map_lock("CudaSkinBackend map"), _gpu_lock("CudaSkinBackend gpu") {
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

// ---- the hook (integration/geomVertexData.patch)
static bool
hook_update(const GeomVertexData *source, GeomVertexData *animated, Thread *current_thread) {
  CudaSkinBackend *backend = CudaSkinBackend::get_global_ptr();
  return backend != nullptr && backend->update(source, animated, current_thread);
}

static void
hook_release(const GeomVertexData *source) {
  CudaSkinBackend *backend = CudaSkinBackend::get_global_ptr();
  if (backend != nullptr) {
    backend->release(source);
  }
}

bool CudaSkinBackend::
install() {
  if (get_global_ptr() == nullptr) {
    return false;
  }
  GeomVertexData::_update_animated_vertices_backend = &hook_update;
  GeomVertexData::_release_animated_vertices_backend = &hook_release;
  return true;
}

void CudaSkinBackend::
uninstall() {
  GeomVertexData::_update_animated_vertices_backend = nullptr;
  GeomVertexData::_release_animated_vertices_backend = nullptr;
}

size_t CudaSkinBackend::
get_num_entries() const {
  LightMutexHolder holder(_map_lock);
  return _entries.size();
}

size_t CudaSkinBackend::
get_num_meshes() const {
  LightMutexHolder holder(_map_lock);
  return _meshes.size();
}

CudaSkinBackend::Entry *CudaSkinBackend::
find_entry(const GeomVertexData *source, bool create) {
  LightMutexHolder holder(_map_lock);
  auto it = _entries.find(source);
  if (it != _entries.end()) {
    return it->second;
  }
  if (!create) {
    return nullptr;
  }
  Entry *entry = new Entry;
  _entries[source] = entry;
  return entry;
}

/**
 * Takes the entry out of its mesh (the last instance of a mesh takes the device copy with it).
 */
void CudaSkinBackend::
detach(Entry &entry) {
  for (auto &reg : entry._registered) {
    if (reg.first != nullptr) {
      _api->host_unregister(_ctx, reg.first);
    }
  }
  entry._registered.clear();
  SharedMesh *mesh = entry._mesh;
  if (mesh != nullptr) {
    LightMutexHolder holder(_map_lock);
    // the last instance moves into the vacated slot
    Entry *last = mesh->_instances.back();
    mesh->_instances[entry._slot] = last;
    last->_slot = entry._slot;
    mesh->_instances.pop_back();
    mesh->_pointers_stale = true;
    if (mesh->_instances.empty()) {
      LightMutexHolder gpu(_gpu_lock);
      if (mesh->_group != nullptr) {
        _api->group_destroy(mesh->_group);
      }
      if (mesh->_mesh != nullptr) {
        _api->mesh_destroy(mesh->_mesh);
      }
      _meshes.erase(std::find(_meshes.begin(), _meshes.end(), mesh));
      delete mesh;
    }
  }
  entry._mesh = nullptr;
  entry._slot = -1;
  entry._prepared = false;
}

void CudaSkinBackend::
release(const GeomVertexData *source) {
  Entry *entry = nullptr;
  {
    LightMutexHolder holder(_map_lock);
    if (_entries.empty()) {
      return;   // the common case for the thousands of vertex datas that are not animated
    }
    auto it = _entries.find(source);
    if (it == _entries.end()) {
      return;
    }
    entry = it->second;
    _entries.erase(it);
  }
  detach(*entry);
  delete entry;
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
 * (Re)creates the mesh's group for `capacity` instances.
 */
bool CudaSkinBackend::
grow(SharedMesh &mesh, int capacity) {
  LightMutexHolder gpu(_gpu_lock);
  if (mesh._group != nullptr) {
    _api->group_destroy(mesh._group);
    mesh._group = nullptr;
  }
  if (_api->group_create(mesh._mesh, capacity, nullptr, &mesh._group) != P3CS_OK) {
    return false;
  }
  mesh._capacity = capacity;
  mesh._palette_values.resize((size_t)capacity * mesh._num_palette * 16);
  mesh._slider_values.resize((size_t)capacity * mesh._num_sliders);
  mesh._pointers_stale = true;
  return true;
}

/**
 * Flattens the static half of a GeomVertexData and attaches the entry to the device mesh it
 * describes: an existing one when another vertex data has the same arrays and a table of the
 * same shape (the copies of one Character), else a new upload.
 */
bool CudaSkinBackend::
attach(const GeomVertexData *source, Entry &entry, Thread *current_thread) {
  detach(entry);
  const GeomVertexFormat *format = source->get_format();
  CPT(GeomVertexFormat) post = format->get_post_animated_format();
  CPT(TransformBlendTable) table = source->get_transform_blend_table();
  CPT(SliderTable) sliders = source->get_slider_table();
  int num_rows = source->get_num_rows();

  // arrays: the ones that survive get_post_animated_format() are outputs
  pvector<p3cs_array_desc> arrays;
  pvector<CPT(GeomVertexArrayDataHandle)> handles;
  pvector<const GeomVertexArrayData *> array_objects;
  pvector<int> out_arrays;
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
    CPT(GeomVertexArrayData) array = source->get_array(a);
    CPT(GeomVertexArrayDataHandle) handle = array->get_handle(current_thread);
    p3cs_array_desc desc;
    desc.data = handle->get_read_pointer(true);
    desc.stride = array_format->get_stride();
    desc.is_output = kept > 0;
    arrays.push_back(desc);
    handles.push_back(handle);
    array_objects.push_back(array.p());
    if (kept > 0) {
      out_arrays.push_back((int)a);
    }
  }
  nassertr(out_arrays.size() == post->get_num_arrays(), false);

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
      // Palette slots in order of first appearance while the table is walked (blends in order, entries in their
      // stored order == accumulation order): copies of a Character, whose joints are different objects at
      // different addresses, then flatten to the same arrays and share one device mesh.
      pmap<const VertexTransform *, int> slot_of;
      for (size_t b = 0; b < table->get_num_blends(); ++b) {
        const TransformBlend &blend = table->get_blend(b);
        for (size_t k = 0; k < blend.get_num_transforms(); ++k) {
          const VertexTransform *t = blend.get_transform(k);
          auto found = slot_of.find(t);
          int slot;
          if (found == slot_of.end()) {
            slot = (int)entry._palette.size();
            slot_of[t] = slot;
            entry._palette.push_back(t);
          } else {
            slot = found->second;
          }
          transform_index.push_back(slot);
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

  const int num_sliders = sliders != nullptr ? (int)sliders->get_num_sliders() : 0;

  // ---- an existing device mesh of the same description?
  SharedMesh *mesh = nullptr;
  {
    LightMutexHolder holder(_map_lock);
    for (SharedMesh *candidate : _meshes) {
      if (candidate->_format == format && candidate->_arrays == array_objects && candidate->_offsets == offsets &&
          candidate->_index == transform_index && candidate->_rows == row_ranges && candidate->_num_sliders == num_sliders &&
          candidate->_weights.size() == weights.size() &&
          memcmp(candidate->_weights.data(), weights.data(), weights.size() * sizeof(float)) == 0 &&
          (num_sliders == 0 || candidate->_slider_tables[0] == sliders.p())) {
        mesh = candidate;
        break;
      }
    }
  }
  if (mesh == nullptr) {
    // morph applications in the reference's own iteration order (geomVertexData.cxx:1467-1480)
    pvector<p3cs_morph_desc> morphs;
    pvector<pvector<int32_t> > morph_rows;
    if (sliders != nullptr) {
      desc.num_sliders = num_sliders;
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

    mesh = new SharedMesh;
    {
      LightMutexHolder gpu(_gpu_lock);
      if (_api->mesh_create(_ctx, &desc, &mesh->_mesh) != P3CS_OK) {
        gobj_cat.error()
          << "p3cudaskin cannot animate " << source->get_name() << ": " << _api->last_error() << "\n";
        delete mesh;
        return false;
      }
    }
    mesh->_format = format;
    mesh->_arrays = array_objects;
    mesh->_offsets = offsets;
    mesh->_index = transform_index;
    mesh->_weights = weights;
    mesh->_rows = row_ranges;
    mesh->_num_palette = (int)entry._palette.size();
    mesh->_num_sliders = num_sliders;
    mesh->_num_outputs = (int)out_arrays.size();
    mesh->_out_arrays = out_arrays;
    if (num_sliders > 0) {
      mesh->_slider_tables.push_back(sliders.p());
    }
    if (!grow(*mesh, 1)) {
      gobj_cat.error() << "p3cudaskin: " << _api->last_error() << "\n";
      LightMutexHolder gpu(_gpu_lock);
      _api->mesh_destroy(mesh->_mesh);
      delete mesh;
      return false;
    }
    LightMutexHolder holder(_map_lock);
    _meshes.push_back(mesh);
  }
  {
    LightMutexHolder holder(_map_lock);
    entry._slot = (int)mesh->_instances.size();
    mesh->_instances.push_back(&entry);
    mesh->_pointers_stale = true;
  }
  if ((int)mesh->_instances.size() > mesh->_capacity && !grow(*mesh, std::max(2 * mesh->_capacity, (int)mesh->_instances.size()))) {
    gobj_cat.error() << "p3cudaskin: " << _api->last_error() << "\n";
    entry._mesh = mesh;
    detach(entry);
    return false;
  }
  entry._mesh = mesh;
  entry._format = format;
  entry._num_rows = num_rows;
  entry._static_modified = source->get_modified(current_thread);
  entry._table = table;
  entry._sliders = sliders;
  return true;
}

UpdateSeq CudaSkinBackend::
animation_stamp(const GeomVertexData *source, Thread *current_thread) {
  // as GeomVertexData::animate_vertices, geomVertexData.cxx:936-954
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
  }
  return modified;
}

/**
 * "Character joint-matrix export": the instance's joint matrices and slider values, in palette
 * order, into its slot of the mesh's host buffers.
 */
void CudaSkinBackend::
gather(const Entry &entry, Thread *current_thread) {
  SharedMesh &mesh = *entry._mesh;
  float *palette = mesh._palette_values.data() + (size_t)entry._slot * mesh._num_palette * 16;
  for (size_t j = 0; j < entry._palette.size(); ++j) {
    LMatrix4 mat;
    entry._palette[j]->get_matrix(mat);
    LMatrix4f matf = LCAST(float, mat);
    memcpy(palette + j * 16, matf.get_data(), 64);
  }
  float *values = mesh._slider_values.data() + (size_t)entry._slot * mesh._num_sliders;
  for (int s = 0; s < mesh._num_sliders; ++s) {
    values[s] = (float)entry._sliders->get_slider(s)->get_slider();
  }
}

/**
 * Makes the result object's arrays the destination of the instance: sized, stamped as modified
 * (get_write_pointer, so the GSG re-uploads them: geomVertexArrayData.cxx:601-606) and
 * page-locked.  Returns false when an array cannot be written in place by the kernel.
 */
bool CudaSkinBackend::
bind_outputs(Entry &entry, GeomVertexData *animated) {
  SharedMesh &mesh = *entry._mesh;
  if (entry._animated != animated) {
    entry._animated = animated;
    entry._prepared = false;
  }
  animated->unclean_set_num_rows(entry._num_rows);
  bool in_place = cuda_skinning_register_outputs;
  if (entry._registered.size() != (size_t)mesh._num_outputs) {
    entry._registered.resize(mesh._num_outputs, std::make_pair((void *)nullptr, (size_t)0));
  }
  for (int o = 0; o < mesh._num_outputs; ++o) {
    PT(GeomVertexArrayDataHandle) handle = animated->modify_array_handle(o);
    void *pointer = handle->get_write_pointer();
    size_t bytes = (size_t)handle->get_data_size_bytes();
    if (entry._registered[o].first != pointer || entry._registered[o].second != bytes) {
      LightMutexHolder gpu(_gpu_lock);
      if (entry._registered[o].first != nullptr) {
        _api->host_unregister(_ctx, entry._registered[o].first);
      }
      entry._registered[o] = std::make_pair((void *)nullptr, (size_t)0);
      if (cuda_skinning_register_outputs && bytes > 0 && bytes % 16 == 0 && ((uintptr_t)pointer % 16) == 0 &&
          _api->host_register(_ctx, pointer, bytes) == P3CS_OK) {
        entry._registered[o] = std::make_pair(pointer, bytes);
      }
      mesh._pointers_stale = true;
    }
    in_place = in_place && entry._registered[o].first != nullptr;
  }
  return in_place;
}

bool CudaSkinBackend::
update(const GeomVertexData *source, GeomVertexData *animated, Thread *current_thread) {
  Entry *entry = find_entry(source, true);
  CPT(TransformBlendTable) table = source->get_transform_blend_table();
  CPT(SliderTable) sliders = source->get_slider_table();
  if (entry->_mesh == nullptr || entry->_format != source->get_format() || entry->_num_rows != source->get_num_rows() ||
      entry->_static_modified != source->get_modified(current_thread) || entry->_table != table || entry->_sliders != sliders) {
    PStatTimer upload_timer(_cuda_upload_pcollector, current_thread);
    if (!attach(source, *entry, current_thread)) {
      animated->copy_from(source, true);   // logged; deliberately NOT computed on the CPU
      return true;
    }
  }
  const UpdateSeq modified = animation_stamp(source, current_thread);
  if (entry->_prepared && entry->_prepared_modified == modified && entry->_animated == animated) {
    return true;   // animate_frame() has already written this frame's result into `animated`
  }
  bind_outputs(*entry, animated);
  SharedMesh &mesh = *entry->_mesh;
  // this thread's own copy of the joint matrices / slider values: other threads may be animating (or attaching)
  // other instances of the same mesh right now
  pvector<float> palette(entry->_palette.size() * 16), slider_values((size_t)mesh._num_sliders);
  _cuda_gather_pcollector.start(current_thread);
  for (size_t j = 0; j < entry->_palette.size(); ++j) {
    LMatrix4 mat;
    entry->_palette[j]->get_matrix(mat);
    LMatrix4f matf = LCAST(float, mat);
    memcpy(&palette[j * 16], matf.get_data(), 64);
  }
  for (int s = 0; s < mesh._num_sliders; ++s) {
    slider_values[s] = (float)entry->_sliders->get_slider(s)->get_slider();
  }
  _cuda_gather_pcollector.stop(current_thread);
  pvector<void *> out_pointers;
  pvector<PT(GeomVertexArrayDataHandle)> out_handles;
  for (int o = 0; o < mesh._num_outputs; ++o) {
    PT(GeomVertexArrayDataHandle) handle = animated->modify_array_handle(o);
    out_pointers.push_back(handle->get_write_pointer());
    out_handles.push_back(handle);
  }
  int rc;
  {
    PStatTimer animate_timer(_cuda_animate_pcollector, current_thread);
    LightMutexHolder gpu(_gpu_lock);   // the GPU section only: flattening above ran unlocked
    rc = _api->animate_vertices(mesh._group, entry->_slot, palette.data(), slider_values.data(), out_pointers.data());
  }
  if (rc != P3CS_OK) {
    gobj_cat.error() << "p3cudaskin: " << _api->last_error() << "\n";
    animated->copy_from(source, true);
    return true;
  }
  entry->_result_modified = modified;
  entry->_prepared = false;
  return true;
}

int CudaSkinBackend::
animate_frame(const GeomVertexData *const *sources, size_t count, Thread *current_thread) {
  PStatTimer frame_timer(_cuda_frame_pcollector, current_thread);
  pvector<SharedMesh *> touched;
  _last_frame_instances = 0;
  for (size_t i = 0; i < count; ++i) {
    const GeomVertexData *source = sources[i];
    Entry *entry = find_entry(source, false);
    if (entry == nullptr || entry->_mesh == nullptr || entry->_animated == nullptr) {
      continue;   // never animated yet: its first animate_vertices() uploads it through the hook
    }
    if (entry->_format != source->get_format() || entry->_num_rows != source->get_num_rows() ||
        entry->_static_modified != source->get_modified(current_thread) ||
        entry->_table != source->get_transform_blend_table() || entry->_sliders != source->get_slider_table()) {
      continue;   // edited: the hook re-attaches it
    }
    const UpdateSeq modified = animation_stamp(source, current_thread);
    if (entry->_result_modified == modified || (entry->_prepared && entry->_prepared_modified == modified)) {
      continue;   // not stale
    }
    if (!bind_outputs(*entry, entry->_animated)) {
      continue;   // cannot be written in place: single path
    }
    gather(*entry, current_thread);
    SharedMesh *mesh = entry->_mesh;
    if (mesh->_dirty.empty()) {
      touched.push_back(mesh);
    }
    mesh->_dirty.push_back(entry->_slot);
    entry->_prepared = true;
    entry->_prepared_modified = modified;
    entry->_result_modified = modified;
    ++_last_frame_instances;
  }
  int launches = 0;
  if (touched.empty()) {
    return 0;
  }
  LightMutexHolder gpu(_gpu_lock);
  for (SharedMesh *mesh : touched) {
    bool ok = true;
    const int n = (int)mesh->_instances.size();
    if (mesh->_pointers_stale) {
      // every slot of the group needs a valid destination; the ones past the instances are never animated
      for (int o = 0; o < mesh->_num_outputs && ok; ++o) {
        pvector<void *> table((size_t)mesh->_capacity, nullptr);
        for (int s = 0; s < mesh->_capacity; ++s) {
          const Entry *e = mesh->_instances[s < n ? s : 0];
          table[s] = e->_registered.size() > (size_t)o ? e->_registered[o].first : nullptr;
          if (table[s] == nullptr) {
            // an instance that was never bound (it is not in the dirty list either): point it at a bound one
            table[s] = mesh->_instances[mesh->_dirty[0]]->_registered[o].first;
          }
        }
        ok = _api->group_set_output_pointers(mesh->_group, o, table.data()) == P3CS_OK;
      }
      mesh->_pointers_stale = !ok;
    }
    ok = ok && _api->group_set_palettes(mesh->_group, 0, n, mesh->_palette_values.data()) == P3CS_OK;
    if (mesh->_num_sliders > 0) {
      ok = ok && _api->group_set_sliders(mesh->_group, 0, n, mesh->_slider_values.data()) == P3CS_OK;
    }
    ok = ok && _api->group_animate_instances(mesh->_group, mesh->_dirty.data(), (int)mesh->_dirty.size()) == P3CS_OK;
    if (!ok) {
      gobj_cat.error() << "p3cudaskin: " << _api->last_error() << "\n";
      for (int slot : mesh->_dirty) {   // their animate_vertices() calls go through the single path
        mesh->_instances[slot]->_prepared = false;
        mesh->_instances[slot]->_result_modified = UpdateSeq();
      }
    } else {
      ++launches;
    }
    mesh->_dirty.clear();
  }
  if (_api->context_sync(_ctx) != P3CS_OK) {
    gobj_cat.error() << "p3cudaskin: " << _api->last_error() << "\n";
  }
  return launches;
}

static void
collect_vertex_datas(PandaNode *node, pvector<CPT(GeomVertexData)> &out, Thread *current_thread) {
  if (node->is_of_type(Character::get_class_type())) {
    DCAST(Character, node)->update();   // what Character::cull_callback does (char/character.cxx:164-202)
  }
  if (node->is_geom_node()) {
    GeomNode *geom_node = DCAST(GeomNode, node);
    for (int i = 0; i < geom_node->get_num_geoms(); ++i) {
      CPT(GeomVertexData) vdata = geom_node->get_geom(i)->get_vertex_data(current_thread);
      if (vdata->get_format()->get_animation().get_animation_type() == GeomEnums::AT_panda &&
          (vdata->get_transform_blend_table() != nullptr || vdata->get_slider_table() != nullptr)) {
        out.push_back(vdata);
      }
    }
  }
  PandaNode::Children children = node->get_children(current_thread);
  for (size_t c = 0; c < children.get_num_children(); ++c) {
    collect_vertex_datas(children.get_child(c), out, current_thread);
  }
}

int CudaSkinBackend::
animate_frame(PandaNode *root, Thread *current_thread) {
  pvector<CPT(GeomVertexData)> held;
  collect_vertex_datas(root, held, current_thread);
  pvector<const GeomVertexData *> sources;
  sources.reserve(held.size());
  for (const CPT(GeomVertexData) &vdata : held) {
    sources.push_back(vdata.p());
  }
  return animate_frame(sources.data(), sources.size(), current_thread);
}
