/**
 * cudaSkinBackend.h -- Panda-side shim of the p3cudaskin backend (C++, built against the real
 * Panda3D 1.11.0 headers plus the 27-line hook of integration/geomVertexData.patch).
 *
 * This is the host side of the drop-in: it keeps the published API untouched
 * (GeomVertexData, TransformBlendTable, SliderTable, VertexTransform, Character, PRC
 * configuration) and replaces only what GeomVertexData::update_animated_vertices
 * (panda/src/gobj/geomVertexData.cxx:1433-1752) computes, by flattening the objects into
 * p3cs_mesh_desc and calling the C-ABI of libp3cudaskin.so (include/p3cudaskin.h).
 *
 * Selection, in place of `hardware-animated-vertices #f`:
 *     vertex-animation-backend cuda        (default `cpu` = the reference's own loops)
 *     cuda-skinning-library p3cudaskin     (resolved with load_dso on plugin-path, like
 *                                           load-display / audio-library-name plugins:
 *                                           panda/src/audio/audioManager.cxx:63-104)
 *     cuda-skinning-device 0
 *
 * How it is reached.  CudaSkinBackend::install() stores two function pointers in
 * GeomVertexData (the hook): update_animated_vertices() then calls update() instead of its CPU
 * loops, and ~GeomVertexData / clear_animated_vertices() call release().  Everything above the
 * loops stays the reference's: the AT_panda gate, the UpdateSeq staleness test, the `force` /
 * residency rule, the per-object CData locks and the `_animated_vertices` object that is
 * created once and re-used in place.  So every stock caller -- CullableObject::munge_geom
 * (pgraph/cullableObject.cxx:170-175), Geom::get_animated_vertex_data (gobj/geom.cxx:287-288),
 * AnimateVerticesRequest (gobj/animateVerticesRequest.cxx:28) -- lands on the GPU.
 *
 * Crowds.  Vertex datas that share their static half (the copies Character::copy_subgraph makes:
 * same GeomVertexArrayData objects, TransformBlendTables of the same shape over each copy's own
 * joints) are instances of ONE device mesh.  animate_frame() takes a frame's worth of them,
 * gathers the stale ones' joint matrices and animates each mesh's instances with ONE kernel
 * launch, every instance's rows going straight into its own result object's (page-locked)
 * arrays (p3cs_group_set_output_pointers).  The stock animate_vertices() calls of the same
 * frame then find their result already computed.  Trigger: once per frame before the cull
 * traversal (where Character::cull_callback, char/character.cxx:164-202, would update the
 * joints), e.g. from a task; animate_frame(root) walks a scene graph for it.
 */
#ifndef CUDASKINBACKEND_H
#define CUDASKINBACKEND_H

#include "pandabase.h"
#include "geomVertexData.h"
#include "updateSeq.h"
#include "pmap.h"
#include "pvector.h"
#include "lightMutex.h"

#include "p3cudaskin.h"

class VertexTransform;
class PandaNode;

class CudaSkinBackend {
public:
  // The process-wide backend, created from the PRC variables on first use.  Returns nullptr
  // (after logging through gobj's Notify category) when the backend is not selected, the
  // library cannot be loaded or there is no CUDA device: there is NO CPU fallback in here.
  static CudaSkinBackend *get_global_ptr();

  // Points GeomVertexData's hook at the global backend (uninstall: back to the CPU loops).
  static bool install();
  static void uninstall();

  // The hook.  update: fills `animated` (created by update_animated_vertices with the
  // post-animated format, re-used in place from then on) from `source`; called with the
  // source's CData write lock held.  Returns false only when this backend is not in charge
  // (then the reference's loops run); a failure inside is logged and leaves `animated` a plain
  // copy of the source -- deliberately not computed on the CPU.
  bool update(const GeomVertexData *source, GeomVertexData *animated, Thread *current_thread);
  // release: `source` is being destructed, or its animated cache was cleared.
  void release(const GeomVertexData *source);

  // One frame of a crowd: every stale vertex data of `sources` that has been animated before
  // is recomputed, one launch per shared mesh, into its own result object.  Call it from the
  // thread that owns the frame, before the cull traversal.  Returns the number of launches.
  int animate_frame(const GeomVertexData *const *sources, size_t count, Thread *current_thread);
  // The same for the scene graph below `root`: Characters are updated (Character::update, what
  // their cull_callback does), the vertex datas of the GeomNodes below them collected.
  int animate_frame(PandaNode *root, Thread *current_thread);

  // bookkeeping a test can look at
  size_t get_num_entries() const;
  size_t get_num_meshes() const;
  int get_last_frame_instances() const { return _last_frame_instances; }

  const char *last_error() const { return _api->last_error(); }

private:
  CudaSkinBackend(void *dso, const p3cs_api *api, p3cs_context ctx);

  struct SharedMesh;
  struct Entry {
    // what the upload was made from (validated on every call: a recycled address, an edited table)
    const GeomVertexFormat *_format = nullptr;
    int _num_rows = 0;
    UpdateSeq _static_modified;      // GeomVertexData::get_modified() at upload
    const TransformBlendTable *_table = nullptr;
    const SliderTable *_sliders = nullptr;
    SharedMesh *_mesh = nullptr;
    int _slot = -1;                                // instance index in the mesh's group
    pvector<const VertexTransform *> _palette;     // palette slot -> transform (first-appearance order of the table walk)
    PT(GeomVertexData) _animated;                  // the source's result object (its CData's _animated_vertices; kept alive here:
                                                   // many setters drop it there without telling anybody)
    UpdateSeq _result_modified;                    // animation stamp its arrays hold
    pvector<std::pair<void *, size_t> > _registered;  // its arrays, page-locked + mapped (p3cs_host_register)
    UpdateSeq _prepared_modified;                  // stamp of a result animate_frame() has already written
    bool _prepared = false;
  };
  struct SharedMesh {
    p3cs_mesh _mesh = nullptr;
    p3cs_group _group = nullptr;
    int _capacity = 0;
    int _num_palette = 0, _num_sliders = 0, _num_outputs = 0;
    pvector<Entry *> _instances;     // slot -> entry
    pvector<int> _out_arrays;        // source array index per output array
    // the key: static arrays and the flattened table
    pvector<const GeomVertexArrayData *> _arrays;
    const GeomVertexFormat *_format = nullptr;
    pvector<int32_t> _offsets, _index, _rows;
    pvector<float> _weights;
    pvector<const SliderTable *> _slider_tables;   // morphs: shared only with the identical SliderTable
    // per frame
    pvector<float> _palette_values, _slider_values;   // [capacity][T][16], [capacity][S]
    pvector<int32_t> _dirty;
    bool _pointers_stale = true;     // the per-instance output pointer table must be re-sent
  };

  Entry *find_entry(const GeomVertexData *source, bool create);
  bool attach(const GeomVertexData *source, Entry &entry, Thread *current_thread);
  void detach(Entry &entry);
  bool grow(SharedMesh &mesh, int capacity);
  void gather(const Entry &entry, Thread *current_thread);
  bool bind_outputs(Entry &entry, GeomVertexData *animated);
  static UpdateSeq animation_stamp(const GeomVertexData *source, Thread *current_thread);

  void *_dso;
  const p3cs_api *_api;
  p3cs_context _ctx;
  mutable LightMutex _map_lock;   // _entries, _meshes and mesh membership
  LightMutex _gpu_lock;           // calls into the C-ABI (one context, one stream)
  pmap<const GeomVertexData *, Entry *> _entries;
  pvector<SharedMesh *> _meshes;
  int _last_frame_instances = 0;
};

#endif
