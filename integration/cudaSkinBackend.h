/**
 * cudaSkinBackend.h -- Panda-side shim of the p3cudaskin backend (C++, built against the real
 * Panda3D 1.11.0 headers).
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
 * The hook that routes GeomVertexData::animate_vertices here is a 12-line patch,
 * integration/geomVertexData.patch; without it, callers use
 * CudaSkinBackend::animate_vertices(vdata, force, thread) directly (integration/panda_cuda_check.cxx).
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

class CudaSkinBackend {
public:
  // The process-wide backend, created from the PRC variables on first use.  Returns nullptr
  // (after logging through gobj's Notify category) when the backend is not selected, the
  // library cannot be loaded or there is no CUDA device: there is NO CPU fallback in here.
  static CudaSkinBackend *get_global_ptr();

  // Same contract as GeomVertexData::animate_vertices (geomVertexData.cxx:895-975): returns the
  // source itself when it carries no CPU animation, the cached result while no transform /
  // slider was modified, otherwise recomputes INTO THE SAME result object and returns it.
  CPT(GeomVertexData) animate_vertices(const GeomVertexData *source, bool force, Thread *current_thread);

  // Drops the device copy of a vertex data (call when it is destructed or edited wholesale).
  void release(const GeomVertexData *source);

  const char *last_error() const { return _api->last_error(); }

private:
  CudaSkinBackend(void *dso, const p3cs_api *api, p3cs_context ctx);

  struct Entry {
    UpdateSeq _static_modified;      // GeomVertexData::get_modified() at upload
    const TransformBlendTable *_table = nullptr;
    const SliderTable *_sliders = nullptr;
    p3cs_mesh _mesh = nullptr;
    p3cs_group _group = nullptr;
    pvector<const VertexTransform *> _palette;   // ascending address = palette order
    pvector<int> _out_arrays;                    // source array index per output array
    PT(GeomVertexData) _animated;
    UpdateSeq _animated_modified;
    pvector<float> _palette_values, _slider_values;
    // result arrays page-locked + mapped (p3cs_host_register) so the kernel stores into them directly
    pvector<std::pair<void *, size_t> > _registered;
  };
  bool upload(const GeomVertexData *source, Entry &entry, Thread *current_thread);
  void free_entry(Entry &entry);

  void *_dso;
  const p3cs_api *_api;
  p3cs_context _ctx;
  LightMutex _lock;
  pmap<const GeomVertexData *, Entry> _entries;
};

#endif
