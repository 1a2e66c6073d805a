// panda_cuda_check.cxx -- drop-in demonstration and parity check INSIDE real Panda3D, through the hook of
// integration/geomVertexData.patch (the libpanda of oracle/_ref/lib_hook is the reference plus that hook):
// loads an Actor-style model through the reference's own egg loader, poses it, and compares
//   cpu = the reference's loops        (hook not installed: GeomVertexData::update_animated_vertices as shipped)
//   gpu = the CUDA backend             (hook installed: the SAME stock calls land in CudaSkinBackend::update)
// byte for byte, per frame, through every stock caller: GeomVertexData::animate_vertices, Geom::get_animated_vertex_data
// (gobj/geom.cxx:287-288) and AnimateVerticesRequest (gobj/animateVerticesRequest.cxx:28).
//   panda_cuda_check <plugin dir> <model.egg> <anim.egg> <frame>...           one character
//   panda_cuda_check <plugin dir> <model.egg> <anim.egg> --crowd N [frames]   N copies of it, one launch per frame
// Prints one JSON line.  Test infrastructure for integration/ (run by tests/test_panda_integration.py on the GPU box).
#include "cudaSkinBackend.h"

#include "animControlCollection.h"
#include "animateVerticesRequest.h"
#include "asyncTaskManager.h"
#include "auto_bind.h"
#include "bamFile.h"
#include "character.h"
#include "geomNode.h"
#include "geomVertexArrayData.h"
#include "load_egg_file.h"
#include "load_prc_file.h"
#include "nodePath.h"
#include "nodePathCollection.h"
#include "pandaNode.h"
#include "thread.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

typedef std::chrono::steady_clock Clock;
static double micros(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }

struct Compare {
  long long differing = 0, total = 0, integer_differing = 0;
  double max_abs = 0.0, max_rel = 0.0;
  bool shape_ok = true;
  void add(const GeomVertexData *cpu, const GeomVertexData *gpu) {
    if (gpu->get_format() != cpu->get_format() || gpu->get_num_rows() != cpu->get_num_rows()) {
      shape_ok = false;
      return;
    }
    for (size_t o = 0; o < cpu->get_num_arrays(); ++o) {
      CPT(GeomVertexArrayDataHandle) hc = cpu->get_array(o)->get_handle(), hg = gpu->get_array(o)->get_handle();
      const unsigned char *pc = hc->get_read_pointer(true), *pg = hg->get_read_pointer(true);
      size_t nbytes = (size_t)cpu->get_num_rows() * cpu->get_format()->get_array(o)->get_stride();
      total += (long long)nbytes;
      for (size_t i = 0; i < nbytes; ++i) differing += pc[i] != pg[i];
      // float32 columns: numeric error per component; every other column (packed colours, integers): bytes, counted apart
      const GeomVertexArrayFormat *af = cpu->get_format()->get_array(o);
      const size_t stride = af->get_stride();
      for (int ci = 0; ci < af->get_num_columns(); ++ci) {
        const GeomVertexColumn *col = af->get_column(ci);
        for (int row = 0; row < cpu->get_num_rows(); ++row) {
          const size_t at = (size_t)row * stride + col->get_start();
          if (col->get_numeric_type() != GeomEnums::NT_float32) {
            for (int b = 0; b < col->get_total_bytes(); ++b) integer_differing += pc[at + b] != pg[at + b];
            continue;
          }
          for (int k = 0; k < col->get_num_components(); ++k) {
            float fc, fg;
            memcpy(&fc, pc + at + 4 * k, 4);
            memcpy(&fg, pg + at + 4 * k, 4);
            double err = std::fabs((double)fc - (double)fg);
            if (err > max_abs) max_abs = err;
            double rel = err / std::fmax(std::fabs((double)fc), 1e-3);
            if (rel > max_rel) max_rel = rel;
          }
        }
      }
    }
  }
};

// The animated vertex data of a character subgraph: (GeomNode's Geom, its vertex data)
static bool find_animated(const NodePath &np, CPT(Geom) &geom, CPT(GeomVertexData) &vdata) {
  NodePathCollection geom_nodes = np.find_all_matches("**/+GeomNode");
  for (int g = 0; g < geom_nodes.get_num_paths(); ++g) {
    GeomNode *gn = DCAST(GeomNode, geom_nodes.get_path(g).node());
    for (int i = 0; i < gn->get_num_geoms(); ++i) {
      CPT(GeomVertexData) candidate = gn->get_geom(i)->get_vertex_data();
      if (candidate->get_transform_blend_table() != nullptr || candidate->get_slider_table() != nullptr) {
        geom = gn->get_geom(i);
        vdata = candidate;
        return true;
      }
    }
  }
  return false;
}

static int run_crowd(CudaSkinBackend *backend, PandaNode *model, PandaNode *anim, int n, int frames);

// P3CS_CHECK_BAM=<dir>: the loaded subgraph goes through the reference's own .bam writer and reader first
// (TransformBlendTable::write_datagram / fillin, gobj/transformBlendTable.cxx:210-284; SliderTable, JointVertexTransform,
// Character likewise), so what the backend flattens is what a shipped application loads from disk.
static PT(PandaNode) through_bam(PandaNode *node, const std::string &path) {
  {
    BamFile out;
    if (!out.open_write(Filename(path)) || !out.write_object(node)) return nullptr;
    out.close();
  }
  BamFile in;
  if (!in.open_read(Filename(path))) return nullptr;
  PT(PandaNode) back = in.read_node();
  if (back == nullptr || !in.resolve()) return nullptr;
  in.close();
  return back;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: panda_cuda_check <plugin dir> <model.egg> <anim.egg> <frame>... | --crowd N [frames]\n");
    return 2;
  }
  load_prc_file_data("check", std::string("plugin-path ") + argv[1] + "\nvertex-animation-backend cuda\n"
                     "cuda-skinning-library p3cudaskin\nhardware-animated-vertices #f\n");
  if (const char *extra = getenv("P3CS_CHECK_PRC")) load_prc_file_data("check-extra", extra);   // e.g. vertex-colors-prefer-packed 1
  CudaSkinBackend *backend = CudaSkinBackend::get_global_ptr();
  if (backend == nullptr) {
    printf("{\"ok\": false, \"error\": \"backend unavailable (see the gobj error log)\"}\n");
    return 1;
  }
  const bool one_file = strcmp(argv[2], argv[3]) == 0;  // e.g. models/ripple.egg holds model and animation
  PT(PandaNode) model = load_egg_file(argv[2]);
  PT(PandaNode) anim = one_file ? nullptr : load_egg_file(argv[3]);
  if (model == nullptr || (!one_file && anim == nullptr)) {
    printf("{\"ok\": false, \"error\": \"egg load failed\"}\n");
    return 1;
  }
  if (const char *bam_dir = getenv("P3CS_CHECK_BAM")) {
    model = through_bam(model, std::string(bam_dir) + "/p3cs_check_model.bam");
    if (anim != nullptr) anim = through_bam(anim, std::string(bam_dir) + "/p3cs_check_anim.bam");
    if (model == nullptr || (!one_file && anim == nullptr)) {
      printf("{\"ok\": false, \"error\": \"bam round trip failed\"}\n");
      return 1;
    }
  }
  if (strcmp(argv[4], "--crowd") == 0) {
    return run_crowd(backend, model, anim, argc > 5 ? atoi(argv[5]) : 1000, argc > 6 ? atoi(argv[6]) : 20);
  }

  PT(PandaNode) root = new PandaNode("root");
  root->add_child(model);
  if (anim != nullptr) root->add_child(anim);
  AnimControlCollection controls;
  auto_bind(root, controls, ~0);
  NodePath np(root);
  Character *character = DCAST(Character, np.find("**/+Character").node());
  CPT(Geom) geom;
  CPT(GeomVertexData) vdata;
  if (!find_animated(np, geom, vdata) || controls.get_num_anims() < 1) {
    printf("{\"ok\": false, \"error\": \"no animated vertex data\"}\n");
    return 1;
  }
  // the CPU side animates a copy of the vertex data (same arrays, same tables), so that each side keeps its own
  // cached result object the way a running application would
  PT(GeomVertexData) cpu_vdata = new GeomVertexData(*vdata);
  Thread *thread = Thread::get_current_thread();
  Compare cmp;
  double cpu_us = 0.0, gpu_us = 0.0;
  int frames = 0;
  bool same_object = true, stock_callers_agree = true;
  CPT(GeomVertexData) first_gpu;
  for (int a = 4; a < argc; ++a, ++frames) {
    controls.get_anim(0)->pose(atoi(argv[a]));
    character->force_update();
    CudaSkinBackend::uninstall();
    auto t0 = Clock::now();
    CPT(GeomVertexData) cpu = cpu_vdata->animate_vertices(true, thread);
    auto t1 = Clock::now();
    CudaSkinBackend::install();
    CPT(GeomVertexData) gpu = vdata->animate_vertices(true, thread);
    auto t2 = Clock::now();
    cpu_us += micros(t0, t1);
    gpu_us += micros(t1, t2);
    if (gpu == vdata || cpu == cpu_vdata) {
      printf("{\"ok\": false, \"error\": \"not animated: %s\"}\n", backend->last_error());
      return 1;
    }
    if (first_gpu == nullptr) first_gpu = gpu;
    // same result object, re-used in place; a second call without changes is the cache hit of geomVertexData.cxx:956-960
    same_object = same_object && (gpu == first_gpu) && (vdata->animate_vertices(true, thread) == gpu);
    // the other stock callers reach the same object
    stock_callers_agree = stock_callers_agree && geom->get_animated_vertex_data(true, thread) == gpu;
    cmp.add(cpu, gpu);
    if (!cmp.shape_ok) {
      printf("{\"ok\": false, \"error\": \"format / row count mismatch\"}\n");
      return 1;
    }
  }
  // AnimateVerticesRequest: pose a new frame, let the task animate, then the direct call must be a cache hit on it
  bool request_ok = false;
  {
    controls.get_anim(0)->pose(13);
    character->force_update();
    CudaSkinBackend::uninstall();
    CPT(GeomVertexData) cpu = cpu_vdata->animate_vertices(true, thread);
    CudaSkinBackend::install();
    PT(AnimateVerticesRequest) request = new AnimateVerticesRequest(const_cast<GeomVertexData *>(vdata.p()));
    AsyncTaskManager *manager = AsyncTaskManager::get_global_ptr();
    manager->add(request);
    for (int spin = 0; spin < 1000 && !request->is_ready(); ++spin) manager->poll();
    CPT(GeomVertexData) gpu = vdata->animate_vertices(true, thread);
    Compare one;
    one.add(cpu, gpu);
    request_ok = request->is_ready() && gpu == first_gpu && one.shape_ok && one.max_rel <= 1e-5;
    cmp.add(cpu, gpu);
    ++frames;
  }
  // steady state: the first call above included the one-time upload of the static half
  double cpu_steady = 0.0, gpu_steady = 0.0;
  const int reps = 50;
  for (int r = 0; r < reps; ++r) {
    controls.get_anim(0)->pose(r % 60);
    character->force_update();
    CudaSkinBackend::uninstall();
    auto t0 = Clock::now();
    CPT(GeomVertexData) cpu = cpu_vdata->animate_vertices(true, thread);
    auto t1 = Clock::now();
    CudaSkinBackend::install();
    CPT(GeomVertexData) gpu = vdata->animate_vertices(true, thread);
    auto t2 = Clock::now();
    cpu_steady += micros(t0, t1);
    gpu_steady += micros(t1, t2);
  }
  // lifetime: dropping the vertex data's animated cache releases the backend's entry
  const size_t entries_before = backend->get_num_entries();
  const_cast<GeomVertexData *>(vdata.p())->clear_animated_vertices();
  const size_t entries_after = backend->get_num_entries();
  std::string columns;   // "vertex:3x5 normal:3x5 color:4x0 ...": name : values x GeomEnums::NumericType, animated arrays only
  for (size_t o = 0; o < first_gpu->get_num_arrays(); ++o) {
    const GeomVertexArrayFormat *af = first_gpu->get_format()->get_array(o);
    for (int ci = 0; ci < af->get_num_columns(); ++ci) {
      const GeomVertexColumn *col = af->get_column(ci);
      columns += (columns.empty() ? "" : " ") + col->get_name()->get_name() + ":" + std::to_string(col->get_num_values()) + "x" +
                 std::to_string((int)col->get_numeric_type());
    }
  }
  printf("{\"ok\": true, \"columns\": \"%s\", \"integer_column_differing_bytes\": %lld, \"morphs\": %d, \"cpu_us_steady\": %.1f, \"cuda_us_steady_e2e\": %.1f, \"frames\": %d, \"rows\": %d, \"bytes_compared\": %lld, \"differing_bytes\": %lld, "
         "\"max_abs_err\": %.3e, \"max_rel_err\": %.3e, \"same_result_object\": %s, \"stock_callers_agree\": %s, \"animate_vertices_request\": %s, "
         "\"entries_before_clear\": %zu, \"entries_after_clear\": %zu, \"cpu_us_per_frame\": %.1f, \"cuda_us_per_frame_e2e\": %.1f}\n",
         columns.c_str(), cmp.integer_differing, (int)vdata->get_format()->get_num_morphs(), cpu_steady / reps, gpu_steady / reps, frames, vdata->get_num_rows(), cmp.total, cmp.differing, cmp.max_abs, cmp.max_rel,
         same_object ? "true" : "false", stock_callers_agree ? "true" : "false", request_ok ? "true" : "false", entries_before, entries_after,
         cpu_us / (frames - 1), gpu_us / (frames - 1));
  return 0;
}

// A crowd of N copies of the character (NodePath::copy_to -> Character::copy_subgraph: every copy has its own joints,
// vertex data, TransformBlendTable over ITS joints, and shares the static arrays), each playing its own frame.
// Per frame: pose every copy, CudaSkinBackend::animate_frame(root) -- ONE launch for the whole crowd --, then the stock
// Geom::get_animated_vertex_data of every copy (a cache hit).  The CPU side is the same loop with the hook off.
static int run_crowd(CudaSkinBackend *backend, PandaNode *model, PandaNode *anim, int n, int frames) {
  Thread *thread = Thread::get_current_thread();
  PT(PandaNode) root = new PandaNode("crowd");
  NodePath root_np(root);
  struct Actor {
    NodePath np;
    AnimControlCollection controls;
    Character *character = nullptr;
    CPT(Geom) geom;
    CPT(GeomVertexData) vdata;
    PT(GeomVertexData) cpu_vdata;
  };
  std::vector<Actor *> actors;
  NodePath proto(model);
  for (int i = 0; i < n; ++i) {
    Actor *actor = new Actor;
    PT(PandaNode) holder = new PandaNode("actor");
    root->add_child(holder);
    actor->np = NodePath(holder);
    proto.copy_to(actor->np);
    if (anim != nullptr) NodePath(anim).copy_to(actor->np);
    auto_bind(holder, actor->controls, ~0);
    actor->character = DCAST(Character, actor->np.find("**/+Character").node());
    if (!find_animated(actor->np, actor->geom, actor->vdata) || actor->controls.get_num_anims() < 1) {
      printf("{\"ok\": false, \"error\": \"copy %d has no animated vertex data\"}\n", i);
      return 1;
    }
    actor->cpu_vdata = new GeomVertexData(*actor->vdata);
    actors.push_back(actor);
  }
  const int rows = actors[0]->vdata->get_num_rows();
  Compare cmp;
  double cpu_us = 0.0, gpu_us = 0.0, gpu_first_us = 0.0;
  int launches_last = 0, batched_last = 0;
  bool cache_hits = true;
  for (int f = 0; f < frames; ++f) {
    for (int i = 0; i < n; ++i) {
      actors[i]->controls.get_anim(0)->pose((f * 3 + i * 7) % 60);
      actors[i]->character->force_update();
    }
    // ---- CPU: the reference's loops, one call per character
    CudaSkinBackend::uninstall();
    auto t0 = Clock::now();
    for (int i = 0; i < n; ++i) actors[i]->cpu_vdata->animate_vertices(true, thread);
    auto t1 = Clock::now();
    // ---- CUDA: one batch call for the frame, then the stock calls
    CudaSkinBackend::install();
    launches_last = backend->animate_frame(root, thread);
    batched_last = backend->get_last_frame_instances();
    auto t2 = Clock::now();
    std::vector<CPT(GeomVertexData)> results(n);
    for (int i = 0; i < n; ++i) results[i] = actors[i]->geom->get_animated_vertex_data(true, thread);
    auto t3 = Clock::now();
    if (f == 0) {
      gpu_first_us = micros(t1, t3);   // first frame: uploads, registrations, one call per character
    } else {
      cpu_us += micros(t0, t1);
      gpu_us += micros(t1, t3);
      cache_hits = cache_hits && micros(t2, t3) < micros(t1, t2) * 4 + 1e5;
    }
    // compare a spread of the copies every frame, all of them on the last
    for (int i = 0; i < n; i += (f + 1 == frames ? 1 : std::max(1, n / 16))) {
      CPT(GeomVertexData) cpu = actors[i]->cpu_vdata->animate_vertices(true, thread);   // cache hit (hook state does not matter)
      cmp.add(cpu, results[i]);
    }
  }
  const double steady = frames > 1 ? frames - 1 : 1;
  printf("{\"ok\": %s, \"crowd\": %d, \"rows\": %d, \"frames\": %d, \"device_meshes\": %zu, \"launches_last_frame\": %d, \"batched_last_frame\": %d, "
         "\"bytes_compared\": %lld, \"differing_bytes\": %lld, \"max_rel_err\": %.3e, \"cpu_ms_per_frame\": %.3f, \"cuda_ms_per_frame_e2e\": %.3f, "
         "\"cuda_first_frame_ms\": %.3f, \"cpu_vertices_per_s\": %.4g, \"cuda_vertices_per_s\": %.4g}\n",
         cmp.shape_ok ? "true" : "false", n, rows, frames, backend->get_num_meshes(), launches_last, batched_last, cmp.total, cmp.differing, cmp.max_rel,
         cpu_us / steady / 1e3, gpu_us / steady / 1e3, gpu_first_us / 1e3, (double)n * rows / (cpu_us / steady * 1e-6),
         (double)n * rows / (gpu_us / steady * 1e-6));
  return cmp.shape_ok ? 0 : 1;
}
