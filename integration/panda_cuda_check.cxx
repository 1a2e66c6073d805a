// panda_cuda_check.cxx -- drop-in demonstration and parity check INSIDE real Panda3D:
// loads an Actor-style model through the reference's own egg loader, poses it, and compares
//   cpu = GeomVertexData::animate_vertices(true, thread)          (the unmodified reference)
//   gpu = CudaSkinBackend::animate_vertices(vdata, true, thread)  (the shim -> C-ABI -> B200)
// byte for byte, per frame; also times both.  Prints one JSON line.  Test infrastructure for
// integration/ (run by tests/test_panda_integration.py on the GPU box).
#include "cudaSkinBackend.h"

#include "animControlCollection.h"
#include "auto_bind.h"
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
#include <cstring>

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: panda_cuda_check <plugin dir> <model.egg> <anim.egg> <frame>...\n");
    return 2;
  }
  load_prc_file_data("check", std::string("plugin-path ") + argv[1] + "\nvertex-animation-backend cuda\n"
                     "cuda-skinning-library p3cudaskin\nhardware-animated-vertices #f\n");
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
  PT(PandaNode) root = new PandaNode("root");
  root->add_child(model);
  if (anim != nullptr) root->add_child(anim);
  AnimControlCollection controls;
  auto_bind(root, controls, ~0);
  NodePath np(root);
  Character *character = DCAST(Character, np.find("**/+Character").node());
  NodePathCollection geom_nodes = np.find_all_matches("**/+GeomNode");
  CPT(GeomVertexData) vdata;
  for (int g = 0; g < geom_nodes.get_num_paths() && vdata == nullptr; ++g) {
    GeomNode *gn = DCAST(GeomNode, geom_nodes.get_path(g).node());
    for (int i = 0; i < gn->get_num_geoms(); ++i) {
      CPT(GeomVertexData) candidate = gn->get_geom(i)->get_vertex_data();
      if (candidate->get_transform_blend_table() != nullptr) {
        vdata = candidate;
        break;
      }
    }
  }
  if (vdata == nullptr || controls.get_num_anims() < 1) {
    printf("{\"ok\": false, \"error\": \"no animated vertex data\"}\n");
    return 1;
  }
  Thread *thread = Thread::get_current_thread();
  long long differing = 0, total = 0;
  double max_abs = 0.0, max_rel = 0.0, cpu_us = 0.0, gpu_us = 0.0;
  int frames = 0;
  bool same_object = true;
  CPT(GeomVertexData) first_gpu;
  for (int a = 4; a < argc; ++a, ++frames) {
    controls.get_anim(0)->pose(atoi(argv[a]));
    character->force_update();
    auto t0 = std::chrono::steady_clock::now();
    CPT(GeomVertexData) cpu = vdata->animate_vertices(true, thread);
    auto t1 = std::chrono::steady_clock::now();
    CPT(GeomVertexData) gpu = backend->animate_vertices(vdata, true, thread);
    auto t2 = std::chrono::steady_clock::now();
    cpu_us += std::chrono::duration<double, std::micro>(t1 - t0).count();
    gpu_us += std::chrono::duration<double, std::micro>(t2 - t1).count();
    if (gpu == vdata) {
      printf("{\"ok\": false, \"error\": \"%s\"}\n", backend->last_error());
      return 1;
    }
    if (first_gpu == nullptr) first_gpu = gpu;
    same_object = same_object && (gpu == first_gpu) && (backend->animate_vertices(vdata, true, thread) == gpu);
    if (gpu->get_format() != cpu->get_format() || gpu->get_num_rows() != cpu->get_num_rows()) {
      printf("{\"ok\": false, \"error\": \"format / row count mismatch\"}\n");
      return 1;
    }
    for (int o = 0; o < cpu->get_num_arrays(); ++o) {
      CPT(GeomVertexArrayDataHandle) hc = cpu->get_array(o)->get_handle(), hg = gpu->get_array(o)->get_handle();
      const unsigned char *pc = hc->get_read_pointer(true), *pg = hg->get_read_pointer(true);
      size_t nbytes = (size_t)cpu->get_num_rows() * cpu->get_format()->get_array(o)->get_stride();
      total += (long long)nbytes;
      for (size_t i = 0; i < nbytes; ++i) differing += pc[i] != pg[i];
      for (size_t i = 0; i + 4 <= nbytes; i += 4) {   // all columns of the panda model are float32
        float fc, fg;
        memcpy(&fc, pc + i, 4);
        memcpy(&fg, pg + i, 4);
        double err = std::fabs((double)fc - (double)fg);
        if (err > max_abs) max_abs = err;
        double rel = err / std::fmax(std::fabs((double)fc), 1e-3);
        if (rel > max_rel) max_rel = rel;
      }
    }
  }
  // steady state: the first call above included the one-time upload of the static half
  double cpu_steady = 0.0, gpu_steady = 0.0;
  const int reps = 50;
  for (int r = 0; r < reps; ++r) {
    controls.get_anim(0)->pose(r % 60);
    character->force_update();
    auto t0 = std::chrono::steady_clock::now();
    CPT(GeomVertexData) cpu = vdata->animate_vertices(true, thread);
    auto t1 = std::chrono::steady_clock::now();
    CPT(GeomVertexData) gpu = backend->animate_vertices(vdata, true, thread);
    auto t2 = std::chrono::steady_clock::now();
    cpu_steady += std::chrono::duration<double, std::micro>(t1 - t0).count();
    gpu_steady += std::chrono::duration<double, std::micro>(t2 - t1).count();
  }
  printf("{\"ok\": true, \"cpu_us_steady\": %.1f, \"cuda_us_steady_e2e\": %.1f, \"frames\": %d, \"rows\": %d, \"bytes_compared\": %lld, \"differing_bytes\": %lld, "
         "\"max_abs_err\": %.3e, \"max_rel_err\": %.3e, \"same_result_object\": %s, \"cpu_us_per_frame\": %.1f, "
         "\"cuda_us_per_frame_e2e\": %.1f}\n",
         cpu_steady / reps, gpu_steady / reps, frames, vdata->get_num_rows(), total, differing, max_abs, max_rel, same_object ? "true" : "false",
         cpu_us / frames, gpu_us / frames);
  return 0;
}
