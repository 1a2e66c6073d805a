/*
 * plugin_host.c -- a plain C host of libp3cudaskin.so, loaded the way Panda3D loads its plug-ins
 * (load_dso + get_dso_symbol, panda/src/audio/audioManager.cxx:63-104): dlopen, dlsym("p3cs_get_api"),
 * then everything through the version-stamped table of function pointers.
 *
 *   plugin_host <path/to/libp3cudaskin.so> [--run]
 *
 * Without --run it only loads the library and checks the table (works without a GPU); with --run it skins a
 * four-vertex mesh -- the survey's probe: v = (2,4,1) under 0.25 * T(1,2,3) + 0.75 * S(2) must give
 * (3.75, 7.5, 2.5) -- and prints the result.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "p3cudaskin.h"

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <libp3cudaskin.so> [--run]\n", argv[0]);
    return 2;
  }
  void *dso = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (dso == NULL) {
    fprintf(stderr, "dlopen failed: %s\n", dlerror());
    return 1;
  }
  const p3cs_api *(*get_api)(void) = (const p3cs_api *(*)(void))dlsym(dso, "p3cs_get_api");
  if (get_api == NULL) {
    fprintf(stderr, "p3cs_get_api missing\n");
    return 1;
  }
  const p3cs_api *api = get_api();
  if (api == NULL || api->version != P3CS_API_VERSION) {
    fprintf(stderr, "api version mismatch\n");
    return 1;
  }
  printf("p3cs api version %d, %d CUDA device(s)\n", api->version, api->device_count());
  if (argc < 3 || strcmp(argv[2], "--run") != 0) return 0;

  p3cs_context ctx = NULL;
  if (api->context_create(0, NULL, &ctx) != P3CS_OK) {
    fprintf(stderr, "context_create: %s\n", api->last_error());  /* no CPU fallback: this is the end */
    return 1;
  }
  /* [vertex(3f) normal(3f)] rows + a transform_blend(u16) array; two transforms, two blends */
  float rows[4][6] = {{2, 4, 1, 0, 0, 1}, {1, 0, 0, 1, 0, 0}, {0, 1, 0, 0, 1, 0}, {0, 0, 1, 0, 0, 1}};
  unsigned short blend_index[4] = {0, 1, 1, 0};
  p3cs_array_desc arrays[2] = {{rows, 24, 1}, {blend_index, 2, 0}};
  p3cs_column_desc skin[2] = {{0, 0, 3, P3CS_NT_FLOAT32, P3CS_C_POINT}, {0, 12, 3, P3CS_NT_FLOAT32, P3CS_C_NORMAL}};
  int32_t blend_offsets[3] = {0, 2, 3}, blend_transform[3] = {0, 1, 0}, row_ranges[2] = {0, 4};
  float blend_weight[3] = {0.25f, 0.75f, 1.0f};
  p3cs_mesh_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.struct_size = sizeof(desc);
  desc.num_rows = 4;
  desc.num_arrays = 2;
  desc.arrays = arrays;
  desc.num_skin_columns = 2;
  desc.skin_columns = skin;
  desc.blend_column.array = 1;
  desc.blend_column.start = 0;
  desc.blend_column.num_values = 1;
  desc.blend_column.numeric_type = P3CS_NT_UINT16;
  desc.blend_column.contents = P3CS_C_INDEX;
  desc.num_transforms = 2;
  desc.num_blends = 2;
  desc.blend_offsets = blend_offsets;
  desc.blend_transform = blend_transform;
  desc.blend_weight = blend_weight;
  desc.num_row_ranges = 1;
  desc.row_ranges = row_ranges;
  p3cs_mesh mesh = NULL;
  p3cs_group group = NULL;
  if (api->mesh_create(ctx, &desc, &mesh) != P3CS_OK || api->group_create(mesh, 1, NULL, &group) != P3CS_OK) {
    fprintf(stderr, "upload: %s\n", api->last_error());
    return 1;
  }
  /* row vectors, v' = v * M: transform 0 translates by (1,2,3), transform 1 scales by 2 */
  float palette[2][16] = {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 1, 2, 3, 1}, {2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 1}};
  float out[4][6];
  void *outputs[1] = {out};
  if (api->animate_vertices(group, 0, &palette[0][0], NULL, outputs) != P3CS_OK) {
    fprintf(stderr, "animate_vertices: %s\n", api->last_error());
    return 1;
  }
  printf("vertex 0 -> (%g, %g, %g), normal (%g, %g, %g)\n", out[0][0], out[0][1], out[0][2], out[0][3], out[0][4], out[0][5]);
  api->group_destroy(group);
  api->mesh_destroy(mesh);
  api->context_destroy(ctx);
  return (out[0][0] == 3.75f && out[0][1] == 7.5f && out[0][2] == 2.5f) ? 0 : 3;
}
