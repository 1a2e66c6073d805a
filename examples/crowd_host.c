/* crowd_host.c -- a plain C99 host driving a crowd over several GPUs through the C-ABI of libp3cudaskin.so
 * (include/p3cudaskin.h, p3cs_crowd_*): what a Panda3D built with `cuda-skinning-devices 0 1 2 ...` does per frame.
 *
 *   cc -std=c99 -O2 -Iinclude examples/crowd_host.c -Lpanda3d_b200 -lp3cudaskin -lm -o crowd_host
 *   ./crowd_host <instances> <rows> <joints> <steps> <device> [<device> ...]
 *
 * Builds a synthetic skinned mesh (vertex + normal, 4 joints per vertex in runs of 6 rows), deals the crowd to the
 * devices (devices[0] = the rendering GPU, sink-aware split), tunes the split from measured times, then times `steps`
 * frames (set palettes on the owning devices, one launch per device, rows land in GPU 0's buffer).  Prints one JSON line.
 */
#include "p3cudaskin.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static unsigned int rng_state = 12345u;
static float frand(void) {
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)(rng_state >> 8) / 16777216.0f;
}

int main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: crowd_host <instances> <rows> <joints> <steps> <device> [<device> ...]\n");
    return 2;
  }
  const int n = atoi(argv[1]), rows = atoi(argv[2]), joints = atoi(argv[3]), steps = atoi(argv[4]);
  int32_t devices[16];
  int ndev = 0;
  for (int a = 5; a < argc && ndev < 16; ++a) devices[ndev++] = atoi(argv[a]);
  if (p3cs_device_count() < ndev) {
    printf("{\"ok\": false, \"error\": \"%d CUDA device(s), %d asked for\"}\n", p3cs_device_count(), ndev);
    return 1;
  }
  /* ---- the mesh: array 0 = vertex(3f) normal(3f), array 1 = transform_blend(u16) */
  const int blends = rows / 6 + 1;
  float *a0 = malloc((size_t)rows * 24);
  uint16_t *a1 = malloc((size_t)rows * 2);
  for (int r = 0; r < rows; ++r) {
    float *v = a0 + (size_t)r * 6;
    v[0] = frand() * 2 - 1; v[1] = frand() * 2 - 1; v[2] = frand() * 2 - 1;
    float nx = frand() - 0.5f, ny = frand() - 0.5f, nz = frand() - 0.5f, l = sqrtf(nx * nx + ny * ny + nz * nz) + 1e-9f;
    v[3] = nx / l; v[4] = ny / l; v[5] = nz / l;
    a1[r] = (uint16_t)(r / 6);
  }
  int32_t *offsets = malloc(sizeof(int32_t) * (size_t)(blends + 1)), *xf = malloc(sizeof(int32_t) * 4 * (size_t)blends);
  float *w = malloc(sizeof(float) * 4 * (size_t)blends);
  for (int b = 0; b < blends; ++b) {
    offsets[b] = 4 * b;
    int j0 = (int)(frand() * (joints - 4));
    float sum = 0.0f;
    for (int k = 0; k < 4; ++k) { xf[4 * b + k] = j0 + k; w[4 * b + k] = 0.05f + frand(); sum += w[4 * b + k]; }
    for (int k = 0; k < 4; ++k) w[4 * b + k] /= sum;
  }
  offsets[blends] = 4 * blends;
  p3cs_array_desc arrays[2] = {{a0, 24, 1}, {a1, 2, 0}};
  p3cs_column_desc skin[2] = {{0, 0, 3, P3CS_NT_FLOAT32, P3CS_C_POINT}, {0, 12, 3, P3CS_NT_FLOAT32, P3CS_C_NORMAL}};
  int32_t row_range[2] = {0, rows};
  p3cs_mesh_desc mesh;
  memset(&mesh, 0, sizeof(mesh));
  mesh.struct_size = sizeof(mesh);
  mesh.num_rows = rows;
  mesh.num_arrays = 2;
  mesh.arrays = arrays;
  mesh.num_skin_columns = 2;
  mesh.skin_columns = skin;
  mesh.blend_column.array = 1;
  mesh.blend_column.start = 0;
  mesh.blend_column.num_values = 1;
  mesh.blend_column.numeric_type = P3CS_NT_UINT16;
  mesh.blend_column.contents = P3CS_C_INDEX;
  mesh.num_transforms = joints;
  mesh.num_blends = blends;
  mesh.blend_offsets = offsets;
  mesh.blend_transform = xf;
  mesh.blend_weight = w;
  mesh.num_row_ranges = 1;
  mesh.row_ranges = row_range;

  p3cs_crowd_desc cd;
  memset(&cd, 0, sizeof(cd));
  cd.struct_size = sizeof(cd);
  cd.num_devices = ndev;
  cd.devices = devices;
  cd.num_instances = n;
  p3cs_crowd crowd = NULL;
  if (p3cs_crowd_create(&cd, &mesh, &crowd) != P3CS_OK) {
    printf("{\"ok\": false, \"error\": \"%s\"}\n", p3cs_last_error());
    return 1;
  }
  /* joint matrices: rotation about z by a per-joint angle + translation (row vectors, translation in row 3) */
  float *pal = malloc((size_t)n * joints * 64);
  for (size_t j = 0; j < (size_t)n * joints; ++j) {
    float *m = pal + j * 16, a = frand() * 6.2831853f, c = cosf(a), s = sinf(a);
    memset(m, 0, 64);
    m[0] = c; m[1] = s; m[4] = -s; m[5] = c; m[10] = 1.0f; m[15] = 1.0f;
    m[12] = frand(); m[13] = frand(); m[14] = frand();
  }
  int rc = p3cs_crowd_set_palettes(crowd, 0, n, pal);
  if (rc == P3CS_OK) rc = p3cs_crowd_tune(crowd, 3);
  if (rc == P3CS_OK) rc = p3cs_crowd_set_palettes(crowd, 0, n, pal);
  for (int s = 0; s < 3 && rc == P3CS_OK; ++s) rc = p3cs_crowd_animate(crowd);
  if (rc == P3CS_OK) rc = p3cs_crowd_sync(crowd);
  float ms[16] = {0};
  double device_ms = 0.0;
  const double t0 = now_ms();
  for (int s = 0; s < steps && rc == P3CS_OK; ++s) {
    rc = p3cs_crowd_animate(crowd);
    if (rc == P3CS_OK) rc = p3cs_crowd_sync(crowd);
    if (rc == P3CS_OK) rc = p3cs_crowd_last_ms(crowd, ms);
    double worst = 0.0;
    for (int d = 0; d < ndev; ++d) worst = ms[d] > worst ? ms[d] : worst;
    device_ms += worst;
  }
  const double wall_ms = (now_ms() - t0) / (steps > 0 ? steps : 1);
  if (rc != P3CS_OK) {
    printf("{\"ok\": false, \"error\": \"%s\"}\n", p3cs_last_error());
    return 1;
  }
  /* one instance back from the rendering GPU: its first vertex must be the blend of its joints applied to the source */
  float *row = malloc((size_t)rows * 24);
  rc = p3cs_crowd_read_output(crowd, 0, n - 1, 1, row);
  printf("{\"ok\": %s, \"devices\": %d, \"instances\": %d, \"rows\": %d, \"ms_per_step_device\": %.4f, \"ms_per_step_wall\": %.4f, "
         "\"vertices_per_s\": %.4g, \"split\": [",
         rc == P3CS_OK ? "true" : "false", ndev, n, rows, device_ms / steps, wall_ms, (double)n * rows / (device_ms / steps * 1e-3));
  for (int d = 0; d < ndev; ++d) {
    int32_t first, count;
    p3cs_crowd_range(crowd, d, &first, &count);
    printf("%s%d", d ? ", " : "", count);
  }
  printf("], \"last_ms\": [");
  for (int d = 0; d < ndev; ++d) printf("%s%.4f", d ? ", " : "", ms[d]);
  printf("]}\n");
  p3cs_crowd_destroy(crowd);
  return rc == P3CS_OK ? 0 : 1;
}
