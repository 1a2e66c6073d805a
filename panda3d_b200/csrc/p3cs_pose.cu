// p3cs_pose.cu -- "next" row f1: the joint hierarchy on the GPU.
//
// One CTA per Character instance evaluates, for the instance's animation frame,
//   AnimChannelMatrixXfmTable::get_value  (panda/src/chan/animChannelMatrixXfmTable.cxx:112-124:
//                                          12 table components -> compose_matrix),
//   CharacterJoint::update_internals      (panda/src/char/characterJoint.cxx:110-150:
//                                          net = value * parent net, skinning = initial^-1 * net)
// (local values for all joints in parallel, then the net transforms level by level: one
// __syncthreads per tree level) and writes
// the palette straight into the group's device buffer: with it a crowd needs one int per
// instance and frame from the host instead of T x 64 bytes of matrices.
// Covers the reference's normal case: one AnimControl, no frame blending
// (chan/movingPartMatrix.cxx:71-76); same -fmad=false arithmetic as the rest of the library.
#include "p3cs_device.cuh"

namespace p3cs {

// MATRIX4_PRODUCT, panda/src/linmath/lmatrix4_src.I:874-893: res = a * b
__device__ __forceinline__ void mul4(float *res, const float *a, const float *b) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      res[i * 4 + j] = a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j] + a[i * 4 + 2] * b[2 * 4 + j] + a[i * 4 + 3] * b[3 * 4 + j];
}

__global__ void __launch_bounds__(256) pose_kernel(SkeletonDev skel, float *palettes, int num_transforms, const int *frames,
                                                    int first_instance) {
  extern __shared__ float smem_pose[];  // [J][16] local values, then [J][16] net transforms of this instance
  const int inst = first_instance + blockIdx.x;
  const int frame = __ldg(frames + blockIdx.x);
  const int J = skel.num_joints;
  float *local = smem_pose;
  float *net = smem_pose + (size_t)J * 16;

  // phase 1, all joints in parallel: the joint's own value at this frame
  // (AnimChannelMatrixXfmTable::get_value -> compose_matrix, or the default value)
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float value[16];
    if (__ldg(skel.has_channel + j)) {
      float comp[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int len = __ldg(skel.table_length + j * 12 + i);
        comp[i] = len == 0 ? (i < 3 ? 1.0f : 0.0f) : __ldg(skel.tables + __ldg(skel.table_offset + j * 12 + i) + frame % len);
      }
      Mat3 up3;
      compose3(up3, comp + 0, comp + 3, comp + 6, skel.coordinate_system);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) value[r * 4 + c] = up3.m[r][c];
        value[r * 4 + 3] = 0.0f;
      }
      value[12] = comp[9];
      value[13] = comp[10];
      value[14] = comp[11];
      value[15] = 1.0f;
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) value[k] = __ldg(skel.default_value + j * 16 + k);
    }
    float4 *dst = reinterpret_cast<float4 *>(local + j * 16);
#pragma unroll
    for (int r = 0; r < 4; ++r) dst[r] = make_float4(value[r * 4], value[r * 4 + 1], value[r * 4 + 2], value[r * 4 + 3]);
  }
  __syncthreads();

  // phase 2, level by level: _net_transform = _value * parent's _net_transform (or the root transform).
  // Joints are stored sorted by level, so level L is the contiguous range [level_begin[L], level_begin[L+1]).
  for (int level = 0; level <= skel.max_level; ++level) {
    const int lb = __ldg(skel.level_begin + level), le = __ldg(skel.level_begin + level + 1);
    for (int k = lb + threadIdx.x; k < le; k += blockDim.x) {
      const int j = __ldg(skel.level_order + k);
      const int parent = __ldg(skel.parent + j);
      float value[16], par[16], out[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        value[q] = local[j * 16 + q];
        par[q] = parent >= 0 ? net[parent * 16 + q] : __ldg(skel.root_xform + q);
      }
      mul4(out, value, par);
#pragma unroll
      for (int q = 0; q < 16; ++q) net[j * 16 + q] = out[q];
    }
    __syncthreads();
  }
  // _skinning_matrix = _initial_net_transform_inverse * _net_transform, gathered in palette order
  float *pal = palettes + (size_t)inst * num_transforms * 16;
  for (int t = threadIdx.x; t < num_transforms; t += blockDim.x) {
    const int j = __ldg(skel.palette_joint + t);
    float inv[16], n[16], out[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      inv[k] = __ldg(skel.initial_inverse + j * 16 + k);
      n[k] = net[j * 16 + k];
    }
    mul4(out, inv, n);
    float4 *dst = reinterpret_cast<float4 *>(pal + t * 16);
#pragma unroll
    for (int r = 0; r < 4; ++r) dst[r] = make_float4(out[r * 4], out[r * 4 + 1], out[r * 4 + 2], out[r * 4 + 3]);
  }
}

void launch_pose(const SkeletonDev &skel, float *palettes, int num_transforms, const int *frames_dev, int first_instance, int count,
                 cudaStream_t stream) {
  if (count == 0 || skel.num_joints == 0) return;
  const size_t smem = (size_t)skel.num_joints * 128;
  if (smem > 48 * 1024) cudaFuncSetAttribute(pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int threads = ((skel.num_joints + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  pose_kernel<<<count, threads, smem, stream>>>(skel, palettes, num_transforms, frames_dev, first_instance);
}

}  // namespace p3cs
