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
// Covers one AnimControl, without frame blending (chan/movingPartMatrix.cxx:71-76) or with it
// (PartBundle::set_frame_blend_flag: BT_linear and BT_normalized_linear, :78-198); same -fmad=false arithmetic
// as the rest of the library.
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

// One table component of a joint at a frame: table[frame % size], or the default of an empty table
// (1 for the scales, 0 otherwise; animChannelMatrixXfmTable.I:75-85).
__device__ __forceinline__ float channel_component(const SkeletonDev &skel, int j, int i, int frame) {
  const int len = __ldg(skel.table_length + j * 12 + i);
  return len == 0 ? (i < 3 ? 1.0f : 0.0f) : __ldg(skel.tables + __ldg(skel.table_offset + j * 12 + i) + frame % len);
}

// AnimChannelMatrixXfmTable::get_value (animChannelMatrixXfmTable.cxx:112-124), or get_value_no_scale_shear
// (:131-149) with unit scale and zero shear: compose_matrix of the components -> LMatrix4f(upper3, translate).
__device__ __forceinline__ void channel_value(const SkeletonDev &skel, int j, int frame, bool no_scale_shear, float *value) {
  float comp[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) comp[i] = (no_scale_shear && i < 6) ? (i < 3 ? 1.0f : 0.0f) : channel_component(skel, j, i, frame);
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
}

// next_frames == nullptr: one AnimControl, no frame blending (movingPartMatrix.cxx:71-76).  Otherwise the bundle's
// frame-blend flag is on: the values at frame and next_frame are blended by frac, blend_type 0 = BT_linear
// (:82-122), 1 = BT_normalized_linear (:125-198, the PRC default), one control of effect 1.
template <bool BLEND>
__global__ void __launch_bounds__(256) pose_kernel(SkeletonDev skel, float *palettes, int num_transforms, const int *frames,
                                                    const int *next_frames, const float *fracs, int blend_type, int first_instance) {
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
    if (BLEND && __ldg(skel.has_channel + j)) {
      const int next_frame = __ldg(next_frames + blockIdx.x);
      const float frac = __ldg(fracs + blockIdx.x);
      // weights e0 = effect * (1 - frac), e1 = effect * frac with effect = 1; the sums start from zero and are
      // divided by net_effect through a reciprocal (LMatrix4 / LVecBase3 operator /=)
      const float effect = 1.0f;
      const float e0 = effect * (1.0f - frac), e1 = effect * frac;
      const float recip = 1.0f / (0.0f + effect);
      float v0[16], v1[16];
      channel_value(skel, j, frame, blend_type != 0, v0);
      channel_value(skel, j, next_frame, blend_type != 0, v1);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float n = 0.0f;
        n = n + v0[k] * e0;
        n = n + v1[k] * e1;
        value[k] = n * recip;
      }
      if (blend_type != 0) {
        float scale[3], shear[3], false_scale[3], false_shear[3], hpr[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float sc = 0.0f, sh = 0.0f;
          sc = sc + channel_component(skel, j, k, frame) * e0;
          sc = sc + channel_component(skel, j, k, next_frame) * e1;
          sh = sh + channel_component(skel, j, 3 + k, frame) * e0;
          sh = sh + channel_component(skel, j, 3 + k, next_frame) * e1;
          scale[k] = sc * recip;
          shear[k] = sh * recip;
        }
        // decompose_matrix(net_value, false_scale, false_shear, hpr, translate), then
        // compose_matrix(_value, scale, shear, hpr, translate)
        Mat3 up3, out3;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) up3.m[r][c] = value[r * 4 + c];
        decompose3(up3, skel.coordinate_system, false_scale, false_shear, hpr);
        compose3(out3, scale, shear, hpr, skel.coordinate_system);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) value[r * 4 + c] = out3.m[r][c];
          value[r * 4 + 3] = 0.0f;
        }
        value[15] = 1.0f;  // translate = row 3 of the blended matrix stays
      }
    } else if (__ldg(skel.has_channel + j)) {
      channel_value(skel, j, frame, false, value);
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

void launch_pose(const SkeletonDev &skel, float *palettes, int num_transforms, const int *frames_dev, const int *next_frames_dev,
                 const float *fracs_dev, int blend_type, int first_instance, int count, cudaStream_t stream) {
  if (count == 0 || skel.num_joints == 0) return;
  const size_t smem = (size_t)skel.num_joints * 128;
  int threads = ((skel.num_joints + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (next_frames_dev != nullptr) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(pose_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pose_kernel<true><<<count, threads, smem, stream>>>(skel, palettes, num_transforms, frames_dev, next_frames_dev, fracs_dev, blend_type,
                                                        first_instance);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(pose_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pose_kernel<false><<<count, threads, smem, stream>>>(skel, palettes, num_transforms, frames_dev, nullptr, nullptr, blend_type,
                                                         first_instance);
  }
}

}  // namespace p3cs
