// p3cs_pose.cu -- "next" row f1: the joint hierarchy on the GPU.
//
// One CTA per Character instance evaluates, for the instance's animation frame,
//   AnimChannelMatrixXfmTable::get_value  (panda/src/chan/animChannelMatrixXfmTable.cxx:112-124:
//                                          12 table components -> compose_matrix),
//   CharacterJoint::update_internals      (panda/src/char/characterJoint.cxx:110-150:
//                                          net = value * parent net, skinning = initial^-1 * net)
// (local values for all joints in parallel, then every joint's net transform as the product down its own chain of
// ancestors, outermost first -- the reference's products in the reference's order, no barrier per tree level) and writes
// the palette straight into the group's device buffer: with it a crowd needs one int per
// instance and frame from the host instead of T x 64 bytes of matrices.
// Covers one AnimControl without blending (chan/movingPartMatrix.cxx:71-76), the frame-blend flag
// (PartBundle::set_frame_blend_flag) and several AnimControls blended by their effects (set_anim_blend_flag /
// set_control_effect), for all four blend types (:78-357); same -fmad=false arithmetic as the rest
// of the library.
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

// frame % len as the reference evaluates it on AnimControl's non-negative frames; a negative frame handed in by a
// careless host wraps into the table instead of reading in front of it
__device__ __forceinline__ int wrap_frame(int frame, int len) {
  const int r = frame % len;
  return r < 0 ? r + len : r;
}

// One table component of a joint at a frame: table[frame % size], or the default of an empty table
// (1 for the scales, 0 otherwise; animChannelMatrixXfmTable.I:75-85).
__device__ __forceinline__ float channel_component(const AnimationDev &an, int j, int i, int frame) {
  const int len = __ldg(an.table_length + j * 12 + i);
  return len == 0 ? (i < 3 ? 1.0f : 0.0f) : __ldg(an.tables + __ldg(an.table_offset + j * 12 + i) + wrap_frame(frame, len));
}

// AnimChannelMatrixXfmTable::get_value (animChannelMatrixXfmTable.cxx:112-124), or get_value_no_scale_shear
// (:131-149) with unit scale and zero shear: compose_matrix of the components -> LMatrix4f(upper3, translate).
__device__ __forceinline__ void channel_value(const AnimationDev &an, int cs, int j, int frame, bool no_scale_shear, float *value) {
  float comp[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) comp[i] = (no_scale_shear && i < 6) ? (i < 3 ? 1.0f : 0.0f) : channel_component(an, j, i, frame);
  Mat3 up3;
  compose3(up3, comp + 0, comp + 3, comp + 6, cs);
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

// ---- BT_componentwise_quat (chan/movingPartMatrix.cxx:273-357): the rotation blends as a quaternion
// LQuaternionf::multiply (linmath/lquaternion_src.I:78-85): a = *this, b = rhs; components (r, i, j, k)
__device__ __forceinline__ void quat_mul(float *res, const float *a, const float *b) {
  const float r = (b[0] * a[0]) - (b[1] * a[1]) - (b[2] * a[2]) - (b[3] * a[3]);
  const float i = (b[1] * a[0]) + (b[0] * a[1]) - (b[3] * a[2]) + (b[2] * a[3]);
  const float j = (b[2] * a[0]) + (b[3] * a[1]) + (b[0] * a[2]) - (b[1] * a[3]);
  const float k = (b[3] * a[0]) - (b[2] * a[1]) + (b[1] * a[2]) + (b[0] * a[3]);
  res[0] = r; res[1] = i; res[2] = j; res[3] = k;
}

// LQuaternionf::set_hpr (linmath/lquaternion_src.cxx:95-120), what AnimChannelMatrixXfmTable::get_quat
// (animChannelMatrixXfmTable.cxx:186-197) returns: half-angle quaternions about up / right / forward, multiplied
// as quat_r * quat_p * quat_h (right-handed) or invert(quat_h * quat_p * quat_r) (left-handed; invert_from negates
// the real part, lquaternion_src.I:474-478).
static __device__ __noinline__ void quat_set_hpr(float *q, const float *hpr, int cs) {
  const bool left = (cs == 3 || cs == 4), zup = (cs == 1 || cs == 3);
  float fwd[3] = {0.0f, 0.0f, 0.0f}, up[3] = {0.0f, 0.0f, 0.0f};
  if (zup) { fwd[1] = left ? -1.0f : 1.0f; up[2] = 1.0f; }
  else { fwd[2] = left ? 1.0f : -1.0f; up[1] = 1.0f; }
  float qh[4], qp[4], qr[4], t[4], sn, cn;
  sincosf((hpr[0] * 0.5f) * P3CS_DEG_2_RAD, &sn, &cn);
  qh[0] = cn; qh[1] = up[0] * sn; qh[2] = up[1] * sn; qh[3] = up[2] * sn;
  sincosf((hpr[1] * 0.5f) * P3CS_DEG_2_RAD, &sn, &cn);
  qp[0] = cn; qp[1] = 1.0f * sn; qp[2] = 0.0f * sn; qp[3] = 0.0f * sn;
  sincosf((hpr[2] * 0.5f) * P3CS_DEG_2_RAD, &sn, &cn);
  qr[0] = cn; qr[1] = fwd[0] * sn; qr[2] = fwd[1] * sn; qr[3] = fwd[2] * sn;
  if (!left) {
    quat_mul(t, qr, qp);
    quat_mul(q, t, qh);
  } else {
    quat_mul(t, qh, qp);
    quat_mul(q, t, qr);
    q[0] = -q[0];
  }
}

// _value = LMatrix4f::scale_shear_mat(scale, shear) * quat; _value.set_row(3, pos)  (movingPartMatrix.cxx:352-353):
// LQuaternionf::extract_to_matrix (lquaternion_src.cxx:74-89), the full 4x4 product of operator * (LMatrix4f,
// LQuaternionf) (lquaternion_src.I:547-561) with the left matrix' row 3 and column 3 put back, then the translation.
static __device__ __noinline__ void quat_compose_value(float *value, const float *scale, const float *shear, const float *q, const float *pos,
                                                       int cs) {
  const float zero_hpr[3] = {0.0f, 0.0f, 0.0f};
  Mat3 s3;
  compose3(s3, scale, shear, zero_hpr, cs);  // zero angles skip every rotation: set_scale_shear_mat alone
  float m[16], qm[16];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) m[r * 4 + c] = s3.m[r][c];
    m[r * 4 + 3] = 0.0f;
  }
  m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
  const float N = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const float s = (N == 0.0f) ? 0.0f : (2.0f / N);
  const float xs = q[1] * s, ys = q[2] * s, zs = q[3] * s;
  const float wx = q[0] * xs, wy = q[0] * ys, wz = q[0] * zs;
  const float xx = q[1] * xs, xy = q[1] * ys, xz = q[1] * zs;
  const float yy = q[2] * ys, yz = q[2] * zs, zz = q[3] * zs;
  qm[0] = 1.0f - (yy + zz); qm[1] = xy + wz; qm[2] = xz - wy; qm[3] = 0.0f;
  qm[4] = xy - wz; qm[5] = 1.0f - (xx + zz); qm[6] = yz + wx; qm[7] = 0.0f;
  qm[8] = xz + wy; qm[9] = yz - wx; qm[10] = 1.0f - (xx + yy); qm[11] = 0.0f;
  qm[12] = 0.0f; qm[13] = 0.0f; qm[14] = 0.0f; qm[15] = 1.0f;
  mul4(value, m, qm);
  value[3] = 0.0f; value[7] = 0.0f; value[11] = 0.0f;  // column 3 and row 3 of scale_shear_mat
  value[12] = pos[0]; value[13] = pos[1]; value[14] = pos[2]; value[15] = 1.0f;
}

// BLEND = false: one AnimControl of animation 0 per instance, no frame blending (movingPartMatrix.cxx:71-76).
// BLEND = true: `num_controls` AnimControls per instance (PartBundle::set_anim_blend_flag / set_control_effect) in
// the order of the bundle's blend map, and / or the frame-blend flag; per joint
// (MovingPartBase::determine_effective_channels, movingPartBase.cxx:303-333; MovingPartMatrix::get_blend_value,
// movingPartMatrix.cxx:52-198): no control with a channel -> the default value; exactly one and no frame blending ->
// its value; else the blend, blend_type 0 = BT_linear (:82-122), 1 = BT_normalized_linear (:125-198, the PRC default),
// 2 = BT_componentwise (:200-271), 3 = BT_componentwise_quat (:273-357).
// AnimChannelScalarTable::get_value (chan/animChannelScalarTable.cxx:88-94)
__device__ __forceinline__ float slider_channel_value(const AnimationDev &an, int s, int frame) {
  const int len = __ldg(an.slider_table_length + s);
  return len == 0 ? 0.0f : __ldg(an.tables + __ldg(an.slider_table_offset + s) + wrap_frame(frame, len));
}

template <bool BLEND>
__global__ void __launch_bounds__(256) pose_kernel(SkeletonDev skel, float *palettes, int num_transforms, float *sliders, const int *frames,
                                                    const ControlState *states, int num_controls, int frame_blend, int blend_type,
                                                    ForcedDev forced, int first_instance) {
  extern __shared__ float smem_pose[];  // [J][16] local values, then [J][16] net transforms of this instance
  const int inst = first_instance + blockIdx.x;
  const int J = skel.num_joints;
  float *local = smem_pose;
  float *net = smem_pose + (size_t)J * 16;
  const int cs = skel.coordinate_system;
  // morph sliders: CharacterSlider's value from its scalar channel(s) (MovingPartScalar::get_blend_value,
  // chan/movingPartScalar.cxx:43-105): same selection rule as the joints; the blend is a plain weighted sum from
  // zero followed by a true division by the net effect
  for (int sl = threadIdx.x; sl < skel.num_sliders; sl += blockDim.x) {
    float value = __ldg(skel.slider_default + sl);
    if (!BLEND) {
      if (__ldg(skel.anim[0].slider_has_channel + sl)) value = slider_channel_value(skel.anim[0], sl, __ldg(frames + blockIdx.x));
    } else {
      const ControlState *st = states + (size_t)blockIdx.x * num_controls;
      int n_eff = 0, only = -1;
      for (int c = 0; c < num_controls; ++c)
        if (st[c].effect != 0.0f && __ldg(skel.anim[st[c].anim].slider_has_channel + sl)) {
          ++n_eff;
          only = c;
        }
      if (n_eff == 1 && !frame_blend) {
        value = slider_channel_value(skel.anim[st[only].anim], sl, st[only].frame);
      } else if (n_eff > 0) {
        float sum = 0.0f, net_effect = 0.0f;
        for (int c = 0; c < num_controls; ++c) {
          const ControlState cst = st[c];
          const AnimationDev &an = skel.anim[cst.anim];
          if (cst.effect == 0.0f || !__ldg(an.slider_has_channel + sl)) continue;
          if (!frame_blend) {
            sum = sum + slider_channel_value(an, sl, cst.frame) * cst.effect;
          } else {
            sum = sum + slider_channel_value(an, sl, cst.frame) * (cst.effect * (1.0f - cst.frac));
            sum = sum + slider_channel_value(an, sl, cst.next_frame) * (cst.effect * cst.frac);
          }
          net_effect = net_effect + cst.effect;
        }
        value = sum / net_effect;
      }
    }
    sliders[(size_t)inst * skel.num_sliders + sl] = value;
  }

  // phase 1, all joints in parallel: the joint's own value
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float value[16];
    bool have = false;
    // a forced channel (freeze_joint / control_joint) wins over whatever is playing (movingPartMatrix.cxx:53-60)
    const int fslot = forced.count > 0 ? __ldg(forced.slot + j) : -1;
    if (fslot >= 0) {
      const float4 *fv = reinterpret_cast<const float4 *>(forced.values + ((size_t)inst * forced.count + fslot) * 16);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 q = __ldg(fv + r);
        value[r * 4] = q.x; value[r * 4 + 1] = q.y; value[r * 4 + 2] = q.z; value[r * 4 + 3] = q.w;
      }
      have = true;
    } else if (!BLEND) {
      if (__ldg(skel.anim[0].has_channel + j)) {
        channel_value(skel.anim[0], cs, j, __ldg(frames + blockIdx.x), false, value);
        have = true;
      }
    } else {
      const ControlState *st = states + (size_t)blockIdx.x * num_controls;
      int n_eff = 0, only = -1;
      for (int c = 0; c < num_controls; ++c)
        if (st[c].effect != 0.0f && __ldg(skel.anim[st[c].anim].has_channel + j)) {
          ++n_eff;
          only = c;
        }
      if (n_eff == 1 && !frame_blend) {
        channel_value(skel.anim[st[only].anim], cs, j, st[only].frame, false, value);
        have = true;
      } else if (n_eff > 0) {
        // sums start from zero, in map order; the division by net_effect multiplies by its reciprocal
        // (LMatrix4 / LVecBase3 operator /=)
        float nv[16], scale[3] = {0.0f, 0.0f, 0.0f}, shear[3] = {0.0f, 0.0f, 0.0f}, net_effect = 0.0f;
        float comps[12];  // BT_componentwise (:200-271): scale, shear, hpr and pos blend on their own
        float quat[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // BT_componentwise_quat (:273-357): the rotation as a quaternion
#pragma unroll
        for (int k = 0; k < 16; ++k) nv[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < 12; ++k) comps[k] = 0.0f;
        for (int c = 0; c < num_controls; ++c) {
          const ControlState cst = st[c];
          const AnimationDev &an = skel.anim[cst.anim];
          if (cst.effect == 0.0f || !__ldg(an.has_channel + j)) continue;
          for (int pass = 0; pass < (frame_blend ? 2 : 1); ++pass) {
            const int fr = pass == 0 ? cst.frame : cst.next_frame;
            const float e = !frame_blend ? cst.effect : pass == 0 ? cst.effect * (1.0f - cst.frac) : cst.effect * cst.frac;
            if (blend_type == 2) {
#pragma unroll
              for (int k = 0; k < 12; ++k) comps[k] = comps[k] + channel_component(an, j, k, fr) * e;
              continue;
            }
            if (blend_type == 3) {
              float ihpr[3], iq[4];
#pragma unroll
              for (int k = 0; k < 3; ++k) ihpr[k] = channel_component(an, j, 6 + k, fr);
              quat_set_hpr(iq, ihpr, cs);
#pragma unroll
              for (int k = 0; k < 12; ++k)
                if (k < 6 || k >= 9) comps[k] = comps[k] + channel_component(an, j, k, fr) * e;
#pragma unroll
              for (int k = 0; k < 4; ++k) quat[k] = quat[k] + iq[k] * e;
              continue;
            }
            float v[16];
            channel_value(an, cs, j, fr, blend_type != 0, v);
#pragma unroll
            for (int k = 0; k < 16; ++k) nv[k] = nv[k] + v[k] * e;
            if (blend_type != 0) {
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                scale[k] = scale[k] + channel_component(an, j, k, fr) * e;
                shear[k] = shear[k] + channel_component(an, j, 3 + k, fr) * e;
              }
            }
          }
          net_effect = net_effect + cst.effect;
        }
        const float recip = 1.0f / net_effect;
#pragma unroll
        for (int k = 0; k < 16; ++k) value[k] = nv[k] * recip;
        if (blend_type == 3) {
#pragma unroll
          for (int k = 0; k < 12; ++k) comps[k] = comps[k] * recip;
#pragma unroll
          for (int k = 0; k < 4; ++k) quat[k] = quat[k] * recip;
          quat_compose_value(value, comps + 0, comps + 3, quat, comps + 9, cs);
        } else if (blend_type == 2) {
#pragma unroll
          for (int k = 0; k < 12; ++k) comps[k] = comps[k] * recip;
          Mat3 out3;
          compose3(out3, comps + 0, comps + 3, comps + 6, cs);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) value[r * 4 + c] = out3.m[r][c];
            value[r * 4 + 3] = 0.0f;
          }
          value[12] = comps[9];
          value[13] = comps[10];
          value[14] = comps[11];
          value[15] = 1.0f;
        } else if (blend_type != 0) {
          // decompose_matrix(net_value, false_scale, false_shear, hpr, translate), then
          // compose_matrix(_value, scale, shear, hpr, translate)
          float false_scale[3], false_shear[3], hpr[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            scale[k] = scale[k] * recip;
            shear[k] = shear[k] * recip;
          }
          Mat3 up3, out3;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) up3.m[r][c] = value[r * 4 + c];
          decompose3(up3, cs, false_scale, false_shear, hpr);
          compose3(out3, scale, shear, hpr, cs);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) value[r * 4 + c] = out3.m[r][c];
            value[r * 4 + 3] = 0.0f;
          }
          value[15] = 1.0f;  // translate = row 3 of the blended matrix stays
        }
        have = true;
      }
    }
    if (!have) {
#pragma unroll
      for (int k = 0; k < 16; ++k) value[k] = __ldg(skel.default_value + j * 16 + k);
    }
    float4 *dst = reinterpret_cast<float4 *>(local + j * 16);
#pragma unroll
    for (int r = 0; r < 4; ++r) dst[r] = make_float4(value[r * 4], value[r * 4 + 1], value[r * 4 + 2], value[r * 4 + 3]);
  }
  __syncthreads();

  // phase 2: _net_transform = _value * parent's _net_transform (or the root transform), CharacterJoint::update_internals.
  // Every joint multiplies down its own chain of ancestors, outermost first -- the same products in the same order as
  // the parents-first walk of the reference (each thread recomputes its ancestors' net transforms, bit for bit what
  // their own threads compute), without a CTA-wide barrier per tree level.
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float cur[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) cur[q] = __ldg(skel.root_xform + q);
    const int pb = __ldg(skel.path_offsets + j), pe = __ldg(skel.path_offsets + j + 1);
    for (int e = pb; e < pe; ++e) {
      const float4 *v4 = reinterpret_cast<const float4 *>(local + __ldg(skel.path_joints + e) * 16);
      float value[16], out[16];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 q = v4[r];
        value[r * 4] = q.x; value[r * 4 + 1] = q.y; value[r * 4 + 2] = q.z; value[r * 4 + 3] = q.w;
      }
      mul4(out, value, cur);
#pragma unroll
      for (int q = 0; q < 16; ++q) cur[q] = out[q];
    }
    float4 *dst = reinterpret_cast<float4 *>(net + j * 16);
#pragma unroll
    for (int r = 0; r < 4; ++r) dst[r] = make_float4(cur[r * 4], cur[r * 4 + 1], cur[r * 4 + 2], cur[r * 4 + 3]);
  }
  __syncthreads();
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

void launch_pose(const SkeletonDev &skel, float *palettes, int num_transforms, float *sliders, const int *frames_dev,
                 const ControlState *states_dev, int num_controls, int frame_blend, int blend_type, const ForcedDev &forced,
                 int first_instance, int count, cudaStream_t stream) {
  if (count == 0 || skel.num_joints == 0) return;
  const size_t smem = (size_t)skel.num_joints * 128;
  int threads = ((skel.num_joints + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (states_dev != nullptr) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(pose_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pose_kernel<true><<<count, threads, smem, stream>>>(skel, palettes, num_transforms, sliders, nullptr, states_dev, num_controls, frame_blend,
                                                        blend_type, forced, first_instance);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(pose_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pose_kernel<false><<<count, threads, smem, stream>>>(skel, palettes, num_transforms, sliders, frames_dev, nullptr, 0, 0, blend_type,
                                                         forced, first_instance);
  }
}

}  // namespace p3cs
