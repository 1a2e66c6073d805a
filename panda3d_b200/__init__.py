"""panda3d_b200 -- B200 (sm_100a) backend for Panda3D's CPU vertex-animation hot path.

Only what the path needs lives here:
  csrc/      hand-written CUDA kernels + the C-ABI (libp3cudaskin.so, include/p3cudaskin.h)
  _capi      ctypes binding of that C-ABI (fails loudly when the library is missing)
  flat       the flattened mesh description handed across the C-ABI
  gobj       host-side mirror of the reference's GeomVertexData / TransformBlendTable /
             SliderTable interface for this path (same names, argument meaning, errors)
  crowd      sharding of independent Character instances over GPUs + the NCCL gather
  synth      seeded synthetic skinned meshes of the BASELINE.json shapes
  casefile   the small binary container golden fixtures are stored in

There is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
