"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper of oracle/libskin_oracle.so (oracle/skin_oracle.c), the plain-C
restatement of the reference's CPU vertex animation.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libskin_oracle.so")
_lib = None


class _Column(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("array", "start", "num_values", "numeric_type", "contents")]


class _Morph(C.Structure):
    _fields_ = [("slider", C.c_int32), ("base", _Column), ("delta", _Column),
                ("num_row_ranges", C.c_int32), ("row_ranges", C.POINTER(C.c_int32))]


class _Mesh(C.Structure):
    _fields_ = [
        ("num_rows", C.c_int32), ("coordinate_system", C.c_int32), ("num_arrays", C.c_int32),
        ("array_data", C.POINTER(C.c_void_p)), ("array_stride", C.POINTER(C.c_int32)),
        ("array_is_output", C.POINTER(C.c_int32)),
        ("num_skin_columns", C.c_int32), ("skin_columns", C.POINTER(_Column)),
        ("blend_column", _Column),
        ("num_transforms", C.c_int32), ("num_blends", C.c_int32),
        ("blend_offsets", C.POINTER(C.c_int32)), ("blend_transform", C.POINTER(C.c_int32)),
        ("blend_weight", C.POINTER(C.c_float)),
        ("num_row_ranges", C.c_int32), ("row_ranges", C.POINTER(C.c_int32)),
        ("num_sliders", C.c_int32), ("num_morphs", C.c_int32), ("morphs", C.POINTER(_Morph)),
    ]


def build(force: bool = False) -> str:
    """Compiles the oracle (gcc, oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "skin_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libskin_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.p3o_animate.restype = C.c_int
        _lib.p3o_normal_xform.restype = C.c_int
        _lib.p3o_invert4.restype = C.c_int
        _lib.p3o_invert3.restype = C.c_int
        _lib.p3o_decompose3.restype = C.c_int
    return _lib


def _col(c) -> _Column:
    return _Column(*c.as_tuple()) if c is not None else _Column(-1, 0, 0, 0, 0)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class OracleMesh:
    """Keeps the ctypes view of a FlatMesh alive."""

    def __init__(self, mesh):
        self.mesh = mesh
        self._keep = []
        m = _Mesh()
        m.num_rows = mesh.num_rows
        m.coordinate_system = mesh.coordinate_system
        na = len(mesh.arrays)
        m.num_arrays = na
        self._arrays = [np.ascontiguousarray(a) for a in mesh.arrays]
        ptrs = (C.c_void_p * na)(*[a.ctypes.data for a in self._arrays])
        strides = np.asarray(mesh.strides, np.int32)
        is_out = np.asarray([1 if o else 0 for o in mesh.is_output], np.int32)
        cols = (_Column * max(1, len(mesh.skin_columns)))(*[_col(c) for c in mesh.skin_columns])
        self._keep += [ptrs, strides, is_out, cols]
        m.array_data = C.cast(ptrs, C.POINTER(C.c_void_p))
        m.array_stride = _ip(strides)
        m.array_is_output = _ip(is_out)
        m.num_skin_columns = len(mesh.skin_columns)
        m.skin_columns = C.cast(cols, C.POINTER(_Column))
        m.blend_column = _col(mesh.blend_column)
        m.num_transforms = mesh.num_transforms
        m.num_blends = mesh.num_blends
        bo = np.ascontiguousarray(mesh.blend_offsets, np.int32)
        bx = np.ascontiguousarray(mesh.blend_transform, np.int32)
        bw = np.ascontiguousarray(mesh.blend_weight, np.float32)
        rr = np.ascontiguousarray(mesh.row_ranges, np.int32).reshape(-1)
        self._keep += [bo, bx, bw, rr]
        m.blend_offsets, m.blend_transform = _ip(bo), _ip(bx)
        m.blend_weight = bw.ctypes.data_as(C.POINTER(C.c_float))
        m.num_row_ranges = rr.size // 2
        m.row_ranges = _ip(rr)
        m.num_sliders = mesh.num_sliders
        morphs = (_Morph * max(1, len(mesh.morphs)))()
        for i, mo in enumerate(mesh.morphs):
            r = np.ascontiguousarray(mo.row_ranges, np.int32).reshape(-1)
            self._keep.append(r)
            morphs[i] = _Morph(mo.slider, _col(mo.base), _col(mo.delta), r.size // 2, _ip(r))
        self._keep.append(morphs)
        m.num_morphs = len(mesh.morphs)
        m.morphs = C.cast(morphs, C.POINTER(_Morph))
        self.c = m

    def animate(self, palette: np.ndarray, sliders: Optional[np.ndarray] = None,
                want_blends: bool = False, want_selection: bool = False):
        """Returns (outputs, blends, selection): the animated output arrays
        (uint8 [num_rows, stride] each), the blended matrices and the per-row selected blend."""
        mesh = self.mesh
        pal = np.ascontiguousarray(palette, np.float32).reshape(-1)
        sl = np.ascontiguousarray(sliders if sliders is not None else np.zeros(max(1, mesh.num_sliders)), np.float32)
        outs = [np.empty((mesh.num_rows, mesh.strides[a]), np.uint8) for a in mesh.output_arrays]
        optrs = (C.c_void_p * max(1, len(outs)))(*[o.ctypes.data for o in outs])
        blends = np.empty((mesh.num_blends, 16), np.float32) if want_blends else None
        sel = np.empty(mesh.num_rows, np.int32) if want_selection else None
        rc = lib().p3o_animate(
            C.byref(self.c), pal.ctypes.data_as(C.c_void_p), sl.ctypes.data_as(C.c_void_p), optrs,
            blends.ctypes.data_as(C.c_void_p) if blends is not None else None,
            sel.ctypes.data_as(C.c_void_p) if sel is not None else None)
        if rc != 0:
            raise RuntimeError(f"oracle: unsupported input (rc={rc})")
        return outs, blends, sel


def unexplained_differences(mesh, got, palette, sliders=None, om: Optional["OracleMesh"] = None):
    """Bytes of `got` (output arrays of one instance) that differ from the oracle's, split into `explained` -- inside a
    C_normal column of a row whose blended matrix takes the uniform-scale branch (decompose_matrix / compose_matrix,
    geomVertexData.cxx:1825-1827), where libm's and CUDA's sinf / cosf / atan2f may differ in the last place -- and
    `unexplained` (anything else: must be 0).  Returns (explained, unexplained, expect)."""
    from panda3d_b200.flat import C_normal, NUMERIC_BYTES
    om = om or OracleMesh(mesh)
    expect, blends, sel = om.animate(palette, sliders, want_blends=True, want_selection=True)
    explained = unexplained = 0
    for o, (g, e) in enumerate(zip(got, expect)):
        g = np.asarray(g).reshape(mesh.num_rows, -1)
        diff = g != e
        if not diff.any():
            continue
        normal_bytes = np.zeros(g.shape[1], bool)
        for c in mesh.skin_columns:
            if c.array == mesh.output_arrays[o] and c.contents == C_normal:
                normal_bytes[c.start:c.start + c.num_bytes] = True
        unexplained += int(diff[:, ~normal_bytes].sum())
        nd = diff[:, normal_bytes]
        rows = np.nonzero(nd.any(axis=1))[0]
        libm_blend = {}
        for r in rows:
            b = int(sel[r])
            if b not in libm_blend:
                libm_blend[b] = b >= 0 and normal_xform(blends[b], mesh.coordinate_system)[0] == 1
            if libm_blend[b]:
                explained += int(nd[r].sum())
            else:
                unexplained += int(nd[r].sum())
    return explained, unexplained, expect


def animate(mesh, palette, sliders=None, want_blends=False, want_selection=False):
    return OracleMesh(mesh).animate(palette, sliders, want_blends, want_selection)


def normal_xform(mat16: np.ndarray, cs: int = 1) -> Tuple[int, np.ndarray, bool]:
    m = np.ascontiguousarray(mat16, np.float32).reshape(16)
    out = np.empty(16, np.float32)
    norm = C.c_int(0)
    branch = lib().p3o_normal_xform(m.ctypes.data_as(C.c_void_p), C.c_int(cs), out.ctypes.data_as(C.c_void_p), C.byref(norm))
    return branch, out.reshape(4, 4), bool(norm.value)


def invert4(mat16: np.ndarray) -> Tuple[bool, np.ndarray]:
    m = np.ascontiguousarray(mat16, np.float32).reshape(16)
    out = np.empty(16, np.float32)
    ok = lib().p3o_invert4(m.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return bool(ok), out.reshape(4, 4)


def decompose3(mat9: np.ndarray, cs: int = 1):
    m = np.ascontiguousarray(mat9, np.float32).reshape(9)
    scale, shear, hpr = (np.empty(3, np.float32) for _ in range(3))
    ok = lib().p3o_decompose3(m.ctypes.data_as(C.c_void_p), C.c_int(cs), scale.ctypes.data_as(C.c_void_p),
                              shear.ctypes.data_as(C.c_void_p), hpr.ctypes.data_as(C.c_void_p))
    return bool(ok), scale, shear, hpr


def compose3(scale, shear, hpr, cs: int = 1) -> np.ndarray:
    out = np.empty(9, np.float32)
    a = [np.ascontiguousarray(x, np.float32) for x in (scale, shear, hpr)]
    lib().p3o_compose3(out.ctypes.data_as(C.c_void_p), a[0].ctypes.data_as(C.c_void_p),
                       a[1].ctypes.data_as(C.c_void_p), a[2].ctypes.data_as(C.c_void_p), C.c_int(cs))
    return out.reshape(3, 3)


# ---- the real reference, when oracle/_ref was built here (oracle/build_ref.sh) ----

REF_HARNESS = os.path.join(_HERE, "_ref", "bin", "ref_harness")


def have_ref() -> bool:
    return os.path.exists(REF_HARNESS)


def ref_run(recipe_path: str, golden_path: str) -> None:
    """Runs the UNMODIFIED reference's animate_vertices on a recipe (oracle/ref_harness.cxx)."""
    subprocess.run([REF_HARNESS, "run", recipe_path, golden_path], check=True, capture_output=True)


class _Skeleton(C.Structure):
    _fields_ = [("num_joints", C.c_int32), ("coordinate_system", C.c_int32),
                ("parent", C.c_void_p), ("has_channel", C.c_void_p), ("table_length", C.c_void_p),
                ("table_offset", C.c_void_p), ("tables", C.c_void_p), ("default_value", C.c_void_p),
                ("initial_inverse", C.c_void_p), ("root_xform", C.c_void_p),
                ("num_sliders", C.c_int32), ("slider_has_channel", C.c_void_p), ("slider_table_length", C.c_void_p),
                ("slider_table_offset", C.c_void_p), ("slider_default", C.c_void_p)]


def _skeleton_struct(skel, keep: list) -> "_Skeleton":
    """ctypes view of a FlatSkeleton; `keep` receives the arrays that must outlive the call."""
    arrs = [np.ascontiguousarray(a) for a in (skel.parent, skel.has_channel, skel.table_length, skel.table_offset,
                                              skel.tables if skel.tables.size else np.zeros(1, np.float32),
                                              skel.default_value, skel.initial_inverse, skel.root_xform)]
    S = skel.num_sliders
    sl = [np.ascontiguousarray(a) if S else np.zeros(1, dt) for a, dt in (
        (skel.slider_has_channel, np.int32), (skel.slider_table_length, np.int32), (skel.slider_table_offset, np.int32),
        (skel.slider_default, np.float32))]
    keep.append(arrs + sl)
    return _Skeleton(skel.num_joints, skel.coordinate_system, *[a.ctypes.data for a in arrs], S, *[a.ctypes.data for a in sl])


def pose(skel, frame: int, next_frame: Optional[int] = None, frac: float = 0.0, blend_type: int = 1) -> np.ndarray:
    """CharacterJoint::_skinning_matrix of every joint at `frame` ([J,16]); palette = result[skel.palette_joint].
    With `next_frame` the bundle's frame-blend flag is on: values blend towards next_frame by `frac`
    (blend_type 0 = BT_linear, 1 = BT_normalized_linear, the reference's default)."""
    keep: list = []
    s = _skeleton_struct(skel, keep)
    out = np.empty((skel.num_joints, 16), np.float32)
    lib().p3o_pose_blend.restype = C.c_int
    rc = lib().p3o_pose_blend(C.byref(s), C.c_int(int(frame)), C.c_int(int(frame if next_frame is None else next_frame)),
                              C.c_float(float(frac)), C.c_int(0 if next_frame is None else 1), C.c_int(int(blend_type)), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle pose failed rc={rc}")
    return out


def pose_controls(skels, frames, next_frames, fracs, effects, frame_blend: bool, blend_type: int = 1,
                  forced_joints=None, forced_values=None) -> np.ndarray:
    """Several AnimControls blended (p3o_pose_controls): skels[k] carries control k's channel tables.
    forced_joints / forced_values ([F], [F,16]): joints with a forced channel (freeze_joint / control_joint) and its value."""
    keep: list = []
    structs = [_skeleton_struct(skel, keep) for skel in skels]
    K = len(skels)
    ptrs = (C.POINTER(_Skeleton) * K)(*[C.pointer(s) for s in structs])
    f = np.ascontiguousarray(frames, np.int32)
    nf = np.ascontiguousarray(next_frames, np.int32)
    fr = np.ascontiguousarray(fracs, np.float32)
    ef = np.ascontiguousarray(effects, np.float32)
    out = np.empty((skels[0].num_joints, 16), np.float32)
    fj = np.ascontiguousarray(forced_joints if forced_joints is not None else [], np.int32).reshape(-1)
    fv = np.ascontiguousarray(forced_values if forced_values is not None else [], np.float32).reshape(-1, 16)
    assert fv.shape[0] == fj.size
    lib().p3o_pose_controls_forced.restype = C.c_int
    rc = lib().p3o_pose_controls_forced(ptrs, C.c_int(K), f.ctypes.data_as(C.c_void_p), nf.ctypes.data_as(C.c_void_p),
                                        fr.ctypes.data_as(C.c_void_p), ef.ctypes.data_as(C.c_void_p), C.c_int(1 if frame_blend else 0),
                                        C.c_int(int(blend_type)), C.c_int(int(fj.size)), fj.ctypes.data_as(C.c_void_p),
                                        fv.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle pose_controls failed rc={rc}")
    return out


def pose_sliders(skels, frames, next_frames, fracs, effects, frame_blend: bool) -> np.ndarray:
    """Slider values of the same controls (p3o_pose_sliders): [S], SliderTable order."""
    keep: list = []
    structs = [_skeleton_struct(skel, keep) for skel in skels]
    K = len(skels)
    ptrs = (C.POINTER(_Skeleton) * K)(*[C.pointer(s) for s in structs])
    f = np.ascontiguousarray(frames, np.int32)
    nf = np.ascontiguousarray(next_frames, np.int32)
    fr = np.ascontiguousarray(fracs, np.float32)
    ef = np.ascontiguousarray(effects, np.float32)
    out = np.empty(max(1, skels[0].num_sliders), np.float32)
    lib().p3o_pose_sliders.restype = C.c_int
    lib().p3o_pose_sliders(ptrs, C.c_int(K), f.ctypes.data_as(C.c_void_p), nf.ctypes.data_as(C.c_void_p), fr.ctypes.data_as(C.c_void_p),
                           ef.ctypes.data_as(C.c_void_p), C.c_int(1 if frame_blend else 0), out.ctypes.data_as(C.c_void_p))
    return out[:skels[0].num_sliders]
