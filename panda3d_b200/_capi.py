"""ctypes binding of libp3cudaskin.so (include/p3cudaskin.h).

This is the only way the package computes anything: there is no Python / numpy / torch
implementation of the path behind it.  A missing library raises at import of the library
handle (`lib()`), a missing GPU raises P3csError(P3CS_ERR_NO_DEVICE) at Context creation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .flat import Column, FlatMesh

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libp3cudaskin.so")

P3CS_OK, P3CS_ERR_INVALID, P3CS_ERR_UNSUPPORTED, P3CS_ERR_CUDA, P3CS_ERR_NOMEM, P3CS_ERR_NO_DEVICE = range(6)

# every symbol include/p3cudaskin.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "p3cs_api_version", "p3cs_device_count", "p3cs_last_error",
    "p3cs_context_create", "p3cs_context_destroy", "p3cs_context_sync", "p3cs_context_stream",
    "p3cs_context_enable_peer_access", "p3cs_device_alloc", "p3cs_device_free", "p3cs_device_read",
    "p3cs_ipc_export", "p3cs_ipc_open", "p3cs_ipc_close",
    "p3cs_mesh_create", "p3cs_mesh_destroy", "p3cs_mesh_num_outputs", "p3cs_mesh_output_stride", "p3cs_mesh_num_rows",
    "p3cs_group_create", "p3cs_group_destroy", "p3cs_group_output_pitch", "p3cs_group_palettes_device",
    "p3cs_group_sliders_device", "p3cs_group_output_device", "p3cs_group_set_palettes", "p3cs_group_set_sliders",
    "p3cs_group_animate", "p3cs_group_read_output", "p3cs_group_read_selection", "p3cs_group_read_blends",
    "p3cs_group_set_profiling", "p3cs_group_last_timings", "p3cs_animate_vertices", "p3cs_get_api",
    "p3cs_skeleton_create", "p3cs_skeleton_destroy", "p3cs_group_pose", "p3cs_group_read_palettes",
    "p3cs_group_enable_bounds", "p3cs_group_read_bounds", "p3cs_animate_groups",
    "p3cs_host_register", "p3cs_host_unregister",
    "p3cs_shareable_alloc", "p3cs_shareable_import", "p3cs_shareable_free", "p3cs_group_animate_host",
    "p3cs_hardware_from_blends", "p3cs_blends_from_hardware", "p3cs_group_animate_instances",
    "p3cs_group_pose_blend", "p3cs_skeleton_add_animation", "p3cs_group_pose_controls",
    "p3cs_group_set_forced_joints", "p3cs_group_set_forced_values", "p3cs_group_set_bounds_rows",
    "p3cs_group_set_output_pointers",
    "p3cs_crowd_create", "p3cs_crowd_destroy", "p3cs_crowd_num_devices", "p3cs_crowd_range", "p3cs_crowd_set_palettes",
    "p3cs_crowd_set_sliders", "p3cs_crowd_animate", "p3cs_crowd_sync", "p3cs_crowd_last_ms", "p3cs_crowd_tune",
    "p3cs_crowd_output_device", "p3cs_crowd_output_pitch", "p3cs_crowd_read_output",
]


class P3csError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"p3cudaskin error {code}: {message}")
        self.code = code


class _ArrayDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("stride", C.c_int32), ("is_output", C.c_int32)]


class _ColumnDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("array", "start", "num_values", "numeric_type", "contents")]


class _MorphDesc(C.Structure):
    _fields_ = [("slider", C.c_int32), ("base", _ColumnDesc), ("delta", _ColumnDesc),
                ("num_row_ranges", C.c_int32), ("row_ranges", C.POINTER(C.c_int32))]


class _MeshDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_rows", C.c_int32), ("coordinate_system", C.c_int32),
        ("num_arrays", C.c_int32), ("arrays", C.POINTER(_ArrayDesc)),
        ("num_skin_columns", C.c_int32), ("skin_columns", C.POINTER(_ColumnDesc)),
        ("blend_column", _ColumnDesc),
        ("num_transforms", C.c_int32), ("num_blends", C.c_int32),
        ("blend_offsets", C.POINTER(C.c_int32)), ("blend_transform", C.POINTER(C.c_int32)),
        ("blend_weight", C.POINTER(C.c_float)),
        ("num_row_ranges", C.c_int32), ("row_ranges", C.POINTER(C.c_int32)),
        ("num_sliders", C.c_int32), ("num_morphs", C.c_int32), ("morphs", C.POINTER(_MorphDesc)),
    ]


class _SkeletonDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("num_joints", C.c_int32), ("coordinate_system", C.c_int32),
                ("num_palette", C.c_int32), ("parent", C.c_void_p), ("has_channel", C.c_void_p),
                ("table_length", C.c_void_p), ("table_offset", C.c_void_p), ("num_table_values", C.c_int32),
                ("tables", C.c_void_p), ("default_value", C.c_void_p), ("initial_inverse", C.c_void_p),
                ("root_xform", C.c_void_p), ("palette_joint", C.c_void_p),
                ("num_sliders", C.c_int32), ("slider_has_channel", C.c_void_p), ("slider_table_length", C.c_void_p),
                ("slider_table_offset", C.c_void_p), ("slider_default", C.c_void_p)]


class _GroupBuffers(C.Structure):
    _fields_ = [("palettes", C.c_void_p), ("sliders", C.c_void_p), ("outputs", C.POINTER(C.c_void_p))]


class Timings(C.Structure):
    _fields_ = [("blend_ms", C.c_float), ("skin_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_int32),
                ("blend_launches", C.c_int32), ("skin_launches", C.c_int32)]


_lib = None


def lib():
    """Loads libp3cudaskin.so; raises if it has not been built (python -m panda3d_b200.build_ext)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m panda3d_b200.build_ext` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.p3cs_last_error.restype = C.c_char_p
        L.p3cs_context_stream.restype = C.c_void_p
        L.p3cs_group_output_pitch.restype = C.c_size_t
        for n in ("p3cs_group_palettes_device", "p3cs_group_sliders_device", "p3cs_group_output_device"):
            getattr(L, n).restype = C.c_void_p
        L.p3cs_get_api.restype = C.c_void_p
        L.p3cs_context_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.p3cs_context_destroy.argtypes = [C.c_void_p]
        L.p3cs_context_sync.argtypes = [C.c_void_p]
        L.p3cs_context_stream.argtypes = [C.c_void_p]
        L.p3cs_context_enable_peer_access.argtypes = [C.c_void_p, C.c_int]
        L.p3cs_device_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.p3cs_device_free.argtypes = [C.c_void_p, C.c_void_p]
        L.p3cs_device_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.p3cs_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.p3cs_ipc_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.p3cs_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
        L.p3cs_mesh_create.argtypes = [C.c_void_p, C.POINTER(_MeshDesc), C.POINTER(C.c_void_p)]
        L.p3cs_mesh_destroy.argtypes = [C.c_void_p]
        L.p3cs_mesh_num_outputs.argtypes = [C.c_void_p]
        L.p3cs_mesh_output_stride.argtypes = [C.c_void_p, C.c_int]
        L.p3cs_mesh_num_rows.argtypes = [C.c_void_p]
        L.p3cs_group_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(_GroupBuffers), C.POINTER(C.c_void_p)]
        L.p3cs_group_destroy.argtypes = [C.c_void_p]
        L.p3cs_group_output_pitch.argtypes = [C.c_void_p, C.c_int]
        L.p3cs_group_palettes_device.argtypes = [C.c_void_p]
        L.p3cs_group_sliders_device.argtypes = [C.c_void_p]
        L.p3cs_group_output_device.argtypes = [C.c_void_p, C.c_int]
        L.p3cs_group_set_palettes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.p3cs_group_set_sliders.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.p3cs_group_animate.argtypes = [C.c_void_p]
        L.p3cs_group_read_output.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.p3cs_group_read_selection.argtypes = [C.c_void_p, C.c_void_p]
        L.p3cs_group_read_blends.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.p3cs_group_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.p3cs_group_last_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
        L.p3cs_animate_vertices.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.p3cs_skeleton_create.argtypes = [C.c_void_p, C.POINTER(_SkeletonDesc), C.POINTER(C.c_void_p)]
        L.p3cs_skeleton_destroy.argtypes = [C.c_void_p]
        L.p3cs_group_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.p3cs_group_read_palettes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.p3cs_group_enable_bounds.argtypes = [C.c_void_p, C.POINTER(_ColumnDesc)]
        L.p3cs_group_read_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.p3cs_animate_groups.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.p3cs_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.p3cs_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
        L.p3cs_shareable_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.p3cs_shareable_import.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
        L.p3cs_shareable_free.argtypes = [C.c_void_p, C.c_void_p]
        L.p3cs_group_animate_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != P3CS_OK:
        raise P3csError(rc, lib().p3cs_last_error().decode())


def device_count() -> int:
    return int(lib().p3cs_device_count())


def _col(c: Optional[Column]) -> _ColumnDesc:
    return _ColumnDesc(*c.as_tuple()) if c is not None else _ColumnDesc(-1, 0, 0, 0, 0)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Context:
    """p3cs_context: one CUDA device + the stream all work is ordered on."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.handle = C.c_void_p()
        _check(lib().p3cs_context_create(device, C.c_void_p(stream) if stream else None, C.byref(self.handle)))
        self.device = device

    @property
    def stream(self) -> int:
        return int(lib().p3cs_context_stream(self.handle) or 0)

    def sync(self) -> None:
        _check(lib().p3cs_context_sync(self.handle))

    def enable_peer_access(self, peer_device: int) -> None:
        _check(lib().p3cs_context_enable_peer_access(self.handle, peer_device))

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _check(lib().p3cs_device_alloc(self.handle, nbytes, C.byref(p)))
        return int(p.value)

    def device_free(self, ptr: int) -> None:
        _check(lib().p3cs_device_free(self.handle, ptr))

    def device_read(self, ptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, np.uint8)
        _check(lib().p3cs_device_read(self.handle, ptr, out.ctypes.data, nbytes))
        return out

    def ipc_export(self, ptr: int) -> bytes:
        buf = (C.c_ubyte * 64)()
        _check(lib().p3cs_ipc_export(self.handle, ptr, buf))
        return bytes(buf)

    # ---- "next" row f2: memory a renderer can import through a POSIX fd
    def shareable_alloc(self, nbytes: int):
        """Returns (device pointer, fd, allocated bytes)."""
        p, fd, size = C.c_void_p(), C.c_int(-1), C.c_size_t(0)
        _check(lib().p3cs_shareable_alloc(self.handle, nbytes, C.byref(p), C.byref(fd), C.byref(size)))
        return int(p.value), int(fd.value), int(size.value)

    def shareable_import(self, fd: int, allocated_bytes: int) -> int:
        p = C.c_void_p()
        _check(lib().p3cs_shareable_import(self.handle, fd, allocated_bytes, C.byref(p)))
        return int(p.value)

    def shareable_free(self, ptr: int) -> None:
        _check(lib().p3cs_shareable_free(self.handle, ptr))

    def host_register(self, array: np.ndarray) -> None:
        """Page-locks + maps a numpy buffer so p3cs_animate_vertices can store into it directly."""
        _check(lib().p3cs_host_register(self.handle, array.ctypes.data, array.nbytes))

    def host_unregister(self, array: np.ndarray) -> None:
        _check(lib().p3cs_host_unregister(self.handle, array.ctypes.data))

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(lib().p3cs_ipc_open(self.handle, buf, C.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int) -> None:
        _check(lib().p3cs_ipc_close(self.handle, ptr))

    def close(self) -> None:
        if self.handle:
            lib().p3cs_context_destroy(self.handle)
            self.handle = C.c_void_p()


def mesh_desc(mesh: FlatMesh, keep: list) -> _MeshDesc:
    """Fills a p3cs_mesh_desc from a FlatMesh; `keep` receives the buffers that must outlive the call."""
    d = _MeshDesc()
    d.struct_size = C.sizeof(_MeshDesc)
    d.num_rows = mesh.num_rows
    d.coordinate_system = mesh.coordinate_system
    arrays = [np.ascontiguousarray(a) for a in mesh.arrays]
    ad = (_ArrayDesc * max(1, len(arrays)))(*[
        _ArrayDesc(a.ctypes.data, int(a.shape[1]), 1 if o else 0) for a, o in zip(arrays, mesh.is_output)])
    cols = (_ColumnDesc * max(1, len(mesh.skin_columns)))(*[_col(c) for c in mesh.skin_columns])
    bo = np.ascontiguousarray(mesh.blend_offsets, np.int32)
    bx = np.ascontiguousarray(mesh.blend_transform, np.int32)
    bw = np.ascontiguousarray(mesh.blend_weight, np.float32)
    rr = np.ascontiguousarray(mesh.row_ranges, np.int32).reshape(-1)
    morphs = (_MorphDesc * max(1, len(mesh.morphs)))()
    for i, mo in enumerate(mesh.morphs):
        r = np.ascontiguousarray(mo.row_ranges, np.int32).reshape(-1)
        keep.append(r)
        morphs[i] = _MorphDesc(mo.slider, _col(mo.base), _col(mo.delta), r.size // 2, _ip(r))
    keep += [arrays, ad, cols, bo, bx, bw, rr, morphs]
    d.num_arrays = len(arrays)
    d.arrays = C.cast(ad, C.POINTER(_ArrayDesc))
    d.num_skin_columns = len(mesh.skin_columns)
    d.skin_columns = C.cast(cols, C.POINTER(_ColumnDesc))
    d.blend_column = _col(mesh.blend_column)
    d.num_transforms = mesh.num_transforms
    d.num_blends = mesh.num_blends
    d.blend_offsets, d.blend_transform = _ip(bo), _ip(bx)
    d.blend_weight = bw.ctypes.data_as(C.POINTER(C.c_float))
    d.num_row_ranges = rr.size // 2
    d.row_ranges = _ip(rr)
    d.num_sliders = mesh.num_sliders
    d.num_morphs = len(mesh.morphs)
    d.morphs = C.cast(morphs, C.POINTER(_MorphDesc))
    return d


def hardware_from_blends(mesh: FlatMesh):
    """p3cs_hardware_from_blends: (weights [rows,4] f32, indices [rows,4] i32, table [n] i32) -- what
    GeomVertexData::copy_from writes when converting to an indexed AT_hardware format."""
    keep: list = []
    d = mesh_desc(mesh, keep)
    w = np.empty((mesh.num_rows, 4), np.float32)
    ix = np.empty((mesh.num_rows, 4), np.int32)
    table = np.empty(max(1, mesh.num_transforms), np.int32)
    n = C.c_int32(0)
    L = lib()
    L.p3cs_hardware_from_blends.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    rc = L.p3cs_hardware_from_blends(C.byref(d), w.ctypes.data, ix.ctypes.data, table.ctypes.data, C.byref(n))
    if rc != P3CS_OK:
        raise P3csError(rc, "p3cs_hardware_from_blends failed")
    return w, ix, table[:n.value].copy()


def blends_from_hardware(weights: np.ndarray, indices: np.ndarray, table: np.ndarray):
    """p3cs_blends_from_hardware: (row_blend [rows] i32, blend_offsets, blend_transform, blend_weight)."""
    w = np.ascontiguousarray(weights, np.float32).reshape(-1, 4)
    ix = np.ascontiguousarray(indices, np.int32).reshape(-1, 4)
    tb = np.ascontiguousarray(table, np.int32)
    rows = w.shape[0]
    row_blend = np.empty(rows, np.int32)
    off = np.empty(rows + 1, np.int32)
    xf = np.empty(4 * max(rows, 1), np.int32)
    bw = np.empty(4 * max(rows, 1), np.float32)
    nb = C.c_int32(0)
    L = lib()
    L.p3cs_blends_from_hardware.argtypes = [C.c_int32] + [C.c_void_p] * 3 + [C.c_int32] + [C.c_void_p] * 4 + [C.POINTER(C.c_int32)]
    rc = L.p3cs_blends_from_hardware(rows, w.ctypes.data, ix.ctypes.data, tb.ctypes.data, int(tb.size), row_blend.ctypes.data,
                                     off.ctypes.data, xf.ctypes.data, bw.ctypes.data, C.byref(nb))
    if rc != P3CS_OK:
        raise P3csError(rc, "p3cs_blends_from_hardware failed")
    n = nb.value
    return row_blend, off[:n + 1].copy(), xf[:off[n]].copy(), bw[:off[n]].copy()


class Mesh:
    """p3cs_mesh: the static half of a GeomVertexData, resident in HBM."""

    def __init__(self, ctx: Context, mesh: FlatMesh):
        self.ctx = ctx
        self.flat = mesh
        keep: list = []
        d = mesh_desc(mesh, keep)
        self.handle = C.c_void_p()
        _check(lib().p3cs_mesh_create(ctx.handle, C.byref(d), C.byref(self.handle)))
        self.num_outputs = int(lib().p3cs_mesh_num_outputs(self.handle))
        self.out_strides = [int(lib().p3cs_mesh_output_stride(self.handle, o)) for o in range(self.num_outputs)]
        self.num_rows = int(lib().p3cs_mesh_num_rows(self.handle))

    def close(self) -> None:
        if self.handle:
            lib().p3cs_mesh_destroy(self.handle)
            self.handle = C.c_void_p()


class Skeleton:
    """p3cs_skeleton: the joint hierarchy + animation tables of a Character, resident in HBM."""

    def __init__(self, ctx: Context, skel):
        self.ctx, self.flat = ctx, skel
        a = [np.ascontiguousarray(x) for x in (
            skel.parent.astype(np.int32), skel.has_channel.astype(np.int32), skel.table_length.astype(np.int32),
            skel.table_offset.astype(np.int32), skel.tables.astype(np.float32) if skel.tables.size else np.zeros(1, np.float32),
            skel.default_value.astype(np.float32), skel.initial_inverse.astype(np.float32),
            skel.root_xform.astype(np.float32), skel.palette_joint.astype(np.int32))]
        S = getattr(skel, "num_sliders", 0)
        sl = [np.ascontiguousarray(x) for x in (
            skel.slider_has_channel.astype(np.int32), skel.slider_table_length.astype(np.int32),
            skel.slider_table_offset.astype(np.int32), skel.slider_default.astype(np.float32))] if S else []
        self._keep = a + sl
        d = _SkeletonDesc(C.sizeof(_SkeletonDesc), skel.num_joints, skel.coordinate_system, len(skel.palette_joint),
                          a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, int(skel.tables.size),
                          a[4].ctypes.data, a[5].ctypes.data, a[6].ctypes.data, a[7].ctypes.data, a[8].ctypes.data,
                          S, *([x.ctypes.data for x in sl] if S else [None] * 4))
        self.handle = C.c_void_p()
        _check(lib().p3cs_skeleton_create(ctx.handle, C.byref(d), C.byref(self.handle)))

    def add_animation(self, skel) -> int:
        """p3cs_skeleton_add_animation: the channel tables of another animation bound to the same joints (a
        FlatSkeleton whose table fields describe it); returns its index for p3cs_control_state::anim."""
        a = [np.ascontiguousarray(x) for x in (
            skel.has_channel.astype(np.int32), skel.table_length.astype(np.int32), skel.table_offset.astype(np.int32),
            skel.tables.astype(np.float32) if skel.tables.size else np.zeros(1, np.float32))]

        class _AnimDesc(C.Structure):
            _fields_ = [("struct_size", C.c_uint32), ("num_table_values", C.c_int32), ("has_channel", C.c_void_p),
                        ("table_length", C.c_void_p), ("table_offset", C.c_void_p), ("tables", C.c_void_p),
                        ("slider_has_channel", C.c_void_p), ("slider_table_length", C.c_void_p), ("slider_table_offset", C.c_void_p)]
        S = getattr(skel, "num_sliders", 0)
        sl = [np.ascontiguousarray(x) for x in (skel.slider_has_channel.astype(np.int32), skel.slider_table_length.astype(np.int32),
                                                skel.slider_table_offset.astype(np.int32))] if S else []
        d = _AnimDesc(C.sizeof(_AnimDesc), int(skel.tables.size), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data,
                      *([x.ctypes.data for x in sl] if S else [None] * 3))
        idx = C.c_int32(-1)
        L = lib()
        L.p3cs_skeleton_add_animation.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        _check(L.p3cs_skeleton_add_animation(self.handle, C.byref(d), C.byref(idx)))
        return int(idx.value)

    def close(self) -> None:
        if self.handle:
            lib().p3cs_skeleton_destroy(self.handle)
            self.handle = C.c_void_p()


class Group:
    """p3cs_group: n animated instances (a crowd of Characters) of one mesh."""

    def __init__(self, mesh: Mesh, num_instances: int = 1, palettes_dev: int = 0, sliders_dev: int = 0,
                 outputs_dev: Optional[Sequence[int]] = None):
        self.mesh = mesh
        self.n = num_instances
        self.handle = C.c_void_p()
        bufs = None
        if palettes_dev or sliders_dev or outputs_dev:
            outs = (C.c_void_p * max(1, mesh.num_outputs))(*[(outputs_dev[o] if outputs_dev else 0) or None
                                                             for o in range(mesh.num_outputs)])
            self._outs = outs
            bufs = _GroupBuffers(palettes_dev or None, sliders_dev or None, C.cast(outs, C.POINTER(C.c_void_p)))
        _check(lib().p3cs_group_create(mesh.handle, num_instances, C.byref(bufs) if bufs else None, C.byref(self.handle)))

    # ---- device pointers / geometry
    def output_pitch(self, o: int = 0) -> int:
        return int(lib().p3cs_group_output_pitch(self.handle, o))

    @property
    def palettes_device(self) -> int:
        return int(lib().p3cs_group_palettes_device(self.handle))

    @property
    def sliders_device(self) -> int:
        return int(lib().p3cs_group_sliders_device(self.handle))

    def output_device(self, o: int = 0) -> int:
        return int(lib().p3cs_group_output_device(self.handle, o))

    # ---- per-frame inputs
    def set_palettes(self, matrices: np.ndarray, first: int = 0) -> None:
        m = np.ascontiguousarray(matrices, np.float32)
        t16 = self.mesh.flat.num_transforms * 16
        count = m.size // t16 if t16 else 0
        _check(lib().p3cs_group_set_palettes(self.handle, first, count, m.ctypes.data))

    def set_palettes_ptr(self, host_ptr: int, first: int, count: int) -> None:
        _check(lib().p3cs_group_set_palettes(self.handle, first, count, host_ptr))

    def set_sliders(self, values: np.ndarray, first: int = 0) -> None:
        v = np.ascontiguousarray(values, np.float32)
        s = self.mesh.flat.num_sliders
        count = v.size // s if s else 0
        _check(lib().p3cs_group_set_sliders(self.handle, first, count, v.ctypes.data))

    def pose(self, skeleton: "Skeleton", frames, first: int = 0) -> None:
        """Joint hierarchy on the GPU: one animation frame per instance in, palettes written on the device."""
        f = np.ascontiguousarray(frames, np.int32)
        _check(lib().p3cs_group_pose(self.handle, skeleton.handle, first, int(f.size), f.ctypes.data))

    def pose_blend(self, skeleton, frames, next_frames, fracs, blend_type: int = 1, first: int = 0) -> None:
        """p3cs_group_pose_blend: frame blending on (blend_type 0 BT_linear, 1 BT_normalized_linear)."""
        f = np.ascontiguousarray(frames, np.int32)
        nf = np.ascontiguousarray(next_frames, np.int32)
        fr = np.ascontiguousarray(fracs, np.float32)
        L = lib()
        L.p3cs_group_pose_blend.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _check(L.p3cs_group_pose_blend(self.handle, skeleton.handle, first, int(f.size), f.ctypes.data, nf.ctypes.data, fr.ctypes.data,
                                       int(blend_type)))

    def pose_controls(self, skeleton, states: np.ndarray, frame_blend: bool, blend_type: int = 1, first: int = 0) -> None:
        """p3cs_group_pose_controls: `states` = [count, K, 5] (anim, frame, next_frame, frac, effect) per control."""
        st = np.asarray(states, np.float64)
        count, K = st.shape[0], st.shape[1]
        packed = np.zeros((count, K), np.dtype([("anim", np.int32), ("frame", np.int32), ("next_frame", np.int32),
                                                  ("frac", np.float32), ("effect", np.float32)]))
        for i, name in enumerate(("anim", "frame", "next_frame", "frac", "effect")):
            packed[name] = st[:, :, i]
        L = lib()
        L.p3cs_group_pose_controls.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _check(L.p3cs_group_pose_controls(self.handle, skeleton.handle, first, count, K, packed.ctypes.data,
                                          1 if frame_blend else 0, int(blend_type)))

    def set_forced_joints(self, skeleton, joints) -> None:
        """p3cs_group_set_forced_joints: PartBundle::freeze_joint / control_joint on these joints (empty = release all)."""
        j = np.ascontiguousarray(joints, np.int32).reshape(-1)
        L = lib()
        L.p3cs_group_set_forced_joints.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _check(L.p3cs_group_set_forced_joints(self.handle, skeleton.handle, int(j.size), j.ctypes.data if j.size else None))

    def set_forced_values(self, values, first: int = 0) -> None:
        """p3cs_group_set_forced_values: [count, num_forced, 16] matrices the forced channels return."""
        v = np.ascontiguousarray(values, np.float32)
        L = lib()
        L.p3cs_group_set_forced_values.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _check(L.p3cs_group_set_forced_values(self.handle, first, int(v.shape[0]), v.ctypes.data))

    def read_palettes(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.empty((count, self.mesh.flat.num_transforms, 16), np.float32)
        _check(lib().p3cs_group_read_palettes(self.handle, first, count, out.ctypes.data))
        return out

    # ---- "next" row f4: animated tight bounds, reduced by the same kernel
    def enable_bounds(self, vertex_column=None, enable: bool = True) -> None:
        """`vertex_column`: a flat.Column (default: the first animated C_point column)."""
        if not enable:
            _check(lib().p3cs_group_enable_bounds(self.handle, None))
            return
        col = vertex_column if vertex_column is not None else self.mesh.flat.skin_columns[0]
        desc = _ColumnDesc(*col.as_tuple())
        _check(lib().p3cs_group_enable_bounds(self.handle, C.byref(desc)))

    def set_bounds_rows(self, referenced=None) -> None:
        """p3cs_group_set_bounds_rows: `referenced` = bool per row (the vertices the Geom's primitives index); None = all."""
        L = lib()
        L.p3cs_group_set_bounds_rows.argtypes = [C.c_void_p, C.c_void_p]
        if referenced is None:
            _check(L.p3cs_group_set_bounds_rows(self.handle, None))
            return
        r = np.asarray(referenced, bool).reshape(-1)
        assert r.size == self.mesh.flat.num_rows
        bits = np.packbits(np.pad(r, (0, (-r.size) % 32)), bitorder="little").view(np.uint32).copy()
        _check(L.p3cs_group_set_bounds_rows(self.handle, bits.ctypes.data))

    def read_bounds(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """[count, 8]: min xyz, max xyz, max |v|^2, found_any -- of the last animate()."""
        count = self.n - first if count is None else count
        out = np.empty((count, 8), np.float32)
        _check(lib().p3cs_group_read_bounds(self.handle, first, count, out.ctypes.data))
        return out

    # ---- the hot path
    def set_output_pointers(self, pointers, o: int = 0) -> None:
        """p3cs_group_set_output_pointers: per-instance destinations (device or page-locked host addresses) of output `o`;
        None restores the group's own buffer."""
        L = lib()
        L.p3cs_group_set_output_pointers.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        if pointers is None:
            _check(L.p3cs_group_set_output_pointers(self.handle, o, None))
            return
        arr = (C.c_void_p * self.n)(*[int(p) for p in pointers])
        _check(L.p3cs_group_set_output_pointers(self.handle, o, arr))

    def animate(self) -> None:
        _check(lib().p3cs_group_animate(self.handle))

    def animate_instances(self, instances) -> None:
        """p3cs_group_animate_instances: recompute only the listed instances."""
        idx = np.ascontiguousarray(instances, np.int32)
        L = lib()
        L.p3cs_group_animate_instances.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _check(L.p3cs_group_animate_instances(self.handle, idx.ctypes.data, int(idx.size)))

    def set_profiling(self, enable: bool) -> None:
        _check(lib().p3cs_group_set_profiling(self.handle, 1 if enable else 0))

    def timings(self) -> Timings:
        t = Timings()
        _check(lib().p3cs_group_last_timings(self.handle, C.byref(t)))
        return t

    # ---- results
    def read_output(self, o: int = 0, first: int = 0, count: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        count = self.n - first if count is None else count
        stride = self.mesh.out_strides[o]
        if out is None:
            out = np.empty((count, self.mesh.num_rows, stride), np.uint8)
        _check(lib().p3cs_group_read_output(self.handle, o, first, count, out.ctypes.data))
        self.mesh.ctx.sync()
        return out

    def read_output_ptr(self, o: int, first: int, count: int, host_ptr: int) -> None:
        _check(lib().p3cs_group_read_output(self.handle, o, first, count, host_ptr))

    def read_selection(self) -> np.ndarray:
        sel = np.empty(self.mesh.num_rows, np.int32)
        _check(lib().p3cs_group_read_selection(self.handle, sel.ctypes.data))
        return sel

    def read_blends(self, instance: int = 0) -> np.ndarray:
        b = np.empty((self.mesh.flat.num_blends, 16), np.float32)
        _check(lib().p3cs_group_read_blends(self.handle, instance, b.ctypes.data))
        return b

    def animate_vertices(self, palette: np.ndarray, sliders: Optional[np.ndarray] = None, instance: int = 0,
                         outs: Optional[List[np.ndarray]] = None) -> List[np.ndarray]:
        """GeomVertexData::animate_vertices(force=True) with host buffers, one instance, synchronous.
        `outs`: destination arrays to reuse (e.g. buffers registered with Context.host_register)."""
        pal = np.ascontiguousarray(palette, np.float32)
        sl = np.ascontiguousarray(sliders, np.float32) if sliders is not None else None
        if outs is None:
            outs = [np.empty((self.mesh.num_rows, s), np.uint8) for s in self.mesh.out_strides]
        ptrs = (C.c_void_p * max(1, len(outs)))(*[o.ctypes.data for o in outs])
        _check(lib().p3cs_animate_vertices(self.handle, instance, pal.ctypes.data,
                                           sl.ctypes.data if sl is not None else None, ptrs))
        return outs

    def animate_host_ptr(self, palettes_ptr: int, sliders_ptr: Optional[int], output_ptrs: List[int], chunk: int = 0) -> None:
        """p3cs_group_animate_host on raw host pointers (page-locked for full overlap); asynchronous."""
        outs = (C.c_void_p * max(1, len(output_ptrs)))(*output_ptrs)
        _check(lib().p3cs_group_animate_host(self.handle, palettes_ptr, sliders_ptr, outs, chunk))

    def animate_host(self, palettes: np.ndarray, sliders: Optional[np.ndarray] = None, chunk: int = 0) -> List[np.ndarray]:
        """The crowd form of animate_vertices: every instance, host buffers in and out, pipelined; synchronous."""
        pal = np.ascontiguousarray(palettes, np.float32)
        sl = np.ascontiguousarray(sliders, np.float32) if sliders is not None else None
        outs = [np.empty((self.n, self.mesh.num_rows, s), np.uint8) for s in self.mesh.out_strides]
        self.animate_host_ptr(pal.ctypes.data, sl.ctypes.data if sl is not None else None, [o.ctypes.data for o in outs], chunk)
        self.mesh.ctx.sync()
        return outs

    def close(self) -> None:
        if self.handle:
            lib().p3cs_group_destroy(self.handle)
            self.handle = C.c_void_p()


class _CrowdDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("num_devices", C.c_int32), ("devices", C.POINTER(C.c_int32)),
                ("num_instances", C.c_int32), ("root_share", C.c_float), ("link_gbs", C.c_float)]


class Crowd:
    """p3cs_crowd: instances of one mesh over several GPUs of THIS process; devices[0] is the rendering GPU, whose
    buffer every device's kernel stores its rows into (sink-aware split, see include/p3cudaskin.h)."""

    def __init__(self, devices: Sequence[int], mesh: FlatMesh, num_instances: int, root_share: float = 0.0, link_gbs: float = 0.0):
        self.flat, self.n = mesh, num_instances
        self.handle = C.c_void_p()
        keep: list = []
        desc = mesh_desc(mesh, keep)
        devs = (C.c_int32 * len(devices))(*devices)
        cd = _CrowdDesc(C.sizeof(_CrowdDesc), len(devices), devs, num_instances, root_share, link_gbs)
        L = lib()
        L.p3cs_crowd_output_pitch.restype = C.c_size_t
        L.p3cs_crowd_output_device.restype = C.c_void_p
        _check(L.p3cs_crowd_create(C.byref(cd), C.byref(desc), C.byref(self.handle)))
        self.num_devices = int(L.p3cs_crowd_num_devices(self.handle))

    def ranges(self) -> List[Tuple[int, int]]:
        out = []
        for d in range(self.num_devices):
            first, count = C.c_int32(), C.c_int32()
            _check(lib().p3cs_crowd_range(self.handle, d, C.byref(first), C.byref(count)))
            out.append((first.value, count.value))
        return out

    def set_palettes(self, matrices: np.ndarray, first: int = 0) -> None:
        m = np.ascontiguousarray(matrices, np.float32)
        count = m.size // (self.flat.num_transforms * 16)
        _check(lib().p3cs_crowd_set_palettes(self.handle, first, count, m.ctypes.data_as(C.c_void_p)))

    def set_sliders(self, values: np.ndarray, first: int = 0) -> None:
        v = np.ascontiguousarray(values, np.float32)
        count = v.size // max(1, self.flat.num_sliders)
        _check(lib().p3cs_crowd_set_sliders(self.handle, first, count, v.ctypes.data_as(C.c_void_p)))

    def animate(self) -> None:
        _check(lib().p3cs_crowd_animate(self.handle))

    def sync(self) -> None:
        _check(lib().p3cs_crowd_sync(self.handle))

    def tune(self, steps: int = 3) -> None:
        _check(lib().p3cs_crowd_tune(self.handle, steps))

    def last_ms(self) -> List[float]:
        ms = (C.c_float * self.num_devices)()
        _check(lib().p3cs_crowd_last_ms(self.handle, ms))
        return list(ms)

    def read_output(self, o: int = 0, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n - first if count is None else count
        stride = self.flat.strides[self.flat.output_arrays[o]]
        out = np.empty((count, self.flat.num_rows, stride), np.uint8)
        _check(lib().p3cs_crowd_read_output(self.handle, o, first, count, out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self) -> None:
        if self.handle:
            lib().p3cs_crowd_destroy(self.handle)
            self.handle = C.c_void_p()


def animate_groups(groups) -> None:
    """p3cs_animate_groups: one frame of a mixed crowd; independent groups overlap on side streams."""
    arr = (C.c_void_p * len(groups))(*[g.handle.value for g in groups])
    _check(lib().p3cs_animate_groups(arr, len(groups)))
