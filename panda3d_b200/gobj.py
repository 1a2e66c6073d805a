"""Host-side mirror of the reference's interface for the vertex-animation path.

Same class names, method names, argument meaning and error behaviour as Panda3D's published
API (panda/src/gobj, panda/src/putil, panda/src/linmath), restricted to what
GeomVertexData::animate_vertices touches, so parity tests read like tests written against
`panda3d.core`.  Storage is numpy; the computation is NOT here: animate_vertices flattens the
objects (the job of the Panda-side shim described in INTEGRATION.md) and calls the C-ABI of
libp3cudaskin.so.  With `vertex-animation-backend cpu` (the reference's own CPU loops) this
package raises: it has no CPU fallback.

Citations are file:line in the Panda3D 1.11.0 tree.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _capi, flat as F
from .flat import Column, FlatMesh, Morph

# ----------------------------------------------------------------------------- PRC config
# dtool/src/prc ConfigVariable*: name -> value, set through load_prc_file_data like the reference
# (panda/src/putil/load_prc_file.h:38-50).
_prc: Dict[str, str] = {
    "hardware-animated-vertices": "#f",       # gobj/config_gobj.cxx:178-188 (default false)
    "vertex-animation-backend": "cuda",       # NEW switch: cpu | cuda
    "cuda-skinning-library": "p3cudaskin",    # NEW: libp3cudaskin.so, loaded like load-display plugins
    "cuda-skinning-device": "0",              # NEW: CUDA device ordinal of the rendering GPU
    "vertex-animation-align-16": "#f",        # gobj/config_gobj.cxx:286-301 (false without Eigen)
    "vertex-column-alignment": "4",           # gobj/config_gobj.cxx:273
    "coordinate-system": "zup-right",         # linmath/coordinateSystem.cxx:30
}


def load_prc_file_data(name: str, data: str) -> None:
    for line in data.splitlines():
        line = line.split("#!", 1)[0].strip()
        if not line:
            continue
        key, _, value = line.partition(" ")
        _prc[key.strip()] = value.strip()


def get_config(name: str) -> str:
    return _prc[name]


def _config_bool(name: str) -> bool:
    return _prc[name].lower() in ("#t", "1", "true", "t")


_CS = {"zup-right": F.CS_zup_right, "yup-right": F.CS_yup_right, "zup-left": F.CS_zup_left, "yup-left": F.CS_yup_left}


def get_default_coordinate_system() -> int:
    return _CS[_prc["coordinate-system"]]


# ----------------------------------------------------------------------------- UpdateSeq
class UpdateSeq:
    """panda/src/putil/updateSeq.h: sequence number with three special values."""
    SC_initial, SC_old, SC_fresh = 0, 1, 0xFFFFFFFF

    def __init__(self, other: Optional["UpdateSeq"] = None):
        self._seq = other._seq if other is not None else self.SC_initial

    @classmethod
    def _make(cls, v: int) -> "UpdateSeq":
        s = cls()
        s._seq = v & 0xFFFFFFFF
        return s

    @classmethod
    def initial(cls): return cls._make(cls.SC_initial)
    @classmethod
    def old(cls): return cls._make(cls.SC_old)
    @classmethod
    def fresh(cls): return cls._make(cls.SC_fresh)

    def clear(self): self._seq = self.SC_initial
    def is_initial(self): return self._seq == self.SC_initial
    def is_old(self): return self._seq == self.SC_old
    def is_fresh(self): return self._seq == self.SC_fresh
    def is_special(self): return self._special(self._seq)

    @staticmethod
    def _special(v: int) -> bool:          # updateSeq.I:224-227
        return ((v + 1) & 0xFFFFFFFF) <= 2

    @classmethod
    def _lt(cls, a: int, b: int) -> bool:  # updateSeq.I:233-242: circular compare unless special
        if cls._special(a) or cls._special(b):
            return a < b
        d = (a - b) & 0xFFFFFFFF
        return d >= 0x80000000

    def __eq__(self, o): return self._seq == o._seq
    def __ne__(self, o): return self._seq != o._seq
    def __lt__(self, o): return self._lt(self._seq, o._seq)
    def __le__(self, o): return self._seq == o._seq or self._lt(self._seq, o._seq)
    def __gt__(self, o): return o < self
    def __ge__(self, o): return o <= self
    def __hash__(self): return hash(self._seq)

    def increment(self) -> "UpdateSeq":    # operator ++ (), updateSeq.I:145-163
        n = (self._seq + 1) & 0xFFFFFFFF
        if self._special(n):
            n = self.SC_old + 1
        self._seq = n
        return UpdateSeq(self)

    @property
    def seq(self) -> int: return self._seq
    def get_seq(self) -> int: return self._seq
    def __repr__(self): return f"UpdateSeq({self._seq})"


# ----------------------------------------------------------------------------- SparseArray
class SparseArray:
    """panda/src/putil/sparseArray.h: sorted disjoint [begin,end) subranges + inverse flag."""

    def __init__(self):
        self._ranges: List[Tuple[int, int]] = []
        self._inverse = False

    @classmethod
    def all_on(cls):
        s = cls(); s._inverse = True; return s
    @classmethod
    def all_off(cls): return cls()
    @classmethod
    def lower_on(cls, on_bits: int):
        s = cls(); s.set_range(0, on_bits); return s
    @classmethod
    def bit(cls, index: int):
        s = cls(); s.set_range(index, 1); return s
    @classmethod
    def range(cls, low_bit: int, size: int):
        s = cls(); s.set_range(low_bit, size); return s

    def _set(self, begin: int, end: int, value: bool) -> None:
        if end <= begin:
            return
        out: List[Tuple[int, int]] = []
        for b, e in self._ranges:
            if e < begin or b > end:       # disjoint and not adjacent
                out.append((b, e))
            elif value:                    # merge
                begin, end = min(begin, b), max(end, e)
            else:                          # cut
                if b < begin: out.append((b, begin))
                if e > end: out.append((end, e))
        if value:
            out.append((begin, end))
        self._ranges = sorted(r for r in out if r[1] > r[0])

    def set_range(self, low_bit: int, size: int) -> None: self._set(low_bit, low_bit + size, not self._inverse)
    def clear_range(self, low_bit: int, size: int) -> None: self._set(low_bit, low_bit + size, self._inverse)
    def set_range_to(self, value: bool, low_bit: int, size: int) -> None:
        (self.set_range if value else self.clear_range)(low_bit, size)
    def set_bit(self, index: int) -> None: self.set_range(index, 1)
    def clear_bit(self, index: int) -> None: self.clear_range(index, 1)
    def set_bit_to(self, index: int, value: bool) -> None: self.set_range_to(value, index, 1)

    def get_bit(self, index: int) -> bool:
        inside = any(b <= index < e for b, e in self._ranges)
        return inside != self._inverse

    def has_any_of(self, low_bit: int, size: int) -> bool:
        return any(self.get_bit(i) for i in range(low_bit, low_bit + size)) if size < 4096 else \
            (self & SparseArray.range(low_bit, size)).is_zero() is False
    def has_all_of(self, low_bit: int, size: int) -> bool:
        probe = SparseArray.range(low_bit, size)
        return (self & probe) == probe

    def get_highest_bits(self) -> bool: return self._inverse
    def is_zero(self) -> bool: return not self._inverse and not self._ranges
    def is_all_on(self) -> bool: return self._inverse and not self._ranges
    def is_inverse(self) -> bool: return self._inverse
    def get_num_subranges(self) -> int: return len(self._ranges)
    def get_subrange_begin(self, n: int) -> int: return self._ranges[n][0]
    def get_subrange_end(self, n: int) -> int: return self._ranges[n][1]

    def get_num_on_bits(self) -> int:
        return -1 if self._inverse else sum(e - b for b, e in self._ranges)
    def get_num_off_bits(self) -> int:
        return sum(e - b for b, e in self._ranges) if self._inverse else -1
    def get_lowest_on_bit(self) -> int:
        if self._inverse:
            return -1
        return self._ranges[0][0] if self._ranges else -1
    def get_highest_on_bit(self) -> int:
        if self._inverse:
            return -1
        return self._ranges[-1][1] - 1 if self._ranges else -1

    def invert_in_place(self) -> None: self._inverse = not self._inverse
    def clear(self) -> None:
        self._ranges, self._inverse = [], False

    def _bits_fn(self):
        return lambda i: self.get_bit(i)

    def _combine(self, other: "SparseArray", op) -> "SparseArray":
        pts = sorted({p for b, e in self._ranges + other._ranges for p in (b, e)})
        res = SparseArray()
        res._inverse = op(self._inverse, other._inverse)
        # evaluate on each elementary interval; store the ranges whose value differs from the tail
        edges = [-(1 << 62)] + pts + [(1 << 62)]
        for lo, hi in zip(edges[:-1], edges[1:]):
            probe = lo if lo > -(1 << 62) else (pts[0] - 1 if pts else 0)
            v = op(self.get_bit(probe), other.get_bit(probe))
            if v != res._inverse and lo > -(1 << 62) and hi < (1 << 62):
                res._ranges.append((lo, hi))
            elif v != res._inverse:
                raise ValueError("unbounded range differs from the inverse flag")
        merged: List[Tuple[int, int]] = []
        for b, e in res._ranges:
            if merged and merged[-1][1] == b:
                merged[-1] = (merged[-1][0], e)
            else:
                merged.append((b, e))
        res._ranges = merged
        return res

    def __and__(self, o): return self._combine(o, lambda a, b: a and b)
    def __or__(self, o): return self._combine(o, lambda a, b: a or b)
    def __xor__(self, o): return self._combine(o, lambda a, b: a != b)
    def __invert__(self):
        s = SparseArray(); s._ranges = list(self._ranges); s._inverse = not self._inverse; return s
    def __eq__(self, o): return self._inverse == o._inverse and self._ranges == o._ranges
    def __ne__(self, o): return not self == o
    def __hash__(self): return hash((self._inverse, tuple(self._ranges)))
    def __repr__(self): return f"SparseArray({'~' if self._inverse else ''}{self._ranges})"

    def ranges(self) -> np.ndarray:
        return np.asarray(self._ranges, np.int32).reshape(-1, 2)


# ----------------------------------------------------------------------------- names / formats
class InternalName:
    """panda/src/gobj/internalName.h, as plain dotted strings."""
    @staticmethod
    def make(name: str) -> str: return name
    @staticmethod
    def get_vertex(): return "vertex"
    @staticmethod
    def get_normal(): return "normal"
    @staticmethod
    def get_tangent(): return "tangent"
    @staticmethod
    def get_binormal(): return "binormal"
    @staticmethod
    def get_texcoord(): return "texcoord"
    @staticmethod
    def get_color(): return "color"
    @staticmethod
    def get_transform_blend(): return "transform_blend"
    @staticmethod
    def get_morph(column: str, slider: str) -> str:   # internalName.I:300-306
        return f"{column}.morph.{slider}"


class GeomEnums:
    NT_uint8, NT_uint16, NT_uint32, NT_packed_dcba, NT_packed_dabc, NT_float32, NT_float64, NT_stdfloat, \
        NT_int8, NT_int16, NT_int32, NT_packed_ufloat = range(12)
    C_other, C_point, C_clip_point, C_vector, C_texcoord, C_color, C_index, C_morph_delta, C_matrix, C_normal = range(10)
    AT_none, AT_panda, AT_hardware = range(3)
    UH_client, UH_stream, UH_dynamic, UH_static, UH_unspecified = range(5)


class GeomVertexColumn:
    """panda/src/gobj/geomVertexColumn.cxx:154-234 (setup): alignment and byte size rules."""

    def __init__(self, name: str, num_components: int, numeric_type: int, contents: int, start: int,
                 column_alignment: int = 0):
        if numeric_type == GeomEnums.NT_stdfloat:
            numeric_type = GeomEnums.NT_float32
        self._name, self._num_components = name, num_components
        self._numeric_type, self._contents = numeric_type, contents
        self._component_bytes = F.NUMERIC_BYTES[numeric_type]
        self._num_values = num_components * (4 if numeric_type in (3, 4) else 3 if numeric_type == 11 else 1)
        if column_alignment < 1:
            column_alignment = max(self._component_bytes, int(_prc["vertex-column-alignment"]))
        self._column_alignment = column_alignment
        self._start = (start + column_alignment - 1) // column_alignment * column_alignment
        self._total_bytes = self._component_bytes * num_components

    def get_name(self): return self._name
    def get_num_components(self): return self._num_components
    def get_num_values(self): return self._num_values
    def get_numeric_type(self): return self._numeric_type
    def get_contents(self): return self._contents
    def get_start(self): return self._start
    def get_column_alignment(self): return self._column_alignment
    def get_component_bytes(self): return self._component_bytes
    def get_total_bytes(self): return self._total_bytes
    def has_homogeneous_coord(self):          # geomVertexColumn.I:183-192
        return self._contents in (GeomEnums.C_point, GeomEnums.C_texcoord)
    def overlaps_with(self, start: int, nbytes: int) -> bool:
        return start < self._start + self._total_bytes and self._start < start + nbytes

    def flat(self, array: int) -> Column:
        return Column(array, self._start, self._num_values, self._numeric_type, self._contents)


class GeomVertexArrayFormat:
    """panda/src/gobj/geomVertexArrayFormat.cxx: add_column :216-272, remove_column :278-304,
    align_columns_for_animation :349-373."""

    def __init__(self, *args):
        self._columns: List[GeomVertexColumn] = []
        self._stride = self._total_bytes = 0
        self._pad_to = 1
        for i in range(0, len(args), 4):
            self.add_column(*args[i:i + 4])

    def add_column(self, name: str, num_components: int, numeric_type: int, contents: int, start: int = -1,
                   column_alignment: int = 0) -> int:
        if start < 0:
            start = self._total_bytes
        col = GeomVertexColumn(name, num_components, numeric_type, contents, start, column_alignment)
        self.remove_column(name)
        for other in [c for c in self._columns if c.overlaps_with(col.get_start(), col.get_total_bytes())]:
            self.remove_column(other.get_name())
        self._total_bytes = max(self._total_bytes, col.get_start() + col.get_total_bytes())
        self._pad_to = max(self._pad_to, col.get_column_alignment())
        self._stride = max(self._stride, self._total_bytes)
        if self._pad_to > 1:
            self._stride = (self._stride + self._pad_to - 1) // self._pad_to * self._pad_to
        self._columns.append(col)
        return len(self._columns) - 1

    def remove_column(self, name: str) -> None:
        self._columns = [c for c in self._columns if c.get_name() != name]
        self._total_bytes = max((c.get_start() + c.get_total_bytes() for c in self._columns), default=0)

    def copy(self) -> "GeomVertexArrayFormat":
        f = GeomVertexArrayFormat()
        f._columns = list(self._columns)
        f._stride, f._total_bytes, f._pad_to = self._stride, self._total_bytes, self._pad_to
        return f

    def align_columns_for_animation(self) -> None:
        orig = sorted(self._columns, key=lambda c: c.get_start())
        self._columns, self._stride, self._total_bytes, self._pad_to = [], 0, 0, 1
        for c in orig:
            if c.get_contents() in (GeomEnums.C_point, GeomEnums.C_vector, GeomEnums.C_normal) and \
                    c.get_numeric_type() in (GeomEnums.NT_float32, GeomEnums.NT_float64) and c.get_num_components() >= 3:
                self.add_column(c.get_name(), 4, c.get_numeric_type(), c.get_contents(), -1, 16)
            else:
                self.add_column(c.get_name(), c.get_num_components(), c.get_numeric_type(), c.get_contents(), -1,
                                c.get_column_alignment())

    def get_stride(self): return self._stride
    def set_stride(self, stride: int): self._stride = stride
    def get_total_bytes(self): return self._total_bytes
    def get_pad_to(self): return self._pad_to
    def get_num_columns(self): return len(self._columns)
    def get_column(self, key):
        if isinstance(key, int):
            return sorted(self._columns, key=lambda c: c.get_start())[key]
        return next((c for c in self._columns if c.get_name() == key), None)
    def has_column(self, name: str): return self.get_column(name) is not None


class GeomVertexAnimationSpec:
    def __init__(self): self._type = GeomEnums.AT_none
    def set_none(self): self._type = GeomEnums.AT_none
    def set_panda(self): self._type = GeomEnums.AT_panda
    def set_hardware(self, num_transforms: int = 4, indexed: bool = True): self._type = GeomEnums.AT_hardware
    def get_animation_type(self): return self._type


class GeomVertexFormat:
    """panda/src/gobj/geomVertexFormat.cxx: do_register :753-811 (points / vectors / morphs),
    get_post_animated_format :106-135."""

    def __init__(self, array_format: Optional[GeomVertexArrayFormat] = None):
        self._arrays: List[GeomVertexArrayFormat] = []
        self._animation = GeomVertexAnimationSpec()
        self._registered = False
        self._post: Optional["GeomVertexFormat"] = None
        if array_format is not None:
            self.add_array(array_format)

    def add_array(self, array_format: GeomVertexArrayFormat) -> int:
        assert not self._registered
        self._arrays.append(array_format)
        return len(self._arrays) - 1

    def set_animation(self, spec: GeomVertexAnimationSpec) -> None: self._animation = spec
    def get_animation(self) -> GeomVertexAnimationSpec: return self._animation
    def get_num_arrays(self): return len(self._arrays)
    def get_array(self, i: int): return self._arrays[i]
    def is_registered(self): return self._registered

    def align_columns_for_animation(self) -> None:
        assert not self._registered
        self._arrays = [a.copy() for a in self._arrays]
        for a in self._arrays:
            a.align_columns_for_animation()

    @staticmethod
    def register_format(fmt: "GeomVertexFormat") -> "GeomVertexFormat":
        if fmt._registered:
            return fmt
        if fmt._animation.get_animation_type() == GeomEnums.AT_panda and _config_bool("vertex-animation-align-16"):
            fmt.align_columns_for_animation()  # maybe_align_columns_for_animation, geomVertexFormat.cxx:578-585
        fmt._points, fmt._vectors, fmt._morphs = [], [], []
        names = sorted((c.get_name(), a) for a, af in enumerate(fmt._arrays) for c in af._columns)
        for name, a in names:
            col = fmt._arrays[a].get_column(name)
            if col.get_contents() == GeomEnums.C_point:
                fmt._points.append(name)
            elif col.get_contents() in (GeomEnums.C_vector, GeomEnums.C_normal):
                fmt._vectors.append(name)
            elif col.get_contents() == GeomEnums.C_morph_delta:
                parts = name.split(".")
                if "morph" in parts:
                    n = parts.index("morph")
                    base, slider = ".".join(parts[:n]), ".".join(parts[n + 1:])
                    if fmt.get_array_with(base) >= 0:
                        fmt._morphs.append((slider, base, name))
        fmt._registered = True
        return fmt

    def get_array_with(self, name: str) -> int:
        return next((a for a, af in enumerate(self._arrays) if af.has_column(name)), -1)
    def get_column(self, name: str) -> Optional[GeomVertexColumn]:
        a = self.get_array_with(name)
        return self._arrays[a].get_column(name) if a >= 0 else None
    def has_column(self, name: str) -> bool: return self.get_array_with(name) >= 0

    def get_num_points(self): return len(self._points)
    def get_point(self, i): return self._points[i]
    def get_num_vectors(self): return len(self._vectors)
    def get_vector(self, i): return self._vectors[i]
    def get_num_morphs(self): return len(self._morphs)
    def get_morph_slider(self, i): return self._morphs[i][0]
    def get_morph_base(self, i): return self._morphs[i][1]
    def get_morph_delta(self, i): return self._morphs[i][2]

    def get_post_animated_format(self) -> "GeomVertexFormat":
        assert self._registered
        if self._post is None:
            new = GeomVertexFormat()
            for af in self._arrays:
                c = af.copy()
                c.remove_column(InternalName.get_transform_blend())
                for _, _, delta in self._morphs:
                    c.remove_column(delta)
                if c.get_num_columns() > 0:      # emptied arrays are dropped (geomVertexFormat.cxx:527-529)
                    new._arrays.append(c)
            new._animation.set_none()
            self._post = GeomVertexFormat.register_format(new)
        return self._post


# ----------------------------------------------------------------------------- transforms / sliders
_global_modified = UpdateSeq()           # VertexTransform::_next_modified, vertexTransform.cxx:96-103
_addresses = itertools.count(1)          # stand-in for object addresses (ordering of TransformBlend entries)


class VertexTransform:
    """panda/src/gobj/vertexTransform.h:36-109."""

    def __init__(self):
        self._modified = UpdateSeq()
        self._address = next(_addresses)

    def get_matrix(self) -> np.ndarray:
        raise NotImplementedError

    def get_modified(self, current_thread=None) -> UpdateSeq: return self._modified

    def mark_modified(self, current_thread=None) -> None:      # vertexTransform.cxx:110-119
        self._modified = VertexTransform.get_next_modified()

    @staticmethod
    def get_next_modified(current_thread=None) -> UpdateSeq: return _global_modified.increment()
    @staticmethod
    def get_global_modified(current_thread=None) -> UpdateSeq: return UpdateSeq(_global_modified)


class UserVertexTransform(VertexTransform):
    """panda/src/gobj/userVertexTransform.I:17-31."""

    def __init__(self, name: str = ""):
        super().__init__()
        self._name = name
        self._matrix = np.eye(4, dtype=np.float32)

    def get_name(self): return self._name
    def set_matrix(self, matrix) -> None:
        self._matrix = np.asarray(matrix, np.float32).reshape(4, 4).copy()
        self.mark_modified()
    def get_matrix(self) -> np.ndarray: return self._matrix


class TransformBlend:
    """panda/src/gobj/transformBlend.h:32-146; entries sorted by transform address (.I:354-357)."""

    def __init__(self, *args):
        self._entries: List[List] = []          # [transform, weight], ascending transform address
        if len(args) == 2:                      # (transform0, weight0): the weight is ignored, 1.0 is used (.I:24-27)
            self.add_transform(args[0], 1.0)
        else:
            for i in range(0, len(args), 2):
                self.add_transform(args[i], args[i + 1])

    def add_transform(self, transform: VertexTransform, weight: float) -> None:   # transformBlend.cxx:51-72
        weight = float(np.float32(weight))
        if abs(weight) < 1e-6:
            return
        for e in self._entries:
            if e[0] is transform:
                e[1] = float(np.float32(e[1] + weight))
                if abs(e[1]) < 1e-6:
                    self._entries.remove(e)
                return
        self._entries.append([transform, weight])
        self._entries.sort(key=lambda e: e[0]._address)

    def remove_transform(self, transform: VertexTransform) -> None:
        self._entries = [e for e in self._entries if e[0] is not transform]

    def limit_transforms(self, max_transforms: int) -> None:                     # transformBlend.cxx:95-115
        if max_transforms <= 0:
            self._entries = []
            return
        while len(self._entries) > max_transforms:
            least = min(range(len(self._entries)), key=lambda i: (self._entries[i][1], i))
            del self._entries[least]

    def normalize_weights(self) -> None:                                          # transformBlend.cxx:122-137
        net = np.float32(0)
        for _, w in self._entries:
            net = np.float32(net + np.float32(w))
        if net != 0:
            for e in self._entries:
                e[1] = float(np.float32(np.float32(e[1]) / net))

    def has_transform(self, transform) -> bool: return any(e[0] is transform for e in self._entries)
    def get_weight(self, key) -> float:
        if isinstance(key, int):
            return self._entries[key][1]
        return next((e[1] for e in self._entries if e[0] is key), 0.0)
    def get_num_transforms(self) -> int: return len(self._entries)
    def get_transform(self, n: int) -> VertexTransform: return self._entries[n][0]
    def set_transform(self, n: int, transform) -> None:
        self._entries[n][0] = transform
        self._entries.sort(key=lambda e: e[0]._address)
    def set_weight(self, n: int, weight: float) -> None: self._entries[n][1] = float(np.float32(weight))

    def get_modified(self, current_thread=None) -> UpdateSeq:                     # transformBlend.I:337-347
        seq = UpdateSeq()
        for t, _ in self._entries:
            if t.get_modified() > seq:
                seq = t.get_modified()
        return seq

    def _key(self):
        return tuple((e[0]._address, round(e[1] / 1e-6)) for e in self._entries)   # compare_to, transformBlend.cxx:24-45


class TransformBlendTable:
    """panda/src/gobj/transformBlendTable.h:45-157."""

    def __init__(self):
        self._blends: List[TransformBlend] = []
        self._index: Dict[tuple, int] = {}
        self._rows = SparseArray()
        self._modified = UpdateSeq()

    def get_num_blends(self): return len(self._blends)
    def get_blend(self, n: int) -> TransformBlend: return self._blends[n]
    def set_blend(self, n: int, blend: TransformBlend) -> None:
        self._blends[n] = blend
        self._index = {b._key(): i for i, b in enumerate(self._blends)}
        self._touch()
    def remove_blend(self, n: int) -> None:
        del self._blends[n]
        self._index = {b._key(): i for i, b in enumerate(self._blends)}
        self._touch()

    def add_blend(self, blend: TransformBlend) -> int:            # transformBlendTable.cxx:86-119 (dedup)
        k = blend._key()
        if k in self._index:
            return self._index[k]
        self._blends.append(blend)
        self._index[k] = len(self._blends) - 1
        self._touch()
        return len(self._blends) - 1

    def _touch(self): self._static_modified = VertexTransform.get_next_modified()

    def get_num_transforms(self) -> int: return len(self._palette())
    def get_max_simultaneous_transforms(self) -> int:
        return max((b.get_num_transforms() for b in self._blends), default=0)
    def set_rows(self, rows: SparseArray) -> None:
        self._rows = rows
        self._touch()
    def get_rows(self) -> SparseArray: return self._rows

    def get_modified(self, current_thread=None) -> UpdateSeq:     # transformBlendTable.cxx:172-186
        seq = UpdateSeq()
        for b in self._blends:
            m = b.get_modified()
            if m > seq:
                seq = m
        return seq

    def _palette(self) -> List[VertexTransform]:
        seen = {}
        for b in self._blends:
            for t, _ in b._entries:
                seen[t._address] = t
        return [seen[a] for a in sorted(seen)]


class VertexSlider:
    """panda/src/gobj/vertexSlider.h."""

    def __init__(self, name: str):
        self._name = name
        self._modified = UpdateSeq()

    def get_name(self): return self._name
    def get_slider(self) -> float: raise NotImplementedError
    def get_modified(self, current_thread=None): return self._modified
    def mark_modified(self, current_thread=None): self._modified = VertexTransform.get_next_modified()


class UserVertexSlider(VertexSlider):
    """panda/src/gobj/userVertexSlider.I:17-23."""

    def __init__(self, name: str):
        super().__init__(name)
        self._slider = 0.0

    def set_slider(self, value: float) -> None:
        self._slider = float(np.float32(value))
        self.mark_modified()
    def get_slider(self) -> float: return self._slider


class SliderTable:
    """panda/src/gobj/sliderTable.h, .I:52-117."""

    def __init__(self):
        self._sliders: List[Tuple[VertexSlider, SparseArray]] = []
        self._registered = False

    @staticmethod
    def register_table(table: "SliderTable") -> "SliderTable":
        table._registered = True
        return table

    def is_registered(self): return self._registered
    def get_num_sliders(self): return len(self._sliders)
    def get_slider(self, n: int) -> VertexSlider: return self._sliders[n][0]
    def get_slider_rows(self, n: int) -> SparseArray: return self._sliders[n][1]
    def add_slider(self, slider: VertexSlider, rows: SparseArray) -> int:
        assert not self._registered
        self._sliders.append((slider, rows))
        return len(self._sliders) - 1
    def find_sliders(self, name: str) -> SparseArray:
        s = SparseArray()
        for i, (sl, _) in enumerate(self._sliders):
            if sl.get_name() == name:
                s.set_bit(i)
        return s
    def has_slider(self, name: str) -> bool: return not self.find_sliders(name).is_zero()
    def get_modified(self, current_thread=None) -> UpdateSeq:     # sliderTable.I:123-127
        seq = UpdateSeq()
        for sl, _ in self._sliders:
            if sl.get_modified() > seq:
                seq = sl.get_modified()
        return seq


# ----------------------------------------------------------------------------- GeomVertexData
class _Backend:
    """Process-wide CUDA backend state (what the Panda-side shim keeps per GSG / cull thread)."""
    ctx: Optional[_capi.Context] = None

    @classmethod
    def context(cls) -> _capi.Context:
        if cls.ctx is None:
            cls.ctx = _capi.Context(int(_prc["cuda-skinning-device"]))
        return cls.ctx


class GeomVertexData:
    """panda/src/gobj/geomVertexData.h: the subset animate_vertices needs."""

    def __init__(self, name: str, format: GeomVertexFormat, usage_hint: int = GeomEnums.UH_static):
        assert format.is_registered(), "format must be registered (GeomVertexFormat.register_format)"
        self._name, self._format, self._usage_hint = name, format, usage_hint
        self._arrays = [np.zeros((0, af.get_stride()), np.uint8) for af in format._arrays]
        self._tbt: Optional[TransformBlendTable] = None
        self._sliders: Optional[SliderTable] = None
        self._animated: Optional["GeomVertexData"] = None
        self._animated_modified = UpdateSeq()
        self._modified = UpdateSeq()
        self._device = None          # (static key, _capi.Mesh, _capi.Group)

    def get_name(self): return self._name
    def get_format(self): return self._format
    def get_usage_hint(self): return self._usage_hint
    def get_num_rows(self): return int(self._arrays[0].shape[0]) if self._arrays else 0
    def get_num_arrays(self): return len(self._arrays)

    def set_num_rows(self, n: int) -> bool:
        for i, af in enumerate(self._format._arrays):
            old = self._arrays[i]
            new = np.zeros((n, af.get_stride()), np.uint8)
            new[:min(n, old.shape[0])] = old[:n]
            self._arrays[i] = new
        self._touch()
        return True
    unclean_set_num_rows = set_num_rows

    def get_array(self, i: int) -> np.ndarray: return self._arrays[i]
    def modify_array(self, i: int) -> np.ndarray:
        self._touch()
        return self._arrays[i]
    def set_array(self, i: int, data: np.ndarray) -> None:
        self._arrays[i] = np.ascontiguousarray(data, np.uint8).reshape(-1, self._format.get_array(i).get_stride())
        self._touch()

    def _touch(self): self._modified = VertexTransform.get_next_modified()
    def get_modified(self): return self._modified

    def set_column(self, name: str, values: np.ndarray) -> None:
        """Convenience writer (GeomVertexWriter over a whole column)."""
        a = self._format.get_array_with(name)
        col = self._format.get_column(name)
        dt = {GeomEnums.NT_float32: np.float32, GeomEnums.NT_uint8: np.uint8, GeomEnums.NT_uint16: np.uint16,
              GeomEnums.NT_uint32: np.uint32}[col.get_numeric_type()]
        v = np.ascontiguousarray(values, dt).reshape(self.get_num_rows(), col.get_num_values())
        self._arrays[a][:, col.get_start():col.get_start() + col.get_total_bytes()] = v.view(np.uint8)
        self._touch()

    def get_column(self, name: str) -> np.ndarray:
        a = self._format.get_array_with(name)
        col = self._format.get_column(name)
        dt = {GeomEnums.NT_float32: np.float32, GeomEnums.NT_uint8: np.uint8, GeomEnums.NT_uint16: np.uint16,
              GeomEnums.NT_uint32: np.uint32}[col.get_numeric_type()]
        raw = np.ascontiguousarray(self._arrays[a][:, col.get_start():col.get_start() + col.get_total_bytes()])
        return raw.view(dt).reshape(self.get_num_rows(), col.get_num_values())

    def set_transform_blend_table(self, table: Optional[TransformBlendTable]) -> None:
        self._tbt = table
        self.clear_animated_vertices()
    def get_transform_blend_table(self): return self._tbt
    def set_slider_table(self, table: Optional[SliderTable]) -> None:
        assert table is None or table.is_registered(), "slider table must be registered (geomVertexData.cxx:403)"
        self._sliders = table
        self.clear_animated_vertices()
    def get_slider_table(self): return self._sliders

    def clear_animated_vertices(self) -> None:            # geomVertexData.cxx:982-989
        self._animated_modified = UpdateSeq()
        self._animated = None

    # ---- the flattening the Panda-side shim performs -------------------------
    def flatten(self) -> Tuple[FlatMesh, List[VertexTransform]]:
        fmt, post = self._format, self._format.get_post_animated_format()
        kept = [af.copy() for af in fmt._arrays]
        is_output = []
        for af in kept:
            af.remove_column(InternalName.get_transform_blend())
            for _, _, delta in fmt._morphs:
                af.remove_column(delta)
            is_output.append(af.get_num_columns() > 0)
        skin = [fmt.get_column(n).flat(fmt.get_array_with(n)) for n in post._points + post._vectors]
        tbt = self._tbt
        blend_column, palette = None, []
        offsets, xf, w = [0], [], []
        rows = np.zeros((0, 2), np.int32)
        if tbt is not None:
            a = fmt.get_array_with(InternalName.get_transform_blend())
            if a < 0:
                raise ValueError(f"Vertex data {self._name} has a transform_blend_table, but no transform_blend data.")
            blend_column = fmt.get_column(InternalName.get_transform_blend()).flat(a)
            palette = tbt._palette()
            slot = {t._address: i for i, t in enumerate(palette)}
            for b in tbt._blends:
                for t, wt in b._entries:          # stored order = accumulation order
                    xf.append(slot[t._address])
                    w.append(wt)
                offsets.append(len(xf))
            assert not tbt.get_rows().is_inverse()
            rows = tbt.get_rows().ranges()
        morphs: List[Morph] = []
        st = self._sliders
        if st is not None:
            for slider_name, base, delta in fmt._morphs:       # geomVertexData.cxx:1467-1480
                found = st.find_sliders(slider_name)
                for sni in range(found.get_num_subranges()):
                    for sn in range(found.get_subrange_begin(sni), found.get_subrange_end(sni)):
                        r = st.get_slider_rows(sn)
                        assert not r.is_inverse()
                        morphs.append(Morph(sn, fmt.get_column(base).flat(fmt.get_array_with(base)),
                                            fmt.get_column(delta).flat(fmt.get_array_with(delta)), r.ranges()))
        mesh = FlatMesh(
            num_rows=self.get_num_rows(), arrays=[np.ascontiguousarray(a) for a in self._arrays], is_output=is_output,
            skin_columns=skin if tbt is not None else [], blend_column=blend_column, num_transforms=len(palette),
            blend_offsets=np.asarray(offsets, np.int32), blend_transform=np.asarray(xf, np.int32),
            blend_weight=np.asarray(w, np.float32), row_ranges=rows,
            num_sliders=st.get_num_sliders() if st is not None else 0, morphs=morphs,
            coordinate_system=get_default_coordinate_system())
        return mesh, palette

    # ---- the hot path ---------------------------------------------------------------------------
    def animate_vertices(self, force: bool = True, current_thread=None) -> "GeomVertexData":
        """GeomVertexData::animate_vertices (geomVertexData.cxx:912-975): same gating, same
        UpdateSeq caching, same returned object reused and mutated in place."""
        if self._format.get_animation().get_animation_type() != GeomEnums.AT_panda:
            return self
        if self._tbt is not None:
            modified = self._tbt.get_modified()
            if self._sliders is not None and self._sliders.get_modified() > modified:
                modified = self._sliders.get_modified()
        elif self._sliders is not None:
            modified = self._sliders.get_modified()
        else:
            return self
        if self._animated is not None and self._animated_modified == modified:
            return self._animated
        self._animated_modified = UpdateSeq(modified)
        self._update_animated_vertices()
        return self._animated

    def _update_animated_vertices(self) -> None:
        backend = _prc["vertex-animation-backend"]
        if backend != "cuda":
            raise RuntimeError("vertex-animation-backend cpu selects the reference's own CPU loops "
                               "(GeomVertexData::update_animated_vertices); panda3d_b200 has no CPU fallback")
        if self._animated is None:                       # geomVertexData.cxx:1448-1453
            post = self._format.get_post_animated_format()
            self._animated = GeomVertexData(self._name, post, min(self._usage_hint, GeomEnums.UH_dynamic))
        static_key = (self._modified.seq, getattr(self._tbt, "_static_modified", UpdateSeq()).seq if self._tbt else 0,
                      id(self._tbt), id(self._sliders))
        if self._device is None or self._device[0] != static_key:
            if self._device is not None:
                self._device[2].close()
                self._device[1].close()
            mesh, palette = self.flatten()
            dmesh = _capi.Mesh(_Backend.context(), mesh)
            self._device = (static_key, dmesh, _capi.Group(dmesh, 1), palette)
        _, dmesh, group, palette = self._device
        pal = np.stack([t.get_matrix().reshape(16) for t in palette]) if palette else np.zeros((0, 16), np.float32)
        st = self._sliders
        sl = np.asarray([st.get_slider(i).get_slider() for i in range(st.get_num_sliders())], np.float32) if st else None
        outs = group.animate_vertices(pal, sl)
        self._animated._arrays = outs                    # same object, new contents (geomVertexData.cxx:895-911)
        self._animated._touch()
