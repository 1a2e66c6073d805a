"""Seeded synthetic skinned meshes of the BASELINE.json shapes.

The vertex formats follow what the reference's egg loader emits
(panda/src/egg2pg/eggLoader.cxx:2287-2369): array 0 = vertex(3f) [normal(3f)]
[texcoord(2f)] [tangent(3f) binormal(3f)], array 1 = transform_blend(1 x u16)
followed by the dense C_morph_delta columns; with `align16` the layout is what
GeomVertexFormat::align_columns_for_animation produces (4 x f32 @ 16 B,
panda/src/gobj/geomVertexArrayFormat.cxx:349-373).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import flat as F
from .flat import Column, FlatMesh, Morph


@dataclass
class SynthMesh:
    flat: FlatMesh
    columns: List[Tuple[str, Column]]           # every column of the source format, with its name
    slider_names: List[str] = field(default_factory=list)
    slider_rows: List[np.ndarray] = field(default_factory=list)   # int32 [k,2] per slider
    name: str = "synthetic"

    def recipe(self, palettes: np.ndarray, sliders: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        """The recipe oracle/ref_harness.cxx `run`/`bench` rebuilds real Panda objects from."""
        m = self.flat
        frames = palettes.shape[0]
        if sliders is None:
            sliders = np.zeros((frames, m.num_sliders), np.float32)
        case: Dict[str, np.ndarray] = {
            "meta": np.asarray([m.num_rows, len(m.arrays), len(m.output_arrays), m.num_transforms,
                                m.num_blends, m.num_sliders, m.coordinate_system, frames], np.int32),
            "columns": np.asarray([c.as_tuple() for _, c in self.columns], np.int32).reshape(-1, 5),
            "column_names": np.frombuffer(("\n".join(n for n, _ in self.columns) + "\n").encode(), np.uint8),
            "src_strides": np.asarray(m.strides, np.int32),
            "blend_column": np.asarray(m.blend_column.as_tuple() if m.blend_column else (-1, 0, 0, 0, 0), np.int32),
            "blend_offsets": m.blend_offsets, "blend_transform": m.blend_transform, "blend_weight": m.blend_weight,
            "rows": np.asarray(m.row_ranges, np.int32).reshape(-1, 2),
            "palette": np.ascontiguousarray(palettes, np.float32).reshape(frames, m.num_transforms, 16),
            "sliders": np.ascontiguousarray(sliders, np.float32).reshape(frames, m.num_sliders),
        }
        for a, arr in enumerate(m.arrays):
            case[f"src_array{a}"] = arr
        if m.num_sliders:
            case["slider_names"] = np.frombuffer(("\n".join(self.slider_names) + "\n").encode(), np.uint8)
            off = np.zeros(m.num_sliders + 1, np.int32)
            for s, r in enumerate(self.slider_rows):
                off[s + 1] = off[s] + len(r)
            case["slider_row_offsets"] = off
            case["slider_rows"] = np.concatenate([np.asarray(r, np.int32).reshape(-1, 2) for r in self.slider_rows]) \
                if self.slider_rows else np.zeros((0, 2), np.int32)
        return case


def _unit(v: np.ndarray) -> np.ndarray:
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def _rotation(hpr_deg: np.ndarray) -> np.ndarray:
    """Z-up right-handed heading/pitch/roll rotation (row-vector convention), float64."""
    h, p, r = np.radians(hpr_deg[:, 0]), np.radians(hpr_deg[:, 1]), np.radians(hpr_deg[:, 2])
    n = len(h)
    def rot(axis, a):
        c, s = np.cos(a), np.sin(a)
        m = np.zeros((n, 3, 3))
        m[:, 0, 0] = m[:, 1, 1] = m[:, 2, 2] = 1
        i, j = {0: (1, 2), 1: (2, 0), 2: (0, 1)}[axis]
        m[:, i, i] = c; m[:, i, j] = s; m[:, j, i] = -s; m[:, j, j] = c
        return m
    return rot(1, r) @ rot(0, p) @ rot(2, h)


def make_palettes(num_transforms: int, frames: int, seed: int, *, scale=1.0, jitter: float = 0.0) -> np.ndarray:
    """Joint matrices = scale * rotation(hpr) with translation in row 3 (LMatrix4f layout).

    `scale` is a float (uniform) or a 3-tuple; `jitter` adds a per-joint relative scale noise.
    Every frame re-randomises every joint, so every blend is dirty every frame.
    """
    rng = np.random.default_rng(seed)
    n = frames * num_transforms
    hpr = np.stack([rng.uniform(-180, 180, n), rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)], 1)
    rot = _rotation(hpr)
    sc = np.broadcast_to(np.asarray(scale, np.float64).reshape(-1), (3,)) if np.ndim(scale) else np.full(3, float(scale))
    sc = sc[None, :] * (1.0 + jitter * rng.uniform(-1, 1, (n, 1)))
    m = np.zeros((n, 4, 4))
    m[:, :3, :3] = rot * sc[:, :, None]
    m[:, 3, :3] = rng.uniform(-1, 1, (n, 3))
    m[:, 3, 3] = 1
    return m.astype(np.float32).reshape(frames, num_transforms, 16)


def make_blend_table(num_transforms: int, num_blends: int, weights_per_blend: int, rng) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """`num_blends` unique blends of `weights_per_blend` distinct transforms, entries in ascending
    transform order (= ascending VertexTransform* order in the reference harness), raw weights in
    [0.05, 1.05) normalised the way TransformBlend::normalize_weights does (transformBlend.cxx:122-137)."""
    k = min(weights_per_blend, num_transforms)
    if k == 1 and num_blends > num_transforms:
        raise ValueError("a rigid (1-transform) table has at most num_transforms unique blends")
    seen = set()
    xf = np.zeros((num_blends, k), np.int32)
    w = np.zeros((num_blends, k), np.float32)
    b = 0
    while b < num_blends:
        joints = np.sort(rng.choice(num_transforms, size=k, replace=False)).astype(np.int32)
        raw = (0.05 + rng.integers(0, 1000, size=k) / 1000.0).astype(np.float32)
        net = np.float32(0)
        for x in raw:
            net = np.float32(net + x)
        ww = (raw / net).astype(np.float32)
        key = (joints.tobytes(), ww.tobytes())
        if key in seen:
            continue
        seen.add(key)
        xf[b], w[b] = joints, ww
        b += 1
    offsets = (np.arange(num_blends + 1) * k).astype(np.int32)
    return offsets, xf.reshape(-1), w.reshape(-1)


def make_mesh(num_rows: int, num_transforms: int, num_blends: int, *, seed: int = 1234,
              weights_per_blend: int = 4, tbn: bool = False, texcoord: bool = False,
              align16: bool = False, index_type: str = "u16", index_order: str = "random",
              rows: Optional[Sequence[Tuple[int, int]]] = None,
              num_sliders: int = 0, morph_columns: Sequence[str] = ("vertex",),
              slider_band: float = 0.04, coordinate_system: int = F.CS_zup_right,
              float64: bool = False, name: str = "synthetic") -> SynthMesh:
    """`float64`: the animated and morph-delta columns are NT_float64 (a `vertices-float64` style
    format): the reference then takes its GeomVertexRewriter branches (geomVertexData.cxx:1782-1799)."""
    rng = np.random.default_rng(seed)
    nv = 4 if align16 else 3
    fnt = F.NT_float64 if float64 else F.NT_float32
    fb = 8 if float64 else 4
    cols: List[Tuple[str, Column]] = []
    start = 0

    def add(name_, n, contents, array=0, nt=F.NT_float32, align=4):
        nonlocal start
        start = (start + align - 1) // align * align
        c = Column(array, start, n, nt, contents)
        cols.append((name_, c))
        start += n * F.NUMERIC_BYTES[nt]
        return c

    al = 16 if align16 else fb
    c_vertex = add("vertex", nv, F.C_point, nt=fnt, align=al)
    c_normal = add("normal", nv, F.C_normal, nt=fnt, align=al)
    if texcoord:
        add("texcoord", 2, F.C_texcoord)
    c_tan = c_bin = None
    if tbn:
        c_tan = add("tangent", nv, F.C_vector, nt=fnt, align=al)
        c_bin = add("binormal", nv, F.C_vector, nt=fnt, align=al)
    stride0 = (start + al - 1) // al * al
    arr0 = np.zeros((num_rows, stride0), np.uint8)
    f0 = arr0.view(np.float32)

    def put(col: Column, data: np.ndarray, w: float):
        if float64:   # doubles that are NOT float-representable, so float round trips would show
            vals = np.zeros((num_rows, col.num_values), np.float64)
            vals[:, :3] = data.astype(np.float64) * (1.0 + 1e-9)
            if col.num_values == 4:
                vals[:, 3] = w * (1.0 + 1e-9)
            arr0[:, col.start:col.start + 8 * col.num_values] = vals.view(np.uint8)
            return
        o = col.start // 4
        f0[:, o:o + 3] = data
        if col.num_values == 4:
            f0[:, o + 3] = w

    put(c_vertex, rng.uniform(-1, 1, (num_rows, 3)).astype(np.float32), 1.0)
    put(c_normal, _unit(rng.normal(size=(num_rows, 3))), 0.0)
    if texcoord:
        o = dict(cols)["texcoord"].start // 4
        f0[:, o:o + 2] = rng.uniform(0, 1, (num_rows, 2)).astype(np.float32)
    if tbn:
        put(c_tan, _unit(rng.normal(size=(num_rows, 3))), 0.0)
        put(c_bin, _unit(rng.normal(size=(num_rows, 3))), 0.0)

    # array 1: transform_blend index (+ dense morph delta columns)
    start = 0
    nt = {"u8": F.NT_uint8, "u16": F.NT_uint16, "u32": F.NT_uint32}[index_type]
    ib = F.NUMERIC_BYTES[nt]
    c_blend = add("transform_blend", 1, F.C_index, array=1, nt=nt, align=ib)
    slider_names, slider_rows, morph_cols = [], [], []
    base_by_name = {"vertex": c_vertex, "normal": c_normal, "tangent": c_tan, "binormal": c_bin}
    for s in range(num_sliders):
        sname = f"s{s}"
        slider_names.append(sname)
        band = max(1, int(round(slider_band * num_rows)))
        b0 = min(num_rows - 1, s * num_rows // max(1, num_sliders))
        slider_rows.append(np.asarray([[b0, min(num_rows, b0 + band)]], np.int32))
        for base in morph_columns:
            morph_cols.append((s, base, add(f"{base}.morph.{sname}", 3, F.C_morph_delta, array=1, nt=fnt, align=fb)))
    stride1 = max(ib, (start + fb - 1) // fb * fb) if morph_cols else ib
    arr1 = np.zeros((num_rows, stride1), np.uint8)

    if index_order == "sorted":
        bidx = (np.arange(num_rows, dtype=np.int64) * num_blends // num_rows)
    elif index_order == "runs":     # short runs, like the egg loader's grouped vertices (avg ~7)
        runs = rng.integers(0, num_blends, size=num_rows // 6 + 2)
        bidx = np.repeat(runs, rng.integers(1, 13, size=runs.size))[:num_rows]
        if bidx.size < num_rows:
            bidx = np.resize(bidx, num_rows)
    elif index_order.startswith("fixed"):   # runs of exactly N rows, e.g. "fixed5" (kernel experiments)
        n = int(index_order[5:])
        bidx = np.repeat(rng.integers(0, num_blends, size=num_rows // n + 1), n)[:num_rows]
    else:
        bidx = rng.integers(0, num_blends, size=num_rows)
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[ib]
    arr1[:, :ib] = bidx.astype(dt).reshape(-1, 1).view(np.uint8)

    morphs: List[Morph] = []
    for s, base, dcol in morph_cols:
        r = slider_rows[s]
        d = np.zeros((num_rows, 3), np.float32)
        b0, b1 = int(r[0, 0]), int(r[0, 1])
        d[b0:b1] = rng.uniform(-0.05, 0.05, (b1 - b0, 3)).astype(np.float32)
        if float64:
            arr1[:, dcol.start:dcol.start + 24] = (d.astype(np.float64) * (1.0 + 1e-9)).view(np.uint8)
        else:
            arr1[:, dcol.start:dcol.start + 12] = d.view(np.uint8)
        morphs.append(Morph(s, base_by_name[base], dcol, r))
    # the reference iterates morph column outer, slider inner; the harness re-exports its own
    # order, which only matters when two morphs hit the same base value of one row
    offsets, xf, w = make_blend_table(num_transforms, num_blends, weights_per_blend, rng)
    skin = [c_vertex] + [c for c in (c_normal, c_tan, c_bin) if c is not None]
    rr = np.asarray(rows if rows is not None else [(0, num_rows)], np.int32).reshape(-1, 2)
    fm = FlatMesh(num_rows=num_rows, arrays=[arr0, arr1], is_output=[True, False], skin_columns=skin,
                  blend_column=c_blend, num_transforms=num_transforms, blend_offsets=offsets,
                  blend_transform=xf, blend_weight=w, row_ranges=rr, num_sliders=num_sliders,
                  morphs=morphs, coordinate_system=coordinate_system)
    return SynthMesh(fm, cols, slider_names, slider_rows, name)


def _random_column(rng, num_rows: int, nv: int, nt: int, contents: int, *, delta: bool = False) -> np.ndarray:
    """uint8 [num_rows, column bytes] of plausible values for one column: finite, and far enough inside the
    integer types' ranges that an animated value still fits the C conversion the reference stores it with."""
    unit = contents in (F.C_vector, F.C_normal)
    if nt in (F.NT_float32, F.NT_float64):
        if delta:
            v = rng.uniform(-0.3, 0.3, (num_rows, nv))
        elif contents == F.C_color:
            v = rng.uniform(0, 1, (num_rows, nv))
        elif unit:
            v = rng.normal(size=(num_rows, nv))
            v /= np.linalg.norm(v[:, :min(nv, 3)], axis=1, keepdims=True)
        else:
            v = rng.uniform(-1, 1, (num_rows, nv))
            if nv == 4:
                v[:, 3] = rng.uniform(0.5, 2.0, num_rows)
        if nt == F.NT_float64:
            return np.ascontiguousarray(v.astype(np.float32).astype(np.float64) * (1.0 + 1e-9)).view(np.uint8).reshape(num_rows, -1)
        return np.ascontiguousarray(v.astype(np.float32)).view(np.uint8).reshape(num_rows, -1)
    if nt in (F.NT_packed_dcba, F.NT_packed_dabc):
        return rng.integers(0, 256, (num_rows, 4), dtype=np.uint8)
    if nt == F.NT_packed_ufloat:
        # r11 g11 b10 with exponents 0 (denormals, zero) .. 17: finite, and small enough to stay finite when animated
        r = rng.integers(0, 18 << 6, num_rows).astype(np.uint32)
        g = rng.integers(0, 18 << 6, num_rows).astype(np.uint32)
        b = rng.integers(0, 18 << 5, num_rows).astype(np.uint32)
        return np.ascontiguousarray(r | (g << 11) | (b << 22)).view(np.uint8).reshape(num_rows, 4)
    dt, lo, hi = {F.NT_uint8: (np.uint8, 0, 256), F.NT_int8: (np.int8, -128, 128),
                  F.NT_uint16: (np.uint16, 0, 40000), F.NT_int16: (np.int16, -9000, 9000),
                  F.NT_uint32: (np.uint32, 0, 1 << 29), F.NT_int32: (np.int32, -(1 << 28), 1 << 28)}[nt]
    if delta:
        lo, hi = (0, 12) if lo == 0 else (-12, 12)
    return np.ascontiguousarray(rng.integers(lo, hi, (num_rows, nv)).astype(dt)).view(np.uint8).reshape(num_rows, -1)


def make_typed_mesh(num_rows: int, num_transforms: int, num_blends: int, *, seed: int, columns, morphs=(),
                    num_sliders: int = 0, slider_band: float = 0.5, weights_per_blend: int = 4,
                    index_order: str = "runs", rows: Optional[Sequence[Tuple[int, int]]] = None,
                    coordinate_system: int = F.CS_zup_right, name: str = "typed") -> SynthMesh:
    """A mesh whose columns are of ANY numeric type: the reference animates everything that is not float32 x 3 / x 4
    through GeomVertexRewriter (geomVertexData.cxx:1782-1799, 1862-1879) and morphs every base column through it
    (:1489-1537), so integer, packed and short columns go through the GeomVertexColumn::Packer family.

    columns: (name, num_values, numeric_type, contents[, alignment]) of array 0, in order; C_point / C_vector /
             C_normal columns are skinned.  num_values counts VALUES (4 for the packed colour words, 3 for packed_ufloat).
    morphs:  (base column name, delta num_values, delta numeric_type): every slider gets one such delta column in
             array 1, overlapping bands of rows, so rows are hit by several morphs in a row.
    """
    rng = np.random.default_rng(seed)
    cols: List[Tuple[str, Column]] = []
    by_name: Dict[str, Column] = {}
    blocks = []
    start = 0
    for spec in columns:
        cname, nv, nt, contents = spec[:4]
        cb = F.NUMERIC_BYTES[nt]
        align = spec[4] if len(spec) > 4 else max(cb, 4)
        start = (start + align - 1) // align * align
        c = Column(0, start, nv, nt, contents)
        cols.append((cname, c))
        by_name[cname] = c
        blocks.append((c, _random_column(rng, num_rows, nv, nt, contents)))
        start += c.num_bytes
    stride0 = (start + 3) // 4 * 4
    arr0 = np.zeros((num_rows, stride0), np.uint8)
    for c, data in blocks:
        arr0[:, c.start:c.start + c.num_bytes] = data

    start = 0
    c_blend = Column(1, 0, 1, F.NT_uint16, F.C_index)
    cols.append(("transform_blend", c_blend))
    start = 2
    slider_names, slider_rows, morph_list, dblocks = [], [], [], []
    band = max(1, int(round(slider_band * num_rows)))
    for sidx in range(num_sliders):
        sname = f"s{sidx}"
        slider_names.append(sname)
        b0 = min(num_rows - 1, sidx * num_rows // max(1, 2 * num_sliders))
        r = np.asarray([[b0, min(num_rows, b0 + band)]], np.int32)
        slider_rows.append(r)
        for base, dnv, dnt in morphs:
            cb = F.NUMERIC_BYTES[dnt]
            start = (start + cb - 1) // cb * cb
            dc = Column(1, start, dnv, dnt, F.C_morph_delta)
            cols.append((f"{base}.morph.{sname}", dc))
            data = _random_column(rng, num_rows, dnv, dnt, F.C_morph_delta, delta=True)
            data[:int(r[0, 0])] = 0
            data[int(r[0, 1]):] = 0
            dblocks.append((dc, data))
            morph_list.append(Morph(sidx, by_name[base], dc, r))
            start += dc.num_bytes
    stride1 = max(2, (start + 3) // 4 * 4) if dblocks else 2
    arr1 = np.zeros((num_rows, stride1), np.uint8)
    for dc, data in dblocks:
        arr1[:, dc.start:dc.start + dc.num_bytes] = data

    if index_order == "runs":
        runs = rng.integers(0, num_blends, size=num_rows // 6 + 2)
        bidx = np.resize(np.repeat(runs, rng.integers(1, 13, size=runs.size)), num_rows)
    else:
        bidx = rng.integers(0, num_blends, size=num_rows)
    arr1[:, :2] = bidx.astype(np.uint16).reshape(-1, 1).view(np.uint8)

    offsets, xf, w = make_blend_table(num_transforms, num_blends, weights_per_blend, rng)
    order = {F.C_point: 0, F.C_normal: 1, F.C_vector: 1}
    skin = sorted((c for _, c in cols if c.array == 0 and c.contents in order), key=lambda c: order[c.contents])
    rr = np.asarray(rows if rows is not None else [(0, num_rows)], np.int32).reshape(-1, 2)
    fm = FlatMesh(num_rows=num_rows, arrays=[arr0, arr1], is_output=[True, False], skin_columns=skin,
                  blend_column=c_blend, num_transforms=num_transforms, blend_offsets=offsets,
                  blend_transform=xf, blend_weight=w, row_ranges=rr, num_sliders=num_sliders,
                  morphs=morph_list, coordinate_system=coordinate_system)
    return SynthMesh(fm, cols, slider_names, slider_rows, name)


def make_sliders(num_sliders: int, frames: int, seed: int) -> np.ndarray:
    """Slider values U(0,1) with 25 % forced to exactly 0 (the `!= 0` skip, geomVertexData.cxx:1483)."""
    rng = np.random.default_rng(seed)
    v = rng.uniform(0, 1, (frames, num_sliders)).astype(np.float32)
    v[rng.uniform(0, 1, (frames, num_sliders)) < 0.25] = 0.0
    return v


# ---- the BASELINE.json configurations (full sizes) -------------------------------------

def config_c2(index_order: str = "random", **kw) -> SynthMesh:
    """configs[1]: 100k-vertex, 64-joint, 4-weight skinned mesh (16 384 unique blends)."""
    return make_mesh(100_000, 64, 16_384, seed=1234, index_order=index_order, name="C2", **kw)


def config_c3_mesh(**kw) -> SynthMesh:
    """configs[2]: the 20k-vertex, 128-joint character of the 10k-instance crowd (4096 blends)."""
    kw.setdefault("index_order", "runs")
    return make_mesh(20_000, 128, 4096, seed=33, name="C3", **kw)


def config_c4(**kw) -> SynthMesh:
    """configs[3]: 250k-vertex face mesh, 64 morph sliders + 4-weight skinning."""
    kw.setdefault("index_order", "runs")
    return make_mesh(250_000, 64, 16_384, seed=44, num_sliders=64, name="C4", **kw)


def config_c5_meshes() -> List[Tuple[SynthMesh, int]]:
    """configs[4]: 4 character types with normals+tangents+binormals, 2 M vertices/frame total."""
    shapes = [(5_000, 32, 1_000, 100), (20_000, 64, 4_000, 25), (50_000, 128, 8_000, 10), (125_000, 256, 16_000, 4)]
    out = []
    for i, (n, t, b, count) in enumerate(shapes):
        out.append((make_mesh(n, t, b, seed=500 + i, tbn=True, texcoord=True, index_order="runs", name=f"C5.{i}"), count))
    return out


def make_skeleton(num_joints: int, num_frames: int, seed: int, *, palette_joints: Optional[int] = None,
                  coordinate_system: int = F.CS_zup_right):
    """A random joint tree (parents-first) animated by hpr + xyz tables of `num_frames` frames, constant
    scale tables, no shear -- the table shapes an egg `<Xfm$Anim_S$>` bundle has (cf. panda-walk4.egg).
    initial_inverse is the inverse of the net transform at frame 0, so frame 0 skins to the bind pose."""
    from .flat import FlatSkeleton
    rng = np.random.default_rng(seed)
    J = num_joints
    parent = np.asarray([-1] + [int(rng.integers(max(0, j - 16), j)) for j in range(1, J)], np.int32)   # depth ~ J/8, like a humanoid rig
    tlen = np.zeros((J, 12), np.int32)
    toff = np.zeros((J, 12), np.int32)
    tables = []
    pos = 0
    for j in range(J):
        for i in range(12):
            if i < 3:
                vals = np.full(1, 1.0, np.float32)                       # constant unit scale
            elif i < 6:
                vals = np.zeros(0, np.float32)                           # no shear table
            elif i < 9:
                vals = rng.uniform(-60, 60, num_frames).astype(np.float32)  # h p r in degrees
            else:
                vals = (rng.uniform(-0.3, 0.3) + 0.05 * rng.standard_normal(num_frames)).astype(np.float32)
            tlen[j, i], toff[j, i] = len(vals), pos
            tables.append(vals)
            pos += len(vals)
    tables = np.concatenate(tables).astype(np.float32)
    ident = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (J, 1))
    skel = FlatSkeleton(parent=parent, has_channel=np.ones(J, np.int32), table_length=tlen, table_offset=toff, tables=tables,
                        default_value=ident.copy(), initial_inverse=ident.copy(), root_xform=np.eye(4, dtype=np.float32).reshape(16),
                        palette_joint=np.arange(palette_joints if palette_joints is not None else J, dtype=np.int32) % J,
                        coordinate_system=coordinate_system)
    return skel
