#!/usr/bin/env bash
# Builds a libpanda WITH the vertex-animation-backend hook (integration/geomVertexData.patch: two function
# pointers in GeomVertexData, called from update_animated_vertices / the destructor / clear_animated_vertices),
# so that the CUDA backend is reached through the STOCK callers of GeomVertexData::animate_vertices
# (Geom::get_animated_vertex_data, CullableObject::munge_geom, AnimateVerticesRequest).
# The patch is applied to a scratch copy of the reference (never to /root/reference, nothing copied into the
# repository); the libraries land in oracle/_ref/lib_hook/ (git-ignored, travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${P3D_REFERENCE:-/root/reference}"
SRC="${P3D_HOOK_SRC:-/tmp/p3d_hook_src}"
BUILD="$ROOT/oracle/_ref/build_hook"
OUT="$ROOT/oracle/_ref/lib_hook"
[ -d "$REF/panda/src/gobj" ] || { echo "reference tree not found at $REF" >&2; exit 3; }
if [ ! -f "$SRC/.hook_applied" ]; then
  rm -rf "$SRC"
  cp -a "$REF" "$SRC"
  chmod -R u+w "$SRC"
  patch -p1 -d "$SRC" < "$HERE/geomVertexData.patch"
  touch "$SRC/.hook_applied"
fi
mkdir -p "$BUILD" "$OUT"
if [ ! -f "$BUILD/lib/libpandaegg.so" ]; then
  ( cd "$BUILD" && cmake "$SRC" -G Ninja -DCMAKE_BUILD_TYPE=Release -DBUILD_DIRECT=OFF \
      -DBUILD_PANDATOOL=OFF -DBUILD_CONTRIB=OFF -DBUILD_MODELS=OFF -DBUILD_TESTS=OFF \
      -DBUILD_TOOLS=OFF -DHAVE_PYTHON=OFF -DINTERROGATE_PYTHON_INTERFACE=OFF -DHAVE_AUDIO=OFF \
      > cmake.log 2>&1 && ninja panda pandaegg > ninja.log 2>&1 )
fi
cp -a "$BUILD"/lib/*.so* "$OUT/"
echo "built $OUT (hooked libpanda from $SRC)"
