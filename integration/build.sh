#!/usr/bin/env bash
# Builds the Panda-side shim (cudaSkinBackend) and its check driver against the Panda3D of
# integration/build_hook.sh: the reference plus the vertex-animation-backend hook
# (integration/geomVertexData.patch, applied to a scratch copy).  Outputs into integration/_build/
# (git-ignored, travels to the GPU box): panda_cuda_check + the egg files of config C1 it is run on.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
SRC="${P3D_HOOK_SRC:-/tmp/p3d_hook_src}"
REFBUILD="$ROOT/oracle/_ref/build_hook"
OUT="$HERE/_build"
[ -f "$ROOT/oracle/_ref/lib_hook/libpanda.so" ] && [ -f "$SRC/.hook_applied" ] || bash "$HERE/build_hook.sh"
mkdir -p "$OUT/models"
DIRS="dtool/src/dtoolbase dtool/src/dtoolutil dtool/src/prc dtool/src/dconfig
 panda/src/pandabase panda/src/express panda/src/putil panda/src/linmath panda/src/mathutil
 panda/src/pipeline panda/src/gobj panda/src/event panda/src/pstatclient panda/src/pnmimage
 panda/src/gsgbase panda/src/pgraph panda/src/chan panda/src/char panda/src/egg2pg panda/src/egg
 panda/src/downloader panda/src/nativenet"
INC="-I$REFBUILD/include -I$ROOT/include -I$HERE"
for d in $DIRS; do INC="$INC -I$SRC/$d -I$REFBUILD/cmake/$d -I$REFBUILD/$d"; done
g++ -std=gnu++17 -O2 -fno-rtti -fno-exceptions $INC "$HERE/cudaSkinBackend.cxx" "$HERE/panda_cuda_check.cxx" \
    -L"$ROOT/oracle/_ref/lib_hook" -lpandaegg -lpanda -lpandaexpress -lp3dtool -lp3prc -lpthread -ldl \
    -Wl,-rpath,'$ORIGIN/../../oracle/_ref/lib_hook' -o "$OUT/panda_cuda_check"
cp "$SRC/models/panda-model.egg" "$SRC/models/panda-walk4.egg" "$SRC/models/ripple.egg" "$OUT/models/"
# ripple.egg with <DRGBA> / <Duv> morphs added: colour and texcoord columns as morph bases (integration/make_morph_egg.py)
python "$HERE/make_morph_egg.py" "$SRC/models/ripple.egg" "$OUT/models/ripple_rgba_uv.egg"
echo "built $OUT/panda_cuda_check"
