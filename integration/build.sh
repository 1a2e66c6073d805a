#!/usr/bin/env bash
# Builds the Panda-side shim (cudaSkinBackend) and its check driver against the real Panda3D
# built by oracle/build_ref.sh.  Outputs into integration/_build/ (git-ignored, travels to the GPU
# box): panda_cuda_check + the two egg files of config C1 it is run on.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${P3D_REFERENCE:-/root/reference}"
REFBUILD="$ROOT/oracle/_ref/build"
OUT="$HERE/_build"
[ -f "$ROOT/oracle/_ref/lib/libpanda.so" ] || bash "$ROOT/oracle/build_ref.sh"
mkdir -p "$OUT/models"
DIRS="dtool/src/dtoolbase dtool/src/dtoolutil dtool/src/prc dtool/src/dconfig
 panda/src/pandabase panda/src/express panda/src/putil panda/src/linmath panda/src/mathutil
 panda/src/pipeline panda/src/gobj panda/src/event panda/src/pstatclient panda/src/pnmimage
 panda/src/gsgbase panda/src/pgraph panda/src/chan panda/src/char panda/src/egg2pg panda/src/egg
 panda/src/downloader panda/src/nativenet"
INC="-I$REFBUILD/include -I$ROOT/include -I$HERE"
for d in $DIRS; do INC="$INC -I$REF/$d -I$REFBUILD/cmake/$d -I$REFBUILD/$d"; done
g++ -std=gnu++17 -O2 -fno-rtti -fno-exceptions $INC "$HERE/cudaSkinBackend.cxx" "$HERE/panda_cuda_check.cxx" \
    -L"$ROOT/oracle/_ref/lib" -lpandaegg -lpanda -lpandaexpress -lp3dtool -lp3prc -lpthread -ldl \
    -Wl,-rpath,'$ORIGIN/../../oracle/_ref/lib' -o "$OUT/panda_cuda_check"
cp "$REF/models/panda-model.egg" "$REF/models/panda-walk4.egg" "$REF/models/ripple.egg" "$OUT/models/"
echo "built $OUT/panda_cuda_check"
