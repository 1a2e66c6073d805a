#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Builds the UNMODIFIED reference (Panda3D 1.11.0) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored),
# then the harness oracle/ref_harness.cxx against it.  Outputs:
#   oracle/_ref/lib/*.so        the reference core (libpanda, libpandaegg, ...)
#   oracle/_ref/bin/ref_harness the driver used to pin the oracle / time the CPU path
# The recipe is SURVEY.md section 8c's (verified: 71 ninja steps, ~90 s on 8 vCPU).
# The reference's path does not compile from "a few source files" (libpanda is
# ~1500 translation units with generated headers), so its own CMake is used,
# out of tree; nothing is copied from /root/reference into the repository.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${P3D_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
BUILD="$OUT/build"
[ -d "$REF/panda/src/gobj" ] || { echo "reference tree not found at $REF" >&2; exit 3; }
mkdir -p "$BUILD" "$OUT/lib" "$OUT/bin"

if [ ! -f "$BUILD/lib/libpandaegg.so" ]; then
  ( cd "$BUILD" && cmake "$REF" -G Ninja -DCMAKE_BUILD_TYPE=Release -DBUILD_DIRECT=OFF \
      -DBUILD_PANDATOOL=OFF -DBUILD_CONTRIB=OFF -DBUILD_MODELS=OFF -DBUILD_TESTS=OFF \
      -DBUILD_TOOLS=OFF -DHAVE_PYTHON=OFF -DINTERROGATE_PYTHON_INTERFACE=OFF -DHAVE_AUDIO=OFF \
      > cmake.log 2>&1 && ninja panda pandaegg > ninja.log 2>&1 )
fi
cp -a "$BUILD"/lib/*.so* "$OUT/lib/"

DIRS="dtool/src/dtoolbase dtool/src/dtoolutil dtool/src/prc dtool/src/dconfig
 panda/src/pandabase panda/src/express panda/src/putil panda/src/linmath panda/src/mathutil
 panda/src/pipeline panda/src/gobj panda/src/event panda/src/pstatclient panda/src/pnmimage
 panda/src/gsgbase panda/src/pgraph panda/src/chan panda/src/char panda/src/egg2pg panda/src/egg
 panda/src/downloader panda/src/nativenet"
INC="-I$BUILD/include"
for d in $DIRS; do INC="$INC -I$REF/$d -I$BUILD/cmake/$d -I$BUILD/$d"; done

g++ -std=gnu++17 -O2 -fno-rtti -fno-exceptions $INC "$HERE/ref_harness.cxx" \
    -L"$OUT/lib" -lpandaegg -lpanda -lpandaexpress -lp3dtool -lp3prc -lpthread \
    -Wl,-rpath,'$ORIGIN/../lib' -o "$OUT/bin/ref_harness"
echo "built $OUT/bin/ref_harness"
