#!/bin/bash
# TEST INFRASTRUCTURE ONLY: builds oracle/_ref/libmythtracer_ref.so from the reference's own sources.
#
# The sources are read from where they lie under /root/reference (never copied into the repo).  The one
# textual change is made in a throw-away directory under /tmp: `const int MAX_RECURSION_LEVEL = 5;`
# (reference mythtracer.h:11) becomes `extern int MAX_RECURSION_LEVEL;` so that depth 2/3/5/8 configs run
# from one binary (ref_shim.cc defines it, default 5).  SDL2_image is absent here, so texture.cc is built
# unmodified against oracle/sdl_stub (a PPM-backed IMG_Load).  Flags follow the reference Makefile:2-5,15
# (-O3 -std=c++1z -D_USE_MATH_DEFINES -fopenmp, no -march => no FMA contraction).
#
# On the GPU box /root/reference does not exist: the prebuilt .so (git-ignored, but shipped by gpurun)
# is used as is.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${MTB_REFERENCE_DIR:-/root/reference/VerStarting}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF" ]; then
  if [ -f "$OUT/libmythtracer_ref.so" ]; then
    echo "build_ref: $REF absent; keeping prebuilt $OUT/libmythtracer_ref.so"
    exit 0
  fi
  echo "build_ref: $REF absent and no prebuilt library" >&2
  exit 1
fi
if [ -f "$OUT/libmythtracer_ref.so" ] && [ "$OUT/libmythtracer_ref.so" -nt "$HERE/ref_shim.cc" ] &&
   [ "$OUT/libmythtracer_ref.so" -nt "$HERE/sdl_stub/sdl_stub.cc" ] && [ "${1:-}" != "--force" ]; then
  echo "build_ref: up to date"
  exit 0
fi
TMP="$(mktemp -d /tmp/mtb_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
for f in mythtracer objreader octtree primitive_triangle aabb camera texture; do
  cp "$REF/$f.cc" "$TMP/"
done
cp "$REF"/*.h "$TMP/"
sed -i 's/^const int MAX_RECURSION_LEVEL = 5;/extern int MAX_RECURSION_LEVEL;/' "$TMP/mythtracer.h"
grep -q '^extern int MAX_RECURSION_LEVEL;' "$TMP/mythtracer.h"
g++ -O3 -fno-omit-frame-pointer -std=c++1z -D_USE_MATH_DEFINES -fopenmp \
    -fPIC -shared -fno-semantic-interposition -Wl,-Bsymbolic \
    -I "$HERE/sdl_stub" -I "$TMP" \
    "$TMP"/mythtracer.cc "$TMP"/objreader.cc "$TMP"/octtree.cc "$TMP"/primitive_triangle.cc \
    "$TMP"/aabb.cc "$TMP"/camera.cc "$TMP"/texture.cc \
    "$HERE/ref_shim.cc" "$HERE/sdl_stub/sdl_stub.cc" \
    -o "$OUT/libmythtracer_ref.so" 2> "$OUT/build.log" || { cat "$OUT/build.log" >&2; exit 1; }
echo "build_ref: built $OUT/libmythtracer_ref.so"
