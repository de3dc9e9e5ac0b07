#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never imported by the product path.
#
# Compiles the UNMODIFIED reference CPU/Embree path (BachiLi/redner, `src/*.cpp`) out-of-tree
# into `oracle/_ref/redner<ext-suffix>.so`, straight from the sources where they lie under
# /root/reference (no source is copied into this repository; `oracle/_ref/` is git-ignored).
# The prebuilt Embree 3 / TBB shared objects the reference links against are copied next to
# the module (binaries, not sources) so that the oracle also runs on the GPU box, where
# /root/reference does not exist.
#
# Recipe follows SURVEY.md section 8(c): g++ -O3, vendored Thrust with the CPP device backend
# (no CUDA needed), pip pybind11 (the vendored v2.4 predates Python 3.12).
set -euo pipefail
REF=${REDNER_REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
if [ ! -d "$REF/src" ]; then
    echo "[oracle] $REF not present; keeping prebuilt oracle/_ref as is" >&2
    exit 0
fi
mkdir -p "$OBJ"
PY=${PYTHON:-python}
PYINC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
PBINC=$($PY -c 'import pybind11;print(pybind11.get_include())')
EXT=$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')
TARGET="$OUT/redner$EXT"
SRCS="aabb active_pixels atomic automatic_uv_map bsdf_sample camera camera_distortion channels edge edge_tree \
load_serialized material parallel path_contribution pathtracer pcg_sampler primary_contribution \
primary_intersection rebuild_topology redner scene shape sobol_sampler"
CXXFLAGS="-std=c++14 -O3 -fPIC -fvisibility=hidden -w -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
-I$REF/thrust -I$REF/redner-dependencies/embree/include -I$REF -I$PYINC -I$PBINC"
# The reference's Python package, unmodified, next to its native module (git-ignored like everything under oracle/_ref): the
# GPU box has no /root/reference, and the drop-in test there runs THIS pyredner on redner_b200/dropin/redner.py.
mkdir -p "$OUT"
[ -d "$OUT/pyredner" ] && chmod -R u+w "$OUT/pyredner" && rm -rf "$OUT/pyredner"   # (the checkout is read-only and cp keeps its modes)
cp -r "$REF/pyredner" "$OUT/pyredner"
chmod -R u+w "$OUT/pyredner"
find "$OUT/pyredner" -name __pycache__ -prune -exec rm -rf {} + 2>/dev/null || true
if [ -f "$TARGET" ] && [ -z "${FORCE:-}" ]; then
    echo "[oracle] $TARGET already built"
    exit 0
fi
pids=()
for s in $SRCS; do
    ( g++ $CXXFLAGS -c "$REF/src/$s.cpp" -o "$OBJ/$s.o" ) &
    pids+=($!)
    # at most 8 compiles in flight
    if [ ${#pids[@]} -ge 8 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
done
( gcc -O3 -fPIC -w -c "$REF/src/miniz.c" -o "$OBJ/miniz.o" ) &
pids+=($!)
( g++ $CXXFLAGS -c "$REF/xatlas/xatlas.cpp" -o "$OBJ/xatlas.o" ) &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
cp -f "$REF"/redner-dependencies/embree/lib-linux/libembree3.so.3 \
      "$REF"/redner-dependencies/embree/lib-linux/libtbb.so.2 \
      "$REF"/redner-dependencies/embree/lib-linux/libtbbmalloc.so.2 "$OUT"/
g++ -shared -o "$TARGET" "$OBJ"/*.o -L"$OUT" -l:libembree3.so.3 -lpthread -Wl,--disable-new-dtags -Wl,-rpath,'$ORIGIN'
rm -rf "$OBJ"
echo "[oracle] built $TARGET"
