#!/bin/bash
# Host setup code (csrc/femb_setup*.cu) under AddressSanitizer + UBSan, no GPU needed: builds a sanitised copy of the library
# next to a scratch directory and runs the hierarchy tests that exercise every host entry point against it.
#   bash scripts/asan_setup.sh [scratch_dir]
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/femb_asan}
mkdir -p "$OUT"
SRC=$(cd "$ROOT" && python -c "
import os, fem_simulation_b200
from fem_simulation_b200 import _lib
print(' '.join(os.path.join('$ROOT/fem-simulation_b200/csrc', s) for s in _lib.SOURCES))")
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -shared -ldl -lgomp -I "$ROOT/include" \
     -Xcompiler -fPIC,-fopenmp,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer $SRC -o "$OUT/libfemb200_asan.so"
cat > "$OUT/run.py" <<PY
import sys; sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
import fem_simulation_b200
from fem_simulation_b200 import _lib
_lib.LIB_PATH = "$OUT/libfemb200_asan.so"
_lib._needs_build = lambda: False
import test_hierarchy as T
for case in ["jag3", "cube12_corrected", "slab_origin"]:
    T.test_one_call_hierarchy_matches_python_build(case)
for case in ["cube40", "jag_sparse", "offset_origin", "corrected"]:
    T.test_native_hierarchy_matches_numpy_restatement(case)
T.test_native_hierarchy_hashed_lattice_and_stencil_table()
T.test_native_hierarchy_random_voxel_sets()
T.test_native_hierarchy_rejects_bad_input()
T.test_host_setup_routines_match_numpy_restatements()
T.test_coo_to_csr_pair_matches_numpy_sort()
T.test_galerkin_plan_emulation()
import test_voxelize as V
for case in ["tetra", "hull", "box_on_lattice", "empty"]:
    V.test_native_inside_test_matches_numpy_restatement(case)
import pathlib, tempfile
V.test_native_obj_reader_matches_python_parser(pathlib.Path(tempfile.mkdtemp()))
V.test_native_mask_mesh_matches_numpy_restatement()
print("host setup under ASan + UBSan: clean")
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 python "$OUT/run.py"
