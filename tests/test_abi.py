"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/femb200.h declares
(no compute calls here -- there is no GPU in the build container)."""
import os
import re

from fem_simulation_b200 import _lib


def _declared(repo_root):
    txt = open(os.path.join(repo_root, "include", "femb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(femb_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(repo_root):
    lib = _lib.load()
    names = _declared(repo_root)
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in femb200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.femb_abi_version() == 5


def test_no_device_is_an_error_not_a_fallback():
    import pytest
    import torch
    from fem_simulation_b200 import Object
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.FembError):
        Object("cube:4", 0.25, [])


def test_sass_has_bulk_copy_and_no_legacy_tensor_path():
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in out            # cp.async.bulk staging of the row pass
    assert "UTMASTG" in out           # TMA tensor store of the element tile (assembly)
    assert "FFMA2" in out             # packed fp32 in the element kernels
    # no kernel spills: ptxas assigned zero local (stack) bytes everywhere
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    local = [int(v) for v in re.findall(r"\bLOCAL:(\d+)", usage)]
    assert len(local) >= 60 and max(local) == 0
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout


def test_ctypes_mirrors_match_header_layout(repo_root, tmp_path):
    """The header compiles as plain C (no CUDA/torch types in the signatures) and the ctypes mirrors in _lib.py have
    the same sizeof / offsetof as the C structs, field by field."""
    import ctypes as C
    import subprocess
    structs = {"FembLevel": _lib.FembLevel, "FembTransfer": _lib.FembTransfer, "FembGalerkinPlan": _lib.FembGalerkinPlan,
               "FembHalo": _lib.FembHalo, "FembHierarchyLevel": _lib.FembHierarchyLevel, "FembHierarchy": _lib.FembHierarchy,
               "FembContext": _lib.FembContext, "FembDynamic": _lib.FembDynamic,
               "FembCloth": _lib.FembCloth}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "femb200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f, _ in cls._fields_:
            lines.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines += ['  printf("consts %d %d %d %d %d %d %d %d\\n", FEMB_ABI_VERSION, FEMB_ASM_CHUNK_HEXES, FEMB_ELEM_K_FLOATS, '
              'FEMB_FOLD_PLAN_BYTES, FEMB_PART_ALL, FEMB_ROW_REC_BYTES, FEMB_MAX_NEIGHBOURS, FEMB_COMM_ID_BYTES);',
              '  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(repo_root, "include"), str(src), "-o", str(exe)],
                   check=True)
    got = dict(l.split(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for f, _ in cls._fields_:
            assert int(got[f"{name}.{f}"]) == getattr(cls, f).offset, f"{name}.{f}"
    assert got["consts"].split() == ["5", str(_lib.ASM_CHUNK_HEXES), str(_lib.ELEM_K_FLOATS), "128", str(_lib.PART_ALL),
                                     str(_lib.ROW_REC_BYTES), str(_lib.MAX_NEIGHBOURS), str(_lib.COMM_ID_BYTES)]


def test_prototypes_match_header_signatures(repo_root):
    """Every ctypes prototype has the arity and per-argument class (int / float / pointer) of its declaration."""
    import ctypes as C
    txt = open(os.path.join(repo_root, "include", "femb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    decls = dict(re.findall(r"\b(femb_\w+)\s*\(([^)]*)\)\s*;", txt))
    assert sorted(decls) == sorted(_lib.PROTOTYPES)

    def klass_c(param):
        param = " ".join(param.split())
        if "*" in param:
            return "p"
        return {"int": "i", "float": "f", "size_t": "z", "long long": "l"}[param.rsplit(" ", 1)[0].replace("const ", "")]

    def klass_ct(t):
        return "i" if t is C.c_int else "f" if t is C.c_float else "z" if t is C.c_size_t else "l" if t is C.c_longlong else "p"

    for name, params in decls.items():
        params = params.strip()
        want = [] if params in ("", "void") else [klass_c(q) for q in params.split(",")]
        assert [klass_ct(t) for t in _lib.PROTOTYPES[name]] == want, name


def test_plain_c_consumer_builds_the_hierarchy(repo_root, tmp_path):
    """examples/setup_from_c.c: a C99 program linked against libfemb200.so (no Python, no torch, no device) calls
    femb_setup_hierarchy_host and reports the level sizes `hierarchy.build_hierarchy` gives for the same cube."""
    import subprocess
    import numpy as np
    from fem_simulation_b200 import hierarchy, mesh
    _lib.load()
    exe = tmp_path / "setup_from_c"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(repo_root, "include"),
                    os.path.join(repo_root, "examples", "setup_from_c.c"), "-L", libdir, "-lfemb200", "-lm", f"-Wl,-rpath,{libdir}",
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "12", "3"], capture_output=True, text=True, check=True).stdout.splitlines()
    H = hierarchy.build_hierarchy(*mesh.cube_mesh(12), 1.0 / 12, layers=3)
    assert out[0] == "abi 5 levels 3"
    for l in range(3):
        colours = " ".join(str(c) for c in np.diff(H.color_vertex_num[l]))
        assert out[1 + l].startswith(f"level {l}: {H.dims[l]} vertices, {len(H.hexs[l])} hexahedra, {H.pattern[l].nnz} blocks, "
                                     f"colours {colours},"), out[1 + l]
    assert f"interpolation entries {H.Us_hat[0].nnz} " in out[1]
    # second mode: a closed surface mesh through the whole native front end (OBJ parser, inside test, kept cells -> mesh,
    # hierarchy), against voxelize.voxelize_obj + build_hierarchy
    from fem_simulation_b200 import voxelize
    obj = tmp_path / "tetra.obj"
    obj.write_text("v 0.013 0.021 0.037\nv 1.071 0.113 0.091\nv 0.233 0.957 0.171\nv 0.317 0.289 1.043\n"
                   "f 1 3 2\nf 1 2 4\nf 2 3 4\nf 1 4 3\n")
    out = subprocess.run([str(exe), str(obj), "0.05", "2"], capture_output=True, text=True, check=True).stdout.splitlines()
    pts, hx = voxelize.voxelize_obj(str(obj), 0.05)
    H = hierarchy.build_hierarchy(pts, hx, 0.05, layers=2)
    assert out[0].startswith("obj: 4 vertices, 4 triangles -> ") and out[0].endswith(f" cells, {len(hx)} kept")
    for l in range(2):
        colours = " ".join(str(c) for c in np.diff(H.color_vertex_num[l]))
        assert out[2 + l].startswith(f"level {l}: {H.dims[l]} vertices, {len(H.hexs[l])} hexahedra, {H.pattern[l].nnz} blocks, "
                                     f"colours {colours},"), out[2 + l]
