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
    assert lib.femb_abi_version() == 2


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
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
