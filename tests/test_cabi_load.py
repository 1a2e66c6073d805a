"""CPU-side checks of the C-ABI library: it loads, exports every symbol of include/p3cudaskin.h,
and fails loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes
import os
import re

import pytest

from panda3d_b200 import _capi, build_ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    build_ext.build()
    return _capi.lib()


def test_header_symbols_all_exported(library):
    header = open(os.path.join(ROOT, "include", "p3cudaskin.h")).read()
    declared = sorted(set(re.findall(r"P3CS_EXPORT[^;]*?\b(p3cs_\w+)\s*\(", header)))
    assert declared, "no declarations found in the header"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(library, name), f"{name} declared in include/p3cudaskin.h but not exported"


def test_api_version_and_plugin_table(library):
    assert library.p3cs_api_version() == 3
    assert library.p3cs_get_api()  # non-NULL table of function pointers (dlsym("p3cs_get_api") convention)


def test_struct_layout_matches_header(library):
    # p3cs_mesh_desc carries struct_size; a mismatch between the ctypes mirror and the C struct is rejected
    assert ctypes.sizeof(_capi._MeshDesc) == 128
    assert ctypes.sizeof(_capi._MorphDesc) == 56
    assert ctypes.sizeof(_capi._ArrayDesc) == 16


def test_no_device_fails_loudly(library):
    if library.p3cs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_capi.P3csError) as e:
        _capi.Context(0)
    assert e.value.code == _capi.P3CS_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU implementation)."""
    pkg = os.path.join(ROOT, "panda3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "skin_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_plain_c_host_loads_the_plugin_table(tmp_path):
    """examples/plugin_host.c: a C99 program that dlopen()s the library and walks p3cs_get_api() -- the way
    Panda3D's load_dso / get_dso_symbol convention would.  Compiled here with gcc."""
    import subprocess
    exe = tmp_path / "plugin_host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "plugin_host.c"), "-ldl", "-o", str(exe)], check=True)
    res = subprocess.run([str(exe), _capi.LIB_PATH], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stderr
    assert "p3cs api version 3" in res.stdout
