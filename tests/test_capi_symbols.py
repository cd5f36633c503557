"""CPU test: the C-ABI library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every function
include/musev_b200.h declares. No compute call is made."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "musev_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mvb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names))


def test_header_declares_expected_surface():
    names = _declared_functions()
    for must in ("mvb_create", "mvb_load_weight", "mvb_finalize", "mvb_workspace_bytes", "mvb_unet_forward",
                 "mvb_fuse_cfg_ddim", "mvb_destroy", "mvb_last_error", "mvb_op_conv_gemm", "mvb_op_attention", "mvb_debug_attention_trace", "mvb_tensor_map_cache_stats"):
        assert must in names


def test_library_loads_and_exports_every_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/musev_b200.h but not exported"
    lib.mvb_version.restype = ctypes.c_int
    assert lib.mvb_version() >= 1


def test_library_is_native_sm100a(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):          # tcgen05.mma / TMA / tcgen05.ld
        assert mnemonic in sass, mnemonic


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "musev_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt, f
