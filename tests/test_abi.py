"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/afs_b200.h declares."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "afs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(afs_[a-z0-9_]+)\s*\(", text)))


def test_build_and_exports():
    from audio_fewshot_b200 import build
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(handle, name), "missing export %s" % name
    handle.afs_abi_version.restype = ctypes.c_int
    assert handle.afs_abi_version() == 1
    handle.afs_status_string.restype = ctypes.c_char_p
    assert handle.afs_status_string(-4) == b"workspace too small"


def test_binding_covers_header():
    from audio_fewshot_b200 import _lib
    assert sorted(_lib.SIGNATURES.keys()) == declared_symbols()
    h = _lib.lib()  # loads and type-checks every symbol
    assert h.afs_proto_workspace_bytes(1, 5, 5, 1600) == 0
    assert h.afs_proto_workspace_bytes(2, 5, 1, 12800) == 2 * 5 * 12800 * 4
    assert h.afs_dn4_workspace_bytes(100, 1, 5, 5, 64, 20) > 100 * 64 * 20 * 4


def test_sass_is_sm100a():
    from audio_fewshot_b200 import build
    path = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_invalid_arguments_are_rejected_without_a_gpu():
    from audio_fewshot_b200 import _lib
    h = _lib.lib()
    # null pointers / bad sizes must be refused before any CUDA call
    assert h.afs_proto_fwd(None, 0, None, 0, 0, 5, 5, 1600, 0, None, None, None, 0, None) == -1
    assert h.afs_vote_acc(None, 5, None, 0, None, 1, None, None, None, None) == -1
    assert h.afs_bdc_fwd(None, 1, 64, 10, None, 1, None, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "audio_fewshot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f
