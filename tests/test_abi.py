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


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/afs_b200.h must be usable from a C host (no C++ or torch types in any signature): a C99 translation
    unit that includes it, takes the address of every declared entry point and calls the GPU-free ones compiles with
    gcc, links against libafs_b200.so and runs."""
    from audio_fewshot_b200 import build
    lib = build.build()
    names = declared_symbols()
    src = tmp_path / "host.c"
    src.write_text(
        '#include <stdio.h>\n#include "afs_b200.h"\n'
        "typedef void (*fn_t)(void);\n"
        "int main(void) {\n"
        "  fn_t fns[] = {%s};\n"
        "  size_t n = sizeof(fns) / sizeof(fns[0]);\n"
        "  for (size_t i = 0; i < n; ++i) if (fns[i] == 0) return 2;\n"
        "  if (afs_abi_version() != AFS_ABI_VERSION) return 3;\n"
        "  if (afs_proto_fwd(0, 0, 0, 0, 0, 5, 5, 1600, 0, 0, 0, 0, 0, 0) != AFS_ERR_INVALID_ARG) return 4;\n"
        '  printf("%%s %%zu\\n", afs_status_string(AFS_ERR_WORKSPACE), n);\n'
        "  return 0;\n}\n" % ", ".join("(fn_t)%s" % n for n in names))
    exe = tmp_path / "host"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), lib, "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out
    assert out.stdout.split() == ["workspace", "too", "small", str(len(names))]
