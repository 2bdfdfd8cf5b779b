# SPDX-License-Identifier: Apache-2.0
"""The C-ABI library loads without a GPU and exports exactly what include/wcn_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wcn_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wcn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("wcn_hash_insert", "wcn_kernel_map_search", "wcn_kernel_map_scatter",
                 "wcn_gather_gemm", "wcn_wgrad", "wcn_weight_image", "wcn_build_tiles"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from warpconvnet_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(_declared())


def test_no_torch_types_in_abi():
    out = subprocess.run(["nm", "-D", "--defined-only",
                          os.path.join(ROOT, "warpconvnet_b200", "csrc", "libwcn_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(not s.startswith("_ZN2at") and "torch" not in s for s in exported)
    assert set(_declared()) <= set(exported)


def test_version_and_arch():
    from warpconvnet_b200 import _lib
    assert "sm_100a" in _lib.version()
    assert _lib.lib.wcn_built_for_sm100a() == 1


def test_sass_is_sm100a_tcgen05():
    """cuobjdump: the library holds sm_100a code with tcgen05 (UTCHMMA) and bulk-copy (UBLKCP)."""
    so = os.path.join(ROOT, "warpconvnet_b200", "csrc", "libwcn_b200.so")
    r = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in r.stdout
    assert "UTCHMMA" in r.stdout or "UTCMMA" in r.stdout
    assert "UBLKCP" in r.stdout


def test_argument_errors_without_gpu():
    """Pure host-side argument validation returns the documented negative status codes."""
    from warpconvnet_b200._lib import lib
    assert lib.wcn_hash_prepare(None, None, 16, None) == -1
    assert lib.wcn_hash_insert(None, None, None, 4, 16, None, None) == -1
    assert lib.wcn_kernel_map_num_blocks(1000) == 4
    assert lib.wcn_gather_gemm(None, 0, 0, None, None, 0, None, None, None, None, None, 4, 128, 512,
                               27, 1, 64, 64, 0, None, 0, 0, 0, None, 0, None, None) == -1
    assert lib.wcn_wgrad(None, 0, None, 0, None, None, None, None, 27, 1, 64, 64, 0, 1.0, 0, 0,
                         None, 0, 1, 1, -1, None, 0, 0, None) == -1
    assert lib.wcn_depthwise_conv(None, 0, None, 0, None, None, None, 8, 27, 64, 0, 0, 0,
                                  None) == -1
    assert lib.wcn_bn_stats(None, 0, 8, 64, 0, None, None) == -1
    assert lib.wcn_depthwise_conv_plan(None, 0, None, 0, None, None, None, None, None, None, 4, 256,
                                       27, 64, 0, 0, 0, None) == -1
    n_slabs, gps = ctypes.c_int(0), ctypes.c_int(0)
    nbytes = lib.wcn_weight_image_bytes(27, 1, 64, 128, 0, 0, ctypes.byref(n_slabs),
                                        ctypes.byref(gps))
    assert nbytes == 27 * 1 * 128 * 128 and n_slabs.value == 1 and gps.value == 1
    nbytes = lib.wcn_weight_image_bytes(27, 64, 8, 8, 0, 0, ctypes.byref(n_slabs),
                                        ctypes.byref(gps))
    assert gps.value == 16 and n_slabs.value == 4 and nbytes == 4 * 27 * 2 * 128 * 128
    assert lib.wcn_weight_image_bytes(27, 1, 64, 20, 0, 0, None, None) == 0  # cout % 16 != 0


def test_missing_library_fails_loudly(tmp_path):
    code = ("import importlib.util, sys\n"
            f"sys.path.insert(0, {ROOT!r})\n"
            "import warpconvnet_b200._lib as m\n")
    # simulate a missing library by pointing LIB lookup at an empty copy of the module
    mod = tmp_path / "pkg"
    (mod / "csrc").mkdir(parents=True)
    src = open(os.path.join(ROOT, "warpconvnet_b200", "_lib.py")).read()
    (mod / "_lib.py").write_text(src)
    (mod / "__init__.py").write_text("")
    r = subprocess.run([os.sys.executable, "-c",
                        f"import sys; sys.path.insert(0, {str(tmp_path)!r}); import pkg._lib"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "native library not found" in r.stderr


def test_ctypes_signatures_match_header_argument_counts():
    """Every ctypes signature lists exactly as many arguments as the prototype in
    include/wcn_b200.h (a short argtypes list makes ctypes pass the surplus pointers as 32-bit
    ints: a truncated stream / device pointer, i.e. a crash that depends on address bits)."""
    import os
    import re
    from warpconvnet_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "wcn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", hdr, flags=re.S)
        assert m is not None, f"{name} is bound but not declared in the header"
        body = m.group(1).strip()
        n = 0 if body in ("", "void") else body.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, ctypes lists {len(args)}"
