"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/xvc_b200.h declares, the struct layouts match the Python mirrors, and the product
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from xvc_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = lib.load()
    header = open(os.path.join(ROOT, "include", "xvc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(xvcb200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) > 50
    for name in declared:
        assert hasattr(L, name), "libxvc_b200.so does not export %s" % name
    assert sorted(set(lib.EXPORTS)) == declared
    assert b"sm_100a" in L.xvcb200_version()


def test_struct_layouts_match_header():
    L = lib.load()
    for which, (name, dt) in abi.ABI_STRUCTS.items():
        assert L.xvcb200_abi_sizeof(which) == dt.itemsize, name
    assert abi.cu_dtype.fields["mv"][1] == 12 and abi.cu_dtype.fields["flags"][1] == 7


def test_qp_init_host_side():
    # pure host arithmetic (Qp::Qp, quantize.cc:48-92): 4:2:0 chroma table at qp 32 / 10 bit
    q = lib.qp_init(32, 10, lam=57.9)
    assert list(q["qp_raw"]) == [32, 31, 31] and list(q["qp_bitdepth"]) == [44, 43, 43]
    assert q["distortion_weight"][0] == 1.0 and abs(q["distortion_weight"][1] - 2 ** (1 / 3.0)) < 1e-12
    assert q["lambda_sqrt"] == np.sqrt(57.9)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(lib.XvcB200Error):
        lib.Context(64, 64)
    a = np.zeros((8, 8), dtype=np.uint16)
    with pytest.raises(lib.XvcB200Error):
        lib.sad(a, a, 8, 8)


def test_header_is_plain_c_and_cpp():
    """include/xvc_b200.h is the drop-in boundary: it must compile on its own as C99 and as C++11
    (what the reference is built with), without warnings under -Wall -Wextra -pedantic."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "xvc_b200.h")
    for cmd in (["gcc", "-std=c99", "-x", "c"], ["g++", "-std=c++11", "-x", "c++"]):
        out = subprocess.run(cmd + ["-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr


def test_cpp_host_example_builds_and_links(tmp_path):
    """examples/encode_step.cc -- the batched step from plain C++ -- compiles against the header, links
    against libxvc_b200.so and, on a machine without a CUDA device, fails loudly (exit 3: no CPU fallback);
    with a device it runs the step (exit 0)."""
    import subprocess
    lib.load()
    exe = str(tmp_path / "encode_step")
    libdir = os.path.join(ROOT, "xvc_b200")
    cmd = ["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "encode_step.cc"),
           "-o", exe, "-L", libdir, "-lxvc_b200", "-Wl,-rpath," + libdir]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode in (0, 3), (run.returncode, run.stdout, run.stderr)
    if run.returncode == 3:
        assert "no CUDA device" in run.stderr
    else:
        assert "recon checksum" in run.stdout
