"""The C-ABI shared library loads and exports every symbol include/fdal.h declares
(no compute calls: there is no GPU on the CPU test box)."""
import ctypes
import os
import re

import pytest

from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import build, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fdal.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fdal_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    for must in ("fdal_create", "fdal_set_csr", "fdal_amg_set_level", "fdal_finalize", "fdal_apply_aug",
                 "fdal_apply_prec", "fdal_apply_system", "fdal_solve", "fdal_solve_dev", "fdal_destroy"):
        assert must in syms


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    dll = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(dll, s), f"{s} declared in fdal.h but not exported by libfdal.so"


def test_binding_table_matches_header():
    api = lib.load()
    names = {s[len("fdal_"):] for s in declared_symbols()}
    bound = set(b.SIGNATURES) | set(b.DEVICE_SIGNATURES)
    assert names == bound, (names - bound, bound - names)
    assert api.version().startswith(b"fdal")


def test_struct_sizes_match_the_c_side():
    # fdal_control: 2*int32 + 2*double ; fdal_config: 13 int32/double fields + 3 controls
    assert ctypes.sizeof(b.Control) == 24
    assert ctypes.sizeof(b.Config) == 8 + 24 + 10 * 4 + 3 * 24
    assert ctypes.sizeof(b.SolveInfo) == 8 * 4 + 3 * 8 + 8 + 8 * b.MAX_HISTORY
    assert ctypes.sizeof(b.CsrView) == 48


def test_no_cpu_fallback_without_a_gpu():
    """On a box without CUDA devices fdal_create must fail loudly (FDAL_ERR_CUDA)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fictitious_domain_al_preconditioners_b200 import ALConfig, ALContext, FdalError

    with pytest.raises(FdalError) as e:
        ALContext(ALConfig())
    assert e.value.status == b.ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fictitious_domain_al_preconditioners_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "fdal_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """include/fdal.h compiles as C99 and as C++17, and a C program (examples/c_abi_smoke.c) links
    against libfdal.so with nothing but the C ABI; on a box without a GPU it reports
    FDAL_ERR_CUDA and exits 0, on a GPU box it runs a tiny solve."""
    import subprocess

    inc = os.path.join(ROOT, "include")
    hdr = os.path.join(inc, "fdal.h")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    lib = build.build()
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(lib)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", inc, os.path.join(ROOT, "examples", "c_abi_smoke.c"), "-L", libdir,
                    "-lfdal", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("fdal ")


def test_dealii_example_compiles():
    """examples/dealii_immersed_laplace_solve.cc — the function a maintainer adds to immersed_laplace.cc —
    goes through a C++17 compiler against the stand-in deal.II / Trilinos-ML types (export_amg included)."""
    import subprocess

    stub = os.path.join(ROOT, "oracle", "ref_harness")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                    "-I", os.path.join(stub, "dealii_stub"), "-I", os.path.join(stub, "trilinos_stub"),
                    "-I", os.path.join(ROOT, "include"), "-DFDAL_STUB_TRILINOS",
                    os.path.join(ROOT, "examples", "dealii_immersed_laplace_solve.cc")], check=True)


def test_stand_alone_setup_entry_points_validate_before_touching_a_device():
    """fdal_assemble_al_term / fdal_csr_to_bsr (no context): bad arguments are refused with FDAL_ERR_INVALID /
    FDAL_ERR_SHAPE, and without a CUDA device the calls fail loudly with FDAL_ERR_CUDA (no CPU fallback)."""
    import numpy as np
    import torch

    api = lib.load()
    p64, p32, pd = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
    rp = np.array([0, 1, 2, 3, 4, 5, 6], dtype=np.int64)
    ci = np.arange(6, dtype=np.int32)
    v = np.ones(6)
    brp = np.zeros(4, dtype=np.int32)
    a = (rp.ctypes.data_as(p64), ci.ctypes.data_as(p32), v.ctypes.data_as(pd))
    # block size out of range / rows not a multiple of it / null arrays
    assert api.csr_to_bsr(0, 6, 6, *a, 4, 1.35, brp.ctypes.data_as(p32), 0, None, None) == -b.ERR_INVALID
    assert api.csr_to_bsr(0, 5, 5, *a, 2, 1.35, brp.ctypes.data_as(p32), 0, None, None) == -b.ERR_INVALID
    assert api.csr_to_bsr(0, 6, 6, None, a[1], a[2], 2, 1.35, brp.ctypes.data_as(p32), 0, None, None) == -b.ERR_INVALID
    dofs = np.array([[0, 7]], dtype=np.int32)  # dof 7 >= n_rows = 6
    phi = np.array([[0.5, 0.5]])
    w = np.ones(1)
    miss = ctypes.c_int64(0)
    al = lambda n_pts, dpc, d: api.assemble_al_term(0, 6, a[0], a[1], a[2], n_pts, dpc, d.ctypes.data_as(p32),  # noqa: E731
                                                      phi.ctypes.data_as(pd), w.ctypes.data_as(pd), ctypes.byref(miss))
    assert al(1, 2, dofs) == b.ERR_SHAPE
    assert al(1, 0, dofs) == b.ERR_INVALID
    assert al(-1, 2, dofs) == b.ERR_INVALID
    if not torch.cuda.is_available():
        ok = np.array([[0, 1]], dtype=np.int32)
        assert al(1, 2, ok) == b.ERR_CUDA
        assert api.csr_to_bsr(0, 6, 6, *a, 2, 1.35, brp.ctypes.data_as(p32), 0, None, None) == -b.ERR_CUDA
