"""CPU checks of the boundary: the C-ABI library loads, exports every symbol include/cggp_b200.h declares, and the
product path refuses to run without a GPU (no CPU fallback, no oracle import)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cggp_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cggp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    from cggp_b200 import build

    return ctypes.CDLL(build.build())


def test_header_declares_expected_surface():
    fns = header_functions()
    for must in ["cggp_cg_solve", "cggp_cg_fused_step", "cggp_kuf_kfu_matvec", "cggp_kernel_matrix",
                 "cggp_nearest_center", "cggp_allreduce_sum", "cggp_block_cholesky", "cggp_symm_matmul"]:
        assert must in fns


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_signatures_cover_header():
    from cggp_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_functions()


def test_struct_layout_matches_header():
    from cggp_b200 import _lib

    assert ctypes.sizeof(_lib.Operator) == 168
    assert ctypes.sizeof(_lib.Precond) == 48
    assert _lib.Operator.struct_size.offset == 0 and _lib.Operator.n.offset == 16
    assert _lib.Operator.dev_PZ.offset == 88 and _lib.Operator.variant.offset == 44
    assert _lib.Operator().struct_size == 168  # the constructor stamps it; cggp_cg_solve rejects any other value


def test_header_struct_fields_match_ctypes_and_integration_doc():
    """Field names and order of struct cggp_operator: header == cggp_b200._lib.Operator == the ctypes stub published
    in INTEGRATION.md (a maintainer copying a stale stub would make the library read past the struct)."""
    from cggp_b200 import _lib

    src = open(HEADER).read()
    body = src[src.index("typedef struct cggp_operator {"):src.index("} cggp_operator;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    header_fields = re.findall(r"\b(\w+)\s*;", body)
    lib_fields = [f[0] for f in _lib.Operator._fields_]
    assert header_fields == lib_fields
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = doc[doc.index("class Operator(C.Structure):"):]
    stub = stub[:stub.index("op = Operator(")]
    doc_fields = re.findall(r'\("(\w+)",\s*C\.(\w+)\)', stub)
    assert [f for f, _ in doc_fields] == lib_fields
    alias = {"c_uint": "c_uint32", "c_int": "c_int32", "c_long": "c_int64"}  # ctypes names on LP64
    assert [t for _, t in doc_fields] == [alias.get(f[1].__name__, f[1].__name__) for f in _lib.Operator._fields_]


def test_prepared_ld_and_version(lib):
    lib.cggp_prepared_ld.restype = ctypes.c_int64
    assert [lib.cggp_prepared_ld(d) for d in (1, 2, 3, 4, 11, 90)] == [4, 4, 4, 8, 12, 92]
    lib.cggp_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.cggp_version()


def test_tensor_core_buffer_sizes(lib):
    """Host-side size queries of the float32 tensor-core arrays (no GPU needed): rows padded to 128, features to 32,
    one 1 KB trailer of per-row scalars per 128-row tile; the 3xFP16 mode adds one scale per row."""
    lib.cggp_tf32_rows.restype = ctypes.c_int64
    lib.cggp_tf32_rows.argtypes = [ctypes.c_int64]
    lib.cggp_tf32_sizes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                    ctypes.POINTER(ctypes.c_int64)]
    assert [lib.cggp_tf32_kp(d) for d in (1, 32, 33, 90)] == [32, 32, 64, 96]
    assert [lib.cggp_tf32_rows(n) for n in (0, 1, 128, 129)] == [0, 128, 128, 256]
    ns, nr = ctypes.c_int64(), ctypes.c_int64()
    for nsplit, parts, esz, rows_extra in ((1, 1, 4, False), (3, 2, 4, False), (16, 2, 2, True)):
        assert lib.cggp_tf32_sizes(nsplit, 1000, 90, ctypes.byref(ns), ctypes.byref(nr)) == 0
        tiles, kp = 8, 96
        assert ns.value * 4 == tiles * ((kp // 32) * parts * 4096 * esz + 1024)
        assert nr.value == (1024 if rows_extra else 1)
    assert lib.cggp_tf32_sizes(2, 1000, 90, ctypes.byref(ns), ctypes.byref(nr)) != 0
    # the FP16 stream is half the bytes of the 3xTF32 one (the SASS must hold the kind::f16 tensor-core path)
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "cggp_b200", "libcggp_b200.so")],
                         capture_output=True, text=True)
    if out.returncode == 0:
        assert "UTCHMMA" in out.stdout or "UTCMMA" in out.stdout or "tcgen05" in out.stdout.lower()


def test_tensor_core_ring_plans_keep_their_invariants(lib):
    """The tcgen05 kernel releases a tile's stages together and deals tiles round-robin to its MMA-issuing threads:
    the ring must hold whole tiles, a whole multiple of `niss` of them (every stage always consumed by the same
    issuer), and fit the 227 KB of shared memory - for every feature count and arithmetic mode."""
    lib.cggp_tf32_ring_plan.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)] * 3 + \
        [ctypes.POINTER(ctypes.c_int64)]
    gc, st, ni, sm = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    seen = set()
    for nsplit in (1, 3, 16):
        for nb in (1, 2):
            for D in range(1, 129):
                assert lib.cggp_tf32_ring_plan(nsplit, D, nb, ctypes.byref(gc), ctypes.byref(st), ctypes.byref(ni),
                                               ctypes.byref(sm)) == 0
                nchunk = lib.cggp_tf32_kp(D) // 32
                parts_cols = (1 if nsplit == 1 else 2) * lib.cggp_tf32_kp(D) // (2 if nsplit == 16 else 1)
                if st.value == 0:
                    # only when the row tile does not fit tensor memory next to the accumulators
                    assert parts_cols > (128 if nsplit == 16 else 256), (nsplit, nb, D)
                    continue
                assert nchunk % gc.value == 0
                ng = nchunk // gc.value
                assert 1 <= ni.value <= 2
                assert st.value % (ng * ni.value) == 0 and st.value >= ng
                assert sm.value <= 227 * 1024
                seen.add((nsplit, gc.value, ni.value))
    assert (16, 3, 2) in seen and (3, 1, 2) in seen and (1, 3, 1) in seen  # the c5 shape (D = 90) in the three modes
    assert lib.cggp_tf32_ring_plan(2, 10, 1, None, None, None, None) != 0


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch

    import cggp_b200

    if not torch.cuda.is_available():
        with pytest.raises(cggp_b200._lib.CggpError):
            cggp_b200._lib.context()
        with pytest.raises(cggp_b200._lib.CggpError):
            cggp_b200.conjugate_gradient(torch.eye(3, dtype=torch.float64), torch.ones(1, 3, dtype=torch.float64),
                                         None, 1e-6)
        with pytest.raises(cggp_b200._lib.CggpError):
            cggp_b200.CoverTree(None, (torch.zeros(4, 2, dtype=torch.float64), torch.zeros(4, 1, dtype=torch.float64)),
                                spatial_resolution=0.5)
    # nothing under cggp_b200/ may import the oracle
    pkg = os.path.join(ROOT, "cggp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_sass_has_dmma_and_no_cpu_path():
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "cggp_b200", "libcggp_b200.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "DMMA" in out.stdout
    assert "sm_100a" in out.stdout


def test_header_is_plain_c():
    """The boundary is a C ABI: include/cggp_b200.h must compile as C99 (no C++-only constructs, no torch types)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    header = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "cggp_b200.h")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-pedantic", "-Werror", header],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
