"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what include/*.h declares,
and fails loudly (no fallback) when no GPU is present.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "onesweep_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"OSB200_API\s+[\w\s\*]+?\b(osb200_\w+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("osb200_create", "osb200_destroy", "osb200_sort_keys_u32", "osb200_sort_pairs_u32",
                 "osb200_sort_keys_u64", "osb200_global_histogram", "osb200_digit_binning_pass",
                 "osb200_validate", "osb200_sharded_sort_keys_u32"):
        assert must in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    import gpusorting_b200 as g

    cdll = ctypes.CDLL(g.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(cdll, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_python_binding_covers_header():
    from gpusorting_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_status_strings():
    import gpusorting_b200 as g

    assert g.lib.osb200_version() >= 1000
    assert g.status_string(0) == "ok"
    assert "max_n" in g.status_string(-2)
    assert g.status_string(-1001) != "unknown status"  # cuda error text


def test_workspace_bytes_accounts_for_alt_buffers():
    import gpusorting_b200 as g

    n = 1 << 20
    assert g.lib.osb200_workspace_bytes(n, 4, 0) >= 4 * n
    assert g.lib.osb200_workspace_bytes(n, 4, 4) >= 8 * n
    assert g.lib.osb200_workspace_bytes(n, 8, 0) >= 8 * n
    assert g.lib.osb200_workspace_bytes(n, 3, 0) == 0  # invalid widths


def test_create_rejects_bad_arguments_without_touching_a_device():
    import gpusorting_b200 as g

    h = ctypes.c_void_p()
    assert g.lib.osb200_create(None, 1024, 4, 0) == -1
    assert g.lib.osb200_create(ctypes.byref(h), 1024, 2, 0) == -1
    assert g.lib.osb200_create(ctypes.byref(h), 1024, 8, 4) == -3
    assert g.lib.osb200_create(ctypes.byref(h), 0, 4, 0) == -1
    assert h.value is None


def test_no_cpu_fallback_without_gpu():
    import torch

    import gpusorting_b200 as g

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert g.lib.osb200_create(ctypes.byref(h), 1024, 4, 0) == -4  # OSB200_ERR_NO_DEVICE, loudly
    with pytest.raises(RuntimeError):
        g.OneSweepSorter(1024)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gpusorting_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gpusorting_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oraclelib" not in text and "oracle/" not in text, f


def test_missing_library_fails_loudly():
    """No fallback of any kind: if the CUDA library is absent, importing the package raises."""
    import subprocess
    import sys

    code = "import os; os.environ['OSB200_LIB']='/nonexistent/libonesweep_b200.so'; import gpusorting_b200"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0
    assert "ImportError" in r.stderr and "no" in r.stderr.lower() and "fallback" in r.stderr.lower()


def test_cpp_mirror_header_compiles_against_the_library():
    """include/OneSweepB200.hpp (OneSweep::Sort, the north-star call shape) and the reference-style driver build with a
    plain host compiler against the C-ABI: no CUDA headers are needed on the caller's side."""
    import subprocess
    import tempfile

    src = r'''
#include "OneSweepB200.hpp"
int main() {
    // never executed here (no GPU): instantiate the overloads and take their addresses
    void (*a)(uint32_t*, uint64_t, void*) = &OneSweep::Sort;
    void (*b)(uint64_t*, uint64_t, void*) = &OneSweep::Sort;
    void (*c)(uint32_t*, uint32_t*, uint64_t, void*) = &OneSweep::Sort;
    return (a && b && c && osb200_version() >= 1000) ? 0 : 1;
}
'''
    libdir = os.path.join(ROOT, "gpusorting_b200", "lib")
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "t.cpp")
        open(cpp, "w").write(src)
        exe = os.path.join(td, "t")
        r = subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), cpp, "-o", exe, "-L", libdir,
                            "-lonesweep_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_host_side_units_of_the_device_headers():
    """tests/host_unit.cu: descriptor packing and the typed-key codec, compiled with nvcc and run on the CPU."""
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "host_unit")
        r = subprocess.run(["nvcc", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host_unit.cu")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "host_unit: ok" in r.stdout, r.stdout + r.stderr
