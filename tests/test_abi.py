"""The C-ABI boundary without a GPU: libb200lp.so loads, exports exactly what include/b200lp.h
declares, the ctypes mirror has the C layout, the host-only entry points work, and every
device entry point fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from linear_programming_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200lp.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(b200lp_[a-z0-9_]+)\s*\(", text))


def test_every_declared_symbol_is_bound_and_exported():
    declared = declared_functions()
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    lib = _ffi.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True)
    exported = set(re.findall(r"\bT (b200lp_\w+)", out.stdout))
    assert exported == declared, exported ^ declared


def test_struct_layout_matches_the_header(tmp_path):
    src = tmp_path / "layout.c"
    fields_o = [n for n, _ in _ffi.Opts._fields_]
    fields_r = [n for n, _ in _ffi.Result._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
             'printf("%zu %zu\\n", sizeof(b200lp_opts), sizeof(b200lp_result));']
    for f in fields_o:
        lines.append(f'printf("%zu\\n", offsetof(b200lp_opts, {f}));')
    for f in fields_r:
        lines.append(f'printf("%zu\\n", offsetof(b200lp_result, {f}));')
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert [int(out[0]), int(out[1])] == [ctypes.sizeof(_ffi.Opts), ctypes.sizeof(_ffi.Result)]
    offs = [int(x) for x in out[2:]]
    want = [getattr(_ffi.Opts, f).offset for f in fields_o] + \
           [getattr(_ffi.Result, f).offset for f in fields_r]
    assert offs == want


def test_host_only_entry_points():
    lib = _ffi.lib()
    assert lib.b200lp_version() >= 101
    assert _ffi.strerror(_ffi.UNBOUNDED) == "Problem is unbounded"          # src/conditions.lisp:47-53
    assert _ffi.strerror(_ffi.INFEASIBLE) == "Problem has no feasible region"  # :55-60
    assert "peer" in _ffi.strerror(_ffi.ERR_PEER_TIMEOUT)
    eps = float.fromhex("0x1.0000000000001p-53")                             # CL double-float-epsilon
    assert _ffi.thresholds(1024) == (128 * eps, 512 * eps, 1024 * eps)
    assert _ffi.thresholds(8) == (eps, 4 * eps, 8 * eps)
    assert [_ffi.partition(10, 3, r) for r in range(3)] == [(0, 4), (4, 8), (8, 10)]
    assert [_ffi.partition(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]


@pytest.mark.skipif(_ffi.device_count() > 0, reason="a GPU is present")
def test_device_entry_points_fail_loudly_without_a_gpu():
    tab = np.array([[2, 1, 0, 1, 0, 8], [0, 1, 1, 0, 1, 7], [-1, -4, -3, 0, 0, 0]], dtype=np.float64)
    basis = np.array([3, 4], dtype=np.int32)
    before = tab.copy()
    with pytest.raises(_ffi.B200DeviceError) as e:
        _ffi.solve(tab, basis, True)
    assert e.value.code == _ffi.ERR_NO_DEVICE and np.array_equal(tab, before)
    with pytest.raises(_ffi.B200DeviceError):
        _ffi.DeviceTableau(3, 6)
    with pytest.raises(_ffi.B200DeviceError):
        _ffi.solve_two_phase(tab.copy(), basis.copy(), tab.copy(), basis.copy(), True)


def test_missing_library_is_an_error_not_a_fallback(monkeypatch):
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", os.path.join(ROOT, "does-not-exist.so"))
    with pytest.raises(_ffi.B200LibraryError):
        _ffi.lib()


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "linear-programming_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".lisp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
    out = subprocess.run(["ldd", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_argument_validation_happens_before_any_device_work():
    with pytest.raises(ValueError):
        _ffi.solve(np.zeros((3, 6), dtype=np.float32), np.zeros(2, np.int32))
    with pytest.raises(ValueError):
        _ffi.solve(np.zeros((3, 6)), np.zeros(5, np.int32))
    with pytest.raises(ValueError):
        _ffi.make_opts(devices=list(range(9)))


def _build_c_consumer(tmp_path):
    exe = tmp_path / "c_abi_consumer"
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_consumer.c"), "-o", str(exe),
                           "-L", libdir, "-l:libb200lp.so", f"-Wl,-rpath,{libdir}"])
    return exe


def test_plain_c_consumer_links_and_is_refused_without_a_gpu(tmp_path):
    """include/b200lp.h is valid C99 and the library links from C with no torch / Python around."""
    exe = _build_c_consumer(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    if _ffi.device_count() == 0:
        assert "no device" in out.stdout


@pytest.mark.gpu
def test_plain_c_consumer_solves_on_the_gpu(tmp_path):
    exe = _build_c_consumer(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "solved: objective 28.5 in 2 pivots" in out.stdout, out.stdout + out.stderr
