"""The drop-in boundary: both C-ABI libraries load without a GPU and export every symbol their
header declares; the device path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "navier-stokes-dealii_b200")


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ns[tg]_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("header,lib", [("nst.h", "libnst.so"), ("nsg.h", "libnsg.so")])
def test_every_declared_symbol_is_exported(header, lib):
    names = declared_functions(header)
    assert len(names) > 25
    L = ctypes.CDLL(os.path.join(PKG, lib))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_bindings_cover_the_headers(pkg):
    from importlib import import_module
    lib = import_module("navier-stokes-dealii_b200._lib")
    assert set(declared_functions("nst.h")) == set(lib._NST_SIGS)
    assert set(declared_functions("nsg.h")) == set(lib._NSG_SIGS)
    lib.nst(), lib.nsg()      # binding every signature must succeed


def test_no_cpu_fallback_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from importlib import import_module
    lib = import_module("navier-stokes-dealii_b200._lib")
    h = ctypes.c_void_p()
    rc = lib.nsg().nsg_create(0, ctypes.byref(h))
    assert rc == -1 and b"no CUDA device" in lib.nsg().nsg_last_error()
    m = pkg.Mesh.read_msh(os.path.join(ROOT, "tests/golden/square_h0.1.msh"))
    part = pkg.Part(pkg.Dofs(m), 0)
    with pytest.raises(lib.NsgError, match="no CUDA device"):
        pkg.DeviceProblem(part, 0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f in ("nst.cpp",) and "not the oracle" in txt, (dirpath, f)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port, timed on the host cores) must print ONE
    JSON line with the keys the driver reads; runs without a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--mesh", "cmy", "--cpu-level", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "jacobian_assembly_mdofs" and j["unit"] == "MDoF/s"
    assert j["value"] > 0 and j["higher_is_better"] is True and j["dtype"] == "f64"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "MDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]
