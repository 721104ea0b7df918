"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports
every symbol include/cudaqr_b200.h declares; no compute call is made here."""
import os
import re

import pytest

from conftest import ROOT, load_pkg

HEADER = os.path.join(ROOT, "include", "cudaqr_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)))


def test_library_loads_and_exports_every_declared_symbol():
    pkg = load_pkg()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(pkg.lib, name), f"{name} declared in cudaqr_b200.h but not exported"
    assert sorted(pkg.EXPORTS) == names


def test_version_string_names_the_arch():
    pkg = load_pkg()
    assert "sm_100a" in pkg.version()


def test_get_panel_dims_matches_reference_formula():
    """qr.cu:49-55 with PR=64, PC=4 (pure integer host code: safe without a GPU)."""
    import oracle
    pkg = load_pkg()
    for m, n in [(64, 64), (124, 64), (4084, 4084), (16384, 16384), (131044, 64), (6, 4), (65, 1)]:
        assert pkg.getPanelDims(m, n) == oracle.panel_dims(m, n, 64, 4)
        assert pkg.tau_size(m, n) >= n


def test_compute_call_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = load_pkg()
    with pytest.raises(pkg.CudaQRError):
        pkg.Context(0)


def test_no_product_import_of_oracle():
    """The product package must never route through oracle/ (CPU fallback voids parity)."""
    pkg_dir = os.path.join(ROOT, "cuda-qr_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src and "oracle/_ref" not in src, f


def test_cli_builds_and_prints_the_reference_usage_line():
    """qr_device is built next to the library (csrc/Makefile) and, like qr.cu:715-719, exits 1 with the usage line
    before touching the GPU when the sizes are missing."""
    import subprocess
    exe = os.path.join(ROOT, "cuda-qr_b200", "qr_device")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 1 and out.stdout.strip() == "Usage: ./qr_device m n"
