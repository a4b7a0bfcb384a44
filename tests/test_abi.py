"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/tmc_b200.h declares
(no compute calls: this runs without a GPU)."""

import ctypes
import os
import re

import pytest

from torch_motion_correction_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def header_symbols():
    with open(os.path.join(ROOT, "include", "tmc_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def library():
    path = _build.build()  # nvcc cross-compiles without a GPU; no-op when the .so is current
    return ctypes.CDLL(path)


def test_header_and_binding_agree(header_symbols):
    assert header_symbols == _lib.exported_symbols()


def test_library_exports_every_declared_symbol(header_symbols, library):
    missing = [name for name in header_symbols if not hasattr(library, name)]
    assert not missing, missing


def test_library_loads_through_the_binding_and_answers_queries():
    lib = _lib.load()
    assert lib.tmc_version() >= 100
    assert _lib.query("tmc_fft_supported_length", 1024) == 1
    assert _lib.query("tmc_fft_supported_length", 959) == 1  # Bluestein
    assert _lib.query("tmc_fft_supported_length", 5000) == 1  # 2 x 2500: decimated axis, full spectrum fits on chip
    assert _lib.query("tmc_fft_supported_length", 5760) == 1 and _lib.query("tmc_fft_supported_length", 11520) == 1
    assert _lib.query("tmc_fft_supported_length", 16384) == 1  # 4 x 4096
    assert _lib.query("tmc_fft_supported_length", 13000) == 2  # 4 x 3250: band-limited transforms only
    assert _lib.query("tmc_fft_supported_length", 32768) == 2
    assert _lib.query("tmc_fft_supported_length", 8198) == 0  # 2 x 4099 (prime)
    # partial-maximum slots per correlation surface: an upper bound over the row kernels that may serve the width
    assert _lib.query("tmc_xc_peak_partials", 1024, 1024) == 32    # polyphase kernel: 32 rows per CTA
    assert _lib.query("tmc_xc_peak_partials", 4096, 4096) == 512   # half-length kernel: 2 sequences x 4 rows per CTA
    assert _lib.query("tmc_xc_peak_partials", 8192, 8192) == 2048  # half-length kernel: 1 sequence x 4 rows per CTA
    assert _lib.query("tmc_fft_plan_elems", 5760) == 2 * 8192 + 2880  # plan of the 2880-point sub-transforms
    assert _lib.query("tmc_fft_plan_elems", 1024) == 1024
    assert _lib.query("tmc_fft_plan_elems", 96) == 2 * 256 + 96
    assert _lib.query("tmc_spline_workspace_floats", 2, 3, 5, 5) == 2 * 5 * 7 * 7
    assert _lib.query("tmc_spline_workspace_floats", 2, 40, 1, 1) == 2 * 42 * 4 * 4
    assert _lib.query("tmc_warp_workspace_floats", 40, 4096, 50) == 40 * 2 * (50 + 3) * 4096  # reflection-padded rows
    assert _lib.query("tmc_xc_peak_partials", 1024, 1024) == 32


def test_bad_arguments_are_reported_not_crashed():
    # argument validation happens before any CUDA call, so this is safe without a GPU
    with pytest.raises(ValueError, match="null pointer"):
        _lib.call("tmc_stack_stats", None, 1, 8, 8, 2, 6, 2, 6, None, None, None)
    with pytest.raises(ValueError, match="kind must be"):
        _lib.call("tmc_spline_eval", 1, 2, 2, 2, 2, 7, None, 0, None, 1, None)
    assert "kind must be" in _lib.load().tmc_last_error().decode()


def test_product_never_imports_the_oracle():
    """The shipped package must not route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "torch_motion_correction_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith(".py"):
                with open(os.path.join(dirpath, name)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), name


def test_header_binding_and_definitions_agree_on_arity():
    """Every entry point has the same number of parameters in include/tmc_b200.h, in the ctypes binding and in its
    definition under csrc/ (a drift would corrupt the call silently)."""
    import glob

    with open(os.path.join(ROOT, "include", "tmc_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    decls = re.findall(r"\b(tmc_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(decls) == len(_lib._SIGNATURES)

    def arity(args):
        args = args.strip()
        return 0 if args in ("void", "") else args.count(",") + 1

    for name, args in decls:
        assert arity(args) == len(_lib._SIGNATURES[name][1]), name
    src = ""
    for path in glob.glob(os.path.join(ROOT, "torch_motion_correction_b200", "csrc", "*.cu")):
        with open(path) as f:
            src += f.read()
    defs = re.findall(r"TMC_API\s+[a-z\s\*]+?\b(tmc_[a-z0-9_]+)\s*\(([^{;]*?)\)\s*\{", src, flags=re.S)
    assert {n for n, _ in defs} == set(_lib._SIGNATURES)
    for name, args in defs:
        assert arity(args) == len(_lib._SIGNATURES[name][1]), name
