"""The jax.ffi shim (probdiffeq_b200/csrc/ffi/pdeq_xla_ffi.cc) cannot be built against XLA here (no JAX). It is
compile-checked against a local stand-in of the xla::ffi names it uses (tests/stubs/xla/ffi/api/ffi.h): that pins its
use of the C ABI (include/probdiffeq_b200.h) and that every handler's signature equals its binding. NOT a test of XLA."""

import pathlib
import shutil
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
SHIM = ROOT / "probdiffeq_b200" / "csrc" / "ffi" / "pdeq_xla_ffi.cc"
CUDA_INC = pathlib.Path("/usr/local/cuda/include")


def _gxx(args, stdin=None):
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-DPDEQ_WITH_XLA_FFI", "-I", str(ROOT / "tests" / "stubs"),
           "-I", str(ROOT / "include"), "-I", str(CUDA_INC), *args]  # fmt: skip
    return subprocess.run(cmd, input=stdin, capture_output=True, text=True)


needs_toolchain = pytest.mark.skipif(shutil.which("g++") is None or not (CUDA_INC / "cuda_runtime_api.h").exists(),
                                     reason="g++ / CUDA headers not available")


@needs_toolchain
def test_shim_compiles_against_the_stub_header():
    res = _gxx([str(SHIM)])
    assert res.returncode == 0, res.stderr


@needs_toolchain
def test_stub_rejects_a_handler_that_does_not_match_its_binding():
    src = """
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;
static ffi::Error Handler(double t0, ffi::Buffer<ffi::F64> x) { return ffi::Error::Success(); }
XLA_FFI_DEFINE_HANDLER_SYMBOL(Sym, Handler, ffi::Ffi::Bind().Arg<ffi::Buffer<ffi::F64>>().Attr<double>("t0"));
"""
    res = _gxx(["-x", "c++", "-"], stdin=src)
    assert res.returncode != 0 and "does not match its binding" in res.stderr
