"""CPU-only checks of the C-ABI boundary: the library builds, loads, exports every symbol include/bkf.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""

import ctypes
import os
import re

import pytest

import pymc_experimental_b200 as pkg
from pymc_experimental_b200 import build, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return engine.load_library()


def test_header_symbols_are_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "bkf.h")).read()
    declared = set(re.findall(r"\b(bkf_[a-z_]+)\s*\(", header))
    assert declared == set(engine.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/bkf.h but not exported by libbkf.so"


def test_version_matches_header(lib):
    header = open(os.path.join(ROOT, "include", "bkf.h")).read()
    assert lib.bkf_version() == int(re.search(r"#define BKF_VERSION (\d+)", header).group(1))


def test_problem_struct_layout_matches_header():
    header = open(os.path.join(ROOT, "include", "bkf.h")).read()
    body = re.search(r"typedef struct bkf_problem \{(.*?)\} bkf_problem;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip() for n in re.sub(r"^(unsigned long long|unsigned|int|double)\s+", "", decl).split(",")]
    assert names == [f[0] for f in engine.BkfProblem._fields_]
    assert ctypes.sizeof(engine.BkfProblem) == 11 * 4 + 4 + 2 * 8 + 2 * 8  # 11 ints/unsigned + pad + 2 doubles + 2 u64


def test_no_cpu_fallback_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.BkfError):
        engine.KalmanEngine(0)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg_dir = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
