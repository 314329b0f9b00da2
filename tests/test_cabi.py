"""The C-ABI shared library loads on a machine without a GPU and exports every
symbol include/rbg_b200.h declares; calls fail loudly (negative code + message)
instead of falling back to a CPU path.  No compute is attempted here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build_library()
    import routing_board_generation_b200 as pkg

    return pkg._lib.load()


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "rbg_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rbg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    import routing_board_generation_b200 as pkg

    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rbg_b200.h but not exported"
    # the ctypes table covers the whole header, nothing more
    assert sorted(pkg._lib.SYMBOLS) == names


def test_version_and_error_string(lib):
    assert lib.rbg_version() == 200
    rc = lib.rbg_prw_generate(None, 4, 10, 5, None, None, None, None, None)
    assert rc == -1 and b"NULL" in lib.rbg_last_error()
    rc = lib.rbg_prw_generate(None, 4, 99, 5, None, None, None, None, None)
    assert rc == -1 and b"grid size" in lib.rbg_last_error()
    rc = lib.rbg_validate(None, 4, 10, 5, None, None)
    assert rc == -1


def test_struct_layouts_match_header():
    import routing_board_generation_b200 as pkg

    L = pkg._lib
    assert C.sizeof(L.rbg_state) == 7 * C.sizeof(C.c_void_p)
    assert C.sizeof(L.rbg_timestep) == 9 * C.sizeof(C.c_void_p)
    assert C.sizeof(L.rbg_env_params) == 40
    assert [f[0] for f in L.rbg_state._fields_] == ["grid", "step_count", "agent_id", "start", "target", "position", "key"]


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import routing_board_generation_b200 as pkg

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.ParallelRandomWalkBoard(10, 10, 5).generate_board(np.zeros(2, np.uint32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.Connector().reset(np.zeros((4, 2), np.uint32))


def test_product_package_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "routing-board-generation_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle|rbg_oracle|librbg_oracle|orc_[a-z]", txt, flags=re.M), f"{fn} uses the oracle"
