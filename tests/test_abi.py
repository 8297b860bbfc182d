"""CPU test: the product library loads and exports every symbol include/zkm_b200.h declares; with
no GPU it must fail loudly (no CPU fallback)."""
import ctypes as C
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _built():
    from zkm_b200 import build
    return build.build(verbose=False)


def test_header_symbols_exported():
    lib = C.CDLL(str(_built()))
    header = (ROOT / "include/zkm_b200.h").read_text()
    names = set(re.findall(r"\b(zkm_b200_\w+)\s*\(", header))
    assert len(names) >= 16
    from zkm_b200.lib import EXPORTS
    assert names == set(EXPORTS), names ^ set(EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_standard_fast_config_matches_reference():
    # reference prover/src/config.rs:17-29
    from zkm_b200.lib import load, StarkConfig
    lib = load()
    c = StarkConfig()
    lib.zkm_b200_standard_fast_config(C.byref(c))
    assert (c.rate_bits, c.cap_height, c.pow_bits, c.num_queries, c.num_challenges, c.arity_bits, c.final_poly_bits) == \
        (2, 4, 16, 37, 2, 4, 5)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zkm_b200.lib import load
    lib = load()
    err = C.c_void_p()
    rc = lib.zkm_b200_init(0, C.byref(err))
    assert rc == -1 and err.value
    msg = C.cast(err, C.c_char_p).value.decode()
    assert "no CPU fallback" in msg or "CUDA" in msg
    lib.zkm_b200_free_string(err)
