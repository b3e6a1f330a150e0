"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, mirrors the struct layouts, and refuses to compute without an sm_100 GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from mocha_sigasia2023_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mocha_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mocha_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mocha_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert lib.mocha_version() >= 100


def test_struct_layouts_match():
    lib = _lib.load()
    sizes = (C.c_size_t * 10)()
    assert lib.mocha_struct_sizes(sizes, 10) == 0
    assert [C.sizeof(s) for s in _lib._STRUCTS] == list(sizes)
    assert lib.mocha_struct_sizes(sizes, 3) != 0
    assert b"mocha_struct_sizes" in lib.mocha_last_error()


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    rc = lib.mocha_linear(None, None, None, None, None, 4, 4, 4, 0, 0, None, 0, None)
    assert rc == -1 and b"mocha_linear" in lib.mocha_last_error()
    rc = lib.mocha_match_exact(None, 1, None, 1, 1, 1, 0, None, None, None, 0, None)
    assert rc == -1
    assert lib.mocha_match_exact_workspace_bytes(0, 10, 1) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.mocha_check_device() != 0
    from mocha_sigasia2023_b200.model import Generator
    from mocha_sigasia2023_b200.weights import DEFAULT_MODEL_CFG
    from mocha_sigasia2023_b200.transformer import mean_variance_norm
    g = Generator(DEFAULT_MODEL_CFG)
    with pytest.raises(_lib.MochaError):
        g.mot_embedding(torch.zeros(1, 60, 24, 15))
    with pytest.raises(_lib.MochaError):
        mean_variance_norm(torch.zeros(1, 256, 90))


def test_state_dict_keys_match_reference_surface():
    from mocha_sigasia2023_b200.model import Generator
    from mocha_sigasia2023_b200.model_CVAE import CVAE
    from mocha_sigasia2023_b200 import weights
    g = Generator(weights.DEFAULT_MODEL_CFG)
    sd = weights.generator_state_dict(1777)
    assert list(g.state_dict().keys()).sort() == list(sd.keys()).sort()
    assert set(g.state_dict().keys()) == set(sd.keys()) and len(sd) == 71
    assert sum(v.numel() for k, v in g.named_parameters()) == 6116559
    g.load_state_dict(sd, strict=True)
    c = CVAE(output_seq=90)
    csd = weights.cvae_state_dict(1778)
    assert set(c.state_dict().keys()) == set(csd.keys()) and len(csd) == 91
    assert sum(v.numel() for k, v in c.named_parameters()) == 3691008
    c.load_state_dict(csd, strict=True)
