"""
CPU-only tests of the host side: the C-ABI library loads and exports every symbol the
header declares, the ctypes prototypes cover them, and the Python entry points mirror the
reference's argument checking (errors are raised before any device work).
"""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "africanus_b200.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(afr_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def built_lib():
    from codex_africanus_b200 import build

    return build.build()


def test_abi_exports_every_declared_symbol(built_lib):
    names = _declared_symbols()
    assert len(names) >= 14
    handle = ctypes.CDLL(built_lib)
    for name in names:
        assert hasattr(handle, name), "missing export: %s" % name
    from codex_africanus_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == names
    lib = _lib.lib()
    assert lib.afr_version() == 100
    assert lib.afr_device_count() >= 0  # no compute without a GPU


def test_freq_uniform_rule(built_lib):
    from codex_africanus_b200 import _lib, _plumbing as pl

    assert pl.channel_mode(np.linspace(0.856e9, 1.712e9, 4096)) == _lib.AFR_CHAN_UNIFORM
    assert pl.channel_mode(0.856e9 + 208984.375 * np.arange(4096)) == _lib.AFR_CHAN_UNIFORM
    assert pl.channel_mode(np.array([1.0e9])) == _lib.AFR_CHAN_UNIFORM
    assert pl.channel_mode(np.array([1.0, 2.0, 4.0])) == _lib.AFR_CHAN_EXACT
    f = np.linspace(1e9, 2e9, 64)
    f[10] *= 1 + 1e-12
    assert pl.channel_mode(f) == _lib.AFR_CHAN_EXACT
    assert pl.channel_mode(np.linspace(1e9, 2e9, 64).astype(np.float32)) == _lib.AFR_CHAN_EXACT
    assert pl.channel_mode(np.linspace(2e9, 1e9, 64)) == _lib.AFR_CHAN_UNIFORM  # descending


def test_errors_match_reference_before_device_work():
    import codex_africanus_b200.dft as dft
    import codex_africanus_b200.rime as rime

    z3 = np.zeros((2, 1, 1))
    # rime/phase.py:33-34, dft/kernels.py:39,115
    with pytest.raises(ValueError, match="convention not in"):
        rime.phase_delay(np.zeros((1, 2)), np.zeros((1, 3)), np.ones(1), convention="bad")
    with pytest.raises(ValueError, match="convention not in"):
        dft.im_to_vis(z3, np.zeros((1, 3)), np.zeros((2, 2)), np.ones(1), convention="bad")
    with pytest.raises(ValueError, match="convention not in"):
        dft.vis_to_im(z3, np.zeros((2, 3)), np.zeros((1, 2)), np.ones(1), z3 > 0, convention="bad")
    # dft/kernels.py:97-98 and :102
    with pytest.raises(TypeError):
        dft.vis_to_im(z3, np.zeros((2, 3)), np.zeros((1, 2)), np.ones(1), z3 > 0, dtype=np.complex64)
    with pytest.raises(AssertionError):
        dft.vis_to_im(z3, np.zeros((2, 3)), np.zeros((1, 2)), np.ones(1), np.zeros((2, 1, 2), bool))
    # rime/predict.py:403-461, 557-563
    ti = a1 = a2 = np.zeros(2, np.int32)
    dde = np.zeros((1, 1, 1, 1, 2, 2), np.complex128)
    coh = np.zeros((1, 2, 1, 2), np.complex128)
    with pytest.raises(ValueError, match="Both dde1_jones and dde2_jones"):
        rime.predict_vis(ti, a1, a2, dde1_jones=dde)
    with pytest.raises(ValueError, match="Both die1_jones and die2_jones"):
        rime.predict_vis(ti, a1, a2, die2_jones=dde[0])
    with pytest.raises(ValueError, match="ndim 3 not in"):
        rime.predict_vis(ti, a1, a2, source_coh=coh[0])
    with pytest.raises(ValueError, match="pre-conditions|mismatched"):
        rime.predict_vis(ti, a1, a2, dde, coh, dde)
    with pytest.raises(ValueError, match="No Jones Matrices were supplied"):
        rime.predict_vis(ti, a1, a2)
    # rime/fast_beam_cubes.py:74-75
    with pytest.raises(ValueError, match="must be >= 2"):
        rime.beam_cube_dde(np.zeros((1, 2, 2, 1), np.complex128), np.zeros((2, 2)), np.zeros(2),
                           np.zeros((1, 2)), np.zeros((1, 1)), np.zeros((1, 1, 1, 2)),
                           np.ones((1, 1, 2)), np.ones(1))


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import codex_africanus_b200.dft as dft
    from codex_africanus_b200._lib import AfricanusB200Error

    with pytest.raises(AfricanusB200Error, match="no CPU fallback"):
        dft.im_to_vis(np.zeros((1, 1, 1)), np.zeros((1, 3)), np.zeros((1, 2)), np.ones(1))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no module of the package may reference it."""
    pkg = os.path.join(ROOT, "codex_africanus_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn
                assert "afr_oracle" not in txt, fn
