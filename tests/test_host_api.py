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


def test_convert_schema_resolution_and_errors():
    """Host logic of model.convert / spectral_model (conversion.py:93-215, spec_model.py:77-171):
    schema flattening, rule choice, output dtype, and the reference's errors -- all before any
    device work."""
    from codex_africanus_b200.model import coherency as coh, convert, spectral_model
    from codex_africanus_b200.model.spectral import promote_bases

    names, shape = coh.schema_elements([["XX", "XY"], ["YX", "YY"]])
    assert names == {"XX": 0, "XY": 1, "YX": 2, "YY": 3} and shape == (2, 2)
    assert coh.schema_elements([9, 12]) == ({"XX": 0, "YY": 1}, (2,))
    assert coh.schema_elements("I") == ({"I": 0}, (1,))
    iquv = coh.schema_elements(["I", "Q", "U", "V"])[0]
    s1, s2, op = coh.resolve(iquv, names, False)
    assert (list(s1), list(s2), list(op)) == ([0, 2, 2, 0], [1, 3, 3, 1], [0, 2, 3, 1])
    s1, s2, op = coh.resolve(iquv, coh.schema_elements([["RR", "RL"], ["LR", "LL"]])[0], False)
    assert (list(s1), list(s2), list(op)) == ([0, 1, 1, 0], [3, 2, 2, 3], [0, 2, 3, 1])
    # implicit Stokes: missing inputs are the default zero (-1)
    s1, s2, op = coh.resolve({"I": 0}, names, True)
    assert (list(s1), list(s2)) == ([0, -1, -1, 0], [-1, -1, -1, -1])
    # both circular and linear inputs: first listed rule wins ties (XX,YY for I)
    both = coh.schema_elements(["XX", "YY", "RR", "LL"])[0]
    assert [list(x) for x in coh.resolve(both, {"I": 0, "V": 1}, False)] == [[0, 2], [1, 3], [4, 5]]
    assert coh._output_dtype(np.dtype(np.float64), [0, 2]) == np.complex128
    assert coh._output_dtype(np.dtype(np.float32), [0, 2]) == np.complex64
    assert coh._output_dtype(np.dtype(np.float32), [4, 5]) == np.float32
    assert coh._output_dtype(np.dtype(np.float64), [4, 6]) == np.complex128
    with pytest.raises(ValueError, match="defined multiple times"):
        coh.schema_elements(["I", "I"])
    with pytest.raises(coh.DimensionMismatch):
        coh.schema_elements([["XX", "XY"], ["YX"]])
    with pytest.raises(TypeError):
        coh.schema_elements([1.5])
    with pytest.raises(ValueError, match="Invalid id"):
        coh.schema_elements([99])
    with pytest.raises(coh.MissingConversionInputs):
        coh.resolve({"I": 0}, names, False)
    with pytest.raises(ValueError, match="Unknown output"):
        coh.resolve(iquv, {"PP": 0}, False)
    with pytest.raises(ValueError, match="doesn't match input schema"):
        convert(np.zeros((3, 3)), ["I", "Q", "U", "V"], ["XX"])
    assert list(promote_bases("log", 3)) == [1, 1, 1]
    assert list(promote_bases(["std", 2], 4)) == [0, 2, 2, 2]
    with pytest.raises(ValueError, match="Invalid base"):
        promote_bases("ln", 2)
    with pytest.raises(TypeError):
        promote_bases(1.0, 2)
    st, spi = np.ones((3, 4)), np.ones((3, 2, 4))
    with pytest.raises(ValueError, match="Dimensions on stokes and spi"):
        spectral_model(st, spi[:, :, 0], np.ones(3), np.ones(5))
    with pytest.raises(ValueError, match="Correlations on stokes and spi"):
        spectral_model(st, spi[:, :, :2], np.ones(3), np.ones(5))


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


def test_bench_reference_arm_emits_one_json_line():
    """bench.py contract: exactly ONE JSON line on stdout (library banners go to stderr), with the
    keys the driver reads; the reference arm runs on the CPU, so this is checkable without a GPU."""
    import json
    import subprocess
    import sys

    env = dict(os.environ, BENCH_REF_SECONDS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:300]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Gterms/s" and line["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    # the reference's own numba kernel when baseline/_ref holds it, else the oracle's C port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline_port"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def test_layout_adapter_host_logic():
    """Host pieces of the DDE layout adapter (rime/fused.py): diagonal terms embedded as diagonal 2x2
    matrices in complex128, the time-order test on numpy and torch indices."""
    import torch
    from codex_africanus_b200.rime import fused
    rng = np.random.default_rng(0)
    x2 = torch.from_numpy((rng.standard_normal((3, 4, 2)) + 1j * rng.standard_normal((3, 4, 2))).astype(np.complex64))
    e = fused._embed_2x2(x2, (2,))
    assert e.dtype == torch.complex128 and tuple(e.shape) == (3, 4, 2, 2)
    assert torch.equal(e[..., 0, 0], x2[..., 0].to(torch.complex128))
    assert torch.equal(e[..., 1, 1], x2[..., 1].to(torch.complex128))
    assert not e[..., 0, 1].any() and not e[..., 1, 0].any()
    x1 = x2[..., :1]
    e = fused._embed_2x2(x1, (1,))
    assert torch.equal(e[..., 0, 0], x1[..., 0].to(torch.complex128)) and not e[..., 1, 1].any()
    full = torch.from_numpy(rng.standard_normal((2, 2, 2)) + 1j * rng.standard_normal((2, 2, 2)))
    assert torch.equal(fused._embed_2x2(full, (2, 2)), full)
    assert fused._time_ordered(np.array([3, 3, 4, 9])) and not fused._time_ordered(np.array([3, 2, 4]))
    assert fused._time_ordered(torch.tensor([1, 1, 2])) and not fused._time_ordered(torch.tensor([2, 1]))
    assert fused._time_ordered(np.array([], dtype=np.int64)) and fused._time_ordered([5])


def test_fused_spec_parser_and_term_rules():
    """Grammar and term rules of the fused-RIME specification front end (reference:
    experimental/rime/fused/specification.py:78-115,166-185,440-453) -- host logic only."""
    from codex_africanus_b200.rime import fused_spec as fs
    terms, stokes, corrs = fs.parse_rime("(Lp, Ep, Kpq, Bpq, Eq, Lq): [I,Q,U,V] -> [XX,XY,YX,YY]")
    assert terms == ["Lp", "Ep", "Kpq", "Bpq", "Eq", "Lq"] and stokes == ["I", "Q", "U", "V"]
    assert corrs == ["XX", "XY", "YX", "YY"]
    assert fs.parse_rime(fs.DEFAULT_SPEC)[0] == ["Kpq", "Bpq"]
    assert fs.parse_rime("[Kpq, Bpq,]: [I, Q] -> [RR, LL]") == (["Kpq", "Bpq"], ["I", "Q"], ["RR", "LL"])
    for bad in ("(Kpq, Bpq)", "(Kpq, Bpq): [I,Q,U,V]", "Kpq, Bpq: [I] -> [XX]", "(Kpq, (Bpq)): [I] -> [XX]",
                "(Kpq, Bpq): I,Q -> [XX]"):
        with pytest.raises(fs.RimeParseError):
            fs.parse_rime(bad)
    assert fs.split_terms(["Lp", "Ep", "Kpq", "Bpq", "Eq", "Lq"]) == (["L", "E"], ["K", "B"], ["E", "L"])
    assert fs.split_terms(["Bpq"]) == ([], ["B"], [])
    for bad in (["Ep", "Kpq", "Bpq", "Lq"], ["Kpq", "Ep", "Bpq", "Eq"], ["Kp", "Bpq", "Kq"], ["Zp", "Kpq", "Bpq", "Zq"],
                ["Kpq", "Bpq", "Kpq"], ["kpq", "Bpq"]):
        with pytest.raises(fs.RimeSpecificationError):
            fs.split_terms(bad)
    with pytest.raises(NotImplementedError):
        fs.split_terms(["Cpq", "Kpq", "Bpq"])
    assert fs.feed_type_of(["XX", "YY"]) == "linear" and fs.feed_type_of(["RR", "RL", "LR", "LL"]) == "circular"
    with pytest.raises(fs.RimeSpecificationError):
        fs.feed_type_of(["XX", "RR"])
    m = fs.consolidate_args(({"TIME": 1, "uvw": 2}, 10, 11), {"convention": "casa"})
    assert m == {"time": 10, "uvw": 2, "antenna1": 11, "convention": "casa"}
    # lm of transformers/lm.py against the direction cosines of a small offset
    lm = fs.radec_to_lm(np.array([[0.01, -0.5], [0.0, -0.49]]), np.array([0.0, -0.5]))
    assert np.allclose(lm[0], [np.cos(-0.5) * np.sin(0.01), np.sin(-0.5) * np.cos(-0.5) * (1 - np.cos(0.01))])
    assert np.allclose(lm[1], [0.0, np.sin(0.01)])
