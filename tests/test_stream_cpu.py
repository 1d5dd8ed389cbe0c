"""CPU tests of the row-block streaming driver (host logic only; the per-block predict is
replaced by the oracle through ``local_fn``, exactly as tests/test_distributed_cpu.py does for the
multi-GPU drivers)."""
import numpy as np
import pytest

_LIN = [["XX", "XY"], ["YX", "YY"]]
_IQUV = ["I", "Q", "U", "V"]


def test_timestep_row_blocks():
    from codex_africanus_b200.rime.stream import timestep_row_blocks as trb

    ti = np.repeat(np.arange(5) + 7, [3, 3, 2, 4, 3])  # 15 rows, ragged timesteps, offset 7
    assert trb(ti, 6) == [(0, 6, 0, 2), (6, 12, 2, 4), (12, 15, 4, 5)]
    assert trb(ti, 1) == [(0, 3, 0, 1), (3, 6, 1, 2), (6, 8, 2, 3), (8, 12, 3, 4), (12, 15, 4, 5)]
    assert trb(ti, 100) == [(0, 15, 0, 5)]
    assert trb(ti[:0], 4) == []
    gap = np.array([0, 0, 3, 3, 9])  # missing timesteps: the t-range spans the values present
    assert trb(gap, 2) == [(0, 2, 0, 1), (2, 4, 3, 4), (4, 5, 9, 10)]
    with pytest.raises(ValueError):
        trb(np.array([1, 0]), 2)
    # every row exactly once, in order
    for n in (2, 5, 7, 15):
        blocks = trb(ti, n)
        assert [b[0] for b in blocks][1:] == [b[1] for b in blocks][:-1]
        assert blocks[0][0] == 0 and blocks[-1][1] == ti.size


def test_stream_predict_blocks_equal_whole(oracle, golden):
    """Blocks of whole timesteps, each predicted with its own (time, ...) slices of the DDE / DIE
    arrays, concatenate to the one-shot predict (reference golden)."""
    from codex_africanus_b200.rime.stream import stream_predict_vis_stokes

    g = golden("brightness")
    st, spi, rf, fr = g["stokes"], g["spi2"], g["ref_freq"], g["freq"]
    lm, uvw, ti, a1, a2 = g["p_lm"], g["p_uvw"], g["p_time_index"], g["p_ant1"], g["p_ant2"]
    die, bvis = g["p_die"], g["p_base_vis"]
    rng = np.random.default_rng(8)
    ntime, na = die.shape[:2]
    dde = 1.0 + 0.2 * (rng.standard_normal((lm.shape[0], ntime, na, fr.shape[0], 2, 2))
                       + 1j * rng.standard_normal((lm.shape[0], ntime, na, fr.shape[0], 2, 2)))
    calls = []

    def local_fn(lm_, uvw_, fr_, st_, spi_, rf_, ti_, a1_, a2_, e1, e2, g1, bv, g2, **kw):
        calls.append((uvw_.shape[0], None if g1 is None else g1.shape[0]))
        b = oracle.convert(oracle.spectral_model(st_, spi_, rf_, fr_, base=kw.get("base", 0)), _IQUV,
                           kw.get("corr_schema", _LIN))
        return oracle.fused_predict(lm_, uvw_, fr_, b, ti_, a1_, a2_, e1, e2, g1, bv, g2)

    nbl = uvw.shape[0] // ntime
    out = np.zeros_like(g["p_linear"])
    for (r0, r1), blk in stream_predict_vis_stokes(lm, uvw, fr, st, spi, rf, ti + 5, a1, a2, None, None, die,
                                                   bvis, die, rows_per_block=nbl, local_fn=local_fn):
        out[r0:r1] = blk
    assert calls == [(nbl, 1)] * ntime
    assert np.array_equal(out, g["p_linear"])
    # with DDEs, two timesteps per block (last block has one)
    b = oracle.convert(oracle.spectral_model(st, spi, rf, fr), _IQUV, _LIN)
    ref = oracle.fused_predict(lm, uvw, fr, b, ti, a1, a2, dde, dde, die, None, die)
    calls.clear()
    got = np.concatenate([blk for _, blk in stream_predict_vis_stokes(
        lm, uvw, fr, st, spi, rf, ti, a1, a2, dde, dde, die, None, die, rows_per_block=2 * nbl,
        local_fn=local_fn)])
    assert calls == [(2 * nbl, 2), (nbl, 1)]
    assert np.array_equal(got, ref)


def test_stream_fused_predict_blocks_equal_whole(oracle, golden):
    """The plain-brightness variant: blocks concatenate to the one-shot fused predict."""
    from codex_africanus_b200.rime.stream import stream_fused_predict_vis

    g = golden("brightness")
    fr, lm, uvw, ti, a1, a2 = g["freq"], g["p_lm"], g["p_uvw"], g["p_time_index"], g["p_ant1"], g["p_ant2"]
    die, bvis, bright = g["p_die"], g["p_base_vis"], g["b_circular"]
    ntime = die.shape[0]
    nbl = uvw.shape[0] // ntime
    rows = []
    parts = []
    for (r0, r1), blk in stream_fused_predict_vis(lm, uvw, fr, bright, ti, a1, a2, None, None, die, bvis, die,
                                                  convention="fourier", rows_per_block=nbl + 1,
                                                  local_fn=oracle.fused_predict):
        rows.append((r0, r1))
        parts.append(blk)
    assert rows == [(k * nbl, (k + 1) * nbl) for k in range(ntime)]
    assert np.array_equal(np.concatenate(parts), g["p_circular"])
    with pytest.raises(ValueError):
        stream_fused_predict_vis(lm, uvw, fr, bright, ti[:-1], a1, a2, local_fn=oracle.fused_predict)


def test_stream_beam_predict_blocks_equal_whole(oracle, golden):
    """Beam-interpolated variant: parallactic angles / pointing errors / DIEs sliced per block; the
    blocks concatenate to the reference composition with the rotated DDE (golden)."""
    from codex_africanus_b200.rime.stream import stream_fused_predict_vis_beam

    g = golden("feeds")
    seen = []

    def local_fn(lm, uvw, fr, b, ti, a1, a2, beam, ext, bfm, pa, pe, asc, g1, bv, g2, feed_type=None):
        seen.append((uvw.shape[0], pa.shape[0], pe.shape[0], g1.shape[0]))
        dde = oracle.beam_cube_dde(beam, ext, bfm, lm, pa, pe, asc, fr)
        dde = np.einsum("stafij,tajk->stafik", dde, oracle.feed_rotation(pa, feed_type))
        return oracle.fused_predict(lm, uvw, fr, b, ti, a1, a2, dde, dde, g1, bv, g2)

    ntime = g["pa"].shape[0]
    nbl = g["uvw"].shape[0] // ntime
    for ft in ("linear", "circular"):
        seen.clear()
        parts = [blk for _, blk in stream_fused_predict_vis_beam(
            g["lm"], g["uvw"], g["freq"], g["bright"], g["time_index"], g["ant1"], g["ant2"], g["beam"],
            g["ext"], g["bfm"], g["pa"], g["pe"], g["asc"], g["die"], None, g["die"], rows_per_block=nbl,
            local_fn=local_fn, feed_type=ft)]
        assert seen == [(nbl, 1, 1, 1)] * ntime
        assert np.array_equal(np.concatenate(parts), g["vis_" + ft])
