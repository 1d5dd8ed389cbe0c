"""GPU test of the row-block streaming driver: numpy blocks through the rotating page-locked
buffers and CUDA-tensor blocks both equal the one-shot predict (reference golden)."""
import numpy as np
import pytest

from conftest import assert_c128_close

pytestmark = pytest.mark.gpu


def test_stream_predict_vis_stokes(golden, oracle):
    import torch

    from codex_africanus_b200.rime.stream import stream_predict_vis_stokes

    g = golden("brightness")
    st, spi, rf, fr = g["stokes"], g["spi2"], g["ref_freq"], g["freq"]
    lm, uvw, ti, a1, a2 = g["p_lm"], g["p_uvw"], g["p_time_index"], g["p_ant1"], g["p_ant2"]
    die, bvis = g["p_die"], g["p_base_vis"]
    ntime = die.shape[0]
    nbl = uvw.shape[0] // ntime
    # numpy in -> numpy blocks; hold every block until the end of its validity window (nbuf - 1
    # further requests) before copying it out, so a premature buffer reuse would be seen
    for nbuf in (2, 3):
        out = np.zeros_like(g["p_linear"])
        held = []
        for (r0, r1), blk in stream_predict_vis_stokes(lm, uvw, fr, st, spi, rf, ti + 3, a1, a2, None, None,
                                                       die, bvis, die, rows_per_block=nbl, nbuf=nbuf,
                                                       source_chunk=7):
            assert isinstance(blk, np.ndarray) and blk.shape == (r1 - r0,) + out.shape[1:]
            held.append(((r0, r1), blk))
            if len(held) == nbuf - 1:
                (h0, h1), hb = held.pop(0)
                out[h0:h1] = hb
        for (h0, h1), hb in held:
            out[h0:h1] = hb
        assert_c128_close(out, g["p_linear"])
    # CUDA tensors in -> CUDA tensor blocks, two timesteps per block, DDEs sliced by time
    rng = np.random.default_rng(8)
    na = die.shape[1]
    dde = 1.0 + 0.2 * (rng.standard_normal((lm.shape[0], ntime, na, fr.shape[0], 2, 2))
                       + 1j * rng.standard_normal((lm.shape[0], ntime, na, fr.shape[0], 2, 2)))
    b = oracle.convert(oracle.spectral_model(st, spi, rf, fr), ["I", "Q", "U", "V"], [["XX", "XY"], ["YX", "YY"]])
    ref = oracle.fused_predict(lm, uvw, fr, b, ti, a1, a2, dde, dde, die, None, die)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    d_dde, d_die = T(dde), T(die)
    blocks = list(stream_predict_vis_stokes(T(lm), T(uvw), T(fr), T(st), T(spi), T(rf), T(ti), T(a1), T(a2),
                                            d_dde, d_dde, d_die, None, d_die, rows_per_block=2 * nbl))
    assert [rng_ for rng_, _ in blocks] == [(0, 2 * nbl), (2 * nbl, 3 * nbl)]
    assert all(isinstance(blk, torch.Tensor) and blk.is_cuda for _, blk in blocks)
    assert_c128_close(torch.cat([blk for _, blk in blocks]).cpu().numpy(), ref)


def test_stream_fused_predict_vis(golden):
    """Plain-brightness streaming: numpy blocks through the page-locked buffers."""
    from codex_africanus_b200.rime.stream import stream_fused_predict_vis

    g = golden("brightness")
    fr, lm, uvw, ti, a1, a2 = g["freq"], g["p_lm"], g["p_uvw"], g["p_time_index"], g["p_ant1"], g["p_ant2"]
    die, bvis = g["p_die"], g["p_base_vis"]
    nbl = uvw.shape[0] // die.shape[0]
    for feed in ("linear", "circular"):
        parts = [blk.copy() for _, blk in stream_fused_predict_vis(
            lm, uvw, fr, g["b_" + feed], ti, a1, a2, None, None, die, bvis, die, rows_per_block=nbl)]
        assert len(parts) == die.shape[0]
        assert_c128_close(np.concatenate(parts), g["p_" + feed])


def test_stream_fused_predict_vis_beam(golden):
    """Beam-interpolated streaming against the reference outputs (rotated DDE through the predict)."""
    from codex_africanus_b200.rime.stream import stream_fused_predict_vis_beam

    g = golden("feeds")
    ntime = g["pa"].shape[0]
    nbl = g["uvw"].shape[0] // ntime
    for ft, rpb in (("linear", nbl), ("circular", 2 * nbl)):
        parts = [blk.copy() for _, blk in stream_fused_predict_vis_beam(
            g["lm"], g["uvw"], g["freq"], g["bright"], g["time_index"], g["ant1"], g["ant2"], g["beam"],
            g["ext"], g["bfm"], g["pa"], g["pe"], g["asc"], g["die"], None, g["die"], rows_per_block=rpb,
            feed_type=ft, source_chunk=4)]
        assert_c128_close(np.concatenate(parts), g["vis_" + ft])
