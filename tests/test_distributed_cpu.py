"""
world_size-2 gloo tests (CPU) of the multi-GPU host logic: timestep-aligned row sharding,
the gather of predict blocks and the all_reduce of vis_to_im partial images.  The local
compute is the CPU oracle standing in for the CUDA entry points (tests may use it).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    rng = np.random.default_rng(9)
    na, ntime, nchan, nsrc = 5, 7, 6, 11
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size) + 4
    nrow = ti.size
    uvw = rng.standard_normal((nrow, 3)) * 500.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(1e9, 1.4e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    return dict(ant1=ant1, ant2=ant2, ti=ti, uvw=uvw, lm=lm, freq=freq, na=na, ntime=ntime,
                image=rng.standard_normal((nsrc, nchan, 2)), vis=rc((nrow, nchan, 2)),
                flags=rng.random((nrow, nchan, 2)) < 0.1, bright=rc((nsrc, nchan, 2, 2)),
                dde=rc((nsrc, ntime, na, nchan, 2, 2)), die=rc((ntime, na, nchan, 2, 2)),
                bvis=rc((nrow, nchan, 2, 2)), stokes=rng.standard_normal((nsrc, 4)),
                spi=rng.standard_normal((nsrc, 2, 4)) * 0.3, rf=np.full(nsrc, 1.2e9))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from codex_africanus_b200 import distributed as D

    p = _problem()
    vis, _ = D.sharded_im_to_vis(p["image"], p["uvw"], p["lm"], p["freq"], p["ti"], gather=True,
                                 local_fn=oracle.im_to_vis)
    img = D.sharded_vis_to_im(p["vis"], p["uvw"], p["lm"], p["freq"], p["flags"], p["ti"],
                              local_fn=oracle.vis_to_im)
    pred, _ = D.sharded_fused_predict_vis(
        p["lm"], p["uvw"], p["freq"], p["bright"], p["ti"], p["ant1"], p["ant2"], p["dde"],
        p["dde"], p["die"], p["bvis"], p["die"], gather=True, local_fn=oracle.fused_predict)
    # SKA-style streaming: each rank streams its own shard in one-timestep blocks, brightness from
    # Stokes parameters (oracle spectral_model + convert standing in for the CUDA kernels)
    def stokes_fn(lm, uvw, fr, st, spi, rf, ti, a1, a2, e1, e2, g1, bv, g2, **kw):
        b = oracle.convert(oracle.spectral_model(st, spi, rf, fr), ["I", "Q", "U", "V"],
                           [["XX", "XY"], ["YX", "YY"]])
        return oracle.fused_predict(lm, uvw, fr, b, ti, a1, a2, e1, e2, g1, bv, g2)

    nbl = p["ant1"].size // p["ntime"]
    rows, parts = [], []
    for (b0, b1), blk in D.sharded_stream_predict_vis_stokes(
            p["lm"], p["uvw"], p["freq"], p["stokes"], p["spi"], p["rf"], p["ti"], p["ant1"], p["ant2"],
            p["dde"], p["dde"], p["die"], p["bvis"], p["die"], rows_per_block=nbl, local_fn=stokes_fn):
        rows.append((b0, b1))
        parts.append(blk)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), vis=vis, img=img, pred=pred,
             srows=np.array(rows), spred=np.concatenate(parts))
    dist.barrier()
    dist.destroy_process_group()


def test_row_shards_are_timestep_aligned():
    from codex_africanus_b200.distributed import row_shards

    ti = np.repeat(np.arange(10), 6) + 3
    for world in (1, 2, 3, 4, 8, 16):
        shards = row_shards(ti, world)
        assert len(shards) == world
        assert shards[0][0] == 0 and shards[-1][1] == ti.size
        for (a0, a1), (b0, b1) in zip(shards[:-1], shards[1:]):
            assert a1 == b0
        for s0, s1 in shards:
            assert s0 % 6 == 0 and s1 % 6 == 0  # whole timesteps only
    assert row_shards(np.zeros(0, int), 3) == [(0, 0)] * 3
    with pytest.raises(ValueError):
        row_shards(np.array([1, 0, 2]), 2)
    # ragged timesteps
    ti = np.array([0, 0, 0, 1, 2, 2, 5, 5, 5, 5])
    shards = row_shards(ti, 2)
    assert shards == [(0, 4), (4, 10)]


def test_world2_gloo_matches_single_process(tmp_path):
    import oracle

    oracle.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p = _problem()
    ref_vis = oracle.im_to_vis(p["image"], p["uvw"], p["lm"], p["freq"])
    ref_img = oracle.vis_to_im(p["vis"], p["uvw"], p["lm"], p["freq"], p["flags"])
    ref_pred = oracle.fused_predict(p["lm"], p["uvw"], p["freq"], p["bright"], p["ti"], p["ant1"],
                                    p["ant2"], p["dde"], p["dde"], p["die"], p["bvis"], p["die"])
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        np.testing.assert_array_equal(got["vis"], ref_vis)       # rows are independent
        np.testing.assert_array_equal(got["pred"], ref_pred)
        np.testing.assert_allclose(got["img"], ref_img, rtol=1e-12, atol=1e-12 * np.abs(ref_img).max())
    # streamed shards: rank r's one-timestep blocks tile its shard; together they are the whole predict
    from codex_africanus_b200.distributed import row_shards

    b = oracle.convert(oracle.spectral_model(p["stokes"], p["spi"], p["rf"], p["freq"]), ["I", "Q", "U", "V"],
                       [["XX", "XY"], ["YX", "YY"]])
    ref_s = oracle.fused_predict(p["lm"], p["uvw"], p["freq"], b, p["ti"], p["ant1"], p["ant2"], p["dde"],
                                 p["dde"], p["die"], p["bvis"], p["die"])
    nbl = p["ant1"].size // p["ntime"]
    for rank, (s0, s1) in enumerate(row_shards(p["ti"], 2)):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert got["srows"].tolist() == [[r, r + nbl] for r in range(s0, s1, nbl)]
        np.testing.assert_array_equal(got["spred"], ref_s[s0:s1])
