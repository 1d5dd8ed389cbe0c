"""
Hardware parity of the sharded multi-GPU path: CUDA kernels + whole-timestep row
shards + NCCL collectives on >= 2 GPUs of one box, compared with the ONE-GPU result of the same
call and with the CPU oracle.  Skipped on a single-GPU box (the driver's round-end box); run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q` (log under profiles/).

  * sharded_vis_to_im: per-rank partial dirty image + ONE all_reduce(SUM) (africanus/dft/dask.py:71-90)
  * sharded_fused_predict_vis (DIE + DDE, antenna-consistent uvw -> antenna-mode GEMM kernel) and
    sharded_im_to_vis: no data-path collective, final all_gather of the row blocks
    (africanus/rime/dask_predict.py:667-726)
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem(ntime=7):
    rng = np.random.default_rng(77)
    na, nchan, nsrc = 12, 96, 33
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime).astype(np.int32), np.tile(a2, ntime).astype(np.int32)
    ti = (np.repeat(np.arange(ntime), a1.size) + 5).astype(np.int32)
    antpos = rng.standard_normal((ntime, na, 3)) * 2500.0
    uvw = antpos[ti - 5, ant1] - antpos[ti - 5, ant2]
    nrow = ti.size
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    return dict(ant1=ant1, ant2=ant2, ti=ti, uvw=uvw, lm=lm, freq=freq,
                image=rng.standard_normal((nsrc, nchan, 1)), vis=rc((nrow, nchan, 1)),
                flags=rng.random((nrow, nchan, 1)) < 0.05, bright=rc((nsrc, nchan, 2, 2)),
                dde=np.eye(2) + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2)),
                die=np.eye(2) + 0.1 * rc((ntime, na, nchan, 2, 2)), bvis=rc((nrow, nchan, 2, 2)))


def _worker(rank, world, port, out_dir, ntime):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from codex_africanus_b200 import _lib
    from codex_africanus_b200 import distributed as D

    p = _problem(ntime)
    vis, _ = D.sharded_im_to_vis(p["image"], p["uvw"], p["lm"], p["freq"], p["ti"], gather=True)
    img = D.sharded_vis_to_im(p["vis"], p["uvw"], p["lm"], p["freq"], p["flags"], p["ti"])
    pred, _ = D.sharded_fused_predict_vis(p["lm"], p["uvw"], p["freq"], p["bright"], p["ti"], p["ant1"],
                                          p["ant2"], p["dde"], p["dde"], p["die"], p["bvis"], p["die"],
                                          gather=True)
    r0, r1 = D.row_shards(p["ti"], world)[rank]
    path = _lib.lib().afr_last_fused_path() if r1 > r0 else -1  # -1: this rank's shard is empty
    # device-resident variant: torch CUDA tensors in, the all_reduce runs on the kernel's output
    dev = torch.device("cuda", rank)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    img_t = D.sharded_vis_to_im(t(p["vis"]), t(p["uvw"]), t(p["lm"]), t(p["freq"]), t(p["flags"]), t(p["ti"]))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), vis=vis, img=img, pred=pred, path=path,
             img_t=img_t.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ntime", [(2, 7), (2, 1), (4, 7), (8, 7)])
def test_nccl_sharded_paths_match_one_gpu_and_oracle(tmp_path, oracle, world, ntime):
    """7 timesteps over 2 / 4 / 8 ranks (at 8 ranks one shard is empty); 1 timestep over 2 ranks
    (rank 0 empty): an empty shard contributes nothing and must not upset the collectives."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from conftest import assert_c128_close

    from codex_africanus_b200 import dft, rime

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), ntime), nprocs=world, join=True)
    p = _problem(ntime)
    one_vis = dft.im_to_vis(p["image"], p["uvw"], p["lm"], p["freq"])
    one_img = dft.vis_to_im(p["vis"], p["uvw"], p["lm"], p["freq"], p["flags"])
    one_pred = rime.fused_predict_vis(p["lm"], p["uvw"], p["freq"], p["bright"], p["ti"], p["ant1"], p["ant2"],
                                      p["dde"], p["dde"], p["die"], p["bvis"], p["die"])
    ref_img = oracle.vis_to_im(p["vis"], p["uvw"], p["lm"], p["freq"], p["flags"])
    ref_pred = oracle.fused_predict(p["lm"], p["uvw"], p["freq"], p["bright"], p["ti"], p["ant1"], p["ant2"],
                                    p["dde"], p["dde"], p["die"], p["bvis"], p["die"])
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert int(got["path"]) in (6, -1)  # every non-empty shard ran the antenna-mode GEMM kernel
        # rows are independent: the gathered blocks are the one-GPU rows (same kernels, same order
        # of the source sum); the image differs by the summation order over row shards only
        assert_c128_close(got["vis"], one_vis, rtol=1e-13)
        assert_c128_close(got["pred"], one_pred, rtol=1e-13)
        assert_c128_close(got["img"], one_img, rtol=1e-12)
        assert_c128_close(got["img_t"], one_img, rtol=1e-12)
        assert_c128_close(got["img"], ref_img)
        assert_c128_close(got["pred"], ref_pred)
