#!/usr/bin/env python
"""stokes_brightness throughput (TB/s of output) and worst relative error against float64 numpy."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import model  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
nsrc, nchan = 4096, 4096
stokes = rng.standard_normal((nsrc, 4))
freq = np.linspace(0.856e9, 1.712e9, nchan)
rf = np.full(nsrc, 1.284e9)
for nspi, base in ((1, "std"), (2, "std"), (2, "log"), (2, "log10")):
    spi = rng.uniform(-1.0, 0.2, (nsrc, nspi, 4))
    d = [torch.from_numpy(a).to(dev) for a in (stokes, spi, rf, freq)]
    out = model.stokes_brightness(*d, base=base)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(5):
        ev[0].record()
        out = model.stokes_brightness(*d, base=base)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    # numpy restatement of spec_model.py:181-208 + the linear-feed schema (I+Q, U+iV, U-iV, I-Q)
    ratio = freq[None, :] / rf[:, None]
    sm = np.empty((nsrc, nchan, 4))
    for p in range(4):
        if base == "std":
            v = stokes[:, None, p] * np.ones_like(ratio)
            for i in range(nspi):
                v = v * ratio ** spi[:, i, None, p]
        else:
            lr = np.log(ratio) if base == "log" else np.log10(ratio)
            acc = np.zeros_like(ratio)
            for i in range(nspi):
                acc = acc + spi[:, i, None, p] * lr ** (i + 1)
            v = stokes[:, None, p] * (np.exp(acc) if base == "log" else 10.0 ** acc)
        sm[:, :, p] = v
    ref = np.stack([sm[..., 0] + sm[..., 1], sm[..., 2] + 1j * sm[..., 3], sm[..., 2] - 1j * sm[..., 3],
                    sm[..., 0] - sm[..., 1]], axis=-1).reshape(nsrc, nchan, 2, 2)
    got = out.cpu().numpy().reshape(ref.shape)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print("stokes_brightness nspi=%d base=%-5s: %.2f TB/s of output (%.3f ms), max err / max ref %.2e"
          % (nspi, base, nsrc * nchan * 64.0 / best / 1e9, best, err))
