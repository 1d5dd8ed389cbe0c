"""Small end-to-end calls of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import dft, rime
rng = np.random.default_rng(1)
na, ntime, nchan, nsrc = 6, 3, 40, 19
a1, a2 = np.triu_indices(na, 1)
ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
ti = np.repeat(np.arange(ntime), a1.size)
nrow = ti.size
uvw = rng.standard_normal((nrow, 3)) * 2000.0
lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
freq = np.linspace(0.856e9, 1.712e9, nchan)
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
for ncorr in (1, 2, 3):
    img = rng.standard_normal((nsrc, nchan, ncorr))
    v = dft.im_to_vis(img, uvw, lm, freq)
    v = dft.im_to_vis(img + 1j * img, uvw, lm, freq)
    fl = rng.random(v.shape) < 0.1
    dft.vis_to_im(v, uvw, lm, freq, fl)
    dft.vis_to_im(v, uvw, lm, freq, fl, dtype=np.float32)
    dft.im_to_vis(img, uvw, lm, freq, dtype=np.complex64)
dft.im_to_vis(img, uvw, lm, np.sort(rng.uniform(1e9, 2e9, nchan)))
bright = rc((nsrc, nchan, 2, 2)); dde = rc((nsrc, ntime, na, nchan, 2, 2)); die = rc((ntime, na, nchan, 2, 2))
rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2)
rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, None, die)
rime.fused_predict_vis(lm, uvw, freq, bright[..., 0], ti, ant1, ant2, dde[..., 0], dde[..., 0])
K = rime.phase_delay(lm, uvw, freq)
rime.predict_vis(ti, ant1, ant2, dde, np.einsum("srf,sfij->srfij", K, bright), dde, die, None, die)
beam = rc((9, 9, 5, 2, 2))
rime.beam_cube_dde(beam, np.array([[-0.03, 0.03], [-0.03, 0.03]]), np.linspace(0.8e9, 1.8e9, 5), lm,
                   rng.uniform(-1, 1, (ntime, na)), np.zeros((ntime, na, nchan, 2)), np.ones((na, nchan, 2)), freq)
# warp-specialised DDE kernel: antenna-phasor mode (uvw = differences of antenna coordinates),
# per-row mode, E1 != E2, non-uniform channels; TMA-staged adjoint with 16-sample flag groups
antpos = rng.standard_normal((ntime, na, 3)) * 1500.0
uvw_a = antpos[ti, ant1] - antpos[ti, ant2]
dde_b = rc((nsrc, ntime, na, nchan, 2, 2))
rime.fused_predict_vis(lm, uvw_a, freq, bright, ti, ant1, ant2, dde, dde_b, die, None, die)
rime.fused_predict_vis(lm, uvw_a, freq, bright, ti, ant1, ant2, dde, dde)
rime.fused_predict_vis(lm, uvw, np.sort(rng.uniform(1e9, 2e9, nchan)), bright, ti, ant1, ant2, dde, dde_b)
freq48 = np.linspace(0.856e9, 1.712e9, 48)
for ncorr in (1, 2, 4):
    v = rc((nrow, 48, ncorr))
    dft.vis_to_im(v, uvw, lm, freq48, rng.random(v.shape) < 0.1)
    dft.vis_to_im(v, uvw, lm, freq48, np.zeros(v.shape, bool))
# antenna mode with > 512 rows per timestep: 2048-row x 1-channel tiles, row PAIRS sharing antenna 1
# (ragged 4 x 16 antenna tiles, straddling pairs, a partial last tile)
na_w, nt_w, nc_w, ns_w = 37, 2, 3, 4
b1, b2 = np.triu_indices(na_w, 1)
w1, w2 = np.tile(b1, nt_w), np.tile(b2, nt_w)
tw = np.repeat(np.arange(nt_w), b1.size)
pos_w = rng.standard_normal((nt_w, na_w, 3)) * 1500.0
uvw_w = pos_w[tw, w1] - pos_w[tw, w2]
fw = np.linspace(0.856e9, 1.712e9, nc_w)
dde_w = rc((ns_w, nt_w, na_w, nc_w, 2, 2))
rime.fused_predict_vis(lm[:ns_w], uvw_w, fw, rc((ns_w, nc_w, 2, 2)), tw, w1, w2, dde_w, dde_w)
# brightness from Stokes parameters, convert, feed rotation in the beam epilogue, streaming driver
from codex_africanus_b200 import model
stokes = rng.standard_normal((nsrc, 4)); spi = rng.standard_normal((nsrc, 2, 4)) * 0.3; rf = np.full(nsrc, 1.2e9)
for base in ("std", "log", ["log10", "std"]):
    model.spectral_model(stokes, spi, rf, freq, base=base)
    model.stokes_brightness(stokes, spi, rf, freq, base=base)
model.spectral_model(stokes[:, 0], spi[:, :, 0], rf, freq)
model.convert(model.convert(stokes, ["I", "Q", "U", "V"], [["RR", "RL"], ["LR", "LL"]]),
              [["RR", "RL"], ["LR", "LL"]], ["I", "V"])
model.convert(stokes[:, :1], ["I"], ["XX", "XY", "YX", "YY"], implicit_stokes=True)
pa = rng.uniform(-1, 1, (ntime, na))
for ft in ("linear", "circular"):
    rime.beam_cube_dde_rotated(beam, np.array([[-0.03, 0.03], [-0.03, 0.03]]), np.linspace(0.8e9, 1.8e9, 5), lm,
                               pa, np.zeros((ntime, na, nchan, 2)), np.ones((na, nchan, 2)), freq,
                               rime.feed_rotation(pa, ft))
rime.beam_cube_dde(beam[..., 0, :], np.array([[-0.03, 0.03], [-0.03, 0.03]]), np.linspace(0.8e9, 1.8e9, 5), lm,
                   pa, np.zeros((ntime, na, nchan, 2)), np.ones((na, nchan, 2)), freq)
rime.fused_predict_vis_stokes(lm, uvw_a, freq, stokes, spi, rf, ti, ant1, ant2, dde, dde, die, None, die,
                              source_chunk=7)
for _, blk in rime.stream_predict_vis_stokes(lm, uvw_a, freq, stokes, spi, rf, ti, ant1, ant2, None, None, die,
                                             None, die, rows_per_block=a1.size):
    blk.sum()
# plane-interpolated beam kernel (needs >= 64 channels), rows with per-channel pointing errors
# falling back to the element path, out-of-band channels, the feed-rotation epilogue
f80 = np.linspace(0.75e9, 1.85e9, 80)
pe80 = np.zeros((ntime, na, 80, 2)); pe80[1] = rng.uniform(-1e-3, 1e-3, (na, 80, 2))
for bm in (beam, beam[..., 0, :], beam[..., 0, :1]):
    rime.beam_cube_dde(bm, np.array([[-0.03, 0.03], [-0.03, 0.03]]), np.linspace(0.8e9, 1.8e9, 5), lm, pa, pe80,
                       np.ones((na, 80, 2)), f80)
rime.beam_cube_dde_rotated(beam.astype(np.complex64), np.array([[-0.03, 0.03], [-0.03, 0.03]]),
                           np.linspace(0.8e9, 1.8e9, 5), lm, pa, pe80, np.ones((na, 80, 2)), f80,
                           rime.feed_rotation(pa, "linear").astype(np.complex64))
# antenna-mode GEMM kernel over several panels (70 antennas: 9 tile rows in panels of 6), E1 != E2,
# an odd number of sources, rows with antenna1 > antenna2
na_m, ns_m, nc_m = 70, 5, 3
m1, m2 = np.triu_indices(na_m, 1)
m1, m2 = np.where(np.arange(m1.size) % 7 == 0, m2, m1), np.where(np.arange(m1.size) % 7 == 0, m1, m2)
pos_m = rng.standard_normal((1, na_m, 3)) * 1500.0
uvw_m = pos_m[0, m1] - pos_m[0, m2]
dm1, dm2 = rc((ns_m, 1, na_m, nc_m, 2, 2)), rc((ns_m, 1, na_m, nc_m, 2, 2))
rime.fused_predict_vis(lm[:ns_m], uvw_m, fw, rc((ns_m, nc_m, 2, 2)), np.zeros(m1.size, int), m1, m2, dm1, dm2)
rime.fused_predict_vis(lm[:ns_m], uvw_m, fw, rc((ns_m, nc_m, 2, 2)), np.zeros(m1.size, int), m1, m2, dm1, dm1)
# padded image grid beyond the unit disc (NaN sources muted / poisoned), im_to_vis
lm_big = np.stack(np.meshgrid(np.linspace(-1.2, 1.2, 5), np.linspace(-1.2, 1.2, 5)), -1).reshape(-1, 2)
img_big = rng.standard_normal((25, nchan, 2)); img_big[(lm_big ** 2).sum(1) >= 1] = 0; img_big[0, 3, 1] = 1.0
dft.im_to_vis(img_big, uvw, lm_big, freq)
print("sanitize target done")
