"""
Seeded synthetic MeerKAT / SKA-Mid shaped inputs for bench.py and the large-size tests
(SURVEY.md 8d).  Pure numpy; nothing here is on the product path.

Row layout: time-major, within a timestep every a1 < a2 pair (np.triu_indices(na, 1)).
"""
import numpy as np

LIGHTSPEED = 2.99792458e8


def antenna_positions(na, rng, core_sigma=1000.0, max_radius=8000.0):
    """ENU positions (m): 80% in a Gaussian core, 20% out to max_radius."""
    ncore = int(round(0.8 * na))
    core = rng.normal(0.0, core_sigma, (ncore, 2))
    r = rng.uniform(core_sigma, max_radius, na - ncore)
    th = rng.uniform(0, 2 * np.pi, na - ncore)
    outer = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    en = np.concatenate([core, outer], axis=0)
    up = rng.normal(0.0, 5.0, (na, 1))
    return np.concatenate([en, up], axis=1)


def uvw_tracks(na, ntime, rng, t0=0, ntime_total=None, hours=8.0, dec_deg=-30.0,
               lat_deg=-30.7, max_radius=8000.0, seed_positions=1234):
    """(ntime*nbl, 3) uvw in metres for timesteps [t0, t0+ntime) of an `hours`-long track of
    ntime_total steps, plus time_index, antenna1, antenna2."""
    prng = np.random.default_rng(seed_positions)
    enu = antenna_positions(na, prng, max_radius=max_radius)
    lat = np.deg2rad(lat_deg)
    dec = np.deg2rad(dec_deg)
    e, n, u = enu[:, 0], enu[:, 1], enu[:, 2]
    x = -np.sin(lat) * n + np.cos(lat) * u
    y = e
    z = np.cos(lat) * n + np.sin(lat) * u
    a1, a2 = np.triu_indices(na, 1)
    lx, ly, lz = x[a1] - x[a2], y[a1] - y[a2], z[a1] - z[a2]
    ntime_total = ntime_total or ntime
    steps = np.arange(t0, t0 + ntime)
    ha = (steps / max(ntime_total - 1, 1) - 0.5) * hours * (np.pi / 12.0)
    sh, ch = np.sin(ha)[:, None], np.cos(ha)[:, None]
    uu = sh * lx + ch * ly
    vv = -np.sin(dec) * ch * lx + np.sin(dec) * sh * ly + np.cos(dec) * lz
    ww = np.cos(dec) * ch * lx - np.cos(dec) * sh * ly + np.sin(dec) * lz
    uvw = np.stack([uu, vv, ww], axis=-1).reshape(-1, 3)
    nbl = a1.size
    time_index = np.repeat(steps, nbl).astype(np.int32)
    ant1 = np.tile(a1, ntime).astype(np.int32)
    ant2 = np.tile(a2, ntime).astype(np.int32)
    return np.ascontiguousarray(uvw), time_index, ant1, ant2


def sky_lm(nsrc, rng, radius=0.02):
    r = radius * np.sqrt(rng.uniform(0, 1, nsrc))
    th = rng.uniform(0, 2 * np.pi, nsrc)
    return np.stack([r * np.cos(th), r * np.sin(th)], axis=1)


def frequencies(nchan, f0=0.856e9, f1=1.712e9):
    return np.linspace(f0, f1, nchan)


def stokes_image(nsrc, nchan, ncorr, rng, freq):
    """Real brightness (source, chan, corr): |N(0,1)| with a -0.7 spectral index."""
    i0 = np.abs(rng.standard_normal(nsrc))
    spec = (freq / freq[nchan // 2]) ** (-0.7)
    img = i0[:, None, None] * spec[None, :, None] * np.ones((1, 1, ncorr))
    return np.ascontiguousarray(img)


def brightness_2x2(nsrc, nchan, rng, freq):
    """Complex (source, chan, 2, 2) coherency [[I+Q, U+iV], [U-iV, I-Q]]
    (africanus/model/coherency/conversion.py:19-28)."""
    spec = (freq / freq[nchan // 2]) ** (-0.7)
    i = np.abs(rng.standard_normal(nsrc))[:, None] * spec[None, :]
    q, u, v = (0.1 * rng.standard_normal(nsrc)[:, None] * spec[None, :] for _ in range(3))
    b = np.empty((nsrc, nchan, 2, 2), np.complex128)
    b[..., 0, 0] = i + q
    b[..., 0, 1] = u + 1j * v
    b[..., 1, 0] = u - 1j * v
    b[..., 1, 1] = i - q
    return b


def gains(ntime, na, nchan, rng, scale=0.1):
    shape = (ntime, na, nchan, 2, 2)
    g = scale * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    g[..., 0, 0] += 1.0
    g[..., 1, 1] += 1.0
    return g


def beam_cube(npix=257, nud=64, rng=None, extent_deg=1.5):
    """MeqTrees-style analytic beam cos^3(min(65 nu_GHz r, 1.0881)) on the diagonal
    (africanus/testing/beam_factory.py:151-155) with small random leakage."""
    rng = rng or np.random.default_rng(0)
    ext = np.deg2rad(extent_deg)
    x = np.linspace(-ext, ext, npix)
    ll, mm = np.meshgrid(x, x, indexing="ij")
    r = np.rad2deg(np.sqrt(ll**2 + mm**2))
    bfreq = np.linspace(0.856e9, 1.712e9, nud)
    amp = np.cos(np.minimum(65.0 * (bfreq[None, None, :] * 1e-9) * r[:, :, None], 1.0881)) ** 3
    beam = np.zeros((npix, npix, nud, 2, 2), np.complex128)
    beam[..., 0, 0] = amp
    beam[..., 1, 1] = amp
    leak = 0.01 * (rng.standard_normal((npix, npix, nud, 2)) + 1j * rng.standard_normal((npix, npix, nud, 2)))
    beam[..., 0, 1] = leak[..., 0]
    beam[..., 1, 0] = leak[..., 1]
    extents = np.array([[-ext, ext], [-ext, ext]])
    return beam, extents, bfreq
