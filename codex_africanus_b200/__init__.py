"""
codex_africanus_b200 -- B200 (sm_100a) implementation of the codex-africanus
RIME predict / direct-DFT hot path behind the reference's own Python signatures.

    codex_africanus_b200.rime : phase_delay, predict_vis, apply_gains, beam_cube_dde,
                                fused_predict_vis (phase_delay (x) brightness -> predict_vis)
    codex_africanus_b200.dft  : im_to_vis, vis_to_im
    codex_africanus_b200.model: spectral_model, convert, stokes_brightness (brightness from
                                Stokes parameters, generated on the device)

Mirrors ``africanus.rime`` (africanus/rime/__init__.py:3-10) and ``africanus.dft``
(africanus/dft/__init__.py:3) as a sibling backend in the style of
``africanus.rime.cuda`` (africanus/rime/cuda/__init__.py:3-6).

Every function runs hand-written CUDA in ``libafricanus_b200.so`` through a ctypes
C-ABI shim (include/africanus_b200.h).  There is NO CPU fallback: if the library
is missing or no CUDA device is present the call raises.
"""
__version__ = "0.1.0"

from . import constants  # noqa: F401
