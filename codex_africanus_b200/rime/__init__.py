"""RIME terms -- mirrors the hot-path subset of ``africanus.rime``
(africanus/rime/__init__.py:3-10) plus the fused predict."""
from .phase import phase_delay  # noqa: F401
from .predict import apply_gains, predict_vis  # noqa: F401
from .fast_beam_cubes import beam_cube_dde, beam_cube_dde_rotated, freq_grid_interp  # noqa: F401
from .feeds import feed_rotation  # noqa: F401
from .fused import fused_predict_vis  # noqa: F401
from .fused_beam import fused_predict_vis_beam  # noqa: F401
from .fused_stokes import fused_predict_vis_stokes  # noqa: F401
from .fused_spec import rime as rime_from_spec  # noqa: F401  (africanus.experimental.rime.fused.core.rime)
from .stream import (stream_fused_predict_vis, stream_fused_predict_vis_beam,  # noqa: F401
                     stream_predict_vis_stokes, timestep_row_blocks)
from .wsclean_predict import spectra as wsclean_spectra, wsclean_predict  # noqa: F401

__all__ = ["phase_delay", "predict_vis", "apply_gains", "beam_cube_dde", "beam_cube_dde_rotated", "feed_rotation", "freq_grid_interp",
           "fused_predict_vis", "fused_predict_vis_beam", "fused_predict_vis_stokes", "stream_fused_predict_vis", "stream_fused_predict_vis_beam", "stream_predict_vis_stokes", "timestep_row_blocks", "wsclean_predict", "wsclean_spectra"]
