"""Direct Fourier transforms -- mirrors ``africanus.dft`` (africanus/dft/__init__.py:3)."""
from .kernels import im_to_vis, vis_to_im  # noqa: F401

__all__ = ["im_to_vis", "vis_to_im"]
