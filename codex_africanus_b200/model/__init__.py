"""Source models feeding the predict: the spectral model and the
Stokes <-> correlation conversion, mirroring ``africanus.model.spectral`` and
``africanus.model.coherency``."""
from .coherency import convert, stokes_brightness  # noqa: F401
from .spectral import spectral_model  # noqa: F401
