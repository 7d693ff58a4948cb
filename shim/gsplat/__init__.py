"""Drop-in shim: makes ``from gsplat.rendering import rasterization`` (the reference's import at
/root/reference/model/gaussian.py:8) resolve to the B200-native rasterizer.  Put this directory's parent
(``<repo>/shim``) on PYTHONPATH *instead of* installing gsplat; the reference then runs unchanged."""
from .rendering import rasterization  # noqa: F401

__version__ = "1.0.0+egs.b200"
