"""pislam_b200 -- B200 (sm_100a) ORB feature-extraction front-end behind the pislam:: API.

  include/pislam_b200.h      the C ABI (the drop-in boundary)
  include/pislam/*.h         C++ template wrappers with the reference's names and signatures
  pislam_b200.api            ctypes mirror of that API for Python callers and the parity tests
  pislam_b200.pyramid        stacked-pyramid geometry + synthetic inputs (host-side numpy)
"""
from .pyramid import FHD, UHD, VGA, PyramidLayout, pyramid_layout  # noqa: F401

__all__ = ["api", "pyramid", "VGA", "FHD", "UHD", "PyramidLayout", "pyramid_layout"]
