"""latentfusion_b200 — a B200-native (sm_100a) implementation of LatentFusion's
reconstruct->render hot path behind the reference's own Python API.

Layout mirrors the reference package for the modules on the path:
``modules/`` (geometry, blocks, unet, gru, lstm), ``recon/`` (models, fusion, inference),
``pose/`` (estimation, utils), ``three/`` and ``observation``.  The arithmetic lives in
``csrc/*.cu`` behind the C ABI of ``include/lfb200.h`` (``liblfb200.so``).
"""
from . import ops  # noqa: F401
from .ops import (PRECISION_FP32, PRECISION_BF16X3, PRECISION_BF16, set_default_precision,  # noqa: F401
                  get_default_precision)

__version__ = '0.1.0'
