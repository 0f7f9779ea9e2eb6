"""svs-unet-pytorch_b200 — B200 (sm_100a) implementation of the SVS-UNet separation hot path.

Host-side mirror of the reference's Python surface (config / model / data / inference / train) over
the C-ABI library ``libsvs_b200.so`` (``include/svs_b200.h``)."""
from . import config  # noqa: F401

__all__ = ["config"]
