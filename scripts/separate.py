#!/usr/bin/env python
"""Fused wav -> wav separation (STFT -> UNet mask -> iSTFT in one process): see svs_unet_pytorch_b200/separate.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200.separate import main  # noqa: E402

if __name__ == "__main__":
    main()
