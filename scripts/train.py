#!/usr/bin/env python
"""Launcher with the reference's file name: `python scripts/train.py ...` == reference `python train.py ...`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200.train import main  # noqa: E402

if __name__ == "__main__":
    main()
