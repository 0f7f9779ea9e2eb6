#!/usr/bin/env python
"""Launcher with the reference's file name: `python scripts/data.py ...` == reference `python data.py ...`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svs_unet_pytorch_b200.data import main  # noqa: E402

if __name__ == "__main__":
    main()
