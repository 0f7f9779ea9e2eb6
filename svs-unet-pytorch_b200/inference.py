"""Drop-in for reference inference.py (same flags and file conventions): mixture ``*_spec.npy`` ->
vocal (or accompaniment) ``*_spec.npy``.  All patches of a song go through the UNet in batches (the
reference runs one patch at a time, inference.py:79-116) with the mask application fused in."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import _lib, pipeline
from .config import N_BINS
from .model import UNet


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--model_path", type=str, required=True)
    p.add_argument("--tar", type=str, required=True)
    p.add_argument("--mixture_folder", type=str, required=True)
    p.add_argument("--vocal_solo", type=int, default=1)
    return p


@torch.no_grad()
def separate_spectrogram(model, mix_spec: np.ndarray, vocal_solo: bool = True, max_batch: int = 64) -> np.ndarray:
    """(513, T) float32 -> (513, T) float32, the per-song body of reference inference.py:65-127."""
    if mix_spec.shape[0] != N_BINS:
        raise _lib.SvsError(f"expected a ({N_BINS}, T) spectrogram, got {mix_spec.shape}")
    dev = next(model.parameters()).device
    t = mix_spec.shape[1]
    mag = torch.from_numpy(np.ascontiguousarray(mix_spec.T, dtype=np.float32)).to(dev)       # [T][513]
    out = torch.zeros_like(mag)
    offs, valid, _ = pipeline.patch_table([t], np.array([0, t]))
    d_off = torch.from_numpy(offs).to(dev)
    d_valid = torch.from_numpy(valid).to(dev)
    plan = model.plan()
    flags = _lib.FLAG_APPLY_MASK | (0 if vocal_solo else _lib.FLAG_INVERT)
    for a in range(0, len(offs), max_batch):
        b = min(len(offs), a + max_batch)
        iv = _lib.PatchView(mag.data_ptr(), d_off[a:b].data_ptr(), 0, 1, N_BINS)
        ov = _lib.PatchView(out.data_ptr(), d_off[a:b].data_ptr(), 0, 1, N_BINS)
        plan.forward_views(iv, ov, d_valid[a:b], b - a, flags)
    return np.ascontiguousarray(out.cpu().numpy().T)                  # (513, T) C-order like np.vstack


def main(argv=None):
    args = build_parser().parse_args(argv)
    os.makedirs(args.tar, exist_ok=True)
    if not torch.cuda.is_available():
        raise _lib.SvsError("inference needs a B200: svs-unet-pytorch_b200 has no CPU path")
    device = torch.device("cuda")
    print(f"Inference using device: {device}")
    model = UNet().to(device)
    try:
        ckpt = torch.load(args.model_path, map_location=device)
        if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
            model.load_state_dict(ckpt["model_state_dict"])
    except Exception as e:
        print(f"failed to load the model: {e}")
        raise SystemExit(1)
    model.eval()
    files = sorted(f for f in os.listdir(args.mixture_folder) if f.endswith("_spec.npy"))[:20]   # inference.py:59
    print(f"found {len(files)} files, separating...")
    for name in files:
        mix = np.load(os.path.join(args.mixture_folder, name))
        np.save(os.path.join(args.tar, name), separate_spectrogram(model, mix, bool(args.vocal_solo)))
    print("done")


if __name__ == "__main__":
    main()
