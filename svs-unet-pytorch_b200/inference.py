"""Drop-in for reference inference.py (same flags and file conventions): mixture ``*_spec.npy`` ->
vocal (or accompaniment) ``*_spec.npy``.  The patches of ALL songs go through the UNet in staged batches of up to
512 (the reference runs one patch at a time, inference.py:79-116) with the mask application fused in."""
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
def separate_spectrograms(model, mix_specs, vocal_solo: bool = True, max_batch: int = 512):
    """list of (513, T_i) float32 -> list of (513, T_i) float32: the per-song body of reference inference.py:65-127
    for MANY songs at once.  All their patches are staged together (svs_patches_gather: DC row dropped, last patch
    zero padded), run through the UNet in batches of up to ``max_batch`` on the TMA / tensor-core fast path and
    scattered back (crop + DC row of zeros) — the same staged, batched route pipeline.Separator uses."""
    for m in mix_specs:
        if m.shape[0] != N_BINS:
            raise _lib.SvsError(f"expected a ({N_BINS}, T) spectrogram, got {m.shape}")
    dev = next(model.parameters()).device
    frames = [int(m.shape[1]) for m in mix_specs]
    frame_off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    host = np.concatenate([np.ascontiguousarray(m.T, dtype=np.float32) for m in mix_specs], axis=0)   # [sum T][513]
    mag = torch.from_numpy(host).to(dev)
    out = torch.empty_like(mag)
    offs, valid, _ = pipeline.patch_table(frames, frame_off)
    d_off = torch.from_numpy(offs).to(dev)
    d_valid = torch.from_numpy(valid).to(dev)
    plan = model.plan()
    flags = _lib.FLAG_APPLY_MASK | (0 if vocal_solo else _lib.FLAG_INVERT)
    for a in range(0, len(offs), max_batch):
        b = min(len(offs), a + max_batch)
        x = _lib.patches_gather_raw(mag, d_off[a:b], d_valid[a:b], None)
        y = plan.forward_dense(x, flags)
        _lib.patches_scatter_raw(y, d_off[a:b], d_valid[a:b], out, dc_zero=True)
    res = out.cpu().numpy()
    return [np.ascontiguousarray(res[int(frame_off[i]):int(frame_off[i + 1])].T) for i in range(len(frames))]


def separate_spectrogram(model, mix_spec: np.ndarray, vocal_solo: bool = True, max_batch: int = 512) -> np.ndarray:
    """(513, T) float32 -> (513, T) float32 C-order (like the np.vstack of inference.py:123)."""
    return separate_spectrograms(model, [mix_spec], vocal_solo, max_batch)[0]


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
    specs, names = [], []
    for name in files:                                               # per-song isolation like inference.py:63-64
        try:
            specs.append(np.load(os.path.join(args.mixture_folder, name)))
            names.append(name)
        except Exception as e:
            print(f"Error processing {name}: {e}")
    for name, pred in zip(names, separate_spectrograms(model, specs, bool(args.vocal_solo))):
        np.save(os.path.join(args.tar, name), pred)
    print("done")


if __name__ == "__main__":
    main()
