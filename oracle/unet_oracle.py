"""CPU oracle: the reference UNet forward / training loss restated on torch-CPU
(TEST INFRASTRUCTURE ONLY — never imported by the product package).

PINNED: ``tests/golden/make_golden.py`` imports the reference's own ``model.py`` from
/root/reference (stubbing the unused ``import auraloss`` at model.py:6) and stores its outputs
in ``tests/golden/unet_golden.npz``; ``tests/test_oracle_unet.py`` checks this restatement
against those vectors, so the oracle can travel to the GPU box where /root/reference is absent.

The forward is written functionally over a plain ``state_dict`` (the 79-key layout of
reference model.py:47-109) so it is independent of the product's ``UNet`` module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default, reference model.py:49
BN_MOMENTUM = 0.1
LEAKY = 0.2          # reference model.py:50

ENC = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6"]
DEC = ["deconv1", "deconv2", "deconv3", "deconv4", "deconv5", "deconv6"]


def _bn(x, sd, prefix, training, batch_stats=None):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        return F.batch_norm(x, None, None, w, b, True, BN_MOMENTUM, BN_EPS)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b,
                        False, BN_MOMENTUM, BN_EPS)


def unet_forward(sd: dict, mix: torch.Tensor, training: bool = False,
                 dropout_masks: dict | None = None, return_activations: bool = False):
    """``UNet.forward`` of reference model.py:169-201.

    ``training=False`` -> eval semantics (running-stat BN, Dropout2d identity).
    ``training=True``  -> batch-stat BN; Dropout2d(0.5) applied only through the explicit
    per-(sample, channel) keep masks in ``dropout_masks`` (``{"deconv1": bool (B, C)}``) so runs
    are reproducible (torch's Philox stream cannot be matched by another implementation).
    """
    acts = {}
    x = mix
    skips = []
    for name in ENC:                                           # model.py:176-181
        x = F.conv2d(x, sd[f"{name}.0.weight"], sd[f"{name}.0.bias"], stride=2, padding=2)
        x = _bn(x, sd, f"{name}.1", training)
        x = F.leaky_relu(x, LEAKY)
        skips.append(x)
        acts[name] = x
    sizes = [mix] + skips                                      # output_size targets
    x = skips[5]
    for i, name in enumerate(DEC):                             # model.py:183-198
        target = sizes[5 - i]
        # ConvTranspose2d(k=5, s=2, p=2) with output_size = 2*in  => output_padding = 1
        x = F.conv_transpose2d(x, sd[f"{name}.weight"], sd[f"{name}.bias"], stride=2, padding=2,
                               output_padding=1)
        assert x.shape[-2:] == target.shape[-2:], (x.shape, target.shape)
        if name != "deconv6":
            x = _bn(x, sd, f"{name}_BAD.0", training)
            x = F.relu(x)
            if training and dropout_masks is not None and name in dropout_masks:
                keep = dropout_masks[name].to(x.dtype)         # Dropout2d: whole channels, scale 1/(1-p)
                x = x * keep[:, :, None, None] * 2.0
            acts[name] = x
            x = torch.cat([x, skips[4 - i]], 1)
    out = torch.sigmoid(x)                                     # model.py:200
    acts["mask"] = out
    return (out, acts) if return_activations else out


def l1_masked_loss(mask, mix, voc, two_term: bool = True):
    """Loss of reference train.py:274-283 with ``crit = nn.L1Loss()`` (config.py:33,44; the shipped
    WeightedL1Loss of model.py:15-40 raises at call time).  ``two_term=False`` = vocal term only."""
    pred_vocal = mask * mix
    loss = (pred_vocal - voc).abs().mean()
    if two_term:
        pred_accomp = (1 - mask) * mix
        target_accomp = torch.clamp(mix - voc, min=0.0)
        loss = loss + (pred_accomp - target_accomp).abs().mean()
    return loss


def separate_spectrogram(sd: dict, mix_spec, vocal_solo: bool = True, dtype=torch.float32):
    """The per-song body of reference inference.py:65-127 (B=1 loop, DC row dropped and re-inserted).

    ``mix_spec`` numpy float32 ``(513, T)`` -> numpy float32 ``(513, T)``."""
    import numpy as np
    seg_len = 128                                              # config.py:50 INPUT_LEN
    mix_crop = np.asarray(mix_spec)[1:, :]
    num_segments = mix_crop.shape[-1] // seg_len + 1           # inference.py:75
    out = []
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    with torch.no_grad():
        for i in range(num_segments):
            seg = mix_crop[:, i * seg_len:(i + 1) * seg_len]
            cur = seg.shape[1]
            if cur == 0:
                continue                                       # inference.py:88
            if cur < seg_len:
                seg = np.pad(seg, ((0, 0), (0, seg_len - cur)), mode="constant")
            x = torch.from_numpy(np.ascontiguousarray(seg[None, None])).to(dtype)
            msk = unet_forward(sd, x)
            if not vocal_solo:
                msk = 1 - msk                                  # inference.py:102
            pred = (x * msk).squeeze().float().numpy()
            out.append(pred[:, :cur])
    full = np.concatenate(out, axis=1)
    return np.vstack((np.zeros((1, full.shape[1]), dtype=np.float32), full))   # inference.py:123
