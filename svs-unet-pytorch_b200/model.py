"""Drop-in `model.py`: the reference's `UNet` module surface on top of the sm_100a kernels.

Same sub-module names / types / creation order as reference model.py:47-109, so the 79-key fp32
``state_dict`` (and therefore every ``.pth`` the reference's train.py / UNet.save writes) loads with
``strict=True`` and a given ``torch.manual_seed`` produces identical random-init weights.  The
arithmetic of ``forward`` does not run in torch: it is one call into libsvs_b200.so
(``svs_b200::unet_forward`` custom op -> ``svs_unet_forward``).  There is no CPU path.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib

_ENC = [(1, 16), (16, 32), (32, 64), (64, 128), (128, 256), (256, 512)]
_DEC = [(512, 256), (512, 128), (256, 64), (128, 32), (64, 16), (32, 1)]


class MaskedL1Loss(nn.Module):
    """The loss the reference trains with.

    * ``crit(pred, target)``       -> mean |pred - target|  (the ``nn.L1Loss`` form recorded in reference
      config.py:33,44 and called two-argument style at train.py:281-282)
    * ``crit(voc, mix, mask)``     -> the two-term masked L1 of train.py:275-283 (the call shape of
      ``UNet.backward``, reference model.py:213)

    The class shipped in reference model.py:15-40 (``WeightedL1Loss``) raises on both call shapes."""

    def __init__(self, reduction: str = "mean"):
        super().__init__()
        self.reduction = reduction

    def _reduce(self, x):
        if self.reduction == "mean":
            return x.mean()
        if self.reduction == "sum":
            return x.sum()
        return x

    def forward(self, a, b, mask=None):
        if mask is None:
            return self._reduce((a - b).abs())
        target_vocal, target_mix = a, b
        pred_vocal = mask * target_mix
        pred_accomp = (1 - mask) * target_mix
        target_accomp = torch.clamp(target_mix - target_vocal, min=0.0)
        return self._reduce((pred_vocal - target_vocal).abs()) + self._reduce((pred_accomp - target_accomp).abs())


WeightedL1Loss = MaskedL1Loss   # name kept importable for code written against reference model.py:15


def _encoder_block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=(5, 5), stride=(2, 2), padding=2),
                         nn.BatchNorm2d(cout), nn.LeakyReLU(negative_slope=0.2, inplace=True))


class UNet(nn.Module):
    """reference model.py:42-220.  ``precision``: "bf16" (default, tcgen05 kind::f16), "tf32"
    (tcgen05 kind::tf32 on fp32 activations) or "fp32" (CUDA-core exact mode); override with the
    ``SVS_B200_PRECISION`` environment variable."""

    def __init__(self, precision: str | None = None):
        super().__init__()
        for i, (cin, cout) in enumerate(_ENC, start=1):
            setattr(self, f"conv{i}", _encoder_block(cin, cout))
        for i, (cin, cout) in enumerate(_DEC, start=1):
            setattr(self, f"deconv{i}", nn.ConvTranspose2d(cin, cout, kernel_size=(5, 5), stride=(2, 2), padding=2))
            if i < 6:
                setattr(self, f"deconv{i}_BAD", nn.Sequential(nn.BatchNorm2d(cout), nn.ReLU(True), nn.Dropout2d(0.5)))
        self.loss_list_vocal = []
        self.loss_list_accomp = []
        self.loss_list_total = []
        self.optim = torch.optim.Adam(self.parameters(), lr=1e-3)      # reference model.py:116
        self.crit = MaskedL1Loss()
        self.precision = precision or os.environ.get("SVS_B200_PRECISION", "bf16")
        self._plan = None
        self._plan_id = None
        self._plan_key = None

    # ------------------------------------------------------------------ checkpoint IO (model.py:122-152)
    def load(self, path):
        if os.path.exists(path):
            print("Load the pre-trained model from {}".format(path))
            state = torch.load(path, map_location="cpu")
            for key, obj in state.items():
                if "loss_list" in key:
                    setattr(self, key, obj)
            self.load_state_dict(state["model_state_dict"], strict=False)
            if "optim" in state:
                self.optim.load_state_dict(state["optim"])
        else:
            print("Pre-trained model {} is not exist...".format(path))

    def save(self, path):
        state = {"model_state_dict": self.state_dict(), "optim": self.optim.state_dict()}
        for key in self.__dict__:
            if "loss_list" in key:
                state[key] = getattr(self, key)
        torch.save(state, path)

    def getLoss(self, normalize=False):
        loss_dict = {}
        for key in self.__dict__:
            if "loss_list" in key:
                val = getattr(self, key)
                if len(val) > 0:
                    loss_dict[key] = np.mean(val) if normalize else round(val[-1], 6)
        return loss_dict

    # ------------------------------------------------------------------ plan management
    def _tensors(self):
        return list(self.parameters()) + list(self.buffers())

    def _current_key(self):
        return (self.precision,) + tuple((t.data_ptr(), t._version) for t in self._tensors())

    def plan(self) -> "_lib.UNetPlan":
        """The inference plan (BatchNorm folded, weights repacked); rebuilt when any parameter or
        buffer changed in place, was re-assigned, or the precision changed."""
        key = self._current_key()
        if self._plan is None or key != self._plan_key:
            if self._plan_id is not None:
                _lib.release_plan(self._plan_id)
            self._plan = _lib.UNetPlan(self.state_dict(), self.precision)
            self._plan_id = _lib.register_plan(self._plan)
            self._plan_key = key
        return self._plan

    # ------------------------------------------------------------------ forward (model.py:169-201)
    def forward(self, mix):
        """mix (B,1,512,128) float32 on CUDA -> soft mask (B,1,512,128) float32.

        In ``train()`` mode the mask carries an autograd edge to every parameter; the activations the backward
        needs are kept in one per-model workspace, so only the MOST RECENT train-mode forward can be
        back-propagated (a stale ``backward()`` raises; the reference's plain autograd has no such limit)."""
        _lib.require_cuda(mix, "mix")
        if self.training:
            from . import training
            return training.train_forward(self, mix)
        self.plan()
        return _lib.unet_forward_op(mix.float(), self._plan_id, 0)

    def separate(self, mix, vocal_solo: bool = True):
        """mask application of reference inference.py:102,107 fused into the last layer:
        returns ``mix * mask`` (or ``mix * (1 - mask)`` when ``vocal_solo`` is false)."""
        _lib.require_cuda(mix, "mix")
        self.plan()
        flags = _lib.FLAG_APPLY_MASK | (0 if vocal_solo else _lib.FLAG_INVERT)
        return _lib.unet_forward_op(mix.float(), self._plan_id, flags)

    def backward(self, mix, voc):
        """reference model.py:203-220: one optimisation step on (mix, voc)."""
        self.optim.zero_grad()
        msk = self.forward(mix)
        loss = self.crit(voc, mix, msk)
        self.loss_list_total.append(loss.item())
        loss.backward()
        self.optim.step()
