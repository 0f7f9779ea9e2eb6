"""The waveform-domain half of the reference's training objective.

reference train.py:24-26,33-60,287-296:  total = 166.66 * L1 + 0.66 * MRSTFT(istft(mask * mix, phase_mix), istft(voc, phase_voc))
with ``auraloss.freq.MultiResolutionSTFTLoss(sample_rate=8192)`` (auraloss==0.4.0, uv.lock) at its defaults.  auraloss is
not installable here, so ``MultiResolutionSTFTLoss`` below RESTATES its published algorithm (spectral convergence +
log-magnitude L1 over three STFT resolutions) — parity unpinned at that boundary, like librosa's.

This term is NOT on the hand-written hot path: it needs gradients through an inverse STFT and three forward STFTs of
other sizes, and runs on torch's differentiable ``torch.istft`` / ``torch.stft`` (cuFFT) exactly as the reference
does.  The UNet under it is still the sm_100a training path (``UNet.forward`` in train mode is an autograd Function
over svs_unet_train_forward / backward)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .config import HOP_SIZE, WINDOW_SIZE


def specific_istft(magnitude: torch.Tensor, phase: torch.Tensor, window: torch.Tensor | None = None) -> torch.Tensor:
    """reference train.py:33-60: (B,1,512,T) magnitude + phase angle -> (B,1,hop*(T-1)) waveform; the DC row dropped
    by the dataset is re-inserted as zeros."""
    magnitude = F.pad(magnitude, (0, 0, 1, 0), "constant", 0)
    phase = F.pad(phase, (0, 0, 1, 0), "constant", 0)
    spec = torch.polar(magnitude, phase).squeeze(1)
    if window is None:
        window = torch.hann_window(WINDOW_SIZE, device=magnitude.device)
    wav = torch.istft(spec, n_fft=WINDOW_SIZE, hop_length=HOP_SIZE, win_length=WINDOW_SIZE, window=window,
                      return_complex=False)
    return wav.unsqueeze(1)


class STFTLoss(torch.nn.Module):
    """auraloss 0.4.0 ``STFTLoss`` at the settings MultiResolutionSTFTLoss uses: w_sc = w_log_mag = 1, w_lin_mag =
    w_phs = 0, L1 magnitude distance, mean reduction, eps 1e-8, Hann window."""

    def __init__(self, fft_size, hop_size, win_length, eps=1e-8):
        super().__init__()
        self.fft_size, self.hop_size, self.win_length, self.eps = fft_size, hop_size, win_length, eps
        self.register_buffer("window", torch.hann_window(win_length), persistent=False)

    def _mag(self, x):
        s = torch.stft(x, self.fft_size, self.hop_size, self.win_length, self.window.to(x.device), return_complex=True)
        return torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=self.eps))

    def forward(self, x, y):
        bs, chs, seq = x.shape
        x_mag, y_mag = self._mag(x.reshape(-1, seq)), self._mag(y.reshape(-1, seq))
        sc = torch.norm(y_mag - x_mag, p="fro") / torch.norm(y_mag, p="fro")          # SpectralConvergenceLoss
        log = F.l1_loss(torch.log(x_mag), torch.log(y_mag))                           # STFTMagnitudeLoss(log=True)
        return sc + log


class MultiResolutionSTFTLoss(torch.nn.Module):
    """auraloss 0.4.0 ``MultiResolutionSTFTLoss`` defaults: fft 1024 / 2048 / 512, hop 120 / 240 / 50, window
    600 / 1200 / 240; the mean of the three STFTLoss values.  (``sample_rate`` only matters for the perceptual
    weighting / mel options, which the reference leaves off.)"""

    def __init__(self, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50), win_lengths=(600, 1200, 240),
                 sample_rate=None, device=None):
        super().__init__()
        self.losses = torch.nn.ModuleList(STFTLoss(f, h, w) for f, h, w in zip(fft_sizes, hop_sizes, win_lengths))

    def forward(self, x, y):
        return sum(loss(x, y) for loss in self.losses) / len(self.losses)
