"""Minimal WAV reader / PCM_16 writer (host I/O around the hot path; reference data.py:78,94,166 use
librosa.load / soundfile.write, neither of which is installable here).  Decoding and the PCM_16 write are exact.
Files whose rate differs from the target are down-mixed and resampled ON THE GPU (resample.py / svs_resample_poly: the
polyphase Kaiser filter of scipy.signal.resample_poly); the reference's soxr_hq filter cannot be reproduced here, so
resampled samples differ from the reference's the way any two high-quality resamplers differ (a warning says so)."""
from __future__ import annotations

import struct
import warnings

import numpy as np


def read_wav(path: str):
    """-> (float32 array (n,) or (n, channels), sample_rate).  PCM 8/16/24/32-bit and IEEE float32/64."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:                      # WAVE_FORMAT_EXTENSIBLE
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw[: len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v >= 1 << 23, v - (1 << 24), v)
            x = v.astype(np.float32) / float(1 << 23)
        elif bits == 32:
            x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / float(1 << 31)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:
        x = np.frombuffer(raw, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAV format tag {tag}")
    if ch > 1:
        x = x[: len(x) // ch * ch].reshape(-1, ch)
    return x, sr


def load(path: str, sr: int, mono: bool = True) -> np.ndarray:
    """Shape of ``librosa.load(path, sr=sr, mono=True)[0]``: float32 mono at ``sr``."""
    x, file_sr = read_wav(path)
    if file_sr != sr:
        if not mono and x.ndim == 2:
            raise ValueError("resampling keeps mono only (reference data.py:78 loads mono)")
        from . import resample
        warnings.warn(f"{path}: resampling {file_sr} -> {sr} Hz with the GPU polyphase Kaiser filter "
                      "(scipy.signal.resample_poly's design); the reference uses soxr_hq, so samples differ slightly",
                      stacklevel=2)
        return resample.resample(x, file_sr, sr)                        # channel mean + resampling in one kernel
    if mono and x.ndim == 2:
        x = x.mean(axis=1).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def write_wav_pcm16(path: str, y: np.ndarray, sr: int):
    """soundfile.write(path, y, sr) default for .wav: PCM_16 (reference data.py:166)."""
    y = np.asarray(y)
    if y.dtype == np.int16:                                          # already quantised on the device
        q = y.astype("<i2")
    else:                                                            # libsndfile f2les_array: lrintf(x * 0x7FFF)
        q = np.clip(np.rint(y.astype(np.float64) * 32767.0), -32768, 32767).astype("<i2")
    body = q.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(body)) + body)
