"""Device-resident separation pipeline:  audio -> STFT -> /max -> UNet mask x mixture -> iSTFT -> 0.9 peak.

This is the three reference CLI stages (data.py to_spec -> inference.py -> data.py to_wave, joined by
.npy files on disk in the reference: inference.py:135-150) as ONE pass that keeps every spectrogram in
HBM.  The patch logic restates reference inference.py:65-127:

* the DC row is dropped on the way in (a +1 element offset) and re-inserted as zeros on the way out,
* the time axis is cut into 128-frame patches, ``T // 128 + 1`` of them with the empty one skipped,
* the last patch is zero padded (frames >= valid are read as 0 by conv1) and cropped (not written).

Patches are staged by a tiled transpose at HBM speed (svs_patches_gather / svs_patches_scatter, with the
normalisation folded into the gather); with ``staged=False`` the UNet kernels instead read / write the
frame-major spectrogram ([T][513], i.e. librosa's Fortran-ordered (513, T)) through a strided patch view.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .config import INPUT_LEN, N_BINS
from .spectral import SongBatch


def patch_table(frames_per_song, frame_off):
    """Host-side patch table (reference inference.py:75-92).  Returns (elem_off int64[P], valid int32[P],
    song int32[P]) where elem_off indexes a [total_frames][513] float32 spectrogram (DC bin skipped)."""
    offs, valid, song = [], [], []
    for s, t in enumerate(frames_per_song):
        for i in range(t // INPUT_LEN + 1):
            cur = min(INPUT_LEN, t - i * INPUT_LEN)
            if cur <= 0:
                continue                                              # inference.py:88
            offs.append((int(frame_off[s]) + i * INPUT_LEN) * N_BINS + 1)
            valid.append(cur)
            song.append(s)
    return (np.asarray(offs, dtype=np.int64), np.asarray(valid, dtype=np.int32), np.asarray(song, dtype=np.int32))


class Separator:
    """Runs whole songs through the fused path on one GPU."""

    def __init__(self, model, max_batch: int = 512, staged: bool = True):
        self.model = model
        self.max_batch = int(max_batch)
        self.staged = bool(staged)      # False: the UNet reads / writes the spectrogram through strided patch views
        self._tables: dict = {}
        self._streams: dict = {}

    def _lanes(self, dev):
        """The two streams UNet batches alternate on (SVS_SEP_STREAMS=1: the caller's stream only)."""
        import os
        if os.environ.get("SVS_SEP_STREAMS", "2") == "1":
            return [torch.cuda.current_stream(dev)]
        key = str(dev)
        if key not in self._streams:
            self._streams[key] = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        return self._streams[key]

    def _table(self, batch: SongBatch, dev):
        """Patch table of a song batch on the device; the last few batch geometries are cached (a corpus of
        equal-length songs re-uses them, and building one costs a host loop plus three small synchronous H2D copies
        while the GPU idles)."""
        key = (tuple(batch.frames), str(dev))
        val = self._tables.get(key)
        if val is None:
            offs, valid, song = patch_table(batch.frames, batch.frame_off_host)
            val = (offs, torch.from_numpy(offs).to(dev), torch.from_numpy(valid).to(dev),
                   torch.from_numpy(song).to(dev).long())
            if len(self._tables) >= 8:
                self._tables.pop(next(iter(self._tables)))
            self._tables[key] = val
        return val

    @torch.no_grad()
    def separate_batch(self, batch: SongBatch, vocal_solo: bool = True, peak_normalize: bool = True,
                       return_spec: bool = False, pcm16: bool = False):
        """SongBatch (device audio) -> (wave [total_wave] f32 device, peak [n_songs]); optionally also the
        normalised mixture spectrogram, the phase and the masked spectrogram ([F,513] layouts).  ``pcm16``: the
        waveform comes back as int16 PCM (0.9 peak normalisation + the PCM_16 quantiser of data.py:166 in one pass)."""
        plan = self.model.plan()
        mag, phase, smax = batch.stft()
        dev = mag.device
        offs, d_off, d_valid, d_song = self._table(batch, dev)
        flags = _lib.FLAG_APPLY_MASK | (0 if vocal_solo else _lib.FLAG_INVERT)
        n = len(offs)
        if self.staged:
            # gather -> dense UNet batches -> scatter: conv1 and deconv6 stay on their TMA / tensor-core paths, and
            # every frame row of out_mag is written in full.  Unless the caller wants the normalised spectrogram
            # back, the data.py:105 normalisation is folded into the gather instead of a pass of its own.
            d_norm = None
            if return_spec:
                batch.normalize(mag, smax)
            else:
                d_norm = smax[d_song]
            out_mag = torch.empty_like(mag)
            spans = [(a, min(n, a + self.max_batch)) for a in range(0, n, self.max_batch)]
            # Two UNet batches in flight on two streams (each with its own activation workspace): the layers that
            # leave SMs idle fill up from the other batch (+7..15 % at 64-patch batches, bench.py `two_in_flight`).
            cur = torch.cuda.current_stream(dev)
            lanes = self._lanes(dev) if len(spans) > 1 else [cur]
            for st in lanes:
                if st is not cur:
                    st.wait_stream(cur)
            for i, (a, b) in enumerate(spans):
                with torch.cuda.stream(lanes[i % len(lanes)]):
                    x = _lib.patches_gather_raw(mag, d_off[a:b], d_valid[a:b], None if d_norm is None else d_norm[a:b])
                    y = plan.forward_dense(x, flags)
                    _lib.patches_scatter_raw(y, d_off[a:b], d_valid[a:b], out_mag, dc_zero=True)
            for st in lanes:
                if st is not cur:
                    cur.wait_stream(st)
        else:
            batch.normalize(mag, smax)                                # data.py:105
            out_mag = torch.zeros_like(mag)                           # DC row stays 0 (inference.py:123)
            for a in range(0, n, self.max_batch):
                b = min(n, a + self.max_batch)
                iv = _lib.PatchView(mag.data_ptr(), d_off[a:b].data_ptr(), 0, 1, N_BINS)
                ov = _lib.PatchView(out_mag.data_ptr(), d_off[a:b].data_ptr(), 0, 1, N_BINS)
                plan.forward_views(iv, ov, d_valid[a:b], b - a, flags)
        wave, peak = batch.istft(out_mag, phase, peak_normalize=peak_normalize, pcm16=pcm16)   # data.py:159-166
        if return_spec:
            return wave, peak, mag, phase, out_mag
        return wave, peak

    def separate(self, songs, vocal_solo: bool = True, peak_normalize: bool = True):
        """list of 1-D float32 host arrays -> list of float32 numpy waveforms (len 768 * (T - 1))."""
        batch = SongBatch.from_audio(songs, device=next(self.model.parameters()).device)
        wave, _ = self.separate_batch(batch, vocal_solo, peak_normalize)
        host = wave.cpu().numpy()
        return [host[int(batch.wave_off_host[s]): int(batch.wave_off_host[s]) + batch.wave_lengths[s]]
                for s in range(batch.n_songs)]


class SongStreamer:
    """Host-buffer entry point for whole songs (BASELINE configs[3] end to end): pinned host audio -> device ->
    STFT -> UNet mask -> iSTFT -> pinned host waveforms.  Songs are processed in chunks on alternating streams;
    inside a stream a chunk's H2D copy, kernels and D2H copy are ordered, across the streams the copies of one chunk
    overlap the kernels and the opposite-direction copies of the others (one DMA engine per direction).  This is the data.py -> inference.py -> data.py
    chain of the reference with the four .npy files per song replaced by HBM."""

    def __init__(self, model, songs_per_chunk: int = 10, vocal_solo: bool = True, n_streams: int = 4):
        self.sep = Separator(model)
        self.songs_per_chunk = int(songs_per_chunk)
        self.vocal_solo = vocal_solo
        self.dev = next(model.parameters()).device
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(max(1, int(n_streams)))]

    @torch.no_grad()
    def run(self, host_audio: torch.Tensor, lengths, host_wave: torch.Tensor):
        """host_audio: pinned float32 — or int16 PCM — [sum(lengths)] (songs back to back); host_wave: pinned float32
        — or int16 PCM — [sum(768 * (len // 768))] receiving the separated songs back to back.  int16 on either side
        is the reference's real file boundary (PCM_16 .wav in, data.py:78; PCM_16 .wav out, data.py:166) and halves
        the PCIe bytes of that direction.  Returns per-song wave lengths."""
        n = len(lengths)
        pcm_out = host_wave.dtype == torch.int16
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            s.wait_stream(cur)
        in_off, out_off, wave_lengths = 0, 0, []
        keep = []                                                     # device tensors stay referenced until the join
        for ci, a in enumerate(range(0, n, self.songs_per_chunk)):
            lens = [int(v) for v in lengths[a:a + self.songs_per_chunk]]
            n_in = sum(lens)
            st = self.streams[ci % len(self.streams)]
            with torch.cuda.stream(st):
                audio = host_audio[in_off:in_off + n_in].to(self.dev, non_blocking=True)
                batch = SongBatch(audio, lens)
                wave, _ = self.sep.separate_batch(batch, self.vocal_solo, True, pcm16=pcm_out)
                host_wave[out_off:out_off + batch.total_wave].copy_(wave, non_blocking=True)
                keep.append((audio, wave))
            wave_lengths += batch.wave_lengths
            in_off += n_in
            out_off += batch.total_wave
        for s in self.streams:
            cur.wait_stream(s)
        return wave_lengths


class PatchStreamer:
    """Host-buffer entry point for patch batches: pinned host -> device -> UNet -> pinned host, with the
    H2D copy of batch i+1 and the D2H copy of batch i-1 overlapping the kernels of batch i (three
    streams, n_buf-deep device staging).  Mirrors what reference inference.py:97-110 does one
    patch at a time with a synchronous .to(device) / .cpu() pair.

    The loop itself is native (`svs_patch_stream_run`, csrc/stream_host.cu): as ~20 torch calls per step it cost
    0.30 ms of interpreter time against 0.34 ms of PCIe time per 64-patch step, so the host CPU, not the link, set the
    throughput on slower boxes."""

    def __init__(self, model, batch: int = 64, vocal_solo: bool | None = None, n_buf: int = 4):
        self.model = model
        self.batch = batch
        self.flags = 0 if vocal_solo is None else (_lib.FLAG_APPLY_MASK | (0 if vocal_solo else _lib.FLAG_INVERT))
        dev = next(model.parameters()).device
        self.dev = dev
        self.n_buf = max(2, int(n_buf))                                 # staging depth: 2 leaves the copy engines idle
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))   # whenever a step jitters
        shape = (batch, 1, 512, 128)
        self.d_in = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(self.n_buf)]
        self.d_out = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(self.n_buf)]
        self.plan = model.plan()
        with torch.cuda.stream(self.s_cmp):
            self.ws = self.plan.workspace(batch)                        # owned by the compute stream
        self._ptr_in = (ctypes.c_void_p * self.n_buf)(*[t.data_ptr() for t in self.d_in])
        self._ptr_out = (ctypes.c_void_p * self.n_buf)(*[t.data_ptr() for t in self.d_out])
        handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(_lib.load().svs_patch_stream_create(self.n_buf, ctypes.byref(handle)), "svs_patch_stream_create")
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().svs_patch_stream_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @torch.no_grad()
    def run(self, host_in, host_out):
        """host_in / host_out: sequences of pinned float32 tensors (batch,1,512,128).  Returns after the
        last result has been ENQUEUED for download; the caller's current stream waits for it."""
        n = len(host_in)
        if n != len(host_out):
            raise _lib.SvsError("PatchStreamer.run: host_in and host_out differ in length")
        shape = (self.batch, 1, 512, 128)
        for t in list(host_in) + list(host_out):
            if t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous():
                raise _lib.SvsError(f"PatchStreamer.run: host buffers must be contiguous float32 {shape} CPU tensors")
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        p_in = (ctypes.c_void_p * n)(*[t.data_ptr() for t in host_in])
        p_out = (ctypes.c_void_p * n)(*[t.data_ptr() for t in host_out])
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().svs_patch_stream_run(
                self.handle, self.plan.handle, p_in, p_out, n, self.batch, self.flags, self._ptr_in, self._ptr_out,
                self.ws.data_ptr(), self.ws.numel(), self.s_in.cuda_stream, self.s_cmp.cuda_stream,
                self.s_out.cuda_stream), "svs_patch_stream_run")
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_in)


class CopyCeiling:
    """The PCIe ceiling of PatchStreamer / SongStreamer on this box: the same pinned-host <-> device copies with no
    kernels in between, H2D and D2H on two streams (one DMA engine per direction).  bench.py runs it on every rank
    at once, so the figure includes whatever the ranks share (root complex, host memory controllers)."""

    def __init__(self, device, nbytes: int):
        self.dev = device
        self.buf_in = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf_out = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.s_in, self.s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def run(self, host_in: torch.Tensor, host_out: torch.Tensor, steps: int):
        hin = host_in.view(-1).view(torch.uint8)[: self.buf_in.numel()]
        hout = host_out.view(-1).view(torch.uint8)[: self.buf_out.numel()]
        cur = torch.cuda.current_stream(self.dev)
        self.s_in.wait_stream(cur)
        self.s_out.wait_stream(cur)
        for _ in range(steps):
            with torch.cuda.stream(self.s_in):
                self.buf_in.copy_(hin, non_blocking=True)
            with torch.cuda.stream(self.s_out):
                hout.copy_(self.buf_out, non_blocking=True)
        cur.wait_stream(self.s_in)
        cur.wait_stream(self.s_out)


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pins this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers
    are allocated, so that they are first-touched on that node: with eight ranks on one socket's memory every copy
    of the other socket's four GPUs crosses the inter-socket link.  Best effort; returns what it did."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        index = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if ids and all(v.isdigit() for v in ids) and local_rank < len(ids):
            index = int(ids[local_rank])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                               # nvml: 00000000:1b:00.0 -> sysfs 0000:1b:00.0
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        orig = os.sched_getaffinity(0)
        cpus &= orig
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus), original_affinity=sorted(orig))
    except Exception as e:                                            # no nvml / sysfs: run unbound
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info
