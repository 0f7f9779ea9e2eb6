#!/usr/bin/env python
"""bench.py — headline benchmark of the SVS-UNet separation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|tf32|fp32]

Workload (BASELINE.json configs[1]): UNet mask inference on a batch of 64 synthetic 512x128 magnitude
patches (torch.rand, the range of spec/norm), random-init weights (torch.manual_seed(0)).  One "step" =
one pass of the hot path (UNet.forward + mask application) over one 64-patch batch per GPU.
Multi-GPU: patches shard by batch across ranks with NO data-path collective ("scaling": "weak").

Every timed loop runs the requested K steps, repeated back to back until the region lasts >= 0.5 s
(``timed_steps`` = how many steps that was); all figures are per-step means over the whole region,
CUDA-event timed on the launching stream, max over ranks.

* ``value``        patches/s with inputs resident in HBM (CUDA-graph replay of the forward); inputs rotate over a
                   pool larger than L2.
* ``e2e``          the same metric through the public host API (pipeline.PatchStreamer): pinned host buffers,
                   H2D + kernels + D2H all inside the timed region; ``pcie`` holds the copy-only ceiling of the
                   same transfers measured in the same run (all ranks copying at once).
* ``roofline``     whole UNet forward (12 launches, all this repo's kernels): exact valid-tap FLOPs / step time
                   against the measured BURST bf16 peak; ``roofline_tc_family`` = the tcgen05 implicit-GEMM
                   layers alone; ``roofline_stft`` / ``roofline_istft`` = 9,228 B x frames / kernel time against
                   the measured HBM copy bandwidth, on the 150-song corpus.
* ``tf32``         the precision-matched path (reference arithmetic is fp32 / TF32 under cuDNN): graph-timed value,
                   e2e and roofline against a TF32 matmul peak measured here the way MEASURED_PEAKS does bf16.
* ``two_in_flight``   informational: the same forward with TWO 64-patch batches in flight on two streams
* ``cudnn_baseline``  informational: the same nn.Module layer list through torch eager + cuDNN on this B200
                   (fp32/TF32 NCHW as reference inference.py:40 would run it, and bf16 channels_last as its best
                   case) — never on the product path.
* ``train``        BASELINE configs[4]: one optimisation step at batch 64 per GPU (forward, masked L1, backward,
                   NCCL gradient all-reduce when N > 1, Adam).
* ``cpu_baseline`` / ``--impl reference``  the reference's CPU path (oracle restatement of model.py on torch-CPU,
                   all host threads): batch-64 best case and the B=1 loop reference inference.py:79-116 really runs.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
METRIC = "patches_per_sec"
UNIT = "patches/s"
MIN_REGION_S = 0.5
AUDIO_S_PER_PATCH = 128 * 768 / 8192.0           # 12.0 s of 8192 Hz audio per full patch
GFLOP_EXACT_PER_PATCH = 1.3247                   # 2 x 662,350,768 valid-tap MACs
GFLOP_TRAIN_PER_PATCH = 3.961                    # fwd + dgrad + wgrad, no dgrad for conv1
BYTES_PER_FRAME = 9228                           # STFT: 3,072 read + 6,156 written; iSTFT the reverse
# exact (valid-tap) MACs per patch per layer, SURVEY.md section 8(a): conv1..conv6, deconv1..deconv6
LAYER_MMAC = [6.477, 51.205, 49.990, 47.587, 42.893, 33.948, 33.948, 85.787, 95.175, 99.979, 102.409, 12.954]
LAYER_NAMES = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "deconv1", "deconv2", "deconv3", "deconv4",
               "deconv5", "deconv6"]
WORKLOAD = "unet_mask_inference_b64_512x128 (BASELINE configs[1])"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
                "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0,
            "source": "fallback of B200_PROFILING.md"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            try:
                pw.append(float(parts[2]))
            except ValueError:
                pass
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------
# CPU reference arm (the oracle restatement of reference model.py; torch-CPU, all host threads)

def _cpu_setup(batch):
    import torch
    from oracle import unet_oracle
    from svs_unet_pytorch_b200 import model as svs_model
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = svs_model.UNet().eval().state_dict()
    g = torch.Generator().manual_seed(0)
    return torch, unet_oracle, sd, torch.rand(batch, 1, 512, 128, generator=g), threads


def cpu_reference_run(steps: int, warmup: int, batch: int = BATCH):
    """Batch-64 best case: returns (patches_per_s, seconds_per_step, threads)."""
    torch, unet_oracle, sd, x, threads = _cpu_setup(batch)
    with torch.no_grad():
        for _ in range(warmup):
            unet_oracle.unet_forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            m = unet_oracle.unet_forward(sd, x)
            _ = x * m                                                  # inference.py:107
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads


def cpu_reference_b1_loop(n_patches: int = 16, warmup: int = 2):
    """What reference inference.py:79-116 really does: one patch per forward.  Returns patches/s."""
    torch, unet_oracle, sd, x, threads = _cpu_setup(n_patches)
    with torch.no_grad():
        for i in range(warmup):
            unet_oracle.unet_forward(sd, x[i:i + 1])
        t0 = time.perf_counter()
        for i in range(n_patches):
            seg = x[i:i + 1]
            _ = seg * unet_oracle.unet_forward(sd, seg)
        dt = time.perf_counter() - t0
    return n_patches / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 12))                               # bounded sample: ~1-3 s per step
    warm = max(1, min(args.warmup, 2))
    pps, sec, threads = cpu_reference_run(steps, warm)
    pps_b1 = cpu_reference_b1_loop()
    sample = f"{steps} steps x {BATCH} patches after {warm} warm-up (batch-64 best case of reference model.py on torch-CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic (torch.rand patches, random-init weights seed 0)",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "patch": "512x128"},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "b1_loop_value": pps_b1,
                         "b1_loop_sample": "16 patches, one forward per patch as reference inference.py:79-116"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "audio_sec_per_sec": pps * AUDIO_S_PER_PATCH,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def torch_eager_forward(net, x):
    """The reference's UNet.forward (model.py:169-201) through the torch modules of the drop-in nn.Module —
    torch eager + cuDNN, informational comparator only (the product's forward never runs torch convolutions)."""
    import torch
    c1 = net.conv1(x); c2 = net.conv2(c1); c3 = net.conv3(c2); c4 = net.conv4(c3); c5 = net.conv5(c4); c6 = net.conv6(c5)
    d = net.deconv1_BAD(net.deconv1(c6, output_size=c5.size()))
    d = net.deconv2_BAD(net.deconv2(torch.cat([d, c5], 1), output_size=c4.size()))
    d = net.deconv3_BAD(net.deconv3(torch.cat([d, c4], 1), output_size=c3.size()))
    d = net.deconv4_BAD(net.deconv4(torch.cat([d, c3], 1), output_size=c2.size()))
    d = net.deconv5_BAD(net.deconv5(torch.cat([d, c2], 1), output_size=c1.size()))
    d = net.deconv6(torch.cat([d, c1], 1), output_size=x.size())
    return torch.sigmoid(d)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from svs_unet_pytorch_b200 import _lib, model as svs_model, pipeline, spectral, training

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU arm)")
    numa = pipeline.bind_to_gpu_numa_node(local)                       # before any pinned allocation
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(v: float) -> float:
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    K = max(1, args.steps)

    def timed(run_k, k=K, min_s=MIN_REGION_S):
        """run_k() enqueues k steps on the current stream.  Returns (ms per step, steps timed): the k-step loop is
        repeated until the region lasts >= min_s (repeat count agreed across ranks), events on the current stream,
        barrier + synchronize on both sides, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(); run_k(); e1.record()
        barrier()
        once = max_ranks(e0.elapsed_time(e1))
        reps = max(1, int(math.ceil(min_s * 1e3 / max(once, 1e-3))))
        barrier()
        e0.record()
        for _ in range(reps):
            run_k()
        e1.record()
        barrier()
        return max_ranks(e0.elapsed_time(e1)) / (reps * k), reps * k

    peaks = measured_peaks()
    torch.manual_seed(0)
    net = svs_model.UNet(precision=args.precision).eval().to(dev)
    plan = net.plan()
    flags = _lib.FLAG_APPLY_MASK                                      # hot path = mask + mask x mixture
    pool = 10                                                         # 10 x 16.8 MB inputs > 126 MB L2
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.rand(BATCH, 1, 512, 128, device=dev, generator=g) for _ in range(pool)]
    ys = [torch.empty_like(xs[0]) for _ in range(pool)]
    side = torch.cuda.Stream(dev)

    def make_graphs(p, first=0, last=11):
        """(graph of `pool` consecutive steps over the input pool, [one single-step graph per pool slot]).  Replaying
        the pool-sized graph amortises torch's per-replay bookkeeping (a 2 us RNG-offset fill kernel that
        CUDAGraph.replay() always enqueues) over ten steps; the single-step graphs cover K % pool."""
        for i in range(max(3, args.warmup)):
            p.forward_dense(xs[i % pool], flags, ys[i % pool])
        torch.cuda.synchronize()

        def one(i):
            if first == 0 and last == 11:
                p.forward_dense(xs[i], flags, ys[i])
            else:
                iv = _lib.PatchView(xs[i].data_ptr(), None, 512 * 128, 128, 1)
                ov = _lib.PatchView(ys[i].data_ptr(), None, 512 * 128, 128, 1)
                p.forward_views(iv, ov, None, BATCH, flags, first, last)

        singles = []
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            multi = torch.cuda.CUDAGraph()
            with torch.cuda.graph(multi, stream=side):
                for i in range(pool):
                    one(i)
            for i in range(pool):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    one(i)
                singles.append(gr)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        multi.replay()
        for i in range(min(pool, max(3, args.warmup))):
            singles[i].replay()
        return multi, singles

    def replay_k(graphs):
        multi, singles = graphs

        def run():
            for _ in range(K // pool):
                multi.replay()
            for i in range(K % pool):
                singles[i].replay()
        return run

    # ---- device-resident throughput (the headline `value`) ----
    graphs = make_graphs(plan)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, n_dev = timed(replay_k(graphs))
    clocks = sampler.stop() if rank == 0 else None

    # ---- informational: TWO independent 64-patch forwards in flight (two streams, two workspaces, one plan).  The
    # layers that leave SMs idle (128 CTAs on 148 SMs, one-tile-per-CTA tails) fill up from the other batch; `value`
    # stays the one-in-flight figure so that `layer_us` keeps adding up to it.
    two_in_flight = None
    try:
        s2 = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        g2 = []
        for si, st2 in enumerate(s2):
            idx = list(range(si, pool, 2))
            st2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st2):
                for i in idx[:2]:
                    plan.forward_dense(xs[i], flags, ys[i])
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st2):
                with torch.cuda.graph(gr, stream=st2):
                    for i in idx:
                        plan.forward_dense(xs[i], flags, ys[i])
            torch.cuda.synchronize()
            g2.append(gr)

        def run_two():
            cur = torch.cuda.current_stream()
            for st2, gr in zip(s2, g2):
                st2.wait_stream(cur)
                with torch.cuda.stream(st2):
                    gr.replay()
            for st2 in s2:
                cur.wait_stream(st2)
        ms_two, n_two = timed(run_two, k=pool)
        two_in_flight = {"patches_per_sec": BATCH * world / (ms_two * 1e-3), "ms_per_step": ms_two, "timed_steps": n_two,
                         "note": "two CUDA graphs of five 64-patch forwards each replayed on two streams at once"}
        del g2, s2
    except Exception as e:                                            # informational only
        two_in_flight = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- end to end through the host API, and the copy-only ceiling of the same transfers ----
    n_host = 8
    host_in = [torch.rand(BATCH, 1, 512, 128).pin_memory() for _ in range(n_host)]
    host_out = [torch.empty(BATCH, 1, 512, 128).pin_memory() for _ in range(n_host)]
    seq_in = [host_in[i % n_host] for i in range(K)]
    seq_out = [host_out[i % n_host] for i in range(K)]
    streamer = pipeline.PatchStreamer(net, BATCH, vocal_solo=True)
    streamer.run(seq_in[:max(3, min(K, args.warmup))], seq_out[:max(3, min(K, args.warmup))])
    ms_e2e, n_e2e = timed(lambda: streamer.run(seq_in, seq_out))
    copier = pipeline.CopyCeiling(dev, BATCH * 512 * 128 * 4)
    copier.run(host_in[0], host_out[0], 3)
    ms_copy, _ = timed(lambda: copier.run(host_in[0], host_out[0], K))
    step_bytes = BATCH * 512 * 128 * 4

    # ---- the tcgen05 conv family alone (conv2..deconv5) and per-layer device time ----
    g_tc = make_graphs(plan, 1, 10)
    ms_tc, _ = timed(replay_k(g_tc))
    # per-layer device time INSIDE the forward: CUDA graphs of the prefixes conv1..layer, rotating over the input pool
    # like the headline; a layer's figure is the difference of consecutive prefixes (so the twelve add up to the step)
    prefix_ms = []
    for li in range(12):
        if li == 11:
            prefix_ms.append(ms_dev)
            continue
        g_pre = make_graphs(plan, 0, li)
        t_pre, _ = timed(replay_k(g_pre), min_s=0.15)
        prefix_ms.append(t_pre)
        del g_pre
    layer_ms = [prefix_ms[0]] + [prefix_ms[i] - prefix_ms[i - 1] for i in range(1, 12)]
    del g_tc

    # ---- full-song pipeline (BASELINE configs[2]/[3]): STFT -> UNet mask -> iSTFT, songs sharded by rank ----
    corpus, seconds = 150, 180.0
    mine = list(range(rank, corpus, world))                           # song i -> rank i % world, no collective
    n_samp = int(seconds * 8192)
    ga = torch.Generator(device=dev).manual_seed(99 + rank)
    audio = torch.randn(len(mine) * n_samp, device=dev, generator=ga) * 0.1
    sbatch = spectral.SongBatch(audio, [n_samp] * len(mine))
    sep = pipeline.Separator(net)                                     # 512-patch UNet batches, staged patches
    for _ in range(2):
        sep.separate_batch(sbatch)
    ms_pipe, _ = timed(lambda: sep.separate_batch(sbatch), k=1)
    # K1 / K2 alone on this rank's shard of the corpus (>= 350 MB of traffic per launch: larger than L2)
    mag, phase, smax = sbatch.stft()
    ms_stft, _ = timed(lambda: sbatch.stft(), k=1)
    ms_istft, _ = timed(lambda: sbatch.istft(mag, phase, peak_normalize=False), k=1)
    ms_istft_norm, _ = timed(lambda: sbatch.istft(mag, phase, peak_normalize=True), k=1)
    frames = sbatch.total_frames
    del mag, phase, smax
    # the same corpus from pinned host audio to pinned host waveforms
    song_lengths = [n_samp] * len(mine)
    song_streamer = pipeline.SongStreamer(net)
    host_audio = audio.cpu().pin_memory()
    host_wave = torch.empty(sbatch.total_wave, dtype=torch.float32).pin_memory()
    for _ in range(2):
        song_streamer.run(host_audio, song_lengths, host_wave)
    ms_pipe_host, _ = timed(lambda: song_streamer.run(host_audio, song_lengths, host_wave), k=1)
    # PCM_16 in / PCM_16 out: the real file boundary of the reference (data.py:78,166 read and write 16-bit WAV)
    host_pcm = (host_audio.clamp(-1, 1) * 32767.0).round().to(torch.int16).pin_memory()
    host_pcm_out = torch.empty(sbatch.total_wave, dtype=torch.int16).pin_memory()
    for _ in range(2):
        song_streamer.run(host_pcm, song_lengths, host_pcm_out)
    ms_pipe_pcm, _ = timed(lambda: song_streamer.run(host_pcm, song_lengths, host_pcm_out), k=1)
    del host_audio, host_wave, host_pcm, host_pcm_out, audio, sbatch

    # ---- the other precision of BASELINE configs[1] ("fp32/TF32 and bf16"): kind::tf32 path, same measurements ----
    tf32 = None
    if args.precision == "bf16":
        net32 = svs_model.UNet(precision="tf32").eval().to(dev)
        net32.load_state_dict(net.state_dict())
        plan32 = net32.plan()
        g32 = make_graphs(plan32)
        ms_tf32, n_tf32 = timed(replay_k(g32))
        st32 = pipeline.PatchStreamer(net32, BATCH, vocal_solo=True)
        st32.run(seq_in[:3], seq_out[:3])
        ms_tf32_e2e, _ = timed(lambda: st32.run(seq_in, seq_out))
        # TF32 tensor peak, measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3, best of 10
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a32 = torch.randn(8192, 8192, device=dev); b32 = torch.randn(8192, 8192, device=dev)
        best = 1e9
        for i in range(13):
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record(); torch.matmul(a32, b32); q1.record()
            torch.cuda.synchronize()
            if i >= 3:
                best = min(best, q0.elapsed_time(q1))
        torch.backends.cuda.matmul.allow_tf32 = old
        tf32_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a32, b32, g32, st32, plan32, net32
        ach32 = BATCH * GFLOP_EXACT_PER_PATCH / ms_tf32
        tf32 = {"patches_per_sec": BATCH * world / (ms_tf32 * 1e-3), "ms_per_step": ms_tf32,
                "timed_steps": n_tf32, "timing": "CUDA-graph replay, same pool rotation as the bf16 value",
                "e2e": {"value": BATCH * world / (ms_tf32_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_tf32_e2e},
                "roofline": {"bound": "tensor", "achieved": ach32, "peak": tf32_peak, "unit": "TFLOP/s",
                             "frac": ach32 / tf32_peak,
                             "peak_source": "torch.matmul fp32 with allow_tf32, 8192^3, best of 10, measured in this run"}}

    # ---- informational comparator: the same layer list through torch eager + cuDNN on this GPU ----
    cudnn = None
    if rank == 0 and not args.no_cudnn_baseline:
        try:
            with torch.no_grad():
                def eager(n_, x_):
                    def run():
                        for _ in range(K):
                            m = torch_eager_forward(n_, x_)
                            _ = x_ * m
                    return run
                torch.backends.cudnn.allow_tf32 = True                 # torch default for convolutions
                for _ in range(3):
                    torch_eager_forward(net, xs[0])
                ms_c32 = _plain_timed(torch, eager(net, xs[0]), K)
                torch.backends.cudnn.benchmark = True
                netb = svs_model.UNet().eval().to(dev)
                netb.load_state_dict(net.state_dict())
                netb = netb.to(torch.bfloat16).to(memory_format=torch.channels_last)
                xb = xs[0].to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                for _ in range(3):
                    torch_eager_forward(netb, xb)
                ms_cb = _plain_timed(torch, eager(netb, xb), K)
                torch.backends.cudnn.benchmark = False
                del netb, xb
            cudnn = {"fp32_tf32_nchw": {"patches_per_sec": BATCH / (ms_c32 * 1e-3), "ms_per_step": ms_c32},
                     "bf16_channels_last": {"patches_per_sec": BATCH / (ms_cb * 1e-3), "ms_per_step": ms_cb},
                     "note": "informational: the reference's layer list (model.py:47-109,169-201) as torch eager + "
                             "cuDNN on this same B200, batch 64, 1 GPU, >= 0.5 s regions; fp32 NCHW with TF32 convs "
                             "is what reference inference.py:40 would run, bf16 channels_last + cudnn.benchmark is "
                             "its best case.  Not on the product path."}
        except Exception as e:                                        # a comparator must never fail the bench
            cudnn = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- BASELINE configs[4]: training step, batch 64 per GPU (data parallel: NCCL gradient all-reduce) ----
    train = None
    if not args.no_train:
        torch.manual_seed(0)
        tnet = svs_model.UNet().train().to(dev)
        gt = torch.Generator(device=dev).manual_seed(rank)
        tmix = torch.rand(BATCH, 1, 512, 128, device=dev, generator=gt)
        tvoc = tmix * torch.rand(BATCH, 1, 512, 128, device=dev, generator=gt)
        for _ in range(3):
            training.train_step(tnet, tmix, tvoc)
        k_tr = min(K, 20)

        def tr(sync):
            def run():
                for _ in range(k_tr):
                    training.train_step(tnet, tmix, tvoc, sync_grads=sync)
            return run
        ms_tr, n_tr = timed(tr(True), k=k_tr)
        ms_tr_local = ms_tr
        if world > 1:
            ms_tr_local, _ = timed(tr(False), k=k_tr)
        train = {"workload": "UNet training step, two-term masked L1, batch 64 per GPU, live Dropout2d + batch-stat "
                             "BatchNorm, Adam (BASELINE configs[4])",
                 "ms_per_step": ms_tr, "patches_per_sec": BATCH * world / (ms_tr * 1e-3), "timed_steps": n_tr,
                 "tflops_exact": BATCH * GFLOP_TRAIN_PER_PATCH / ms_tr,
                 "allreduce_exposed_ms": max(0.0, ms_tr - ms_tr_local) if world > 1 else 0.0,
                 "allreduce_payload_bytes": 9823313 * 4 if world > 1 else 0,
                 "precision": training.train_precision(tnet), "cuda_graph": True}
        if tf32 is not None:                                          # forward + dgrad + wgrad FLOPs against the TF32 peak
            train["roofline"] = {"bound": "tensor", "achieved": train["tflops_exact"], "peak": tf32["roofline"]["peak"],
                                 "unit": "TFLOP/s", "frac": train["tflops_exact"] / tf32["roofline"]["peak"],
                                 "note": "whole step incl. BatchNorm passes, loss and Adam; 3.961 GFLOP/patch exact"}
        del tnet, tmix, tvoc

    if rank == 0:
        value = BATCH * world / (ms_dev * 1e-3)
        e2e_val = BATCH * world / (ms_e2e * 1e-3)
        launches = plan.launch_count(BATCH)
        whole_tflops = BATCH * GFLOP_EXACT_PER_PATCH / ms_dev         # GFLOP / ms == TFLOP/s
        tc_flops = sum(LAYER_MMAC[li] for li in range(1, 11)) * 2e6 * BATCH
        tc_tflops = tc_flops / (ms_tc * 1e-3) / 1e12
        traffic, traffic_note = None, None
        for name in ("r02b_traffic.json", "r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):                                 # DRAM bytes of the same kernels from the committed ncu capture
                with open(tpath) as f:
                    tj = json.load(f)
                traffic = tj.get("unet_forward_dram_bytes", tj.get("tc_conv_family_dram_bytes_per_forward"))
                traffic_note = f"profiles/{name}: dram read+write bytes of one 64-patch forward (ncu --set full, cold L2)"
                break
        copy_gbs = 2 * step_bytes / (ms_copy * 1e-3) / 1e9            # both directions, this rank
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "timed_steps": n_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.precision],
            "data": "synthetic (torch.rand patches, random-init weights seed 0)",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH,
                       "patch": "512x128", "precision": args.precision, "parallelism": f"patch-batch sharding x{world}, no collective",
                       "l2": "inputs/outputs rotate over 10 distinct 16.8 MB batches (336 MB > 126 MB L2); CUDA-graph replay "
                             "(one graph = the 10 steps of a pool rotation)",
                       "timed_region": f">= {MIN_REGION_S} s per measurement: the K-step loop is repeated back to back"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": step_bytes,
                    "d2h_bytes_per_step": step_bytes, "ms_per_step": ms_e2e, "timed_steps": n_e2e,
                    "api": "pipeline.PatchStreamer.run -> svs_patch_stream_run (pinned host -> H2D -> svs_unet_forward -> D2H, 3 streams, native loop)",
                    "frac_of_pcie": ms_copy / ms_e2e},
            "pcie": {"copy_only_ms_per_step": ms_copy, "gbs_per_rank_both_directions": copy_gbs,
                     "ceiling_patches_per_sec": BATCH * world / (ms_copy * 1e-3),
                     "note": f"the step's H2D + D2H copies alone (pinned, two streams, no kernels), all {world} ranks at once"},
            "gpu_launches": launches * n_dev,
            "launches_per_step": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": whole_tflops, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": whole_tflops / peaks["bf16_burst"], "traffic": traffic, "traffic_note": traffic_note,
                         "kernel": f"whole UNet forward: all {launches} launches of svs_unet_forward (conv1_zc, zc_conv, "
                                   "tc_conv(_ck), deconv6_tc), 64 patches",
                         "ms_per_step": ms_dev,
                         "peak_source": peaks["source"] + " bf16_tflops (burst)",
                         "flops": "exact valid-tap count, 1.3247 GFLOP/patch (SURVEY.md 8d)",
                         "frac_of_sustained_peak": whole_tflops / peaks["bf16_sustained"]},
            "roofline_tc_family": {"bound": "tensor", "achieved": tc_tflops, "peak": peaks["bf16_burst"],
                                   "unit": "TFLOP/s", "frac": tc_tflops / peaks["bf16_burst"], "ms_per_step": ms_tc,
                                   "kernel": "tcgen05 implicit-GEMM layers conv2..deconv5 alone (graph of 10 layers)"},
            "roofline_stft": {"bound": "hbm", "achieved": frames * BYTES_PER_FRAME / (ms_stft * 1e-3) / 1e9,
                              "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": frames * BYTES_PER_FRAME / (ms_stft * 1e-3) / 1e9 / peaks["hbm_gbs"],
                              "ms": ms_stft, "frames": frames, "kernel": "stft_mag_phase_kernel (K1)",
                              "bytes_per_frame": BYTES_PER_FRAME},
            "roofline_istft": {"bound": "hbm", "achieved": frames * BYTES_PER_FRAME / (ms_istft * 1e-3) / 1e9,
                               "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": frames * BYTES_PER_FRAME / (ms_istft * 1e-3) / 1e9 / peaks["hbm_gbs"],
                               "ms": ms_istft, "frames": frames, "kernel": "istft_ola_kernel (K2)",
                               "peak_normalize_extra_ms": ms_istft_norm - ms_istft},
            "audio_sec_per_sec": value * AUDIO_S_PER_PATCH,
            "pipeline": {"workload": "150 synthetic 3-min songs (BASELINE configs[3]) sharded by song, device resident: "
                                     "STFT -> /max -> UNet mask x mixture -> iSTFT -> 0.9 peak",
                         "audio_sec_per_sec": corpus * seconds / (ms_pipe * 1e-3), "ms_per_corpus": ms_pipe,
                         "patches_per_sec": corpus * 16 / (ms_pipe * 1e-3),
                         "unet_batch": sep.max_batch, "patch_staging": "svs_patches_gather / svs_patches_scatter",
                         "host_to_host": {"audio_sec_per_sec": corpus * seconds / (ms_pipe_host * 1e-3),
                                          "ms_per_corpus": ms_pipe_host,
                                          "h2d_bytes": len(mine) * n_samp * 4, "d2h_bytes": len(mine) * n_samp * 4,
                                          "api": "pipeline.SongStreamer.run (pinned float32 host audio -> pinned "
                                                 "float32 host waveforms, song chunks on four streams)"},
                         "host_to_host_pcm16": {"audio_sec_per_sec": corpus * seconds / (ms_pipe_pcm * 1e-3),
                                                "ms_per_corpus": ms_pipe_pcm,
                                                "h2d_bytes": len(mine) * n_samp * 2, "d2h_bytes": len(mine) * n_samp * 2,
                                                "api": "pipeline.SongStreamer.run with int16 PCM in / out (the "
                                                       "reference's file boundary, data.py:78,166): int16 -> float in "
                                                       "K1's load, 0.9/peak + round to int16 in the normalise pass"}},
            "audio_sec_per_sec_host_to_host": corpus * seconds / (ms_pipe_pcm * 1e-3),
            "tflops_exact_whole_net": whole_tflops,
            "tf32": tf32,
            "two_in_flight": two_in_flight,
            "cudnn_baseline": cudnn,
            "train": train,
            "layer_us": {LAYER_NAMES[li]: round(layer_ms[li] * 1e3, 1) for li in range(12)},
            "layer_us_note": "incremental device time of each layer inside the 64-patch forward: differences of CUDA-graph "
                             "replays of the prefixes conv1..layer (same input rotation as `value`)",
            "layer_frac_of_burst_peak": {LAYER_NAMES[li]: round(LAYER_MMAC[li] * 2e6 * BATCH / (max(layer_ms[li], 1e-6) * 1e-3)
                                                                / 1e12 / peaks["bf16_burst"], 3) for li in range(12)},
        }
        if cudnn and "bf16_channels_last" in cudnn:
            line["vs_cudnn"] = {"bf16_vs_cudnn_bf16_channels_last": (BATCH / (ms_dev * 1e-3)) / cudnn["bf16_channels_last"]["patches_per_sec"],
                                "tf32_vs_cudnn_fp32_tf32_nchw": None if tf32 is None else
                                (BATCH / (tf32["ms_per_step"] * 1e-3)) / cudnn["fp32_tf32_nchw"]["patches_per_sec"]}
        line["numa"] = {k: v for k, v in numa.items() if k != "original_affinity"}
        if world == 1 and not args.no_cpu_baseline:
            if numa.get("original_affinity"):                         # the CPU arm gets every host core back
                os.sched_setaffinity(0, numa["original_affinity"])
            pps, sec, threads = cpu_reference_run(steps=3, warmup=1)
            pps_b1 = cpu_reference_b1_loop()
            line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "3 forwards of the same 64-patch batch after 1 warm-up, oracle "
                                              "restatement of reference model.py on torch-CPU fp32",
                                    "b1_loop_value": pps_b1,
                                    "b1_loop_sample": "16 patches, one forward per patch as reference "
                                                      "inference.py:79-116 really runs"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _plain_timed(torch, run_k, k, min_s=MIN_REGION_S):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); run_k(); e1.record()
    torch.cuda.synchronize()
    reps = max(1, int(math.ceil(min_s * 1e3 / max(e0.elapsed_time(e1), 1e-3))))
    e0.record()
    for _ in range(reps):
        run_k()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cudnn-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
