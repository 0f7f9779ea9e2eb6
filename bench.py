#!/usr/bin/env python
"""bench.py — headline benchmark of the SVS-UNet separation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|tf32|fp32]

Workload (BASELINE.json configs[1]): UNet mask inference on a batch of 64 synthetic 512x128 magnitude
patches (torch.rand, the range of spec/norm), random-init weights (torch.manual_seed(0)).  One "step" =
one pass of the hot path (UNet.forward + mask application) over one 64-patch batch per GPU.
Multi-GPU: patches shard by batch across ranks with NO data-path collective ("scaling": "weak").

* ``value``  patches/s with inputs resident in HBM (CUDA-graph replay of the forward, CUDA-event timed,
  max over ranks); inputs rotate over a pool larger than L2.
* ``e2e``    the same metric through the public host API (pipeline.PatchStreamer): pinned host buffers,
  H2D + kernels + D2H all inside the timed region.
* ``roofline``  the tcgen05 implicit-GEMM conv kernels (the dense contraction): exact-tap FLOPs of the
  layers they execute / their summed device time, measured live with CUDA events.
* ``cpu_baseline`` / ``--impl reference``  the reference's CPU path (oracle restatement of model.py on
  torch-CPU, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
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
AUDIO_S_PER_PATCH = 128 * 768 / 8192.0           # 12.0 s of 8192 Hz audio per full patch
GFLOP_EXACT_PER_PATCH = 1.3247                   # 2 x 662,350,768 valid-tap MACs
# exact (valid-tap) MACs per patch per layer, SURVEY.md section 8(a): conv1..conv6, deconv1..deconv6
LAYER_MMAC = [6.477, 51.205, 49.990, 47.587, 42.893, 33.948, 33.948, 85.787, 95.175, 99.979, 102.409, 12.954]
LAYER_NAMES = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "deconv1", "deconv2", "deconv3", "deconv4",
               "deconv5", "deconv6"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
                "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


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
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, batch: int = BATCH):
    """The reference's CPU implementation of the path (oracle restatement of model.py, torch-CPU fp32,
    all host threads): returns (patches_per_s, seconds_per_step, threads)."""
    import torch
    from oracle import unet_oracle
    from svs_unet_pytorch_b200 import model as svs_model
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = svs_model.UNet().eval()
    sd = net.state_dict()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(batch, 1, 512, 128, generator=g)
    with torch.no_grad():
        for _ in range(warmup):
            unet_oracle.unet_forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            m = unet_oracle.unet_forward(sd, x)
            _ = x * m                                                  # inference.py:107
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 12))                               # bounded sample: ~1-3 s per step
    warm = max(1, min(args.warmup, 2))
    pps, sec, threads = cpu_reference_run(steps, warm)
    sample = f"{steps} steps x {BATCH} patches after {warm} warm-up (batch-64 best case of reference model.py on torch-CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic (torch.rand patches, random-init weights seed 0)",
        "config": {"workload": "unet_mask_inference_b64_512x128 (BASELINE configs[1])", "batch_per_gpu": BATCH,
                   "patch": "512x128"},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "audio_sec_per_sec": pps * AUDIO_S_PER_PATCH,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from svs_unet_pytorch_b200 import _lib, model as svs_model, pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(0)
    net = svs_model.UNet(precision=args.precision).eval().to(dev)
    plan = net.plan()
    flags = _lib.FLAG_APPLY_MASK                                      # hot path = mask + mask x mixture
    pool = 10                                                         # 10 x 16.8 MB inputs > 126 MB L2
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.rand(BATCH, 1, 512, 128, device=dev, generator=g) for _ in range(pool)]
    ys = [torch.empty_like(xs[0]) for _ in range(pool)]
    for i in range(max(3, args.warmup)):
        plan.forward_dense(xs[i % pool], flags, ys[i % pool])
    torch.cuda.synchronize()

    # ---- CUDA graphs of the forward, one per pool slot ----
    graphs = []
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(pool):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                plan.forward_dense(xs[i], flags, ys[i])
            graphs.append(gr)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for i in range(max(3, args.warmup)):
        graphs[i % pool].replay()
    barrier()

    # ---- device-resident throughput ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        graphs[i % pool].replay()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host API ----
    n_host = min(args.steps, 8)
    host_in = [torch.rand(BATCH, 1, 512, 128).pin_memory() for _ in range(n_host)]
    host_out = [torch.empty(BATCH, 1, 512, 128).pin_memory() for _ in range(n_host)]
    streamer = pipeline.PatchStreamer(net, BATCH, vocal_solo=True)
    seq_in = [host_in[i % n_host] for i in range(args.steps)]
    seq_out = [host_out[i % n_host] for i in range(args.steps)]
    streamer.run(seq_in[:max(3, args.warmup)], seq_out[:max(3, args.warmup)])
    barrier()
    t0 = time.perf_counter()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    streamer.run(seq_in, seq_out)
    h1.record()
    barrier()
    ms_e2e = max(h0.elapsed_time(h1), 0.0)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e, 0.0)

    # ---- per-layer device time (roofline of the dominant kernel family) ----
    iv = _lib.PatchView(xs[0].data_ptr(), None, 512 * 128, 128, 1)
    ov = _lib.PatchView(ys[0].data_ptr(), None, 512 * 128, 128, 1)
    layer_ms = [0.0] * 12
    reps = 10
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)]
           for _ in range(reps)]
    for r in range(reps):
        for li in range(12):
            a, b = evs[r][li]
            a.record()
            plan.forward_views(iv, ov, None, BATCH, flags, li, li)
            b.record()
    torch.cuda.synchronize()
    for li in range(12):
        layer_ms[li] = sorted(evs[r][li][0].elapsed_time(evs[r][li][1]) for r in range(reps))[reps // 2]

    # ---- full-song pipeline (BASELINE configs[2]/[3]): STFT -> UNet mask -> iSTFT, songs sharded by rank ----
    from svs_unet_pytorch_b200 import spectral
    corpus, seconds = 150, 180.0
    mine = list(range(rank, corpus, world))                           # song i -> rank i % world, no collective
    n_samp = int(seconds * 8192)
    ga = torch.Generator(device=dev).manual_seed(99 + rank)
    audio = torch.randn(len(mine) * n_samp, device=dev, generator=ga) * 0.1
    sbatch = spectral.SongBatch(audio, [n_samp] * len(mine))
    sep = pipeline.Separator(net)                                     # 512-patch UNet batches, staged patches
    for _ in range(2):
        sep.separate_batch(sbatch)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p_reps = 5
    p0.record()
    for _ in range(p_reps):
        sep.separate_batch(sbatch)
    p1.record()
    barrier()
    ms_pipe = p0.elapsed_time(p1) / p_reps
    # the same corpus from pinned host audio to pinned host waveforms: chunks of songs on
    # alternating streams so the PCIe copies of one chunk overlap the kernels of the others
    host_audio = audio.cpu().pin_memory()
    host_wave = torch.empty(sbatch.total_wave, dtype=torch.float32).pin_memory()
    song_streamer = pipeline.SongStreamer(net)
    song_lengths = [n_samp] * len(mine)
    for _ in range(2):
        song_streamer.run(host_audio, song_lengths, host_wave)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(p_reps):
        song_streamer.run(host_audio, song_lengths, host_wave)
    h1.record()
    barrier()
    ms_pipe_host = h0.elapsed_time(h1) / p_reps

    # ---- the tcgen05 conv family alone (layers conv2..deconv5) as one CUDA graph: its device time per step ----
    g_tc = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g_tc, stream=side):
            plan.forward_views(iv, ov, None, BATCH, flags, 1, 10)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(5):
        g_tc.replay()
    torch.cuda.synchronize()
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tc_reps = min(args.steps, 200)
    t0e.record()
    for _ in range(tc_reps):
        g_tc.replay()
    t1e.record()
    torch.cuda.synchronize()
    ms_tc = t0e.elapsed_time(t1e) / tc_reps

    # ---- the other precision of BASELINE configs[1] ("fp32/TF32 and bf16"): TF32 tensor-core path, same workload ----
    ms_tf32 = None
    if args.precision == "bf16":
        net32 = svs_model.UNet(precision="tf32").eval().to(dev)
        net32.load_state_dict(net.state_dict())
        plan32 = net32.plan()
        for i in range(5):
            plan32.forward_dense(xs[i % pool], flags, ys[i % pool])
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n32 = min(args.steps, 100)
        q0.record()
        for i in range(n32):
            plan32.forward_dense(xs[i % pool], flags, ys[i % pool])
        q1.record()
        torch.cuda.synchronize()
        ms_tf32 = q0.elapsed_time(q1) / n32
        del plan32, net32

    times = torch.tensor([ms_dev, ms_e2e, ms_pipe, ms_pipe_host], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_pipe, ms_pipe_host = (float(times[i]) for i in range(4))

    if rank == 0:
        peaks = measured_peaks()
        total_patches = BATCH * args.steps * world
        value = total_patches / (ms_dev * 1e-3)
        e2e_val = total_patches / (ms_e2e * 1e-3)
        tc_layers = [li for li in range(12) if 1 <= li <= 10]
        tc_flops = sum(LAYER_MMAC[li] for li in tc_layers) * 2e6 * BATCH
        tc_ms = ms_tc                                                 # graph replay of exactly these launches
        achieved = tc_flops / (tc_ms * 1e-3) / 1e12
        launches = plan.launch_count(BATCH)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):                                     # DRAM bytes of the same kernels from the committed ncu capture
            with open(tpath) as f:
                traffic = json.load(f).get("tc_conv_family_dram_bytes_per_forward")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.precision],
            "data": "synthetic (torch.rand patches, random-init weights seed 0)",
            "config": {"workload": "unet_mask_inference_b64_512x128 (BASELINE configs[1])", "batch_per_gpu": BATCH,
                       "patch": "512x128", "precision": args.precision, "parallelism": f"patch-batch sharding x{world}, no collective",
                       "l2": "inputs/outputs rotate over 10 distinct 16.8 MB batches (336 MB > 126 MB L2); CUDA-graph replay"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": BATCH * 512 * 128 * 4,
                    "d2h_bytes_per_step": BATCH * 512 * 128 * 4, "ms_per_step": ms_e2e / args.steps,
                    "api": "pipeline.PatchStreamer.run (pinned host -> H2D -> svs_unet_forward -> D2H, 3 streams)",
                    "wall_ms": wall_e2e},
            "gpu_launches": launches * args.steps,
            "launches_per_step": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_sustained"], "traffic": traffic,
                         "traffic_note": "dram read+write bytes summed over the 10 conv2..deconv5 launches of one 64-patch forward (ncu --set full, cold L2); algorithmic activation bytes: 2 x 64 x 0.98 M bf16 elements = 251 MB + 41 MB weights",
                         "kernel": "zc_conv_kernel + tc_conv_kernel + tc_conv_ck_kernel (tcgen05 implicit GEMM, conv2..deconv5; "
                                   f"10 layers, {plan.launch_count(BATCH) - 2} launches)",
                         "ms_per_step": ms_tc,
                         "peak_source": peaks["source"] + " bf16_tflops_sustained",
                         "flops": "exact valid-tap count of the layers the kernel executes"},
            "audio_sec_per_sec": value * AUDIO_S_PER_PATCH,
            "pipeline": {"workload": "150 synthetic 3-min songs (BASELINE configs[3]) sharded by song, device resident: "
                                     "STFT -> /max -> UNet mask x mixture -> iSTFT -> 0.9 peak",
                         "audio_sec_per_sec": corpus * seconds / (ms_pipe * 1e-3), "ms_per_corpus": ms_pipe,
                         "patches_per_sec": corpus * 16 / (ms_pipe * 1e-3),
                         "unet_batch": sep.max_batch, "patch_staging": "svs_patches_gather / svs_patches_scatter",
                         "host_to_host": {"audio_sec_per_sec": corpus * seconds / (ms_pipe_host * 1e-3),
                                          "ms_per_corpus": ms_pipe_host,
                                          "h2d_bytes": len(mine) * n_samp * 4, "d2h_bytes": int(sbatch.total_wave) * 4,
                                          "api": "pipeline.SongStreamer.run (pinned host audio -> pinned host "
                                                 "waveforms, 10-song chunks on four streams; PCIe bound)"}},
            "tflops_exact_whole_net": value * GFLOP_EXACT_PER_PATCH / 1e3,
            "tf32": None if ms_tf32 is None else {"patches_per_sec_per_gpu": BATCH / (ms_tf32 * 1e-3), "ms_per_step": ms_tf32,
                                                  "note": "same workload on the kind::tf32 path, direct launches (no graph), rank 0"},
            "layer_us": {LAYER_NAMES[li]: round(layer_ms[li] * 1e3, 1) for li in range(12)},
        }
        if world == 1 and not args.no_cpu_baseline:
            pps, sec, threads = cpu_reference_run(steps=3, warmup=1)
            line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "3 forwards of the same 64-patch batch after 1 warm-up, oracle "
                                              "restatement of reference model.py on torch-CPU fp32"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
