#!/usr/bin/env python
"""Benchmark of the HILCodec encode -> RVQ -> decode hot path (BASELINE.json metric:
audio frames/s, 1 frame = 320 samples of 24 kHz audio = 1/75 s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one one-shot pass (zero caches) of the fused path over one batch of synthetic
clips.  Workload per GPU: BASELINE.json configs[2] = hil_music, 256 clips x 24000 samples,
n_q = 12 (the configuration the north-star target is quoted on); at N GPUs each rank runs its
own 256-clip shard (weak scaling; N = 8 is configs[4], 2048 clips) with no data-path
collective -- clips are independent.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the launch stream,
inputs resident in HBM, max over ranks; `e2e` goes through the C-ABI host-buffer call
(`hil_codec_forward_host`: pinned-host H2D, forward, D2H of indices + PCM, stream sync).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, clips per GPU, samples, n_q, BASELINE.json config it is)
    "music256": ("hil_music", 256, 24000, 12, "configs[2]: hil_music, batch=256x24000 @24 kHz, n_q=12"),
    "speech64": ("hil_speech", 64, 24000, 8, "configs[1]: hil_speech, batch=64x24000 @24 kHz, n_q=8"),
    # streaming (not a default bench line): `clips` concurrent streams fed hop by hop, one step = 75 hops = 1 s of audio
    "stream1": ("hil_music", 1, 24000, 12, "configs[3]: hil_music streaming, hop 320, per-frame causal cache, 1 stream"),
    "stream64": ("hil_music", 64, 24000, 12, "configs[3]: hil_music streaming, hop 320, per-frame causal cache, 64 streams"),
}
FLOP_PER_FRAME = {"hil_speech": 456.257e6, "hil_music": 457.306e6}
CATEGORIES = ["pointwise_gemm", "stft_gemm", "depthwise", "depthwise_transposed", "conv_pre", "conv_post_tanh",
              "rvq", "misc"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="music256", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_weights(model_name):
    from hilcodec_b200 import weights as W

    cfg = W.CONFIGS[model_name]
    if W.have_pretrained(model_name):
        return cfg, W.load_pretrained(model_name), "published weights (from the reference's ONNX files)"
    return cfg, W.random_weights(cfg, 0), "random-init weights of the same architecture"


def synth(batch, samples, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """`nvidia-smi -lms 200` in the background during the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_run(model_name, n_q, samples, steps, warmup, budget_s):
    """Time the CPU oracle port (oracle/hilcodec_oracle.py: the reference's deployment graph as
    torch CPU ops -- the reference itself is a Python package that cannot travel to the GPU
    box) on all host threads, on a bounded sample of the workload."""
    import numpy as np
    import torch

    from oracle import hilcodec_oracle as O

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    cfg, w, _ = load_weights(model_name)
    p = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}
    ocfg = O.CodecConfig(num_quantizers=cfg.num_quantizers)

    def run(x):
        t0 = time.perf_counter()
        with torch.no_grad():
            O.codec_forward(ocfg, p, x, n_q)
        return time.perf_counter() - t0

    run(synth(1, samples, 99))                      # page-in / thread pool start
    # the CPU path's throughput depends on the batch it is given (allocation / cache effects):
    # probe a few sub-batch sizes and give the reference its best one
    best_b, best_rate = 1, 0.0
    for b in (1, 2, 4, 8):
        t = run(synth(b, samples, 90 + b))
        if b / t > best_rate:
            best_b, best_rate = b, b / t
    total_steps = max(1, steps + warmup)
    reps = int(max(1, min(256 // best_b, budget_s / total_steps * best_rate / best_b)))
    batch = reps * best_b
    xs = [synth(best_b, samples, 1234 + i) for i in range(reps)]

    def step():
        return sum(run(x) for x in xs)

    for _ in range(warmup):
        step()
    times = [step() for _ in range(max(1, steps))]
    frames = batch * (samples // cfg.hop)
    sec = sum(times) / len(times)
    return {
        "value": frames / sec, "unit": "frames/s", "cores": cores, "kind": "port",
        "sample": f"{reps} x {best_b} clips x {samples} samples of the workload per step (sub-batch {best_b} = fastest "
                  f"of 1/2/4/8), {len(times)} timed steps ({sec:.2f} s/step), torch {torch.__version__} CPU, "
                  f"{cores} threads",
        "ms_per_step": sec * 1e3, "batch": batch,
    }


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model_name, clips, samples, n_q, desc = WORKLOADS[args.workload]
    r = cpu_reference_run(model_name, n_q, samples, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "audio frames/sec (24 kHz enc+RVQ+dec)", "value": r["value"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "model": model_name, "clips_per_step": r["batch"], "samples": samples, "n_q": n_q,
                   "note": "CPU: bounded sample of the workload per step"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- CUDA arm, streaming workloads
def main_stream(args):
    """BASELINE configs[3]: frame-by-frame streaming with GPU-resident caches.  One step = 75 sequential hops (1 s of
    audio) of every stream; a hop is one CUDA-graph replay (`hil_codec_forward_graph`).  Single GPU (streams of
    different GPUs are independent; `--gpus N` runs the same thing per rank and adds the rates up)."""
    import torch
    import torch.distributed as dist

    from hilcodec_b200 import _lib
    from hilcodec_b200 import streaming as S
    from hilcodec_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU path in hilcodec_b200)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model_name, B, samples, n_q, desc = WORKLOADS[args.workload]
    cfg, w, wdesc = load_weights(model_name)
    model = S.HILCodec.from_weights(w, cfg.num_quantizers).cuda()
    lib = _lib.load()
    hop, hops = cfg.hop, samples // cfg.hop
    warm = max(args.warmup, 3)
    x_host = synth(B, samples * (args.steps + warm + 2), 1234 + rank).pin_memory()
    x = x_host.to(dev)
    core = model._core
    hmodel, hstate = core.model(dev), core.state(dev, B)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    xin = torch.empty(B, 1, hop, dtype=torch.float32, device=dev)
    idx = torch.empty(n_q, B, 1, dtype=torch.int64, device=dev)
    y = torch.empty(B, 1, hop, dtype=torch.float32, device=dev)
    _lib.check(lib.hil_state_reset(hstate, sp))
    pos = [0]

    def step():
        for _ in range(hops):
            xin.copy_(x[:, :, pos[0]:pos[0] + hop])   # the hop that "arrives" (device-resident clip)
            _lib.check(lib.hil_codec_forward_graph(hmodel, hstate, xin.data_ptr(), B, hop, n_q, idx.data_ptr(), y.data_ptr(), sp))
            pos[0] += hop

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.hil_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = (lib.hil_launch_count() - l0) // max(args.steps, 1)
    ms = e0.elapsed_time(e1)

    # e2e: every hop comes from pinned host memory and its indices + PCM go back to the host (C-ABI host call)
    idx_host = torch.empty(n_q, B, 1, dtype=torch.int64).pin_memory()
    y_host = torch.empty(B, 1, hop, dtype=torch.float32).pin_memory()
    chunk_host = torch.empty(B, 1, hop, dtype=torch.float32).pin_memory()
    checksum = 0.0
    hpos = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps * hops):
        chunk_host.copy_(x_host[:, :, hpos:hpos + hop])
        _lib.check(lib.hil_codec_forward_host(hmodel, hstate, chunk_host.data_ptr(), B, hop, n_q, idx_host.data_ptr(),
                                              y_host.data_ptr(), sp))
        checksum += float(y_host[0, 0, 0])
        hpos += hop
    ms_e2e = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    ms, ms_e2e = sharding.reduce_max([ms, ms_e2e], device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    frames = B * hops * world
    weight_bytes = float(sum(v.size * 4 for v in w.values()))
    ms_hop = ms / (args.steps * hops)
    line = {
        "metric": "audio frames/sec (24 kHz enc+RVQ+dec)", "value": frames * args.steps / (ms * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": f"synthetic 0.1*randn audio, {wdesc}",
        "config": {"workload": desc, "model": model_name, "streams_per_gpu": B, "hop": hop, "hops_per_step": hops,
                   "n_q": n_q, "executor": "hil_codec_forward_graph (one CUDA-graph replay per hop)",
                   "l2": "no flush: a hop re-reads the ~50 MB of weights, which fit the 126 MB L2, by design"},
        "ms_per_hop": ms_hop, "x_realtime_per_stream": (hop / 24000.0) / (ms_hop * 1e-3),
        "e2e": {"value": frames * args.steps / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * hop * 4 * hops,
                "d2h_bytes_per_step": (n_q * B * 8 + B * hop * 4) * hops, "ms_per_step": ms_e2e / args.steps,
                "api": "hil_codec_forward_host per hop (pinned host buffers, eager launches)", "checksum": checksum},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # what a hop cannot avoid is reading every weight once; everything above that is latency
        "roofline": {"bound": "hbm", "kernel": "whole hop (latency-bound chain of ~115 dependent launches)",
                     "achieved": weight_bytes / (ms_hop * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": weight_bytes / (ms_hop * 1e-3) / 1e9 / hbm_peak, "traffic": None},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- CUDA arm
def main_ours(args):
    import torch
    import torch.distributed as dist

    from hilcodec_b200 import _lib
    from hilcodec_b200 import streaming as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU path in hilcodec_b200)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model_name, clips, samples, n_q, desc = WORKLOADS[args.workload]
    cfg, w, wdesc = load_weights(model_name)
    model = S.HILCodec.from_weights(w, cfg.num_quantizers).cuda()
    lib = _lib.load()
    B, T = clips, samples
    F = T // cfg.hop
    x_host = synth(B, T, 1234 + rank).pin_memory()
    x = x_host.to(dev)
    core = model._core
    hmodel = core.model(dev)
    hstate = core.state(dev, B)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    idx = torch.empty(n_q, B, F, dtype=torch.int64, device=dev)
    y = torch.empty(B, 1, T, dtype=torch.float32, device=dev)

    def step():
        _lib.check(lib.hil_state_reset(hstate, sp))
        _lib.check(lib.hil_codec_forward(hmodel, hstate, x.data_ptr(), B, T, n_q, None, idx.data_ptr(), y.data_ptr(), sp))

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    l0 = lib.hil_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    launches = (lib.hil_launch_count() - l0) // max(args.steps, 1)
    ms = e0.elapsed_time(e1)

    # ---- e2e: host buffers through the C-ABI call, copies inside the timed region
    idx_host = torch.empty(n_q, B, F, dtype=torch.int64).pin_memory()
    y_host = torch.empty(B, 1, T, dtype=torch.float32).pin_memory()

    def step_host():
        _lib.check(lib.hil_state_reset(hstate, sp))
        _lib.check(lib.hil_codec_forward_host(hmodel, hstate, x_host.data_ptr(), B, T, n_q, idx_host.data_ptr(),
                                              y_host.data_ptr(), sp))

    for _ in range(2):
        step_host()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e_checksum = float(y_host.abs().sum())  # the device->host result is really read on the host

    from hilcodec_b200 import sharding
    ms, ms_e2e = sharding.reduce_max([ms, ms_e2e], device=dev)  # a multi-GPU step is as slow as its slowest rank

    # ---- per-category kernel timing (separate pass, CUDA events around every launch)
    n_cat = len(CATEGORIES)
    arr_ms, arr_fl, arr_by = (C.c_double * n_cat)(), (C.c_double * n_cat)(), (C.c_double * n_cat)()
    arr_n = (C.c_int64 * n_cat)()
    prof_steps = 2
    torch.cuda.synchronize()
    _lib.check(lib.hil_profile_begin())
    for _ in range(prof_steps):
        step()
    _lib.check(lib.hil_profile_end(arr_ms, arr_fl, arr_by, arr_n, n_cat))
    cats = {}
    for i, name in enumerate(CATEGORIES):
        if arr_n[i]:
            cats[name] = {"ms_per_step": arr_ms[i] / prof_steps, "launches_per_step": arr_n[i] // prof_steps,
                          "tflops": arr_fl[i] / (arr_ms[i] * 1e-3) / 1e12 if arr_ms[i] > 0 else 0.0,
                          "gbs": arr_by[i] / (arr_ms[i] * 1e-3) / 1e9 if arr_ms[i] > 0 else 0.0}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = {}, "fallback (B200_PROFILING.md): 6650 GB/s, 1590 TFLOP/s bf16"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_src = "MEASURED_PEAKS.json"
    except Exception:
        pass
    bf16_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    traffic, traffic_step = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_step = tj.get("dram_bytes_per_launch_avg"), tj.get("dram_bytes_per_step")
    except Exception:
        pass
    pw = cats.get("pointwise_gemm", {"tflops": 0.0, "gbs": 0.0, "ms_per_step": 0.0, "launches_per_step": 0})
    if traffic_step is not None and model_name == "hil_music" and B == 256:
        traffic = traffic_step                        # ncu DRAM bytes of the category's launches in one music256 step
    elif traffic is not None:
        traffic = traffic * pw["launches_per_step"]   # other workloads: per-launch average x launches (rough)
    dominant = max(cats, key=lambda k: cats[k]["ms_per_step"]) if cats else None
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # The dominant kernels are the tensor-core GEMMs (plain 1x1, fused DWS block, fused ResBlock, fused upsampling
    # layer).  Two rooflines apply to them; the binding one is the one with the higher time floor for the step's
    # algorithmic work: HBM (bytes / measured copy bandwidth) or the tensor pipe (3 fp16 MMAs per fp32-accurate
    # product -> FLOPs * 3 / measured bf16 throughput).  Both fractions are reported.
    MMA_PER_PRODUCT = 3.0
    frac_hbm = pw["gbs"] / hbm_peak if hbm_peak else 0.0
    frac_tensor = pw["tflops"] * MMA_PER_PRODUCT / bf16_peak if bf16_peak else 0.0
    bound = "hbm" if frac_hbm >= frac_tensor else "tensor"
    roofline = {
        "bound": bound,
        "kernel": "th::gemm_h_kernel / rb::resblock_kernel (tcgen05 kind::f16 hi/lo-split GEMMs: 1x1 conv, fused DWS "
                  "block, fused ResBlock, fused upsampling layer; all launches of the step)",
        "achieved": pw["gbs"] if bound == "hbm" else pw["tflops"],
        "peak": hbm_peak if bound == "hbm" else bf16_peak,
        "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
        "frac": frac_hbm if bound == "hbm" else pw["tflops"] / bf16_peak,
        "traffic": traffic,
        "peak_source": f"{peak_src}: hbm_gbs / bf16_tflops_sustained (kernels timed inside a long step)",
        "hbm": {"achieved_gbs": pw["gbs"], "peak_gbs": hbm_peak, "frac": frac_hbm,
                "bytes": "algorithmic: every launch reads its inputs once and writes its outputs once (codec.cu HIL_LAUNCH)"},
        "tensor": {"achieved_tflops": pw["tflops"], "peak_tflops": bf16_peak, "frac_of_bf16_peak": pw["tflops"] / bf16_peak,
                   "mma_per_product": MMA_PER_PRODUCT, "frac_of_reachable": frac_tensor},
        "share_of_step": pw["ms_per_step"] / (sum(c["ms_per_step"] for c in cats.values()) or 1.0),
        "dominant_category": dominant,
        "note": "fp32-accurate arithmetic is required for bit-exact VQ indices: every product is 3 fp16 MMAs "
                "(hi*hi, hi*lo, lo*hi) into two fp32 TMEM accumulators, so the tensor ceiling is bf16 peak / 3; the "
                "binding roofline is the one with the larger fraction",
    }

    frames_total = B * F * world
    value = frames_total * args.steps / (ms * 1e-3)
    e2e_value = frames_total * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "audio frames/sec (24 kHz enc+RVQ+dec)", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": f"synthetic 0.1*randn audio, {wdesc}",
        "config": {"workload": desc, "model": model_name, "clips_per_gpu": B, "samples": T, "n_q": n_q,
                   "sharding": f"batch-sharded x{world}, no data-path collective",
                   "l2": "no explicit flush: one step streams >10 GB of activations through the 126 MB L2"},
        "rtf_x_realtime": value / 75.0,
        "model_tflops": value * FLOP_PER_FRAME[model_name] / 1e12,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * T * 4,
                "d2h_bytes_per_step": n_q * B * F * 8 + B * T * 4, "ms_per_step": ms_e2e / args.steps,
                "api": "hil_codec_forward_host (pinned host buffers)", "checksum": e2e_checksum},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernel_categories": cats,
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(model_name, n_q, samples, steps=2, warmup=0, budget_s=40.0)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    elif a.workload.startswith("stream"):
        main_stream(a)
    else:
        main_ours(a)
