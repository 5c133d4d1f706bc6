#!/usr/bin/env python
"""Benchmark of the HILCodec encode -> RVQ -> decode hot path (BASELINE.json metric:
audio frames/s, 1 frame = 320 samples of 24 kHz audio = 1/75 s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

Headline workload (one "step" = one one-shot pass, zero caches, of the fused path over one batch of synthetic clips):
BASELINE.json configs[2] = hil_music, 256 clips x 24000 samples, n_q = 12 per GPU -- the configuration the north-star
target is quoted on; at N GPUs each rank runs its own 256-clip shard (weak scaling; N = 8 is configs[4], 2048 clips)
with no data-path collective -- clips are independent.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the launch stream, inputs resident in HBM, max
over ranks; `e2e` goes through the C-ABI host-buffer call (`hil_codec_forward_host`: pinned-host H2D, forward, D2H of
indices + PCM, stream sync).  At N = 1 with no --workload the same line also carries `other_workloads`: configs[1]
(speech64), configs[0] (speech1: the single clip the CPU reference is quoted on) and configs[3] (stream1 / stream64: hil_music fed hop by hop, GPU-resident caches), each with its own value /
e2e / roofline / cpu_baseline.  `--workload X` runs only X.

CPU legs (`cpu_baseline`, `--impl reference`): the reference's OWN classes (models/hilcodec/streaming.py, unmodified
copy in oracle/_ref made by oracle/build_ref.py) with the published weights on the box's host cores, `kind:
"reference"`; if that copy is absent, the oracle port (`kind: "port"`).  Streaming is timed with the reference's
protocol (scripts/HILCodec Onnx.ipynb cell 3, test_onnx.py:75-135): one 320-sample hop per call of encoder, quantizer,
dequantizer and decoder, caches handed back as tensors -- with 1 thread (the reference's published setting,
HILCodec Onnx.ipynb:35, test_onnx.py:4-8) and with all host cores.
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
    "speech1": ("hil_speech", 1, 24000, 8, "configs[0]: hil_speech, 1 utterance x 24000 @24 kHz, n_q=8 (the CPU reference's case)"),
    # streaming: `clips` concurrent streams fed hop by hop, one step = 75 hops = 1 s of audio per stream
    "stream1": ("hil_music", 1, 24000, 12, "configs[3]: hil_music streaming, hop 320, per-frame causal cache, 1 stream"),
    "stream64": ("hil_music", 64, 24000, 12, "configs[3]: hil_music streaming, hop 320, per-frame causal cache, 64 streams"),
}
FLOP_PER_FRAME = {"hil_speech": 456.257e6, "hil_music": 457.306e6}
FUSED_UNIT_BYTES_PER_FRAME = 4.2e6     # only ResBlock / resample-block inputs + outputs touch HBM
LAYER_BOUNDARY_BYTES_PER_FRAME = 15.77e6   # every conv output written once and read once
CATEGORIES = ["pointwise_gemm_narrow", "stft_gemm", "depthwise", "depthwise_transposed", "conv_pre", "conv_post_tanh",
              "rvq", "misc", "pointwise_gemm_wide", "resblock_fused"]
GEMM_CLASSES = {   # layer classes of the tensor-core GEMM kernels and the roofline that binds each
    "pointwise_gemm_narrow": ("hbm", "th::gemm_h_kernel on layers with Cin, Cout < 384 (1x1, fused DWSBlock, fused upsampling)"),
    "resblock_fused": ("hbm", "rb::resblock_kernel<128> (whole ResBlock, C <= 128)"),
    "pointwise_gemm_wide": ("tensor", "th::gemm_h_kernel on layers with Cin or Cout >= 384"),
}
MMA_PER_PRODUCT = 3.0   # fp32-accurate product = hi*hi + hi*lo + lo*hi on kind::f16


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
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


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {}, "fallback (B200_PROFILING.md): 6650 GB/s, 1590 TFLOP/s bf16"


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
        return self

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


# ----------------------------------------------------------------------------- CPU legs
class CpuCodec:
    """The reference's CPU implementation of the path: its own streaming.py classes when oracle/_ref (or the full
    reference tree) is present, else the oracle port.  Test / measurement infrastructure: never on the product path."""

    def __init__(self, model_name):
        import numpy as np
        import torch

        from oracle import ref_shim

        self.cfg, w, self.wdesc = load_weights(model_name)
        self.n_q = self.cfg.num_quantizers
        self.kind = "reference" if ref_shim.deploy_available() else "port"
        if self.kind == "reference":
            self.model = ref_shim.build_reference_model(w, self.n_q)
            self.what = ("the reference's own models/hilcodec/streaming.py classes (unmodified copy, oracle/_ref), "
                         f"{self.wdesc}, torch {torch.__version__} CPU")
        else:
            from oracle import hilcodec_oracle as O

            self.O = O
            self.p = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}
            self.ocfg = O.CodecConfig(num_quantizers=self.n_q)
            self.what = f"oracle port of the reference's deployment graph, {self.wdesc}, torch {torch.__version__} CPU"

    def one_shot(self, x, n):
        import torch

        from oracle import ref_shim

        with torch.no_grad():
            if self.kind == "reference":
                return ref_shim.reference_forward(self.model, x, n)
            return self.O.codec_forward(self.ocfg, self.p, x, n)

    def stream(self, x, n, hop, frames):
        """Frame by frame, caches as lists of tensors: notebook cell 3 / test_onnx.py:75-135.  Returns seconds."""
        import torch

        from oracle import ref_shim

        with torch.no_grad():
            if self.kind == "reference":
                ce, cd = self.model.initialize_cache(x[:, :, :hop])
            else:
                ce = self.O.zero_caches(self.O.encoder_cache_shapes(self.ocfg, x.shape[0]))
                cd = self.O.zero_caches(self.O.decoder_cache_shapes(self.ocfg, x.shape[0]))
            t0 = time.perf_counter()
            for f in range(frames):
                chunk = x[:, :, f * hop:(f + 1) * hop]
                if self.kind == "reference":
                    r = ref_shim.reference_forward(self.model, chunk, n, ce, cd)
                else:
                    r = self.O.codec_forward(self.ocfg, self.p, chunk, n, ce, cd)
                ce, cd = r["enc_caches"], r["dec_caches"]
            return time.perf_counter() - t0


def cpu_reference_run(model_name, n_q, samples, steps, warmup, budget_s, clips=None):
    """One-shot batches on all host threads, on a bounded sample of the workload (`clips`: the workload's batch size;
    sub-batches larger than it are not tried, so configs[0] is timed as the single clip it is)."""
    import torch

    cores = host_cores()
    torch.set_num_threads(cores)
    codec = CpuCodec(model_name)

    def run(x):
        t0 = time.perf_counter()
        codec.one_shot(x, n_q)
        return time.perf_counter() - t0

    run(synth(1, samples, 99))                      # page-in / thread pool start
    # the CPU path's throughput depends on the batch it is given (allocation / cache effects):
    # probe a few sub-batch sizes and give the reference its best one
    best_b, best_rate = 1, 0.0
    for b in (1, 2, 4, 8):
        if clips is not None and b > clips:
            break
        t = run(synth(b, samples, 90 + b))
        if b / t > best_rate:
            best_b, best_rate = b, b / t
    total_steps = max(1, steps + warmup)
    reps = int(max(1, min((clips or 256) // best_b, budget_s / total_steps * best_rate / best_b)))
    batch = reps * best_b
    xs = [synth(best_b, samples, 1234 + i) for i in range(reps)]

    def step():
        return sum(run(x) for x in xs)

    for _ in range(warmup):
        step()
    times = [step() for _ in range(max(1, steps))]
    frames = batch * (samples // codec.cfg.hop)
    sec = sum(times) / len(times)
    return {
        "value": frames / sec, "unit": "frames/s", "cores": cores, "kind": codec.kind,
        "sample": f"{reps} x {best_b} clips x {samples} samples of the workload per step (sub-batch {best_b} = fastest "
                  f"of 1/2/4/8), {len(times)} timed steps ({sec:.2f} s/step), {cores} threads; {codec.what}",
        "ms_per_step": sec * 1e3, "batch": batch,
    }


def cpu_stream_run(model_name, n_q, streams, hop, budget_s, threads_list=None):
    """The reference's frame-by-frame protocol on the host: 1 thread (its published setting) and all cores."""
    import torch

    cores = host_cores()
    codec = CpuCodec(model_name)
    out = {}
    for threads in (threads_list or [1, cores]):
        torch.set_num_threads(threads)
        x = synth(streams, hop * 12, 77)
        t = codec.stream(x, n_q, hop, 2)                         # warm-up, and a first estimate of a frame's cost
        per = max(t / 2, 1e-4)
        frames = int(max(2, min(75, budget_s / 2 / per)))
        x = synth(streams, hop * frames, 1234)
        sec = codec.stream(x, n_q, hop, frames)
        out[threads] = {"value": streams * frames / sec, "unit": "frames/s", "cores": threads, "kind": codec.kind,
                        "ms_per_frame": sec / frames * 1e3, "x_realtime_per_stream": (hop / 24000.0) / (sec / frames),
                        "sample": f"{frames} sequential hops of {streams} stream(s), one encoder / quantizer / dequantizer / "
                                  f"decoder call per hop with the caches handed back as tensors, {threads} thread(s); "
                                  f"{codec.what}"}
    torch.set_num_threads(cores)
    best = out[cores] if cores in out else out[max(out)]
    r = dict(best)
    if 1 in out and cores != 1:
        r["one_thread"] = out[1]     # the reference's own published protocol (RTF 0.41 on the authors' server)
    return r


def reference_on_gpu_run(ctx, model_name, n_q, samples, clips, reps=5):
    """Second, stated baseline: the reference's OWN classes moved to the same B200 with
    `.cuda()`, i.e. stock PyTorch -> cuDNN / cuBLAS in fp32 (TF32 off, as the parity bars need), one-shot on a sub-batch
    of the workload.  Not the `--impl reference` arm (that one is the CPU path BASELINE.json names)."""
    import torch

    from oracle import ref_shim

    if not ref_shim.deploy_available():
        return None
    cfg, w, wdesc = load_weights(model_name)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = ref_shim.build_reference_model(w, n_q).to(ctx.dev)
    x = synth(clips, samples, 1234).to(ctx.dev)
    for _ in range(2):
        ref_shim.reference_forward(model, x, n_q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ref_shim.reference_forward(model, x, n_q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    del model, x
    torch.cuda.empty_cache()
    return {"value": clips * (samples // cfg.hop) / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "clips": clips,
            "what": f"the reference's streaming.py classes .cuda() (torch {torch.__version__}, cuDNN / cuBLAS fp32, TF32 off), "
                    f"{clips} x {samples} one-shot, {reps} timed calls, {wdesc}"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload or "music256"
    model_name, clips, samples, n_q, desc = WORKLOADS[wl]
    if wl.startswith("stream"):
        hop = 320
        r = cpu_stream_run(model_name, n_q, clips, hop, budget_s=min(150.0, 12.0 * max(1, args.steps)))
        ms_step = r["ms_per_frame"] * (samples // hop)
        cfgd = {"workload": desc, "model": model_name, "streams": clips, "hop": hop, "n_q": n_q,
                "note": "CPU: bounded number of sequential hops, scaled to a 75-hop step"}
    else:
        r = cpu_reference_run(model_name, n_q, samples, args.steps, args.warmup, budget_s=150.0, clips=clips)
        ms_step = r["ms_per_step"]
        cfgd = {"workload": desc, "model": model_name, "clips_per_step": r["batch"], "samples": samples, "n_q": n_q,
                "note": "CPU: bounded sample of the workload per step"}
    line = {
        "impl": "reference", "metric": "audio frames/sec (24 kHz enc+RVQ+dec)", "value": r["value"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfgd,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "one_thread" in r:
        line["cpu_baseline"]["one_thread"] = {k: r["one_thread"][k] for k in ("value", "unit", "cores", "ms_per_frame",
                                                                             "x_realtime_per_stream")}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- CUDA arm: shared set-up
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (there is no CPU path in hilcodec_b200)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from hilcodec_b200 import _lib

        self._lib = _lib
        self.lib = _lib.load()
        self.peaks, self.peak_src = read_peaks()
        self.hbm_peak = float(self.peaks.get("hbm_gbs", 6650.0))
        self.bf16_peak = float(self.peaks.get("bf16_tflops_sustained", self.peaks.get("bf16_tflops", 1590.0)))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def profile_categories(ctx, step, prof_steps):
    """Per-category kernel timing: a separate pass with CUDA events around every launch (hil_profile_begin / end)."""
    n_cat = len(CATEGORIES)
    arr_ms, arr_fl, arr_by = (C.c_double * n_cat)(), (C.c_double * n_cat)(), (C.c_double * n_cat)()
    arr_n = (C.c_int64 * n_cat)()
    ctx.torch.cuda.synchronize()
    ctx._lib.check(ctx.lib.hil_profile_begin())
    for _ in range(prof_steps):
        step()
    ctx._lib.check(ctx.lib.hil_profile_end(arr_ms, arr_fl, arr_by, arr_n, n_cat))
    cats = {}
    for i, name in enumerate(CATEGORIES):
        if arr_n[i]:
            cats[name] = {"ms_per_step": arr_ms[i] / prof_steps, "launches_per_step": arr_n[i] // prof_steps,
                          "tflops": arr_fl[i] / (arr_ms[i] * 1e-3) / 1e12 if arr_ms[i] > 0 else 0.0,
                          "gbs": arr_by[i] / (arr_ms[i] * 1e-3) / 1e9 if arr_ms[i] > 0 else 0.0,
                          "algorithmic_bytes_per_step": arr_by[i] / prof_steps}
    dump = os.environ.get("HILCODEC_DUMP_LAUNCHES")
    if dump and ctx.rank == 0:   # launch-by-launch manifest of the profiled steps, for tools/summarize_launches.py
        cap = 1 << 16
        cat, ms, fl, by = (C.c_int32 * cap)(), (C.c_double * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
        n = ctx.lib.hil_profile_launches(cat, ms, fl, by, cap)
        per = n // prof_steps
        with open(dump, "w") as f:
            json.dump({"launches_per_step": per,
                       "launches": [{"cat": CATEGORIES[cat[i]], "ms": ms[i], "flops": fl[i], "bytes": by[i]}
                                    for i in range(n - per, n)]}, f)
    return cats


def traffic_by_class():
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic_by_class.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def class_roofline(ctx, name, c, traffic, workload):
    bound, kernel = GEMM_CLASSES[name]
    frac_hbm = c["gbs"] / ctx.hbm_peak
    frac_tensor = c["tflops"] / ctx.bf16_peak
    t = traffic.get(workload, {}).get(name) if traffic else None
    per_launch = c["algorithmic_bytes_per_step"] / max(c["launches_per_step"], 1)
    r = {"bound": bound, "kernel": kernel, "launches_per_step": c["launches_per_step"], "ms_per_step": c["ms_per_step"],
         "achieved": c["gbs"] if bound == "hbm" else c["tflops"], "peak": ctx.hbm_peak if bound == "hbm" else ctx.bf16_peak,
         "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": frac_hbm if bound == "hbm" else frac_tensor,
         "algorithmic_bytes_per_launch": per_launch,
         "traffic": t["dram_bytes_per_launch_avg"] if t else None,
         "hbm_frac": frac_hbm, "tensor_frac_of_bf16_peak": frac_tensor,
         "tensor_frac_of_reachable": frac_tensor * MMA_PER_PRODUCT}
    if t:
        r["traffic_source"] = t.get("source")
        r["traffic_over_algorithmic"] = t["dram_bytes_per_launch_avg"] / per_launch if per_launch else None
    return r


# ----------------------------------------------------------------------------- CUDA arm, streaming workloads
def run_stream(ctx, wl, steps, warmup, with_cpu, cpu_budget=None):
    """BASELINE configs[3]: frame-by-frame streaming with GPU-resident caches.  One step = 75 sequential hops (1 s of
    audio) of every stream; a hop is one CUDA-graph replay (`hil_codec_forward_graph`).  Streams of different GPUs are
    independent: `--gpus N` runs the same thing per rank and adds the rates up."""
    torch, lib, _lib = ctx.torch, ctx.lib, ctx._lib
    from hilcodec_b200 import sharding
    from hilcodec_b200 import streaming as S

    dev = ctx.dev
    model_name, B, samples, n_q, desc = WORKLOADS[wl]
    cfg, w, wdesc = load_weights(model_name)
    model = S.HILCodec.from_weights(w, cfg.num_quantizers).cuda()
    hop, hops = cfg.hop, samples // cfg.hop
    warm = max(warmup, 3)
    x_host = synth(B, samples * (steps + warm + 2), 1234 + ctx.rank).pin_memory()
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
    sampler = ClockSampler(ctx.local).start() if ctx.rank == 0 else None
    ctx.barrier()
    torch.cuda.synchronize()
    l0 = lib.hil_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ctx.barrier()
    launches = (lib.hil_launch_count() - l0) // max(steps * hops, 1)
    ms = e0.elapsed_time(e1)

    # e2e: every hop comes from pinned host memory and its indices + PCM go back to the host (C-ABI host call: H2D,
    # graph replay, D2H, stream sync -- the latency a caller sees per hop)
    idx_host = torch.empty(n_q, B, 1, dtype=torch.int64).pin_memory()
    y_host = torch.empty(B, 1, hop, dtype=torch.float32).pin_memory()
    chunk_host = torch.empty(B, 1, hop, dtype=torch.float32).pin_memory()
    checksum, hpos = 0.0, 0
    for _ in range(3):   # the host call's own staging buffers / graph key
        chunk_host.copy_(x_host[:, :, hpos:hpos + hop])
        _lib.check(lib.hil_codec_forward_host(hmodel, hstate, chunk_host.data_ptr(), B, hop, n_q, idx_host.data_ptr(),
                                              y_host.data_ptr(), sp))
        hpos += hop
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps * hops):
        chunk_host.copy_(x_host[:, :, hpos:hpos + hop])
        _lib.check(lib.hil_codec_forward_host(hmodel, hstate, chunk_host.data_ptr(), B, hop, n_q, idx_host.data_ptr(),
                                              y_host.data_ptr(), sp))
        checksum += float(y_host[0, 0, 0])
        hpos += hop
    ms_e2e = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    # where a hop goes (eager launches, events around each: shares matter, the sum carries ~2 us per launch of overhead)
    st2 = core.state(dev, B)

    def eager_hop():
        _lib.check(lib.hil_codec_forward(hmodel, st2, xin.data_ptr(), B, hop, n_q, None, idx.data_ptr(), y.data_ptr(), sp))

    cats = profile_categories(ctx, eager_hop, 20)
    ms, ms_e2e = sharding.reduce_max([ms, ms_e2e], device=dev)
    if ctx.rank != 0:
        return None
    frames = B * hops * ctx.world
    weight_bytes = float(sum(v.size * 4 for v in w.values()))
    ms_hop = ms / (steps * hops)
    line = {
        "metric": "audio frames/sec (24 kHz enc+RVQ+dec)", "value": frames * steps / (ms * 1e-3), "unit": "frames/s",
        "n_gpus": ctx.world, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": f"synthetic 0.1*randn audio, {wdesc}",
        "config": {"workload": desc, "model": model_name, "streams_per_gpu": B, "hop": hop, "hops_per_step": hops,
                   "n_q": n_q, "executor": "hil_codec_forward_graph (one CUDA-graph replay per hop)",
                   "l2": "no flush: a hop re-reads the ~50 MB of weights, which fit the 126 MB L2, by design",
                   "deviation": "hop 320, not BASELINE's 300: 300 is AudioDec's hop and cannot be fed to HILCodec (BASELINE.md 2)"},
        "ms_per_hop": ms_hop, "x_realtime_per_stream": (hop / 24000.0) / (ms_hop * 1e-3),
        "e2e": {"value": frames * steps / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": B * hop * 4 * hops,
                "d2h_bytes_per_step": (n_q * B * 8 + B * hop * 4) * hops, "ms_per_step": ms_e2e / steps,
                "ms_per_hop": ms_e2e / (steps * hops),
                "api": "hil_codec_forward_host per hop (pinned host buffers, graph replay, stream sync)", "checksum": checksum},
        "gpu_launches": int(launches) * hops, "gpu_launches_per_hop": int(launches),
        "clocks": clocks,
        # what a hop cannot avoid is reading every weight once; everything above that is launch / dependency latency
        "roofline": {"bound": "hbm", "kernel": f"whole hop: a chain of {int(launches)} dependent launches; weights (L2-resident) read once per hop",
                     "achieved": weight_bytes / (ms_hop * 1e-3) / 1e9, "peak": ctx.hbm_peak, "unit": "GB/s",
                     "frac": weight_bytes / (ms_hop * 1e-3) / 1e9 / ctx.hbm_peak, "traffic": None,
                     "algorithmic_bytes_per_hop": weight_bytes},
        "kernel_categories_per_hop": {k: {"ms": v["ms_per_step"], "launches": v["launches_per_step"]} for k, v in cats.items()},
    }
    if with_cpu and ctx.world == 1:
        r = cpu_stream_run(model_name, n_q, B, hop, budget_s=cpu_budget or (16.0 if B == 1 else 24.0))
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "ms_per_frame")}
        if "one_thread" in r:
            line["cpu_baseline"]["one_thread"] = {k: r["one_thread"][k] for k in ("value", "unit", "cores", "ms_per_frame",
                                                                                 "x_realtime_per_stream", "sample")}
    return line


# ----------------------------------------------------------------------------- CUDA arm, one-shot batches
def run_batch(ctx, wl, steps, warmup, with_cpu, cpu_budget=None):
    torch, lib, _lib = ctx.torch, ctx.lib, ctx._lib
    from hilcodec_b200 import sharding
    from hilcodec_b200 import streaming as S

    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    model_name, clips, samples, n_q, desc = WORKLOADS[wl]
    cfg, w, wdesc = load_weights(model_name)
    model = S.HILCodec.from_weights(w, cfg.num_quantizers).cuda()
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
    warm = max(warmup, 3)

    def step():
        _lib.check(lib.hil_state_reset(hstate, sp))
        _lib.check(lib.hil_codec_forward(hmodel, hstate, x.data_ptr(), B, T, n_q, None, idx.data_ptr(), y.data_ptr(), sp))

    for _ in range(warm):
        step()
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(ctx.local).start() if rank == 0 else None
    ctx.barrier()
    torch.cuda.synchronize()
    l0 = lib.hil_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ctx.barrier()
    launches = (lib.hil_launch_count() - l0) // max(steps, 1)
    ms = e0.elapsed_time(e1)

    # ---- e2e: host buffers through the C-ABI call, copies inside the timed region
    idx_host = torch.empty(n_q, B, F, dtype=torch.int64).pin_memory()
    y_host = torch.empty(B, 1, T, dtype=torch.float32).pin_memory()

    def step_host():
        _lib.check(lib.hil_state_reset(hstate, sp))
        _lib.check(lib.hil_codec_forward_host(hmodel, hstate, x_host.data_ptr(), B, T, n_q, idx_host.data_ptr(),
                                              y_host.data_ptr(), sp))

    for _ in range(2 if B * T > 64 * 3200 else 5):   # small batches replay CUDA graphs: two keys (cache generations) to capture first
        step_host()
    ctx.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    ctx.barrier()
    clocks = sampler.stop() if sampler else None
    e2e_checksum = float(y_host.abs().sum())  # the device->host result is really read on the host
    ms, ms_e2e = sharding.reduce_max([ms, ms_e2e], device=dev)  # a multi-GPU step is as slow as its slowest rank

    # ---- config 5 "via NCCL": all-gather of every rank's indices + PCM (outside the step: the path itself has no
    # collective; this is what collecting the results of the batch split costs)
    gather = None
    if world > 1:
        placeholder = torch.empty(world * B, 1, 0, device=dev)   # forward_sharded only needs the full batch size

        def gather_once():
            return sharding.forward_sharded(lambda xs: (idx, y), placeholder, gather=True)

        gather_once()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        g0.record(stream)
        for _ in range(5):
            gi, gy = gather_once()
        g1.record(stream)
        torch.cuda.synchronize()
        gms = sharding.reduce_max([g0.elapsed_time(g1) / 5], device=dev)[0]
        by = world * (idx.numel() * 8 + y.numel() * 4)
        gather = {"ms": gms, "bytes_received_per_rank": by, "gbs_per_rank": by / (gms * 1e-3) / 1e9,
                  "what": "torch.distributed.all_gather (NCCL) of idx [n,B,F] int64 + wav [B,1,T] fp32 from every rank, "
                          "after the step; not inside `value`", "shape_ok": bool(gi.shape[1] == world * B and gy.shape[0] == world * B)}

    cats = profile_categories(ctx, step, 2)
    if rank != 0:
        return None

    traffic = traffic_by_class()
    total_ms = sum(c["ms_per_step"] for c in cats.values()) or 1.0
    by_class = {k: class_roofline(ctx, k, cats[k], traffic, wl) for k in GEMM_CLASSES if k in cats}
    for k, r in by_class.items():
        r["share_of_step"] = cats[k]["ms_per_step"] / total_ms
    dominant = max(by_class, key=lambda k: by_class[k]["ms_per_step"]) if by_class else None
    frames_total = B * F * world
    value = frames_total * steps / (ms * 1e-3)
    e2e_value = frames_total * steps / (ms_e2e * 1e-3)
    frames_rank = B * F
    step_s = ms / steps * 1e-3
    roofline = dict(by_class[dominant]) if dominant else {}
    roofline.update({
        "dominant_class": dominant,
        "peak_source": f"{ctx.peak_src}: hbm_gbs / bf16_tflops_sustained (kernels timed inside a long step)",
        "bytes": "algorithmic: every launch reads its inputs once and writes its outputs once (codec.cu HIL_LAUNCH)",
        "note": "fp32-accurate arithmetic is required for bit-exact VQ indices: every product is 3 fp16 MMAs (hi*hi, "
                "hi*lo, lo*hi) into two fp32 TMEM accumulators, so the reachable tensor ceiling is the bf16 peak / 3",
    })
    line = {
        "metric": "audio frames/sec (24 kHz enc+RVQ+dec)", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": f"synthetic 0.1*randn audio, {wdesc}",
        "config": {"workload": desc, "model": model_name, "clips_per_gpu": B, "samples": T, "n_q": n_q,
                   "sharding": f"batch-sharded x{world}, no data-path collective",
                   "l2": ("no explicit flush: one step streams >10 GB of activations through the 126 MB L2" if B >= 32 else
                          "no flush: a single clip's activations and the ~50 MB of weights fit the 126 MB L2, as they would "
                          "in a service that keeps the model loaded")},
        "rtf_x_realtime": value / 75.0,
        "model_tflops": value * FLOP_PER_FRAME[model_name] / 1e12,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * T * 4,
                "d2h_bytes_per_step": n_q * B * F * 8 + B * T * 4, "ms_per_step": ms_e2e / steps,
                "api": "hil_codec_forward_host (pinned host buffers)", "checksum": e2e_checksum},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_by_class": by_class,
        # the whole step against both rooflines, with SURVEY 8(d)'s two byte models next to the per-launch bytes
        "step_roofline": {
            "tensor_frac_of_reachable": frames_rank * FLOP_PER_FRAME[model_name] / step_s / 1e12 * MMA_PER_PRODUCT / ctx.bf16_peak,
            "hbm_frac_fused_unit_bytes": frames_rank * FUSED_UNIT_BYTES_PER_FRAME / step_s / 1e9 / ctx.hbm_peak,
            "hbm_frac_layer_boundary_bytes": frames_rank * LAYER_BOUNDARY_BYTES_PER_FRAME / step_s / 1e9 / ctx.hbm_peak,
            "hbm_frac_per_launch_bytes": sum(c["algorithmic_bytes_per_step"] for c in cats.values()) / step_s / 1e9 / ctx.hbm_peak,
            "bytes_per_frame": {"fused_unit_model": FUSED_UNIT_BYTES_PER_FRAME, "layer_boundary_model": LAYER_BOUNDARY_BYTES_PER_FRAME,
                                "per_launch_algorithmic": sum(c["algorithmic_bytes_per_step"] for c in cats.values()) / frames_rank},
        },
        "kernel_categories": {k: {kk: vv for kk, vv in v.items() if kk != "algorithmic_bytes_per_step"} for k, v in cats.items()},
    }
    if gather:
        line["gather"] = gather
    if world == 1 and with_cpu:
        r = cpu_reference_run(model_name, n_q, samples, steps=4, warmup=1, budget_s=cpu_budget or 40.0, clips=B)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        try:
            g = reference_on_gpu_run(ctx, model_name, n_q, samples, clips=min(B, 32))
            if g:
                line["reference_on_same_gpu"] = g
        except Exception as e:   # informative extra: never takes the line down
            line["reference_on_same_gpu"] = {"error": f"{type(e).__name__}: {e}"}
    return line


def condensed(line):
    keep = ("value", "unit", "ms_per_step", "ms_per_hop", "x_realtime_per_stream", "rtf_x_realtime", "e2e", "roofline",
            "cpu_baseline", "reference_on_same_gpu", "gpu_launches", "gpu_launches_per_hop", "config", "data", "steps",
            "warmup")
    return {k: line[k] for k in keep if k in line}


def main_ours(args):
    ctx = Ctx()
    wl = args.workload or "music256"
    with_cpu = not args.no_cpu_baseline
    run = run_stream if wl.startswith("stream") else run_batch
    line = run(ctx, wl, args.steps, args.warmup, with_cpu)
    if line is not None and args.workload is None and ctx.world == 1 and not args.no_other_workloads:
        # the other BASELINE configs that fit one GPU, in the same contract (shorter runs: they are not the headline)
        others = {}
        for name, st in (("speech64", 10), ("speech1", 10), ("stream1", 4), ("stream64", 4)):
            try:
                r = (run_stream if name.startswith("stream") else run_batch)(ctx, name, st, 3, with_cpu, cpu_budget=14.0)
                others[name] = condensed(r)
            except Exception as e:  # a secondary workload must not take the headline line down with it
                others[name] = {"error": f"{type(e).__name__}: {e}"}
        line["other_workloads"] = others
    if line is not None:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
