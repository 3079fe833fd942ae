#!/usr/bin/env python
"""Benchmark of the hot path: A-scans/sec of batched PAUT signal-model inference on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model msc]

A step = one pass of forward + post-processing over one resident synthetic volume
(default: BASELINE.json configs[1], MultiSignalClassifier over 1 000 200 A-scans = 3334 sets x 300 x 320,
bf16 in HBM).  N > 1 is launched by torchrun, one rank per GPU; every rank owns its own shard of the same
size (sets are independent: no data-path collective, weak scaling) and the time is the max over ranks.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of each key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SETS = {"msc": (3334, 300), "msc_n": (3334, 300), "conv1d_msc": (3334, 300),
        "ssd": (20000, 50), "enhanced": (20000, 50), "two_stage": (20000, 50),
        # SURVEY section 8 "next" rows (f2 / f3): same volume as the MSC config
        "msc_legacy": (3334, 300), "improved": (3334, 300), "hybrid": (3334, 300), "complex": (3334, 300)}
S = 320

# Algorithmic work per A-scan of the fused kernels: (FLOPs = 2*MAC, HBM bytes, bound).
# "bound" is the roofline the kernel is judged against: its algorithmic intensity is above the ridge
# (~210 FLOP/B with the measured peaks) for the tensor-bound ones.
def kernel_work(kind, n_per):
    if kind in ("msc", "msc_n"):
        attn = 2 * 64 * 192 + 2 * 2 * n_per * 64 + 2 * 64 * 64
        return {
            "msc_encoder_tc": (2 * (8 * 3 + 16 * 24) * S + 2 * S * 128 + 2 * 128 * 64, 2 * S + 4 * 64, "tensor"),
            "msc_attn_block": (attn, 2 * 4 * 64, "tensor"),
            "msc_ffn_head": (2 * 2 * 64 * 32 + 2 * 64 * 3, 4 * 64 + 12, "hbm"),
            "msc_front": (2 * (8 * 3 + 16 * 24) * S, 2 * 4 * S, "hbm"),
        }
    if kind in ("msc_legacy", "improved"):
        return {}                                   # CUDA-core front end + library-shaped linears: no fused-kernel figure yet
    conv = {"hybrid": 2 * (32 * 64 * 3 + 64 * 64 * 5) * S, "complex": 2 * (32 * 64 * 7 + 64 * 64 * 15) * S,
            "two_stage": 2 * 32 * 32 * (3 + 5 + 7 + 11) * S, "ssd": 2 * (64 * 128 * 5 + 128 * 256 * 3) * S,
            "conv1d_msc": 2 * (64 * 128 * 3 + 128 * 128) * S,
            # branches + combine + six residual convs at full length; the stride-2 pyramid at S/2 and S/4 outputs
            "enhanced": 2 * (4 * 64 * 32 * 3 + 128 * 128 + 6 * 128 * 128 * 3) * S +
                        2 * (128 * 256 * 3 * (S // 2) + 256 * 256 * 3 * (S // 4))}
    return {"conv_tc": (conv[kind], 0, "tensor")}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
# (profiles/), per A-scan; None until a capture of the current kernel exists
NCU_TRAFFIC_PER_ASCAN = {"msc_encoder_tc": 827.0}   # profiles/r01_s3/ncu_msc_encoder_tc_team_handoff.txt: (230.6 + 67.1) MB / 360 000 A-scans


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    sm_max=d.get("sm_max_mhz"), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, sm_max=1965.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a thread
    (nvidia-smi -lms cannot sample a 30 ms region); falls back to one nvidia-smi query."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 5 ms poll"}
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(v) for v in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi after the region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}


def make_volume(kind, n_sets, n_per_set, seed):
    from defectdetection_viaobjectdetection_b200 import synthetic as synth
    x = synth.synth_paut_sets(n_sets, n_per_set, S, seed=seed, defect_frac=0.01)     # [sets, N, S] fp32 in [0,1]
    if kind == "conv1d_msc":
        x = np.ascontiguousarray(x.transpose(0, 2, 1))
    return torch.from_numpy(x)


def oracle_forward(kind, sd, x, threshold=0.5):
    # the ONLY use of oracle/ in this file: the CPU legs (cpu_baseline and --impl reference)
    from oracle import models as om
    from oracle import postprocess as opp
    with torch.no_grad():
        out = om.FORWARD[kind](sd, x)
    return opp.postprocess(kind, out, threshold, S)


def cpu_reference_rate(kind, sd, x_host_f32, budget_s, n_per_set):
    """Oracle (the CPU port of the reference path, torch fp32 on all host threads) on a bounded sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    chunk = 8
    n_sets = x_host_f32.shape[0]
    oracle_forward(kind, sd, x_host_f32[:chunk])                       # warm-up
    done, t0, i = 0, time.perf_counter(), 0
    while True:
        lo = (i * chunk) % max(n_sets - chunk, 1)
        oracle_forward(kind, sd, x_host_f32[lo:lo + chunk])
        done += chunk * n_per_set
        i += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return done / el, done, el


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="msc", choices=sorted(SETS))
    ap.add_argument("--sets", type=int, default=0, help="sets per GPU (default: the BASELINE config)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: at least 3 warm-up steps"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind = args.model
    n_sets, n_per = SETS[kind]
    if args.sets:
        n_sets = args.sets
    from defectdetection_viaobjectdetection_b200 import synthetic as synth      # data + weights: not oracle code
    sd = synth.synth_state_dict(kind, seed=0)
    workload = f"{kind} over {n_sets * n_per} A-scans per GPU ({n_sets} sets x {n_per} x {S}), synthetic PAUT volume"

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        sample_sets = 64
        x = make_volume(kind, sample_sets, n_per, seed=42)
        per_step = max(2.0, min(20.0, 90.0 / max(args.steps + args.warmup, 1)))
        for _ in range(args.warmup):
            cpu_reference_rate(kind, sd, x, 0.5, n_per)
        t0, done = time.perf_counter(), 0
        for _ in range(args.steps):
            _, d, _ = cpu_reference_rate(kind, sd, x, per_step, n_per)
            done += d
        el = time.perf_counter() - t0
        rate = done / el
        cores = torch.get_num_threads()
        print(json.dumps({
            "impl": "reference", "metric": "A-scans/sec", "value": rate, "unit": "A-scans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "l2": "sample larger than L2 is not relevant on the CPU arm"},
            "cpu_baseline": {"value": rate, "unit": "A-scans/s", "cores": cores, "kind": "port",
                             "sample": f"{sample_sets} sets x {n_per} A-scans cycled for {per_step:.0f} s per step, "
                                       "oracle port of the reference forward + predict post-processing, fp32"},
            "e2e": {"value": rate, "unit": "A-scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm
    import defectdetection_viaobjectdetection_b200 as paut
    from defectdetection_viaobjectdetection_b200 import runtime
    from defectdetection_viaobjectdetection_b200.modules import FACTORIES as MODELS

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    model = MODELS[kind](dict(signal_length=S))
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    model.precision = args.precision
    in_dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32

    x_host = make_volume(kind, n_sets, n_per, seed=42 + rank).to(in_dtype).pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    ctx = paut.get_context(dev)
    n_ascans = n_sets * n_per

    def step_resident():
        native, (outs, struct, (B, N, S_)) = model._run(x_dev)
        return native.postprocess(struct, B, N, S_, 0.5, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        det, count = step_resident()     # same reference pattern as the timed loop: the caching allocator reaches its
    barrier()                            # steady state (two record buffers alive) before the timed region
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = ctx.launch_count
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        det, count = step_resident()
    e1.record()
    barrier()
    prof = ctx.profile_end()
    launches = ctx.launch_count - launches0
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    n_found = int(count.item())

    # ---- end to end through the public API with host buffers: H2D + forward + post-process + D2H records
    # (VolumeScanner: resident chunks on two streams, H2D / kernels / D2H of kept records overlapped)
    from defectdetection_viaobjectdetection_b200.streaming import VolumeScanner
    scan_chunk = int(os.environ.get("PAUT_BENCH_CHUNK_ASCANS", "38400"))
    scanner = VolumeScanner(model, chunk_sets=max(1, (scan_chunk // n_per)), lanes=int(os.environ.get("PAUT_BENCH_LANES", "4")),
                            reuse_output=True)

    def step_e2e():
        return scanner.scan(x_host, threshold=0.5)

    for _ in range(3):                   # warm-up: lanes, both record arenas, allocator steady state
        rec = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rec = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        value = world * n_ascans * args.steps / (ms * 1e-3)
        e2e = world * n_ascans * args.steps / e2e_s
        # dominant kernel and its roofline
        total_ms = sum(v[1] for v in prof.values()) or 1.0
        top = max(prof.items(), key=lambda kv: kv[1][1])
        name, (n_launch, k_ms) = top
        work = kernel_work(kind, n_per).get(name)
        roof = {"kernel": name, "share_of_step": k_ms / total_ms, "launches": n_launch,
                "avg_launch_ms": k_ms / n_launch, "peak_source": pk["source"]}
        if work:
            flops, nbytes, bound = work
            per_launch_ascans = n_ascans * args.steps / n_launch
            t = (k_ms / n_launch) * 1e-3
            if bound == "tensor":
                ach = flops * per_launch_ascans / t / 1e12
                roof.update(bound="tensor", achieved=ach, peak=pk["tf_sust"], unit="TFLOP/s", frac=ach / pk["tf_sust"],
                            hbm_gbs=nbytes * per_launch_ascans / t / 1e9, flop_per_ascan=flops, bytes_per_ascan=nbytes)
            else:
                ach = nbytes * per_launch_ascans / t / 1e9
                roof.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                            flop_per_ascan=flops, bytes_per_ascan=nbytes)
            tr = NCU_TRAFFIC_PER_ASCAN.get(name)
            roof["traffic"] = tr * per_launch_ascans if tr else None
        else:
            roof.update(bound="hbm", achieved=None, peak=pk["hbm"], unit="GB/s", frac=None)
        roof.setdefault("traffic", None)
        line = {
            "metric": "A-scans/sec", "value": value, "unit": "A-scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": workload, "l2": "input volume (%.0f MB) larger than the 126 MB L2" %
                       (x_dev.numel() * x_dev.element_size() / 1e6), "sharding": f"dp{world} by scan position",
                       "threshold": 0.5, "detections_last_step": n_found},
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "A-scans/s", "h2d_bytes_per_step": scanner.h2d_bytes,
                    "d2h_bytes_per_step": scanner.d2h_bytes, "api": "VolumeScanner(reuse_output=True).scan(pinned host tensor)"},
            "roofline": roof,
            "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
        }
        if world == 1 and args.cpu_seconds > 0:
            x_cpu = x_host[:64].float()
            rate, done, el = cpu_reference_rate(kind, sd, x_cpu, args.cpu_seconds, n_per)
            line["cpu_baseline"] = {"value": rate, "unit": "A-scans/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{done} A-scans (64 sets cycled) in {el:.1f} s, oracle port of the "
                                              "reference forward + predict post-processing, fp32, all host threads"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
