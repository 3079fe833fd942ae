#!/usr/bin/env python
"""Benchmark of the hot path: A-scans/sec of batched PAUT signal-model inference on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model msc] [--no-extra]

A step = one pass of forward + post-processing over one resident synthetic volume
(default: BASELINE.json configs[1], MultiSignalClassifier over 1 000 200 A-scans = 3334 sets x 300 x 320,
bf16 in HBM).  N > 1 is launched by torchrun, one rank per GPU; every rank owns its own shard of the same
size (sets are independent: no data-path collective, weak scaling) and the time is the max over ranks.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of each key.

Besides the contract keys the line carries
  roofline / roofline_2nd : the two kernels with the largest share of the step (algorithmic work / measured time)
  extra.models            : the other BASELINE configs on one GPU -- two_stage with predict (configs[3]),
                            enhanced (configs[2]), ssd -- each >= 5 timed steps after warm-up
  extra.strong_scaling    : the fixed 8 M-A-scan whole-weld volume of configs[4] sharded over the N ranks
  extra.gpu_eager_port    : the oracle port (plain torch ops) on the SAME GPU, fp32 and bf16 autocast -- the
                            "reference modules on cuda, eager" comparator of BASELINE.md 4.5
  e2e.fp32_host           : the end-to-end figure when the host volume is fp32 (what the reference's loaders hold)
  e2e.h2d_ceiling_gbs     : pinned cudaMemcpyAsync rate of this box at this N (tools/h2d_ceiling.py), next to the
                            rate the scanner achieved
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
STRONG_SETS = 26667          # configs[4]: 160 000 scan positions x 50 beams = 8 M A-scans = 26 667 MSC sets of 300


# Algorithmic work per A-scan of the fused kernels: (FLOPs = 2*MAC, HBM bytes, bound).
# "bound" is the roofline the kernel is judged against: its algorithmic intensity is above the ridge
# (~210 FLOP/B with the measured peaks) for the tensor-bound ones.
def kernel_work(kind, n_per):
    if kind in ("msc", "msc_n"):
        attn = 2 * 64 * 192 + 2 * 2 * n_per * 64 + 2 * 64 * 64
        return {
            "msc_encoder_tc": (2 * (8 * 3 + 16 * 24) * S + 2 * S * 128 + 2 * 128 * 64, 2 * S + 4 * 64, "tensor"),
            "msc_attn_block": (attn, 2 * 4 * 64, "tensor"),
            "msc_attn_tc": (attn, 2 * 4 * 64, "tensor"),
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
    work = {"conv_tc": (conv[kind], 0, "tensor")}
    if kind == "two_stage":
        # fused encoder: both convolutions of the four branches on tcgen05; x in (bf16) + pooled features out (fp32)
        work["ts_encoder"] = (conv[kind] + 2 * 32 * (3 + 5 + 7 + 11) * S, 2 * S + 4 * 128, "tensor")
    return work


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per A-scan of the fused kernels, from this round's committed
    `ncu --set full` captures (profiles/r02/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep files)."""
    p = os.path.join(ROOT, "profiles", "r02", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {}


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
    # oracle/ is used by the baseline legs only (cpu_baseline, --impl reference, extra.gpu_eager_port): as the thing
    # our path is compared WITH, never inside the product path
    from oracle import models as om
    from oracle import postprocess as opp
    with torch.no_grad():
        out = om.FORWARD[kind](sd, x)
    return opp.postprocess(kind, out, threshold, S)


_REF_MODELS = {}


def reference_forward(kind, sd, x, threshold=0.5):
    """The reference's own class (oracle/_ref, see oracle/build_ref.py) on the CPU: forward + its predict() loop
    (SSD family) or the callers' threshold loop (model_pred.py:82-85).  None when oracle/_ref is absent."""
    if kind not in _REF_MODELS:
        from oracle import ref_models
        m = ref_models.reference_model(kind, S)
        if m is not None:
            m.load_state_dict(sd, strict=True)
            m.eval()
        _REF_MODELS[kind] = m
    m = _REF_MODELS[kind]
    if m is None:
        return None
    with torch.no_grad():
        if hasattr(m, "predict"):
            return m.predict(x, threshold=threshold)
        out = m(x)
        prob = out[0] if isinstance(out, tuple) else out
        return [[i for i, p in enumerate(row) if p > threshold] for row in prob.tolist()]


def cpu_kind(kind, sd):
    return "reference" if reference_forward(kind, sd, torch.zeros(1, 4, S) if kind != "conv1d_msc" else torch.zeros(1, S, 4)) is not None else "port"


def cpu_reference_rate(kind, sd, x_host_f32, budget_s, n_per_set):
    """The reference classes (oracle/_ref) or, without them, the oracle port, in torch fp32 on all host threads, on a
    bounded sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    chunk = 8
    n_sets = x_host_f32.shape[0]
    use_ref = cpu_kind(kind, sd) == "reference"
    fwd = (lambda xs: reference_forward(kind, sd, xs)) if use_ref else (lambda xs: oracle_forward(kind, sd, xs))
    fwd(x_host_f32[:chunk])                                            # warm-up
    done, t0, i = 0, time.perf_counter(), 0
    while True:
        lo = (i * chunk) % max(n_sets - chunk, 1)
        fwd(x_host_f32[lo:lo + chunk])
        done += chunk * n_per_set
        i += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return done / el, done, el


def gpu_eager_port(kind, sd, x_dev_f32, n_per_set):
    """The oracle port (the reference forward restated with plain torch ops) on the GPU: eager fp32 and bf16
    autocast, forward only (the reference's predict() loop is a per-element .item() loop and would dominate)."""
    from oracle import models as om
    sd_dev = {k: v.to(x_dev_f32.device) for k, v in sd.items()}
    out = {}
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        def run():
            with torch.no_grad():
                if ctx is None:
                    return om.FORWARD[kind](sd_dev, x_dev_f32)
                with ctx:
                    return om.FORWARD[kind](sd_dev, x_dev_f32)
        try:
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                run()
            e1.record()
            torch.cuda.synchronize()
            out[name] = x_dev_f32.shape[0] * n_per_set * 5 / (e0.elapsed_time(e1) * 1e-3)
        except Exception as ex:                                       # e.g. an op autocast cannot run in bf16
            out[name] = f"failed: {type(ex).__name__}"
    return out


def pin_rank_to_cores(local_rank, local_world):
    """Disjoint host-core sets per rank (all GPUs of this pool report one NUMA node): the scanner's host threads
    and the first-touch of its pinned buffers then do not migrate between the ranks' copy loops."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if local_world > 1 and len(cores) >= 2 * local_world:
            per = len(cores) // local_world
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
            return per
    except Exception:
        pass
    return None


def h2d_ceiling(dev, nbytes, reps, barrier):
    """Pinned host -> device copy rate of this rank while all ranks copy at the same time (one cudaMemcpyAsync
    per repetition on a side stream)."""
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.zero_()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        dst.copy_(host, non_blocking=True)
    s.synchronize()
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(reps):
            dst.copy_(host, non_blocking=True)
    s.synchronize()
    dt = time.perf_counter() - t0
    return nbytes * reps / dt / 1e9


def roofline_entries(kind, n_per, prof, n_ascans, steps, pk, traffic, top=2):
    total_ms = sum(v[1] for v in prof.values()) or 1.0
    out = []
    for name, (n_launch, k_ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:top]:
        work = kernel_work(kind, n_per).get(name)
        roof = {"kernel": name, "share_of_step": k_ms / total_ms, "launches": n_launch,
                "avg_launch_ms": k_ms / n_launch, "peak_source": pk["source"]}
        per_launch_ascans = n_ascans * steps / n_launch
        if work:
            flops, nbytes, bound = work
            t = (k_ms / n_launch) * 1e-3
            if bound == "tensor":
                ach = flops * per_launch_ascans / t / 1e12
                roof.update(bound="tensor", achieved=ach, peak=pk["tf_sust"], unit="TFLOP/s", frac=ach / pk["tf_sust"],
                            frac_of_burst_peak=ach / pk["tf_burst"], hbm_gbs=nbytes * per_launch_ascans / t / 1e9,
                            flop_per_ascan=flops, bytes_per_ascan=nbytes)
            else:
                ach = nbytes * per_launch_ascans / t / 1e9
                roof.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                            flop_per_ascan=flops, bytes_per_ascan=nbytes)
        else:
            roof.update(bound="hbm", achieved=None, peak=pk["hbm"], unit="GB/s", frac=None)
        tr = traffic.get(name)
        roof["traffic"] = tr["bytes_per_ascan"] * per_launch_ascans if tr else None
        if tr:
            roof["traffic_source"] = tr.get("source")
        out.append(roof)
    while len(out) < top:
        out.append(None)
    return out


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
    ap.add_argument("--no-extra", action="store_true", help="skip the extra models / strong scaling / comparators")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: at least 3 warm-up steps"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
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
            "cpu_baseline": {"value": rate, "unit": "A-scans/s", "cores": cores, "kind": cpu_kind(kind, sd),
                             "sample": f"{sample_sets} sets x {n_per} A-scans cycled for {per_step:.0f} s per step, "
                                       + ("the reference's own class (oracle/_ref) forward + its predict / threshold loop, fp32"
                                          if cpu_kind(kind, sd) == "reference" else
                                          "oracle port of the reference forward + predict post-processing, fp32")},
            "e2e": {"value": rate, "unit": "A-scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm
    cores_per_rank = pin_rank_to_cores(local_rank, local_world)      # before any pinned allocation (first touch)
    import defectdetection_viaobjectdetection_b200 as paut
    from defectdetection_viaobjectdetection_b200.modules import FACTORIES as MODELS
    from defectdetection_viaobjectdetection_b200.sharding import shard_range
    from defectdetection_viaobjectdetection_b200.streaming import VolumeScanner

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    in_dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32
    ctx = paut.get_context(dev)
    pk = peaks()
    traffic = ncu_traffic()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in t)

    def build(k):
        m = MODELS[k](dict(signal_length=S))
        m.load_state_dict(synth.synth_state_dict(k, seed=0), strict=True)
        m = m.to(dev).eval()
        m.precision = args.precision
        return m

    def time_resident(model, x_dev, steps, warmup):
        """forward + post-processing on a resident volume: CUDA events on the library's stream, profiling OFF;
        a second pass with one event per launch gives the per-kernel times."""
        def step():
            native, (outs, struct, (B, N, S_)) = model._run(x_dev)
            return native.postprocess(struct, B, N, S_, 0.5, dev)
        for _ in range(warmup):
            det, count = step()          # the caching allocator reaches its steady state before the timed region
        barrier()
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            det, count = step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count - launches0
        ctx.profile_begin()
        for _ in range(steps):
            det, count = step()
        torch.cuda.synchronize()
        prof = ctx.profile_end()
        return ms, launches, prof, int(count.item())

    model = build(kind)
    x_host = make_volume(kind, n_sets, n_per, seed=42 + rank).to(in_dtype).pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    n_ascans = n_sets * n_per

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms, launches, prof, n_found = time_resident(model, x_dev, args.steps, args.warmup)
    clk = clocks.stop()

    # ---- end to end through the public API with host buffers: H2D + forward + post-process + D2H records
    # (VolumeScanner: resident chunks on a few streams, H2D / kernels / D2H of kept records overlapped)
    # 4 lanes x 128-set chunks at every N (measured at 4 GPUs: 274 M A-scans/s against 243 M with 2 lanes x 256 sets)
    scan_chunk = int(os.environ.get("PAUT_BENCH_CHUNK_ASCANS", "38400"))
    lanes = int(os.environ.get("PAUT_BENCH_LANES", "4"))
    scanner = VolumeScanner(model, chunk_sets=max(1, (scan_chunk // n_per)), lanes=lanes, reuse_output=True)

    def time_e2e(xh, steps):
        for _ in range(3):               # warm-up: lanes, both record arenas, allocator steady state
            scanner.scan(xh, threshold=0.5)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            scanner.scan(xh, threshold=0.5)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e_s = time_e2e(x_host, args.steps)
    h2d_b, d2h_b = scanner.h2d_bytes, scanner.d2h_bytes
    ms, e2e_s = max_over_ranks(ms, e2e_s)
    e2e_extra = {}
    if not args.no_extra:
        # the same volume held as fp32 on the host (what json_dataset.py:112-116 produces): twice the PCIe bytes,
        # the cast to bf16 runs on the device inside the scanner's forward
        if in_dtype == torch.bfloat16:
            x32 = x_host.float().pin_memory()
            steps32 = max(3, args.steps // 4)
            s32 = time_e2e(x32, steps32)
            (s32,) = max_over_ranks(s32)
            e2e_extra["fp32_host"] = {"value": world * n_ascans * steps32 / s32, "unit": "A-scans/s",
                                      "h2d_bytes_per_step": scanner.h2d_bytes, "steps": steps32}
            del x32
        ceil = h2d_ceiling(dev, 640 << 20, 5, barrier)
        t = torch.tensor([ceil], device=dev, dtype=torch.float64)
        tmin = t.clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        e2e_extra["h2d_ceiling_gbs"] = float(t[0])
        # every rank streams the same number of bytes and the step ends with the slowest rank: with equal shards the
        # ceiling of the box is N x the slowest rank's rate (at 8 GPUs this pool gives 4 ranks 23 GB/s and 4 ranks 35 GB/s)
        e2e_extra["h2d_ceiling_equal_shards_gbs"] = world * float(tmin[0])
        e2e_extra["h2d_achieved_gbs"] = world * h2d_b * args.steps / e2e_s / 1e9
        e2e_extra["host_cores_per_rank"] = cores_per_rank

    extra = {}
    if not args.no_extra:
        # ---- strong scaling: the fixed whole-weld volume of configs[4] (8 M A-scans), sharded by scan position
        if kind == "msc":
            lo, hi = shard_range(STRONG_SETS, world, rank)
            reps = -(-(hi - lo) // n_sets)
            xs = x_dev.repeat(reps, 1, 1)[:hi - lo].contiguous() if hi - lo > n_sets else x_dev[:hi - lo]
            steps_s = 5
            ms_s, _, _, _ = time_resident(model, xs, steps_s, 3)
            (ms_s,) = max_over_ranks(ms_s)
            extra["strong_scaling"] = {"workload": f"msc over a fixed {STRONG_SETS * n_per}-A-scan volume "
                                                   f"({STRONG_SETS} sets x {n_per} x {S}) sharded over {world} GPU(s) by scan position "
                                                   "(the 1 M-A-scan synthetic block repeated)",
                                       "value": STRONG_SETS * n_per * steps_s / (ms_s * 1e-3), "unit": "A-scans/s",
                                       "ms_per_step": ms_s / steps_s, "steps": steps_s, "scaling": "strong",
                                       "sets_on_rank0": hi - lo}
            del xs
    if not args.no_extra and world == 1:
        # ---- the other BASELINE configs on one GPU (driver-visible): two_stage + predict, enhanced, ssd
        del x_dev
        torch.cuda.empty_cache()
        models = {}
        ns2, np2 = SETS["two_stage"]
        xh2 = make_volume("two_stage", ns2, np2, seed=43).to(in_dtype)
        xd2 = xh2.to(dev)
        for k2 in ("two_stage", "enhanced", "ssd"):
            try:
                m2 = build(k2)
                st2 = 8 if k2 != "enhanced" else 5
                ms2, l2, prof2, found2 = time_resident(m2, xd2, st2, 3)
                r1, _ = roofline_entries(k2, np2, prof2, ns2 * np2, st2, pk, traffic)
                models[k2] = {"workload": f"{k2} over {ns2 * np2} A-scans ({ns2} sets x {np2} x {S}), forward + predict post-processing",
                              "value": ns2 * np2 * st2 / (ms2 * 1e-3), "unit": "A-scans/s", "ms_per_step": ms2 / st2,
                              "steps": st2, "warmup": 3, "gpu_launches": l2, "detections_last_step": found2,
                              "roofline": r1,
                              "kernels_ms_per_step": {n: round(v[1] / st2, 4) for n, v in sorted(prof2.items(), key=lambda kv: -kv[1][1])[:6]}}
                del m2
            except Exception as ex:
                models[k2] = {"error": f"{type(ex).__name__}: {ex}"}
        extra["models"] = models
        del xd2, xh2
        torch.cuda.empty_cache()
        # ---- the eager comparator on the same GPU
        try:
            xg = x_host[:256].to(dev).float()
            extra["gpu_eager_port"] = {"workload": f"{kind}, oracle port with torch ops on cuda, forward only, 256 sets x {n_per}",
                                       "unit": "A-scans/s", **gpu_eager_port(kind, sd, xg, n_per)}
            del xg
        except Exception as ex:
            extra["gpu_eager_port"] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        value = world * n_ascans * args.steps / (ms * 1e-3)
        e2e = world * n_ascans * args.steps / e2e_s
        roof1, roof2 = roofline_entries(kind, n_per, prof, n_ascans, args.steps, pk, traffic)
        line = {
            "metric": "A-scans/sec", "value": value, "unit": "A-scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": workload, "l2": "input volume (%.0f MB) larger than the 126 MB L2" %
                       (n_ascans * S * (2 if in_dtype == torch.bfloat16 else 4) / 1e6), "sharding": f"dp{world} by scan position",
                       "threshold": 0.5, "detections_last_step": n_found},
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "A-scans/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                    "api": f"VolumeScanner(lanes={lanes}, reuse_output=True).scan(pinned host tensor, "
                           f"{'bf16' if in_dtype == torch.bfloat16 else 'fp32'} as stored on the host)", **e2e_extra},
            "roofline": roof1, "roofline_2nd": roof2,
            "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
        }
        if extra:
            line["extra"] = extra
        if world == 1 and args.cpu_seconds > 0:
            x_cpu = x_host[:64].float()
            rate, done, el = cpu_reference_rate(kind, sd, x_cpu, args.cpu_seconds, n_per)
            ck = cpu_kind(kind, sd)
            line["cpu_baseline"] = {"value": rate, "unit": "A-scans/s", "cores": torch.get_num_threads(), "kind": ck,
                                    "sample": f"{done} A-scans (64 sets cycled) in {el:.1f} s, "
                                              + ("the reference's own class (oracle/_ref) forward + threshold loop"
                                                 if ck == "reference" else "oracle port of the reference forward + predict post-processing")
                                              + ", fp32, all host threads"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
