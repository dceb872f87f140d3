#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 recursive-filter engine.

Workload (BASELINE.json configs[2], the one the metric is quoted on): apps/gaussian --
3rd-order van Vliet-Young-Verbeek Gaussian (sigma 5), causal + anticausal along x and y,
clamped border, 8192 x 8192 float32 images.  One "step" filters a batch of BATCH distinct
synthetic images (so consecutive launches never re-read a cached input; every image,
268 MB, is itself larger than the 126 MB L2).

  N = 1   the whole image on one GPU (rf_plan_execute)
  N > 1   every image is cut into N horizontal strips, one per rank; x scans are strip
          local, the y scans exchange only their order-3 boundary tails (2*3*8192 floats
          per image and rank) with ONE NCCL all-gather per step (rf_plan_stage1 /
          all_gather / rf_plan_stage2).  A step is --batch x N images, so the samples per GPU per
          step are fixed: "scaling": "weak".  --strong keeps --batch images per step; that figure
          and the collective-free batch sharding are also reported (`strip_sharded_strong`,
          `batch_sharded`).

Metric: Gsamples/s = filtered output samples per second over all ranks (device time,
CUDA events, max over ranks).  `roofline` is the dominant kernel (the final tile kernel,
which reads the image and writes the result) against the measured HBM copy bandwidth;
`e2e` is the same metric through the host-buffer C-ABI call (rf_plan_execute_host: H2D +
kernels + D2H inside the timed region); `cpu_baseline` is the oracle's serial recurrence
loops (the reference's CPU path stand-in) on the host cores.

`--impl reference` times that CPU path alone (rank 0 only), same metric and config.
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

import numpy as np  # noqa: E402

W = H = 8192
SIGMA, ORDER = 5.0, 3
BATCH = 8
METRIC = "Gsamples/s and % of HBM roofline for 2-D r=3 Gaussian 8192^2 fp32"
WORKLOAD = "apps/gaussian: 3rd-order VYV Gaussian sigma=5, +x,-x,+y,-y, clamped border, 8192x8192 fp32"


def scans_c3():
    from recfilter_b200 import gaussian_weights
    w3 = gaussian_weights(SIGMA, ORDER)
    return [(0, True, w3), (0, False, w3), (1, True, w3), (1, False, w3)]


def bench_config(N: int, batch: int, strong: bool, exchange: str, graph: bool, overlap: int = 1) -> dict:
    """The workload description of a run -- the same dict for the GPU arm and for `--impl reference` (the CPU arm reports
    the configuration it stands beside; what it actually samples is said in its `cpu_baseline.sample`)."""
    weak = N > 1 and not strong
    B = batch * (N if weak else 1)
    if N == 1:
        sharding = "none"
    else:
        # ShardedFilter's rule: peer-to-peer windows on request; NCCL all-gather, two column-chunked all-to-alls from 4 ranks
        # on when the all-gather would deliver 8 MB or more per rank (2 d scans x order 3 x 8192 columns x B strips x 4 B)
        tail_bytes = 2 * ORDER * W * B * 4
        chunked = exchange == "alltoall" or (exchange == "auto" and N >= 4 and tail_bytes * N >= (8 << 20))
        how = ("peer-to-peer exchange windows over NVLink (rf_xchg_put / rf_xchg_wait)" if exchange == "p2p" else
               "peer-to-peer exchange windows, neighbouring strips only: causal tails to rank + 1, anticausal to rank - 1 "
               "(rf_xchg_put_part / rf_xchg_wait_from; the filter forgets a strip before it leaves it)" if exchange == "neighbor" else
               "two column-chunked NCCL all-to-alls" if chunked else "one NCCL all-gather")
        sharding = (f"every image cut into {N} row strips, one per GPU; the order-3 strip tails travel once per step ({how}); "
                    f"{batch} images' worth of samples per GPU per step")
    groups = overlap if (B > 1 and overlap > 1 and B % overlap == 0) else 1
    return {"workload": WORKLOAD, "images_per_step": B, "sharding": sharding,
            "l2": "every image (268 MB) exceeds L2 and a step sweeps %d distinct images (one stack)" % B,
            "tile": "128x128 register tiles (fused engine)",
            "launch": "one CUDA graph per step (captured once, replayed)" if (graph and B > 1 and exchange not in ("p2p", "neighbor")) else "eager launches",
            "streams": f"{groups} sub-stacks of {B // groups} images on their own CUDA streams" if B > 1 else "1"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons, polled through NVML every few ms while the timed region runs
    (the region lasts tens of ms: `nvidia-smi -lms` would not produce a sample in time)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.thread, self.samples, self.reasons, self.stop_flag = index, None, [], set(), False
        self.max_mhz, self.err = None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    idx = self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception as e:          # noqa: BLE001
            self.err = f"{type(e).__name__}: {e}"

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:       # noqa: BLE001  (older binding name)
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:      # noqa: BLE001
                self.err = f"{type(e).__name__}: {e}"
                return
            time.sleep(0.002)

    def stop(self) -> dict:
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": [],
                    "error": self.err or "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(self.reasons), "source": "NVML polled every 2 ms during the timed region"}


# ------------------------------------------------------------------------------------------------
# CPU path (oracle): the reference-arm and the cpu_baseline leg
# ------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Host cores this process may use.  Passed to the oracle explicitly (an OpenMP num_threads clause): torchrun
    exports OMP_NUM_THREADS=1, which must not shrink the CPU arm."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_rate(rows: int, reps: int, threads: int):
    """Gsamples/s of the serial recurrence loops on a `rows` x 8192 strip of the workload."""
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(2)
    img = rng.random((rows, W), dtype=np.float32)
    sc = scans_c3()
    best = float("inf")
    for _ in range(reps):
        work = img.copy()
        t0 = time.perf_counter()
        oracle.apply_filter(work, sc, "clamp", threads=threads, inplace=True)
        best = min(best, time.perf_counter() - t0)
    return img.size / best / 1e9, best


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path stand-in (oracle port, all host threads)."""
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    threads = host_threads()
    rng = np.random.default_rng(2)
    rows = H
    img = rng.random((rows, W), dtype=np.float32)
    sc = scans_c3()
    for _ in range(max(args.warmup, 1)):
        oracle.apply_filter(img, sc, "clamp", threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.apply_filter(img, sc, "clamp", threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * img.size / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the GPU arm's configuration at this N; the CPU sample of it is one image per step (cpu_baseline.sample)
        "config": bench_config(args.gpus, args.batch, args.strong, args.exchange, bool(args.graph), args.overlap),
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": threads, "kind": "port",
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                         "sample": "one full 8192x8192 image per step, oracle/oracle.c serial recurrence loops, "
                                   f"OpenMP over independent lines with an explicit num_threads({threads}) "
                                   "(Halide x86 JIT of the reference cannot be built here)"},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def ncu_traffic_per_launch():
    """dram read+write bytes per launch of the dominant kernel, from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    try:
        d = json.load(open(path))
        return float(d["dram_bytes_per_launch"]), d.get("source")
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs (C1, C2, C4, C5) beside the headline one: device time, inputs resident
# ------------------------------------------------------------------------------------------------
W2 = [[1.0, 0.5, 0.25], [1.0, 0.5, 0.125], [1.0, 0.5, 0.0625], [1.0, 0.5, 0.125], [1.0, 0.5, 0.25], [1.0, 0.5, 0.0625]]   # tests/test_generic_xyz.cpp:24-30
A8 = [1.0] + [0.01] * 8                                                                                                  # apps/audio/audio_filter_high_order.cpp:41-42


def _time_calls(fn, iters, barrier, dist, world):
    import torch
    for _ in range(3):
        fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def other_configs(N, rank, barrier, peak, exchange):
    """C1 / C2 (one GPU: replicas only), C4 (channels split over the ranks, no collective), C5 (z slabs,
    one tail exchange): Gsamples/s over all ranks, fraction of N x the HBM peak by the 8 B/sample measure."""
    import torch
    import torch.distributed as dist
    from recfilter_b200 import Plan, Scan
    from recfilter_b200.sharded import ShardedFilter
    out = {}

    def entry(name, samples, ms, gpus, how, launches):
        out[name] = {"value": samples / (ms * 1e-3) / 1e9, "unit": "Gsamples/s", "us": ms * 1e3, "n_gpus": gpus,
                     "hbm_frac_of_peak": 8.0 * samples / (ms * 1e-3) / 1e9 / (peak * gpus), "launches_per_call": launches, "how": how}

    sat = [Scan(0, True, [1, 1]), Scan(1, True, [1, 1])]
    if N == 1:
        # the headline filter on ONE image per call (the step of the headline number is a stack of images)
        plan = Plan((W, H), "f32", [Scan(*s) for s in scans_c3()], "clamp")
        src = [torch.rand(W * H, device="cuda") for _ in range(2)]
        dst = torch.empty_like(src[0])
        state = {"i": 0}

        def one():
            i = state["i"] = (state["i"] + 1) % 2
            plan.execute(src[i], dst)
        ms = _time_calls(one, 40, barrier, dist, 1)
        entry("C3 Gaussian 8192^2, one image per call", W * H, ms, 1, plan.describe().splitlines()[1].strip()[:140], plan.num_launches)
        plan.close()
        # apps/usm: unsharp mask (1+w)*image - w*blur, the pointwise stage fused into pass 2's store (rf_options.epilogue)
        plan = Plan((W, H), "f32", [Scan(*s) for s in scans_c3()], "clamp", epilogue=(2.0, -1.0))

        def usm():
            i = state["i"] = (state["i"] + 1) % 2
            plan.execute(src[i], dst)
        ms = _time_calls(usm, 40, barrier, dist, 1)
        entry("apps/usm unsharp mask 8192^2 (Gaussian + fused pointwise epilogue)", W * H, ms, 1,
              plan.describe().splitlines()[1].strip()[-110:], plan.num_launches)
        plan.close()
        del src, dst
        for name, ext, dt, tdt in (("C1 summed-area table 2048^2 u32", (2048, 2048), "u32", torch.int32),
                                   ("C2 box-filter integral image 4096^2 f32", (4096, 4096), "f32", torch.float32)):
            plan = Plan(ext, dt, sat)
            n = ext[0] * ext[1]
            # several distinct inputs in turn, together larger than L2 (C1 / C2 alone would sit in the 126 MB L2)
            reps = max(2, int(300e6 // (4 * n)) + 1)
            src = [(torch.rand(n, device="cuda") if dt == "f32" else torch.randint(0, 256, (n,), device="cuda", dtype=tdt)) for _ in range(reps)]
            dst = [torch.empty_like(x) for x in src]
            state = {"i": 0}

            def call():
                i = state["i"] = (state["i"] + 1) % reps
                plan.execute(src[i], dst[i])
            ms = _time_calls(call, 40, barrier, dist, 1)
            plan.check()
            entry(name, n, ms, 1, plan.describe().splitlines()[1].strip()[:120] + f" ({reps} inputs in turn, > L2)", plan.num_launches)
            plan.close()
            del src, dst
    # C4: 64 channels x 2^24 samples, order 8, channels split over the ranks (no collective)
    ch = 64 // N
    plan = Plan((1 << 24, ch), "f32", [Scan(0, True, A8)])
    src = torch.rand((ch, 1 << 24), device="cuda") * 2 - 1
    dst = torch.empty_like(src)
    ms = _time_calls(lambda: plan.execute(src, dst), 5, barrier, dist, N)
    plan.check()
    entry("C4 audio 64 ch x 2^24 r=8", 64 * (1 << 24), ms, N,
          (f"{ch} channels per GPU, no collective; " if N > 1 else "") + plan.describe().splitlines()[1].strip()[:140], plan.num_launches)
    plan.close()
    del src, dst
    # C5: 512^3, six order-2 scans, z slabs over the ranks
    sc5 = [Scan(0, True, W2[0]), Scan(0, False, W2[1]), Scan(1, True, W2[2]), Scan(1, False, W2[3]), Scan(2, True, W2[4]), Scan(2, False, W2[5])]
    flt = ShardedFilter((512, 512, 512), "f32", sc5, "zero", rank=rank, world=N, shard_dim=2, batch=1, exchange=exchange)
    z = flt.local_extents[2]
    src = torch.rand((z, 512, 512), device="cuda")
    dst = torch.empty_like(src)
    ms = _time_calls(lambda: flt.run([src], [dst]), 10, barrier, dist, N)
    entry("C5 volume 512^3 r=2 xyz", 512 ** 3, ms, N,
          (f"z slabs of {z} planes per GPU, order-2 plane tails exchanged once per call ({'p2p windows' if flt.p2p else 'NCCL'}); " if N > 1 else "") +
          " | ".join(l.strip()[:60] for l in flt.plans[0].describe().splitlines()[1:]), flt.plans[0].num_launches)
    flt.close()
    del src, dst
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C1 / C2 / C4 / C5 figures")
    ap.add_argument("--overlap", type=int, default=1,
                    help="sub-stacks of a step that run on their own CUDA streams (carry stage of one beside the tile "
                         "kernels of another); 1 = one stream")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "neighbor", "allgather", "alltoall"],
                    help="N > 1: how the strip tails travel (auto: NCCL all-gather, or two column-chunked all-to-alls from 4 "
                         "ranks on; p2p: peer-to-peer exchange windows over NVLink, every rank to every rank; neighbor: the "
                         "windows, adjacent strips only -- for filters that forget a strip before they leave it)")
    ap.add_argument("--graph", type=int, default=1,
                    help="1 (default): a step is captured once in a CUDA graph and replayed (a sharded step is a dozen short "
                         "launches: from 4 ranks on the host cannot issue them as fast as the GPUs finish them); 0: eager launches")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: keep --batch images per step (strong scaling) instead of --batch x N (weak scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import recfilter_b200 as rf
    from recfilter_b200 import Scan
    from recfilter_b200.sharded import ShardedFilter

    if not torch.cuda.is_available() or rf.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    rf.lib().rf_set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = world
    assert N == args.gpus or world == 1, "--gpus must match the torchrun world size"

    scans = [Scan(*s) for s in scans_c3()]
    # N > 1: every image is cut into N row strips (one per rank, carries exchanged).  Weak scaling (default): a step
    # is --batch x N images, so every rank holds --batch images' worth of samples whatever N is; --strong keeps
    # --batch images per step.
    weak = N > 1 and not args.strong
    B = args.batch * (N if weak else 1)
    # a step = B distinct 8192x8192 images, held as one dense stack [B][rows][W] and filtered as one filter whose
    # outermost dimension carries no scans (the reference allows that: lib/split.cpp:1888-1898): one launch
    # sequence -- and, sharded, one tail exchange -- per step
    flt = ShardedFilter((W, H), "f32", scans, "clamp", rank=rank, world=N, shard_dim=1, batch=B, stacked=B > 1,
                        overlap=args.overlap, exchange=args.exchange, graph=bool(args.graph))
    rows = flt.local_extents[1]
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    src_stack = torch.rand((B, rows, W), device="cuda", dtype=torch.float32, generator=gen)
    dst_stack = torch.empty_like(src_stack)
    srcs = [src_stack[i] for i in range(B)]
    dsts = [dst_stack[i] for i in range(B)]
    launches_per_image = flt.launches_per_image

    def step():
        if flt.stacked:
            flt.run_stacked(src_stack, dst_stack)
        else:
            flt.run(srcs, dsts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    samples_per_step = B * W * H
    value = args.steps * samples_per_step / (ms * 1e-3) / 1e9

    # ---- per-kernel timing: same steps, same inputs, CUDA events around every launch of the plan's stream ----
    ksteps = max(2, min(args.steps, 5))
    for p in flt.plans:
        p.stage_timing(True)
    use_graph, flt.use_graph = flt.use_graph, False          # per-kernel events need eager launches
    for _ in range(ksteps):
        step()
    torch.cuda.synchronize()
    stage = {}
    for p in flt.plans:
        for k, v in p.stage_times().items():
            s = stage.setdefault(k, {"ms": 0.0, "launches": 0})
            s["ms"] += v["ms"]; s["launches"] += v["launches"]
        p.stage_timing(False)
    flt.use_graph = use_graph
    peak, peak_src = measured_peaks()
    fin = stage["tile_final"]
    k_ms = fin["ms"] / max(fin["launches"], 1)
    imgs_per_launch = flt.sub if flt.stacked else 1
    alg_bytes = 8.0 * W * rows * imgs_per_launch     # 4 B read + 4 B written per sample of this rank's strips, per launch
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    total_stage_ms = sum(v["ms"] for v in stage.values())
    traffic, traffic_src = ncu_traffic_per_launch() if N == 1 else (None, None)
    if traffic is not None:
        traffic *= imgs_per_launch                   # the capture is of a one-image launch; a stack is B of them
    roofline = {"bound": "hbm", "kernel": "fused_tile_kernel<float,3,128,P2> (pass 2: re-scan from carries, store)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_us": k_ms * 1e3, "algorithmic_bytes_per_launch": alg_bytes,
                "share_of_step": fin["ms"] / total_stage_ms if total_stage_ms else None,
                "stage_us_per_image": {k: v["ms"] * 1e3 / (ksteps * B) for k, v in stage.items() if v["launches"]},
                "images_per_launch": imgs_per_launch,
                "kernel_timing": f"CUDA events around every launch on the plan's stream, {ksteps} steps after the timed region",
                # whole filter, all launches: algorithmic bytes of the step / step time, per GPU, over one GPU's peak
                "whole_filter_frac_of_peak": (8.0 * samples_per_step * args.steps / (ms * 1e-3) / 1e9) / (peak * N)}

    # ---- end to end: pinned host buffers, H2D + filter + D2H inside the timed region ----------------------
    e2e_steps = max(2, min(args.steps, 5))
    host_in = [torch.empty((rows, W), dtype=torch.float32).pin_memory() for _ in range(B)]
    host_out = [torch.empty((rows, W), dtype=torch.float32).pin_memory() for _ in range(B)]
    for hi, s in zip(host_in, srcs):
        hi.copy_(s)

    img_plan = None
    if N == 1:
        from recfilter_b200 import Plan as _Plan
        img_plan = _Plan((W, H), "f32", scans, "clamp")          # per-image plan: frames stream through realize()

    # N > 1: the step's strips stream through the GPU in G groups: H2D of group g+1 and D2H of group g-1 run beside the
    # kernels and the tail exchange of group g (three streams, a ring of three device buffers per rank)
    G = 4 if (N > 1 and B % 4 == 0) else 1
    if N > 1:
        gb = B // G
        gflt = ShardedFilter((W, H), "f32", scans, "clamp", rank=rank, world=N, shard_dim=1, batch=gb, stacked=gb > 1,
                             exchange=args.exchange)
        ring = [torch.empty((gb, rows, W), device="cuda", dtype=torch.float32) for _ in range(3)]
        s_up, s_run, s_down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        ev_up = [torch.cuda.Event() for _ in range(3)]
        ev_run = [torch.cuda.Event() for _ in range(3)]
        ev_down = [torch.cuda.Event() for _ in range(3)]

    def e2e_step():
        if N == 1:
            # rf_plan_execute_host_batch: a batch of frames through the host-buffer entry point of the C ABI, copies
            # pipelined with the kernels (three device buffers)
            img_plan.realize_batch_ptr([h.data_ptr() for h in host_in], [h.data_ptr() for h in host_out])
            return
        for g in range(G):
            b = g % 3
            with torch.cuda.stream(s_up):
                if g >= 3:
                    s_up.wait_event(ev_down[b])                          # ring buffer b is free again
                for i in range(gb):
                    ring[b][i].copy_(host_in[g * gb + i], non_blocking=True)
                ev_up[b].record(s_up)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_up[b])
                if gflt.stacked:
                    gflt.run_stacked(ring[b], ring[b])
                else:
                    gflt.run([ring[b][0]], [ring[b][0]])
                ev_run[b].record(s_run)
            with torch.cuda.stream(s_down):
                s_down.wait_event(ev_run[b])
                for i in range(gb):
                    host_out[g * gb + i].copy_(ring[b][i], non_blocking=True)
                ev_down[b].record(s_down)
        s_down.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    e2e_value = e2e_steps * samples_per_step / dt / 1e9
    e2e = {"value": e2e_value, "unit": "Gsamples/s", "h2d_bytes_per_step": 4 * samples_per_step,
           "d2h_bytes_per_step": 4 * samples_per_step, "steps": e2e_steps,
           "api": ("rf_plan_execute_host_batch (host-buffer entry point of the C ABI, one call per step of %d images; H2D / "
                   "kernels / D2H pipelined over 3 device buffers), pinned host buffers" % B) if N == 1 else
                  ("per rank: pinned H2D / rf_plan_stage1 + tail exchange + rf_plan_stage2 / D2H of %d groups of %d strips "
                   "pipelined on three streams over a ring of 3 device buffers" % (G, B // G))}
    if N > 1:
        gflt.close()
        del ring

    # ---- N > 1, weak run: the strong-scaling figure (fixed --batch images per step) beside it -----------------
    strong = None
    if weak:
        sflt = ShardedFilter((W, H), "f32", scans, "clamp", rank=rank, world=N, shard_dim=1, batch=args.batch,
                             stacked=args.batch > 1, exchange=args.exchange, graph=bool(args.graph))
        ssrc, sdst = src_stack[:args.batch].contiguous(), dst_stack[:args.batch].contiguous()

        def sstep():
            if sflt.stacked:
                sflt.run_stacked(ssrc, sdst)
            else:
                sflt.run([ssrc[i] for i in range(args.batch)], [sdst[i] for i in range(args.batch)])
        for _ in range(args.warmup):
            sstep()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            sstep()
        s1.record()
        barrier()
        ts_ = torch.tensor([s0.elapsed_time(s1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        sms = float(ts_.item())
        strong = {"value": args.steps * args.batch * W * H / (sms * 1e-3) / 1e9, "unit": "Gsamples/s", "scaling": "strong",
                  "ms_per_step": sms / args.steps,
                  "config": f"{args.batch} images per step whatever N is, each cut into {N} row strips"}
        sflt.close()
        del ssrc, sdst

    # ---- N > 1: the batched configuration beside the strip-sharded one (north_star: "strip-sharded and
    # batched configs"): every rank filters its own B full images, no collective -> weak scaling ----------
    batch_sharded = None
    if N > 1:
        from recfilter_b200 import Plan
        Bb = args.batch
        bplan = Plan((W, H, Bb), "f32", scans, "clamp")
        bsrc = torch.rand((Bb, H, W), device="cuda", dtype=torch.float32, generator=gen)
        bdst = torch.empty_like(bsrc)
        for _ in range(args.warmup):
            bplan.execute(bsrc, bdst)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            bplan.execute(bsrc, bdst)
        b1.record()
        barrier()
        tb = torch.tensor([b0.elapsed_time(b1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        bms = float(tb.item())
        batch_sharded = {"value": args.steps * Bb * W * H * N / (bms * 1e-3) / 1e9, "unit": "Gsamples/s",
                         "scaling": "weak", "ms_per_step": bms / args.steps,
                         "config": f"every rank filters its own stack of {Bb} full 8192x8192 images per step (no collective)"}
        del bsrc, bdst
        bplan.close()

    # ---- the other BASELINE configs -----------------------------------------------------------------------
    del src_stack, dst_stack, srcs, dsts, host_in, host_out
    torch.cuda.empty_cache()
    others = None
    if not args.no_other_configs:
        others = other_configs(N, rank, barrier, peak, args.exchange)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        th = host_threads()
        v_all, _ = cpu_rate(H, 3, th)
        v_one, _ = cpu_rate(H // 4, 2, 1)
        cpu = {"value": v_all, "unit": "Gsamples/s", "cores": th, "kind": "port",
               "single_thread_value": v_one,
               "sample": f"one full 8192x8192 image (best of 3) with {th} OpenMP threads; single-thread figure on a "
                         "2048x8192 strip; oracle/oracle.c serial recurrence loops (the reference's Halide x86 JIT "
                         "cannot be built here)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (N > 1 and not weak) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(N, args.batch, args.strong, args.exchange, bool(args.graph), args.overlap),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "batch_sharded": batch_sharded,
            "strip_sharded_strong": strong, "other_configs": others,
            "gpu_launches": int(args.steps * B * launches_per_image), "clocks": clocks,
            "hbm_roofline_pct": 100.0 * (8.0 * value) / (peak * N),
            "hbm_roofline_pct_of_8TBs": 100.0 * (8.0 * value) / (8000.0 * N),
        }
        print(json.dumps(line), flush=True)
    flt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
