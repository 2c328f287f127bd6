#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense optical-flow hot path (BASELINE.json metric:
1080p frame-pairs/s).

    python bench.py --gpus N --steps K --warmup W [--workload tvl1|farneback] [--impl reference]

One rank per GPU (torchrun for N > 1).  A *step* is one pass of the hot path over one batch of
synthetic 1080p frame pairs (`--pairs` per GPU, default 32: 133 MB of u8 input per step, larger
than the 126 MB L2).  Frame pairs are independent, so ranks share nothing on the data path
(weak scaling); with N > 1 the step ends with the NCCL gather of every flow field to rank 0.

Printed JSON line (rank 0): the base contract + `roofline` (dominant kernel, CUDA events per launch
on the launching stream, separate profiled step), `cpu_baseline` (oracle timed on the host cores,
bounded sample), `e2e` (same metric through the host-buffer C-ABI call b2f_calc_host, pinned host
memory, H2D + D2H inside the timed region), `gpu_launches`, `clocks`.
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
# keep stdout to the one JSON line (some boxes export NCCL_DEBUG=VERSION, which prints to stdout)
if os.environ.get("B2F_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = os.environ["B2F_NCCL_DEBUG"]
else:  # NCCL prints its version banner to stdout at both VERSION and WARN level
    os.environ.pop("NCCL_DEBUG", None)

H, W = 1080, 1920
WORKLOADS = {
    # BASELINE.json configs[2] / north_star target: TV-L1 1080p, 5 scales / 10 warps / 30 iters, epsilon = 0
    # (fixed work)
    "tvl1": dict(name="cv::cuda::OpticalFlowDual_TVL1 1920x1080 u8, 5 scales/10 warps/30 iters, epsilon=0",
                 dtype="f32"),
    # BASELINE.json configs[1]: Farneback 1080p, 5 pyramid levels (perf_optflow.cpp:242-258)
    "farneback": dict(name="cv::cuda::FarnebackOpticalFlow 1920x1080 u8, numLevels=5 pyrScale=0.5 winSize=13 "
                           "numIters=10 polyN=5 polySigma=1.1", dtype="f32"),
}


BATCH_PARAMS = {
    "tvl1": dict(tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=10, epsilon=0.0, iterations=30, scale_step=0.8,
                 gamma=0.0, use_initial_flow=0),
    "farneback": dict(num_levels=5, pyr_scale=0.5, fast_pyramids=0, win_size=13, num_iters=10, poly_n=5, poly_sigma=1.1,
                      flags=0),
}


def make_alg(workload: str):
    import opencv_contrib_b200 as ocb
    if workload == "tvl1":
        return ocb.OpticalFlowDual_TVL1_create(tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=10, epsilon=0.0,
                                               iterations=30, scaleStep=0.8, gamma=0.0, useInitialFlow=False)
    return ocb.FarnebackOpticalFlow_create(numLevels=5, pyrScale=0.5, fastPyramids=False, winSize=13, numIters=10,
                                           polyN=5, polySigma=1.1, flags=0)


def synth_frames(n_frames: int, seed: int = 0):
    """n_frames distinct 1080p u8 frames; pair i = (frame i, frame i+1) of a drifting texture."""
    import numpy as np
    from oracle import synth
    T = synth.texture(H + 64, W + 64 + 3 * n_frames, seed)
    frames = []
    for i in range(n_frames):
        dy = int(round(8 * np.sin(i * 0.7)))
        frames.append(np.clip(np.rint(T[32 + dy:32 + dy + H, 3 * i:3 * i + W]), 0, 255).astype(np.uint8))
    return frames


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def sample_now(self):
        """One synchronous query (used while work is still queued on the GPU, so short timed regions -- a few
        Farneback steps finish faster than the 200 ms polling period -- still get a sample under load)."""
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            self.lines.extend(l.strip() for l in out.splitlines() if l.strip())
        except Exception:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def peak_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_reference_run(workload: str, steps: int, warmup: int, budget_s: float = 25.0):
    """Times the reference's CPU implementation of the path on the host cores, on a bounded sample
    of the workload.  Farneback: the LIVE reference cv2.calcOpticalFlowFarneback (kind 'reference').
    TV-L1: the oracle port of modules/optflow/src/tvl1flow.cpp (kind 'port'; C/OpenMP build when
    oracle/_build/libtvl1_cpu.so exists, else the numpy restatement)."""
    import numpy as np
    import cv2
    frames = synth_frames(2)
    I0, I1 = frames[0], frames[1]
    cores = os.cpu_count() or 1
    if workload == "farneback":
        cv2.setNumThreads(-1)

        def one():
            cv2.calcOpticalFlowFarneback(I0, I1, None, 0.5, 5, 13, 10, 5, 1.1, 0)

        kind, sample, scale = "reference", "full 1920x1080 pair per step, cv2 %s, %d threads" % (
            cv2.__version__, cv2.getNumThreads()), 1.0
        used = cv2.getNumThreads()
    else:
        from oracle import tvl1_cpu
        native = None
        try:
            from oracle import tvl1_cpu_native
            native = tvl1_cpu_native if tvl1_cpu_native.available() else None
        except Exception:
            native = None
        P = tvl1_cpu.TVL1Params(tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=10, epsilon=0.0,
                                innerIterations=1, outerIterations=30, scaleStep=0.8, gamma=0.0, medianFiltering=1)
        if native is not None:
            used = native.set_threads(native.usable_cpus())

            def one():
                native.calc(I0, I1, P)
            kind, sample, scale = "port", ("full 1920x1080 pair per step, C/OpenMP port of optflow/src/tvl1flow.cpp, "
                                           "%d threads (host reports %d logical CPUs)" % (used, cores)), 1.0
        else:
            # numpy restatement: bounded to a 480x270 crop (1/16 of the pixels), throughput scaled by area
            c0, c1 = I0[:270, :480].copy(), I1[:270, :480].copy()

            def one():
                tvl1_cpu.calc(c0, c1, P)
            kind, scale = "port", (270 * 480) / float(H * W)
            sample = "480x270 crop (1/16 of a 1080p pair) per step, numpy restatement, value scaled by area"
            used = 1
    for _ in range(max(0, min(warmup, 1))):
        one()
    t0 = time.perf_counter()
    n = 0
    while n < steps:
        one()
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    pairs_per_s = n * scale / dt
    return dict(value=pairs_per_s, unit="1080p frame-pairs/s", cores=used, kind=kind, sample=sample,
                steps_run=n, seconds=dt)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    r = cpu_reference_run(args.workload, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "1080p frame-pairs/sec (%s)" % args.workload, "value": r["value"],
        "unit": "1080p frame-pairs/s", "n_gpus": args.gpus, "steps": r["steps_run"], "warmup": min(args.warmup, 1),
        "ms_per_step": 1000.0 * r["seconds"] / max(r["steps_run"], 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": WORKLOADS[args.workload]["dtype"], "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["name"], "pairs_per_step": 1},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "1080p frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def side_measurements(workload: str, pairs, flow_views, dev):
    """Secondary single-stream numbers for the other BASELINE configs (not the headline `value`):
    device-resident, CUDA events, after the main timed region."""
    import torch
    import opencv_contrib_b200 as ocb
    out = {}

    def time_alg(alg, a, b, f, n):
        for _ in range(2):
            alg.calc(a, b, f)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            alg.calc(a, b, f)
        e1.record()
        torch.cuda.synchronize()
        return 1000.0 * n / e0.elapsed_time(e1)

    a, b = pairs[0]
    f = flow_views[0]
    try:
        if workload != "farneback":
            out["farneback_1080p_default_pairs_per_s_1stream"] = time_alg(ocb.FarnebackOpticalFlow_create(), a, b, f, 10)
        if workload != "tvl1":
            out["tvl1_1080p_5x10x30_eps0_pairs_per_s_1stream"] = time_alg(
                ocb.OpticalFlowDual_TVL1_create(nscales=5, warps=10, epsilon=0.0, iterations=30), a, b, f, 5)
        # the reference's create() defaults (5 warps, <= 300 iterations, epsilon 0.01: data-dependent early exit)
        alg = ocb.OpticalFlowDual_TVL1_create()
        out["tvl1_1080p_reference_defaults_eps0.01_pairs_per_s_1stream"] = time_alg(alg, a, b, f, 5)
        out["tvl1_1080p_reference_defaults_iterations_run"] = alg.getStats()["iterations_run"]
        # BASELINE configs[3]: Brox 1280x720, the reference's only parameter set (10, 77, 10)
        bx = (a[:720, :1280].float() / 255.0).contiguous()
        by = (b[:720, :1280].float() / 255.0).contiguous()
        bf = torch.empty((720, 1280, 2), dtype=torch.float32, device=dev)
        out["brox_720p_10_77_10_pairs_per_s_1stream"] = time_alg(
            ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 10, 77, 10), bx, by, bf, 3)
        out["denselk_1080p_default_pairs_per_s_1stream"] = time_alg(ocb.DensePyrLKOpticalFlow_create(), a, b, f, 3)
        # video front end (one upload per frame, 3-stream pipeline): host frames in, host flows out
        import time
        import numpy as np
        hf = [x.cpu().numpy() for (x, _) in pairs[:9]]
        for name, make in (("tvl1_5x10x30_eps0", lambda: ocb.OpticalFlowDual_TVL1_create(nscales=5, warps=10, epsilon=0.0,
                                                                                       iterations=30)),
                           ("farneback_default", lambda: ocb.FarnebackOpticalFlow_create())):
            vf = ocb.VideoFlow(make(), H, W, dtype=np.uint8, depth=3)
            seq = [hf[i % len(hf)] for i in range(25)]
            for _ in vf.run(seq[:4], copy=False):
                pass
            t0 = time.perf_counter()
            n = sum(1 for _ in vf.run(seq, copy=False))
            out["video_%s_1080p_pairs_per_s_host_to_host" % name] = n / (time.perf_counter() - t0)
            vf.close()
        # interpolateFrames, the consumer right after calc (cudalegacy), 1080p
        u, v = (torch.randn((H, W), device=dev) * 3 for _ in range(2))
        f0, f1 = a.float() / 255.0, b.float() / 255.0
        mid, buf = torch.empty_like(f0), torch.empty((6 * H, W), device=dev)
        for _ in range(2):
            ocb.interpolateFrames(f0, f1, u, v, -u, -v, 0.5, mid, buf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ocb.interpolateFrames(f0, f1, u, v, -u, -v, 0.5, mid, buf)
        e1.record()
        torch.cuda.synchronize()
        out["interpolate_frames_1080p_frames_per_s"] = 20000.0 / e0.elapsed_time(e1)
    except Exception as e:  # side numbers must never break the headline line
        out["error"] = repr(e)
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist
    from opencv_contrib_b200.batch import NativeFlowBatch, gather_flows

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.pairs
    frames_h = synth_frames(B + 1, seed=rank)
    frames = [torch.from_numpy(f).to(dev) for f in frames_h]
    pairs = [(frames[i], frames[i + 1]) for i in range(B)]
    flows = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev)
    flow_views = [flows[i] for i in range(B)]
    # native batch front end (csrc/batch.cu): N engine handles on N streams, one C call per batch
    batcher = NativeFlowBatch(args.workload, BATCH_PARAMS[args.workload], n_streams=args.streams)

    def step():
        batcher.run_device(pairs, flow_views)
        if world > 1:
            gather_flows(flows, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    batcher.reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    if rank == 0:
        sampler.sample_now()  # the steps above are asynchronous: the GPU is still working through them
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(ms.item())
    launches = batcher.launches()
    value = world * B * args.steps / (total_ms / 1000.0)

    # ---- roofline of the dominant kernel: one profiled pair (per-launch CUDA events) ----
    roofline = None
    e2e = None
    cpu = None
    extras = None
    if rank == 0:
        alg = make_alg(args.workload)
        alg.calc(*pairs[0], flow_views[0])
        torch.cuda.synchronize()
        alg.setProfiling(True)
        alg.resetStats()
        for i in range(min(3, B)):
            alg.calc(*pairs[i], flow_views[i])
        torch.cuda.synchronize()
        st = alg.getStats()
        dom_name, dom = max(st["classes"].items(), key=lambda kv: kv[1]["ms"])
        peak, peak_src = peak_hbm_gbs()
        ach = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
        share = dom["ms"] / max(sum(c["ms"] for c in st["classes"].values()), 1e-9)
        roofline = {"bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "launches_timed": dom["launches"], "avg_launch_us": 1e3 * dom["ms"] / max(dom["launches"], 1),
                    "share_of_step": share,
                    "algorithmic_bytes_per_launch": dom["bytes"] / max(dom["launches"], 1),
                    "how": "separate profiled calls, CUDA events around every launch on the launching stream",
                    "all_classes_ms": {k: v["ms"] for k, v in st["classes"].items()}}
        alg.setProfiling(False)
        # DRAM traffic of the same kernel class from the committed ncu --set full capture (profiles/)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(dom_name)
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_launch"] = tr["launch"]
                roofline["traffic_algorithmic_bytes"] = tr["algorithmic_bytes_per_launch"]
        except Exception:
            pass
    # ---- e2e: host (pinned) buffers through b2f_calc_host, copies inside the timed region; every rank
    #      runs its own shard at the same time, the slowest rank sets the time ----
    nb = min(B, 16)
    h_in = [torch.from_numpy(f).pin_memory() for f in frames_h[:nb + 1]]
    h_out = torch.empty((nb, H, W, 2), dtype=torch.float32).pin_memory()
    hp = [(h_in[i].numpy(), h_in[i + 1].numpy()) for i in range(nb)]
    ho = [h_out[i].numpy() for i in range(nb)]
    batcher.run_host(hp, ho)
    n_e2e = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        batcher.run_host(hp, ho)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        e2e = {"value": world * nb * n_e2e / float(dt.item()), "unit": "1080p frame-pairs/s",
               "h2d_bytes_per_step": world * nb * 2 * H * W, "d2h_bytes_per_step": world * nb * H * W * 8,
               "pairs_per_step": world * nb, "streams": args.streams,
               "timing": "host wall clock (max over ranks) around b2f_calc_host calls; each call returns after its D2H completed"}
        if not args.no_cpu:
            r = cpu_reference_run(args.workload, steps=3, warmup=1, budget_s=25.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        extras = side_measurements(args.workload, pairs, flow_views, dev) if not args.no_extras else None

    if rank == 0:
        line = {
            "metric": "1080p frame-pairs/sec (%s)" % args.workload, "value": value, "unit": "1080p frame-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": WORKLOADS[args.workload]["dtype"], "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["name"], "pairs_per_step_per_gpu": B,
                       "streams_per_gpu": args.streams, "parallelism": "pairs sharded over %d rank(s)%s" % (
                           world, ", NCCL gather of flows to rank 0 inside the step" if world > 1 else ""),
                       "l2": "inputs per step (%.0f MB u8) exceed the 126 MB L2; engine working set ~0.3 GB/pair" % (
                           B * 2 * H * W / 1e6)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tvl1", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=32, help="frame pairs per step per GPU")
    ap.add_argument("--streams", type=int, default=0,
                    help="engine instances / CUDA streams per GPU (0 = per workload: 4 for tvl1, whose persistent kernels "
                         "fill the GPU on their own; 8 for farneback, whose coarse levels are launch-bound)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary single-stream measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.streams <= 0:
        args.streams = 8 if args.workload == "farneback" else 4
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
