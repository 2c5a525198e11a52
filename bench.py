#!/usr/bin/env python
"""bench.py -- headline measurement of the Lyapunov hot path (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload frame1080|frame4k|bake512] [--mode exact|fast|host]

One "step" = one frame of BASELINE.json's configs[1] (1920x1080, default params/scene:
sequence BCABA, 18 settle + 1008 accumulate iterations per sample) rendered by ONE GPU.
With N GPUs every rank renders one such frame per step (whole frames dealt to ranks, as
the reference's animation driver would; no communication inside the render) and the
frames are gathered to rank 0 over NCCL inside the timed region: weak scaling.

metric  = Giga-iterations/s: logistic-map steps (settle + accumulate) of all exponent
          evaluations the frame needed, counted by the kernel, per second.
value   = kernel path with device-resident buffers (CUDA events, max over ranks).
e2e     = the same frame through the C-ABI host-buffer call lyap_render_host()
          (allocation, H2D of lights, kernel, D2H of the RGBA frame into pinned memory).
roofline= the render kernel against the measured MUFU.LG2 issue peak (exact mode is
          SFU-bound: one lg2.approx per accumulate step) resp. the measured FFMA peak
          (fast mode); peaks are measured live by lyap_probe_peaks().
cpu_baseline = the CPU oracle (C restatement, OpenMP, all host threads) on a bounded
          sample of rows of the same frame.

--impl reference times the reference's own CPU implementation (oracle/_ref/libref_host.so,
the unmodified sources host-compiled; the C restatement if that is absent) the same way.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "frame1080": dict(w=1920, h=1080, seq="BCABA", settle=18, accum=1008, what="1920x1080 frame, default params/scene"),
    "frame4k": dict(w=3840, h=2160, seq="A6B6C6", settle=72, accum=4032, what="3840x2160 frame, A6B6C6, 72+4032 iterations"),
    "bake512": dict(n=512, seq="BCABA", settle=18, accum=1008, what="512^3 voxel bake, default params"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="frame1080", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="exact", choices=["exact", "fast", "host"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=12, help="rows of the frame in the CPU sample")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------ scene setup
def scene_for(wl, lp):
    prm, cam, lights, n_lights, _, _ = lp.params_init()
    prm.settle, prm.accum = wl["settle"], wl["accum"]
    lp.scene_lights_recalculate(lights, n_lights)
    seq = lp.scene_convert_sequence(wl["seq"])
    if "w" in wl:
        lp.scene_cam_recalculate(cam, wl["w"], wl["h"], 1)
    return prm, cam, lights, n_lights, seq


def sample_rows(h, n):
    n = max(1, min(n, h))
    return [int((i + 0.5) * h / n) for i in range(n)]


def cpu_sample(checker, counter, wl, scene, rows, known_calls=None):
    """Render the given rows on the CPU.  Returns (iterations, seconds, calls).  `checker` does
    the timed work; `counter` (the C restatement, bit-identical march) counts the exponent
    evaluations once, outside the timed part, if the checker cannot count them itself."""
    prm, cam, lights, n_lights, seq = scene
    w, h = wl["w"], wl["h"]
    calls = 0
    t0 = time.perf_counter()
    for y in rows:
        _, _, c = checker.render(cam, prm, seq, lights, n_lights, w, h, y0=y, y1=y + 1)
        calls += c or 0
    dt = time.perf_counter() - t0
    if not calls:
        calls = known_calls
    if not calls:
        calls = sum(counter.render(cam, prm, seq, lights, n_lights, w, h, y0=y, y1=y + 1)[2] for y in rows)
    return calls * (prm.settle + prm.accum), dt, calls


# --------------------------------------------------------------------- reference arm
def run_reference(args, wl, rank):
    if rank != 0:
        return
    if "w" not in wl:
        print(json.dumps({"impl": "reference", "unavailable": "reference arm implemented for frame workloads"}))
        return
    import lyapunov3d_b200 as lp
    from oracle import Oracle, RefHost
    port = Oracle()
    if RefHost.available():
        checker, kind = RefHost(), "reference"
    else:
        checker, kind = port, "port"
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    port.set_threads(ncpu)       # torchrun exports OMP_NUM_THREADS=1
    checker.set_threads(ncpu)
    scene = scene_for(wl, lp)
    # bounded sample: ~0.6 s per row on 16 cores; keep the whole K-step run near two minutes
    rows = sample_rows(wl["h"], min(args.cpu_rows, max(2, 160 // max(args.steps, 1))))
    for _ in range(min(args.warmup, 1)):
        cpu_sample(checker, port, wl, scene, rows[:1], known_calls=1)
    iters, secs, calls = 0, 0.0, None
    for _ in range(args.steps):
        i, s, calls = cpu_sample(checker, port, wl, scene, rows, known_calls=calls)
        iters += i
        secs += s
    val = iters / secs / 1e9
    sample = f"{len(rows)} evenly spaced rows of the frame per step ({len(rows) * wl['w']} pixels)"
    print(json.dumps({
        "impl": "reference", "metric": "lyapunov_giga_iters_per_s", "value": val, "unit": "Giter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["what"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Giter/s", "cores": port.threads(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "Giter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_reference_cuda(args, wl, rank):
    """Extra arm (not part of the driver contract): the UNMODIFIED reference kernel.cu, compiled with
    the reference's nvcc flags for sm_100 (oracle/_ref/libref_cuda.so), timed on this GPU with CUDA
    events around its own <<<(W/16,H/16),(16,16)>>> launch -- the kernel this library replaces."""
    if rank != 0:
        return
    import lyapunov3d_b200 as lp
    from oracle import Oracle, RefCuda
    if not RefCuda.available() or "w" not in wl:
        print(json.dumps({"impl": "reference-cuda", "unavailable": "oracle/_ref/libref_cuda.so not built or not a frame workload"}))
        return
    prm, cam, lights, n_lights, seq = scene_for(wl, lp)
    rc = RefCuda()
    _, _, ms = rc.render(cam, prm, seq, lights, n_lights, wl["w"], wl["h"], reps=max(1, min(args.steps, 3)))
    # evaluations: counted by our exact-mode kernel, whose march is bit-identical to the reference kernel's
    evals = int(lp.render(cam, prm, seq, lights, n_lights, wl["w"], wl["h"], mode="exact")[2].item())
    val = evals * (prm.settle + prm.accum) / (ms * 1e-3) / 1e9
    print(json.dumps({"impl": "reference-cuda", "metric": "lyapunov_giga_iters_per_s", "value": val, "unit": "Giter/s", "n_gpus": 1,
                      "ms_per_step": ms, "higher_is_better": True, "dtype": "f32/f64 mixed (reference arithmetic)", "data": "synthetic",
                      "config": {"workload": wl["what"], "kernel": "kernel_calc_render<<<(W/16,H/16),(16,16)>>>, best of %d" % max(1, min(args.steps, 3))},
                      "frames_per_s": 1e3 / ms}))


# --------------------------------------------------------------------------- our arm
def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if args.impl == "reference-cuda":
        run_reference_cuda(args, wl, rank)
        return

    import torch
    import torch.distributed as dist
    import lyapunov3d_b200 as lp
    from lyapunov3d_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    scene = scene_for(wl, lp)
    prm, cam, lights, n_lights, seq = scene
    iters_per_eval = prm.settle + prm.accum
    d_lights = api.upload_lights(lights, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    from lyapunov3d_b200 import dist as ld
    peer = None
    if "w" in wl:
        w, h = wl["w"], wl["h"]
        rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
        pts = torch.zeros((h, w, 36), dtype=torch.uint8, device=dev)
        frame_bytes = w * h * 4
        if world > 1:
            # rank 0 owns one RGBA slot per rank; every rank's kernel stores its frame straight
            # into its slot over NVLink (CUDA IPC peer mapping): there is no gather step
            peer = ld.PeerBuffer(frame_bytes * world)

        def step():
            pts.zero_()     # the frame contract: miss pixels shade a zeroed LyapPoint
            if world == 1:
                return lp.render(cam, prm, seq, d_lights, n_lights, w, h, mode=args.mode, rgba=rgba, points=pts)[2]
            ev = torch.zeros(1, dtype=torch.int64, device=dev)
            api.render_into(peer.ptr + rank * frame_bytes, pts.data_ptr(), cam, prm, seq, d_lights, n_lights, w, h,
                            mode=args.mode, evals=ev)
            torch.cuda.current_stream().synchronize()
            dist.barrier()          # rank 0 may consume all frames of the step from here on
            return ev
        launches_per_step = 1
        units = w * h
    else:
        n = wl["n"]
        z0, z1 = ld.slab_range(n, rank, world)     # z-slab sharding: strong scaling
        if world > 1:
            peer = ld.PeerBuffer(n ** 3 * 4)       # the full volume lives on rank 0 only
        else:
            vol = torch.zeros((n, n, n), dtype=torch.float32, device=dev)

        def step():
            if world == 1:
                lp.bake(prm, seq, n, mode=args.mode, out=vol)
            else:
                ld.bake_sharded_peer(peer, prm, seq, n, n, n, mode=args.mode)
            return torch.tensor([(z1 - z0) * n * n], device=dev)
        launches_per_step = 1
        units = n ** 3

    for _ in range(max(args.warmup, 0)):
        ev = step()
    torch.cuda.synchronize()
    peaks = api.probe_peaks() if rank == 0 else None

    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    evals = 0
    t_wall = time.perf_counter()
    e0.record()
    ev_list = []
    for i in range(args.steps):
        flush.zero_()                 # L2 flush between timed iterations (inside the timed region)
        k0[i].record()
        ev_list.append(step())
        k1[i].record()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in zip(k0, k1)) / args.steps
    evals = sum(int(e.item()) for e in ev_list)
    stats = torch.tensor([ms, float(evals), kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, kernel_ms = float(mx[0]), float(mx[2])
        total_evals = float(sm[1]) if "w" in wl else float(units) * args.steps
    else:
        total_evals = float(evals)
    total_iters = total_evals * iters_per_eval
    value = total_iters / (ms * 1e-3) / 1e9

    # ---- end-to-end through the C ABI with host buffers (every rank, its own frame)
    e2e = None
    if "w" in wl:
        host_rgba = torch.zeros((wl["h"], wl["w"], 4), dtype=torch.uint8).pin_memory().numpy()
        lp.render_host(cam, prm, seq, lights, n_lights, wl["w"], wl["h"], mode=args.mode, device=local, want_points=False, rgba=host_rgba)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e_evals = 0
        for _ in range(args.steps):
            _, _, e = lp.render_host(cam, prm, seq, lights, n_lights, wl["w"], wl["h"], mode=args.mode, device=local,
                                     want_points=False, rgba=host_rgba)
            e_evals += e
        dt = time.perf_counter() - t0
        t = torch.tensor([dt, float(e_evals)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            dt, e_evals = float(tm[0]), float(ts[1])
        e2e = {"value": e_evals * iters_per_eval / dt / 1e9, "unit": "Giter/s",
               "h2d_bytes_per_step": int(224 * n_lights + 224 + 56 + seq.nbytes),
               "d2h_bytes_per_step": int(wl["w"] * wl["h"] * 4 + 8),
               "ms_per_step": dt / args.steps * 1e3, "frames_per_s": args.steps * world / dt,
               "api": "lyap_render_host (C ABI, host buffers: H2D of lights + kernel + D2H of the frame + sync per call)"}

    if peer is not None:
        peer.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from this run's own events and live peaks
    sms = peaks["sm_count"]
    if args.mode == "fast":
        fp32_per_iter = (2.0 * prm.settle + 4.0 * prm.accum) / iters_per_eval
        peak = peaks["ffma_lane_ops_per_s"] / fp32_per_iter / 1e9
        bound, note = "fp32", "FFMA-issue bound: %.3f FP32 lane-ops per iteration" % fp32_per_iter
    else:
        mufu_per_iter = prm.accum / iters_per_eval
        peak = peaks["mufu_lane_ops_per_s"] / mufu_per_iter / 1e9
        bound, note = "sfu", "MUFU.LG2-issue bound: %.4f lg2 per iteration" % mufu_per_iter
    per_gpu_iters_per_step = total_iters / args.steps / world
    achieved = per_gpu_iters_per_step / (kernel_ms * 1e-3) / 1e9
    traffic_path = os.path.join(ROOT, "profiles", "traffic_%s_%s.json" % (args.workload, args.mode))
    traffic = json.load(open(traffic_path)).get("dram_bytes_per_launch") if os.path.exists(traffic_path) else None
    out_bytes = units * (40 if "w" in wl else 4) / world
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "Giter/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "render_kernel<%s>" % args.mode if "w" in wl else "bake_kernel<%s>" % args.mode,
                "peak_source": "measured live by lyap_probe_peaks (register-only MUFU.LG2/FFMA loops on this GPU): "
                               "%.2f T MUFU/s, %.2f T FFMA/s" % (peaks["mufu_lane_ops_per_s"] / 1e12, peaks["ffma_lane_ops_per_s"] / 1e12),
                "note": note, "algorithmic_bytes_per_launch": out_bytes,
                "hbm_gbs_sanity": out_bytes / (kernel_ms * 1e-3) / 1e9}

    cpu = None
    if not args.no_cpu_baseline and "w" in wl and world == 1:
        from oracle import Oracle   # the checker, timed as the reported CPU baseline only
        port = Oracle()
        port.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
        rows = sample_rows(wl["h"], args.cpu_rows)
        it, secs, _ = cpu_sample(port, port, wl, scene, rows)
        cpu = {"value": it / secs / 1e9, "unit": "Giter/s", "cores": port.threads(), "kind": "port",
               "sample": f"{len(rows)} evenly spaced rows of the frame ({len(rows) * wl['w']} pixels), {secs:.1f} s"}

    line = {
        "metric": "lyapunov_giga_iters_per_s", "value": value, "unit": "Giter/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if "w" in wl else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["what"] + (", one frame per GPU per step" if "w" in wl else ", z-slabs across GPUs"),
                   "mode": args.mode, "sequence": wl["seq"], "settle": prm.settle, "accum": prm.accum,
                   "l2": "flushed between timed steps (256 MiB device write inside the timed region)",
                   "gather": ("none needed: every rank's kernel stores its shard directly into rank 0's buffer over NVLink "
                              "(CUDA IPC peer mapping); stream sync + barrier per step inside the timed region") if world > 1 else "none"},
        "frames_per_s": (args.steps * world / (ms * 1e-3)) if "w" in wl else None,
        "evaluations_per_step": total_evals / args.steps, "wall_ms": wall_ms,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps * world,
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
