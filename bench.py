#!/usr/bin/env python
"""bench.py -- headline measurement of the Lyapunov hot path (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cuda]
                    [--workload frame1080|frame4k|bake512] [--mode exact|fast|host|hybrid|hybrid_host]
                    [--jitter J] [--no-subrecords] [--no-cpu-baseline] [--no-reference-cuda]

One "step" = ONE frame of BASELINE.json's configs[1] (1920x1080, default params/scene: sequence
BCABA, 18 settle + 1008 accumulate iterations per sample).  With N GPUs the SAME frame is cut into
8x8-pixel tiles dealt round-robin to the ranks (north_star: "load-balanced interleaved image tiles");
every rank's kernel stores its tiles straight into rank 0's frame over NVLink (CUDA IPC peer mapping),
so there is no gather step: STRONG scaling, value(N) / value(1) is the speed-up of one frame.  Rank 0
checks, outside the timed region, that the assembled frame equals its own single-GPU render byte for byte.

metric  = Giga-iterations/s: logistic-map steps (settle + accumulate) of all exponent evaluations the
          frame needed, counted by the kernels, per second.
value   = kernel path with device-resident buffers (CUDA events on the launching stream, max over ranks).
e2e     = the same frame through the public host-buffer API: N=1 the C-ABI call lyap_render_host()
          (H2D of the lights, kernel, D2H of the frame into pinned memory, sync); N>1 the peer-sharded
          render plus rank 0's D2H of the assembled frame.
roofline= the dominant kernel against the measured MUFU.LG2 issue peak (exact: one lg2.approx per
          accumulate step, SFU-bound) resp. the measured FFMA peak (fast / hybrid); peaks measured live
          by lyap_probe_peaks().
cpu_baseline = the CPU oracle (C restatement, OpenMP, all host threads) on a bounded sample of rows.
reference_cuda = the UNMODIFIED reference kernel.cu (nvcc --use_fast_math, sm_100; oracle/_ref/libref_cuda.so)
          timed on this GPU with CUDA events around its own <<<(W/16,H/16),(16,16)>>> launch (N=1 only).
sub     = the north-star's other multi-GPU configs, measured the same way in the same run:
          frame4k (3840x2160, A6B6C6, 72+4032, interleaved tiles) and bake512 (512^3, z-slabs written
          straight into rank 0's volume), each with ms, Giter/s, speed-up base and the byte-equality check.

--impl reference times the reference's own CPU implementation (oracle/_ref/libref_host.so, the unmodified
sources host-compiled; the C restatement if that is absent) on sampled rows of the same frame.
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
TILE = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="frame1080", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="exact", choices=["exact", "fast", "host", "hybrid", "hybrid_host"])
    ap.add_argument("--jitter", type=float, default=None, help="override prm.jitter (hybrid modes need 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true")
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
def scene_for(wl, host, jitter=None):
    """The workload's scene from a host layer with the reference's interface: `host` is the product
    package (our arm) or the host-compiled reference itself (reference arm)."""
    prm, cam, lights, n_lights, _, _ = host.params_init()
    prm.settle, prm.accum = wl["settle"], wl["accum"]
    if jitter is not None:
        prm.jitter = jitter
    if hasattr(host, "scene_lights_recalculate"):
        host.scene_lights_recalculate(lights, n_lights)
        seq = host.scene_convert_sequence(wl["seq"])
        if "w" in wl:
            host.scene_cam_recalculate(cam, wl["w"], wl["h"], 1)
    else:
        host.lights_recalculate(lights, n_lights)
        seq = host.convert_sequence(wl["seq"])
        if "w" in wl:
            host.cam_recalculate(cam, wl["w"], wl["h"], 1)
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


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# --------------------------------------------------------------------- reference arm
def run_reference(args, wl, rank):
    """The reference's own CPU implementation of the path, nothing of this repo's product on it: the
    scene comes from the reference's params_init / scene_* (host-compiled), the rows from its raymarch /
    shade; the C restatement only counts evaluations, outside the timed part."""
    if rank != 0:
        return
    if "w" not in wl:
        print(json.dumps({"impl": "reference", "unavailable": "reference arm implemented for frame workloads"}))
        return
    from oracle import Oracle, RefHost
    port = Oracle()
    if RefHost.available():
        checker, kind = RefHost(), "reference"
    else:
        checker, kind = port, "port"
    ncpu = host_threads()
    port.set_threads(ncpu)       # torchrun exports OMP_NUM_THREADS=1
    checker.set_threads(ncpu)
    scene = scene_for(wl, checker, args.jitter)
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
    sample = (f"{len(rows)} evenly spaced rows of the frame per step ({len(rows) * wl['w']} pixels); the metric is per "
              "iteration, so a row sample measures the same Giter/s as the whole frame would")
    print(json.dumps({
        "impl": "reference", "metric": "lyapunov_giga_iters_per_s", "value": val, "unit": "Giter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["what"], "sample": sample, "scene_from": "the reference's own params_init/scene_* (host-compiled)"
                   if kind == "reference" else "the C restatement's params_init/scene_*"},
        "cpu_baseline": {"value": val, "unit": "Giter/s", "cores": port.threads(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "Giter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def time_reference_cuda(wl, scene, evals, reps):
    """The unmodified reference kernel on this GPU (checker library, timed as a reported baseline)."""
    from oracle import RefCuda
    if not RefCuda.available() or "w" not in wl:
        return None
    prm, cam, lights, n_lights, seq = scene
    _, _, ms = RefCuda().render(cam, prm, seq, lights, n_lights, wl["w"], wl["h"], reps=reps)
    return {"ms": ms, "giter_s": evals * (prm.settle + prm.accum) / (ms * 1e-3) / 1e9, "frames_per_s": 1e3 / ms,
            "kernel": "kernel_calc_render<<<(W/16,H/16),(16,16)>>> of the unmodified kernel.cu, nvcc --use_fast_math -arch=sm_100, "
                      "best of %d launch(es), CUDA events" % reps}


def run_reference_cuda(args, wl, rank):
    """Extra arm (not part of the driver contract): the reference CUDA kernel alone."""
    if rank != 0:
        return
    import lyapunov3d_b200 as lp
    scene = scene_for(wl, lp, args.jitter)
    prm, cam, lights, n_lights, seq = scene
    if "w" not in wl:
        print(json.dumps({"impl": "reference-cuda", "unavailable": "not a frame workload"}))
        return
    # evaluations: counted by our exact-mode kernel, whose march is bit-identical to the reference kernel's
    evals = int(lp.render(cam, prm, seq, lights, n_lights, wl["w"], wl["h"], mode="exact")[2].item())
    rc = time_reference_cuda(wl, scene, evals, max(1, min(args.steps, 3)))
    if rc is None:
        print(json.dumps({"impl": "reference-cuda", "unavailable": "oracle/_ref/libref_cuda.so not built"}))
        return
    print(json.dumps({"impl": "reference-cuda", "metric": "lyapunov_giga_iters_per_s", "value": rc["giter_s"], "unit": "Giter/s", "n_gpus": 1,
                      "ms_per_step": rc["ms"], "higher_is_better": True, "dtype": "f32/f64 mixed (reference arithmetic)", "data": "synthetic",
                      "config": {"workload": wl["what"], "kernel": rc["kernel"]}, "frames_per_s": rc["frames_per_s"]}))


# --------------------------------------------------------------------------- our arm
class Runner:
    """One workload on this rank's GPU: step() queues one step on the current stream and returns the
    device counter of exponent evaluations this rank performed; check() compares the sharded result
    on rank 0 with a single-GPU run."""

    def __init__(self, wl, mode, jitter, rank, world, dev):
        import torch
        import lyapunov3d_b200 as lp
        from lyapunov3d_b200 import api
        from lyapunov3d_b200 import dist as ld
        self.torch, self.lp, self.api, self.ld = torch, lp, api, ld
        self.wl, self.mode, self.rank, self.world, self.dev = wl, mode, rank, world, dev
        self.scene = scene_for(wl, lp, jitter)
        self.prm, self.cam, self.lights, self.n_lights, self.seq = self.scene
        self.iters_per_eval = self.prm.settle + self.prm.accum
        self.d_lights = api.upload_lights(self.lights, dev)
        self.frame = "w" in wl
        self.peers = []
        if self.frame:
            w, h = wl["w"], wl["h"]
            self.units = w * h
            self.out_bytes = w * h * 40                     # 4 B RGBA + 36 B LyapPoint per pixel
            if world == 1:
                self.rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
                self.pts = torch.zeros((h, w, 36), dtype=torch.uint8, device=dev)
            else:
                # rank 0 owns the frame; every rank's kernel stores its tiles into it over NVLink
                self.p_rgba = ld.PeerBuffer(w * h * 4)
                self.p_pts = ld.PeerBuffer(w * h * 36)
                self.peers = [self.p_rgba, self.p_pts]
        else:
            n = wl["n"]
            self.units = n ** 3
            self.out_bytes = n ** 3 * 4
            self.z0, self.z1 = ld.slab_range(n, rank, world)
            if world == 1:
                self.vol = torch.zeros((n, n, n), dtype=torch.float32, device=dev)
            else:
                self.p_vol = ld.PeerBuffer(n ** 3 * 4)     # the full volume lives on rank 0 only
                self.peers = [self.p_vol]

    def step(self):
        torch, lp, api, ld, wl = self.torch, self.lp, self.api, self.ld, self.wl
        import torch.distributed as dist
        if self.frame:
            w, h = wl["w"], wl["h"]
            if self.world == 1:
                self.pts.zero_()     # the frame contract: miss pixels shade a zeroed LyapPoint
                return lp.render(self.cam, self.prm, self.seq, self.d_lights, self.n_lights, w, h, mode=self.mode,
                                 rgba=self.rgba, points=self.pts)[2]
            self.p_pts.zero_()       # rank 0 clears the shared point buffer ...
            torch.cuda.current_stream().synchronize()
            dist.barrier()           # ... before any rank's kernel may store into it
            ev = torch.zeros(1, dtype=torch.int64, device=self.dev)
            ld.render_frame_sharded_peer(self.p_rgba, self.p_pts, self.cam, self.prm, self.seq, self.d_lights, self.n_lights,
                                         w, h, mode=self.mode, tile=TILE, evals=ev)   # kernel + stream sync + barrier
            return ev
        n = wl["n"]
        if self.world == 1:
            lp.bake(self.prm, self.seq, n, mode=self.mode, out=self.vol)
        else:
            ld.bake_sharded_peer(self.p_vol, self.prm, self.seq, n, n, n, mode=self.mode)
        return torch.tensor([(self.z1 - self.z0) * n * n], device=self.dev)

    def check(self):
        """Rank 0: the result assembled from all ranks' shards == this GPU's own single-launch result."""
        torch, lp, wl = self.torch, self.lp, self.wl
        import torch.distributed as dist
        if self.world == 1:
            return {"checked": False, "note": "single GPU: nothing to assemble"}
        res = None
        if self.rank == 0:
            if self.frame:
                w, h = wl["w"], wl["h"]
                one_rgba, one_pts, _ = lp.render(self.cam, self.prm, self.seq, self.d_lights, self.n_lights, w, h, mode=self.mode)
                got_rgba = self.p_rgba.view((h, w, 4), "|u1").tensor()
                got_pts = self.p_pts.view((h, w, 36), "|u1").tensor()
                res = {"checked": True, "rgba_bytes_equal": bool(torch.equal(got_rgba, one_rgba)),
                       "points_bytes_equal": bool(torch.equal(got_pts, one_pts)),
                       "checksum_rgba": int(got_rgba.to(torch.int64).sum().item())}
                res["equal_to_single_gpu"] = res["rgba_bytes_equal"] and res["points_bytes_equal"]
            else:
                n = wl["n"]
                one = lp.bake(self.prm, self.seq, n, mode=self.mode)
                got = self.p_vol.view((n, n, n), "<f4").tensor()
                eq = bool(torch.equal(got.view(torch.int32), one.view(torch.int32)))
                res = {"checked": True, "volume_bytes_equal": eq, "equal_to_single_gpu": eq}
            torch.cuda.synchronize()
        dist.barrier()
        return res

    def close(self):
        for p in self.peers:
            p.close()
        self.peers = []


def timed_run(run, steps, warmup, flush, sampler=None):
    """W warm-up steps, then K timed steps bracketed by barrier + synchronize; CUDA events around the
    whole region and around each step; max over ranks.  Returns a dict of totals."""
    import torch
    import torch.distributed as dist
    world = run.world
    for _ in range(max(warmup, 0)):
        run.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t_wall = time.perf_counter()
    e0.record()
    ev_list = []
    for i in range(steps):
        flush.zero_()                 # L2 flush between timed iterations (inside the timed region)
        k0[i].record()
        ev_list.append(run.step())
        k1[i].record()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    step_ms = sum(a.elapsed_time(b) for a, b in zip(k0, k1)) / steps
    evals = sum(int(e.item()) for e in ev_list)
    stats = torch.tensor([ms, float(evals), step_ms], dtype=torch.float64, device=run.dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, step_ms, total_evals = float(mx[0]), float(mx[2]), float(sm[1])
    else:
        total_evals = float(evals)
    total_iters = total_evals * run.iters_per_eval
    return {"ms": ms, "ms_per_step": ms / steps, "step_ms": step_ms, "total_evals": total_evals, "total_iters": total_iters,
            "giter_s": total_iters / (ms * 1e-3) / 1e9, "wall_ms": wall_ms, "clocks": clocks}


def sub_record(name, mode, steps, warmup, rank, world, dev, flush):
    """One of the north-star's other configs, measured like the headline (strong scaling)."""
    run = Runner(WORKLOADS[name], mode, None, rank, world, dev)
    t = timed_run(run, steps, warmup, flush)
    chk = run.check()
    run.close()
    rec = {"workload": WORKLOADS[name]["what"], "mode": mode, "steps": steps, "warmup": warmup, "ms": t["ms_per_step"],
           "giter_s": t["giter_s"], "scaling": "strong", "shard_check": chk,
           "partition": ("%dx%d tiles round-robin over ranks" % (TILE, TILE)) if run.frame else "contiguous z-slabs"}
    rec["frames_per_s" if run.frame else "volumes_per_s"] = 1e3 / t["ms_per_step"]
    return rec


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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    run = Runner(wl, args.mode, args.jitter, rank, world, dev)
    prm, cam, lights, n_lights, seq = run.scene
    iters_per_eval = run.iters_per_eval
    # peaks first (also warms the clocks), then the timed region
    for _ in range(min(args.warmup, 1)):
        run.step()
    torch.cuda.synchronize()
    peaks = api.probe_peaks() if rank == 0 else None
    sampler = ClockSampler(local) if rank == 0 else None
    t = timed_run(run, args.steps, args.warmup, flush, sampler)
    shard_check = run.check()
    ms, total_evals, total_iters = t["ms"], t["total_evals"], t["total_iters"]
    if not run.frame:
        total_iters = float(run.units) * args.steps * iters_per_eval
        total_evals = float(run.units) * args.steps
    value = total_iters / (ms * 1e-3) / 1e9

    # ---- end to end with HOST buffers
    e2e = None
    if run.frame:
        w, h = wl["w"], wl["h"]
        host_rgba = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()
        h2d = int(224 * n_lights + 224 + 56 + seq.nbytes)
        if world == 1:
            hr = host_rgba.numpy()
            lp.render_host(cam, prm, seq, lights, n_lights, w, h, mode=args.mode, device=local, want_points=False, rgba=hr)
            t0 = time.perf_counter()
            e_evals = 0
            for _ in range(args.steps):
                e_evals += lp.render_host(cam, prm, seq, lights, n_lights, w, h, mode=args.mode, device=local, want_points=False, rgba=hr)[2]
            dt = time.perf_counter() - t0
            how = "lyap_render_host (C ABI, host buffers: H2D of lights + kernel + D2H of the frame + sync per call)"
        else:
            host_ev = torch.zeros(1, dtype=torch.int64).pin_memory()

            def e2e_step():
                d_l = api.upload_lights(lights, dev)                      # H2D of this step's inputs
                run.d_lights = d_l
                ev = run.step()                                           # zero + barrier + sharded render + barrier
                host_ev.copy_(ev, non_blocking=True)
                if rank == 0:
                    host_rgba.copy_(run.p_rgba.view((h, w, 4), "|u1").tensor(), non_blocking=True)   # D2H of the assembled frame
                torch.cuda.current_stream().synchronize()
                return int(host_ev.item())
            e2e_step()
            dist.barrier()
            t0 = time.perf_counter()
            e_evals = 0
            for _ in range(args.steps):
                e_evals += e2e_step()
            dist.barrier()
            dt = time.perf_counter() - t0
            how = ("peer-sharded render (dist.render_frame_sharded_peer: per step H2D of lights on every rank, zero + barrier, tile kernels "
                   "storing into rank 0's frame over NVLink, barrier) + rank 0's D2H of the assembled frame into pinned memory")
        tt = torch.tensor([dt, float(e_evals)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = tt.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = tt.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            dt, e_evals = float(tm[0]), float(ts[1])
        e2e = {"value": e_evals * iters_per_eval / dt / 1e9, "unit": "Giter/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": int(w * h * 4 + 8 * world), "ms_per_step": dt / args.steps * 1e3,
               "frames_per_s": args.steps / dt, "api": how}
    run.close()

    # ---- the north-star's other multi-GPU configs in the same run (strong scaling, byte-checked)
    sub = None
    if not args.no_subrecords and args.workload == "frame1080":
        sub = {}
        sub["frame4k"] = sub_record("frame4k", args.mode if args.mode in ("exact", "fast") else "exact", 2, 1, rank, world, dev, flush)
        sub["bake512_fast"] = sub_record("bake512", "fast", 5, 2, rank, world, dev, flush)
        sub["bake512_exact"] = sub_record("bake512", "exact", 5, 2, rank, world, dev, flush)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from this run's own events and live peaks
    if args.mode in ("fast", "hybrid", "hybrid_host") and (args.mode == "fast" or prm.jitter == 0.0):
        fp32_per_iter = (2.0 * prm.settle + 4.0 * prm.accum) / iters_per_eval
        peak = peaks["ffma_lane_ops_per_s"] / fp32_per_iter / 1e9
        bound, note = "fp32", "FFMA-issue bound: %.3f FP32 lane-ops per iteration" % fp32_per_iter
        if args.mode != "fast":
            note += " (march samples, ~97.7 % of the evaluations; refinement and normals run on the parity evaluator)"
    elif args.mode in ("host", "hybrid_host"):
        # bit-exact glibc logf per step: 2 conversions (F2F f32<->f64) on the XU pipe, the pipe MUFU shares
        # (16 lanes/clk/SM, measured by the MUFU.LG2 probe), 6 FP64 ops, 1 LDS.128; the tightest pipe is XU
        xu_per_iter = 2.0 * prm.accum / iters_per_eval
        peak = peaks["mufu_lane_ops_per_s"] / xu_per_iter / 1e9
        bound, note = "xu", ("XU-pipe bound: %.3f F2F conversions per iteration (glibc-exact logf in FP64); the issue model "
                             "(17.2 instructions + 6 double-dispatch FP64 per step) caps it at ~0.69 of this" % xu_per_iter)
    else:
        mufu_per_iter = prm.accum / iters_per_eval
        peak = peaks["mufu_lane_ops_per_s"] / mufu_per_iter / 1e9
        bound, note = "sfu", "MUFU.LG2-issue bound: %.4f lg2 per iteration" % mufu_per_iter
    per_gpu_iters_per_step = total_iters / args.steps / world
    achieved = per_gpu_iters_per_step / (t["step_ms"] * 1e-3) / 1e9
    traffic_path = os.path.join(ROOT, "profiles", "traffic_%s_%s.json" % (args.workload, args.mode))
    traffic = json.load(open(traffic_path)).get("dram_bytes_per_launch") if (os.path.exists(traffic_path) and world == 1) else None
    out_bytes = run.out_bytes / world          # every rank's launch writes its 1/N of the frame / volume
    kernel = ("march_fast2_kernel + render_kernel" if args.mode.startswith("hybrid") and prm.jitter == 0.0 else
              "render_fast2_kernel" if args.mode == "fast" else "render_kernel<%s>" % args.mode.replace("hybrid_host", "host").replace("hybrid", "exact"))
    if not run.frame:
        kernel = "bake_kernel<%s>" % args.mode
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "Giter/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel,
                "peak_source": "measured live by lyap_probe_peaks (register-only MUFU.LG2/FFMA loops on this GPU): "
                               "%.2f T MUFU/s, %.2f T FFMA/s; theoretical %d SMs x 16 resp. 128 lanes x clock"
                               % (peaks["mufu_lane_ops_per_s"] / 1e12, peaks["ffma_lane_ops_per_s"] / 1e12, peaks["sm_count"]),
                "frac_of_theoretical": None, "note": note, "algorithmic_bytes_per_launch": out_bytes,
                "hbm_gbs_sanity": out_bytes / (t["step_ms"] * 1e-3) / 1e9}
    clk = (t["clocks"] or {}).get("sm_mhz") or (t["clocks"] or {}).get("sm_max_mhz")
    if clk:
        lanes = (128.0 / ((2.0 * prm.settle + 4.0 * prm.accum) / iters_per_eval) if bound == "fp32" else
                 16.0 / ((2.0 if bound == "xu" else 1.0) * prm.accum / iters_per_eval))
        roofline["frac_of_theoretical"] = achieved / (peaks["sm_count"] * lanes * clk * 1e6 / 1e9)

    cpu = None
    if not args.no_cpu_baseline and run.frame and world == 1:
        from oracle import Oracle   # the checker, timed as the reported CPU baseline only
        port = Oracle()
        port.set_threads(host_threads())
        rows = sample_rows(wl["h"], args.cpu_rows)
        it, secs, _ = cpu_sample(port, port, wl, run.scene, rows)
        cpu = {"value": it / secs / 1e9, "unit": "Giter/s", "cores": port.threads(), "kind": "port",
               "sample": f"{len(rows)} evenly spaced rows of the frame ({len(rows) * wl['w']} pixels), {secs:.1f} s"}

    ref_cuda = None
    if not args.no_reference_cuda and run.frame and world == 1:
        try:
            # its evaluation count is the exact-mode count (bit-identical march, tests/test_gpu_parity.py)
            ev_exact = total_evals / args.steps if args.mode == "exact" else float(
                lp.render(cam, prm, seq, lights, n_lights, wl["w"], wl["h"], mode="exact")[2].item())
            ref_cuda = time_reference_cuda(wl, run.scene, ev_exact, 1)
            if ref_cuda:
                ref_cuda["speedup"] = ref_cuda["ms"] / t["ms_per_step"]
        except Exception as exc:          # a missing checker must not take the headline down
            ref_cuda = {"unavailable": repr(exc)}

    # kernels of ours per step: the render kernel (march + refine in hybrid mode) and, for a frame, the seven
    # small launches that order its tile queue (tile_key_kernel + CUB radix sort: histogram, scan, 4 passes;
    # ~60 us together -- profiles/r02_launches_bench_exact.csv)
    n_kernels = (2 if (args.mode.startswith("hybrid") and prm.jitter == 0.0) else 1) + (7 if run.frame else 0)
    line = {
        "metric": "lyapunov_giga_iters_per_s", "value": value, "unit": "Giter/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["what"] + (", ONE frame per step: %dx%d tiles round-robin over the GPUs" % (TILE, TILE) if run.frame
                                             else ", ONE volume per step: z-slabs across the GPUs"),
                   "mode": args.mode, "sequence": wl["seq"], "settle": prm.settle, "accum": prm.accum, "jitter": prm.jitter,
                   "l2": "flushed between timed steps (256 MiB device write inside the timed region)",
                   "gather": ("none needed: every rank's kernel stores its shard directly into rank 0's buffer over NVLink "
                              "(CUDA IPC peer mapping); per step inside the timed region: zero of the shared point buffer + barrier, "
                              "kernel, stream sync + barrier") if world > 1 else "none",
                   "limiter_at_n_gpus": "the longest single ray (a strictly sequential chain of ~1.6e6 dependent FP32 steps, ~6.6 ms) and the "
                                        "few rays per lane left at 1/N of a frame -- not a collective" if world > 1 else None},
        "frames_per_s": (args.steps / (ms * 1e-3)) if run.frame else None,
        "evaluations_per_step": total_evals / args.steps, "wall_ms": t["wall_ms"],
        "clocks": t["clocks"], "e2e": e2e, "gpu_launches": n_kernels * args.steps * world,
        "roofline": roofline, "cpu_baseline": cpu, "reference_cuda": ref_cuda, "shard_check": shard_check, "sub": sub,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
