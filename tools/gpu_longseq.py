"""Long-period sequences (A9B9C9D9 = 40 symbols, and a 53-symbol random one): shared-memory multiplier table
against the run-length loop, bake 256^3 and a 960x540 frame, as fractions of the measured peaks."""
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
w, h = 960, 540
lp.scene_cam_recalculate(cam, w, h, 1)
dl = api.upload_lights(lights)
peaks = api.probe_peaks()
it = prm.settle + prm.accum
peak = {"exact": peaks["mufu_lane_ops_per_s"] / (prm.accum / it), "fast": peaks["ffma_lane_ops_per_s"] / ((2.0 * prm.settle + 4.0 * prm.accum) / it)}
rng = np.random.default_rng(5)
seqs = {"BCABA (register table, for reference)": "BCABA", "A9B9C9D9 (register table P=40)": "A9B9C9D9", "A9A8B9B9 (39 symbols)": "A9A8B9B9", "random53": "".join("ABC"[i] for i in rng.integers(0, 3, 53))}
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best, r
for name, txt in seqs.items():
    seq = lp.scene_convert_sequence(txt)
    for mode in ("exact", "fast"):
        for table in ((1, 0) if api.plan_period(seq, prm.settle, prm.accum) == 0 else (1,)):
            api.set_option("seq_table", table)
            vol = torch.empty((256, 256, 256), dtype=torch.float32, device="cuda")
            tb, _ = timed(lambda: lp.bake(prm, seq, 256, mode=mode, out=vol))
            tf, r = timed(lambda: lp.render(cam, prm, seq, dl, n, w, h, mode=mode))
            ev = int(r[2].item())
            gb, gf = 256 ** 3 * it / tb, ev * it / tf
            print(json.dumps({"sequence": name, "mode": mode, "seq_table": table, "bake256_ms": round(tb * 1e3, 2), "bake_giter_s": round(gb / 1e9),
                              "bake_frac": round(gb / peak[mode], 3), "frame_ms": round(tf * 1e3, 2), "frame_giter_s": round(gf / 1e9), "frame_frac": round(gf / peak[mode], 3)}), flush=True)
api.set_option("seq_table", 1)
