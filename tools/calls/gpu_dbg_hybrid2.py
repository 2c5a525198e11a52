import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
prm0, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
w, h = 48, 32
lp.scene_cam_recalculate(cam, w, h, 1)
rng = np.random.default_rng(5)
seqs = ["A9B9C9D9", "A9A8B9B9", "ABCDABCDABCDABCDABCDABCDABCDABCDABCDA", "".join("ABC"[i] for i in rng.integers(0, 3, 53)), "BCABA", "A6B6C6"]
def pts(t): return t.cpu().numpy().view(np.float32).reshape(h, w, 9)
for txt in seqs:
    seq = lp.scene_convert_sequence(txt)
    for settle, accum, d in ((7, 333, 3.2), (18, 1008, 3.2), (18, 1008, 2.1)):
        prm = clone(prm0); prm.settle, prm.accum, prm.d, prm.jitter = settle, accum, d, 0.0
        par = lp.render(cam, prm, seq, lights, n, w, h, mode="exact")
        for table in (1, 0):
            api.set_option("seq_table", table)
            for pct in (100, 1000, 100000):
                api.set_option("hybrid_guard_percent", pct)
                hy = lp.render(cam, prm, seq, lights, n, w, h, mode="hybrid")
                a, b = pts(hy[1]), pts(par[1])
                diff = (a.view(np.uint32)[..., :3] != b.view(np.uint32)[..., :3]).any(-1)
                print(txt[:12], len(txt), (settle, accum, d), "table", table, "guard%", pct, "P differs on", int(diff.sum()), "pixels; evals", int(hy[2].item()), int(par[2].item()))
                if diff.any() and pct == 100:
                    y, x = np.argwhere(diff)[0]
                    print("    pixel", x, y, "hybrid P,l", a[y, x, :3], a[y, x, 8], "parity P,l", b[y, x, :3], b[y, x, 8])
api.set_option("hybrid_guard_percent", 100); api.set_option("seq_table", 1)
