import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
prm0, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
w, h = 48, 32
lp.scene_cam_recalculate(cam, w, h, 1)
def pts(t): return t.cpu().numpy().view(np.uint32).reshape(h, w, 9)
def cmp(a, b):
    pa, pb = pts(a[1]), pts(b[1])
    return {"rgba_eq": float((a[0] == b[0]).all(dim=-1).float().mean()), "pts_fields_eq": [(float((pa[..., k] == pb[..., k]).mean())) for k in range(9)], "evals": (int(a[2].item()), int(b[2].item()))}
# 1. short sequence, forced generic, vs register path
seq = lp.scene_convert_sequence("BCABA")
prm = clone(prm0); prm.settle, prm.accum = 7, 333
ref = lp.render(cam, prm, seq, lights, n, w, h, mode="exact")
api.set_option("force_generic", 1)
for table in (1, 0):
    api.set_option("seq_table", table)
    for tc in (1, 0):
        api.set_option("tail_compaction", tc)
        g = lp.render(cam, prm, seq, lights, n, w, h, mode="exact")
        print("BCABA forced generic table", table, "tailc", tc, json.dumps(cmp(g, ref)))
api.set_option("force_generic", 0); api.set_option("tail_compaction", 1)
for txt in ("A9B9C9D9", "ABCDABCDABCDABCDABCDABCDABCDABCDABCDA"):
    seq = lp.scene_convert_sequence(txt)
    prm.d = 3.2
    out = {}
    for table in (1, 0):
        api.set_option("seq_table", table)
        out[table] = lp.render(cam, prm, seq, lights, n, w, h, mode="exact")
        again = lp.render(cam, prm, seq, lights, n, w, h, mode="exact")
        print(txt[:10], "table", table, "repeatable:", json.dumps(cmp(out[table], again)))
    print(txt[:10], "table vs runlength:", json.dumps(cmp(out[1], out[0])))
    # evaluator itself on the frame's hit points
    P = out[0][1].view(torch.float32).reshape(-1, 9)[:, :3].contiguous()
    api.set_option("seq_table", 1); e1 = lp.exponent_points(P, prm, seq, mode="exact")
    api.set_option("seq_table", 0); e0 = lp.exponent_points(P, prm, seq, mode="exact")
    print("   exponent at hit points equal:", float((e1.view(torch.int32) == e0.view(torch.int32)).float().mean()))
api.set_option("seq_table", 1)
