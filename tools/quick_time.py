"""Scratch timing of the resident step on one GPU (not the bench contract; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import misa_md_b200 as mb
from misa_md_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
P = (n, n, n)
t0 = time.time()
st = synth.create_global_state(P)
print("state built %.1fs" % (time.time() - t0), flush=True)
ctx = mb.Context(P)
ctx.make_offsets()
elec, embed, phi = mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH))
ctx.set_potential(elec, embed, phi)
arr, lay = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
ctx.upload(arr)
ctx.prepare()
th0 = ctx.thermo()
e0 = 0.5 * th0["mvv"] * synth.MVV2E + th0["pe"]
ctx.step(3)
for opt in ((1, 1, 1), (0, 1, 1), (1, 1, 0)):
    ctx.set_option("prune", opt[0]); ctx.set_option("fuse", opt[1]); ctx.set_option("smem", opt[2])
    ctx.step(2)
    ms = ctx.timed_steps(steps)
    print("prune=%d fuse=%d smem=%d: %.3f ms/step  %.3e atom-steps/s" % (opt[0], opt[1], opt[2], ms / steps, ctx.n_owned * steps / (ms * 1e-3)), flush=True)
ctx.set_option("prune", 1); ctx.set_option("fuse", 1); ctx.set_option("smem", 1)
ctx.profile_enable(True)
ctx.step(steps)
pr = ctx.profile_read()
ctx.profile_enable(False)
for k, (ms, cnt) in pr.items():
    if cnt:
        print("  %-8s %8.3f ms/launch x %d" % (k, ms / cnt, cnt))
th = ctx.thermo()
e1 = 0.5 * th["mvv"] * synth.MVV2E + th["pe"]
print("E0 %.6f E1 %.6f drift/atom %.3e eV  runaways %d inter %d" % (e0, e1, (e1 - e0) / ctx.n_owned, th["runaways"], th["n_inter"]))
