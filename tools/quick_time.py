"""Scratch timing of the resident step on one GPU (not the bench contract; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import misa_md_b200 as mb
from misa_md_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ratio = tuple(int(v) for v in sys.argv[4:7]) if len(sys.argv) > 6 else (1, 0, 0)
P = (n, n, n)
t0 = time.time()
st = synth.create_global_state(P, ratio=ratio)
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
equil = int(sys.argv[3]) if len(sys.argv) > 3 else 300
ctx.step(equil)  # let the lattice thermalise so that dmax (=> the pruned stencil level) is stationary
print("after %d steps: dmax %.3f A, n_off %d of %d, novac %d" % (equil, ctx.query("dmax"), ctx.query("n_off"), ctx.query("n_full"), ctx.query("novac")), flush=True)
for opt in ((1, 1, 1, 0, 0), (1, 1, 1, 0, 1), (1, 1, 1, 1, 0), (1, 1, 1, 1, 1), (0, 1, 1, 0, 1), (1, 1, 0, 0, 0)):
    for k, name in enumerate(("prune", "fuse", "smem", "tex", "novac")):
        ctx.set_option(name, opt[k])
    ctx.step(2)
    ms = ctx.timed_steps(steps)
    print("prune=%d fuse=%d smem=%d tex=%d novac=%d: %.3f ms/step  %.3e atom-steps/s" % (opt + (ms / steps, ctx.n_owned * steps / (ms * 1e-3))), flush=True)
    ctx.profile_enable(True); ctx.step(5); pr = ctx.profile_read(); ctx.profile_enable(False)
    print("     rho %.3f ms  force %.3f ms  verlet1 %.3f ms  (n_off %d, dmax %.3f A)" % (pr["rho"][0] / pr["rho"][1], pr["force"][0] / pr["force"][1], pr["verlet1"][0] / pr["verlet1"][1], ctx.query("n_off"), ctx.query("dmax")), flush=True)
for k, name in enumerate(("prune", "fuse", "smem", "tex", "novac")):
    ctx.set_option(name, (1, 1, 1, 0, 1)[k])
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
