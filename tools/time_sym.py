"""Scratch comparison of the pair-symmetric passes against the full-list kernels on one thermalised box.
usage: python tools/time_sym.py [cells=100] [equil=200] [r0 r1 r2]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import misa_md_b200 as mb
from misa_md_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
equil = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ratio = tuple(int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (1, 0, 0)
P = (n, n, n)
st = synth.create_global_state(P, ratio=ratio)
ctx = mb.Context(P)
ctx.make_offsets()
ctx.set_potential(*mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH)))
arr, lay = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
ctx.upload(arr)
ctx.prepare()
ctx.step(equil)
print("dmax %.3f A, n_off %d of %d, novac %d, n_half %d" % (ctx.query("dmax"), ctx.query("n_off"), ctx.query("n_full"), ctx.query("novac"), ctx.query("n_half")), flush=True)
for sym, prune in ((1, 1), (0, 1), (1, 1), (0, 1), (0, 0)):
    ctx.set_option("sym", sym)
    ctx.set_option("prune", prune)
    ctx.step(3)
    ms = ctx.timed_steps(20)
    ctx.profile_enable(True); ctx.step(10); pr = ctx.profile_read(); ctx.profile_enable(False)
    print("sym=%d prune=%d: %.3f ms/step  %.3e atom-steps/s | rho %.3f force %.3f verlet1 %.3f ms" % (
        sym, prune, ms / 20, ctx.n_owned * 20 / (ms * 1e-3), pr["rho"][0] / pr["rho"][1], pr["force"][0] / pr["force"][1],
        pr["verlet1"][0] / pr["verlet1"][1]), flush=True)
