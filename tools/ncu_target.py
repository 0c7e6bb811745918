"""Target for ncu captures: thermalise a bcc Fe box, then run a few production steps.
usage: python tools/ncu_target.py [cells=100] [equil=200] [steps=3] [ratio r0 r1 r2]
With ncu:  -k regex:'k_(force|rho)_s' --launch-skip 2*equil+2 --launch-count 2"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import misa_md_b200 as mb
from misa_md_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
equil = int(sys.argv[2]) if len(sys.argv) > 2 else 200
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ratio = tuple(int(v) for v in sys.argv[4:7]) if len(sys.argv) > 6 else (1, 0, 0)
P = (n, n, n)
st = synth.create_global_state(P, ratio=ratio)
ctx = mb.Context(P)
ctx.make_offsets()
elec, embed, phi = mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH))
ctx.set_potential(elec, embed, phi)
arr, lay = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
ctx.upload(arr)
ctx.prepare()
ctx.step(equil)
print("dmax %.3f A, n_off %d of %d, novac %d" % (ctx.query("dmax"), ctx.query("n_off"), ctx.query("n_full"), ctx.query("novac")), flush=True)
ms = ctx.timed_steps(steps)
print("%.3f ms/step" % (ms / steps))
