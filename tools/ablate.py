"""Dev-only: time the rho / force passes on a STATIC thermally-displaced lattice (no stepping), for A/B of builds
selected with MISA_B200_LIB. usage: python tools/ablate.py [cells] [reps] [sigma]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import misa_md_b200 as mb
from misa_md_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.08
ratio = tuple(int(v) for v in sys.argv[4:7]) if len(sys.argv) > 6 else (1, 0, 0)
P = (n, n, n)
st = synth.create_global_state(P, ratio=ratio)
synth.perturb_positions(st, sigma)
ctx = mb.Context(P)
ctx.make_offsets()
elec, embed, phi = mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH))
ctx.set_potential(elec, embed, phi)
arr, lay = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
ctx.upload(arr)
ctx.prepare()
for fast, tex in ((1, 1), (0, 1), (0, 0)):
    ctx.set_option("fast", fast)
    ctx.set_option("tex", tex)
    for _ in range(3):
        ctx.run_pass("rho"); ctx.run_pass("force")
    ctx.sync()
    ctx.profile_enable(True)
    for _ in range(reps):
        ctx.run_pass("rho"); ctx.run_pass("force")
    ctx.sync()
    pr = ctx.profile_read(); ctx.profile_enable(False)
    print("%s fast=%d tex=%d n_off=%d dmax=%.3f: rho %.4f ms  force %.4f ms" % (os.environ.get("MISA_B200_LIB", "default"), fast, tex, ctx.query("n_off"), ctx.query("dmax"),
          pr["rho"][0] / max(1, pr["rho"][1]), pr["force"][0] / max(1, pr["force"][1])), flush=True)
