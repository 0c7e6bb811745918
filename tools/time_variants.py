"""Dev-only A/B of builds of the same library (nvcc -D variants under build/variants): thermalised 100^3 box, CUDA-event
slots of rho / force and the step time. usage: [RATIO=97,2,1] python tools/time_variants.py lib1.so lib2.so ..."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import sys, os
sys.path.insert(0, os.path.dirname(%r))
import misa_md_b200 as mb
from misa_md_b200 import synth
P = (100, 100, 100)
ratio = tuple(int(v) for v in os.environ.get("RATIO", "1,0,0").split(","))
st = synth.create_global_state(P, ratio=ratio)
ctx = mb.Context(P)
ctx.make_offsets()
ctx.set_potential(*mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH)))
arr, lay = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
ctx.upload(arr)
ctx.prepare()
ctx.step(200)
ms = ctx.timed_steps(50)
ctx.profile_enable(True); ctx.step(20); pr = ctx.profile_read(); ctx.profile_enable(False)
import numpy as np
a = ctx.download()
sl = lambda k: pr[k][0] / max(pr[k][1], 1)
ctx.step(1); ctx.profile_enable(True); [ctx.step(1) for _ in range(10)]; p1 = ctx.profile_read(); ctx.profile_enable(False)   # single-step calls: k_verlet2 runs
print("%%-28s %%.4f ms/step | rho %%.4f force %%.4f verlet1 %%.4f halo_x %%.4f halo_df %%.4f verlet2(1-step calls) %%.4f | checksums f %%.17g x %%.17g" %% (os.path.basename(os.environ.get("MISA_B200_LIB", "default")), ms / 50, sl("rho"), sl("force"), sl("verlet1"), sl("halo_x"), sl("halo_df"), p1["verlet2"][0] / max(p1["verlet2"][1], 1), float(np.abs(a["f"]).sum()), float(np.abs(a["x"]).sum())), flush=True)
''' % here
for lib in [None] + sys.argv[1:]:
    env = dict(os.environ)
    if lib:
        env["MISA_B200_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, "-c", code], env=env)
