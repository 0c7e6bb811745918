"""The real drop-in, timed (bench.py's hooks_whole_step leg runs this in a child process): the UNMODIFIED reference driver
-- its own sources compiled in place, oracle/_ref/libmisa_ref_cuda.so -- running simulate()'s loop body (reference
src/simulation.cpp:164-194) with atom::latRho / latDf / latForce dispatched to the eight cuda_* hooks of arch_cuda/ and
everything else (Verlet, decide, packers, exchange) as the reference's host code on ONE rank. Prints one JSON object.
usage: python tools/hooks_step.py cells steps r0 r1 r2"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_py as R

cells, steps = int(sys.argv[1]), int(sys.argv[2])
ratio = tuple(int(v) for v in sys.argv[3:6]) if len(sys.argv) >= 6 else (1, 0, 0)
if not R.available(hooks=True):
    print(json.dumps({"unavailable": "oracle/_ref/libmisa_ref_cuda.so not built"}))
    sys.exit(0)
w = R.World((cells,) * 3, grid=(1, 1, 1), a=2.85532, crf=1.96125, dt=0.001, hooks=True)
assert w.L.ref_accelerated() == 1
w.build_world(seed=466953, t_set=600.0, ratio=ratio)
w.prepare()
w.step(1)
t0 = time.perf_counter()
w.step(steps)
dt = (time.perf_counter() - t0) / steps
print(json.dumps({"value": 2 * cells ** 3 / dt, "unit": "atom-steps/s", "ms_per_step": dt * 1e3, "steps": steps,
                  "api": "reference simulate() loop on 1 host thread + cuda_eam_{rho,df,force}_calc hooks (AoS over PCIe per hook)"}))
w.close()
