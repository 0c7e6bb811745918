"""CPU arm of bench.py (TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE).

Steps the reference algorithm for the hot path on the host cores: the reference's own sources compiled in
place (oracle/_ref/libmisa_ref.so, kind "reference") when that library was built in the authoring container,
else the plain-C restatement (oracle/liboracle.so, kind "port"). One sub-box per thread, the staged
neighbour exchange done in process ("MPI-equivalent in-process exchange": no MPI exists on the box).
Only bench.py's cpu_baseline / --impl reference legs and tests/ import this module.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class _PortArm:
    kind = "port"

    def __init__(self, cells, grid, a, crf, dt):
        from misa_md_b200 import synth
        from . import oracle_py as O
        self.grid = tuple(grid)
        self.threads = grid[0] * grid[1] * grid[2]
        st = synth.create_global_state((cells,) * 3, a=a)
        self.w = O.World((cells,) * 3, grid=self.grid, a=a, crf=crf, dt=dt, threads=self.threads)
        for r in range(self.w.n_ranks):
            d = self.w.rank(r).dom
            arr, _ = synth.scatter_to_sub_box(st, self.grid, tuple(d.grid_coord), crf)
            self.w.atoms(r)[:] = arr
        self.w.prepare()

    def step(self):
        self.w.step()

    def close(self):
        self.w.close()


class _RefArm(_PortArm):
    kind = "reference"

    def __init__(self, cells, grid, a, crf, dt):
        from misa_md_b200 import synth
        from . import ref_py as R
        self.grid = tuple(grid)
        self.threads = grid[0] * grid[1] * grid[2]
        st = synth.create_global_state((cells,) * 3, a=a)
        self.w = R.World((cells,) * 3, grid=self.grid, a=a, crf=crf, dt=dt, threads=self.threads)
        for r in range(self.w.n_ranks):
            arr, _ = synth.scatter_to_sub_box(st, self.grid, self.w.coord(r), crf)
            self.w.atoms(r)[:] = arr
        self.w.prepare()


def make(cells, grid, a, crf, dt, prefer_ref=True):
    if prefer_ref and os.path.exists(os.path.join(HERE, "_ref", "libmisa_ref.so")) and os.path.exists(os.path.join(HERE, "ref_py.py")):
        try:
            return _RefArm(cells, grid, a, crf, dt)
        except Exception as e:  # a stale/unloadable prebuilt .so: fall back to the port, visibly
            import sys
            print("cpu_arm: oracle/_ref unusable (%s); using the port" % e, file=sys.stderr)
    return _PortArm(cells, grid, a, crf, dt)
