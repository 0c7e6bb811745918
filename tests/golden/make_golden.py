"""Generate the committed golden vectors under tests/golden/ from the REFERENCE ITSELF.

The generator drives oracle/_ref/libmisa_ref.so -- the reference's own src/atom.cpp, src/newton_motion.cpp,
src/pack/*, src/atom/*, src/lattice/* compiled in place (oracle/Makefile target `ref`) -- so it only runs in the
authoring container where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures travel (small .npz files); the tests that consume them (tests/test_golden.py) never touch
/root/reference.  Inputs are stored next to the outputs, so neither the oracle nor the CUDA path has to re-derive
them from a generator.

  thermal_alloy.npz   8x7x9 cells, Fe:Cu:Ni 80:12:8, 3 vacancies, sigma 0.05 A: state after prepare()
                      (exchangeAtomFirst + clearForce + computeEam, reference src/simulation.cpp:137-145) and after
                      5 further steps (src/simulation.cpp:164-194)
  thermal_2ranks.npz  12x6x6 cells cut 2x1x1: both sub-boxes after prepare() and after 4 steps
  pka.npz             10^3 cells Fe, setv at [5,5,5,0] direction [1,3,5] 400 eV, dt 2e-4: lattice + inter-atom list
                      after 150 steps (4 inter atoms) (atom::setv/decide/interRho/interForce, src/atom.cpp:21-84,194-494)
                      + the AtomDump::dump byte stream of that state (frontend/io/atom_dump.cpp:39-75)
  world.npz           WorldBuilder::build on one rank, 6x7x8 cells Fe, seed 466953, 600 K (src/world_builder.cpp:64-199)
  index.npz           NeighbourIndex offsets (src/atom/neighbour_index.inl:13-76) and sendlist/recvlist
                      (src/atom/atom_list.cpp:33-40, src/pack/lat_particle_packer.cpp:65-139) of a 7x8x9 sub-box
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from misa_md_b200 import synth  # noqa: E402
from oracle import ref_py  # noqa: E402
from tests import common as cm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("id", "type", "x", "v", "f", "rho", "df")


def ref_world(state, grid=(1, 1, 1), dt=0.001):
    w = ref_py.World(state["phase_space"], grid=grid, a=cm.A, crf=cm.CRF, dt=dt)
    for r in range(w.n_ranks):
        arr, _ = synth.scatter_to_sub_box(state, grid, w.coord(r), cm.CRF)
        w.atoms(r)[:] = arr
    return w


def snap(w, r, prefix, out):
    a = w.atoms(r)
    for f in FIELDS:
        out["%s_%s" % (prefix, f)] = a[f].copy()


def thermal_alloy():
    st = cm.make_state((8, 7, 9), ratio=(80, 12, 8), sigma=0.05, vacancies=3)
    w = ref_world(st)
    out = {"phase_space": np.array(st["phase_space"]), "grid": np.array([1, 1, 1]), "dt": 0.001}
    snap(w, 0, "in", out)
    w.prepare()
    snap(w, 0, "prep", out)
    w.step(5)
    snap(w, 0, "step5", out)
    w.close()
    np.savez_compressed(os.path.join(HERE, "thermal_alloy.npz"), **out)


def thermal_2ranks():
    st = cm.make_state((12, 6, 6), ratio=(90, 6, 4), sigma=0.04)
    w = ref_world(st, grid=(2, 1, 1))
    out = {"phase_space": np.array(st["phase_space"]), "grid": np.array([2, 1, 1]), "dt": 0.001}
    for r in range(2):
        out["coord%d" % r] = np.array(w.coord(r))
        snap(w, r, "in%d" % r, out)
    w.prepare()
    for r in range(2):
        snap(w, r, "prep%d" % r, out)
    w.step(4)
    for r in range(2):
        snap(w, r, "step4_%d" % r, out)
    w.close()
    np.savez_compressed(os.path.join(HERE, "thermal_2ranks.npz"), **out)


def pka():
    st = cm.make_state((10, 10, 10), t_set=300.0)
    w = ref_world(st, dt=2e-4)
    lat, direction, energy, nsteps = (5, 5, 5, 0), (1.0, 3.0, 5.0), 400.0, 150
    out = {"phase_space": np.array(st["phase_space"]), "grid": np.array([1, 1, 1]), "dt": 2e-4,
           "lat": np.array(lat), "direction": np.array(direction), "energy": energy, "nsteps": nsteps}
    snap(w, 0, "in", out)
    w.prepare()
    w.collision_step(lat, direction, energy)
    w.step(nsteps - 1)
    snap(w, 0, "end", out)
    inter = w.inter(0)
    for f in FIELDS:
        out["inter_%s" % f] = inter[f].copy()
    assert inter.size > 0, "the PKA golden must exercise the inter-atom path"
    # AtomDump::dump of the end state, the reference's own frontend/io sources (oracle/shim/ref_dump.cpp)
    out["dump_step"] = 150
    out["dump_bytes"] = np.frombuffer(w.dump(0, 150).tobytes(), dtype=np.uint8)
    w.close()
    np.savez_compressed(os.path.join(HERE, "pka.npz"), **out)


def world():
    """WorldBuilder::build on one rank (reference src/world_builder.cpp:64-199): pure Fe, seed 466953, 600 K."""
    phase = (6, 7, 8)
    w = ref_py.World(phase, a=cm.A, crf=cm.CRF)
    w.build_world(seed=466953, t_set=600.0, ratio=(1, 0, 0))
    got = w.atoms(0).reshape(w.shape(0))[w.owned_slices(0)]
    out = {"phase_space": np.array(phase), "seed": 466953, "t_set": 600.0, "temperature": w.temperature()}
    for f in ("id", "type", "x", "v"):
        out[f] = got[f].copy()
    w.close()
    np.savez_compressed(os.path.join(HERE, "world.npz"), **out)


def index():
    st = cm.make_state((7, 8, 9))
    w = ref_world(st)
    w.prepare()
    out = {"phase_space": np.array(st["phase_space"])}
    for which, name in enumerate(("even", "odd", "half_even", "half_odd")):
        out["off_" + name] = w.offsets(0, which)
    for i in range(6):
        out["send%d" % i] = w.sendlist(0, i)
        out["recv%d" % i] = w.sendlist(0, i, recv=True)
    w.close()
    np.savez_compressed(os.path.join(HERE, "index.npz"), **out)


if __name__ == "__main__":
    if not os.path.isdir("/root/reference/src"):
        sys.exit("the generator needs /root/reference (it runs the reference's own sources)")
    ref_py.build()
    thermal_alloy()
    thermal_2ranks()
    pka()
    world()
    index()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
