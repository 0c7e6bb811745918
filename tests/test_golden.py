"""Golden vectors produced by the REFERENCE'S OWN sources (tests/golden/make_golden.py drives
oracle/_ref/libmisa_ref.so in the authoring container) -- consumed here without /root/reference.

CPU (`-m "not gpu"`): the C restatement (oracle/md_oracle.c) reproduces every stored field bit for bit, and the
product's host-side planners reproduce the stored index vectors value for value.
GPU (`-m gpu`): the CUDA path, through the C ABI, against the same files: integer state exact, rho/df/force within
the north-star's 1e-10 relative."""
import ctypes as C
import os

import numpy as np
import pytest

from misa_md_b200 import capi, synth
from oracle import oracle_py as O
from tests import common as cm

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("id", "type", "x", "v", "f", "rho", "df")
TOL = 1e-10  # BASELINE.json north_star: per-atom rho, df, forces within 1e-10 relative


def load(name):
    return np.load(os.path.join(G, name))


def aos(g, prefix):
    arr = np.zeros(g[prefix + "_id"].size, dtype=synth.ATOM_DTYPE)
    for f in FIELDS:
        arr[f] = g["%s_%s" % (prefix, f)]
    return arr


def assert_bit_identical(got, g, prefix):
    for f in FIELDS:
        assert np.array_equal(got[f], g["%s_%s" % (prefix, f)]), (prefix, f)


# ------------------------------------------------------------------------------------------- CPU: oracle
def test_oracle_reproduces_thermal_alloy_golden():
    g = load("thermal_alloy.npz")
    w = O.World(tuple(g["phase_space"]), a=cm.A, crf=cm.CRF, dt=float(g["dt"]))
    w.atoms(0)[:] = aos(g, "in")
    w.prepare()
    assert_bit_identical(w.atoms(0), g, "prep")
    for _ in range(5):
        w.step()
    assert_bit_identical(w.atoms(0), g, "step5")
    w.close()


def test_oracle_reproduces_two_rank_golden():
    g = load("thermal_2ranks.npz")
    w = O.World(tuple(g["phase_space"]), grid=tuple(g["grid"]), a=cm.A, crf=cm.CRF, dt=float(g["dt"]))
    ranks = {tuple(w.rank(r).dom.grid_coord): r for r in range(2)}
    order = [ranks[tuple(g["coord%d" % r])] for r in range(2)]
    for r in range(2):
        w.atoms(order[r])[:] = aos(g, "in%d" % r)
    w.prepare()
    for r in range(2):
        assert_bit_identical(w.atoms(order[r]), g, "prep%d" % r)
    for _ in range(4):
        w.step()
    for r in range(2):
        assert_bit_identical(w.atoms(order[r]), g, "step4_%d" % r)
    w.close()


def test_oracle_reproduces_pka_golden():
    g = load("pka.npz")
    w = O.World(tuple(g["phase_space"]), a=cm.A, crf=cm.CRF, dt=float(g["dt"]))
    w.atoms(0)[:] = aos(g, "in")
    w.prepare()
    lat, d = [int(v) for v in g["lat"]], [float(v) for v in g["direction"]]
    w.L.ora_collision_step(w.h, C.byref((C.c_int * 4)(*lat)), C.byref((C.c_double * 3)(*d)), float(g["energy"]))
    for _ in range(int(g["nsteps"]) - 1):
        w.step()
    assert_bit_identical(w.atoms(0), g, "end")
    inter = w.inter(0)
    assert inter.size == g["inter_id"].size > 0
    assert_bit_identical(inter, g, "inter")
    w.close()


def test_host_planners_reproduce_index_golden():
    g = load("index.npz")
    dom = capi.make_domain(tuple(g["phase_space"]), (1, 1, 1), (0, 0, 0), cm.A, cm.CRF)
    for which, name in enumerate(("even", "odd", "half_even", "half_odd")):
        assert np.array_equal(capi.plan_offsets(dom, which), g["off_" + name])
    for dim in range(3):
        for direction in range(2):
            send, recv, _ = capi.plan_halo(dom, dim, direction)
            assert np.array_equal(send, g["send%d" % (2 * dim + direction)])
            assert np.array_equal(recv, g["recv%d" % (2 * dim + direction)])


# ------------------------------------------------------------------------------------------- GPU: C ABI
def gpu_ctx(g, prefix="in", grid=(1, 1, 1), coord=(0, 0, 0)):
    import misa_md_b200 as mb
    ctx = mb.Context(tuple(int(v) for v in g["phase_space"]), grid=grid, coord=coord, a=cm.A, crf=cm.CRF)
    ctx.make_offsets()
    ctx.set_potential(*cm.host_potential())
    ctx.set_timestep(float(g["dt"]))
    ctx.upload(aos(g, prefix))
    return ctx


def compare_gpu(ctx, got, g, prefix, xtol, ftol):
    got, ref = cm.owned(ctx, got), cm.owned(ctx, aos(g, prefix))
    assert np.array_equal(got["type"], ref["type"])
    valid = ref["type"] >= 0
    assert np.array_equal(got["id"][valid], ref["id"][valid])
    if xtol == 0:
        assert np.array_equal(got["x"][valid], ref["x"][valid])
        assert np.array_equal(got["v"][valid], ref["v"][valid])
    else:
        assert cm.rel_err(got["x"][valid], ref["x"][valid]) < xtol
        assert cm.rel_err(got["v"][valid], ref["v"][valid]) < xtol
    for f in ("rho", "df", "f"):
        assert cm.rel_err(got[f][valid], ref[f][valid]) < ftol, f


@pytest.mark.gpu
def test_gpu_thermal_alloy_golden():
    g = load("thermal_alloy.npz")
    ctx = gpu_ctx(g)
    ctx.prepare()
    got = ctx.download()
    # ghost images (type + shifted x) are integer/byte work: exact, ghosts included
    assert np.array_equal(got["type"], g["prep_type"])
    assert np.array_equal(got["x"], g["prep_x"])
    compare_gpu(ctx, got, g, "prep", 0, TOL)
    ctx.step(5)
    compare_gpu(ctx, ctx.download(), g, "step5", TOL, TOL)
    ctx.close()


@pytest.mark.gpu
def test_gpu_pka_golden():
    g = load("pka.npz")
    ctx = gpu_ctx(g)
    ctx.prepare()
    ctx.collision_step([int(v) for v in g["lat"]], [float(v) for v in g["direction"]], float(g["energy"]))
    ctx.step(int(g["nsteps"]) - 1)
    # 150 steps of a cascade amplify last-bit force differences: trajectory bars as in test_gpu_inter.py
    compare_gpu(ctx, ctx.download(), g, "end", 1e-8, 1e-6)
    inter = ctx.download_inter()
    assert np.array_equal(inter["id"], g["inter_id"])      # same atoms, same list order
    assert np.array_equal(inter["type"], g["inter_type"])
    assert cm.rel_err(inter["x"], g["inter_x"]) < 1e-8
    ctx.close()


# --------------------------------------------------------------------- dump record stream / world builder (8f)
def test_oracle_dump_reproduces_reference_byte_stream():
    """ora_dump (restatement of AtomDump::dump + BufferedFileWriter::write) against the byte stream the reference's
    own frontend/io sources produced for the PKA end state (4 inter atoms, 4 vacancies)."""
    g = load("pka.npz")
    w = O.World(tuple(g["phase_space"]), a=cm.A, crf=cm.CRF, dt=float(g["dt"]))
    w.atoms(0)[:] = aos(g, "end")
    inter = aos(g, "inter")
    w.L.ora_test_set_inter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    w.L.ora_test_set_inter(w.h, 0, inter.ctypes.data, len(inter))
    rec = w.dump(0, int(g["dump_step"]))
    assert rec.size == 2000 and rec.dtype.itemsize == 72
    assert rec.tobytes() == g["dump_bytes"].tobytes()
    assert np.array_equal(rec["id"][:len(inter)], inter["id"])  # inter atoms first, list order
    w.close()


def test_host_world_mirror_matches_reference_world_builder():
    """synth.create_global_state (host mirror of csrc/world.cuh) against the stored output of the reference's
    WorldBuilder::build: ids, species, positions exact; velocities to summation rounding."""
    g = load("world.npz")
    st = synth.create_global_state(tuple(g["phase_space"]), seed=int(g["seed"]), t_set=float(g["t_set"]))
    assert np.array_equal(st["id"], g["id"])
    assert np.array_equal(st["type"], g["type"])
    assert np.array_equal(st["x"], g["x"])
    assert np.allclose(st["v"], g["v"], rtol=1e-11, atol=1e-13)


@pytest.mark.gpu
def test_gpu_dump_golden():
    g = load("pka.npz")
    ctx = gpu_ctx(g, prefix="end")
    ctx.upload_inter(aos(g, "inter"))
    rec = ctx.dump_records(int(g["dump_step"]))
    assert rec.tobytes() == g["dump_bytes"].tobytes()   # byte work: identical, padding included
    ctx.close()


@pytest.mark.gpu
def test_gpu_build_world_golden():
    import misa_md_b200 as mb
    g = load("world.npz")
    ctx = mb.Context(tuple(int(v) for v in g["phase_space"]), a=cm.A, crf=cm.CRF)
    ctx.build_world(seed=int(g["seed"]), t_set=float(g["t_set"]), ratio=(1, 0, 0))
    got = cm.owned(ctx, ctx.download())
    assert np.array_equal(got["id"], g["id"])
    assert np.array_equal(got["type"], g["type"])
    assert np.array_equal(got["x"], g["x"])
    assert np.allclose(got["v"], g["v"], rtol=1e-11, atol=1e-13)  # reductions sum in a different order
    assert abs(ctx.temperature()["T"] - float(g["temperature"])) < 1e-9
    ctx.close()
