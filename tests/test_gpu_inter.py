"""GPU parity on the off-lattice ("inter") atom path: PKA kick, run-away detection (atom::decide), inter-atom
rho/force, vacancy re-occupation -- BASELINE.json configs[3] at test size. Occupancy, ids and the inter-atom set
must be exact; positions/velocities/forces within the floating-point bars."""
import numpy as np
import pytest

from tests import common as cm

pytestmark = pytest.mark.gpu


def run_cascade(phase, energy, direction, lat, steps, dt, check_every=0):
    st = cm.make_state(phase)
    w = cm.oracle_world(st, dt=dt)
    w.prepare()
    ctx = cm.gpu_context(st, dt=dt)
    ctx.prepare()
    import ctypes as C
    w.L.ora_collision_step(w.h, C.byref((C.c_int * 4)(*lat)), C.byref((C.c_double * 3)(*direction)), energy)
    ctx.collision_step(lat, direction, energy)
    seen_inter = 0
    for s in range(steps):
        w.step()
        ctx.step(1)
        if check_every and (s + 1) % check_every == 0:
            th = ctx.thermo()
            assert th["n_inter"] == w.total_inter(), s
        seen_inter = max(seen_inter, w.total_inter())
    return st, w, ctx, seen_inter


def compare(w, ctx, xtol=1e-9, ftol=1e-7):
    got = cm.owned(ctx, ctx.download())
    ref = cm.owned(ctx, w.atoms(0))
    assert np.array_equal(got["type"], ref["type"])           # vacancies exactly where the reference has them
    valid = ref["type"] >= 0
    assert np.array_equal(got["id"][valid], ref["id"][valid])   # and the same atoms on the same sites
    assert cm.rel_err(got["x"][valid], ref["x"][valid]) < xtol
    assert cm.rel_err(got["v"][valid], ref["v"][valid]) < 1e-6
    assert cm.rel_err(got["f"][valid], ref["f"][valid]) < ftol
    gi, ri = ctx.download_inter(), w.inter(0)
    assert len(gi) == len(ri)
    if len(ri):
        assert np.array_equal(gi["id"], ri["id"])               # same inter atoms, same list order
        assert np.array_equal(gi["type"], ri["type"])
        assert cm.rel_err(gi["x"], ri["x"]) < xtol
        assert cm.rel_err(gi["f"], ri["f"]) < ftol
        assert cm.rel_err(gi["rho"], ri["rho"]) < 1e-9


def test_pka_creates_inter_atoms_and_tracks_oracle():
    st, w, ctx, seen = run_cascade((10, 10, 10), 300.0, (1.0, 3.0, 5.0), (5, 5, 5, 0), steps=120, dt=2e-4, check_every=10)
    assert seen > 0, "the kick must drive at least one atom off its site"
    compare(w, ctx)
    th = ctx.thermo()
    assert abs(th["mvv"] - w.L.ora_mvv(w.h)) / th["mvv"] < 1e-8
    assert abs(th["pe"] - w.potential_energy()) / abs(th["pe"]) < 1e-9
    ctx.close()
    w.close()


def test_pka_crossing_the_periodic_boundary():
    """A PKA near the box face: inter atoms migrate through the periodic image (exchangeInter with the self
    neighbour) and ghost inter atoms contribute across the boundary (borderInter)."""
    st, w, ctx, seen = run_cascade((8, 8, 8), 400.0, (-3.0, -1.0, -0.5), (0, 1, 1, 0), steps=150, dt=2e-4, check_every=25)
    assert seen > 0
    compare(w, ctx)
    ctx.close()
    w.close()


def test_uploaded_inter_atoms_static_parity():
    """rho / df / force with interstitials and vacancies present from the start (no dynamics): one prepare()."""
    st = cm.make_state((8, 8, 8), ratio=(90, 6, 4), sigma=0.04, vacancies=12)
    rs = np.random.RandomState(5)
    from misa_md_b200 import synth
    inter = np.zeros(9, dtype=synth.ATOM_DTYPE)
    inter["id"] = 100000 + np.arange(9)
    inter["type"] = rs.randint(0, 3, size=9)
    # octahedral-like interstitial positions well inside the box, >1.2 A from lattice atoms
    cells = rs.randint(1, 7, size=(9, 3))
    inter["x"] = (cells + np.array([0.5, 0.0, 0.03])) * cm.A
    w = cm.oracle_world(st)
    rk = w.rank(0)
    import ctypes as C
    buf = (C.c_char * (len(inter) * 104)).from_buffer_copy(inter.tobytes())
    # hand the same list to the oracle: it owns realloc'd storage, so append through its own container
    lib = w.L
    lib.ora_test_set_inter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.ora_test_set_inter(w.h, 0, buf, len(inter))
    w.prepare()
    ctx = cm.gpu_context(st)
    ctx.upload_inter(inter)
    ctx.prepare()
    assert w.total_inter() == 9
    compare(w, ctx, ftol=1e-9)
    ctx.close()
    w.close()


@pytest.mark.parametrize("phase,lat,direction,energy,steps", [((10, 10, 10), (5, 5, 5, 0), (1.0, 3.0, 5.0), 300.0, 120),
                                                               ((8, 8, 8), (0, 1, 1, 0), (-3.0, -1.0, -0.5), 400.0, 150)])
def test_device_resident_list_equals_the_host_list(phase, lat, direction, energy, steps):
    """csrc/inter_dev.cuh (Wigner-Seitz mapping, decide, exchangeInter / borderInter packers and the list itself on the
    device) against csrc/inter.cuh (the reference's list kept on the host): the same cascade: occupancy, ids and the list
    order identical, positions / velocities / sums to accumulated round-off. The second case drives atoms through the periodic faces (packers with the image shift)."""
    st = cm.make_state(phase)
    out = []
    for dev in (1, 0):
        ctx = cm.gpu_context(st, dt=2e-4)
        ctx.set_option("inter_dev", dev)
        ctx.prepare()
        ctx.collision_step(lat, direction, energy)
        seen = 0
        for s in range(steps):
            ctx.step(1)
            seen = max(seen, ctx.thermo()["n_inter"])
        out.append((cm.owned(ctx, ctx.download()).copy(), ctx.download_inter(), seen, ctx.thermo()))
        ctx.close()
    (la, ia, sa, ta), (lb, ib, sb, tb) = out
    assert sa > 0 and sa == sb
    for fld in ("type", "id"):
        assert np.array_equal(la[fld], lb[fld]), fld
    assert len(ia) == len(ib) and np.array_equal(ia["id"], ib["id"]) and np.array_equal(ia["type"], ib["type"])
    # (inter atoms add to lattice sites with atomics on both paths: the order of two contributions to one site is not fixed,
    # so positions agree to accumulated round-off, not bit for bit)
    valid = lb["type"] >= 0
    assert cm.rel_err(la["x"][valid], lb["x"][valid]) < 1e-11 and cm.rel_err(la["v"][valid], lb["v"][valid]) < 1e-8
    if len(ia):
        assert cm.rel_err(ia["x"], ib["x"]) < 1e-11 and cm.rel_err(ia["v"], ib["v"]) < 1e-8
    for fld in ("f", "rho", "df"):
        assert cm.rel_err(la[fld][valid], lb[fld][valid]) < 1e-8, fld
        if len(ia):
            assert cm.rel_err(ia[fld], ib[fld]) < 1e-8, fld
    assert abs(ta["mvv"] - tb["mvv"]) <= 1e-10 * abs(tb["mvv"]) and abs(ta["pe"] - tb["pe"]) <= 1e-10 * abs(tb["pe"])
