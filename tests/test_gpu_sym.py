"""Pair-symmetric rho / force passes (misa_md_b200/csrc/eam_sym.cuh: every near pair evaluated ONCE, its scalar handed
to the partner through the per-pair scratch array) against the full-list kernels and the oracle
(reference src/atom.cpp:151-192, 311-358: the half-list loops whose pairs these are)."""
import numpy as np
import pytest

from tests import common as cm

pytestmark = pytest.mark.gpu
TOL = 1e-10   # BASELINE.json north_star: per-atom rho, df and forces within 1e-10 relative


@pytest.mark.parametrize("shape,ratio,vac,sigma", [
    ((9, 10, 11), (1, 0, 0), 0, 0.07),      # one species, no vacancy: the benchmark path
    ((8, 8, 8), (1, 0, 0), 25, 0.06),       # vacancies: vacant sites hand zeros to their partners
    ((10, 9, 11), (97, 2, 1), 0, 0.06),     # dilute alloy: majority loop symmetric, minority epilogue per atom
    ((10, 9, 11), (92, 5, 3), 7, 0.05),
    ((4, 5, 7), (0, 1, 0), 0, 0.05),        # box narrower than twice the ghost shell
])
def test_symmetric_passes_match_full_list_and_oracle(shape, ratio, vac, sigma):
    st = cm.make_state(shape, ratio=ratio, sigma=sigma, vacancies=vac)
    w = cm.oracle_world(st)
    w.prepare()
    out = {}
    for sym in (1, 0):
        ctx = cm.gpu_context(st)
        ctx.set_option("sym", sym)
        ctx.prepare()
        assert ctx.query("sym") == sym
        assert ctx.query("n_half") * 2 <= ctx.query("n_off")
        first = cm.owned(ctx, ctx.download()).copy()
        ref = cm.owned(ctx, w.atoms(0))
        valid = ref["type"] >= 0
        for fld in ("rho", "df", "f"):
            assert cm.rel_err(first[fld][valid], ref[fld][valid]) < TOL, (sym, fld)
        assert np.all(first["rho"][~valid] == 0.0) and np.all(first["f"][~valid] == 0.0)
        ctx.step(4)
        out[sym] = cm.owned(ctx, ctx.download()).copy()
        ctx.close()
    for _ in range(4):
        w.step()
    ref = cm.owned(ctx, w.atoms(0))
    valid = ref["type"] >= 0
    for sym in (1, 0):
        assert np.array_equal(out[sym]["type"], ref["type"])
        assert cm.rel_err(out[sym]["rho"][valid], ref["rho"][valid]) < TOL
        assert cm.rel_err(out[sym]["f"][valid], ref["f"][valid]) < 1e-9
    assert cm.rel_err(out[1]["f"][valid], out[0]["f"][valid]) < 1e-11
    assert cm.rel_err(out[1]["x"][valid], out[0]["x"][valid]) < 1e-13
    w.close()


def test_symmetric_passes_are_deterministic_and_pipelined_equals_serial():
    """Gathers in a fixed order, no atomics: two runs give identical bits, and so do the sync-free and the serial step."""
    st = cm.make_state((11, 10, 12), sigma=0.03)
    out = []
    for pipe in (1, 1, 0):
        ctx = cm.gpu_context(st)
        ctx.set_option("pipe", pipe)
        ctx.set_option("sym", 1)
        ctx.prepare()
        assert ctx.query("sym") == 1
        ctx.step(5)
        out.append(cm.owned(ctx, ctx.download()).copy())
        ctx.close()
    for other in out[1:]:
        for fld in ("x", "v", "f", "rho", "df"):
            assert np.array_equal(out[0][fld], other[fld]), fld


def test_symmetric_passes_close_pair_below_staged_range():
    """A pair closer than the staged r range is recomputed from the global rows by the site that owns it, and the
    exact scalar reaches the partner through the scratch array."""
    st = cm.make_state((8, 8, 8), sigma=0.0)
    x = st["x"]
    x[4, 4, 8] += (x[4, 4, 9] - x[4, 4, 8]) * 0.22
    x[2, 3, 5] += (x[2, 3, 4] - x[2, 3, 5]) * 0.22     # the same towards a LOWER neighbour
    w = cm.oracle_world(st)
    w.prepare()
    ctx = cm.gpu_context(st)
    ctx.set_option("sym", 1)
    ctx.prepare()
    assert ctx.query("sym") == 1
    got, ref = cm.owned(ctx, ctx.download()), cm.owned(ctx, w.atoms(0))
    for fld in ("rho", "df", "f"):
        assert cm.rel_err(got[fld], ref[fld]) < TOL, fld
    ctx.close()
    w.close()
