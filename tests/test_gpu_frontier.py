"""GPU parity for the rows SURVEY.md section 8(f) ranks next to the hot path: the initial state built on the device
by global atom id, the dump record stream compacted on the device, and the global temperature / rescale entry points
of the stage machine. All through the C ABI, against the oracle or the host mirror on identical inputs."""
import numpy as np
import pytest

import misa_md_b200 as mb
from misa_md_b200 import synth
from tests import common as cm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("phase,grid,coord,ratio", [((9, 8, 7), (1, 1, 1), (0, 0, 0), (1, 0, 0)),
                                                    ((9, 8, 7), (1, 1, 1), (0, 0, 0), (90, 6, 4)),
                                                    ((12, 8, 8), (2, 1, 1), (1, 0, 0), (97, 2, 1)),
                                                    ((12, 12, 8), (2, 2, 1), (0, 1, 0), (0, 1, 0))])
def test_build_world_matches_host_mirror(phase, grid, coord, ratio):
    """Device WorldBuilder vs synth.create_global_state cut to the sub-box: integer state, positions and the raw
    mt19937 stream exact (every sub-box sees the SAME global state); velocities to reduction rounding."""
    st = synth.create_global_state(phase, a=cm.A, seed=466953, t_set=600.0, ratio=ratio, alloy_seed=77)
    want, _ = synth.scatter_to_sub_box(st, grid, coord, cm.CRF)
    ctx = mb.Context(phase, grid=grid, coord=coord, a=cm.A, crf=cm.CRF)
    ctx.build_world(seed=466953, t_set=600.0, ratio=ratio, alloy_seed=77)
    got = ctx.download()
    assert np.array_equal(got["id"], want["id"])
    assert np.array_equal(got["type"], want["type"])       # ghosts INVALID until the first exchange
    assert np.array_equal(got["x"], want["x"])
    assert np.allclose(got["v"], want["v"], rtol=1e-11, atol=1e-13)
    assert not got["f"].any() and not got["rho"].any()
    ctx.close()


def test_build_world_without_rescale_is_bit_exact_per_atom():
    """t_set = 0 leaves only (u - 0.5)/m - vcm/m: with the host mirror's vcm the per-atom arithmetic must agree to
    the last bit wherever vcm rounds identically; the raw draws themselves are checked exactly via v*m + 0.5."""
    phase = (8, 8, 8)
    ctx = mb.Context(phase, a=cm.A, crf=cm.CRF)
    ctx.build_world(seed=12345, t_set=0.0, ratio=(1, 0, 0))
    got = cm.owned(ctx, ctx.download())
    u = synth.mt19937_unit(12345, 3 * 2 * 8 * 8 * 8).reshape(got["v"].shape)
    raw = (u - 0.5) / synth.MASS[0]
    vcm = (raw * synth.MASS[0]).reshape(-1, 3).sum(axis=0) / (2 * 8 * 8 * 8)
    assert np.allclose(got["v"], raw - vcm / synth.MASS[0], rtol=0, atol=1e-17)
    ctx.close()


def test_world_built_on_device_steps_like_the_uploaded_one():
    phase = (9, 9, 9)
    st = synth.create_global_state(phase, a=cm.A, ratio=(90, 6, 4), alloy_seed=5)
    a = cm.gpu_context(st)
    b = cm.gpu_context(st, upload=False)
    b.build_world(ratio=(90, 6, 4), alloy_seed=5)
    for ctx in (a, b):
        ctx.prepare()
        ctx.step(3)
    ga, gb = cm.owned(a, a.download()), cm.owned(b, b.download())
    assert np.array_equal(ga["type"], gb["type"])
    assert cm.rel_err(gb["x"], ga["x"]) < 1e-12
    assert cm.rel_err(gb["f"], ga["f"]) < 1e-9
    a.close()
    b.close()


@pytest.mark.parametrize("vacancies", [0, 37])
def test_dump_records_match_oracle(vacancies):
    st = cm.make_state((9, 10, 11), ratio=(90, 6, 4), sigma=0.05, vacancies=vacancies)
    w = cm.oracle_world(st)
    ctx = cm.gpu_context(st)
    want = w.dump(0, 42)
    got = ctx.dump_records(42)
    assert got.size == 2 * 9 * 10 * 11 - vacancies
    assert got.tobytes() == want.tobytes()
    # a sub-region in ghost-inclusive doubled-x coordinates, ghosts (INVALID before the first exchange) skipped
    sub = ctx.dump_records(7, begin=(3, 2, 1), end=(20, 9, 8))
    arr = ctx.download().reshape(ctx.ext_shape)[1:8, 2:9, 3:20].reshape(-1)
    arr = arr[arr["type"] >= 0]
    assert np.array_equal(sub["id"], arr["id"]) and np.array_equal(sub["x"], arr["x"]) and np.array_equal(sub["v"], arr["v"])
    assert (sub["step"] == 7).all() and (sub["inter_type"] == 0).all()
    ctx.close()
    w.close()


def test_dump_after_steps_is_the_device_state():
    st = cm.make_state((10, 10, 10))
    ctx = cm.gpu_context(st)
    ctx.prepare()
    ctx.step(5)
    rec = ctx.dump_records(5)
    own = cm.owned(ctx, ctx.download()).reshape(-1)
    assert np.array_equal(rec["id"], own["id"]) and np.array_equal(rec["x"], own["x"]) and np.array_equal(rec["v"], own["v"])
    ctx.close()


def test_temperature_and_rescale_match_oracle():
    st = cm.make_state((9, 9, 10), ratio=(90, 6, 4), t_set=450.0, vacancies=3)
    w = cm.oracle_world(st)
    ctx = cm.gpu_context(st)
    th = ctx.temperature()
    assert abs(th["mvv"] - w.L.ora_mvv(w.h)) / th["mvv"] < 1e-13
    assert abs(th["T"] - w.temperature()) / th["T"] < 1e-13
    assert th["n_atoms"] == 2 * 9 * 9 * 10 - 3
    ctx.rescale_to(300.0)
    w.L.ora_rescale(w.h, 300.0)
    got, ref = cm.owned(ctx, ctx.download()), cm.owned(ctx, w.atoms(0))
    valid = ref["type"] >= 0
    assert cm.rel_err(got["v"][valid], ref["v"][valid]) < 1e-13
    assert abs(ctx.temperature()["T"] - 300.0) < 1e-9
    ctx.close()
    w.close()
