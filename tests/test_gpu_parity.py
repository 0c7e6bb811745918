"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.
Bars (BASELINE.json north_star): lattice indexing / occupancy / ownership bit-exact; per-atom rho, df, force
within 1e-10 relative."""
import numpy as np
import pytest

from tests import common as cm

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def alloy_case():
    st = cm.make_state((10, 11, 12), ratio=(90, 6, 4), sigma=0.05)
    w = cm.oracle_world(st)
    w.prepare()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    got = ctx.download()
    yield st, w, ctx, got
    ctx.close()
    w.close()


def test_halo_positions_bit_exact(alloy_case):
    st, w, ctx, got = alloy_case
    ref = w.atoms(0)
    assert np.array_equal(got["type"], ref["type"])
    assert np.array_equal(got["x"], ref["x"])  # ghosts carry the periodic image shift, bit-exact


def test_rho_df_force_parity(alloy_case):
    st, w, ctx, got = alloy_case
    ref = cm.owned(ctx, w.atoms(0))
    g = cm.owned(ctx, got)
    assert cm.rel_err(g["rho"], ref["rho"]) < TOL
    assert cm.rel_err(g["df"], ref["df"]) < TOL
    assert cm.rel_err(g["f"], ref["f"]) < TOL
    assert np.max(np.abs(ref["f"])) > 1e-2  # the test is not vacuous


def test_pruned_and_full_stencil_agree(alloy_case):
    st, w, ctx, got = alloy_case
    ctx2 = cm.gpu_context(st)
    ctx2.set_option("prune", 0)
    ctx2.set_option("fuse", 0)
    ctx2.prepare()
    full = cm.owned(ctx2, ctx2.download())
    g = cm.owned(ctx, got)
    ctx2.close()
    assert cm.rel_err(g["rho"], full["rho"]) < 1e-13
    assert cm.rel_err(g["f"], full["f"]) < 1e-12


@pytest.mark.parametrize("ratio", [(1, 0, 0), (0, 1, 0), (90, 6, 4)])
def test_smem_tables_match_generic_tables(ratio):
    """eam_smem.cuh (Hermite tables staged in shared memory by TMA) vs kernels.cuh (7-coefficient rows in global
    memory, the reference's evaluation order) -- both against the oracle, single- and multi-species."""
    st = cm.make_state((9, 10, 11), ratio=ratio, sigma=0.06)
    w = cm.oracle_world(st)
    w.prepare()
    ref = None
    out = {}
    for smem in (1, 0):
        ctx = cm.gpu_context(st)
        ctx.set_option("smem", smem)
        ctx.prepare()
        out[smem] = cm.owned(ctx, ctx.download())
        ref = cm.owned(ctx, w.atoms(0)).copy()  # w.atoms() is a view into the oracle's memory
        ctx.close()
    w.close()
    for smem in (1, 0):
        for fld in ("rho", "df", "f"):
            assert cm.rel_err(out[smem][fld], ref[fld]) < TOL, (smem, fld)
    assert cm.rel_err(out[1]["f"], out[0]["f"]) < 1e-11


@pytest.mark.parametrize("tex,novac", [(1, 0), (0, 1), (1, 1)])
def test_kernel_variants_match_oracle(tex, novac):
    """TEX-pipe neighbour loads and the no-vacancy fast path (type test dropped) against the oracle, then with a
    vacancy present (the census must switch the fast path off)."""
    for vac in (0, 5):
        st = cm.make_state((9, 9, 10), sigma=0.06, vacancies=vac)
        w = cm.oracle_world(st)
        w.prepare()
        ctx = cm.gpu_context(st)
        ctx.set_option("tex", tex)
        ctx.set_option("novac", novac)
        ctx.prepare()
        for _ in range(3):
            w.step()
        ctx.step(3)
        got = cm.owned(ctx, ctx.download())
        ref = cm.owned(ctx, w.atoms(0))
        valid = ref["type"] >= 0
        assert np.array_equal(got["type"], ref["type"])
        assert cm.rel_err(got["rho"][valid], ref["rho"][valid]) < TOL
        assert cm.rel_err(got["f"][valid], ref["f"][valid]) < 1e-9
        ctx.close()
        w.close()


@pytest.mark.parametrize("ratio", [(1, 0, 0), (90, 6, 4)])
def test_kernel_generations_agree(ratio):
    """Third-generation kernels (eam_fast.cuh: warp-uniform range test, Hermite basis form, near/far lists) against
    the second generation (eam_smem.cuh) and the oracle on the same state."""
    st = cm.make_state((9, 10, 11), ratio=ratio, sigma=0.07)
    w = cm.oracle_world(st)
    w.prepare()
    ref = None
    out = {}
    for fast in (0, 1, 2):   # 2: also the multi-species force of generation three
        ctx = cm.gpu_context(st)
        ctx.set_option("fast", fast)
        ctx.prepare()
        out[fast] = cm.owned(ctx, ctx.download())
        if ref is None:
            ref = cm.owned(ctx, w.atoms(0))
        ctx.close()
    for fast in (0, 1, 2):
        assert cm.rel_err(out[fast]["rho"], ref["rho"]) < TOL
        assert cm.rel_err(out[fast]["df"], ref["df"]) < TOL
        assert cm.rel_err(out[fast]["f"], ref["f"]) < 1e-9
    assert cm.rel_err(out[1]["f"], out[0]["f"]) < 1e-11
    assert cm.rel_err(out[2]["f"], out[0]["f"]) < 1e-11
    w.close()


@pytest.mark.parametrize("ratio", [(1, 0, 0), (90, 6, 4)])
def test_pipelined_step_is_the_serial_step(ratio):
    """The sync-free step (device-side choice of the pruned list, activity word checked before verlet2) and its
    interior / boundary split must reproduce the serial step BIT FOR BIT: same kernels, same lists, same order of
    the per-atom sums."""
    st = cm.make_state((11, 10, 12), ratio=ratio, sigma=0.02)
    out = {}
    # sym 0: the interior / boundary launches keep the full-list kernels, the whole-box launch would otherwise take the
    # pair-symmetric passes (another summation order; tests/test_gpu_sym.py compares those)
    for name, opts in (("serial", {"pipe": 0, "sym": 0}), ("pipe", {"pipe": 1, "sym": 0}), ("split", {"pipe": 1, "overlap": 2, "reserve": 8, "sym": 0})):
        ctx = cm.gpu_context(st)
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.prepare()
        ctx.step(6)
        out[name] = cm.owned(ctx, ctx.download())
        if name != "serial":
            assert ctx.query("pipe_steps") == 6 and ctx.query("pipe_redo") == 0
        ctx.close()
    for name in ("pipe", "split"):
        for fld in ("type", "x", "v", "f", "rho", "df"):
            assert np.array_equal(out[name][fld], out["serial"][fld]), (name, fld)


def test_pipelined_step_falls_back_on_runaway():
    """A PKA makes atoms leave their sites: the speculative rho/force of that step are discarded and the step is
    redone through the off-lattice path, so the trajectory equals the serial one exactly."""
    st = cm.make_state((10, 10, 10))
    out = {}
    for pipe in (0, 1):
        ctx = cm.gpu_context(st, dt=2e-4)
        ctx.set_option("pipe", pipe)
        ctx.prepare()
        ctx.collision_step((5, 5, 5, 0), (3.0, 0.7, 0.4), 400.0)
        ctx.step(60)
        out[pipe] = (cm.owned(ctx, ctx.download()), ctx.download_inter())
        if pipe:
            assert ctx.query("pipe_redo") >= 1
        ctx.close()
    assert len(out[0][1]) > 0
    assert np.array_equal(out[0][1]["id"], out[1][1]["id"])
    for fld in ("type", "x", "v", "f"):
        assert np.array_equal(out[0][0][fld], out[1][0][fld]), fld


def test_close_pairs_below_staged_range():
    """Pairs closer than the staged r range (r < 2 Angstrom) take the global Hermite rows: same results."""
    st = cm.make_state((8, 8, 8), sigma=0.0)
    x = st["x"]
    x[4, 4, 8] += (x[4, 4, 9] - x[4, 4, 8]) * 0.22  # push one atom towards its 1nn: r ~ 1.93 A, still < 0.2a... from its site
    w = cm.oracle_world(st)
    w.prepare()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    got = cm.owned(ctx, ctx.download())
    ref = cm.owned(ctx, w.atoms(0))
    d = np.linalg.norm(st["x"][4, 4, 9] - st["x"][4, 4, 8])
    assert d < 2.0
    for fld in ("rho", "df", "f"):
        assert cm.rel_err(got[fld], ref[fld]) < TOL, fld
    ctx.close()
    w.close()


def test_vacancies(alloy_case):
    st = cm.make_state((8, 8, 8), ratio=(90, 6, 4), sigma=0.05, vacancies=40)
    w = cm.oracle_world(st)
    w.prepare()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    got = ctx.download()
    ref = cm.owned(ctx, w.atoms(0))
    g = cm.owned(ctx, got)
    assert np.array_equal(g["type"], ref["type"])
    valid = ref["type"] >= 0
    assert cm.rel_err(g["rho"][valid], ref["rho"][valid]) < TOL
    assert cm.rel_err(g["f"][valid], ref["f"][valid]) < TOL
    assert np.all(g["rho"][~valid] == 0.0) and np.all(g["f"][~valid] == 0.0)
    ctx.close()
    w.close()


def test_verlet_bit_exact():
    """firststep / secondstep with identical forces must reproduce x and v to the last bit."""
    st = cm.make_state((8, 9, 10), ratio=(90, 6, 4), sigma=0.05)
    w = cm.oracle_world(st)
    w.prepare()
    ctx = cm.gpu_context(st, upload=False)
    ctx.upload(w.atoms(0).copy())  # same f, v, x as the oracle
    w.L.ora_first_step(w.h)
    ctx.run_pass("verlet1")
    got = cm.owned(ctx, ctx.download())
    ref = cm.owned(ctx, w.atoms(0))
    assert np.array_equal(got["x"], ref["x"])
    assert np.array_equal(got["v"], ref["v"])
    w.L.ora_second_step(w.h)
    ctx.run_pass("verlet2")
    got = cm.owned(ctx, ctx.download())
    assert np.array_equal(got["v"], cm.owned(ctx, w.atoms(0))["v"])
    ctx.close()
    w.close()


def test_ten_steps_track_oracle():
    st = cm.make_state((10, 10, 10))
    w = cm.oracle_world(st)
    w.prepare()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    for _ in range(10):
        w.step()
    ctx.step(10)
    got = cm.owned(ctx, ctx.download())
    ref = cm.owned(ctx, w.atoms(0))
    assert np.array_equal(got["type"], ref["type"])
    assert cm.rel_err(got["x"], ref["x"]) < 1e-12
    assert cm.rel_err(got["v"], ref["v"]) < 1e-9
    assert cm.rel_err(got["f"], ref["f"]) < 1e-8
    th = ctx.thermo()
    assert abs(th["mvv"] - w.L.ora_mvv(w.h)) / th["mvv"] < 1e-10
    assert abs(th["pe"] - w.potential_energy()) / abs(th["pe"]) < 1e-10
    ctx.close()
    w.close()


def test_compat_hooks_match_cpu_branch():
    """The three reference hooks on a HOST AoS array (ghosts already exchanged by the host) reproduce the
    CPU branch of computeEam on owned sites."""
    st = cm.make_state((8, 8, 9), ratio=(90, 6, 4), sigma=0.05)
    w = cm.oracle_world(st)
    w.L.ora_exchange_atom_first(w.h)
    w.L.ora_clear_force(w.h)
    host = w.atoms(0).copy()  # host array as the driver would hand it to the hooks
    w.L.ora_compute_eam(w.h)
    ref = w.atoms(0)
    ctx = cm.gpu_context(st, upload=False)
    ctx.eam_rho_calc(host)
    own = ctx.owned
    h3 = host.reshape(ctx.ext_shape)
    r3 = ref.reshape(ctx.ext_shape)
    assert cm.rel_err(h3[own]["rho"], r3[own]["rho"]) < TOL
    ctx.eam_df_calc(host)
    assert cm.rel_err(h3[own]["df"], r3[own]["df"]) < TOL
    # the host's forward df halo (DfEmbedPacker) fills the ghosts before the force hook
    h3["df"][...] = r3["df"]
    ctx.eam_force_calc(host)
    assert cm.rel_err(h3[own]["f"], r3[own]["f"]) < TOL
    assert np.array_equal(host["x"], ref["x"]) and np.array_equal(host["id"], ref["id"])
    ctx.close()
    w.close()


@pytest.mark.parametrize("ratio,vac", [((97, 2, 1), 0), ((92, 5, 3), 7), ((3, 95, 2), 0)])
def test_dilute_alloy_kernels_match_oracle(ratio, vac):
    """Dilute-alloy kernels of eam_fast.cuh (majority tables branch-free, minority pairs listed and corrected, minority atoms one warp
    each) against the oracle and against the general multi-species kernels, over a few steps."""
    st = cm.make_state((10, 9, 11), ratio=ratio, sigma=0.06, vacancies=vac)
    w = cm.oracle_world(st)
    w.prepare()
    out = {}
    for dilute in (1, 0):
        ctx = cm.gpu_context(st)
        ctx.set_option("dilute", dilute)
        ctx.prepare()
        assert ctx.query("dilute") == dilute
        first = cm.owned(ctx, ctx.download()).copy()
        ref = cm.owned(ctx, w.atoms(0))
        valid = ref["type"] >= 0
        for fld in ("rho", "df", "f"):
            assert cm.rel_err(first[fld][valid], ref[fld][valid]) < TOL, (dilute, fld)
        ctx.step(3)
        out[dilute] = cm.owned(ctx, ctx.download()).copy()
        ctx.close()
    for _ in range(3):
        w.step()
    ref = cm.owned(ctx, w.atoms(0))
    valid = ref["type"] >= 0
    for dilute in (1, 0):
        assert np.array_equal(out[dilute]["type"], ref["type"])
        assert cm.rel_err(out[dilute]["rho"][valid], ref["rho"][valid]) < TOL
        assert cm.rel_err(out[dilute]["f"][valid], ref["f"][valid]) < 1e-9
    assert cm.rel_err(out[1]["x"][valid], out[0]["x"][valid]) < 1e-13
    w.close()


def test_dilute_alloy_solute_cluster_overflows_the_pair_list():
    """A compact Cu precipitate: atoms next to it have more than EAM_LIST_CAP minority neighbours, so the per-lane
    list overflows and the atom is recomputed by the generic routine; Cu atoms inside it go through k_force_minor."""
    st = cm.make_state((12, 12, 12), ratio=(1, 0, 0), sigma=0.05)
    st["type"][4:8, 4:8, 8:16] = 1      # 4x4x4 cells of Cu = 128 atoms of 3456 (3.7 %)
    w = cm.oracle_world(st)
    w.prepare()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    assert ctx.query("dilute") == 1
    got, ref = cm.owned(ctx, ctx.download()), cm.owned(ctx, w.atoms(0))
    for fld in ("rho", "df", "f"):
        assert cm.rel_err(got[fld], ref[fld]) < TOL, fld
    ctx.close()
    w.close()


def test_fused_half_kicks_are_bit_identical():
    """Inside one multi-step call the second half-kick of step k runs inside k_verlet1 of step k+1 (one pass over v
    and f instead of two). Same two rounded adds: identical bits to stepping one call at a time, and to the option off."""
    st = cm.make_state((10, 10, 10), ratio=(97, 2, 1), sigma=0.04)
    out = []
    for mode in ("one_call", "single_steps", "option_off"):
        ctx = cm.gpu_context(st)
        if mode == "option_off":
            ctx.set_option("fuse_verlet", 0)
        ctx.prepare()
        if mode == "single_steps":
            for _ in range(6):
                ctx.step(1)
        else:
            ctx.step(6)
        assert ctx.query("pipe_steps") == 6
        out.append(cm.owned(ctx, ctx.download()).copy())
        ctx.close()
    for other in out[1:]:
        for fld in ("x", "v", "f", "rho", "df"):
            assert np.array_equal(out[0][fld], other[fld]), fld


def test_step_host_moves_only_the_owned_box_and_equals_resident_steps():
    """misa_b200_step_host after its first call uploads / downloads the owned records only (one pitched 3-D copy each
    way): same trajectory, bit for bit, as resident stepping; ghost records of the host array are left alone."""
    st = cm.make_state((9, 8, 10), sigma=0.04)
    ctx = cm.gpu_context(st)
    ctx.prepare()
    ctx.step(4)
    want = cm.owned(ctx, ctx.download()).copy()
    ctx.close()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    host = ctx.download()
    ctx.host_register(host)
    ctx.step_host(host, 1)                      # resident state exists: already the owned-box path
    ghost = np.ones(ctx.ext_shape, dtype=bool)
    ghost[ctx.owned] = False
    h3 = host.reshape(ctx.ext_shape)
    h3["x"][ghost] = 12345.0                    # poison: must neither be read nor overwritten
    h3["rho"][ghost] = -1.0
    for _ in range(3):
        ctx.step_host(host, 1)
    got = cm.owned(ctx, host)
    for fld in ("type", "id", "x", "v", "f", "rho", "df"):
        assert np.array_equal(got[fld], want[fld]), fld
    assert np.all(h3["x"][ghost] == 12345.0) and np.all(h3["rho"][ghost] == -1.0)
    ctx.host_unregister(host)
    ctx.close()


@pytest.mark.parametrize("planes", [4, 3])
def test_step_host_slab_pipeline_equals_resident_steps_and_falls_back_on_a_runaway(planes):
    """The slab-pipelined misa_b200_step_host (z-slabs uploaded, computed and downloaded concurrently, both PCIe directions at
    once) gives the trajectory of resident stepping bit for bit, owned records only; when an atom runs away inside such a
    step the input records are restored on the device and the step is redone through the serial path -- same state as
    resident stepping through the cascade (occupancy, ids, inter-atom list)."""
    st = cm.make_state((10, 9, 26), sigma=0.04)
    ref = cm.gpu_context(st)
    ref.prepare()
    ctx = cm.gpu_context(st)
    ctx.set_option("slab_planes", planes)
    ctx.prepare()
    host = ctx.download()
    ctx.host_register(host)
    ctx.step_host(host, 1)
    ref.step(1)
    for _ in range(4):
        ctx.step_host(host, 1)
        ref.step(1)
    assert ctx.query("host_slab_steps") == 5 and ctx.query("host_slab_redo") == 0
    want = cm.owned(ref, ref.download())
    got = cm.owned(ctx, host)
    for fld in ("type", "id", "x", "v", "f", "rho", "df"):
        assert np.array_equal(got[fld], want[fld]), fld
    # kick one atom hard enough to leave its site within a few steps: both contexts from the same (bit-identical) state
    lat, direction = (5, 4, 13, 0), (1.0, 2.0, 3.0)
    ctx.setv(lat, direction, 300.0)
    ref.setv(lat, direction, 300.0)
    host = ctx.download(host)
    for _ in range(12):
        ctx.step_host(host, 1)
        ref.step(1)
    assert ref.thermo()["n_inter"] > 0 and ctx.query("host_slab_redo") >= 1
    want = cm.owned(ref, ref.download())
    got = cm.owned(ctx, host)
    for fld in ("type", "id", "x", "v", "f", "rho", "df"):
        assert np.array_equal(got[fld], want[fld]), fld
    a, b = ctx.download_inter(), ref.download_inter()
    assert len(a) == len(b) and np.array_equal(a["id"], b["id"]) and np.array_equal(a["x"], b["x"])
    ctx.host_unregister(host)
    ctx.close()
    ref.close()


def test_hot_cell_marks_bound_the_partner_rigorously():
    """Per-warp stencil prefixes bound the partner atom by the marking level T unless an atom above T sits within reach
    (k_verlet1 marks those cells). One atom 0.19a off its site, a partner in its <311>/2 shell (2.18a) 0.04a towards it:
    the pair is in range (1.949a < crf) although partner level + T says it cannot be -- only the mark keeps it. Also:
    marks on / off give identical bits on a thermal state."""
    st = cm.make_state((14, 14, 14), t_set=1e-6)
    x = st["x"]
    a_site, c_site = (7, 7, 14), (7, 8, 17)             # corner of cell (7,7,7); centre of cell (8,8,7): separation (1.5,1.5,0.5)a
    u = x[c_site] - x[a_site]
    assert abs(np.linalg.norm(u) / cm.A - 2.179) < 1e-3
    u /= np.linalg.norm(u)
    x[a_site] += 0.19 * cm.A * u
    x[c_site] -= 0.04 * cm.A * u
    w = cm.oracle_world(st)
    w.prepare()
    w.step()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    ctx.step(1)
    assert 0 <= ctx.query("mark_level") < 19 and 1 <= ctx.query("mark_count") <= 4
    got, ref = cm.owned(ctx, ctx.download()), cm.owned(ctx, w.atoms(0))
    for fld in ("rho", "df", "f"):
        assert cm.rel_err(got[fld], ref[fld]) < TOL, fld
    # the pair matters at the tolerance: without it rho of the partner is off by far more than 1e-10
    r = np.linalg.norm(st["x"][c_site] - st["x"][a_site])
    assert r < cm.A * cm.CRF
    ctx.close()
    w.close()
    st = cm.make_state((12, 13, 14), sigma=0.05)
    out = []
    for mark in (1, 0):
        ctx = cm.gpu_context(st)
        ctx.set_option("mark", mark)
        ctx.prepare()
        ctx.step(5)
        assert (ctx.query("mark_level") >= 0) == bool(mark)
        out.append(cm.owned(ctx, ctx.download()).copy())
        ctx.close()
    for fld in ("x", "v", "f", "rho", "df"):
        assert np.array_equal(out[0][fld], out[1][fld]), fld


def test_minority_atoms_force_matches_oracle():
    """k_force_minor (one warp per minority-species atom, pairs from the global monomial block, address-ordered offset list)
    against the oracle after two steps, on the minority atoms alone and on the whole box."""
    st = cm.make_state((10, 11, 9), ratio=(92, 5, 3), sigma=0.06, vacancies=5)
    w = cm.oracle_world(st)
    w.prepare()
    for _ in range(2):
        w.step()
    ctx = cm.gpu_context(st)
    ctx.prepare()
    assert ctx.query("dilute") == 1 and ctx.query("n_minor") > 0
    ctx.step(2)
    out = cm.owned(ctx, ctx.download()).copy()
    ref = cm.owned(ctx, w.atoms(0))
    ctx.close()
    valid = ref["type"] >= 0
    minor = valid & (ref["type"] != 0)
    assert minor.sum() > 100
    assert cm.rel_err(out["f"][valid], ref["f"][valid]) < 1e-9
    assert cm.rel_err(out["f"][minor], ref["f"][minor]) < 1e-9
    w.close()
