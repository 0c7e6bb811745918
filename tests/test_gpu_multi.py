"""Multi-GPU parity: one sub-box per GPU, ghost positions and df halos over NCCL send/recv, compared with the
oracle's multi-sub-box world on the same global state. Needs >= 2 GPUs (run under `gpurun --gpus 2`)."""
import os
import time

import numpy as np
import pytest

import misa_md_b200 as mb
from tests import common as cm

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import torch.multiprocessing as mp


def _worker(rank, world, phase, grid, ratio, steps, out_dir, pka, opts="", vacancies=0):
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if opts:
        os.environ["MISA_B200_OPTS"] = opts
    lib = mb.load()
    mb.capi._ck(lib.misa_b200_env_init(rank))
    coord = (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])
    st = cm.make_state(phase, ratio=ratio, sigma=0.03, vacancies=vacancies)
    ctx = cm.gpu_context(st, grid=grid, coord=coord, dt=pka["dt"] if pka else 0.001)
    uid_path = os.path.join(out_dir, "uid.bin")
    if rank == 0:
        uid = ctx.comm_unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(uid_path + ".tmp", uid_path)
    else:
        t0 = time.time()
        while not os.path.exists(uid_path):
            assert time.time() - t0 < 60
            time.sleep(0.01)
        uid = open(uid_path, "rb").read()
    ctx.comm_init(uid, rank, world)
    ctx.prepare()
    if pka:
        ctx.collision_step(pka["lat"], pka["dir"], pka["energy"])
    ctx.step(steps)
    np.save(os.path.join(out_dir, "lat%d.npy" % rank), ctx.download())
    np.save(os.path.join(out_dir, "inter%d.npy" % rank), ctx.download_inter())
    np.save(os.path.join(out_dir, "p2p%d.npy" % rank), np.array([ctx.query("p2p"), ctx.query("p2p_error")]))
    ctx.close()


def _run(tmp_path, phase, grid, ratio, steps, pka=None, opts="", vacancies=0):
    world = grid[0] * grid[1] * grid[2]
    if mb.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    mp.spawn(_worker, args=(world, phase, grid, ratio, steps, str(tmp_path), pka, opts, vacancies), nprocs=world, join=True)
    st = cm.make_state(phase, ratio=ratio, sigma=0.03, vacancies=vacancies)
    w = cm.oracle_world(st, grid=grid, dt=pka["dt"] if pka else 0.001, threads=world)
    w.prepare()
    if pka:
        import ctypes as C
        w.L.ora_collision_step(w.h, C.byref((C.c_int * 4)(*pka["lat"])), C.byref((C.c_double * 3)(*pka["dir"])), pka["energy"])
    for _ in range(steps):
        w.step()
    return w


def _compare(tmp_path, w, xtol, ftol):
    for r in range(w.n_ranks):
        got = np.load(os.path.join(str(tmp_path), "lat%d.npy" % r)).reshape(w.shape(r))[w.owned_slices(r)]
        ref = w.atoms(r).reshape(w.shape(r))[w.owned_slices(r)]
        assert np.array_equal(got["type"], ref["type"]), r  # ownership / occupancy exact on every sub-box
        valid = ref["type"] >= 0
        assert np.array_equal(got["id"][valid], ref["id"][valid])
        assert cm.rel_err(got["x"][valid], ref["x"][valid]) < xtol
        assert cm.rel_err(got["rho"][valid], ref["rho"][valid]) < 1e-9
        assert cm.rel_err(got["f"][valid], ref["f"][valid]) < ftol
        gi, ri = np.load(os.path.join(str(tmp_path), "inter%d.npy" % r)), w.inter(r)
        assert len(gi) == len(ri), r                          # migration: every inter atom on the right owner
        if len(ri):
            assert np.array_equal(gi["id"], ri["id"])
            assert cm.rel_err(gi["x"], ri["x"]) < xtol


def test_two_gpus_track_oracle(tmp_path):
    w = _run(tmp_path, (12, 8, 8), (2, 1, 1), (90, 6, 4), steps=5)
    _compare(tmp_path, w, 1e-12, 1e-9)
    w.close()


@pytest.mark.parametrize("opts", ["pipe=0", "pipe=1,overlap=1,reserve=8"])
def test_two_gpus_serial_and_overlapped_exchange(tmp_path, opts):
    """The serial step and the interior/boundary split with the NCCL exchange on a second stream (sub-boxes of
    8x8x8 cells: a 2x2x2 interior) both track the oracle."""
    w = _run(tmp_path, (16, 8, 8), (2, 1, 1), (90, 6, 4), steps=5, opts=opts)
    _compare(tmp_path, w, 1e-12, 1e-9)
    w.close()


def test_two_gpus_direct_push_equals_staged_nccl_exchange(tmp_path):
    """csrc/p2p.cuh (one push kernel into the neighbour's HBM + flag handshakes) against the three staged NCCL
    exchanges: the ghosts get the same bits, so the whole trajectory is identical in every field."""
    out = {}
    for name, opts in (("p2p", ""), ("nccl", "p2p=0")):
        d = tmp_path / name
        d.mkdir()
        w = _run(d, (16, 10, 8), (2, 1, 1), (90, 6, 4), steps=6, opts=opts)
        _compare(d, w, 1e-12, 1e-9)
        w.close()
        out[name] = [np.load(os.path.join(str(d), "lat%d.npy" % r)) for r in range(2)]
        flags = [np.load(os.path.join(str(d), "p2p%d.npy" % r)) for r in range(2)]
        assert all(f[1] == 0 for f in flags)
        if name == "p2p" and flags[0][0] != 1:
            pytest.skip("no peer access between the two GPUs: the NCCL path ran")
        assert all(f[0] == (1 if name == "p2p" else 0) for f in flags)
    for r in range(2):
        for fld in ("type", "x", "v", "f", "rho", "df"):
            assert np.array_equal(out["p2p"][r][fld], out["nccl"][r][fld]), (r, fld)


def test_two_gpus_wait_for_the_push_inside_the_stencil_kernels(tmp_path):
    """Sub-boxes long enough in x (40 cells) to have an interior: rho / force start on it while the neighbour's push is in
    flight and wait for the arrive flags right before their first boundary unit. Same bits as waiting in front of the
    launch, and the oracle's trajectory."""
    out = {}
    for name, opts in (("late", ""), ("front", "late=0")):
        d = tmp_path / name
        d.mkdir()
        w = _run(d, (80, 8, 8), (2, 1, 1), (1, 0, 0), steps=6, opts=opts)
        _compare(d, w, 1e-12, 1e-9)
        w.close()
        out[name] = [np.load(os.path.join(str(d), "lat%d.npy" % r)) for r in range(2)]
        flags = [np.load(os.path.join(str(d), "p2p%d.npy" % r)) for r in range(2)]
        assert all(f[1] == 0 for f in flags)
        if flags[0][0] != 1:
            pytest.skip("no peer access between the two GPUs: the NCCL path ran")
    for r in range(2):
        for fld in ("type", "x", "v", "f", "rho", "df"):
            assert np.array_equal(out["late"][r][fld], out["front"][r][fld]), (r, fld)


@pytest.mark.parametrize("phase,ratio,vacancies", [((80, 8, 8), (1, 0, 0), 0), ((24, 10, 8), (97, 2, 1), 0), ((16, 8, 10), (1, 0, 0), 4)])
def test_two_gpus_push_from_inside_the_producers_equals_the_push_kernels(tmp_path, phase, ratio, vacancies):
    """The sync-free step's ghost pushes done by k_verlet1 (positions) and k_rho_f's epilogue (df), ARRIVE posted by the
    consuming stencil kernel (or a tiny kernel when it cannot wait inside: vacancies), against the push kernels of
    csrc/p2p.cuh: the ghosts get the same bits, so every field of the trajectory is identical -- and both track the oracle.
    Pure Fe with an interior, the dilute alloy of BASELINE config 3, pure Fe with vacancies."""
    out = {}
    for name, opts in (("fused", ""), ("kernels", "push_fused=0")):
        d = tmp_path / name
        d.mkdir()
        w = _run(d, phase, (2, 1, 1), ratio, steps=6, opts=opts, vacancies=vacancies)
        _compare(d, w, 1e-12, 1e-9)
        w.close()
        out[name] = [np.load(os.path.join(str(d), "lat%d.npy" % r)) for r in range(2)]
        flags = [np.load(os.path.join(str(d), "p2p%d.npy" % r)) for r in range(2)]
        assert all(f[1] == 0 for f in flags)
        if flags[0][0] != 1:
            pytest.skip("no peer access between the two GPUs: the NCCL path ran")
    for r in range(2):
        for fld in ("type", "x", "v", "f", "rho", "df"):
            assert np.array_equal(out["fused"][r][fld], out["kernels"][r][fld]), (r, fld)


def test_two_gpus_pka_migrates_across_sub_boxes(tmp_path):
    """PKA launched next to the sub-box interface: inter atoms cross ranks (exchangeInter) and act as ghost
    inter atoms on the neighbour (borderInter + the inter part of the df halo)."""
    pka = dict(lat=(5, 4, 4, 0), dir=(3.0, 0.7, 0.4), energy=400.0, dt=2e-4)
    w = _run(tmp_path, (12, 8, 8), (2, 1, 1), (1, 0, 0), steps=150, pka=pka)
    assert w.total_inter() > 0
    _compare(tmp_path, w, 1e-9, 1e-7)
    w.close()


def test_eight_gpus_track_oracle(tmp_path):
    w = _run(tmp_path, (12, 12, 12), (2, 2, 2), (90, 6, 4), steps=3)
    _compare(tmp_path, w, 1e-12, 1e-9)
    w.close()


def _frontier_worker(rank, world, phase, grid, ratio, out_dir):
    """SURVEY.md section 8f entry points on a decomposed box: device world build by global id, global temperature and
    rescale (NCCL all-reduce), dump record stream per sub-box."""
    lib = mb.load()
    mb.capi._ck(lib.misa_b200_env_init(rank))
    coord = (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])
    ctx = mb.Context(phase, grid=grid, coord=coord, a=cm.A, crf=cm.CRF)
    ctx.make_offsets()
    ctx.set_potential(*cm.host_potential())
    uid_path = os.path.join(out_dir, "uid.bin")
    if rank == 0:
        with open(uid_path + ".tmp", "wb") as f:
            f.write(ctx.comm_unique_id())
        os.rename(uid_path + ".tmp", uid_path)
    else:
        t0 = time.time()
        while not os.path.exists(uid_path):
            assert time.time() - t0 < 60
            time.sleep(0.01)
    ctx.comm_init(open(uid_path, "rb").read(), rank, world)
    ctx.build_world(seed=466953, t_set=600.0, ratio=ratio, alloy_seed=9)
    t0 = ctx.temperature()
    ctx.prepare()
    ctx.step(4)
    t1 = ctx.temperature()
    ctx.rescale_to(350.0)
    t2 = ctx.temperature()
    np.save(os.path.join(out_dir, "lat%d.npy" % rank), ctx.download())
    np.save(os.path.join(out_dir, "dump%d.npy" % rank), ctx.dump_records(4))
    np.save(os.path.join(out_dir, "temps%d.npy" % rank), np.array([t0["T"], t1["T"], t2["T"], t0["n_atoms"]]))
    ctx.close()


def test_two_gpus_world_thermo_rescale_dump(tmp_path):
    from misa_md_b200 import synth
    phase, grid, ratio = (12, 8, 8), (2, 1, 1), (90, 6, 4)
    if mb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_frontier_worker, args=(2, phase, grid, ratio, str(tmp_path)), nprocs=2, join=True)
    st = synth.create_global_state(phase, a=cm.A, seed=466953, t_set=600.0, ratio=ratio, alloy_seed=9)
    w = cm.oracle_world(st, grid=grid, threads=2)
    w.prepare()
    for _ in range(4):
        w.step()
    t1 = w.temperature()
    w.L.ora_rescale(w.h, 350.0)
    temps = [np.load(os.path.join(str(tmp_path), "temps%d.npy" % r)) for r in range(2)]
    assert np.array_equal(temps[0], temps[1])                # every rank holds the same global numbers
    assert abs(temps[0][0] - 600.0) < 1e-9 and temps[0][3] == 2 * 12 * 8 * 8
    assert abs(temps[0][1] - t1) / t1 < 1e-10
    assert abs(temps[0][2] - 350.0) < 1e-9
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "lat%d.npy" % r)).reshape(w.shape(r))[w.owned_slices(r)]
        ref = w.atoms(r).reshape(w.shape(r))[w.owned_slices(r)]
        assert np.array_equal(got["id"], ref["id"]) and np.array_equal(got["type"], ref["type"])
        assert cm.rel_err(got["x"], ref["x"]) < 1e-12
        assert cm.rel_err(got["v"], ref["v"]) < 1e-9
        rec, want = np.load(os.path.join(str(tmp_path), "dump%d.npy" % r)), w.dump(r, 4)
        assert np.array_equal(rec["id"], want["id"]) and np.array_equal(rec["type"], want["type"])
        assert cm.rel_err(rec["x"], want["x"]) < 1e-12 and cm.rel_err(rec["v"], want["v"]) < 1e-9
    w.close()
