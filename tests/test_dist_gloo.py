"""N>1 host logic on CPU: two processes over torch.distributed (gloo, world_size 2) run the staged x->y->z ghost
exchange exactly as the CUDA path schedules it (same send/recv index lists and periodic shifts from the C ABI's
host-only planner, same peer ranks, LOWER and HIGHER going to the SAME peer when the grid is 2 wide) and must
reproduce the oracle's two-sub-box world bit for bit. No GPU involved: pack/unpack are numpy gathers here."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from misa_md_b200 import capi, synth
from oracle import oracle_py as O

A, CRF = 2.85532, 1.96125


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(rank, world, port, phase, grid, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coord = (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])
    dom = capi.make_domain(phase, grid, coord, A, CRF)
    assert dom.rank == rank
    st = synth.create_global_state(phase, a=A, ratio=(90, 6, 4))
    synth.perturb_positions(st, 0.05)
    arr, _ = synth.scatter_to_sub_box(st, grid, coord, CRF)
    for dim in range(3):
        plans = [capi.plan_halo(dom, dim, d) for d in range(2)]
        bufs = []
        for send, _, shift in plans:  # pack both directions first (libcomm neiSendReceive order)
            b = np.empty((len(send), 4))
            b[:, :3] = arr["x"][send] + shift
            b[:, 3] = arr["type"][send]
            bufs.append(torch.from_numpy(b))
        recvd = []
        for d in range(2):
            dst = dom.rank_id_neighbours[dim][d]
            src = dom.rank_id_neighbours[dim][(d + 1) % 2]
            got = torch.empty_like(bufs[d])
            if dst == rank:
                got.copy_(bufs[d])
            else:
                req = dist.isend(bufs[d], dst, tag=d)
                dist.recv(got, src, tag=d)
                req.wait()
            recvd.append(got.numpy())
        for d in range(2):
            recv = plans[d][1]
            arr["x"][recv] = recvd[d][:, :3]
            arr["type"][recv] = recvd[d][:, 3].astype(np.int32)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), arr)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("phase,grid", [((12, 8, 8), (2, 1, 1)), ((8, 8, 12), (1, 1, 2))])
def test_two_rank_ghost_exchange_matches_oracle(tmp_path, phase, grid, pot):
    mp.spawn(_exchange, args=(2, _free_port(), phase, grid, str(tmp_path)), nprocs=2, join=True)
    st = synth.create_global_state(phase, a=A, ratio=(90, 6, 4))
    synth.perturb_positions(st, 0.05)
    w = O.World(phase, grid=grid, a=A, crf=CRF, pot=pot)
    for r in range(2):
        arr, _ = synth.scatter_to_sub_box(st, grid, tuple(w.rank(r).dom.grid_coord), CRF)
        w.atoms(r)[:] = arr
    w.L.ora_exchange_atom_first(w.h)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        ref = w.atoms(r)
        assert np.array_equal(got["type"], ref["type"])
        assert np.array_equal(got["x"], ref["x"])
        assert np.all(got["type"] >= 0)  # every ghost site was filled
    w.close()


def _push(rank, world, port, phase, grid, out_dir):
    """The direct push of csrc/p2p.cuh as two processes would run it: every rank scatters each direction group of
    misa_b200_plan_push to the rank at the opposite grid offset (looked up from the grid coordinates, as p2p_setup does
    from the all-gathered blobs) -- ONE message per group instead of three dependent stages."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coord_of = [(r // (grid[1] * grid[2]), (r // grid[2]) % grid[1], r % grid[2]) for r in range(world)]
    rank_at = {c: r for r, c in enumerate(coord_of)}
    coord = coord_of[rank]
    dom = capi.make_domain(phase, grid, coord, A, CRF)
    st = synth.create_global_state(phase, a=A, ratio=(90, 6, 4))
    synth.perturb_positions(st, 0.05)
    arr, _ = synth.scatter_to_sub_box(st, grid, coord, CRF)
    own = arr.copy()                                  # pushes read OWNED sites only: the source never changes under them
    dst, src, code, shift = capi.plan_push(dom)
    sends, recvs = [], []
    for k in sorted(set(code.tolist())):
        s = (k % 3 - 1, (k // 3) % 3 - 1, k // 9 - 1)
        m = code == k
        to = rank_at[tuple((coord[d] - s[d]) % grid[d] for d in range(3))]
        frm = rank_at[tuple((coord[d] + s[d]) % grid[d] for d in range(3))]
        b = np.empty((int(m.sum()), 4))
        b[:, :3] = own["x"][src[m]] + shift[k]
        b[:, 3] = own["type"][src[m]]
        if to == rank:                                # a grid dimension of size 1: the group stays in this sub-box
            assert frm == rank
            arr["x"][dst[m]] = b[:, :3]
            arr["type"][dst[m]] = b[:, 3].astype(np.int32)
            continue
        got = torch.empty(b.shape, dtype=torch.float64)
        sends.append(dist.isend(torch.from_numpy(b), to, tag=int(k)))
        recvs.append((dist.irecv(got, frm, tag=int(k)), got, dst[m]))
    for req, got, where in recvs:
        req.wait()
        arr["x"][where] = got.numpy()[:, :3]
        arr["type"][where] = got.numpy()[:, 3].astype(np.int32)
    for req in sends:
        req.wait()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), arr)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("phase,grid", [((12, 8, 8), (2, 1, 1)), ((8, 8, 12), (1, 1, 2))])
def test_two_rank_direct_push_matches_oracle(tmp_path, phase, grid, pot):
    mp.spawn(_push, args=(2, _free_port(), phase, grid, str(tmp_path)), nprocs=2, join=True)
    st = synth.create_global_state(phase, a=A, ratio=(90, 6, 4))
    synth.perturb_positions(st, 0.05)
    w = O.World(phase, grid=grid, a=A, crf=CRF, pot=pot)
    for r in range(2):
        arr, _ = synth.scatter_to_sub_box(st, grid, tuple(w.rank(r).dom.grid_coord), CRF)
        w.atoms(r)[:] = arr
    w.L.ora_exchange_atom_first(w.h)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        ref = w.atoms(r)
        assert np.array_equal(got["type"], ref["type"])
        assert np.array_equal(got["x"], ref["x"])
    w.close()
