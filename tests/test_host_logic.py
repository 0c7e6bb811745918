"""Host-side logic of the product checked on CPU against the oracle (which is pinned to the reference's own unit
tests in test_oracle_kat.py): lattice indexing, halo ownership lists, setfl parsing + spline rows, initial state.
Integer results must be bit-exact."""
import ctypes as C

import numpy as np
import pytest

import misa_md_b200 as mb
from misa_md_b200 import capi, synth
from oracle import oracle_py as O

A, CRF = 2.85532, 1.96125
CASES = [((8, 9, 10), (1, 1, 1), (0, 0, 0)), ((12, 8, 10), (2, 1, 1), (1, 0, 0)), ((12, 12, 12), (2, 2, 2), (1, 0, 1)),
         ((16, 12, 20), (2, 2, 2), (0, 1, 1)), ((200, 200, 200), (2, 2, 2), (1, 1, 1))]


@pytest.mark.parametrize("phase,grid,coord", CASES)
def test_neighbour_offsets_bit_exact(phase, grid, coord):
    dom = capi.make_domain(phase, grid, coord, A, CRF)
    w = None
    odom = O.make_domain(phase, grid, coord, A, CRF, ghost=3)
    rk = O.Rank()
    e = odom.dbx_ghost_extended_lattice_size
    rk.size_x, rk.size_y, rk.size_z = e[0], e[1], e[2]
    O.lib().ora_nei_make(C.byref(rk), odom.cut_lattice, CRF)
    want = [v.to_numpy() for v in (rk.nei_even, rk.nei_odd, rk.nei_half_even, rk.nei_half_odd)]
    for which in range(4):
        got = capi.plan_offsets(dom, which)
        assert np.array_equal(got, want[which])  # same values in the same order
    assert len(want[0]) == 228 and len(want[2]) == 114


@pytest.mark.parametrize("phase,grid,coord", CASES[:4])
def test_halo_lists_bit_exact(phase, grid, coord, pot):
    """sendlist / recvlist of AtomList::exchangeAtomFirst for this sub-box, and the periodic shifts."""
    w = O.World(phase, grid=grid, a=A, crf=CRF, pot=pot)
    w.fill_lattice()
    w.L.ora_exchange_atom_first(w.h)
    dom = capi.make_domain(phase, grid, coord, A, CRF)
    rk = w.rank(dom.rank)
    assert tuple(rk.dom.grid_coord) == tuple(coord)
    for dim in range(3):
        for direction in range(2):
            send, recv, shift = capi.plan_halo(dom, dim, direction)
            assert np.array_equal(send, rk.sendlist[2 * dim + direction].to_numpy())
            assert np.array_equal(recv, rk.recvlist[2 * dim + direction].to_numpy())
            want = np.zeros(3)
            if coord[dim] == 0 and direction == 0:
                want[dim] = phase[dim] * A
            if coord[dim] == grid[dim] - 1 and direction == 1:
                want[dim] = -(phase[dim] * A)
            assert np.array_equal(shift, want)
    w.close()


def test_domain_matches_oracle_domain():
    for phase, grid, coord in CASES:
        d = capi.make_domain(phase, grid, coord, A, CRF)
        o = O.make_domain(phase, grid, coord, A, CRF, ghost=3)
        assert d.rank == o.rank
        assert [list(r) for r in d.rank_id_neighbours] == [list(r) for r in o.rank_id_neighbours]
        assert list(d.sub_box_lattice_size) == list(o.sub_box_lattice_size)
        assert list(d.lattice_size_ghost) == list(o.lattice_size_ghost)
        r = o.sub_box_lattice_region
        assert list(d.sub_box_lattice_low) == [r.x_low, r.y_low, r.z_low]
        assert list(d.meas_global_length) == list(o.meas_global_length)


def test_setfl_reader_and_spline_rows_bit_exact(pot):
    """The product's host-side setfl parser + array2spline give the oracle's coefficient rows to the last bit
    (both restate libpot's interpolateFile; the GPU consumes exactly these rows)."""
    p = capi.read_setfl(mb.SETFL_PATH)
    assert p["keys"] == [26, 29, 28]
    L = O.lib()

    class PotTable(C.Structure):
        _fields_ = [("n", C.c_int), ("dx", C.c_double), ("inv_dx", C.c_double), ("values", C.POINTER(C.c_double)),
                    ("spline", C.POINTER(C.c_double))]

    class PotEam(C.Structure):
        _fields_ = [("n_ele", C.c_int), ("key", C.c_int * 3), ("mass", C.c_double * 3), ("lat_const", C.c_double * 3),
                    ("n_rho", C.c_int), ("n_r", C.c_int), ("d_rho", C.c_double), ("d_r", C.c_double), ("cutoff", C.c_double),
                    ("embed", PotTable * 3), ("elec", PotTable * 3), ("phi", (PotTable * 3) * 3)]

    pe = C.cast(pot.h, C.POINTER(PotEam)).contents
    assert pe.n_ele == 3

    def rows(t):
        return np.ctypeslib.as_array(t.spline, shape=((t.n + 1) * 7,))

    for i in range(3):
        for mine, theirs in ((p["elec"][i], pe.elec[i]), (p["embed"][i], pe.embed[i])):
            assert mine[0] == theirs.n and mine[1] == theirs.inv_dx
            assert np.array_equal(mine[2], rows(theirs))
        for j in range(3):
            assert np.array_equal(p["phi"][i][j][2], rows(pe.phi[i][j]))


def test_mt19937_stream_is_std_mt19937():
    # std::mt19937 default-seeded (5489) yields 3499211612 first and 4123659995 as its 10000th output
    u = synth.mt19937_unit(5489, 10000)
    assert int(round(u[0] * 4294967295.0)) == 3499211612
    assert int(round(u[9999] * 4294967295.0)) == 4123659995


def test_initial_state_rules():
    st = synth.create_global_state((6, 7, 8), seed=466953, t_set=600.0, ratio=(97, 2, 1))
    n = 2 * 6 * 7 * 8
    m = synth.MASS[st["type"]]
    p = (st["v"] * m[..., None]).reshape(-1, 3).sum(axis=0)
    assert np.all(np.abs(p) < 1e-9)
    t = float(((st["v"] ** 2).sum(axis=-1) * m).sum()) * synth.MVV2E / ((3 * n - 3) * synth.BOLTZ)
    assert abs(t - 600.0) < 1e-9
    # world_builder.cpp:120-124: body-centre sites sit at +a/2 in y and z
    assert st["x"][0, 0, 1].tolist() == [0.5 * A, A / 2, A / 2]
    assert set(np.unique(st["type"]).tolist()) <= {0, 1, 2}
    arr, lay = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
    a3 = arr.reshape(lay["ext_shape"])
    assert np.all(a3["type"][:3] == synth.INVALID) and np.all(a3[lay["owned"]]["type"] >= 0)
    assert np.array_equal(a3[lay["owned"]]["id"].reshape(-1), np.arange(1, n + 1, dtype=np.uint64))


def test_sub_box_state_matches_reference_builder_rule():
    """create_sub_box_state draws per sub-box with one seed (the reference's per-rank rule) and places the
    block at its global lattice position / ids."""
    arr, lay = synth.create_sub_box_state((8, 8, 8), (2, 1, 1), (1, 0, 0))
    own = arr.reshape(lay["ext_shape"])[lay["owned"]]
    assert own.shape == (8, 8, 8)
    assert own["x"][0, 0, 0].tolist() == [4 * A, 0.0, 0.0]
    assert own["id"][0, 0, 0] == 1 + 8
    loc = synth.create_global_state((4, 8, 8))
    assert np.array_equal(own["v"], loc["v"])


@pytest.mark.parametrize("phase,grid", [((12, 12, 12), (2, 2, 2)), ((12, 8, 10), (2, 1, 1)), ((8, 16, 18), (1, 2, 3)), ((8, 9, 10), (1, 1, 1)),
                                        ((12, 3, 8), (4, 1, 2)), ((9, 12, 5), (3, 4, 1)), ((6, 6, 6), (2, 2, 2))])   # incl. sub-boxes as narrow as the ghost shell
def test_direct_push_map_equals_the_staged_exchange(phase, grid, pot):
    """misa_b200_plan_push (what csrc/p2p.cuh applies with one kernel over NVLink peer memory) against the reference's
    three staged exchanges run by the oracle on every sub-box: same ghost positions (bits) and types everywhere."""
    w = O.World(phase, grid=grid, a=A, crf=CRF, pot=pot)
    w.fill_lattice()
    rs = np.random.RandomState(5)
    for r in range(w.n_ranks):                        # distinguishable owned values, ghosts poisoned
        a = w.atoms(r).reshape(w.shape(r))
        own = w.owned_slices(r)
        keep_x, keep_t = a["x"][own].copy(), a["type"][own].copy()
        a["x"][...] = np.nan
        a["type"][...] = -7
        a["x"][own] = keep_x + rs.uniform(-0.2, 0.2, keep_x.shape)
        a["type"][own] = rs.randint(0, 3, keep_t.shape)
    before = [w.atoms(r).copy() for r in range(w.n_ranks)]
    w.L.ora_exchange_atom_first(w.h)
    coord_of = [tuple(w.rank(r).dom.grid_coord) for r in range(w.n_ranks)]
    rank_at = {c: r for r, c in enumerate(coord_of)}
    got = [b.copy() for b in before]
    n_ghost = 0
    for r in range(w.n_ranks):
        dom = capi.make_domain(phase, grid, coord_of[r], A, CRF)
        dst, src, code, shift = capi.plan_push(dom)
        n_ghost = len(dst)
        assert len(np.unique(dst)) == len(dst)
        for k in np.unique(code):
            s = (k % 3 - 1, (k // 3) % 3 - 1, k // 9 - 1)
            to = rank_at[tuple((coord_of[r][d] - s[d]) % grid[d] for d in range(3))]
            m = code == k
            got[to]["x"][dst[m]] = before[r]["x"][src[m]] + shift[k]
            got[to]["type"][dst[m]] = before[r]["type"][src[m]]
    e = w.shape(0)
    assert n_ghost == e[0] * e[1] * e[2] - 2 * (phase[0] // grid[0]) * (phase[1] // grid[1]) * (phase[2] // grid[2])
    for r in range(w.n_ranks):
        ref = w.atoms(r)
        assert np.array_equal(got[r]["type"], ref["type"])
        assert np.array_equal(got[r]["x"].view(np.uint64), ref["x"].view(np.uint64))   # incl. the image shifts, bit for bit
    w.close()


@pytest.mark.parametrize("phase,grid,coord", CASES[:4])
def test_stencil_plan_is_a_permutation_with_exact_prefixes(phase, grid, coord):
    """misa_b200_plan_stencil: what the stencil kernels loop. (1) a permutation of NeighbourIndex::make's full list (so the
    un-pruned sums are the reference's); (2) every pruned list is a prefix that contains EVERY offset whose sites are closer
    than (crf + 0.01 L) a -- the exactness of the per-warp pruning rests on this; (3) the near group is split into one of
    each +v / -v pair and its mirrors, and lower_slot pairs them up (pair-symmetric passes)."""
    dom = capi.make_domain(phase, grid, coord, A, CRF)
    sx = 2 * (dom.sub_box_lattice_size[0] + 2 * dom.lattice_size_ghost[0])
    sy = dom.sub_box_lattice_size[1] + 2 * dom.lattice_size_ghost[1]

    def sep2(off, parity):   # squared site separation in units of a^2, from the doubled-x linear offset
        dx = ((off % sx) + sx + sx // 2) % sx - sx // 2
        r = (off - dx) // sx
        dy = ((r % sy) + sy + sy // 2) % sy - sy // 2
        dz = (r - dy) // sy
        h = (0.5 if parity == 0 else -0.5) if dx & 1 else 0.0
        return (0.5 * dx) ** 2 + (dy + h) ** 2 + (dz + h) ** 2, (dx, 2 * dy + int(2 * h), 2 * dz + int(2 * h))

    plans = [capi.plan_stencil(dom, p) for p in range(2)]
    for p, pl in enumerate(plans):
        ref = capi.plan_offsets(dom, p)
        assert len(pl["sorted"]) == len(ref) == 228
        assert np.array_equal(np.sort(pl["sorted"]), np.sort(ref))
        r2 = np.array([sep2(int(o), p)[0] for o in pl["sorted"]])
        assert np.allclose(r2, pl["site_r2"], rtol=0, atol=1e-12)
        near, half = pl["n_near"], pl["n_half"]
        assert near == 2 * half and half > 0
        assert np.all(r2[:near] < (CRF + 0.05) ** 2) and np.all(r2[near:] >= (CRF + 0.05) ** 2)
        assert np.all(np.diff(r2[near:]) >= 0)                      # the far part is sorted by separation
        prefix = pl["prefix"]
        assert np.all(np.diff(prefix) >= 0) and prefix[0] >= near and prefix[-1] <= 228
        for L in range(41):
            inside = r2 < (CRF + 0.01 * L) ** 2
            assert np.all(np.nonzero(inside)[0] < prefix[L]), L     # nothing that could be in range is cut off
        # decide() keeps every atom within 0.2a of its site: level 40 must hold every pair that can ever be in range
        assert prefix[40] >= np.count_nonzero(r2 < (CRF + 0.4) ** 2)
    for pj, pl in enumerate(plans):
        half = pl["n_half"]
        for m in range(half):
            o = int(pl["sorted"][half + m])                          # a lower near offset of parity pj: neighbour i = j + o
            pi = pj ^ (o & 1)
            assert int(plans[pi]["sorted"][pl["lower_slot"][m]]) == -o
        vec = [sep2(int(o), pj)[1] for o in pl["sorted"][:half]]
        assert all(v[2] > 0 or (v[2] == 0 and (v[1] > 0 or (v[1] == 0 and v[0] > 0))) for v in vec)
        assert sorted(vec) == sorted(tuple(-c for c in sep2(int(o), pj)[1]) for o in pl["sorted"][half:2 * half])


@pytest.mark.parametrize("phase,grid", [((100, 100, 100), (1, 1, 1)), ((80, 8, 8), (2, 1, 1)), ((12, 12, 12), (2, 2, 2)), ((200, 200, 200), (2, 2, 2)),
                                        ((74, 20, 9), (1, 1, 1))])
def test_work_regions_partition_the_owned_box(phase, grid):
    """Every owned cell is computed by exactly one box of a launch; the interior keeps the stencil reach (3 cells) from
    every face -- and 15 more along x when the neighbours' push is awaited inside the kernel (which = 3)."""
    dom = capi.make_domain(phase, grid, (0, 0, 0), A, CRF)
    n = [int(v) for v in dom.sub_box_lattice_size]
    g = [int(v) for v in dom.lattice_size_ghost]

    def cover(boxes):
        cnt = np.zeros((n[2], n[1], n[0]), dtype=np.int32)
        for x0, y0, z0, nx, ny, nz in boxes:
            assert x0 >= 0 and y0 >= 0 and z0 >= 0 and x0 + nx <= n[0] and y0 + ny <= n[1] and z0 + nz <= n[2]
            cnt[z0:z0 + nz, y0:y0 + ny, x0:x0 + nx] += 1
        return cnt

    whole, units, split = capi.plan_regions(dom, 0)
    assert len(whole) == 1 and split == 0 and np.all(cover(whole) == 1)
    assert units == (n[0] * n[1] * n[2] + 31) // 32
    interior, _, _ = capi.plan_regions(dom, 1)
    slabs, _, _ = capi.plan_regions(dom, 2)
    if all(n[k] > 2 * g[k] for k in range(3)):
        assert np.all(cover(list(interior) + list(slabs)) == 1)
        x0, y0, z0, nx, ny, nz = interior[0]
        assert (x0, y0, z0) == tuple(g) and (x0 + nx, y0 + ny, z0 + nz) == tuple(n[k] - g[k] for k in range(3))
    late, units3, split3 = capi.plan_regions(dom, 3)
    assert np.all(cover(late) == 1)
    if split3 > 0:
        x0, y0, z0, nx, ny, nz = late[0]                       # the interior comes first
        assert split3 == (nx * ny * nz + 31) // 32
        assert x0 >= g[0] + 15 and x0 + nx <= n[0] - g[0] - 15   # no 128-byte line (16 cells) with a ghost site is read
        assert y0 >= g[1] and y0 + ny <= n[1] - g[1] and z0 >= g[2] and z0 + nz <= n[2] - g[2]
    else:
        assert len(late) == 1
    assert (split3 > 0) == (n[0] > 2 * (g[0] + 15) and n[1] > 2 * g[1] and n[2] > 2 * g[2])


@pytest.mark.parametrize("phase,which", [((40, 9, 8), 3), ((40, 9, 8), 0), ((12, 12, 12), 3), ((12, 12, 12), 2)])
def test_unit_order_is_a_bijection_with_the_interior_first(phase, which):
    """The kernels' unit_split: every (sub-lattice, unit) exactly once; with which = 3 no boundary unit comes before the
    last interior unit of EITHER sub-lattice (that is when a warp starts to wait for the neighbours' push)."""
    dom = capi.make_domain(phase, (1, 1, 1), (0, 0, 0), A, CRF)
    _, units, split = capi.plan_regions(dom, which)
    seen = [capi.plan_unit_order(dom, which, u) for u in range(2 * units)]
    assert sorted(seen) == [(p, k) for p in range(2) for k in range(units)]
    if split > 0:
        assert all(k < split for _, k in seen[:2 * split]) and all(k >= split for _, k in seen[2 * split:])
    else:
        assert seen == [(p, k) for p in range(2) for k in range(units)]


@pytest.mark.parametrize("cells", [(12, 8, 10), (6, 6, 6), (40, 7, 9), (5, 6, 7)])
def test_arithmetic_push_of_the_fused_kernels_equals_the_composed_map(cells):
    """kernels.cuh:push_site (positions pushed from inside k_verlet1, df from k_rho_f's epilogue) derives a band site's
    destinations arithmetically: per dimension stay, or +n cells from the low band (origin code +1), -n from the high band
    (code -1). That map must be misa_b200_plan_push's -- the composition of the reference's three staged exchanges
    (src/pack/lat_particle_packer.cpp:97-139) -- entry for entry, sub-boxes thinner than two ghost widths included."""
    dom = capi.make_domain(cells, (1, 1, 1), (0, 0, 0), A, CRF)
    dst, src, code, _ = capi.plan_push(dom)
    n = [int(v) for v in dom.sub_box_lattice_size]
    g = [int(v) for v in dom.lattice_size_ghost]
    sxc, sy, sz = n[0] + 2 * g[0], n[1] + 2 * g[1], n[2] + 2 * g[2]
    H = sxc * sy * sz
    to_dev = lambda idx: (idx >> 1) + (idx & 1) * H
    want = set(zip(to_dev(src).tolist(), code.tolist(), to_dev(dst).tolist()))
    got = set()
    opts = lambda c, nn, gg: [0] + ([1] if c < gg else []) + ([-1] if c >= nn - gg else [])
    for par in range(2):
        for z in range(n[2]):
            for y in range(n[1]):
                for cx in range(n[0]):
                    d = par * H + ((z + g[2]) * sy + (y + g[1])) * sxc + cx + g[0]
                    for s2 in opts(z, n[2], g[2]):
                        for s1 in opts(y, n[1], g[1]):
                            for s0 in opts(cx, n[0], g[0]):
                                if s0 or s1 or s2:
                                    got.add((d, (s0 + 1) + 3 * (s1 + 1) + 9 * (s2 + 1), d + s0 * n[0] + (s2 * n[2] * sy + s1 * n[1]) * sxc))
    assert got == want
