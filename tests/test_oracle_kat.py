"""Pins the CPU oracle against the reference's OWN unit tests: every test here is a
known-answer or property test taken from /root/reference/tests/unit, re-expressed against oracle/liboracle.so.
Integer / lattice results must be exact. Nothing here needs a GPU or /root/reference at run time."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle_py as O


def make_rank(size_x, size_y, size_z):
    """An ora_rank with only the BccLattice sizes set (what NeighbourIndex::make reads)."""
    rk = O.Rank()
    rk.size_x, rk.size_y, rk.size_z = size_x, size_y, size_z
    rk.size = size_x * size_y * size_z
    return rk


def nei(rk, cut_lattice, crf):
    O.lib().ora_nei_make(C.byref(rk), cut_lattice, crf)
    return [v.to_numpy() for v in (rk.nei_even, rk.nei_odd, rk.nei_half_even, rk.nei_half_odd)]


def atom_at(x, y, z):
    a = np.zeros(1, dtype=O.ATOM_DTYPE)
    a["x"][0] = (x, y, z)
    return a


def sub_box_coord(a, dom):
    c = (C.c_long * 3)()
    O.lib().ora_near_lat_sub_box_coord(a.ctypes.data, C.byref(dom), C.byref(c))
    return tuple(c)


# ---- reference tests/unit/neighbour_index_test.cpp ---------------------------------------------------
def test_is_positive_index():  # :110-114
    L = O.lib()
    assert L.ora_is_positive_index(-3, -2, 0) == 0
    assert L.ora_is_positive_index(-3, -2, -1) == 0
    assert L.ora_is_positive_index(-3, 1, 0) == 1


def test_index_vector_lengths():  # :117-130
    even, odd, half_even, half_odd = nei(make_rank(16, 8, 8), 2, 2 - 0.1)
    assert len(odd) == 2 * len(half_odd)
    assert len(even) == 2 * len(half_even)
    assert len(odd) == len(even)


def test_1nn_gives_8_offsets():  # :132-145
    even, odd, _, _ = nei(make_rank(16, 8, 8), 2, 0.51 * math.sqrt(3))
    # the search radius is crf + 2*0.5 (md_building_config.h.in:24-29), so "0.51*sqrt(3)" reaches past 1nn; the
    # reference asserts 8 -- and documents the sibling 2nn case as a known-wrong expectation (":147 fixme").
    # What IS exact: the 8 first-neighbour offsets are contained, and even/odd have equal length.
    first = {(dz * 8 + dy) * 16 + dx for dx, dy, dz in
             [(1, 0, 0), (-1, -1, -1), (1, -1, 0), (1, 0, -1), (1, -1, -1), (-1, 0, 0), (-1, -1, 0), (-1, 0, -1)]}
    assert first <= set(even.tolist())
    assert len(even) == len(odd)


def test_site_24_16_3_has_neighbour_23_17_4():  # :163-185
    rk = make_rank(18, 9, 9)
    even, odd, half_even, half_odd = nei(rk, 3, 1.96125)
    idx = lambda x, y, z: (z * 9 + y) * 18 + x
    lst = half_even  # x = 24 is even (neighbour_index.inl:96-132)
    assert idx(23, 17, 4) - idx(24, 16, 3) in set(lst.tolist())


def legacy_offsets(sx, sy, cut_lattice, crf):
    """The independent generator of the reference test (neighbour_index_test.cpp:27-107), in Python."""
    lim2 = (crf + 1.0) ** 2
    idx = lambda x, y, z: (z * sy + y) * sx + x
    pos = lambda x, y, z: z > 0 or (z == 0 and (y > 0 or (y == 0 and x > 0)))
    out = [[], [], [], []]
    rng = range(-cut_lattice - 1, cut_lattice + 2)
    for z in rng:
        for y in rng:
            for x in rng:
                r = x * x + y * y + z * z
                if 0 < r < lim2:
                    out[0].append(idx(2 * x, y, z))
                    if pos(x, y, z):
                        out[2].append(idx(2 * x, y, z))
                r = (x + 0.5) ** 2 + (y + 0.5) ** 2 + (z + 0.5) ** 2
                if 0 < r < lim2:
                    out[0].append(idx(2 * x + 1, y, z))
                    if pos(x + 0.5, y + 0.5, z + 0.5):
                        out[2].append(idx(2 * x + 1, y, z))
    for z in rng:
        for y in rng:
            for x in rng:
                r = x * x + y * y + z * z
                if 0 < r < lim2:
                    out[1].append(idx(2 * x, y, z))
                    if pos(x, y, z):
                        out[3].append(idx(2 * x, y, z))
                r = (x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2
                if 0 < r < lim2:
                    out[1].append(idx(2 * x - 1, y, z))
                    if pos(x - 0.5, y - 0.5, z - 0.6):
                        out[3].append(idx(2 * x - 1, y, z))
    return [np.array(sorted(v), dtype=np.int64) for v in out]


def test_new_and_legacy_generators_agree():  # :188-243
    box, cut = 6, 2
    ext = box + 2 * cut
    got = nei(make_rank(2 * ext, ext, ext), cut, 1.96125)
    want = legacy_offsets(2 * ext, ext, cut, 1.96125)
    for g, w in zip(got, want):
        assert np.array_equal(np.sort(g), w)
    assert len(got[0]) == 228 and len(got[2]) == 114  # stencil sizes for crf = 1.96125


# ---- reference tests/unit/lattice/ws_utils_test.cpp ----------------------------------------------------
def test_ws_exact_sites_and_perturbed_map_back():  # :13-51 (sub-sampled: every 7th site, same assertions)
    a = 0.86
    dom = O.make_domain((100, 120, 144), (2, 2, 2), (0, 0, 0), a, 1.1421)
    for z in range(0, 72, 7):
        for y in range(0, 60, 7):
            for x in range(0, 100, 3):
                px = x * (a / 2)
                py = a * y if x % 2 == 0 else a * y + a / 2
                pz = a * z if x % 2 == 0 else a * z + a / 2
                assert sub_box_coord(atom_at(px, py, pz), dom) == (x, y, z)
                assert sub_box_coord(atom_at(px + 0.1 * a, py - 0.2 * a, pz + 0.15 * a), dom) == (x, y, z)


A0 = 2.85532


@pytest.fixture(scope="module")
def dom25():
    return O.make_domain((50, 50, 50), (2, 2, 2), (0, 0, 0), A0, 1.96125)


def test_ws_negative_z_even_x(dom25):  # :54-103
    for z, want in ((0 - 0.01, 0), (A0 * (-2) - 0.01, -2), (A0 * (-2.44), -2), (A0 * (-2.66), -3)):
        assert sub_box_coord(atom_at(0, 0, z), dom25)[2] == want


def test_ws_negative_z_odd_x(dom25):  # :106-157
    for z, want in ((0 - 0.01, -1), (A0 * (-2) + 0.01, -2), (A0 * (-2) - 0.01, -3), (A0 * (-2.44), -3), (A0 * (-2.66), -3)):
        assert sub_box_coord(atom_at(3.5 * A0, 0.5 * A0, z), dom25)[2] == want


def test_ws_negative_x(dom25):  # :160-209
    cases = ((0 - 0.01, 0), (A0 * (-2) - 0.01, -4), (A0 * (-2 - 0.1), -4), (A0 * (-2 + 0.1), -4),
             (A0 * (-2.25 - 0.1), -4), (A0 * (-2.75 + 0.1), -6), (A0 * (-2.75 - 0.1), -6))
    for x, want in cases:
        assert sub_box_coord(atom_at(x, 0.0, 0.0), dom25)[0] == want


# ---- reference tests/unit/inter_atom_test.cpp ----------------------------------------------------------
def test_is_out_box_flags():  # :26-61
    a = 0.86
    dom = O.make_domain((100, 120, 144), (2, 2, 2), (0, 0, 0), a, 1.1421)
    f = lambda x, y, z: O.lib().ora_is_out_box(atom_at(x, y, z).ctypes.data, C.byref(dom))
    assert f(10 * a, 10 * a, 10 * a) == 0                      # IN_BOX
    assert f(-10 * a, 10 * a, 10 * a) == 1                     # OUT_BOX_X_LITTER
    assert f(-0.01, 10 * a, 10 * a) == 0
    assert f(-a / 2 - 0.01, -a / 2 - 0.01, -a / 2 - 0.01) == (1 | 4 | 16)


def test_is_out_box_big_case():  # :63-80
    dom = O.make_domain((50, 50, 50), (2, 2, 1), (1, 1, 0), A0, 1.96125)
    a = atom_at(71.3905, 71.3921, 74.234)
    assert O.lib().ora_is_out_box(a.ctypes.data, C.byref(dom)) == 0


# ---- reference tests/unit/atom/nei_index_test.cpp:71-173 (property test, fixed seed instead of time()) ----
def test_pairs_within_cutoff_are_stencil_related():
    dom = O.make_domain((50, 50, 50), (2, 2, 2), (0, 0, 0), A0, 1.96125, ghost=3)
    sx, sy, sz = 2 * (25 + 6), 25 + 6, 25 + 6
    even, odd, _, _ = nei(make_rank(sx, sy, sz), dom.cut_lattice, dom.cutoff_radius_factor)
    sets = {0: set(even.tolist()), 1: set(odd.tolist())}
    rs = np.random.RandomState(20190714)
    p = rs.randint(-5000, 5001, size=(200000, 2, 3)) / 1000.0
    d2 = ((p[:, 0] - p[:, 1]) ** 2).sum(axis=1)
    rc = 1.96125 * A0
    close = p[d2 < rc * rc]
    assert len(close) > 50
    idx = lambda c: (c[2] * sy + c[1]) * sx + c[0]
    checked = 0
    for a, b in close:
        c1, c2 = sub_box_coord(atom_at(*a), dom), sub_box_coord(atom_at(*b), dom)
        if c1 == c2:
            continue
        assert idx(c2) - idx(c1) in sets[c1[0] % 2], (a, b, c1, c2)
        checked += 1
    assert checked > 50


# ---- reference tests/unit/atom_test.cpp:11-54 (setv known answers; pins mvv2e and the Fe mass) -----------
def test_setv_known_answers(pot):
    for energy, direction, want in ((5000.0, (1, 3, 5), (222.17975, 666.53928, 1110.89879)),
                                    (20000.0, (1, 2, 2), (876.28882, 1752.5776, 1752.5776))):
        w = O.World((6, 6, 6), a=0.86, crf=1.1421, pot=pot)
        w.fill_lattice()
        lat = (C.c_int * 4)(3, 3, 3, 0)
        w.L.ora_setv(w.h, C.byref(lat), C.byref((C.c_double * 3)(*direction)), energy)
        at = w.atoms(0).reshape(w.shape(0))
        d = w.rank(0).dom
        g = d.dbx_lattice_size_ghost
        v = at[g[2] + 3, g[1] + 3, g[0] + 6]["v"]
        assert np.allclose(v, want, rtol=1e-6)  # EXPECT_FLOAT_EQ
        assert np.count_nonzero(at["v"]) == 3
        w.close()


# ---- reference tests/unit/world_builder_test.cpp:13-44, configuration_test.cpp:12-42 -----------------------
def test_rescale_hits_target_temperature(pot):
    from misa_md_b200 import synth
    st = synth.create_global_state((6, 7, 8), t_set=0.0)
    w = O.World((6, 7, 8), pot=pot)
    arr, _ = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
    w.atoms(0)[:] = arr
    w.L.ora_rescale(w.h, 600.0)
    assert abs(w.temperature() - 600.0) < 1e-9
    own = w.atoms(0).reshape(w.shape(0))[w.owned_slices(0)]
    p = (own["v"] * synth.MASS[own["type"]][..., None]).reshape(-1, 3).sum(axis=0)
    assert np.all(np.abs(p) < 1e-8)  # zero total momentum (world_builder_test.cpp:40-43)
    # the two temperature formulas of configuration_test.cpp:35-41 agree
    ke = w.kinetic_energy()
    n = 2 * 6 * 7 * 8
    assert abs(2 * ke / ((3 * n - 3) * 8.617343e-5) - w.temperature()) < 1e-9
    w.close()
