"""Brute-force check, on CPU, of the rule by which a warp of the stencil kernels cuts its offset list:
    L = max over the warp's 32 atoms of lev_i  +  (lg if the warp sees a marked / edge cell else min(lg, T)),   n_off = prefix[L]
with lev = ceil(|x - site| / 0.01a), lg the level of the global maximum, atoms above the marking level T stamping the 7^3
cells around them. The list order and the prefixes come from the product's own planner (misa_b200_plan_stencil); levels,
marks and the warp -> cells mapping are restated here in numpy as kernels.cuh / eam_smem.cuh compute them. For random thermal
states with outliers, ANY marking level and every warp: no pair that is within the cutoff lies beyond the warp's prefix."""
import numpy as np
import pytest

from misa_md_b200 import capi

A, CRF = 2.85532, 1.96125
N = (100, 10, 10)       # owned cells: one periodic sub-box (1x1x1 grid); rows of 100 cells so that some 32-cell warp units
                        # keep clear of the cells next to the ghost shell (x < 3 or x >= 97) and take the cheap bound
G = 3


def decode(off, sx, sy):
    dx = ((off % sx) + sx + sx // 2) % sx - sx // 2
    r = (off - dx) // sx
    dy = ((r % sy) + sy + sy // 2) % sy - sy // 2
    return dx, dy, (r - dy) // sy


@pytest.mark.parametrize("seed,sigma,n_out", [(1, 0.08, 0), (2, 0.08, 6), (3, 0.12, 20), (4, 0.03, 3)])
def test_no_in_range_pair_beyond_the_warp_prefix(seed, sigma, n_out):
    rng = np.random.default_rng(seed)
    dom = capi.make_domain(N, (1, 1, 1), (0, 0, 0), A, CRF)
    sx, sy = 2 * (N[0] + 2 * G), N[1] + 2 * G
    nx, ny, nz = N
    # displacements per (parity, z, y, x); a few outliers close to decide()'s 0.2a bound
    u = rng.normal(0.0, sigma, size=(2, nz, ny, nx, 3))
    for _ in range(n_out):
        p, z, y, x = rng.integers(2), rng.integers(nz), rng.integers(ny), rng.integers(nx)
        v = rng.normal(size=3)
        u[p, z, y, x] = v / np.linalg.norm(v) * rng.uniform(0.15, 0.199) * A
    if n_out:
        # one crafted pair that ONLY the mark keeps: a body-centre atom 0.19a off its site towards a corner atom of its
        # <311>/2 shell two cells away (separation (1.5, 1.5, 0.5) a), that partner 0.04a towards it: 1.949a < crf
        n_hat = np.array([1.5, 1.5, 0.5]) / np.linalg.norm([1.5, 1.5, 0.5])
        u[1, 3, 3, 50] = 0.19 * A * n_hat          # centre of cell (50, 3, 3)
        u[0, 4, 5, 52] = -0.04 * A * n_hat         # corner of cell (52, 5, 4): its warp unit is cells 44..75 of that row
    norm = np.linalg.norm(u, axis=-1)
    assert norm.max() < 0.2 * A
    lev = np.ceil((norm + 2e-6) * (100.0 / A)).astype(int)                    # kernels.cuh:disp_level
    lg = int(np.ceil((norm.max() + 1e-6) / (0.01 * A)))                        # eam_smem.cuh:base_level
    plans = [capi.plan_stencil(dom, p) for p in range(2)]
    prefix = plans[0]["prefix"]
    cellvec = []
    for p in range(2):
        rows = []
        for o in plans[p]["sorted"]:
            dx, dy, dz = decode(int(o), sx, sy)
            q = p ^ (dx & 1)                                                  # sub-lattice of the neighbour
            rows.append(((p + dx) >> 1, dy, dz, q, 0.5 * dx, dy + (0.5 if (dx & 1 and p == 0) else -0.5 if dx & 1 else 0.0),
                         dz + (0.5 if (dx & 1 and p == 0) else -0.5 if dx & 1 else 0.0)))
        cellvec.append(rows)
    edge = np.ones((nz, ny, nx), dtype=bool)                                  # cells within the stencil reach of the ghost shell
    edge[G:nz - G, G:ny - G, G:nx - G] = False
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    checked = pruned = cheap = 0
    for T in (0, max(0, lg - 6), max(0, lg - 3), lg - 1, lg + 5):
        hot = edge.copy()
        for p, z, y, x in zip(*np.nonzero(lev > T)):                          # kernels.cuh:verlet1_site stamps 7^3 cells
            hot[max(0, z - G):z + G + 1, max(0, y - G):y + G + 1, max(0, x - G):x + G + 1] = True
        for p in range(2):
            flat_lev, flat_hot = lev[p].reshape(-1), hot.reshape(-1)          # cell index c = (z*ny + y)*nx + x, 32 per warp unit
            n_units = (flat_lev.size + 31) // 32
            unit_of = np.arange(flat_lev.size) // 32
            lw = np.zeros(n_units, dtype=int)
            np.maximum.at(lw, unit_of, flat_lev)
            uh = np.zeros(n_units, dtype=bool)
            np.logical_or.at(uh, unit_of, flat_hot)
            L = lw + np.where(uh, lg, min(lg, T))
            cheap += int((~uh).sum()) if T < lg else 0
            n_off = np.where(L < 41, prefix[np.minimum(L, 40)], 228)[unit_of].reshape(nz, ny, nx)
            pruned += int((n_off < 228).sum())
            for q, (dcx, dy, dz, par_j, X, Y, Z) in enumerate(cellvec[p]):
                cut = n_off <= q                                             # warps that do NOT visit offset q
                if not cut.any():
                    continue
                jx, jy, jz = (xx + dcx) % nx, (yy + dy) % ny, (zz + dz) % nz  # periodic partner
                d = np.array([X, Y, Z]) * A + u[par_j, jz, jy, jx] - u[p]
                in_range = np.einsum("...k,...k->...", d, d) < (A * CRF) ** 2
                checked += int(cut.sum())
                assert not np.any(in_range & cut), (T, p, q)
    assert checked > 0 and pruned > 0 and cheap > 0
