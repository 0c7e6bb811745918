"""Shared helpers for the parity tests: identical seeded inputs for the CPU oracle and the CUDA path."""
import numpy as np

import misa_md_b200 as mb
from misa_md_b200 import synth
from oracle import oracle_py as O

A = 2.85532
CRF = 1.96125

_POT_HOST = None


def host_potential():
    """(elec, embed, phi) spline tables in atom_type enum order, built by the product's host code."""
    global _POT_HOST
    if _POT_HOST is None:
        _POT_HOST = mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH))
    return _POT_HOST


def make_state(phase_space, ratio=(1, 0, 0), sigma=0.0, seed=466953, t_set=600.0, vacancies=0):
    st = synth.create_global_state(phase_space, a=A, seed=seed, t_set=t_set, ratio=ratio)
    if sigma:
        synth.perturb_positions(st, sigma)
    if vacancies:
        rs = np.random.RandomState(99)
        flat = st["type"].reshape(-1)
        flat[rs.choice(flat.size, vacancies, replace=False)] = synth.INVALID
    return st


def oracle_world(state, grid=(1, 1, 1), dt=0.001, threads=1):
    w = O.World(state["phase_space"], grid=grid, a=A, crf=CRF, dt=dt, threads=threads)
    for r in range(w.n_ranks):
        d = w.rank(r).dom
        arr, _ = synth.scatter_to_sub_box(state, grid, tuple(d.grid_coord), CRF)
        w.atoms(r)[:] = arr
    return w


def gpu_context(state, grid=(1, 1, 1), coord=(0, 0, 0), dt=0.001, upload=True):
    ctx = mb.Context(state["phase_space"], grid=grid, coord=coord, a=A, crf=CRF)
    ctx.make_offsets()
    ctx.set_potential(*host_potential())
    ctx.set_timestep(dt)
    if upload:
        arr, _ = synth.scatter_to_sub_box(state, grid, coord, CRF)
        ctx.upload(arr)
    return ctx


def owned(ctx_or_shape, arr, owned_slices=None):
    if owned_slices is None:
        return arr.reshape(ctx_or_shape.ext_shape)[ctx_or_shape.owned]
    return arr.reshape(ctx_or_shape)[owned_slices]


def rel_err(got, ref, scale=None):
    """max |got-ref| / max(|ref|, eps*scale) with scale = max|ref| (SURVEY.md section 8d parity gate)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = float(np.max(np.abs(ref))) if scale is None else scale
    if scale == 0.0:
        return float(np.max(np.abs(got)))
    den = np.maximum(np.abs(ref), 1e-3 * scale)
    return float(np.max(np.abs(got - ref) / den))


def per_atom_rel(got, ref, floor=1e-6):
    """north_star's gate "per-atom rho, df and forces within 1e-10 relative": max over atoms of
    |got_i - ref_i|_inf / max(|ref_i|_2, 1e-6 * max_j |ref_j|_2) -- every atom is judged against ITS OWN magnitude; the floor
    only guards atoms whose value vanishes by symmetry (1e-6 of the largest, not rel_err's 1e-3). For df = F'(rho), which
    changes sign near the equilibrium density, pass floor=1e-3: an atom's own |df| says nothing about the accuracy of F' there."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.ndim == 1:
        got, ref = got[:, None], ref[:, None]
    got, ref = got.reshape(-1, got.shape[-1]), ref.reshape(-1, ref.shape[-1])
    if ref.size == 0:
        return 0.0
    mag = np.sqrt((ref * ref).sum(axis=1))
    top = float(mag.max())
    den = np.maximum(mag, floor * top if top > 0 else 1.0)
    return float((np.abs(got - ref).max(axis=1) / den).max())


def oracle_threads(limit=8):
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(limit, n))


def oracle_global(w, fields=("id", "type", "x", "v", "f", "rho", "df")):
    """Owned records of every sub-box of an oracle world, assembled into the global (PZ, PY, 2 PX) box."""
    from misa_md_b200 import synth
    pz, py, px = w.phase_space[2], w.phase_space[1], w.phase_space[0]
    out = np.zeros((pz, py, 2 * px), dtype=synth.ATOM_DTYPE)
    for r in range(w.n_ranks):
        d = w.rank(r).dom
        n = [w.phase_space[k] // w.grid[k] for k in range(3)]
        lo = [d.grid_coord[k] * n[k] for k in range(3)]
        own = w.atoms(r).reshape(w.shape(r))[w.owned_slices(r)]
        out[lo[2]:lo[2] + n[2], lo[1]:lo[1] + n[1], 2 * lo[0]:2 * lo[0] + 2 * n[0]] = own
    return out
