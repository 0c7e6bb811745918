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
