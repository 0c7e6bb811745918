"""ctypes binding of libmisa_b200.so (the C ABI in include/misa_b200.h).

Python is only the test / bench harness here; the product is the shared library. Loading fails loudly when
the library is missing and every compute call fails loudly when no CUDA device is present -- there is no
CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .synth import ATOM_DTYPE, DUMP_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
K_NAMES = ["verlet1", "halo_x", "rho", "df", "halo_df", "force", "verlet2", "inter", "xfer"]
K_COUNT = len(K_NAMES)

# every symbol declared in include/misa_b200.h (checked by tests/test_abi.py)
EXPORTS = [
    "misa_b200_env_init", "misa_b200_env_clean", "misa_b200_device_count", "misa_b200_last_error",
    "misa_b200_create", "misa_b200_destroy", "misa_b200_set_neighbour_offsets", "misa_b200_make_neighbour_offsets",
    "misa_b200_get_neighbour_offsets", "misa_b200_plan_offsets", "misa_b200_plan_halo", "misa_b200_plan_push", "misa_b200_plan_stencil", "misa_b200_plan_regions", "misa_b200_plan_unit_order", "misa_b200_set_potential",
    "misa_b200_eam_rho_calc", "misa_b200_eam_df_calc", "misa_b200_eam_force_calc",
    "misa_b200_site_count", "misa_b200_host_register", "misa_b200_host_unregister",
    "misa_b200_upload_atoms", "misa_b200_download_atoms", "misa_b200_upload_inter", "misa_b200_download_inter",
    "misa_b200_set_timestep", "misa_b200_prepare", "misa_b200_step", "misa_b200_step_host", "misa_b200_setv", "misa_b200_collision_step",
    "misa_b200_rescale", "misa_b200_thermo", "misa_b200_sync",
    "misa_b200_pass_halo_x", "misa_b200_pass_clear", "misa_b200_pass_rho", "misa_b200_pass_df", "misa_b200_pass_halo_df",
    "misa_b200_pass_force", "misa_b200_pass_verlet1", "misa_b200_pass_verlet2", "misa_b200_set_option", "misa_b200_query",
    "misa_b200_comm_unique_id", "misa_b200_comm_init", "misa_b200_comm_destroy",
    "misa_b200_profile_enable", "misa_b200_profile_read", "misa_b200_launch_count", "misa_b200_timed_steps",
    "misa_b200_stencil_stats", "misa_b200_build_world", "misa_b200_temperature", "misa_b200_rescale_to", "misa_b200_dump_records",
]


class Domain(C.Structure):
    _fields_ = [
        ("phase_space", C.c_int64 * 3), ("grid_size", C.c_int32 * 3), ("grid_coord", C.c_int32 * 3),
        ("sub_box_lattice_size", C.c_int32 * 3), ("lattice_size_ghost", C.c_int32 * 3), ("sub_box_lattice_low", C.c_int32 * 3),
        ("rank_id_neighbours", (C.c_int32 * 2) * 3), ("rank", C.c_int32),
        ("lattice_const", C.c_double), ("cutoff_radius_factor", C.c_double), ("meas_global_length", C.c_double * 3),
    ]


class Table(C.Structure):
    _fields_ = [("n", C.c_int32), ("inv_dx", C.c_double), ("spline", C.POINTER(C.c_double))]


class MisaError(RuntimeError):
    pass


_lib = None


def load(build=True):
    """dlopen libmisa_b200.so (building it in-tree first if nvcc is around and sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("MISA_B200_LIB")  # A/B timing of two builds of the SAME library; never a fallback
    if override:
        _build.LIB = override
        build = False
    if build:
        try:
            _build.build()
        except Exception:
            if not os.path.exists(_build.LIB):
                raise
    if not os.path.exists(_build.LIB):
        raise MisaError("libmisa_b200.so is missing: run `python -m misa_md_b200.build` (no CPU fallback exists)")
    L = C.CDLL(_build.LIB, mode=C.RTLD_GLOBAL)
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    L.misa_b200_last_error.restype = C.c_char_p
    L.misa_b200_env_init.argtypes = [i]
    L.misa_b200_create.argtypes = [C.POINTER(Domain), C.POINTER(vp)]
    L.misa_b200_destroy.argtypes = [vp]
    i64p = C.POINTER(C.c_int64)
    L.misa_b200_set_neighbour_offsets.argtypes = [vp, i64p, C.c_size_t, i64p, C.c_size_t, i64p, C.c_size_t, i64p, C.c_size_t]
    L.misa_b200_make_neighbour_offsets.argtypes = [vp, i, d]
    L.misa_b200_get_neighbour_offsets.argtypes = [vp, i, i64p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.misa_b200_plan_offsets.argtypes = [C.POINTER(Domain), i, d, i, i64p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.misa_b200_plan_halo.argtypes = [C.POINTER(Domain), i, i, i64p, i64p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(d * 3)]
    L.misa_b200_plan_stencil.argtypes = [C.POINTER(Domain), i, d, i, i64p, C.POINTER(d), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.misa_b200_plan_regions.argtypes = [C.POINTER(Domain), i, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.misa_b200_plan_unit_order.argtypes = [C.POINTER(Domain), i, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.misa_b200_plan_push.argtypes = [C.POINTER(Domain), i64p, i64p, C.POINTER(C.c_int8), C.c_size_t, C.POINTER(C.c_size_t), vp]
    L.misa_b200_set_potential.argtypes = [vp, i, C.POINTER(Table), C.POINTER(Table), C.POINTER(Table)]
    for fn in ("misa_b200_eam_rho_calc", "misa_b200_eam_df_calc", "misa_b200_eam_force_calc"):
        getattr(L, fn).argtypes = [vp, vp, d]
    L.misa_b200_site_count.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.misa_b200_host_register.argtypes = [vp, C.c_size_t]
    L.misa_b200_host_unregister.argtypes = [vp]
    L.misa_b200_upload_atoms.argtypes = [vp, vp]
    L.misa_b200_download_atoms.argtypes = [vp, vp]
    L.misa_b200_upload_inter.argtypes = [vp, vp, C.c_size_t]
    L.misa_b200_download_inter.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.misa_b200_set_timestep.argtypes = [vp, d]
    L.misa_b200_prepare.argtypes = [vp]
    L.misa_b200_step.argtypes = [vp, i]
    L.misa_b200_step_host.argtypes = [vp, vp, i]
    L.misa_b200_setv.argtypes = [vp, C.POINTER(C.c_int32 * 4), C.POINTER(d * 3), d]
    L.misa_b200_collision_step.argtypes = [vp, C.POINTER(C.c_int32 * 4), C.POINTER(d * 3), d]
    L.misa_b200_rescale.argtypes = [vp, d, d]
    L.misa_b200_thermo.argtypes = [vp, C.POINTER(d * 6)]
    for fn in ("misa_b200_sync", "misa_b200_pass_halo_x", "misa_b200_pass_clear", "misa_b200_pass_rho", "misa_b200_pass_df",
               "misa_b200_pass_halo_df", "misa_b200_pass_force", "misa_b200_pass_verlet1", "misa_b200_pass_verlet2",
               "misa_b200_comm_destroy"):
        getattr(L, fn).argtypes = [vp]
    L.misa_b200_set_option.argtypes = [vp, C.c_char_p, i]
    L.misa_b200_query.argtypes = [vp, C.c_char_p, C.POINTER(d)]
    L.misa_b200_comm_unique_id.argtypes = [vp]
    L.misa_b200_comm_init.argtypes = [vp, vp, i, i]
    L.misa_b200_profile_enable.argtypes = [vp, i]
    L.misa_b200_profile_read.argtypes = [vp, C.POINTER(d * K_COUNT), C.POINTER(C.c_int64 * K_COUNT)]
    L.misa_b200_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    L.misa_b200_timed_steps.argtypes = [vp, i, C.POINTER(d)]
    L.misa_b200_stencil_stats.argtypes = [vp, C.POINTER(d * 4)]
    L.misa_b200_build_world.argtypes = [vp, C.c_uint32, d, C.POINTER(C.c_int32 * 3), C.c_uint64]
    L.misa_b200_temperature.argtypes = [vp, C.c_uint64, C.POINTER(d * 4)]
    L.misa_b200_rescale_to.argtypes = [vp, d, C.c_uint64]
    L.misa_b200_dump_records.argtypes = [vp, C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), C.c_uint64, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise MisaError("misa_b200 error %d: %s" % (rc, load().misa_b200_last_error().decode()))


def device_count():
    return load().misa_b200_device_count()


def make_domain(phase_space, grid=(1, 1, 1), coord=(0, 0, 0), a=2.85532, crf=1.96125, ghost=None):
    """Flatten what comm::BccDomain::Builder computes (reference src/simulation.cpp:41-47) into the ABI struct."""
    import math
    ghost = int(math.ceil(crf)) + 1 if ghost is None else ghost
    dom = Domain()
    rank_of = lambda c: (c[0] * grid[1] + c[1]) * grid[2] + c[2]  # MPI_Cart order
    for k in range(3):
        n = phase_space[k] // grid[k]
        if n * grid[k] != phase_space[k]:
            raise ValueError("phase space must divide evenly by the process grid")
        dom.phase_space[k] = phase_space[k]
        dom.grid_size[k] = grid[k]
        dom.grid_coord[k] = coord[k]
        dom.sub_box_lattice_size[k] = n
        dom.lattice_size_ghost[k] = ghost
        dom.sub_box_lattice_low[k] = coord[k] * n
        dom.meas_global_length[k] = phase_space[k] * a
        lo, hi = list(coord), list(coord)
        lo[k] = (coord[k] - 1) % grid[k]
        hi[k] = (coord[k] + 1) % grid[k]
        dom.rank_id_neighbours[k][0] = rank_of(lo)
        dom.rank_id_neighbours[k][1] = rank_of(hi)
    dom.rank = rank_of(coord)
    dom.lattice_const = a
    dom.cutoff_radius_factor = crf
    return dom


def plan_offsets(dom, which, cut_lattice=None, crf=None):
    """Host-only: neighbour offsets (reference index space) of list `which` (0 even, 1 odd, 2 half_even, 3 half_odd)."""
    import math
    L = load()
    crf = dom.cutoff_radius_factor if crf is None else crf
    cut_lattice = int(math.ceil(crf)) if cut_lattice is None else cut_lattice
    n = C.c_size_t()
    _ck(L.misa_b200_plan_offsets(C.byref(dom), cut_lattice, crf, which, None, 0, C.byref(n)))
    out = np.zeros(n.value, dtype=np.int64)
    _ck(L.misa_b200_plan_offsets(C.byref(dom), cut_lattice, crf, which, out.ctypes.data_as(C.POINTER(C.c_int64)), n.value, C.byref(n)))
    return out


def plan_halo(dom, dim, direction):
    """Host-only: (send indices, recv indices, shift[3]) of halo message (dim, direction), reference index space."""
    L = load()
    n = C.c_size_t()
    _ck(L.misa_b200_plan_halo(C.byref(dom), dim, direction, None, None, 0, C.byref(n), None))
    send = np.zeros(n.value, dtype=np.int64)
    recv = np.zeros(n.value, dtype=np.int64)
    shift = (C.c_double * 3)()
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    _ck(L.misa_b200_plan_halo(C.byref(dom), dim, direction, p(send), p(recv), n.value, C.byref(n), C.byref(shift)))
    return send, recv, np.array(list(shift))


def plan_stencil(dom, parity, cut_lattice=None, crf=None):
    """Host-only: dict(sorted, site_r2, n_near, n_half, prefix[41], lower_slot) -- include/misa_b200.h:misa_b200_plan_stencil."""
    import math
    L = load()
    crf = dom.cutoff_radius_factor if crf is None else crf
    cut_lattice = int(math.ceil(crf)) if cut_lattice is None else cut_lattice
    n = C.c_size_t()
    _ck(L.misa_b200_plan_stencil(C.byref(dom), cut_lattice, crf, parity, None, None, 0, C.byref(n), None, None, None, None))
    srt = np.zeros(n.value, dtype=np.int64)
    r2 = np.zeros(n.value, dtype=np.float64)
    near, half = C.c_int32(), C.c_int32()
    prefix = np.zeros(41, dtype=np.int32)
    slot = np.zeros(n.value, dtype=np.int32)
    _ck(L.misa_b200_plan_stencil(C.byref(dom), cut_lattice, crf, parity, srt.ctypes.data_as(C.POINTER(C.c_int64)),
                                 r2.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n), C.byref(near), C.byref(half),
                                 prefix.ctypes.data_as(C.POINTER(C.c_int32)), slot.ctypes.data_as(C.POINTER(C.c_int32))))
    return dict(sorted=srt, site_r2=r2, n_near=near.value, n_half=half.value, prefix=prefix, lower_slot=slot[:half.value])


def plan_regions(dom, which):
    """Host-only: (boxes [n][6] = x0,y0,z0,nx,ny,nz in owned-cell coordinates, warp units per parity, interior units)."""
    L = load()
    boxes = np.zeros((7, 6), dtype=np.int32)
    n, units, split = C.c_int32(), C.c_int64(), C.c_int64()
    _ck(L.misa_b200_plan_regions(C.byref(dom), which, boxes.ctypes.data_as(C.c_void_p), C.byref(n), C.byref(units), C.byref(split)))
    return boxes[:n.value].copy(), units.value, split.value


def plan_unit_order(dom, which, u):
    """Host-only: (sub-lattice, unit) visited as the u-th warp unit of a launch over plan_regions(dom, which)."""
    L = load()
    par, unit = C.c_int32(), C.c_int64()
    _ck(L.misa_b200_plan_unit_order(C.byref(dom), which, int(u), C.byref(par), C.byref(unit)))
    return par.value, unit.value


def plan_push(dom):
    """Host-only: the staged exchanges composed into one ghost <- owned map: (dst, src, code, shift[27][3]); see
    include/misa_b200.h:misa_b200_plan_push."""
    L = load()
    n = C.c_size_t()
    _ck(L.misa_b200_plan_push(C.byref(dom), None, None, None, 0, C.byref(n), None))
    dst = np.zeros(n.value, dtype=np.int64)
    src = np.zeros(n.value, dtype=np.int64)
    code = np.zeros(n.value, dtype=np.int8)
    shift = np.zeros((27, 3), dtype=np.float64)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    _ck(L.misa_b200_plan_push(C.byref(dom), p(dst), p(src), code.ctypes.data_as(C.POINTER(C.c_int8)), n.value, C.byref(n),
                              shift.ctypes.data_as(C.c_void_p)))
    return dst, src, code, shift


class Context:
    """One sub-box on one GPU (thin OO wrapper over the C ABI)."""

    def __init__(self, phase_space, grid=(1, 1, 1), coord=(0, 0, 0), a=2.85532, crf=1.96125, device=None):
        self.L = load()
        if device is not None:
            _ck(self.L.misa_b200_env_init(device))
        self.dom = make_domain(phase_space, grid, coord, a, crf)
        self.h = C.c_void_p()
        _ck(self.L.misa_b200_create(C.byref(self.dom), C.byref(self.h)))
        g = self.dom.lattice_size_ghost
        n = self.dom.sub_box_lattice_size
        self.ext_shape = (n[2] + 2 * g[2], n[1] + 2 * g[1], 2 * (n[0] + 2 * g[0]))
        self.n_ext = self.ext_shape[0] * self.ext_shape[1] * self.ext_shape[2]
        self.n_owned = 2 * n[0] * n[1] * n[2]
        self.owned = (slice(g[2], g[2] + n[2]), slice(g[1], g[1] + n[1]), slice(2 * g[0], 2 * g[0] + 2 * n[0]))
        self.cutoff_radius = a * crf
        self._keep = []

    def close(self):
        if self.h:
            self.L.misa_b200_destroy(self.h)
            self.h = C.c_void_p()

    # ---- setup -------------------------------------------------------------------------------
    def make_offsets(self, cut_lattice=None, crf=None):
        import math
        crf = self.dom.cutoff_radius_factor if crf is None else crf
        cut_lattice = int(math.ceil(crf)) if cut_lattice is None else cut_lattice
        _ck(self.L.misa_b200_make_neighbour_offsets(self.h, cut_lattice, crf))

    def set_offsets(self, even, odd, half_even, half_odd):
        arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in (even, odd, half_even, half_odd)]
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        _ck(self.L.misa_b200_set_neighbour_offsets(self.h, p(arrs[0]), len(arrs[0]), p(arrs[1]), len(arrs[1]),
                                                   p(arrs[2]), len(arrs[2]), p(arrs[3]), len(arrs[3])))

    def get_offsets(self, which):
        n = C.c_size_t()
        _ck(self.L.misa_b200_get_neighbour_offsets(self.h, which, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int64)
        _ck(self.L.misa_b200_get_neighbour_offsets(self.h, which, out.ctypes.data_as(C.POINTER(C.c_int64)), n.value, C.byref(n)))
        return out

    def set_potential(self, elec, embed, phi):
        """elec/embed: lists (atom_type enum order) of (n, inv_dx, spline[(n+1)*7]); phi: n_types x n_types nested list."""
        nt = len(elec)

        def tab(t):
            n, inv_dx, sp = t
            sp = np.ascontiguousarray(sp, dtype=np.float64)
            assert sp.size == (n + 1) * 7
            self._keep.append(sp)
            return Table(n, inv_dx, sp.ctypes.data_as(C.POINTER(C.c_double)))

        e = (Table * nt)(*[tab(t) for t in elec])
        f = (Table * nt)(*[tab(t) for t in embed])
        ph = (Table * (nt * nt))(*[tab(phi[i][j]) for i in range(nt) for j in range(nt)])
        _ck(self.L.misa_b200_set_potential(self.h, nt, e, f, ph))

    def set_option(self, name, value):
        _ck(self.L.misa_b200_set_option(self.h, name.encode(), int(value)))

    def query(self, name):
        v = C.c_double()
        _ck(self.L.misa_b200_query(self.h, name.encode(), C.byref(v)))
        return v.value

    # ---- compat hooks ------------------------------------------------------------------------
    def eam_rho_calc(self, atoms):
        _ck(self.L.misa_b200_eam_rho_calc(self.h, atoms.ctypes.data, self.cutoff_radius))

    def eam_df_calc(self, atoms):
        _ck(self.L.misa_b200_eam_df_calc(self.h, atoms.ctypes.data, self.cutoff_radius))

    def eam_force_calc(self, atoms):
        _ck(self.L.misa_b200_eam_force_calc(self.h, atoms.ctypes.data, self.cutoff_radius))

    def host_register(self, atoms):
        _ck(self.L.misa_b200_host_register(atoms.ctypes.data, atoms.nbytes))

    def host_unregister(self, atoms):
        _ck(self.L.misa_b200_host_unregister(atoms.ctypes.data))

    # ---- resident mode -----------------------------------------------------------------------
    def upload(self, atoms):
        assert atoms.dtype == ATOM_DTYPE and atoms.size == self.n_ext and atoms.flags["C_CONTIGUOUS"]
        _ck(self.L.misa_b200_upload_atoms(self.h, atoms.ctypes.data))

    def download(self, out=None):
        out = np.zeros(self.n_ext, dtype=ATOM_DTYPE) if out is None else out
        _ck(self.L.misa_b200_download_atoms(self.h, out.ctypes.data))
        return out

    def upload_inter(self, inter):
        inter = np.ascontiguousarray(inter, dtype=ATOM_DTYPE)
        _ck(self.L.misa_b200_upload_inter(self.h, inter.ctypes.data, inter.size))

    def download_inter(self):
        n = C.c_size_t()
        _ck(self.L.misa_b200_download_inter(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=ATOM_DTYPE)
        if n.value:
            _ck(self.L.misa_b200_download_inter(self.h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def set_timestep(self, dt):
        _ck(self.L.misa_b200_set_timestep(self.h, dt))

    def prepare(self):
        _ck(self.L.misa_b200_prepare(self.h))

    def step(self, n=1):
        _ck(self.L.misa_b200_step(self.h, n))

    def step_host(self, atoms, n=1):
        assert atoms.dtype == ATOM_DTYPE and atoms.size == self.n_ext and atoms.flags["C_CONTIGUOUS"]
        _ck(self.L.misa_b200_step_host(self.h, atoms.ctypes.data, n))

    def timed_steps(self, n):
        ms = C.c_double()
        _ck(self.L.misa_b200_timed_steps(self.h, n, C.byref(ms)))
        return ms.value

    def setv(self, lat, direction, energy):
        _ck(self.L.misa_b200_setv(self.h, (C.c_int32 * 4)(*lat), (C.c_double * 3)(*direction), energy))

    def collision_step(self, lat, direction, energy):
        _ck(self.L.misa_b200_collision_step(self.h, (C.c_int32 * 4)(*lat), (C.c_double * 3)(*direction), energy))

    def thermo(self):
        out = (C.c_double * 6)()
        _ck(self.L.misa_b200_thermo(self.h, C.byref(out)))
        return dict(mvv=out[0], pe=out[1], n_atoms=out[2], n_inter=int(out[3]), n_ghost_inter=int(out[4]), runaways=int(out[5]))

    def rescale(self, t_set, t_now):
        _ck(self.L.misa_b200_rescale(self.h, t_set, t_now))

    def sync(self):
        _ck(self.L.misa_b200_sync(self.h))

    # ---- callers / data formats either side of the path ----------------------
    def build_world(self, seed=466953, t_set=600.0, ratio=(1, 0, 0), alloy_seed=1024):
        _ck(self.L.misa_b200_build_world(self.h, seed, t_set, (C.c_int32 * 3)(*ratio), alloy_seed))

    def n_atoms_global(self):
        p = self.dom.phase_space
        return 2 * p[0] * p[1] * p[2]

    def temperature(self):
        out = (C.c_double * 4)()
        _ck(self.L.misa_b200_temperature(self.h, self.n_atoms_global(), C.byref(out)))
        return dict(mvv=out[0], T=out[1], ke=out[2], n_atoms=out[3])

    def rescale_to(self, t_set):
        _ck(self.L.misa_b200_rescale_to(self.h, t_set, self.n_atoms_global()))

    def dump_records(self, time_step, begin=None, end=None, out=None):
        """AtomDump::dump record stream (numpy array of DUMP_DTYPE), compacted on the device."""
        b = C.byref((C.c_int32 * 3)(*begin)) if begin is not None else None
        e = C.byref((C.c_int32 * 3)(*end)) if end is not None else None
        n = C.c_size_t()
        if out is None:
            _ck(self.L.misa_b200_dump_records(self.h, b, e, time_step, None, 0, C.byref(n)))
            out = np.zeros(n.value, dtype=DUMP_DTYPE)
        assert out.dtype == DUMP_DTYPE and out.flags["C_CONTIGUOUS"]
        _ck(self.L.misa_b200_dump_records(self.h, b, e, time_step, out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value]

    def run_pass(self, name):
        _ck(getattr(self.L, "misa_b200_pass_" + name)(self.h))

    # ---- comm / profiling ----------------------------------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        _ck(self.L.misa_b200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid, rank, n_ranks):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        _ck(self.L.misa_b200_comm_init(self.h, buf, rank, n_ranks))

    def profile_enable(self, on=True):
        _ck(self.L.misa_b200_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms = (C.c_double * K_COUNT)()
        n = (C.c_int64 * K_COUNT)()
        _ck(self.L.misa_b200_profile_read(self.h, C.byref(ms), C.byref(n)))
        return {K_NAMES[k]: (ms[k], n[k]) for k in range(K_COUNT)}

    def stencil_stats(self):
        """Per owned site: offsets looped, pair evaluations executed, pairs inside the cutoff (current state)."""
        out = (C.c_double * 4)()
        _ck(self.L.misa_b200_stencil_stats(self.h, C.byref(out)))
        n = max(out[0], 1.0)
        return dict(sites=int(out[0]), offsets_per_atom=out[1] / n, evals_per_atom=out[2] / n, pairs_per_atom=out[3] / n)

    def launch_count(self):
        n = C.c_int64()
        _ck(self.L.misa_b200_launch_count(self.h, C.byref(n)))
        return n.value


# ---- setfl on the host (product side): parse + spline rows, for hosts that do not link libpot ----------
def read_setfl(path):
    """Parse a setfl (eam/alloy) file and build the 7-coefficient spline rows the way libpot's
    eam::interpolateFile() does (LAMMPS array2spline; call site reference src/simulation.cpp:105-131).
    Returns dict(keys, elec, embed, phi) ready for Context.set_potential (file element order)."""
    with open(path) as f:
        for _ in range(3):
            f.readline()
        tok = f.read().split()
    pos = 0
    n_ele = int(tok[pos]); pos += 1 + n_ele
    n_rho, d_rho, n_r, d_r = int(tok[pos]), float(tok[pos + 1]), int(tok[pos + 2]), float(tok[pos + 3]); pos += 5
    keys, embed, elec = [], [], []
    for _ in range(n_ele):
        keys.append(int(tok[pos])); pos += 4
        embed.append(np.array(tok[pos:pos + n_rho], dtype=np.float64)); pos += n_rho
        elec.append(np.array(tok[pos:pos + n_r], dtype=np.float64)); pos += n_r
    phi = [[None] * n_ele for _ in range(n_ele)]
    for i in range(n_ele):
        for j in range(i + 1):
            phi[i][j] = phi[j][i] = np.array(tok[pos:pos + n_r], dtype=np.float64); pos += n_r
    mk = lambda v, dx: (len(v), 1.0 / dx, array2spline(v, dx))
    return dict(keys=keys, elec=[mk(v, d_r) for v in elec], embed=[mk(v, d_rho) for v in embed],
                phi=[[mk(phi[i][j], d_r) for j in range(n_ele)] for i in range(n_ele)])


def array2spline(values, dx):
    n = len(values)
    s = np.zeros((n + 1, 7))
    s[1:, 6] = values
    s[1, 5] = s[2, 6] - s[1, 6]
    s[2, 5] = 0.5 * (s[3, 6] - s[1, 6])
    s[n - 1, 5] = 0.5 * (s[n, 6] - s[n - 2, 6])
    s[n, 5] = s[n, 6] - s[n - 1, 6]
    m = np.arange(3, n - 1)
    s[m, 5] = ((s[m - 2, 6] - s[m + 2, 6]) + 8.0 * (s[m + 1, 6] - s[m - 1, 6])) / 12.0
    m = np.arange(1, n)
    s[m, 4] = 3.0 * (s[m + 1, 6] - s[m, 6]) - 2.0 * s[m, 5] - s[m + 1, 5]
    s[m, 3] = s[m, 5] + s[m + 1, 5] - 2.0 * (s[m + 1, 6] - s[m, 6])
    s[1:, 2] = s[1:, 5] / dx
    s[1:, 1] = 2.0 * s[1:, 4] / dx
    s[1:, 0] = 3.0 * s[1:, 3] / dx
    return s.reshape(-1)


# atom_type enum order (Fe, Cu, Ni) -> atomic numbers, reference src/types/atom_types.h:61-73
TYPE_KEYS = (26, 29, 28)


def potential_in_type_order(pot):
    """Re-order a read_setfl() result from file element order into atom_type enum order."""
    idx = [pot["keys"].index(k) for k in TYPE_KEYS if k in pot["keys"]]
    return ([pot["elec"][i] for i in idx], [pot["embed"][i] for i in idx],
            [[pot["phi"][i][j] for j in idx] for i in idx])
