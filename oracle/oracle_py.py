"""ctypes binding of the TEST-ONLY CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

# reference src/atom/atom_element.h:18-41 (104 B)
ATOM_DTYPE = np.dtype(
    [("id", "<u8"), ("type", "<i4"), ("_pad", "<i4"), ("x", "<f8", 3), ("v", "<f8", 3), ("f", "<f8", 3),
     ("rho", "<f8"), ("df", "<f8")]
)
assert ATOM_DTYPE.itemsize == 104


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("pot.c", "md_oracle.c", "pot.h", "md_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class IRegion(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("x_low", "y_low", "z_low", "x_high", "y_high", "z_high")]


class Domain(C.Structure):
    _fields_ = [
        ("phase_space", C.c_long * 3), ("grid_size", C.c_int * 3), ("grid_coord", C.c_int * 3),
        ("rank", C.c_int), ("n_ranks", C.c_int), ("rank_id_neighbours", (C.c_int * 2) * 3),
        ("lattice_const", C.c_double), ("cutoff_radius_factor", C.c_double), ("cut_lattice", C.c_int),
        ("meas_global_length", C.c_double * 3), ("meas_global_low", C.c_double * 3), ("meas_global_high", C.c_double * 3),
        ("meas_sub_box_low", C.c_double * 3), ("meas_sub_box_high", C.c_double * 3),
        ("sub_box_lattice_size", C.c_int * 3), ("lattice_size_ghost", C.c_int * 3),
        ("ghost_extended_lattice_size", C.c_int * 3),
        ("sub_box_lattice_region", IRegion), ("ghost_ext_lattice_region", IRegion),
        ("dbx_sub_box_lattice_size", C.c_int * 3), ("dbx_lattice_size_ghost", C.c_int * 3),
        ("dbx_ghost_extended_lattice_size", C.c_int * 3),
        ("dbx_sub_box_lattice_region", IRegion), ("dbx_ghost_ext_lattice_region", IRegion),
    ]


class IVec(C.Structure):
    _fields_ = [("v", C.POINTER(C.c_long)), ("n", C.c_size_t), ("cap", C.c_size_t)]

    def to_numpy(self):
        if self.n == 0:
            return np.zeros(0, dtype=np.int64)
        return np.ctypeslib.as_array(self.v, shape=(self.n,)).astype(np.int64)


class Rank(C.Structure):
    _fields_ = [
        ("dom", Domain),
        ("size_x", C.c_long), ("size_y", C.c_long), ("size_z", C.c_long), ("size", C.c_long),
        ("atoms", C.c_void_p),
        ("sendlist", IVec * 6), ("recvlist", IVec * 6),
        ("nei_even", IVec), ("nei_odd", IVec), ("nei_half_even", IVec), ("nei_half_odd", IVec),
        ("inter", C.c_void_p), ("n_inter", C.c_size_t), ("cap_inter", C.c_size_t),
        ("ghost", C.c_void_p), ("n_ghost", C.c_size_t), ("cap_ghost", C.c_size_t),
        ("intersend", IVec * 6), ("interrecv", IVec * 6),
        ("map_site", C.c_void_p), ("map_ref", C.c_void_p), ("n_map", C.c_size_t), ("cap_map", C.c_size_t),
        ("cutoff_radius", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    vp, d, i, l = C.c_void_p, C.c_double, C.c_int, C.c_long
    L.pot_write_synthetic_setfl.argtypes = [C.c_char_p, i, d, i, d, d]
    L.pot_write_synthetic_setfl.restype = i
    L.pot_read_setfl.argtypes = [C.c_char_p]
    L.pot_read_setfl.restype = vp
    L.pot_free.argtypes = [vp]
    for fn, args in (("pot_charge_density", [vp, i, d]), ("pot_d_embed_energy", [vp, i, d]),
                     ("pot_to_force", [vp, i, i, d, d, d]), ("pot_embed_energy", [vp, i, d]),
                     ("pot_pair_energy", [vp, i, i, d])):
        getattr(L, fn).argtypes = args
        getattr(L, fn).restype = d
    L.ora_domain_build.argtypes = [C.POINTER(Domain), C.POINTER(l * 3), C.POINTER(i * 3), C.POINTER(i * 3), d, d, i]
    L.ora_domain_build.restype = i
    L.ora_world_create.argtypes = [C.POINTER(l * 3), C.POINTER(i * 3), d, d, vp, d, i]
    L.ora_world_create.restype = vp
    L.ora_world_free.argtypes = [vp]
    L.ora_world_rank.argtypes = [vp, i]
    L.ora_world_rank.restype = C.POINTER(Rank)
    for fn in ("ora_world_fill_lattice", "ora_exchange_atom_first", "ora_exchange_atom", "ora_exchange_inter",
               "ora_border_inter", "ora_clear_force", "ora_compute_eam", "ora_first_step", "ora_second_step",
               "ora_prepare", "ora_step"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = None
    L.ora_decide.argtypes = [vp]
    L.ora_decide.restype = i
    L.ora_set_dt.argtypes = [vp, d]
    L.ora_setv.argtypes = [vp, C.POINTER(i * 4), C.POINTER(d * 3), d]
    L.ora_collision_step.argtypes = [vp, C.POINTER(i * 4), C.POINTER(d * 3), d]
    for fn in ("ora_mvv", "ora_kinetic_energy", "ora_temperature", "ora_potential_energy"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = d
    L.ora_rescale.argtypes = [vp, d]
    L.ora_total_inter.argtypes = [vp]
    L.ora_total_inter.restype = C.c_size_t
    for fn in ("ora_lat_rho", "ora_lat_df", "ora_lat_force"):
        getattr(L, fn).argtypes = [C.POINTER(Rank), vp]
        getattr(L, fn).restype = None
    L.ora_nei_make.argtypes = [C.POINTER(Rank), i, d]
    L.ora_is_positive_index.argtypes = [d, d, d]
    L.ora_is_positive_index.restype = i
    L.ora_voronoy.argtypes = [d, d, d, d, C.POINTER(l * 3)]
    L.ora_is_out_box.argtypes = [vp, C.POINTER(Domain)]
    L.ora_is_out_box.restype = C.c_uint
    L.ora_near_lat_coord.argtypes = [vp, C.POINTER(Domain), C.POINTER(l * 3)]
    L.ora_near_lat_sub_box_coord.argtypes = [vp, C.POINTER(Domain), C.POINTER(l * 3)]
    L.ora_fw_comm_local_region.argtypes = [C.POINTER(Domain), i, i]
    L.ora_fw_comm_local_region.restype = IRegion
    _lib = L
    return L


DEFAULT_SETFL = dict(n_rho=5000, d_rho=0.02, n_r=5001, d_r=0.00112, cutoff=5.6)


SETFL_PATH = os.path.join(os.path.dirname(HERE), "misa_md_b200", "data", "FeCuNi.synthetic.eam.alloy")


def synthetic_setfl_path():
    """The committed synthetic FeCuNi setfl file (generated once by pot_write_synthetic_setfl with
    DEFAULT_SETFL; it is input DATA shared by the oracle and the product, not code)."""
    return SETFL_PATH


class Pot:
    def __init__(self, path=None):
        self.path = path or synthetic_setfl_path()
        self.h = lib().pot_read_setfl(self.path.encode())
        if not self.h:
            raise RuntimeError("cannot read setfl " + self.path)

    def charge_density(self, key, r2):
        return lib().pot_charge_density(self.h, key, r2)

    def d_embed_energy(self, key, rho):
        return lib().pot_d_embed_energy(self.h, key, rho)

    def to_force(self, k1, k2, r2, df1, df2):
        return lib().pot_to_force(self.h, k1, k2, r2, df1, df2)


def make_domain(phase_space, grid_size, grid_coord, a, crf, ghost=-1):
    d = Domain()
    rc = lib().ora_domain_build(C.byref(d), (C.c_long * 3)(*phase_space), (C.c_int * 3)(*grid_size),
                                (C.c_int * 3)(*grid_coord), a, crf, ghost)
    if rc != 0:
        raise ValueError("phase space not divisible by grid")
    return d


class World:
    """In-process multi-sub-box oracle world (one 'rank' per sub-box)."""

    def __init__(self, phase_space, grid=(1, 1, 1), a=2.85532, crf=1.96125, pot=None, dt=0.001, threads=1):
        self.pot = pot or Pot()
        self.L = lib()
        self.h = self.L.ora_world_create((C.c_long * 3)(*phase_space), (C.c_int * 3)(*grid), a, crf, self.pot.h, dt, threads)
        if not self.h:
            raise ValueError("bad world")
        self.n_ranks = grid[0] * grid[1] * grid[2]
        self.phase_space = tuple(phase_space)
        self.grid = tuple(grid)

    def close(self):
        if self.h:
            self.L.ora_world_free(self.h)
            self.h = None

    def rank(self, r):
        return self.L.ora_world_rank(self.h, r).contents

    def atoms(self, r):
        """numpy structured view (no copy) of rank r's ghost-extended AoS array."""
        rk = self.rank(r)
        buf = (C.c_char * (rk.size * 104)).from_address(rk.atoms)
        return np.frombuffer(buf, dtype=ATOM_DTYPE)

    def inter(self, r):
        rk = self.rank(r)
        if rk.n_inter == 0:
            return np.zeros(0, dtype=ATOM_DTYPE)
        buf = (C.c_char * (rk.n_inter * 104)).from_address(rk.inter)
        return np.frombuffer(buf, dtype=ATOM_DTYPE).copy()

    def shape(self, r):
        rk = self.rank(r)
        return (rk.size_z, rk.size_y, rk.size_x)

    def owned_slices(self, r):
        d = self.rank(r).dom
        g, b = d.dbx_lattice_size_ghost, d.dbx_sub_box_lattice_size
        return (slice(g[2], g[2] + b[2]), slice(g[1], g[1] + b[1]), slice(g[0], g[0] + b[0]))

    def fill_lattice(self):
        self.L.ora_world_fill_lattice(self.h)

    def prepare(self):
        self.L.ora_prepare(self.h)

    def step(self):
        self.L.ora_step(self.h)

    def kinetic_energy(self):
        return self.L.ora_kinetic_energy(self.h)

    def potential_energy(self):
        return self.L.ora_potential_energy(self.h)

    def temperature(self):
        return self.L.ora_temperature(self.h)

    def total_inter(self):
        return self.L.ora_total_inter(self.h)

    def dump(self, r, time_step):
        """AtomDump::dump record stream of rank r (numpy array of DUMP_DTYPE)."""
        from misa_md_b200.synth import DUMP_DTYPE
        self.L.ora_dump.restype = C.c_size_t
        self.L.ora_dump.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        rk = self.L.ora_world_rank(self.h, r)
        n = self.L.ora_dump(rk, time_step, None, 0)
        out = np.zeros(n, dtype=DUMP_DTYPE)
        self.L.ora_dump(rk, time_step, out.ctypes.data, n)
        return out
