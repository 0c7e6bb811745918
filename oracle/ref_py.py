"""ctypes binding of oracle/_ref/libmisa_ref.so: the reference's OWN hot-path sources compiled in place from
/root/reference/src against the shim headers in oracle/shim (recipe: oracle/Makefile target `ref`).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it. The .so is built in the authoring container (where /root/reference exists) and
travels to the GPU box as a prebuilt file; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle_py import ATOM_DTYPE, synthetic_setfl_path

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libmisa_ref.so")
HOOK_LIB_PATH = os.path.join(HERE, "_ref", "libmisa_ref_cuda.so")


def available(hooks=False):
    return os.path.exists(HOOK_LIB_PATH if hooks else LIB_PATH)


def build():
    """(Re)build from the reference sources when they are present; otherwise keep the prebuilt library."""
    if os.path.isdir("/root/reference/src"):
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL, env=env)
    return available()


_libs = {}


def lib(hooks=False):
    if hooks in _libs:
        return _libs[hooks]
    path = HOOK_LIB_PATH if hooks else LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(path + " is missing (built by `make -C oracle ref` where /root/reference exists)")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL if hooks else C.RTLD_LOCAL)
    vp, d, i, l = C.c_void_p, C.c_double, C.c_int, C.c_long
    L.ref_world_create.argtypes = [C.POINTER(l * 3), C.POINTER(i * 3), d, d, C.c_char_p, d]
    L.ref_world_create.restype = vp
    L.ref_world_free.argtypes = [vp]
    L.ref_accelerated.restype = i
    L.ref_atoms.argtypes = [vp, i]
    L.ref_atoms.restype = vp
    L.ref_size.argtypes = [vp, i]
    L.ref_size.restype = l
    L.ref_layout.argtypes = [vp, i, C.POINTER(i * 3), C.POINTER(i * 3), C.POINTER(i * 3), C.POINTER(i * 3)]
    L.ref_build_world.argtypes = [vp, i, d, C.POINTER(i * 3)]
    L.ref_set_dt.argtypes = [vp, d]
    for fn in ("ref_prepare", "ref_exchange_atom_first", "ref_clear_force", "ref_compute_eam", "ref_first_step", "ref_second_step"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = None
    L.ref_step.argtypes = [vp, i]
    L.ref_decide.argtypes = [vp]
    L.ref_decide.restype = i
    L.ref_collision_step.argtypes = [vp, C.POINTER(i * 4), C.POINTER(d * 3), d]
    L.ref_setv.argtypes = [vp, C.POINTER(i * 4), C.POINTER(d * 3), d]
    L.ref_list_len.argtypes = [vp, i, i, i]
    L.ref_list_len.restype = C.c_size_t
    L.ref_list_get.argtypes = [vp, i, i, i, C.POINTER(l)]
    L.ref_nei_len.argtypes = [vp, i, i]
    L.ref_nei_len.restype = C.c_size_t
    L.ref_nei_get.argtypes = [vp, i, i, C.POINTER(l)]
    for fn in ("ref_n_inter", "ref_n_ghost_inter"):
        getattr(L, fn).argtypes = [vp, i]
        getattr(L, fn).restype = C.c_size_t
    L.ref_get_inter.argtypes = [vp, i, vp]
    L.ref_set_inter.argtypes = [vp, i, vp, C.c_size_t]
    for fn in ("ref_mvv", "ref_temperature"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = d
    L.ref_rescale.argtypes = [vp, d]
    L.ref_is_out_box.argtypes = [vp, i, C.POINTER(d * 3)]
    L.ref_is_out_box.restype = C.c_uint
    L.ref_dump.argtypes = [vp, i, C.c_size_t, vp, C.c_size_t]
    L.ref_dump.restype = C.c_size_t
    L.ref_near_lat_sub_box_coord.argtypes = [vp, i, C.POINTER(d * 3), C.POINTER(l * 3)]
    _libs[hooks] = L
    return L


class World:
    """Same surface as oracle_py.World, driven by the reference's own classes. hooks=True loads the build in
    which atom::latRho/latDf/latForce call the cuda_* hooks of arch_cuda/cuda_hooks.cpp (needs a GPU)."""

    def __init__(self, phase_space, grid=(1, 1, 1), a=2.85532, crf=1.96125, pot=None, dt=0.001, threads=None,
                 setfl=None, hooks=False):
        self.L = lib(hooks)
        self.h = self.L.ref_world_create((C.c_long * 3)(*phase_space), (C.c_int * 3)(*grid), a, crf,
                                         (setfl or synthetic_setfl_path()).encode(), dt)
        if not self.h:
            raise ValueError("bad world / setfl")
        self.n_ranks = grid[0] * grid[1] * grid[2]
        self.phase_space, self.grid = tuple(phase_space), tuple(grid)

    def close(self):
        if self.h:
            self.L.ref_world_free(self.h)
            self.h = None

    def _layout(self, r):
        e, b, g, c = (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)()
        self.L.ref_layout(self.h, r, C.byref(e), C.byref(b), C.byref(g), C.byref(c))
        return list(e), list(b), list(g), tuple(c)

    def coord(self, r):
        return self._layout(r)[3]

    def shape(self, r):
        e = self._layout(r)[0]
        return (e[2], e[1], e[0])

    def owned_slices(self, r):
        _, b, g, _ = self._layout(r)
        return (slice(g[2], g[2] + b[2]), slice(g[1], g[1] + b[1]), slice(g[0], g[0] + b[0]))

    def atoms(self, r):
        n = self.L.ref_size(self.h, r)
        buf = (C.c_char * (n * 104)).from_address(self.L.ref_atoms(self.h, r))
        return np.frombuffer(buf, dtype=ATOM_DTYPE)

    def inter(self, r):
        n = self.L.ref_n_inter(self.h, r)
        out = np.zeros(n, dtype=ATOM_DTYPE)
        if n:
            self.L.ref_get_inter(self.h, r, out.ctypes.data)
        return out

    def set_inter(self, r, atoms):
        atoms = np.ascontiguousarray(atoms, dtype=ATOM_DTYPE)
        self.L.ref_set_inter(self.h, r, atoms.ctypes.data, atoms.size)

    def total_inter(self):
        return sum(self.L.ref_n_inter(self.h, r) for r in range(self.n_ranks))

    def sendlist(self, r, index, recv=False):
        n = self.L.ref_list_len(self.h, r, int(recv), index)
        out = (C.c_long * max(n, 1))()
        if n:
            self.L.ref_list_get(self.h, r, int(recv), index, out)
        return np.array(out[:n], dtype=np.int64)

    def offsets(self, r, which):
        n = self.L.ref_nei_len(self.h, r, which)
        out = (C.c_long * max(n, 1))()
        self.L.ref_nei_get(self.h, r, which, out)
        return np.array(out[:n], dtype=np.int64)

    def build_world(self, seed=466953, t_set=600.0, ratio=(1, 0, 0)):
        self.L.ref_build_world(self.h, seed, t_set, (C.c_int * 3)(*ratio))

    def prepare(self):
        self.L.ref_prepare(self.h)

    def step(self, n=1):
        self.L.ref_step(self.h, n)

    def collision_step(self, lat, direction, energy):
        self.L.ref_collision_step(self.h, (C.c_int * 4)(*lat), (C.c_double * 3)(*direction), energy)

    def set_dt(self, dt):
        self.L.ref_set_dt(self.h, dt)

    def mvv(self):
        return self.L.ref_mvv(self.h)

    def temperature(self):
        return self.L.ref_temperature(self.h)

    def rescale(self, t):
        self.L.ref_rescale(self.h, t)

    def dump(self, r, time_step):
        """Byte stream AtomDump::dump hands to the file writer for rank r, as DUMP_DTYPE records. The two padding
        bytes of each record are uninitialised heap in the reference (new AtomInfoDump[]): zeroed here."""
        from misa_md_b200.synth import DUMP_DTYPE
        nbytes = self.L.ref_dump(self.h, r, time_step, None, 0)
        assert nbytes % DUMP_DTYPE.itemsize == 0
        out = np.zeros(nbytes // DUMP_DTYPE.itemsize, dtype=DUMP_DTYPE)
        self.L.ref_dump(self.h, r, time_step, out.ctypes.data, nbytes)
        out["_pad"] = 0
        return out
