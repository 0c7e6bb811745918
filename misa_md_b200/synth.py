"""Synthetic initial states (host side, numpy) -- mirror of the reference's WorldBuilder formulas.

reference src/world_builder.cpp:105-131 (positions, velocity draws), :143-199 (vcm / zeroMomentum /
randomAtomsType), src/system_configuration.cpp:45-111 (temperature / rescale), RNG = std::mt19937 scaled by
1/max (src/utils/random/random.cpp:30-37 with MD_RAND=MT, config.cmake:22).

States are generated on the GLOBAL lattice, indexed by global lattice coordinates, and then cut into
sub-boxes, so that the same physical state is fed to every decomposition (the reference's builder draws
per rank and is therefore grid-dependent -- SURVEY.md section 8c).
"""
import numpy as np

BOLTZ = 8.617343e-5          # reference src/types/pre_define.h:11
MVV2E = 1.0364269e-4         # reference src/types/pre_define.h:18
MASS = np.array([55.845, 63.546, 58.6934])  # reference src/types/atom_types.h:17-19
INVALID = -1

# 104-byte AtomElement record, reference src/atom/atom_element.h:18-41
ATOM_DTYPE = np.dtype(
    [("id", "<u8"), ("type", "<i4"), ("_pad", "<i4"), ("x", "<f8", 3), ("v", "<f8", 3), ("f", "<f8", 3),
     ("rho", "<f8"), ("df", "<f8")]
)

# 72-byte dump record atom_dump::AtomInfoDump, reference frontend/io/atom_info_dump.h:14-22
DUMP_DTYPE = np.dtype(
    [("id", "<u8"), ("step", "<u8"), ("type", "<i4"), ("inter_type", "<i2"), ("_pad", "<i2"), ("x", "<f8", 3), ("v", "<f8", 3)]
)


def mt19937_unit(seed, n):
    """n draws of md_rand::random(): std::mt19937(seed)() * (1.0 / 0xFFFFFFFF)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed))
    return bg.random_raw(n).astype(np.float64) * (1.0 / 4294967295.0)


def species_by_id(ids, ratio, alloy_seed):
    """Species of global atom ids (1-based): counter-based hash of (alloy_seed, id) into the cumulative-ratio rule of
    WorldBuilder::randomAtomsType (reference src/world_builder.cpp:180-199, whose unseeded libc rand() has no
    reproducible stream). Host mirror of misa_md_b200/csrc/world.cuh:species_by_id -- same integers, bit for bit."""
    ratio = np.asarray(ratio, dtype=np.int64)
    ids = np.asarray(ids, dtype=np.uint64)
    if np.count_nonzero(ratio) == 1:
        return np.full(ids.shape, int(np.argmax(ratio)), dtype=np.int32)
    with np.errstate(over="ignore"):
        z = np.uint64(alloy_seed) + ids * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    draw = ((z >> np.uint64(33)) % np.uint64(int(ratio.sum()))).astype(np.int64)
    return np.searchsorted(np.cumsum(ratio), draw, side="right").astype(np.int32)


def create_global_state(phase_space, a=2.85532, seed=466953, t_set=600.0, ratio=(1, 0, 0), alloy_seed=1024):
    """Return dict(id,type,x,v) of arrays shaped (PZ, PY, 2*PX[, 3]) for the whole periodic box."""
    px, py, pz = (int(v) for v in phase_space)
    shape = (pz, py, 2 * px)
    n = 2 * px * py * pz
    k, j, i = np.meshgrid(np.arange(pz), np.arange(py), np.arange(2 * px), indexing="ij")
    x = np.empty(shape + (3,), dtype=np.float64)
    # world_builder.cpp:120-124 (operation order kept)
    x[..., 0] = i * 0.5 * a
    x[..., 1] = j * a + (i % 2) * (a / 2)
    x[..., 2] = k * a + (i % 2) * (a / 2)
    ids = (1 + (k * py + j) * (2 * px) + i).astype(np.uint64)
    types = species_by_id(ids, ratio, alloy_seed)
    mass = MASS[types]
    u = mt19937_unit(seed, 3 * n).reshape(shape + (3,))
    v = (u - 0.5) / mass[..., None]
    # WorldBuilder::build: vcm -> /N -> zeroMomentum (world_builder.cpp:75-90,143-178)
    vcm = (v * mass[..., None]).reshape(-1, 3).sum(axis=0) / n
    v -= vcm[None, None, None, :] / mass[..., None]
    if t_set:
        mvv = float(((v ** 2).sum(axis=-1) * mass).sum())
        t_now = mvv * MVV2E / ((3 * n - 3) * BOLTZ)
        v *= np.sqrt(t_set / t_now)
    return dict(id=ids, type=types, x=x, v=v, a=a, phase_space=(px, py, pz))


def sub_box_layout(phase_space, grid, coord, crf=1.96125, ghost=None):
    """Ghost-extended, doubled-x layout of one sub-box (libcomm BccDomain restated on the host)."""
    ghost = int(np.ceil(crf)) + 1 if ghost is None else ghost
    n = [int(phase_space[d]) // int(grid[d]) for d in range(3)]
    for d in range(3):
        if n[d] * grid[d] != phase_space[d]:
            raise ValueError("phase space must divide evenly by the process grid")
    lo = [coord[d] * n[d] for d in range(3)]
    return dict(n=n, ghost=ghost, lo=lo,
                ext_shape=(n[2] + 2 * ghost, n[1] + 2 * ghost, 2 * (n[0] + 2 * ghost)),
                owned=(slice(ghost, ghost + n[2]), slice(ghost, ghost + n[1]), slice(2 * ghost, 2 * ghost + 2 * n[0])))


def scatter_to_sub_box(state, grid, coord, crf=1.96125):
    """Ghost-extended AoS array (flat, ATOM_DTYPE) of one sub-box; ghost sites are INVALID placeholders
    until the first halo exchange fills them (reference src/atom/atom_list.cpp:25-49)."""
    lay = sub_box_layout(state["phase_space"], grid, coord, crf)
    arr = np.zeros(lay["ext_shape"], dtype=ATOM_DTYPE)
    arr["type"] = INVALID
    n, lo = lay["n"], lay["lo"]
    gsl = (slice(lo[2], lo[2] + n[2]), slice(lo[1], lo[1] + n[1]), slice(2 * lo[0], 2 * lo[0] + 2 * n[0]))
    own = arr[lay["owned"]]
    own["id"] = state["id"][gsl]
    own["type"] = state["type"][gsl]
    own["x"] = state["x"][gsl]
    own["v"] = state["v"][gsl]
    arr[lay["owned"]] = own
    return arr.reshape(-1), lay


def perturb_positions(state, sigma, seed=7):
    """Gaussian displacement (sigma in Angstrom) for force-parity tests that need broken symmetry."""
    rs = np.random.RandomState(seed)
    state["x"] = state["x"] + rs.normal(0.0, sigma, size=state["x"].shape)
    return state


def create_sub_box_state(phase_space, grid, coord, a=2.85532, seed=466953, t_set=600.0, ratio=(1, 0, 0), crf=1.96125):
    """Ghost-extended AoS array of ONE sub-box drawn the way the reference's WorldBuilder does it per rank:
    every rank seeds the same generator and draws for its own owned sites (reference src/world_builder.cpp:
    105-131), so the cost is O(sub-box), not O(global box). For identical sub-boxes the
    global zero-momentum / rescale reductions equal the local ones. Used by bench.py for multi-GPU runs."""
    lay = sub_box_layout(phase_space, grid, coord, crf)
    n, lo = lay["n"], lay["lo"]
    local = create_global_state(n, a=a, seed=seed, t_set=t_set, ratio=ratio)
    px, py = int(phase_space[0]), int(phase_space[1])
    k, j, i = np.meshgrid(np.arange(n[2]) + lo[2], np.arange(n[1]) + lo[1], np.arange(2 * n[0]) + 2 * lo[0], indexing="ij")
    local["id"] = (1 + (k * py + j) * (2 * px) + i).astype(np.uint64)
    x = local["x"]
    x[..., 0] = i * 0.5 * a
    x[..., 1] = j * a + (i % 2) * (a / 2)
    x[..., 2] = k * a + (i % 2) * (a / 2)
    arr = np.zeros(lay["ext_shape"], dtype=ATOM_DTYPE)
    arr["type"] = INVALID
    own = arr[lay["owned"]]
    for fld in ("id", "type", "x", "v"):
        own[fld] = local[fld]
    arr[lay["owned"]] = own
    return arr.reshape(-1), lay
