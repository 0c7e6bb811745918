#!/usr/bin/env python
"""bench.py -- atom-steps/s of the EAM hot path (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, resident mode), N=1 default
    torchrun ... bench.py --gpus N --steps K --warmup W      # N>1: one rank per GPU, weak scaling
    python bench.py --impl reference --steps K --warmup W    # the CPU arm (reference algorithm on host cores)

A "step" is one iteration of simulation::simulate (reference src/simulation.cpp:164-194): firststep, decide,
ghost exchange, rho, df, df halo, force, secondstep. Workload at N=1: BASELINE.json configs[1], bcc Fe 100^3
cells (2 M atoms), synthetic FeCuNi setfl table; at N>1 the same 100^3 cells PER GPU (configs[2] geometry at
N=8: 200^3 cells on a 2x2x2 grid), i.e. weak scaling. See DESIGN.md section 7 for every field of the line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "atom-steps/sec (EAM Fe-Cu-Ni, fp64)"
UNIT = "atom-steps/s"
A, CRF, DT = 2.85532, 1.96125, 0.001
GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
# algorithmic bytes per owned atom per launch
BYTES_PER_ATOM = {"rho": 36 + 8, "df": 20, "force": 60, "verlet1": 124, "verlet2": 76}
STEP_BYTES_PER_ATOM = 316
# fp64 view (SURVEY.md section 8d "algorithmic flops"): fp64 warp instructions per EVALUATED pair and per distance test
# that is rejected, counted in the SASS of the production kernels (profiles/r02_sass_near_loop.txt; DFMA/DMUL/DADD/DSETP/
# MUFU.RSQ64H each one issue slot of the fp64 pipe); peak = measured DFMA lane rate (profiles/r01_fp64_peak.json)
FP64_INST = {"force": {"eval": 50, "test": 7}, "rho": {"eval": 22, "test": 7}}
# the stencil kernels whose ncu capture profiles/traffic.json describes; bench.py emits traffic: null when these sources changed
TRAFFIC_SOURCES = ["misa_md_b200/csrc/eam_fast.cuh", "misa_md_b200/csrc/eam_smem.cuh"]


def workload_name(cells, ratio, atoms_per_gpu):
    species = "Fe" if list(ratio)[1:] == [0, 0] else "Fe-Cu-Ni %d:%d:%d" % tuple(ratio)
    return "bcc %s %d^3 cells (%d atoms) per GPU, NVE, dt 1 fs, T0 600 K, synthetic FeCuNi setfl" % (species, cells, atoms_per_gpu)


def config_block(cells, ratio, n_gpus, equil):
    """The `config` object of the JSON line -- identical for both arms (the reference arm runs a bounded sample of it)."""
    atoms = 2 * cells ** 3
    ext = 2 * (cells + 6) ** 3
    return {"workload": workload_name(cells, ratio, atoms), "cells_per_gpu": [cells] * 3, "grid": list(GRIDS[n_gpus]),
            "atoms_total": n_gpus * atoms, "species_ratio": list(ratio), "equil_steps": equil,
            "l2": "resident state %.0f MB per GPU exceeds the 126 MB L2; no flush between steps" % (ext * 105 / 1e6)}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout. Native libraries write there too (NCCL prints its version banner on
# fd 1 when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the line goes to the saved fd.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in-process every 10 ms when pynvml is
    importable (nvidia_ml_py), else `nvidia-smi -lms 50` as a child process. Started well before the timed region."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "power.draw", "clocks_event_reasons.hw_slowdown",
              "clocks_event_reasons.hw_thermal_slowdown", "clocks_event_reasons.sw_thermal_slowdown",
              "clocks_event_reasons.sw_power_cap"]
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.rows = []       # (t, sm_mhz, sm_max_mhz, set(reasons))
        self.proc = None
        self.stop = False
        self.source = None
        self.t0 = self.t1 = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + ",".join(self.FIELDS), "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.th = threading.Thread(target=self._read_smi, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), sm, self.mx, {n for b, n in self.NVML_REASONS if mask & b}))
            except Exception:
                pass
            time.sleep(0.01)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) != len(self.FIELDS):
                continue
            try:
                sm, mx = float(r[0]), float(r[1])
            except ValueError:
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.rows.append((time.perf_counter(), sm, mx, {n for n, v in zip(names, r[3:]) if v.lower().startswith("active")}))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def finish(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= self.t1 + 0.02]
        where = "timed region"
        if not rows:  # timed region shorter than one sample: take whatever was seen closest to it
            rows, where = self.rows[-3:], "nearest samples"
        reasons = set()
        for r in rows:
            reasons |= r[3]
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source, "window": where}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the box's host cores (oracle port, or oracle/_ref when it was built)
# ---------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_grid(cores):
    best = (1, 1, 1)
    for gx in (1, 2, 4):
        for gy in (1, 2, 4):
            for gz in (1, 2, 4):
                n = gx * gy * gz
                if n <= cores and (n > best[0] * best[1] * best[2] or
                                   (n == best[0] * best[1] * best[2] and max(gx, gy, gz) < max(best))):
                    best = (gx, gy, gz)
    return best


class CpuArm:
    """One in-process world of sub-boxes stepped by OpenMP threads -- the reference's one-MPI-rank-per-sub-box
    model without MPI ("MPI-equivalent in-process exchange")."""

    def __init__(self, cells, cores):
        from oracle import cpu_arm  # the ONLY place bench.py touches oracle/: the reported CPU baseline
        self.impl = cpu_arm.make(cells, cpu_grid(cores), A, CRF, DT)
        self.kind = self.impl.kind
        self.cells = cells
        self.grid = self.impl.grid
        self.cores = self.impl.threads
        self.atoms = 2 * cells ** 3

    def step(self):
        self.impl.step()

    def close(self):
        self.impl.close()

    def sample(self, n_steps):
        return "bcc Fe %d^3 cells (%d atoms), %dx%dx%d in-process sub-boxes on %d threads, %d steps" % (
            self.cells, self.atoms, self.grid[0], self.grid[1], self.grid[2], self.cores, n_steps)


def pick_cpu_cells(cores, steps_total, budget_s):
    """Largest sample of the workload whose (steps_total) steps fit the budget at ~2.5e5 atom-steps/s/core."""
    for cells in (100, 80, 64, 48, 40, 32, 24, 16):
        if 2 * cells ** 3 * steps_total / (2.5e5 * max(cores, 1)) <= budget_s:
            return cells
    return 16


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    cells = pick_cpu_cells(cores, args.steps + args.warmup, 150.0)
    arm = CpuArm(cells, cores)
    for _ in range(args.warmup):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    dt = time.perf_counter() - t0
    value = arm.atoms * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args.cells, args.ratio, args.gpus, args.equil),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample(args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    arm.close()
    emit(line)


def cpu_baseline(budget_s=20.0):
    cores = host_cores()
    cells = pick_cpu_cells(cores, 4, budget_s)
    arm = CpuArm(cells, cores)
    arm.step()  # warm
    t0 = time.perf_counter()
    n = 0
    while True:
        arm.step()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s * 0.5 or n >= 10:
            break
    out = {"value": arm.atoms * n / el, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample(n)}
    arm.close()
    return out


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def source_hash():
    import hashlib
    h = hashlib.sha256()
    for rel in TRAFFIC_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def committed_traffic(kernel):
    """DRAM bytes per launch + pipe utilisation of `kernel` from the committed `ncu --set full` capture -- only when the
    capture was taken on THESE kernel sources (profiles/traffic.json carries their hash); otherwise null, not a stale number."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, None, "no capture committed"
    with open(tp) as f:
        tj = json.load(f)
    if tj.get("source_sha256_16") != source_hash():
        return None, None, "capture is of other kernel sources (%s != %s)" % (tj.get("source_sha256_16"), source_hash())
    return tj.get(kernel), tj.get(kernel + "_pipes"), tj.get("_source")


class Env:
    """torch.distributed plumbing of one rank."""

    def __init__(self, n_gpus):
        import torch
        import torch.distributed as dist
        import misa_md_b200 as mb
        self.torch, self.dist, self.mb = torch, dist, mb
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != n_gpus:
            if self.world == 1 and n_gpus > 1:
                raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (n_gpus, n_gpus))
            n_gpus = self.world
        self.n_gpus = n_gpus
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.grid = GRIDS[n_gpus]
        if os.environ.get("BENCH_GRID"):   # experiments only (e.g. "1,1,2": which faces cross NVLink); the contract's grids are GRIDS
            self.grid = tuple(int(v) for v in os.environ["BENCH_GRID"].split(","))
            assert self.grid[0] * self.grid[1] * self.grid[2] == n_gpus
        g = self.grid
        self.coord = (self.rank // (g[1] * g[2]), (self.rank // g[2]) % g[1], self.rank % g[2])
        self.lib = mb.load()
        mb.capi._ck(self.lib.misa_b200_env_init(self.local_rank))
        self.pot = mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH))

    def context(self, cells_per_gpu, dt=DT):
        """One sub-box of the bench's own process grid on this rank's GPU, communicator and peer mapping included."""
        phase = tuple(c * g for c, g in zip(cells_per_gpu, self.grid))
        ctx = self.mb.Context(phase, grid=self.grid, coord=self.coord, a=A, crf=CRF)
        ctx.make_offsets()
        ctx.set_potential(*self.pot)
        ctx.set_timestep(dt)
        if self.world > 1:
            uid = self.torch.zeros(128, dtype=self.torch.uint8, device="cuda")
            if self.rank == 0:
                uid.copy_(self.torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=self.torch.uint8))
            self.dist.broadcast(uid, 0)
            ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), self.rank, self.world)
        return ctx

    def barrier(self, ctx=None):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        if ctx is not None:
            ctx.sync()

    def reduce(self, x, op="max"):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())


# ---- parity self-check: the CUDA path against the CPU oracle on the bench's OWN process grid, before anything is timed ----
def _per_atom_rel(got, ref, floor=1e-6):
    """max over atoms of |got_i - ref_i|_inf / max(|ref_i|_2, floor * max_j |ref_j|_2): north_star's "1e-10 relative per atom".
    rho and f are judged against the atom's own magnitude (floor 1e-6 of the largest only guards exact zeros); df = F'(rho)
    changes sign near the equilibrium density, where an atom's own |df| says nothing about the accuracy of F' -- floor 1e-3."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.ndim == 1:
        got, ref = got[:, None], ref[:, None]
    if ref.size == 0:
        return 0.0
    mag = np.sqrt((ref * ref).sum(axis=1))
    den = np.maximum(mag, floor * float(mag.max()) if mag.max() > 0 else 1.0)
    return float((np.abs(got - ref).max(axis=1) / den).max())


def parity_check(env, args):
    """simulation::prepareForStart + 5 x simulate's loop body (reference src/simulation.cpp:137-145,164-194) on a small box
    cut over the SAME process grid as the timed run, CUDA path (ghost push / NCCL exchange included) against the oracle's
    in-process multi-sub-box world. Case "timed_path": the timed run's species ratio, no vacancy, sub-boxes long enough in x
    for the stencil kernels' interior-first order; case "alloy_vacancies": Fe-Cu-Ni 97:2:1 with 5 vacant sites. Occupancy and
    ids exact, rho/df/f at step 0 <= 1e-10 per atom, x/v after 5 steps <= 1e-12, f <= 1e-9 (fp64 round-off, amplified by the
    dynamics). The oracle is the checker here, nothing it computes is timed."""
    from misa_md_b200 import synth
    from oracle import oracle_py as O   # checker only
    cases = [("timed_path", (40, 12, 12), tuple(args.ratio), 0), ("alloy_vacancies", (12, 12, 12), (97, 2, 1), 5)]
    out = {"n_ranks": env.world, "grid": list(env.grid), "steps": 5, "cases": {}, "oracle": "oracle/liboracle.so (CPU restatement, bit-identical to the reference sources: tests/test_oracle_vs_ref.py)"}
    all_ok = True
    for name, cells, ratio, vac in cases:
        phase = tuple(c * g for c, g in zip(cells, env.grid))
        st = synth.create_global_state(phase, a=A, seed=466953, t_set=600.0, ratio=ratio)
        synth.perturb_positions(st, 0.03)
        if vac:
            rs = np.random.RandomState(99)
            flat = st["type"].reshape(-1)
            flat[rs.choice(flat.size, vac, replace=False)] = synth.INVALID
        ctx = env.context(cells)
        arr, _ = synth.scatter_to_sub_box(st, env.grid, env.coord, CRF)
        ctx.upload(arr)
        w = O.World(phase, grid=env.grid, a=A, crf=CRF, dt=DT, threads=1)
        for r in range(w.n_ranks):
            sub, _ = synth.scatter_to_sub_box(st, env.grid, tuple(w.rank(r).dom.grid_coord), CRF)
            w.atoms(r)[:] = sub
        ctx.prepare()
        w.prepare()
        sl = ctx.owned
        me = [r for r in range(w.n_ranks) if tuple(w.rank(r).dom.grid_coord) == tuple(env.coord)][0]
        assert tuple(w.shape(me)) == tuple(ctx.ext_shape)
        got0 = ctx.download().reshape(ctx.ext_shape)[sl].copy()
        ref0 = w.atoms(me).reshape(ctx.ext_shape)[sl].copy()
        ctx.step(5)                      # ONE call: the sync-free pipelined step with the fused half-kicks
        for _ in range(5):
            w.step()
        got5 = ctx.download().reshape(ctx.ext_shape)[sl]
        ref5 = w.atoms(me).reshape(ctx.ext_shape)[sl]
        v0, v5 = ref0["type"] >= 0, ref5["type"] >= 0
        res = {
            "occupancy_exact": bool(np.array_equal(got0["type"], ref0["type"]) and np.array_equal(got5["type"], ref5["type"])),
            "ids_exact": bool(np.array_equal(got5["id"][v5], ref5["id"][v5])),
            "max_rel_rho_step0": _per_atom_rel(got0["rho"][v0], ref0["rho"][v0]),
            "max_rel_df_step0": _per_atom_rel(got0["df"][v0], ref0["df"][v0], 1e-3),
            "max_rel_f_step0": _per_atom_rel(got0["f"][v0], ref0["f"][v0]),
            "max_rel_x_step5": _per_atom_rel(got5["x"][v5], ref5["x"][v5]),
            "max_rel_v_step5": _per_atom_rel(got5["v"][v5], ref5["v"][v5]),
            "max_rel_f_step5": _per_atom_rel(got5["f"][v5], ref5["f"][v5]),
            "runaways": int(ctx.thermo()["runaways"]), "pipelined_steps": int(ctx.query("pipe_steps")),
            "exchange": "fill" if env.world == 1 else ("push" if ctx.query("p2p") else "nccl"),
        }
        ok = (res["occupancy_exact"] and res["ids_exact"] and res["max_rel_rho_step0"] <= 1e-10 and res["max_rel_df_step0"] <= 1e-10 and
              res["max_rel_f_step0"] <= 1e-10 and res["max_rel_x_step5"] <= 1e-12 and res["max_rel_v_step5"] <= 1e-9 and res["max_rel_f_step5"] <= 1e-9)
        # every rank checked its own sub-box: fold (worst value, all exact) over the ranks
        for k, v in list(res.items()):
            if isinstance(v, bool):
                res[k] = bool(env.reduce(1.0 if v else 0.0, "min") > 0.5)
            elif isinstance(v, float):
                res[k] = env.reduce(v, "max")
        ok = env.reduce(1.0 if ok else 0.0, "min") > 0.5
        res["ok"] = bool(ok)
        res["atoms"] = int(2 * phase[0] * phase[1] * phase[2])
        res["cells_per_gpu"] = list(cells)
        res["species_ratio"] = list(ratio)
        res["vacancies"] = vac
        out["cases"][name] = res
        all_ok = all_ok and ok
        ctx.close()
        w.close()
        env.barrier()
    out["ok"] = bool(all_ok)
    out["max_rel_f"] = max(c["max_rel_f_step0"] for c in out["cases"].values())
    out["occupancy_exact"] = all(c["occupancy_exact"] for c in out["cases"].values())
    return out


def kernel_table(prof, atoms_per_gpu, peak):
    kernels = {}
    for name, (tot_ms, cnt) in prof.items():
        if cnt and name in BYTES_PER_ATOM:
            t = tot_ms / cnt * 1e-3
            gbs = BYTES_PER_ATOM[name] * atoms_per_gpu / t / 1e9
            kernels[name] = {"ms": t * 1e3, "gbs": gbs, "frac": gbs / peak}
        elif cnt:
            kernels[name] = {"ms": tot_ms / cnt}
    return kernels


def timed_config(env, cells, ratio, args, clocks=None):
    """Build the world on the device, thermalise, time `args.steps` resident steps (CUDA events on the library's stream, max
    over ranks) and the per-kernel slots in a separate pass. Returns (ctx, result dict); the caller closes ctx."""
    ctx = env.context((cells,) * 3)
    t_build = time.perf_counter()
    # initial state by global atom id (misa_b200_build_world = WorldBuilder::build, seed 466953, 600 K): every sub-box cuts
    # its part out of the same global state, no host init, no H2D
    ctx.build_world(seed=466953, t_set=600.0, ratio=tuple(ratio))
    t_build = time.perf_counter() - t_build
    ctx.prepare()
    # Untimed thermalisation: the state starts as a PERFECT lattice with 600 K of kinetic energy; until the positions have
    # thermalised (~100 steps) fewer pairs fall inside the cutoff and the pruned stencil is shorter, which would flatter the
    # timed region. The timed steps are those of the stationary NVE run BASELINE.json names.
    ctx.step(args.equil)
    ctx.step(max(args.warmup, 3))
    env.barrier(ctx)
    l0 = ctx.launch_count()
    if clocks:
        clocks.mark_start()
    ms = ctx.timed_steps(args.steps)
    env.barrier(ctx)
    if clocks:
        clocks.mark_stop()
    launches = ctx.launch_count() - l0
    ms = env.reduce(ms, "max")
    atoms = ctx.n_owned
    peak, _ = peaks()
    ctx.profile_enable(True)
    ctx.step(args.steps)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    stats = ctx.stencil_stats()
    res = {
        "workload": workload_name(cells, ratio, atoms),
        "cells_per_gpu": [cells] * 3, "species_ratio": list(ratio), "atoms_total": env.n_gpus * atoms,
        "value": env.n_gpus * atoms * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps, "steps": args.steps,
        "kernels": kernel_table(prof, atoms, peak), "gpu_launches": int(launches),
        "step_hbm_frac": STEP_BYTES_PER_ATOM * atoms * args.steps / (ms * 1e-3) / 1e9 / peak,
        "stencil": stats, "world_build_ms": 1e3 * t_build,
    }
    return ctx, res


def pka_config(env, cells, args):
    """BASELINE.json configs[3]: PKA collision cascade in bcc Fe (reference stage machine of example/config.yaml:58-65 --
    thermalise at dt 1 fs, rescale to 300 K, setv on one atom, cascade at dt 0.1 fs; frontend/md_simulation.cpp:40-80), all in
    resident mode. Timed: the last `steps` of the cascade, when off-lattice atoms and vacancies are around (serial path, inter
    kernels, low-list recompute). One sub-box only."""
    from misa_md_b200 import synth
    ctx = env.context((cells,) * 3)
    ctx.build_world(seed=466953, t_set=600.0, ratio=(1, 0, 0))
    ctx.set_timestep(DT)
    ctx.prepare()
    ctx.step(args.equil)
    ctx.rescale_to(300.0)
    ctx.step(50)

    def energy():
        th = ctx.thermo()
        return 0.5 * th["mvv"] * synth.MVV2E + th["pe"], th
    ctx.set_timestep(1e-4)
    ctx.collision_step((cells // 2, cells // 2, cells // 2, 0), (1.0, 3.0, 5.0), args.pka_ev)
    e0, _ = energy()
    lead = args.pka_steps
    ctx.step(lead)                                   # the cascade develops (untimed)
    _, th0 = energy()
    ms = ctx.timed_steps(args.steps)
    e1, th1 = energy()
    ctx.profile_enable(True)
    ctx.step(min(args.steps, 50))
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    atoms = ctx.n_owned
    res = {"workload": "PKA cascade %.0f eV in bcc Fe %d^3 cells (%d atoms), dt 0.1 fs, steps %d..%d after the kick" % (args.pka_ev, cells, atoms, lead, lead + args.steps),
           "value": atoms * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps, "steps": args.steps, "atoms_total": atoms,
           "inter_atoms": [int(th0["n_inter"]), int(th1["n_inter"])], "runaways_last_step": int(th1["runaways"]),
           "energy_drift_ev_per_atom": (e1 - e0) / atoms, "pipelined_steps": int(ctx.query("pipe_steps")),
           "slots_ms_x_per_step": {k: [v[0] / max(v[1], 1), v[1] / min(args.steps, 50)] for k, v in prof.items() if v[1]}}
    ctx.close()
    return res


def fp64_view(kernels, stats, atoms):
    """The two stencil kernels against the fp64 pipe: fp64 warp-instruction lanes per second over the measured DFMA lane rate."""
    p = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        pk = json.load(f)
    out = {"peak_dfma_lanes_per_s": pk["dfma_per_s"], "peak_tflops": pk["fp64_tflops"], "peak_source": "profiles/r01_fp64_peak.json (tools/fp64_peak.cu, measured on this pool's B200)",
           "offsets_looped_per_atom": stats["offsets_per_atom"], "pair_evaluations_per_atom": stats["evals_per_atom"], "pairs_in_range_per_atom": stats["pairs_per_atom"]}
    for k in ("rho", "force"):
        if k in kernels:
            inst = FP64_INST[k]["eval"] * stats["evals_per_atom"] + FP64_INST[k]["test"] * (stats["offsets_per_atom"] - stats["evals_per_atom"])
            rate = inst * atoms / (kernels[k]["ms"] * 1e-3)
            out[k] = {"fp64_inst_per_atom": inst, "achieved_lanes_per_s": rate, "achieved_tflops_fma_equiv": 2 * rate / 1e12, "frac": rate / pk["dfma_per_s"]}
            # SURVEY.md section 8d's ALGORITHMIC count (what the reference's loop needs, not what this kernel issues):
            # K1 = N_off x 8 + N_pair x (sqrt + 7 FMA), K3 = N_pair x (sqrt + div + 3 x (index + 8 FMA) + 9); FMA = 2 flop
            n_off, n_pair = stats["offsets_per_atom"], stats["pairs_per_atom"]
            alg = n_off * 8 + n_pair * (1 + 14) if k == "rho" else n_pair * (1 + 1 + 3 * (1 + 16) + 9)
            out[k]["algorithmic_flop_per_atom"] = alg
            out[k]["algorithmic_tflops"] = alg * atoms / (kernels[k]["ms"] * 1e-3) / 1e12
            out[k]["algorithmic_frac"] = out[k]["algorithmic_tflops"] / pk["fp64_tflops"]
    return out


def hooks_whole_step(args):
    """The real drop-in, timed: the UNMODIFIED reference driver (oracle/_ref/libmisa_ref_cuda.so = its own sources compiled
    in place) stepping with the eight cuda_* hooks of arch_cuda/ -> this library; see tools/hooks_step.py. Runs in a child
    process (the reference aborts through MPI_Abort on any error) after this process has released the GPU memory."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "hooks_step.py"), str(args.cells), "3"] + [str(v) for v in args.ratio],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        if r.returncode != 0:
            return {"unavailable": "child exited %d: %s" % (r.returncode, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "")}
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # the leg is informative: never lose the line over it
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_b200(args):
    from misa_md_b200 import synth
    env = Env(args.gpus)
    n_gpus, rank = env.n_gpus, env.rank
    cells = args.cells
    parity = None if args.no_parity else parity_check(env, args)
    if parity is not None and not parity["ok"]:
        log(json.dumps(parity, indent=1))
        raise SystemExit("bench.py: parity self-check FAILED on %d rank(s): no number is reported for a wrong result" % env.world)

    clocks = ClockSampler(env.local_rank)  # started early: it is sampling long before the timed region begins
    ctx, main = timed_config(env, cells, args.ratio, args, clocks)
    clk = clocks.finish()
    atoms_per_gpu = ctx.n_owned
    kernels = main["kernels"]
    peak, peak_src = peaks()
    dom = max((k for k in kernels if "gbs" in kernels[k]), key=lambda k: kernels[k]["ms"])
    traffic, ncu_pipes, traffic_src = committed_traffic(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_atom": BYTES_PER_ATOM[dom], "atoms_per_launch": atoms_per_gpu,
                "note": "rho/force are bound by the LSU data pipe (32 lanes gathering random 16-byte table rows from shared memory), not by "
                        "HBM: ncu_pipes.lsu_wavefronts_pct against a measured saturation of 89 % of that pipe (DESIGN.md section 4.3d, "
                        "profiles/r04a_*); fp64 view in roofline_fp64; whole-step HBM fraction in step_hbm_frac; ncu_pipes = committed capture, % of peak",
                "lsu_wall": {"saturation_pct": 89.0, "source": "profiles/r04a_tex_frontend_ldg_all.csv (all neighbour fields on the LSU pipe)"},
                "ncu_pipes": ncu_pipes}

    # ---- e2e: host AoS buffers through the C ABI, H2D + D2H inside the timed region --------------------
    host = np.zeros(ctx.n_ext, dtype=synth.ATOM_DTYPE)
    ctx.host_register(host)  # pinned: the e2e leg copies from / to this array every step
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    ctx.download(host)
    ctx.step_host(host, 1)
    env.barrier(ctx)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.step_host(host, 1)
    env.barrier(ctx)
    e2e_s = env.reduce(time.perf_counter() - t0, "max")
    e2e = {"value": n_gpus * atoms_per_gpu * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ctx.n_owned * 104,
           "d2h_bytes_per_step": ctx.n_owned * 104, "steps": e2e_steps,
           "api": "misa_b200_step_host(ctx, AtomElement* host, 1): the owned box of the host AoS array (page-locked) goes up, one step runs, "
                  "the owned box comes back -- all 104-byte records both ways; on one sub-box as z-slabs on two copy streams with the "
                  "step's kernels running slab by slab between them (csrc/misa_b200.cu:step_host_slabs)",
           "slab_pipelined_steps": int(ctx.query("host_slab_steps")), "slab_redone": int(ctx.query("host_slab_redo"))}
    # the three reference hooks on the host array (EAM part of a step only; what the unmodified driver calls)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.eam_rho_calc(host); ctx.eam_df_calc(host); ctx.eam_force_calc(host)
    hooks_s = env.reduce((time.perf_counter() - t0) / 3, "max")
    e2e["hooks_eam_only_atom_passes_per_s"] = n_gpus * atoms_per_gpu / hooks_s

    if ctx.query("p2p_error"):
        raise RuntimeError("ghost push over peer memory timed out (p2p_error %d): numbers invalid" % ctx.query("p2p_error"))
    th = ctx.thermo()
    state = {"runaways_last_step": th["runaways"], "inter_atoms": th["n_inter"], "equil_steps": args.equil,
             "world_build_ms": main["world_build_ms"], "stencil_offsets": int(ctx.query("n_off")), "stencil_offsets_full": int(ctx.query("n_full")),
             "ghost_exchange": ("periodic fill in place" if n_gpus == 1 else "direct push over NVLink peer memory (csrc/p2p.cuh)" if ctx.query("p2p") else "staged NCCL send/recv"),
             "max_displacement_A": ctx.query("dmax"), "temperature_K": th["mvv"] * 1.0364269e-4 / ((3 * th["n_atoms"] - 3) * 8.617343e-5)}
    if os.environ.get("BENCH_P2P_DEBUG") and n_gpus > 1:   # experiments: phase timing of the push kernels (globaltimer stamps)
        ctx.set_option("p2p_debug", 1)
        ctx.step(50)
        state["p2p_push_ns"] = {w + "_" + ph: ctx.query("p2p_dbg_%s_%s" % (w, ph)) for w in ("x", "df") for ph in ("head", "body", "tail", "all")}
        ctx.set_option("p2p_debug", 0)
    ctx.host_unregister(host)
    ctx.close()
    del host

    # ---- the other BASELINE.json configurations, same process grid, each with its own timed region ----------------
    configs = {"fe_%d" % cells: {k: main[k] for k in ("workload", "value", "unit", "ms_per_step", "atoms_total", "kernels", "stencil")}}
    extra = [c for c in args.configs.split(",") if c]
    if "alloy" in extra and list(args.ratio) != [97, 2, 1]:
        # configs[2]: random Fe-Cu-Ni 97:2:1 (example/config.yaml:29-32); at N = 8 this is the 16 M-atom box the north-star target names
        c2, r2 = timed_config(env, cells, (97, 2, 1), args)
        r2["dilute_path"] = bool(c2.query("dilute"))
        c2.close()
        configs["alloy_97_2_1"] = r2
    if "cells200" in extra and cells != 200:
        # configs[4]: weak scaling at 16 M atoms per GPU (128 M atoms on 8 GPUs)
        c3, r3 = timed_config(env, 200, args.ratio, args)
        c3.close()
        configs["cells200"] = r3
    if "pka" in extra and n_gpus == 1:
        # configs[3]: PKA collision cascade (inter-atom and run-away paths under load)
        try:
            configs["pka_cascade"] = pka_config(env, cells, args)
        except Exception as e:   # the cascade is an extra configuration: its failure is reported in the line, it does not take the line away
            configs["pka_cascade"] = {"error": "%s: %s" % (type(e).__name__, e)}

    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(cells, args.ratio, n_gpus, args.equil),
        "roofline": roofline, "roofline_fp64": fp64_view(kernels, main["stencil"], atoms_per_gpu), "kernels": kernels,
        "step_hbm_gbs": main["step_hbm_frac"] * peak, "step_hbm_frac": main["step_hbm_frac"],
        "e2e": e2e, "gpu_launches": main["gpu_launches"], "clocks": clk, "state": state, "configs": configs, "parity_check": parity,
    }
    if rank == 0 and n_gpus == 1 and not args.no_hooks:
        line["e2e"]["hooks_whole_step"] = hooks_whole_step(args)
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    if rank == 0:
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=100, help="cells per dimension per GPU")
    ap.add_argument("--ratio", type=int, nargs=3, default=[1, 0, 0], help="Fe Cu Ni ratio (config 3: 97 2 1)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--equil", type=int, default=200, help="untimed thermalisation steps before warm-up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity self-check against the CPU oracle")
    ap.add_argument("--no-hooks", action="store_true", help="skip the reference-driver-on-hooks leg (N=1)")
    ap.add_argument("--configs", default="alloy,cells200,pka", help="extra BASELINE configs timed after the headline one")
    ap.add_argument("--pka-ev", type=float, default=5000.0, help="PKA energy of the cascade config (eV)")
    ap.add_argument("--pka-steps", type=int, default=900, help="untimed cascade steps between the kick and the timed region")
    args = ap.parse_args()
    if args.gpus not in GRIDS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    capture_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
