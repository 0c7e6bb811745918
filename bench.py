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
        "config": {"workload": "bcc Fe 100^3 cells (2M atoms) NVE, synthetic FeCuNi setfl; CPU arm runs a bounded sample",
                   "cells_per_gpu": [100, 100, 100]},
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
def run_b200(args):
    import torch
    import torch.distributed as dist
    import misa_md_b200 as mb
    from misa_md_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    if world != n_gpus:
        if world == 1 and n_gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (n_gpus, n_gpus))
        n_gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    grid = GRIDS[n_gpus]
    coord = (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])
    cells = args.cells
    phase = tuple(cells * g for g in grid)

    lib = mb.load()
    mb.capi._ck(lib.misa_b200_env_init(local_rank))
    ctx = mb.Context(phase, grid=grid, coord=coord, a=A, crf=CRF)
    ctx.make_offsets()
    ctx.set_potential(*mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH)))
    ctx.set_timestep(DT)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    clocks = ClockSampler(local_rank)  # started early: it is sampling long before the timed region begins
    # initial state built on the device by global atom id (misa_b200_build_world = WorldBuilder::build, seed 466953,
    # 600 K): every sub-box cuts its part out of the same global state, no host init, no H2D
    t_build = time.perf_counter()
    ctx.build_world(seed=466953, t_set=600.0, ratio=tuple(args.ratio))
    t_build = time.perf_counter() - t_build
    host = np.zeros(ctx.n_ext, dtype=synth.ATOM_DTYPE)
    ctx.host_register(host)  # pinned: the e2e leg copies from / to this array every step
    ctx.prepare()
    atoms_per_gpu = ctx.n_owned

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident mode: inputs in HBM when the timed region starts ------------------------------------
    # Untimed thermalisation first: the state starts as a PERFECT lattice with 600 K of kinetic energy; until the
    # positions have thermalised (~100 steps) fewer pairs fall inside the cutoff and the pruned stencil is shorter,
    # which would flatter the timed region. The timed steps are those of the stationary NVE run BASELINE.json names.
    ctx.step(args.equil)
    ctx.step(max(args.warmup, 3))
    barrier()
    l0 = ctx.launch_count()
    clocks.mark_start()
    ms = ctx.timed_steps(args.steps)  # CUDA events on the stream every kernel of the step is launched on
    barrier()
    clocks.mark_stop()
    launches = ctx.launch_count() - l0
    ms = max_over_ranks(ms)
    clk = clocks.finish()
    value = n_gpus * atoms_per_gpu * args.steps / (ms * 1e-3)

    # ---- per-kernel CUDA-event durations (separate pass; event pairs around every kernel slot) ---------
    ctx.profile_enable(True)
    ctx.step(args.steps)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    peak, peak_src = peaks()
    kernels = {}
    for name, (tot_ms, cnt) in prof.items():
        if cnt and name in BYTES_PER_ATOM:
            t = tot_ms / cnt * 1e-3
            gbs = BYTES_PER_ATOM[name] * atoms_per_gpu / t / 1e9
            kernels[name] = {"ms": t * 1e3, "gbs": gbs, "frac": gbs / peak}
        elif cnt:
            kernels[name] = {"ms": tot_ms / cnt}
    dom = max((k for k in kernels if "gbs" in kernels[k]), key=lambda k: kernels[k]["ms"])
    traffic, ncu_pipes = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes/launch + pipe utilisation from the committed ncu --set full capture
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic, ncu_pipes = tj.get(dom), tj.get(dom + "_pipes")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "bytes_per_atom": BYTES_PER_ATOM[dom], "atoms_per_launch": atoms_per_gpu,
                "note": "rho/force are fp64-pipe/LSU bound, not HBM bound (DESIGN.md section 4); whole-step "
                        "HBM fraction in step_hbm_frac; ncu_pipes = what does bound the kernel (committed capture, % of peak)",
                "ncu_pipes": ncu_pipes}
    step_gbs = STEP_BYTES_PER_ATOM * atoms_per_gpu * args.steps / (ms * 1e-3) / 1e9

    # ---- e2e: host AoS buffers through the C ABI, H2D + D2H inside the timed region --------------------
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    ctx.download(host)
    ctx.step_host(host, 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.step_host(host, 1)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": n_gpus * atoms_per_gpu * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ctx.n_owned * 104,
           "d2h_bytes_per_step": ctx.n_owned * 104, "steps": e2e_steps,
           "api": "misa_b200_step_host(ctx, AtomElement* host, 1): upload the owned box of the host AoS array (pitched 3-D copy "
                  "from page-locked memory), one step, download the owned box"}
    # the three reference hooks on the host array (EAM part of a step only; what the unmodified driver calls)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.eam_rho_calc(host); ctx.eam_df_calc(host); ctx.eam_force_calc(host)
    hooks_s = max_over_ranks((time.perf_counter() - t0) / 3)
    e2e["hooks_eam_only_atom_passes_per_s"] = n_gpus * atoms_per_gpu / hooks_s

    if ctx.query("p2p_error"):
        raise RuntimeError("ghost push over peer memory timed out (p2p_error %d): numbers invalid" % ctx.query("p2p_error"))
    th = ctx.thermo()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "bcc Fe %d^3 cells (%d atoms) per GPU, NVE, dt 1 fs, T0 600 K, synthetic FeCuNi setfl" % (cells, atoms_per_gpu),
                   "cells_per_gpu": [cells] * 3, "grid": list(grid), "atoms_total": n_gpus * atoms_per_gpu,
                   "species_ratio": list(args.ratio), "equil_steps": args.equil,
                   "l2": "resident state %.0f MB per GPU exceeds the 126 MB L2; no flush between steps" % (ctx.n_ext * 105 / 1e6)},
        "roofline": roofline, "kernels": kernels, "step_hbm_gbs": step_gbs, "step_hbm_frac": step_gbs / peak,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        "state": {"runaways_last_step": th["runaways"], "inter_atoms": th["n_inter"], "equil_steps": args.equil,
                  "world_build_ms": 1e3 * t_build, "stencil_offsets": int(ctx.query("n_off")), "stencil_offsets_full": int(ctx.query("n_full")),
                  "ghost_exchange": ("periodic fill in place" if n_gpus == 1 else "direct push over NVLink peer memory (csrc/p2p.cuh)" if ctx.query("p2p") else "staged NCCL send/recv"),
                  "max_displacement_A": ctx.query("dmax"), "temperature_K": th["mvv"] * 1.0364269e-4 / ((3 * th["n_atoms"] - 3) * 8.617343e-5)},
    }
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    ctx.host_unregister(host)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
