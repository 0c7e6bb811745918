"""BASELINE.json configs[3]: PKA collision cascade in bcc Fe at full size (100^3 cells, 2 M atoms) on one GPU -- the
inter-atom and run-away paths (atom::setv, decide, interRho/interForce, vacancy re-occupation) under load.
Stage machine as in the reference's example (frontend/md_simulation.cpp:40-80, config.yaml:58-65): thermalise at dt 1 fs,
rescale, kick one atom, run the cascade at dt 0.1 fs, all without leaving resident mode.
usage: python tools/pka_cascade.py [cells=100] [energy_eV=5000] [cascade_steps=2000]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import misa_md_b200 as mb
from misa_md_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
energy = float(sys.argv[2]) if len(sys.argv) > 2 else 5000.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
P = (n, n, n)
ctx = mb.Context(P)
ctx.make_offsets()
ctx.set_potential(*mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH)))
ctx.build_world(seed=466953, t_set=600.0, ratio=(1, 0, 0))
ctx.set_timestep(0.001)
ctx.prepare()
ctx.step(200)
ctx.rescale_to(300.0)                      # stage: rescale (configuration::rescale)
ctx.step(50)


def energy_now():
    th = ctx.thermo()
    return 0.5 * th["mvv"] * synth.MVV2E + th["pe"], th


e_before, th = energy_now()
ctx.set_timestep(1e-4)                     # stage: collision, dt 0.1 fs (NewtonMotion::setTimestepLength)
ctx.collision_step((n // 2, n // 2, n // 2, 0), (1.0, 3.0, 5.0), energy)
e0, th = energy_now()
print("kick %.0f eV: E before %.3f, after %.3f (delta %.3f eV)" % (energy, e_before, e0, e0 - e_before), flush=True)
rows, t_all = [], 0.0
chunk = max(steps // 10, 1)
for s in range(chunk, steps + 1, chunk):
    t0 = time.perf_counter()
    ctx.step(chunk)
    ctx.sync()
    dt = time.perf_counter() - t0
    t_all += dt
    e, th = energy_now()
    rows.append(dict(step=s, e=e, de_per_atom=(e - e0) / ctx.n_owned, inter=th["n_inter"], runaways_last=th["runaways"], ms_per_step=1e3 * dt / chunk))
    print("step %5d  E %.3f  dE %.3e eV/atom  inter atoms %d  run-aways(last step) %d  %.3f ms/step" % (
        s, e, (e - e0) / ctx.n_owned, th["n_inter"], th["runaways"], 1e3 * dt / chunk), flush=True)
if os.environ.get("PKA_PROFILE"):          # per-slot CUDA-event times of the last state (inter atoms present)
    ctx.profile_enable(True)
    ctx.step(50)
    pr = ctx.profile_read()
    ctx.profile_enable(False)
    print("slots (ms per launch x launches per step):", {k: (round(v[0] / max(v[1], 1), 4), v[1] / 50) for k, v in pr.items() if v[1]}, flush=True)
    print("stencil stats:", ctx.stencil_stats(), "novac", ctx.query("novac"), "mark level", ctx.query("mark_level"), "dmax", ctx.query("dmax"), flush=True)
rec = ctx.dump_records(steps)
vac = ctx.n_owned - (rec.size - rows[-1]["inter"])
print(json.dumps(dict(cells=n, atoms=ctx.n_owned, pka_ev=energy, dt_ps=1e-4, steps=steps, ms_per_step=1e3 * t_all / steps,
                      atom_steps_per_s=ctx.n_owned * steps / t_all, inter_atoms_end=rows[-1]["inter"], vacancies_end=int(vac),
                      dump_records=int(rec.size), pipelined_steps=ctx.query("pipe_steps"), redone_serially=ctx.query("pipe_redo"),
                      energy_drift_ev_per_atom=rows[-1]["de_per_atom"], samples=rows)))
