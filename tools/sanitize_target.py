"""Target for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel family of the library once,
on boxes small enough for the tools' 10-100x slow-down. Not a parity test (tests/ do that) -- it only has to LAUNCH
everything: thermal steps (pruned lists, fused half-kicks), alloy + vacancies (dilute path, type tests), a PKA cascade
(device-resident inter-atom lists, run-aways, re-occupation), the slab-pipelined host step, the drop-in hook passes,
world builder, dump, thermostat.
usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py [scenario ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import misa_md_b200 as mb
from misa_md_b200 import synth

POT = mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH))


def context(cells):
    ctx = mb.Context(cells)
    ctx.make_offsets()
    ctx.set_potential(*POT)
    return ctx


def state(cells, ratio=(1, 0, 0), vac=0):
    st = synth.create_global_state(cells, ratio=ratio)
    synth.perturb_positions(st, 0.03)
    if vac:
        flat = st["type"].reshape(-1)
        flat[np.random.RandomState(99).choice(flat.size, vac, replace=False)] = synth.INVALID
    return synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))[0]


def thermal():
    ctx = context((40, 12, 12))
    ctx.upload(state((40, 12, 12)))
    ctx.prepare()
    ctx.step(6)
    ctx.step(1)
    print("thermal", ctx.thermo(), ctx.stencil_stats(), flush=True)
    ctx.close()


def alloy():
    ctx = context((16, 12, 12))
    ctx.upload(state((16, 12, 12), (97, 2, 1), 5))
    ctx.prepare()
    ctx.step(5)
    print("alloy", ctx.thermo(), flush=True)
    ctx.close()


def world_dump_thermostat():
    ctx = context((16, 12, 12))
    ctx.build_world(seed=466953, t_set=600.0, ratio=(97, 2, 1))
    ctx.prepare()
    ctx.step(3)
    t = ctx.temperature()
    ctx.rescale_to(300.0)
    rec = ctx.dump_records(3)
    print("world", t, ctx.temperature(), rec.size, flush=True)
    ctx.close()


def cascade():
    ctx = context((16, 16, 16))
    ctx.build_world(seed=466953, t_set=300.0, ratio=(1, 0, 0))
    ctx.set_timestep(1e-4)
    ctx.prepare()
    ctx.step(2)
    ctx.collision_step((8, 8, 8, 0), (1.0, 3.0, 5.0), 800.0)
    for _ in range(6):
        ctx.step(50)
        th = ctx.thermo()
        print("cascade inter %d runaways %d" % (th["n_inter"], th["runaways"]), flush=True)
    ctx.download_inter()
    ctx.dump_records(300)
    ctx.close()


def cascade_host_lists():
    ctx = context((16, 16, 16))
    ctx.set_option("inter_dev", 0)
    ctx.build_world(seed=466953, t_set=300.0, ratio=(1, 0, 0))
    ctx.set_timestep(1e-4)
    ctx.prepare()
    ctx.collision_step((8, 8, 8, 0), (1.0, 3.0, 5.0), 800.0)
    ctx.step(150)
    print("cascade (host lists) inter %d" % ctx.thermo()["n_inter"], flush=True)
    ctx.close()


def host_step():
    ctx = context((16, 12, 24))
    arr = state((16, 12, 24))
    ctx.host_register(arr)
    for _ in range(3):
        ctx.step_host(arr, 1)
    ctx.host_unregister(arr)
    print("host_step slabs", ctx.query("host_slab_steps"), flush=True)
    ctx.close()


def hooks():
    ctx = context((16, 12, 12))
    arr = state((16, 12, 12), (97, 2, 1), 3)
    ctx.eam_rho_calc(arr)
    ctx.eam_df_calc(arr)
    ctx.eam_force_calc(arr)
    print("hooks", float(np.abs(arr["f"]).sum()), flush=True)
    ctx.close()


def sym():
    ctx = context((16, 12, 12))
    ctx.set_option("sym", 1)
    ctx.upload(state((16, 12, 12)))
    ctx.prepare()
    ctx.step(3)
    print("sym", ctx.thermo(), flush=True)
    ctx.close()


ALL = dict(thermal=thermal, alloy=alloy, world=world_dump_thermostat, cascade=cascade, cascade_host=cascade_host_lists,
           host_step=host_step, hooks=hooks, sym=sym)
for name in (sys.argv[1:] or list(ALL)):
    ALL[name]()
print("sanitize_target done", flush=True)
