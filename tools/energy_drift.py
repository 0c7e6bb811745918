"""Total-energy drift of an NVE run on one GPU (north-star: 'total-energy drift over 1000 steps reported').
E = 0.5*mvv2e*sum(m v^2) + E_pot, E_pot = sum F(rho) + 1/2 sum phi (the reference never computes E_pot; ours, checked
against the oracle's in tests/test_gpu_parity.py::test_ten_steps_track_oracle)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import misa_md_b200 as mb
from misa_md_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ratio = tuple(int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (1, 0, 0)
P = (n, n, n)
ctx = mb.Context(P)
ctx.make_offsets()
ctx.set_potential(*mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH)))
ctx.build_world(seed=466953, t_set=600.0, ratio=ratio)   # WorldBuilder on the device, by global atom id
ctx.prepare()


def energy():
    th = ctx.thermo()
    ke = 0.5 * th["mvv"] * synth.MVV2E
    return ke, th["pe"], th["mvv"] * synth.MVV2E / ((3 * ctx.n_owned - 3) * synth.BOLTZ), th


rows = []
ke0, pe0, t0, _ = energy()
e0 = ke0 + pe0
for s in range(0, steps + 1, max(steps // 10, 1)):
    if s:
        ctx.step(max(steps // 10, 1))
    ke, pe, t, th = energy()
    rows.append(dict(step=s, ke=ke, pe=pe, e=ke + pe, T=t, runaways=th["runaways"], inter=th["n_inter"]))
    print("step %5d  KE %.6f  PE %.6f  E %.6f  T %.2f K  dE/atom %.3e eV" % (s, ke, pe, ke + pe, t, (ke + pe - e0) / ctx.n_owned), flush=True)
e1 = rows[-1]["e"]
out = dict(cells=n, atoms=ctx.n_owned, steps=steps, dt_ps=0.001, ratio=list(ratio), e0=e0, e1=e1,
           drift_ev_per_atom=(e1 - e0) / ctx.n_owned, drift_rel=(e1 - e0) / abs(e0),
           max_abs_dev_ev_per_atom=max(abs(r["e"] - e0) for r in rows) / ctx.n_owned, samples=rows)
print(json.dumps(out))
