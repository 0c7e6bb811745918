"""Scratch timing of the section-8f entry points on one GPU: device world build, dump compaction, temperature."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import misa_md_b200 as mb
from misa_md_b200 import synth

for n in [int(a) for a in sys.argv[1:]] or [100]:
    P = (n, n, n)
    ctx = mb.Context(P)
    ctx.make_offsets()
    ctx.set_potential(*mb.capi.potential_in_type_order(mb.capi.read_setfl(mb.SETFL_PATH)))
    for rep in range(2):
        t0 = time.perf_counter(); ctx.build_world(ratio=(97, 2, 1)); ctx.sync(); t1 = time.perf_counter()
    print("cells %d^3 (%d atoms): build_world %.1f ms" % (n, 2 * n ** 3, 1e3 * (t1 - t0)), flush=True)
    t0 = time.perf_counter(); st = synth.create_global_state(P, ratio=(97, 2, 1)); t1 = time.perf_counter()
    arr, _ = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0)); t2 = time.perf_counter()
    ctx2 = mb.Context(P); ctx2.upload(arr); ctx2.sync(); t3 = time.perf_counter()
    print("   host numpy init %.1f ms + scatter %.1f ms + upload %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
    ctx2.close()
    ctx.prepare(); ctx.step(5)
    out = np.zeros(2 * n ** 3, dtype=synth.DUMP_DTYPE)
    ctx.L.misa_b200_host_register(out.ctypes.data, out.nbytes)
    for rep in range(3):
        t0 = time.perf_counter(); rec = ctx.dump_records(5, out=out); t1 = time.perf_counter()
    host = np.zeros(ctx.n_ext, dtype=synth.ATOM_DTYPE)
    ctx.L.misa_b200_host_register(host.ctypes.data, host.nbytes)
    for rep in range(3):
        t2 = time.perf_counter(); ctx.download(host); t3 = time.perf_counter()
    print("   dump_records %.2f ms (%d records, %.1f MB) vs full AoS download %.2f ms (%.1f MB)" % (
        1e3 * (t1 - t0), rec.size, rec.nbytes / 1e6, 1e3 * (t3 - t2), host.nbytes / 1e6), flush=True)
    for rep in range(3):
        t0 = time.perf_counter(); th = ctx.temperature(); t1 = time.perf_counter()
    print("   temperature %.3f ms -> T = %.2f K" % (1e3 * (t1 - t0), th["T"]), flush=True)
    ctx.close()
