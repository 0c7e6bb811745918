"""profiles/traffic.json from an `ncu --set full` raw CSV export: DRAM bytes per launch and pipe utilisation of the stencil
kernels, stamped with the hash of the kernel sources the capture was taken on (bench.py refuses a stale stamp).
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/make_traffic.py raw.csv "<provenance text>" """
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = lambda r, k: float(r[hdr.index(k)].replace(",", ""))
unit = lambda k: rows[1][hdr.index(k)]
out = {"_source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "source_sha256_16": bench.source_hash(), "kernels": {}}
for name, pat in (("rho", "k_rho_f"), ("force", "k_force_f"), ("verlet1", "k_verlet1"), ("verlet2", "k_verlet2")):
    sel = [r for r in data if pat in r[hdr.index("Kernel Name")]]
    if not sel:
        continue
    r = sel[-1]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    dram = sum(col(r, k) * scale[unit(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out[name] = int(dram)
    out["kernels"][name] = r[hdr.index("Kernel Name")]
    out[name + "_pipes"] = {"lsu_wavefronts_pct": round(col(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), 1),
                            "fp64_pipe_pct": round(col(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 1),
                            "tex_wavefronts_pct": round(col(r, "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed"), 1),
                            "issue_active_pct": round(col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                            "dram_throughput_pct": round(col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 1),
                            "duration_us": col(r, "gpu__time_duration.sum")}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
