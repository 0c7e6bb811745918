"""Dev-only: nvcc -D variants of libmisa_b200.so under build/variants/ for tools/time_variants.py (A/B on one box).
usage: python tools/build_variants.py name1:-DEAM_NB_LDG=4 name2:-DEAM_NB_LDG=6,-DEAM_THREADS=512 ...   (built in parallel)"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
sys.path.insert(0, root)
from misa_md_b200 import build as b

out = os.path.join(root, "build", "variants")
os.makedirs(out, exist_ok=True)


def one(spec):
    name, _, flags = spec.partition(":")
    lib = os.path.join(out, name + ".so")
    cmd = [b.nvcc_path()] + [f for f in b.NVCC_FLAGS if not f.startswith("--use_fast_math")] + [f for f in flags.split(",") if f] + \
          ["-o", lib] + [os.path.join(b.CSRC, s) for s in b.SOURCES] + ["-ldl"]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    return name, r.returncode, r.stdout[-2000:]


with ThreadPoolExecutor(max_workers=int(os.environ.get("JOBS", "6"))) as ex:
    for name, rc, log in ex.map(one, sys.argv[1:]):
        print(name, "ok" if rc == 0 else "FAILED\n" + log, flush=True)
