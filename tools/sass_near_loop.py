"""SASS excerpt + instruction accounting of the branch-free near loop of the production stencil kernels (the loop with four
MUFU.RSQ64H: EAM_UNROLL_NEAR = 4 neighbour pairs per lane and iteration). usage: python tools/sass_near_loop.py > profiles/...txt
(needs cuobjdump; runs on the CPU box -- the cubin is in misa_md_b200/libmisa_b200.so)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "misa_md_b200", "libmisa_b200.so")], stdout=subprocess.PIPE, text=True).stdout.split("\n")
KERNELS = [("k_force_f<1,1,0,0,0>  latForce, single species, no per-neighbour type test", "_Z9k_force_fILb1ELb1ELb0ELb0ELb0EE"),
           ("k_rho_f<1,1,1,0,0,0>  latRho + latDf", "_Z7k_rho_fILb1ELb1ELb1ELb0ELb0ELb0EE")]
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def function(mangled):
    out, on = [], False
    for l in sass:
        if "Function :" in l:
            if on:
                break
            on = mangled in l
        if on:
            out.append(l)
    return out


def op(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]


print("# SASS of the branch-free near loop (four neighbour pairs per lane and iteration) of the production stencil kernels")
print("# cuobjdump -sass misa_md_b200/libmisa_b200.so (nvcc 12.9, sm_100a), tools/sass_near_loop.py")
for title, mangled in KERNELS:
    ins = []
    for l in function(mangled):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            body = [u for b, u in ins if int(m.group(1), 16) <= b <= a]
            if sum("MUFU.RSQ64H" in u for u in body) == 4 and (best is None or len(body) < len(best[2])):
                best = (int(m.group(1), 16), a, body)
    lo, hi, body = best
    cnt = collections.Counter(op(t) for t in body)
    fp64 = sum(cnt[k] for k in FP64)
    print("\n== %s\n   near loop 0x%04x..0x%04x: %d instructions for 4 pairs per lane" % (title, lo, hi, len(body)))
    print("   fp64 pipe: %d = %.2f per pair  %s;  MUFU.RSQ64H 4;  LDS %d (%.0f per pair);  TLD %d;  other %d" % (
        fp64, fp64 / 4, {k: cnt[k] for k in FP64 if cnt[k]}, cnt["LDS"], cnt["LDS"] / 4, cnt["TLD"], len(body) - fp64 - 4 - cnt["LDS"] - cnt["TLD"]))
    print("   " + str(dict(cnt.most_common())))
    for t in body:
        print("        " + t)
