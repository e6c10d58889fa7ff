"""development aid: SASS evidence per kernel of the built library (TMA / mbarrier / redux mnemonics, FP64 and memory instruction
counts) -> profiles/r02_sass_summary.txt.  Runs here (cuobjdump, no GPU needed)."""
import os, re, subprocess
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "hyperelasticsolver_b200", "libhyperelastic_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
out = ["# SASS evidence (cuobjdump -sass hyperelasticsolver_b200/libhyperelastic_b200.so, sm_100a), static instruction counts per kernel",
       "# TMA: UTMALDG.2D = cp.async.bulk.tensor.2d tile copies, UBLKCP = cp.async.bulk row copies, SYNCS = mbarrier ops; CREDUX/REDUX = redux.sync max",
       "# regenerate: python tools/sass_summary.py", ""]
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if not any(k in dem for k in ("k_step", "k_bounds", "k_exchange", "k_transpose_rot", "k_dt2d")):
        continue
    c, n = Counter(), 0
    for line in f.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        n += 1
        op = re.sub(r"^@!?U?P\w+\s+", "", m.group(1)).split()[0]
        base = op.split(".")[0]
        if base in ("UTMALDG", "SYNCS"):
            c[".".join(op.split(".")[:3])] += 1
        elif base in ("UBLKCP", "CREDUX", "REDUX", "ATOMG", "BAR", "LDS", "STS", "SHFL", "MUFU", "LDG", "STG"):
            c[base] += 1
        elif base in ("DFMA", "DMUL", "DADD", "DSETP"):
            c["FP64"] += 1
    short = re.sub(r"\(.*", "", dem.replace("hs::", ""))
    out.append(f"{short:60s} total {n:5d}  " + "  ".join(f"{k} {v}" for k, v in sorted(c.items())))
open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-12:]))
