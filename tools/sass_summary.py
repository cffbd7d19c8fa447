"""Opcode evidence per kernel of libmagcache_b200.so (no GPU needed): python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG / UBLKCP = TMA, ELECT = elect.sync, USETMAXREG = setmaxnreg, SYNCS = mbarrier, MUFU.EX2, packed fp32 (FFMA2/FADD2), 256-bit LDG."""
import os
import re
import subprocess
import sys
from collections import Counter

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "magcache_b200", "libmagcache_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "ELECT", "USETMAXREG", "SYNCS", "MUFU.EX2", "MUFU.TANH", "FFMA2", "FADD2", "FMNMX3",
        "F2FP", "LDG.E.256", "STG.E.256", "HMMA", "BAR.SYNC", "FENCE", "LDL", "STL"]
print(f"# SASS opcode summary of {os.path.basename(so)} (cuobjdump -sass, sm_100a); columns = instruction counts in the kernel's code\n")
for f in re.split(r"\n\s*Function : ", sass)[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem)
    ops = Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M))
    tot = sum(ops.values())
    found = {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in KEYS}
    found = {k: v for k, v in found.items() if v}
    print(f"{dem}: {tot} instructions; " + ", ".join(f"{k} {v}" for k, v in found.items()))
