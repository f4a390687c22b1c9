#!/usr/bin/env python
"""Container: Blackwell-only instruction census of a built object (cuobjdump -sass) + an excerpt around the first tcgen05.mma.
Usage: python scripts/sass_summary.py prior_flow_b200/build/pf_volume_tc.o volume_tc_kernel > profiles/rNN_sass_<kernel>.txt"""
import collections
import re
import subprocess
import sys

obj, main_name = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
print(f"cuobjdump -sass {obj}  (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)")
print("Blackwell-only instruction classes per kernel: UTCHMMA = tcgen05.mma (kind::f16/tf32), LDTM = tcgen05.ld, STTM = tcgen05.st,\n"
      "UTMALDG = TMA load (.MULTICAST across the CTA pair), UTMASTG = TMA store, UTCBAR = tcgen05.commit -> mbarrier,\n"
      "UTCATOMSWS = TMEM alloc/dealloc, SYNCS = mbarrier ops, UCGABAR = cluster barrier, ACQBULK = bulk-async acquire.\n")
pat = re.compile(r"(UTC|LDTM|STTM|UTMA|SYNCS|UCGABAR|ACQBULK|ELECT|UBLKCP|UBLKPF|CCTL)")
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter(ops)
    keys = sorted(k for k in c if pat.match(k))
    print(f"{name}\n  {len(ops)} SASS instructions; " + (", ".join(f"{k} x{c[k]}" for k in keys) or "no Blackwell-only instructions"))
    if main_name and main_name in name:
        lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
        idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l or "UTCQMMA" in l]
        if idx:
            print("  excerpt around the first tcgen05.mma group:")
            for l in lines[max(0, idx[0] - 10):idx[0] + 30]:
                print("    " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).strip())
