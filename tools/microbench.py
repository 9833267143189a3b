#!/usr/bin/env python3
"""Integer-pipe microbenchmarks on the current GPU (writes JSON to stdout)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge
pkg = ge.load_package()
H = pkg.host
H.init()
names = {0: "imad_u32_gops", 1: "imad_wide_carry_gops", 2: "fq_mul_wide_gmul", 3: "fr_mul_wide_gmul",
         4: "fq_mul_narrow_gmul", 5: "fr_mul_narrow_gmul", 6: "dfma_gops"}
out = {}
for k, nm in names.items():
    it = 20000 if k < 2 or k == 6 else 2000
    out[nm] = max(H.microbench(k, it) for _ in range(3))
print(json.dumps(out))
