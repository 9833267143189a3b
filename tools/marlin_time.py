#!/usr/bin/env python3
"""bench.py's Marlin entry alone: tools/marlin_time.py [log_h ...]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge
import bench
pkg = ge.load_package(); H, S = pkg.host, pkg.synth
H.init(); H.set_party(0, 1)
for log_h in [int(a) for a in sys.argv[1:]] or [16, 20]:
    r = bench.bench_marlin(pkg, H, S, log_h)
    r.pop("what")
    print(json.dumps(r), flush=True)
