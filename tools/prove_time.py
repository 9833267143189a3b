#!/usr/bin/env python3
"""The 3-party prove entries of bench.py alone (no CPU arm), with library options from the environment:
OPTS="msm_reduce_warp_max=512,msm_task_len=16" tools/prove_time.py"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge
import bench
pkg = ge.load_package(); H, S = pkg.host, pkg.synth
H.init(); H.set_party(0, 1)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("=")
    H.set_option(k, int(v))
args = argparse.Namespace(log_n=int(os.environ.get("LOG_N", "13")), no_cpu=True)
r = bench.bench_prove(pkg, H, S, args)
print(json.dumps({"opts": os.environ.get("OPTS", ""), **{k: round(v["prove_hot_path_s"] * 1e3, 3) for k, v in r.items()}}))
