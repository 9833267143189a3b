#!/usr/bin/env python3
"""Where one party's prove call spends its wall time (single party, no threads): MySecretInputCircuit shape by default,
`tools/prove_profile.py 20` for the synthetic 2^20 instance of bench.py."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, G = pkg.host, pkg.synth, pkg.groth16
H.init(); H.set_party(0, 1)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
n = 1 << log_n
nc, ni, nv = (6574, 5, 6600) if log_n == 13 else (n - 8, 5, n)
mats = S.r1cs_matrices(0xB10 + log_n, nc, nv)
gen1 = lambda seed, c: (lambda b: (b.download().reshape(c, 12), b.free())[0])(H.g1_generate(seed, c))
gen2 = lambda seed, c: (lambda b: (b.download().reshape(c, 24), b.free())[0])(H.g2_generate(seed, c))
pkarr = S.proving_key_arrays(gen1, gen2, 0xB20, nv, ni, n)
z = S.fr_uniform(0xB30, nv)
pk = G.ProvingKey(**pkarr); r1cs = G.R1CS(*mats, num_inputs=ni, num_vars=nv)


class Net:
    party, n_parties = 0, 1
    def exchange(self, p): return [p]


sess = G.ProverSession(pk, r1cs)
# monkey-patch timers around the host calls
times = {}
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); times[name] = times.get(name, 0) + time.perf_counter() - t0; return r
    setattr(mod, name, g)
for nm in ("witness_map_begin_r1cs", "witness_map_masked_payload", "witness_map_open_payloads", "witness_map_finish_dev", "msm_handle_dev", "sum_partials"):
    wrap(H, nm)
for it in range(5):
    times.clear()
    t0 = time.perf_counter()
    sess.prove(z, Net())
    total = time.perf_counter() - t0
print(json.dumps({"total_ms": round(total * 1e3, 3), **{k: round(v * 1e3, 3) for k, v in times.items()}}))
