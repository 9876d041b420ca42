"""Manual soak test (not collected by pytest): random, mostly damaged inputs through the kernels on the CPU emulation,
both paths, random tile size / CTA count / input phase / thread order.   python tests/host_stub/fuzz_kernels.py <seed> <seconds>
(build tests/_build/kernels_harness first: python -m pytest tests/test_kernels_on_cpu.py -k bench_workload)"""
import sys, os, subprocess, time
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np
from test_oracle_fuzz_vs_ref import _records, _damage, FLAGSETS
import test_kernels_on_cpu as t
exe = os.environ.get("KH_EXE") or os.path.join(t.BUILD, "kernels_harness")
seed0 = int(sys.argv[1]); budget = float(sys.argv[2])
rng = np.random.default_rng(seed0)
t0 = time.time(); n = 0; bad = []; stats = {}
p = '/tmp/fz_%d.fq' % seed0
while time.time() - t0 < budget:
    qualtype = ["sanger", "illumina", "solexa"][n % 3]
    clean = bool(os.environ.get("FZ_CLEAN"))        # undamaged reads of 70-260 bases: the fused kernels keep nearly all of them
    lmax = int(rng.choice([90, 120, 151, 200, 260])) if clean else int(rng.choice([8, 25, 60, 110, 151, 260, 700, 3000]))
    nrec = int(rng.integers(2, 40)) if lmax > 700 else int(rng.integers(40, 1500))
    data = _records(rng, nrec, lmax, qualtype)
    if not clean and rng.random() < 0.5:
        data = _damage(rng, data)
    if data and not data.endswith(b"\n") and rng.random() < 0.7:
        data = data[:-1] + b"\n"
    open(p, 'wb').write(data)
    fl = FLAGSETS[int(rng.integers(0, len(FLAGSETS)))]
    mode = ["se", "pei", "peM"][int(rng.integers(0, 3))]
    kw = dict(mode=mode, qualtype=qualtype, q=fl["q"], l=fl["l"], x=fl["x"], n=fl["n"], singles=bool(rng.integers(0, 2)) if mode == "pei" else (mode != "peM"),
              first=int(rng.integers(0, 16)), ctas=int(rng.integers(1, 7)))
    env = {"SIMT_SHUFFLE": str(int(rng.integers(1, 1000)))} if rng.random() < 0.3 else None
    # FZ_ORDER=1: the -a N paths instead -- index pass + routing + K3 (any mode), index pass + ordered emit (single end, N <= 32)
    threads = int(rng.choice([2, 3, 4, 7, 8, 16, 31, 32])) if os.environ.get("FZ_ORDER") else 1
    if threads > 1:
        kernels = ["general", "index%d" % int(rng.choice([3, 5, 7, 9]))] + (["order%d" % int(rng.choice([3, 5, 7, 9]))] if mode == "se" else [])
    else:
        kernels = ("general", ["fused5", "fused7", "fused9", "fused11"][int(rng.integers(0, 4))])
    for k in kernels:
        rc, out, err = t.run(exe, p, kernel=k, env=env, threads=threads, **kw)
        key = (k[:5], out.split()[0] if out else "rc%d" % rc)
        stats[key] = stats.get(key, 0) + 1
        ok = rc == 0 and (out.startswith("OK") or (k != "general" and out.startswith("FASTFAIL")))
        if not ok:
            keep = '/tmp/fz_bad_%d_%d.fq' % (seed0, n)
            open(keep, 'wb').write(data)
            bad.append((keep, k, kw, env, out, err[-300:]))
    n += 1
print("seed", seed0, "cases", n, "stats", stats, "bad", len(bad))
for b in bad[:5]:
    print(b)
