"""Which half / which ingredient is nondeterministic?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashpca_b200.synth import SynthSpec

n, p, miss = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
s = SynthSpec(n, p, seed=20240603, missing_rate=miss)
op = s.create_operator()
rng = np.random.default_rng(0)
x = rng.standard_normal(n)
v = rng.standard_normal(p)
ts = [op.crossprod(x) for _ in range(5)]
ys = [op.prod(v) for _ in range(5)]
def rep(name, arr):
    bad = [i for i in range(1, len(arr)) if not np.array_equal(arr[0], arr[i])]
    d = max([np.abs(arr[i] - arr[0]).max() for i in bad], default=0.0)
    print("  %-10s nondeterministic reps %s, max diff/scale %.3e" % (name, bad, d / np.abs(arr[0]).max()))
print("env GEMV=%s PRIO=%s GATHER=%s miss=%g" % (os.environ.get("FPB_GEMV"), os.environ.get("FPB_PRIO"), os.environ.get("FPB_GATHER"), miss))
rep("crossprod", ts)
rep("prod", ys)
# pattern of differing outputs in crossprod (rows = SNPs): TMA tiles are 256 rows, gather blocks 8 rows
d = np.abs(ts[1] - ts[0]) > 0
idx = np.nonzero(d)[0]
print("  crossprod differing SNPs: %d of %d" % (idx.size, d.size))
if idx.size:
    runs = np.split(idx, np.nonzero(np.diff(idx) > 1)[0] + 1)
    lens = np.array([len(r) for r in runs])
    starts = np.array([r[0] for r in runs])
    print("  runs: %d, run length min/median/max %d/%d/%d" % (len(runs), lens.min(), np.median(lens), lens.max()))
    print("  first run starts:", starts[:12], " starts mod 256:", (starts[:12] % 256), " mod 8:", (starts[:12] % 8))
    print("  first run lengths:", lens[:12])
    rel = np.abs(ts[1] - ts[0])[idx] / np.abs(ts[0]).max()
    print("  rel diff quantiles:", np.quantile(rel, [0.1, 0.5, 0.9]))
