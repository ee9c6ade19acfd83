"""Determinism / race probe: the tensor path is integer-exact, so repeated
perform_op calls on the same input must be bit-identical."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashpca_b200.synth import SynthSpec

n, p = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
s = SynthSpec(n, p, seed=20240603)
op = s.create_operator()
rng = np.random.default_rng(0)
x, z = rng.standard_normal(n), rng.standard_normal(n)
ys = [op.perform_op(x) for _ in range(reps)]
zs = [op.perform_op(z) for _ in range(2)]
bad = [i for i in range(1, reps) if not np.array_equal(ys[0], ys[i])]
print("FPB_GEMV=%s n=%d p=%d: nondeterministic reps: %s" % (os.environ.get("FPB_GEMV"), n, p, bad))
for i in bad:
    d = np.abs(ys[i] - ys[0])
    print("  rep %d: max abs diff %.3e (scale %.3e), #diff %d" % (i, d.max(), np.abs(ys[0]).max(), (d > 0).sum()))
lin = op.perform_op(2 * x - 3 * z)
print("  linearity err / scale: %.3e" % (np.abs(lin - (2 * ys[0] - 3 * zs[0])).max() / np.abs(ys[0]).max()))
t = op.crossprod(x)
print("  prod(crossprod) vs op: %.3e" % (np.abs(op.prod(t) - ys[0]).max() / np.abs(ys[0]).max()))
