"""Regenerates tests/golden/golden_pca.json: dense float64 numpy eigh of
X X'/p (binom2, missing -> 0) on the reference's bundled bed fixtures, copied
from /root/reference/HapMap3/data.* and flashpcaR/inst/extdata/data_chr1.*.
The reference holds no stored eigen-results; its tests
compute this dense decomposition at run time (test_pca.R:47,70)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import load_fixture  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = {}
for name in ("data_chr1", "hapmap3"):
    _, payload, n, p = load_fixture(name)
    x, msd = O.dense_standardise(O.dense_codes(payload, n, p))
    r = O.dense_pca(x, 20)
    out[name] = dict(n=n, p=p, eigenvalues=[float(v) for v in r["d"]], trace_over_p=r["trace"],
                     pve1=float(r["pve"][0]), mean_first3=[float(v) for v in msd[:3, 0]],
                     sd_first3=[float(v) for v in msd[:3, 1]])
json.dump(out, open(os.path.join(HERE, "golden_pca.json"), "w"), indent=1)
