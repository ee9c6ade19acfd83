"""Digest of the dosage matrix the reference keeps in HapMap3/data.RData (`hapmap3$bed`,
957 x 14,389, the matrix HapMap3/test_pca.R and test_cca.R start from).  The .RData file is
5 MB, so the repository holds this digest instead of a copy; run in the build container:

    python tests/golden/make_rdata_digest.py /root/reference/HapMap3/data.RData

`hapmap3$bed` holds no NA: upstream filled the 21,221 genotypes that are missing in
HapMap3/data.bed with imputed 0/1/2 values before saving.  Everywhere else the two must agree,
so the digest is taken with the bed's missing positions masked (set to -1):
tests/golden/hapmap3_rdata_digest.json = shape, SHA-256 of the masked column-major float64
bytes, per-SNP sums of the masked matrix, the histogram of the imputed values, and the row /
column names.  tests/test_reference_fixtures.py checks the oracle's decode of
tests/golden/hapmap3/data.bed (a byte copy of HapMap3/data.bed) against it.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import rdata  # noqa: E402


def masked_sha256(x: np.ndarray, missing: np.ndarray) -> str:
    canon = np.where(missing, -1.0, x).astype(np.float64)
    return hashlib.sha256(np.asfortranarray(canon).tobytes(order="F")).hexdigest()


def digest(x: np.ndarray, missing: np.ndarray, rows, cols) -> dict:
    vals, counts = np.unique(x[missing], return_counts=True)
    return {
        "shape": list(x.shape),
        "na_count_rdata": int(np.isnan(x).sum()),
        "missing_in_bed": int(missing.sum()),
        "imputed_histogram": {str(int(v)): int(c) for v, c in zip(vals, counts)},
        "sha256_masked": masked_sha256(x, missing),
        "col_sums_masked": [int(v) for v in np.where(missing, 0.0, x).sum(axis=0)],
        "rownames_sha256": hashlib.sha256("\n".join(rows).encode()).hexdigest(),
        "colnames_sha256": hashlib.sha256("\n".join(cols).encode()).hexdigest(),
        "rownames_first": rows[:3],
        "colnames_first": cols[:3],
    }


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/HapMap3/data.RData"
    obj = rdata.read_rdata(src)["hapmap3"]
    names = obj.attrs["names"].value
    bed = obj.value[names.index("bed")]
    x = rdata.as_matrix(bed)
    rn, cn = rdata.dimnames(bed)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import oracle as O
    stem = os.path.join(HERE, "hapmap3", "data")
    n = O.count_lines(stem + ".fam")
    payload, _, p = O.read_bed_payload(stem + ".bed", n)
    missing = O.dense_codes(payload, n, p) == 1   # PLINK code 01 = missing (data.cpp:36-45)
    out = digest(x, missing, rn, cn)
    out["source"] = "HapMap3/data.RData::hapmap3$bed"
    with open(os.path.join(HERE, "hapmap3_rdata_digest.json"), "w") as f:
        json.dump(out, f)
    print(out["shape"], out["missing_in_bed"], out["imputed_histogram"], out["sha256_masked"])
