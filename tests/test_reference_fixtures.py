"""Pin the oracle's genotype decode against matrices the REFERENCE itself holds.

The reference keeps two R workspaces with the dosage matrices its own tests start from:
  * flashpcaR/data/hm3.chr1.rda  (`hm3.chr1$bed`, test_pca.R:8: `data(hm3.chr1); X <- hm3.chr1$bed`,
    compared by test_pca.R:45-105 with the PLINK fileset inst/extdata/data_chr1 at :46);
  * HapMap3/data.RData (`hapmap3$bed`, the matrix behind HapMap3/data.{bed,bim,fam}).
They were written by R (plink2R::read_plink), not by this repository, so agreement pins
decode_plink / the code -> dosage table (data.cpp:36-45, 65-126) of oracle/ to reference-held
data: rows a3/a5 of SURVEY section 8.  tests/rdata.py is a stdlib reader for R's XDR format.
The solver (Spectra) has no such vector anywhere upstream and stays unpinned.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import rdata
from conftest import GOLDEN, load_fixture
from oracle import oracle as O


def _field(lst, name):
    return lst.value[lst.attrs["names"].value.index(name)]


def _col(df, k):
    c = df.value[k]
    return c.value if isinstance(c, rdata.RObj) else c


@pytest.fixture(scope="module")
def hm3():
    return rdata.read_rdata(os.path.join(GOLDEN, "data_chr1", "hm3.chr1.rda"))["hm3.chr1"]


def test_rda_dosages_equal_oracle_decode(hm3):
    """hm3.chr1$bed (957 x 1129 doubles, 1660 NA) == oracle decode of data_chr1.bed, exactly."""
    x = rdata.as_matrix(_field(hm3, "bed"))
    stem, payload, n, p = load_fixture("data_chr1")
    d = O.dosage_matrix(O.dense_codes(payload, n, p))
    assert x.shape == d.shape == (957, 1129)
    assert int(np.isnan(x).sum()) == 1660
    assert np.array_equal(np.isnan(x), np.isnan(d))
    assert np.array_equal(np.nan_to_num(x, nan=-1.0), np.nan_to_num(d, nan=-1.0))


def test_rda_dosages_equal_c_oracle_lookup(hm3):
    """The C restatement of read_snp_block (data.cpp:215-335) standardises exactly the dosages
    the reference's matrix holds: lookup[code] == (dosage - mean) / sd with the oracle's mean/sd,
    and mean/sd equal the column statistics of hm3.chr1$bed (test_pca.R:13-22 via scale2)."""
    x = rdata.as_matrix(_field(hm3, "bed"))
    stem, payload, n, p = load_fixture("data_chr1")
    orc = O.COracle(payload, n, p)
    orc.crossprod(np.ones(n), 300)      # statistics are filled on the first visit (data.cpp:257)
    msd = orc.meansd()
    lut = orc.lookup()                  # 4 x P, indexed by raw PLINK code (data.cpp:316-319)
    mean = np.nanmean(x, axis=0)
    pj = mean / 2.0
    assert np.allclose(msd[:, 0], mean, rtol=0, atol=1e-14)
    assert np.allclose(msd[:, 1], np.sqrt(2.0 * pj * (1.0 - pj)), rtol=1e-14, atol=0)
    codes = O.dense_codes(payload, n, p)
    xs = lut[codes.astype(np.int64), np.arange(p)[None, :]]   # X(i, j) = lookup(code_ij, j), data.cpp:330-333
    ref = (x - msd[:, 0]) / msd[:, 1]
    ref[np.isnan(ref)] = 0.0             # missing -> mean -> 0 after centring (data.cpp:319)
    assert np.allclose(xs, ref, rtol=0, atol=1e-13)


def test_rda_ids_match_fam_and_bim(hm3):
    """FID/IID order (fam) and SNP / allele order (bim) of the PLINK fileset equal the ids stored in
    the workspace (HapMap3/test_pca.R:93-106 asserts the same orders on the CLI outputs)."""
    stem, payload, n, p = load_fixture("data_chr1")
    fid, iid = O.read_fam_ids(stem + ".fam")
    fam, bim = _field(hm3, "fam"), _field(hm3, "bim")
    assert list(_col(fam, 0)) == fid and list(_col(fam, 1)) == iid
    rows, cols = rdata.dimnames(_field(hm3, "bed"))
    assert rows == ["%s:%s" % (a, b) for a, b in zip(fid, iid)]   # flashpcaR rownames FID:IID
    snps, a1 = [], []
    with open(stem + ".bim") as f:
        for line in f:
            t = line.split()
            snps.append(t[1])
            a1.append(t[4])
    assert cols == snps == list(_col(bim, 1))
    assert a1 == list(_col(bim, 4))


def test_hapmap3_rdata_digest_matches_oracle_decode():
    """hapmap3$bed of HapMap3/data.RData against the oracle decode of HapMap3/data.bed through the
    committed digest (tests/golden/make_rdata_digest.py; upstream imputed the bed's 21,221 missing
    genotypes before saving, every other entry must agree)."""
    with open(os.path.join(GOLDEN, "hapmap3_rdata_digest.json")) as f:
        g = json.load(f)
    stem, payload, n, p = load_fixture("hapmap3")
    codes = O.dense_codes(payload, n, p)
    d = O.dosage_matrix(codes)
    missing = np.isnan(d)
    assert g["shape"] == [n, p] and g["na_count_rdata"] == 0
    assert int(missing.sum()) == g["missing_in_bed"] == 21221
    assert sum(g["imputed_histogram"].values()) == g["missing_in_bed"]
    assert set(g["imputed_histogram"]) <= {"0", "1", "2"}
    canon = np.where(missing, -1.0, d).astype(np.float64)
    sha = hashlib.sha256(np.asfortranarray(canon).tobytes(order="F")).hexdigest()
    assert sha == g["sha256_masked"]
    assert [int(v) for v in np.where(missing, 0.0, d).sum(axis=0)] == g["col_sums_masked"]
    fid, iid = O.read_fam_ids(stem + ".fam")
    rows = ["%s:%s" % (a, b) for a, b in zip(fid, iid)]
    assert hashlib.sha256("\n".join(rows).encode()).hexdigest() == g["rownames_sha256"]
    snps = [line.split()[1] for line in open(stem + ".bim")]
    assert hashlib.sha256("\n".join(snps).encode()).hexdigest() == g["colnames_sha256"]


@pytest.mark.skipif(not os.path.exists("/root/reference/HapMap3/data.RData"),
                    reason="the reference tree is only present in the build container")
def test_hapmap3_rdata_direct():
    """Same check against the workspace itself where the reference tree is available."""
    obj = rdata.read_rdata("/root/reference/HapMap3/data.RData")["hapmap3"]
    x = rdata.as_matrix(_field(obj, "bed"))
    stem, payload, n, p = load_fixture("hapmap3")
    d = O.dosage_matrix(O.dense_codes(payload, n, p))
    m = np.isnan(d)
    assert np.array_equal(x[~m], d[~m])
    assert set(np.unique(x[m])) <= {0.0, 1.0, 2.0}


def test_overlap_fixture_shapes():
    """HapMap3/test_pca.R:40-76 filesets: N, P, missing rate."""
    for sub, stem, n_exp, p_exp, size in (
            ("hm3_overlap", "HM3_thinned_autosomal_overlap", 957, 14079, 3378963),
            ("kg1_overlap", "1kg.ref.phase1_release_v3.20101123_thinned_autosomal_overlap", 1092, 14079,
             3843570)):
        path = os.path.join(GOLDEN, sub, stem)
        assert os.path.getsize(path + ".bed") == size
        n = O.count_lines(path + ".fam")
        payload, _, p = O.read_bed_payload(path + ".bed", n)
        assert (n, p) == (n_exp, p_exp)
        codes = O.dense_codes(payload, n, p)
        if sub == "kg1_overlap":
            assert int((codes == 1).sum()) == 0      # N multiple of 4, no missing genotypes
