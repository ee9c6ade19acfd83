"""CPU tests: pin the oracle (oracle/) against the reference's own criterion
(dense eigendecomposition on its bundled fixtures, test_pca.R:45-70) and the
session-probe constants recorded in SURVEY.md section 8c / tests/golden."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_fixture
from oracle import oracle as O


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "golden_pca.json")) as f:
        return json.load(f)


def test_decode_plink_matches_reference_table():
    # data.cpp:36-45: 00 -> 2 copies, 10 -> 1, 11 -> 0, 01 -> NA(3); LSB-first within a byte
    import ctypes
    lib = ctypes.CDLL(O.build())
    allbytes = np.arange(256, dtype=np.uint8)
    out = np.zeros(1024, dtype=np.uint8)
    lib.fo_decode_plink(out.ctypes.data, allbytes.ctypes.data, 256)
    table = {0: 2, 1: 3, 2: 1, 3: 0}
    for b in range(256):
        for q in range(4):
            assert out[4 * b + q] == table[(b >> (2 * q)) & 3]
    lib.fo_decode_plink_simple(out.ctypes.data, allbytes.ctypes.data, 256)
    for b in (0, 0x1B, 0xE4, 0xFF):
        assert [int(v) for v in out[4 * b:4 * b + 4]] == [(b >> (2 * q)) & 3 for q in range(4)]


def test_fixture_shapes():
    # SURVEY.md section 4 fixture table
    _, payload, n, p = load_fixture("hapmap3")
    assert (n, p, payload.size) == (957, 14389, 3453363 - 3)
    _, payload, n, p = load_fixture("data_chr1")
    assert (n, p, payload.size) == (957, 1129, 270963 - 3)


@pytest.mark.parametrize("name", ["data_chr1", "hapmap3"])
def test_survey_constants(name, golden):
    """SURVEY.md section 8c probe numbers (independent numpy script) and the
    committed golden file agree with the oracle's dense path."""
    _, payload, n, p = load_fixture(name)
    codes = O.dense_codes(payload, n, p)
    x, msd = O.dense_standardise(codes)
    g = golden[name]
    assert np.allclose(msd[:3, 0], g["mean_first3"], rtol=0, atol=1e-11)
    assert np.allclose(msd[:3, 1], g["sd_first3"], rtol=0, atol=1e-11)
    ndim = len(g["eigenvalues"])
    res = O.dense_pca(x, ndim)
    assert np.allclose(res["d"], g["eigenvalues"], rtol=1e-10)
    assert abs(res["trace"] - g["trace_over_p"]) < 1e-9 * g["trace_over_p"]
    assert abs(res["pve"][0] - g["pve1"]) < 1e-10
    if name == "hapmap3":  # literal values printed in SURVEY.md section 8c
        assert abs(res["d"][0] - 26.467988137205) < 1e-9
        assert abs(res["d"][9] - 2.218548237411) < 1e-9
        assert abs(res["trace"] - 990.4296132830024) < 1e-7
    else:
        assert abs(res["d"][0] - 28.011938222135) < 1e-9
        assert abs(res["trace"] - 987.2553072389829) < 1e-7


@pytest.mark.parametrize("block_size", [0, 1, 300, 1129])
def test_c_oracle_operator_vs_dense(block_size):
    """flashpca_oracle.c (blocked read_snp_block + two GEMVs) against dense numpy,
    over the block sizes the reference would pick (one block; many blocks; B=1)."""
    _, payload, n, p = load_fixture("data_chr1")
    if block_size == 1:
        p = 64
        payload = payload[: p * ((n + 3) // 4)]
    codes = O.dense_codes(payload, n, p)
    x, msd = O.dense_standardise(codes)
    orc = O.COracle(payload, n, p)
    rng = np.random.default_rng(1)
    v = rng.standard_normal(n)
    y = orc.perform_op(v, block_size)
    yd = x @ (x.T @ v)
    assert np.abs(y - yd).max() <= 1e-12 * np.abs(yd).max()
    assert np.array_equal(orc.meansd(), msd)  # bit-exact statistics
    assert abs(orc.trace - np.sum(x * x)) <= 1e-12 * np.sum(x * x)
    m = rng.standard_normal((n, 3))
    assert np.allclose(orc.crossprod(m, block_size), x.T @ m, rtol=0, atol=1e-10)
    w = rng.standard_normal((p, 2))
    assert np.allclose(orc.prod(w, block_size), x @ w, rtol=0, atol=1e-9)
    lk = orc.lookup()
    assert np.all(lk[1] == 0.0)  # missing -> 0 (data.cpp:319)


def test_c_oracle_binom_and_preloaded():
    _, payload, n, p = load_fixture("data_chr1")
    codes = O.dense_codes(payload, n, p)
    x1, msd1 = O.dense_standardise(codes, O.STANDARDISE_BINOM)
    orc = O.COracle(payload, n, p, O.STANDARDISE_BINOM)
    v = np.random.default_rng(2).standard_normal(n)
    y = orc.perform_op(v, 200)
    assert np.abs(y - x1 @ (x1.T @ v)).max() <= 1e-12 * np.abs(y).max()
    # binom sd = binom2 sd / sqrt(2)
    _, msd2 = O.dense_standardise(codes, O.STANDARDISE_BINOM2)
    assert np.allclose(msd1[:, 1] * np.sqrt(2.0), msd2[:, 1], rtol=1e-14)
    # preloaded mean/sd (data.cpp:293-297), maf2meansd's sd = 2p(1-p) (randompca.cpp:745-751)
    maf = msd2[:, 0] / 2.0
    pre = np.stack([2 * maf, 2 * maf * (1 - maf)], axis=1)
    xp, _ = O.dense_standardise(codes, meansd=pre)
    orc2 = O.COracle(payload, n, p, meansd=pre)
    y2 = orc2.perform_op(v, 0)
    assert np.abs(y2 - xp @ (xp.T @ v)).max() <= 1e-12 * np.abs(y2).max()


def test_monomorphic_and_all_missing_columns():
    """sd <= VAR_TOL -> all-zero column (data.cpp:300); untested upstream."""
    n, p = 37, 6
    codes = np.full((n, p), 3, dtype=np.uint8)        # all major homozygous: monomorphic
    codes[:, 1] = 1                                   # all missing -> mean NaN
    rng = np.random.default_rng(3)
    codes[:, 2:] = rng.choice([0, 2, 3, 1], size=(n, p - 2), p=[0.2, 0.3, 0.45, 0.05])
    npb = (n + 3) // 4
    padded = np.zeros((p, npb * 4), dtype=np.uint8)
    padded[:, :n] = codes.T
    q = padded.reshape(p, npb, 4)
    payload = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8).ravel()
    orc = O.COracle(payload, n, p)
    v = rng.standard_normal(n)
    y = orc.perform_op(v, 4)
    with np.errstate(all="ignore"):
        x, _ = O.dense_standardise(codes)
    assert np.all(x[:, 0] == 0) and np.all(x[:, 1] == 0)
    assert np.allclose(y, x @ (x.T @ v), rtol=0, atol=1e-12)
    assert np.isfinite(y).all()


def test_block_size_formula():
    """flashpca.cpp:636-686 at the SURVEY.md section 8 configurations."""
    orc = O.COracle(np.zeros(4, dtype=np.uint8), 4, 1)
    lib = orc.lib
    assert lib.fo_block_size_from_memory(957, 14389, 10, 0, 2048) == 14389
    assert lib.fo_block_size_from_memory(10000, 100000, 20, 0, 2048) == 26286
    assert lib.fo_block_size_from_memory(500000, 100000, 20, 0, 2048) == 465
    assert lib.fo_block_size_from_memory(1000000, 500000, 20, 0, 2048) == 183
    assert lib.fo_block_size_from_memory(500000, 100000, 20, 0, 1) == 0


def test_simple_random_matches_park_miller():
    v = O.simple_random_vec(3, 0)
    # seed 0 -> 1; 16807^k mod (2^31-1)
    assert np.allclose(v, np.array([16807, 282475249, 1622650073]) / 2147483647.0 - 0.5)


@pytest.mark.parametrize("name,ndim", [("data_chr1", 10), ("hapmap3", 10)])
def test_irlm_restatement_vs_dense(name, ndim, golden):
    """Spectra restatement on the C operator converges to the dense eigenpairs
    (HapMap3/test_pca.R:121-246 criterion, tol 1e-6 -> eigenvalues << 1e-6)."""
    _, payload, n, p = load_fixture(name)
    res = O.oracle_pca(payload, n, p, ndim, tol=1e-6, block_size=500)
    g = golden[name]
    assert np.allclose(res["d"], g["eigenvalues"][:ndim], rtol=1e-9)
    assert abs(res["trace"] - g["trace_over_p"]) < 1e-9 * g["trace_over_p"]
    assert np.allclose(res["pve"], res["d"] / res["trace"])
    codes = O.dense_codes(payload, n, p)
    x, _ = O.dense_standardise(codes)
    dres = O.dense_pca(x, ndim)
    u = O.sign_align(res["U"], dres["U"])
    assert np.abs(u - dres["U"]).max() < 5e-6
    assert 1 + 2 * ndim <= res["nops"] <= 400


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
def test_standardise_matrix_restatement(method):
    """util.cpp:24-192 restatement vs the formulas flashpcaR's tests use
    (test_standardisation.R:15-86: scale()/scale2() with NA -> 0)."""
    _, payload, n, p = load_fixture("data_chr1")
    x = O.dosage_matrix(O.dense_codes(payload, n, p))[:, :200]
    s, msd = O.standardise_matrix(x, method)
    mu = np.nanmean(x, axis=0)
    assert np.allclose(msd[:, 0], mu, rtol=1e-13)
    if method == 0:
        assert np.allclose(s, np.where(np.isnan(x), mu, x))
    elif method == 4:
        assert np.allclose(s, np.where(np.isnan(x), 0, x - mu))
    else:
        if method == 1:
            sd = np.nanstd(x, axis=0, ddof=1)
        else:
            sd = np.sqrt((1 if method == 2 else 2) * (mu / 2) * (1 - mu / 2))
        assert np.allclose(msd[:, 1], sd, rtol=1e-12)
        assert np.allclose(s, np.where(np.isnan(x), 0, (x - mu) / sd), rtol=1e-11, atol=1e-12)
    if method == 3:   # the matrix path and the bed path standardise identically
        xb, _ = O.dense_standardise(O.dense_codes(payload, n, p)[:, :200])
        assert np.allclose(s, xb, rtol=1e-13, atol=1e-13)


def test_host_synth_generator_matches_numpy_generator():
    """The C generator bench.py uses for the CPU baseline's input equals synth.py bit for bit
    (which the device generator equals too, test_gpu_parity.py)."""
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(1003, 300, seed=5)
    assert np.array_equal(O.synth_packed_bed(s, 17, 211), s.packed_bed(17, 211))
    assert np.array_equal(O.synth_packed_bed(s, 0, 300), s.packed_bed())
