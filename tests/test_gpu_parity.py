"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs (bit-exact for the integer-derived statistics; the
floating-point tolerances are written next to each assertion)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import FIXTURES, load_fixture
from oracle import oracle as O

pytestmark = pytest.mark.gpu

# y = X X' x: both sides sum ~N*P doubles in different orders (SURVEY.md section 8e:
# "expect ~1e-15 relative differences"); bound used throughout:
OP_RTOL = 1e-12


def _mk(payload, n, p, **kw):
    from flashpca_b200 import SVDWideOnline
    return SVDWideOnline(payload=payload, n=n, nsnps=p, **kw)


@pytest.fixture(params=["imma", "generic"])
def path(request, monkeypatch):
    """Both compute paths of the library: the int8 tensor-core path and the
    generic FP64 path (FPB_PATH is read when an operator is created)."""
    monkeypatch.setenv("FPB_PATH", request.param)
    return request.param


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _pack(codes):
    """(n, p) raw codes -> packed payload with zero pad bits."""
    n, p = codes.shape
    npb = (n + 3) // 4
    padded = np.zeros((p, npb * 4), dtype=np.uint8)
    padded[:, :n] = codes.T
    q = padded.reshape(p, npb, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(
        np.uint8).ravel()


@pytest.mark.parametrize("name", ["data_chr1", "hapmap3"])
@pytest.mark.parametrize("stand", [O.STANDARDISE_BINOM2, O.STANDARDISE_BINOM])
def test_fixture_operator_family(native_lib, path, name, stand):
    _, payload, n, p = load_fixture(name)
    op = _mk(payload, n, p, stand_method=stand)
    orc = O.COracle(payload, n, p, stand)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(n)
    y_ref = orc.perform_op(x, 0)            # also fills oracle statistics
    assert np.array_equal(op.meansd(), orc.meansd())          # bit-exact mean / sd
    assert abs(op.trace - orc.trace) <= 1e-12 * orc.trace
    assert (op.rows(), op.cols()) == (n, n)
    assert _relerr(op.perform_op(x), y_ref) <= OP_RTOL
    assert _relerr(op.crossprod(x), orc.crossprod(x)) <= OP_RTOL
    v = rng.standard_normal(p)
    assert _relerr(op.prod(v), orc.prod(v)) <= OP_RTOL
    m = rng.standard_normal((n, 3))
    assert _relerr(op.perform_op_mat(m), orc.perform_op(m, 0)) <= OP_RTOL
    assert _relerr(op.crossprod2(m), orc.crossprod(m, 0)) <= OP_RTOL
    w = rng.standard_normal((p, 2))
    assert _relerr(op.prod3(w), orc.prod(w, 0)) <= OP_RTOL
    # the staged bytes are the file's bytes for every real individual
    assert np.array_equal(O.dense_codes(op.bed_payload(), n, p), O.dense_codes(payload, n, p))
    op.close()


def test_reference_block_order_vs_one_pass(native_lib):
    """Upstream accumulates block by block (svdwide.cpp:48-59); the GPU sums in
    one pass.  Same result within OP_RTOL for the reference's block sizes."""
    _, payload, n, p = load_fixture("data_chr1")
    op = _mk(payload, n, p)
    x = np.random.default_rng(5).standard_normal(n)
    y = op.perform_op(x)
    for bs in (1129, 302, 37):
        assert _relerr(y, O.COracle(payload, n, p).perform_op(x, bs)) <= OP_RTOL


@pytest.mark.parametrize("n,p", [(1, 3), (3, 1), (4, 5), (63, 9), (64, 64), (65, 130), (1000, 17),
                                 (4099, 257), (16384 + 5, 40)])
def test_ragged_shapes(native_lib, path, n, p):
    """N not a multiple of 4/16/64 (pad genotypes must contribute 0), tiny and
    ragged shapes, heavy missingness."""
    rng = np.random.default_rng(n * 1000 + p)
    codes = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(n, p),
                       p=[0.15, 0.1, 0.35, 0.4])
    payload = _pack(codes)
    # dirty pad bits: upstream ignores whatever sits there (it loops i < N)
    if n % 4:
        npb = (n + 3) // 4
        pl = payload.reshape(p, npb)
        pl[:, -1] |= np.uint8((0xFF << (2 * (n % 4))) & 0xFF)
    op = _mk(payload, n, p)
    orc = O.COracle(payload, n, p)
    x = rng.standard_normal(n)
    y_ref = orc.perform_op(x, 0)
    got, want = op.meansd(), orc.meansd()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    y = op.perform_op(x)
    assert np.isfinite(y).all()
    assert np.abs(y - y_ref).max() <= OP_RTOL * max(np.abs(y_ref).max(), 1.0)
    v = rng.standard_normal(p)
    assert np.abs(op.prod(v) - orc.prod(v)).max() <= OP_RTOL * max(np.abs(v).max() * p, 1.0)
    assert np.abs(op.crossprod(x) - orc.crossprod(x)).max() <= OP_RTOL * n
    op.close()


def test_monomorphic_and_all_missing_columns(native_lib, path):
    n, p = 37, 6
    rng = np.random.default_rng(3)
    codes = np.full((n, p), 3, dtype=np.uint8)
    codes[:, 1] = 1
    codes[:, 2:] = rng.choice([0, 2, 3, 1], size=(n, p - 2), p=[0.2, 0.3, 0.45, 0.05])
    payload = _pack(codes)
    op = _mk(payload, n, p)
    orc = O.COracle(payload, n, p)
    x = rng.standard_normal(n)
    y_ref = orc.perform_op(x, 0)
    y = op.perform_op(x)
    assert np.isfinite(y).all() and np.allclose(y, y_ref, rtol=0, atol=1e-12)
    t = op.crossprod(x)
    assert t[0] == 0.0 and t[1] == 0.0     # sd <= VAR_TOL / NaN -> zero column (data.cpp:300)
    assert abs(op.trace - orc.trace) <= 1e-12 * orc.trace


def test_preloaded_meansd(native_lib, path):
    """Data::use_preloaded_maf (data.cpp:293-297) with maf2meansd's sd
    (randompca.cpp:745-751: 2p(1-p), no sqrt)."""
    _, payload, n, p = load_fixture("data_chr1")
    base = O.COracle(payload, n, p)
    base.perform_op(np.ones(n), 0)
    maf = base.meansd()[:, 0] / 2.0
    pre = np.asfortranarray(np.stack([2 * maf, 2 * maf * (1 - maf)], axis=1))
    op = _mk(payload, n, p, meansd=pre)
    orc = O.COracle(payload, n, p, meansd=pre)
    x = np.random.default_rng(9).standard_normal(n)
    assert _relerr(op.perform_op(x), orc.perform_op(x, 0)) <= OP_RTOL
    assert np.array_equal(op.meansd(), pre)


def test_create_from_file_and_shards(native_lib, path):
    from flashpca_b200 import Data, SVDWideOnline
    stem = FIXTURES["data_chr1"]
    d = Data()
    d.read_pheno(stem + ".fam", 6)
    d.geno_filename = stem + ".bed"
    d.get_size()
    d.prepare()
    op = SVDWideOnline(d, 0, 3)
    _, payload, n, p = load_fixture("data_chr1")
    orc = O.COracle(payload, n, p)
    x = np.random.default_rng(4).standard_normal(n)
    y_ref = orc.perform_op(x, 0)
    assert _relerr(op.perform_op(x), y_ref) <= OP_RTOL
    assert np.array_equal(d.X_meansd, orc.meansd())
    # two SNP shards sum to the whole (svdwide.cpp:48-59 distributed)
    a = SVDWideOnline(d, 0, 3, snp_begin=0, snp_count=500)
    b = SVDWideOnline(d, 0, 3, snp_begin=500, snp_count=0)
    assert a.p + b.p == p
    assert _relerr(a.perform_op(x) + b.perform_op(x), y_ref) <= OP_RTOL
    assert abs(a.trace + b.trace - orc.trace) <= 1e-12 * orc.trace
    assert np.array_equal(np.concatenate([a.meansd(), b.meansd()]), orc.meansd())


def test_tensor_path_is_deterministic_and_default(native_lib, monkeypatch):
    """Low-missingness data takes the tensor path by default; its integer
    accumulation makes repeated calls bit-identical, and it agrees with the
    generic FP64 path to OP_RTOL."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    _, payload, n, p = load_fixture("hapmap3")
    op = _mk(payload, n, p)
    x = np.random.default_rng(21).standard_normal(n)
    y1, y2 = op.perform_op(x), op.perform_op(x)
    assert np.array_equal(y1, y2)
    monkeypatch.setenv("FPB_PATH", "generic")
    og = _mk(payload, n, p)
    assert _relerr(y1, og.perform_op(x)) <= OP_RTOL
    # scale invariance of the fixed-point conversion: tiny, huge and zero vectors
    for sc in (1e-300, 1e-30, 1e30, 1e250):
        assert _relerr(op.perform_op(x * sc) / sc, y1) <= OP_RTOL
    assert np.all(op.perform_op(np.zeros(n)) == 0.0)
    e = np.zeros(n)
    e[17] = 1.0
    orc = O.COracle(payload, n, p)
    assert _relerr(op.perform_op(e), orc.perform_op(e, 0)) <= OP_RTOL
    assert np.isnan(op.perform_op(np.full(n, np.nan))).all()


@pytest.mark.parametrize("variant", ["tma", "tma2", "ldg", "wide", "narrow", "gather_sms", "l2_keep_off"])
def test_contraction_kernel_variants_agree(native_lib, monkeypatch, variant):
    """FPB_GEMV selects the contraction kernels: default single-copy TMA pipeline
    (k_imma_gemv_tma + k_imma_gemv_tma_t), two-copy TMA (tma2) and the register-staged
    LDG kernel (ldg).  All are exact integer contractions per split; the FP64
    recombination of splits rounds differently, so they agree to ~1e-15, not bit
    for bit; each variant on its own is bit-reproducible."""
    _, payload, n, p = load_fixture("hapmap3")
    monkeypatch.delenv("FPB_PATH", raising=False)
    monkeypatch.delenv("FPB_GEMV", raising=False)
    base = _mk(payload, n, p)
    rng = np.random.default_rng(31)
    x, v = rng.standard_normal(n), rng.standard_normal(p)
    y0, t0, z0 = base.perform_op(x), base.crossprod(x), base.prod(v)
    if variant == "wide":       # second half with 256-byte stripes (k_imma_gemv_tma_tw)
        monkeypatch.setenv("FPB_P2WIDE", "1")
        monkeypatch.setenv("FPB_PERSIST", "0")
    elif variant == "narrow":   # one-shot 128-byte stripe grids
        monkeypatch.setenv("FPB_P2WIDE", "0")
        monkeypatch.setenv("FPB_PERSIST", "0")
    elif variant == "gather_sms":   # missing-genotype gathers on 3 dedicated SMs (k_sell_gather_p)
        monkeypatch.setenv("FPB_GATHER_SMS", "3")
    elif variant == "l2_keep_off":  # default: the first half leaves its last SNP rows in L2 for the
        monkeypatch.setenv("FPB_L2_KEEP_MB", "0")   # second (evict_last) when the matrix is small
    else:
        monkeypatch.setenv("FPB_GEMV", variant)
    op = _mk(payload, n, p)
    y1 = op.perform_op(x)
    assert _relerr(y1, y0) <= 1e-13
    if variant in ("gather_sms", "l2_keep_off"):   # same kernels' arithmetic, same summation order
        assert np.array_equal(y1, y0)
    assert _relerr(op.crossprod(x), t0) <= 1e-13
    assert _relerr(op.prod(v), z0) <= 1e-13
    assert np.array_equal(op.perform_op(x), y1)
    orc = O.COracle(payload, n, p)
    assert _relerr(y0, orc.perform_op(x, 0)) <= OP_RTOL


def test_error_paths(native_lib):
    from flashpca_b200 import Data, SVDWideOnline
    from flashpca_b200._lib import FpbError
    _, payload, n, p = load_fixture("data_chr1")
    with pytest.raises(FpbError, match="unknown standardisation method: 1"):
        _mk(payload, n, p, stand_method=1)      # data.cpp:283-288
    d = Data()
    d.N = 10
    d.geno_filename = "/nonexistent/file.bed"
    with pytest.raises(FpbError, match="Error reading file"):
        SVDWideOnline(d, 0, 3)
    op = _mk(payload, n, p)
    with pytest.raises(FpbError, match="dimension mismatch"):
        op.perform_op(np.zeros(n + 1))
    with pytest.raises(FpbError, match="invalid nev/ncv"):
        op.pca(10, 10, 5, 1e-6)


def test_device_synth_matches_host_generator(native_lib):
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(1003, 300, seed=99, missing_rate=0.01)
    op = s.create_operator()
    got = op.bed_payload()
    want = s.packed_bed()
    assert np.array_equal(O.dense_codes(got, s.n, s.p), O.dense_codes(want, s.n, s.p))
    sh = s.create_operator(j0=100, j1=250)
    assert np.array_equal(O.dense_codes(sh.bed_payload(), s.n, 150),
                          O.dense_codes(s.packed_bed(100, 250), s.n, 150))


@pytest.mark.parametrize("name,ndim", [("data_chr1", 10), ("hapmap3", 10), ("hapmap3", 20)])
def test_pca_vs_dense_eigh(native_lib, name, ndim):
    """BASELINE config 0 (bundled HapMap3/data, --ndim 10) and the R-test fixture:
    eigenvalues / pve within 1e-6 relative of dense eigh at the reference's
    default tol 1e-6; eigenvectors and PCs within 1e-6 (sign-aligned) when the
    solver is converged tighter (vector error ~ residual/gap)."""
    from flashpca_b200 import Data, RandomPCA
    stem = FIXTURES[name]
    d = Data()
    d.read_pheno(stem + ".fam", 6)
    d.read_plink_fam(stem + ".fam")
    d.geno_filename = stem + ".bed"
    d.get_size()
    d.prepare()
    _, payload, n, p = load_fixture(name)
    x, msd = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, ndim)
    r = RandomPCA()
    r.pca_fast(d, 0, ndim, 500, 1e-6, 1, do_loadings=True)
    assert np.abs(r.d / ref["d"] - 1).max() < 1e-6
    assert np.abs(r.pve / ref["pve"] - 1).max() < 1e-6
    assert abs(r.trace / ref["trace"] - 1) < 1e-12
    assert np.array_equal(r.X_meansd, msd)
    assert 1 + 2 * ndim <= r.nops <= 500
    r2 = RandomPCA()
    r2.pca_fast(d, 0, ndim, 500, 1e-10, 1, do_loadings=True)
    u = O.sign_align(r2.U, ref["U"])
    assert np.abs(u - ref["U"]).max() < 1e-6
    px = O.sign_align(r2.Px, ref["Px"])
    assert np.abs(px - ref["Px"]).max() < 1e-6 * np.abs(ref["Px"]).max()
    vref = O.dense_loadings(x, ref["U"], ref["d"], ref["div"])
    v = O.sign_align(r2.V, vref)
    assert np.abs(v - vref).max() < 1e-6 * np.abs(vref).max()
    assert np.abs(np.linalg.norm(r2.U, axis=0) - 1).max() < 1e-12
    # FID / IID ordering is the fam order, bit-exact
    fid, iid = O.read_fam_ids(stem + ".fam")
    assert d.fam_ids == fid and d.indiv_ids == iid
    # --check semantics (randompca.cpp:663-703): mse < 1e-8 as README.md:207 expects
    r2.check(d, 0, r2.U, r2.d)
    assert r2.mse < 1e-8
    # --project semantics (randompca.cpp:798-820): projecting the training data
    # with its own loadings and mean/sd reproduces the PCs (test_project.R:23-26)
    d.X_meansd = r2.X_meansd
    d.use_preloaded_maf = True
    r3 = RandomPCA()
    r3.project(d, 0, r2.V)
    assert np.abs(r3.Px - r2.Px).max() < 1e-5 * np.abs(r2.Px).max()


@pytest.mark.parametrize("name,ndim", [("hapmap3", 100), ("data_chr1", 200), ("data_chr1", 478)])
def test_pca_large_ndim(native_lib, name, ndim):
    """ndim up to the reference's own limit (min(N, P) - 1) / 2 (flashpca.cpp:623-633; 478 on the
    R fixture, test_pca.R:5 uses 50): ncv = 2 ndim + 1 goes up to N = 957 here.  Eigenvalues and pve
    vs dense eigh at the reference's default tol, --check mse on the solver's own eigenpairs."""
    _, payload, n, p = load_fixture(name)
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, ndim)
    op = _mk(payload, n, p)
    res = op.pca(ndim, 2 * ndim + 1, 500, 1e-6)
    assert res["nconv"] == ndim
    d = res["values"] / p
    assert np.abs(d / ref["d"] - 1).max() < 1e-6
    err = op.pca_residual(ndim, float(p))
    assert err.sum() / (n * ndim) < 1e-8
    u = res["vectors"]
    assert np.abs(u.T @ u - np.eye(ndim)).max() < 1e-10
    # the device check equals the host formula of randompca.cpp:663-703
    y = op.perform_op_mat(u) / p
    host_err = ((y - u * d[None, :]) ** 2).sum(axis=0)
    assert np.allclose(err, host_err, rtol=1e-6, atol=1e-18)
    op.close()


@pytest.mark.parametrize("name,ndim", [("data_chr1", 10), ("hapmap3", 10), ("hapmap3", 20)])
def test_block_krylov_solver_vs_dense_eigh(native_lib, monkeypatch, name, ndim):
    """fpb_pca_block (extension): block Lanczos on the tcgen05 block operator, Spectra's convergence
    criterion.  Converged eigenpairs against dense eigh with the tolerances of the single-vector
    solver (eigenvalues 1e-6 at tol 1e-6; vectors 1e-6 sign-aligned at a tighter tol), against the
    Spectra-schedule solver, and through the reference's --check criterion."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    _, payload, n, p = load_fixture(name)
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, ndim)
    op = _mk(payload, n, p)
    got = op.pca_block(ndim, 1e-6)
    assert got["nconv"] == ndim and 3 <= got["npasses"] <= 40
    assert np.abs(got["values"] / p / ref["d"] - 1).max() < 1e-6
    err = op.pca_residual(ndim, float(p))
    assert err.sum() / (n * ndim) < 1e-8
    single = op.pca(ndim, 2 * ndim + 1, 500, 1e-6)
    assert np.abs(got["values"] / single["values"] - 1).max() < 1e-6
    tight = op.pca_block(ndim, 1e-10)
    assert tight["nconv"] == ndim
    u = O.sign_align(tight["vectors"], ref["U"])
    assert np.abs(u - ref["U"]).max() < 1e-6
    assert np.abs(tight["vectors"].T @ tight["vectors"] - np.eye(ndim)).max() < 1e-10
    again = op.pca_block(ndim, 1e-6)
    assert np.array_equal(again["values"], got["values"])      # bit-reproducible
    op.close()


def test_block_krylov_solver_synthetic_and_errors(native_lib, monkeypatch):
    monkeypatch.delenv("FPB_PATH", raising=False)
    from flashpca_b200._lib import FpbError
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(20011, 6000, seed=4, fst=0.05)
    op = s.create_operator()
    got = op.pca_block(20, 1e-6)
    single = op.pca(20, 41, 500, 1e-6)
    assert got["nconv"] == 20
    assert np.abs(got["values"] / single["values"] - 1).max() < 1e-6
    assert got["nops"] < 4 * single["nops"]
    err = op.pca_residual(20, float(s.p))
    assert err.sum() / (s.n * 20) < 1e-8
    with pytest.raises(FpbError, match="invalid nev/block"):
        op.pca_block(20, 1e-6, block=9)
    with pytest.raises(FpbError, match="max_passes too small"):
        op.pca_block(20, 1e-6, block=4, max_passes=2)
    few = op.pca_block(20, 1e-12, max_passes=4)         # runs out of passes: reports what converged
    assert few["nconv"] < 20 and few["npasses"] == 4
    op.close()


def test_solve_parity_at_10k_x_100k(native_lib, monkeypatch):
    """BASELINE configs[1] (synthetic 10,000 x 100,000, k = 20): the full GPU solve against the CPU
    oracle's solve (C restatement of the operator under the oracle's Spectra restatement) on the
    same matrix: eigenvalues 1e-6 relative at the reference's tol, eigenvectors sign-aligned at a
    tighter tol (vector error ~ residual / gap), --check mse < 1e-8 (README.md:207)."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    from flashpca_b200.synth import SynthSpec
    n, p, k = 10000, 100000, 20
    s = SynthSpec(n, p)
    op = s.create_operator()
    payload = op.bed_payload()                      # the matrix the GPU holds, fed to the oracle
    got = op.pca(k, 2 * k + 1, 500, 1e-6)
    assert got["nconv"] == k
    err = op.pca_residual(k, float(p))
    assert err.sum() / (n * k) < 1e-8
    # one CPU solve, converged tightly (about 100-150 oracle ops of 1e9 genotypes each)
    want = O.oracle_pca(payload, n, p, k, tol=1e-8, block_size=4096)
    assert np.abs(got["values"] / p / want["d"] - 1).max() < 1e-6
    tight = op.pca(k, 2 * k + 1, 500, 1e-8)
    u = O.sign_align(tight["vectors"], want["U"])
    assert np.abs(u - want["U"]).max() < 1e-6
    assert np.abs(tight["values"] / p / want["d"] - 1).max() < 1e-9
    op.close()


def test_gpu_solver_vs_oracle_solver_same_tol(native_lib, path):
    """Device IRLM and the oracle's Spectra restatement, same start vector and
    tolerance, on the C oracle operator vs the CUDA operator."""
    _, payload, n, p = load_fixture("data_chr1")
    want = O.oracle_pca(payload, n, p, 10, tol=1e-6)
    op = _mk(payload, n, p)
    got = op.pca(10, 21, 500, 1e-6)
    assert got["nconv"] == 10
    assert np.abs(got["values"] / p / want["d"] - 1).max() < 1e-8
    u = O.sign_align(got["vectors"], want["U"])
    assert np.abs(u - want["U"]).max() < 5e-6
    assert abs(int(got["nops"]) - int(want["nops"])) <= 25


def test_dropin_boundary_with_external_solver(native_lib):
    """The Spectra-shaped boundary: an external IRLM (the oracle's restatement
    standing in for Spectra) drives perform_op(x_in, y_out) with host buffers."""
    _, payload, n, p = load_fixture("data_chr1")
    op = _mk(payload, n, p)
    ybuf = np.zeros(n)

    def cb(x):
        op.perform_op(x, ybuf)
        return ybuf.copy()

    res = O.spectra_irlm(cb, n, 10, 21, 500, 1e-6)
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, 10)
    assert res["nconv"] == 10
    assert np.abs(res["values"] / p / ref["d"] - 1).max() < 1e-6


def test_synthetic_structured_pca_vs_dense(native_lib):
    """Synthetic Balding-Nichols matrix (the bench generator) at an oracle-sized
    shape: k = 20 solve vs dense eigh."""
    from flashpca_b200 import RandomPCA
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(1501, 6000, seed=5, npop=25, fst=0.05)
    op = s.create_operator()
    x, _ = O.dense_standardise(np.ascontiguousarray(s.codes().T))
    ref = O.dense_pca(x, 20)
    r = RandomPCA()
    r.pca_fast(None, 0, 20, 500, 1e-6, op=op)
    assert np.abs(r.d / ref["d"] - 1).max() < 1e-6
    r.pca_fast(None, 0, 20, 500, 1e-11, op=op)
    u = O.sign_align(r.U, ref["U"])
    assert np.abs(u - ref["U"]).max() < 1e-6


@pytest.mark.parametrize("n,p", [(10000, 100000), (500000, 100000)])
def test_full_size_properties(native_lib, n, p):
    """BASELINE configs 1 and 2 at full size, through size-independent
    properties of y = X X' x: linearity, symmetry, positive semi-definiteness,
    consistency of the fused op with its two halves, and the trace identity
    E[z' A z] = trace for Rademacher z (checked loosely)."""
    import torch
    if n * p > 2e10 and torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs > 60 GB HBM")
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(n, p, seed=20240601 + (1 if n == 10000 else 2))
    op = s.create_operator()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n)
    z = rng.standard_normal(n)
    ax, az = op.perform_op(x), op.perform_op(z)
    scale = np.abs(ax).max()
    # linearity: A(2x - 3z) = 2Ax - 3Az
    lin = op.perform_op(2 * x - 3 * z)
    assert np.abs(lin - (2 * ax - 3 * az)).max() <= 1e-11 * scale
    # symmetry: z'Ax = x'Az
    assert abs(z @ ax - x @ az) <= 1e-11 * abs(x @ ax)
    # PSD and consistency with the halves: x'Ax = |X'x|^2, Ax = X (X'x)
    t = op.crossprod(x)
    assert abs(x @ ax - t @ t) <= 1e-11 * (t @ t)
    assert np.abs(op.prod(t) - ax).max() <= 1e-11 * scale
    # a slice of SNPs against the oracle (first 64 SNPs of the same generator)
    sub = s.create_operator(j0=0, j1=64)
    orc = O.COracle(s.packed_bed(0, 64), n, 64)
    assert _relerr(sub.perform_op(x), orc.perform_op(x, 0)) <= OP_RTOL
    assert np.array_equal(sub.meansd(), orc.meansd())
    assert np.array_equal(op.meansd()[:64], orc.meansd())
    # trace: every non-monomorphic SNP contributes about its non-missing count * var ratio
    assert 0.5 * n * p < op.trace < 1.5 * n * p


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
def test_in_memory_matrix_path(native_lib, method):
    """RandomPCA::pca_fast(MatrixXd&) with SVDWide + standardise(), all five
    standardisation methods, against eigen(tcrossprod(S)/ncol(S)) of the
    oracle-standardised matrix: the flashpcaR test (test_pca.R:45-105, tol 1e-4;
    tighter here) on its own fixture."""
    from flashpca_b200 import RandomPCA, SVDWide
    _, payload, n, p = load_fixture("data_chr1")
    x = O.dosage_matrix(O.dense_codes(payload, n, p))
    s_ref, msd_ref = O.standardise_matrix(x, method)
    op = SVDWide(x, method)
    assert np.allclose(op.meansd(), msd_ref, rtol=1e-12, atol=1e-13)
    assert np.abs(op.standardised() - s_ref).max() <= 1e-11 * np.abs(s_ref).max()
    v = np.random.default_rng(method).standard_normal(n)
    y_ref = s_ref @ (s_ref.T @ v)
    assert _relerr(op.perform_op(v), y_ref) <= 1e-11          # svdwide.cpp:10
    assert abs(op.trace / np.sum(s_ref * s_ref) - 1) < 1e-12
    ref = O.dense_pca(s_ref, 10)
    r = RandomPCA()
    r.stand_method_x = method
    r.pca_fast_matrix(x, 10, 500, 1e-8, do_loadings=True)
    assert np.abs(r.d / ref["d"] - 1).max() < 1e-6
    assert np.abs(r.pve / ref["pve"] - 1).max() < 1e-6
    u = O.sign_align(r.U, ref["U"])
    assert np.abs(u - ref["U"]).max() < 1e-6
    vref = O.dense_loadings(s_ref, ref["U"], ref["d"], ref["div"])
    assert np.abs(O.sign_align(r.V, vref) - vref).max() < 1e-6 * np.abs(vref).max()
    if method in (2, 3):  # the bed path and the matrix path agree (test_pca.R:108-131)
        opb = _mk(payload, n, p, stand_method=method)
        assert _relerr(opb.perform_op(v), y_ref) <= 1e-11


def test_random_matrices_vs_oracle(native_lib, monkeypatch):
    """Seeded sweep over shapes, missing rates, standardisation methods and both
    compute paths: statistics bit-exact, operator family within OP_RTOL."""
    rng = np.random.default_rng(20240607)
    for case in range(40):
        n = int(rng.integers(1, 700))
        p = int(rng.integers(1, 400))
        miss = float(rng.choice([0.0, 0.002, 0.05, 0.5]))
        stand = int(rng.choice([O.STANDARDISE_BINOM, O.STANDARDISE_BINOM2]))
        path = ["imma", "generic"][case % 2]
        maf = rng.uniform(0.0, 0.5, size=p)          # includes near-monomorphic SNPs
        g = rng.binomial(2, maf[None, :], size=(n, p))
        codes = np.where(g == 2, 0, np.where(g == 1, 2, 3)).astype(np.uint8)
        codes[rng.random((n, p)) < miss] = 1
        payload = _pack(codes)
        monkeypatch.setenv("FPB_PATH", path)
        op = _mk(payload, n, p, stand_method=stand)
        orc = O.COracle(payload, n, p, stand)
        x = rng.standard_normal(n)
        v = rng.standard_normal(p)
        y_ref = orc.perform_op(x, 0)
        tag = "case %d: n=%d p=%d miss=%g stand=%d path=%s" % (case, n, p, miss, stand, path)
        got, want = op.meansd(), orc.meansd()
        assert np.array_equal(np.isnan(got), np.isnan(want)), tag
        assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), tag
        scale = max(np.abs(y_ref).max(), 1e-30)
        assert np.abs(op.perform_op(x) - y_ref).max() <= OP_RTOL * max(scale, 1.0), tag
        t_ref = orc.crossprod(x)
        assert np.abs(op.crossprod(x) - t_ref).max() <= OP_RTOL * max(np.abs(t_ref).max(), 1.0), tag
        z_ref = orc.prod(v)
        assert np.abs(op.prod(v) - z_ref).max() <= OP_RTOL * max(np.abs(z_ref).max(), 1.0), tag
        assert abs(op.trace - orc.trace) <= 1e-12 * max(orc.trace, 1.0), tag
        op.close()


def test_pca_ndim50_like_the_r_tests(native_lib):
    """flashpcaR's test_pca.R:3-6 asks for ndim = 50 on the 957 x 1129 fixture
    (ncv = 101: the solver's generic, non-register-cached kernels)."""
    from flashpca_b200 import RandomPCA
    _, payload, n, p = load_fixture("data_chr1")
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, 50)
    op = _mk(payload, n, p)
    r = RandomPCA()
    r.pca_fast(None, 0, 50, 500, 1e-6, op=op)
    assert np.abs(r.d / ref["d"] - 1).max() < 1e-6          # test_pca.R uses tol 1e-4
    assert np.abs(r.pve / ref["pve"] - 1).max() < 1e-6
    # |cor| of each PC with the dense one ~ 1 (test_pca.R:24-43)
    for j in range(50):
        c = abs(np.corrcoef(r.Px[:, j], ref["Px"][:, j])[0, 1])
        assert c > 1 - 1e-6, (j, c)


# ---------------------------------------------------------------------------
# Fused single-pass perform_op (fpb_fused.cuh), opt-in with FPB_FUSED=1.
# ---------------------------------------------------------------------------
@pytest.fixture
def fused(monkeypatch):
    monkeypatch.delenv("FPB_PATH", raising=False)
    monkeypatch.delenv("FPB_GEMV", raising=False)
    monkeypatch.setenv("FPB_FUSED", "1")


def _is_fused(op):
    from flashpca_b200 import _lib
    return bool(op.path_info() & _lib.PATH_FUSED)


@pytest.mark.parametrize("name", ["data_chr1", "hapmap3"])
def test_fused_op_fixtures(native_lib, fused, monkeypatch, name):
    _, payload, n, p = load_fixture(name)
    op = _mk(payload, n, p)
    assert _is_fused(op)
    orc = O.COracle(payload, n, p)
    rng = np.random.default_rng(41)
    x = rng.standard_normal(n)
    y_ref = orc.perform_op(x, 0)
    y = op.perform_op(x)
    assert _relerr(y, y_ref) <= OP_RTOL
    assert np.array_equal(op.perform_op(x), y)               # bit-reproducible
    m = rng.standard_normal((n, 3))
    assert _relerr(op.perform_op_mat(m), orc.perform_op(m, 0)) <= OP_RTOL
    # the one-sided ops of a fused handle still take the two-kernel path
    assert _relerr(op.crossprod(x), orc.crossprod(x)) <= OP_RTOL
    # same result as the two-kernel path up to the FP64 recombination order
    monkeypatch.setenv("FPB_FUSED", "0")
    two = _mk(payload, n, p)
    assert not _is_fused(two)
    assert _relerr(y, two.perform_op(x)) <= 1e-13
    # scale invariance, zero, unit and non-finite vectors
    for sc in (1e-300, 1e-30, 1e30, 1e250):
        assert _relerr(op.perform_op(x * sc) / sc, y) <= OP_RTOL
    assert np.all(op.perform_op(np.zeros(n)) == 0.0)
    e = np.zeros(n)
    e[17] = 1.0
    assert _relerr(op.perform_op(e), orc.perform_op(e, 0)) <= OP_RTOL
    assert np.isnan(op.perform_op(np.full(n, np.nan))).all()
    assert _relerr(op.perform_op(x), y_ref) <= OP_RTOL       # state is clean after a NaN op


@pytest.mark.parametrize("n,p", [(1, 3), (3, 1), (4, 5), (63, 9), (64, 64), (65, 130), (513, 129),
                                 (1000, 17), (4099, 257), (16384 + 5, 40), (3000, 1100)])
def test_fused_op_ragged_shapes(native_lib, fused, n, p):
    rng = np.random.default_rng(n * 1000 + p + 1)
    codes = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(n, p),
                       p=[0.15, 0.02, 0.38, 0.45])
    payload = _pack(codes)
    op = _mk(payload, n, p)
    assert _is_fused(op)
    orc = O.COracle(payload, n, p)
    x = rng.standard_normal(n)
    y_ref = orc.perform_op(x, 0)
    y = op.perform_op(x)
    assert np.isfinite(y).all()
    assert np.abs(y - y_ref).max() <= OP_RTOL * max(np.abs(y_ref).max(), 1.0)
    assert np.array_equal(op.perform_op(x), y)
    op.close()


def test_fused_op_exponent_growth_and_dead_snps(native_lib, fused):
    """The running exponent of a grows along the SNP axis (rare variants late in
    the file have small sd, hence large a_j): the second half must drain its
    accumulators and carry on; monomorphic and all-missing SNPs contribute 0."""
    n, p = 2100, 1500
    rng = np.random.default_rng(77)
    maf = np.concatenate([rng.uniform(0.3, 0.5, 600), rng.uniform(0.002, 0.01, 400),
                          np.zeros(100), rng.uniform(0.0005, 0.002, 400)])
    g = rng.binomial(2, maf[None, :], size=(n, p))
    codes = np.where(g == 2, 0, np.where(g == 1, 2, 3)).astype(np.uint8)
    codes[rng.random((n, p)) < 0.002] = 1
    codes[:, 1050] = 1                                      # all missing
    payload = _pack(codes)
    op = _mk(payload, n, p)
    assert _is_fused(op)
    orc = O.COracle(payload, n, p)
    x = rng.standard_normal(n)
    y_ref = orc.perform_op(x, 0)
    y = op.perform_op(x)
    assert np.abs(y - y_ref).max() <= OP_RTOL * np.abs(y_ref).max()
    assert np.array_equal(op.perform_op(x), y)


def test_fused_solver_matches_two_kernel_solver(native_lib, fused, monkeypatch):
    _, payload, n, p = load_fixture("hapmap3")
    op = _mk(payload, n, p)
    assert _is_fused(op)
    got = op.pca(10, 21, 500, 1e-6)
    monkeypatch.setenv("FPB_FUSED", "0")
    two = _mk(payload, n, p).pca(10, 21, 500, 1e-6)
    assert got["nconv"] == 10
    assert np.abs(got["values"] / two["values"] - 1).max() < 1e-9
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, 10)
    assert np.abs(got["values"] / p / ref["d"] - 1).max() < 1e-6


def test_fused_at_full_width(native_lib, fused, monkeypatch):
    """500,000 individuals (148 CTAs x 7 stripes, the shape the fused kernel is
    built for) x 20,000 SNPs: checked against the default two-kernel path on the
    same HBM-resident matrix and for bit-reproducibility.  The fused kernel is
    opt-in (it is slower than the two HBM-bound kernels on B200)."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs > 60 GB HBM")
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(500000, 20000, seed=20240603)
    op = s.create_operator()
    assert _is_fused(op)
    x = np.random.default_rng(2).standard_normal(s.n)
    y = op.perform_op(x)
    for _ in range(3):
        assert np.array_equal(op.perform_op(x), y)
    op.close()
    monkeypatch.delenv("FPB_FUSED", raising=False)
    two = s.create_operator()
    assert not _is_fused(two)                  # default path
    assert _relerr(y, two.perform_op(x)) <= 1e-13


def test_persistent_and_one_shot_grids_agree(native_lib, monkeypatch):
    """Small grids use the persistent TMA kernels (one CTA per SM walks a work
    list), large ones the one-item-per-CTA grids; FPB_PERSIST=0|1 forces the
    choice.  Same items, same sums (the split counts differ, hence ~1e-15
    differences)."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    monkeypatch.delenv("FPB_GEMV", raising=False)
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(70001, 3001, seed=7, missing_rate=0.004)
    rng = np.random.default_rng(3)
    x, v = rng.standard_normal(s.n), rng.standard_normal(s.p)
    monkeypatch.setenv("FPB_PERSIST", "1")
    a = s.create_operator()
    ya, ta, za = a.perform_op(x), a.crossprod(x), a.prod(v)
    assert np.array_equal(a.perform_op(x), ya)
    monkeypatch.setenv("FPB_PERSIST", "0")
    b = s.create_operator()
    assert _relerr(b.perform_op(x), ya) <= 1e-13
    assert _relerr(b.crossprod(x), ta) <= 1e-13
    assert _relerr(b.prod(v), za) <= 1e-13
    orc = O.COracle(s.packed_bed(0, 64), s.n, 64)
    sub = s.create_operator(j0=0, j1=64)
    assert _relerr(sub.perform_op(x), orc.perform_op(x, 0)) <= OP_RTOL


# ---------------------------------------------------------------------------
# Out-of-HBM streaming mode (fpb_create_streaming): slabs of SNP columns in
# pinned host memory, streamed through two device buffers on every op.
# ---------------------------------------------------------------------------
def _streaming_op(name, slab, **kw):
    from flashpca_b200 import Data, SVDWideOnline
    stem = FIXTURES[name]
    d = Data()
    d.read_pheno(stem + ".fam", 6)
    d.geno_filename = stem + ".bed"
    d.get_size()
    d.prepare()
    return d, SVDWideOnline(d, 0, 3, snps_per_slab=slab, **kw)


@pytest.mark.parametrize("name,slab", [("data_chr1", 300), ("data_chr1", 1129), ("data_chr1", 1),
                                       ("hapmap3", 5000)])
def test_streaming_operator_family(native_lib, path, name, slab):
    """Ragged last slab, one slab, one SNP per slab; both compute paths.  The slab
    loop is upstream's block loop (svdwide.cpp:48-59): same sums, block order."""
    from flashpca_b200 import _lib
    if slab == 1 and name == "data_chr1":
        slab = 97   # 1129 = 11 * 97 + 62 slabs would be slow to stage; keep a ragged small slab
    _, payload, n, p = load_fixture(name)
    d, op = _streaming_op(name, slab)
    assert op.path_info() & _lib.PATH_STREAMING
    orc = O.COracle(payload, n, p)
    rng = np.random.default_rng(61)
    x, v = rng.standard_normal(n), rng.standard_normal(p)
    y_ref = orc.perform_op(x, slab)                       # the reference's block size = the slab
    assert np.array_equal(op.meansd(), orc.meansd())
    assert np.array_equal(d.X_meansd, orc.meansd())
    assert abs(op.trace - orc.trace) <= 1e-12 * orc.trace
    y = op.perform_op(x)
    assert _relerr(y, y_ref) <= OP_RTOL
    assert np.array_equal(op.perform_op(x), y)            # bit-reproducible on both compute paths
    assert _relerr(op.crossprod(x), orc.crossprod(x)) <= OP_RTOL
    assert _relerr(op.prod(v), orc.prod(v)) <= OP_RTOL
    m = rng.standard_normal((n, 2))
    assert _relerr(op.perform_op_mat(m), orc.perform_op(m, 0)) <= OP_RTOL
    with pytest.raises(_lib.FpbError, match="streaming"):
        op.bed_payload()
    op.close()


def test_streaming_many_slabs_share_one_scratch_set(native_lib, monkeypatch, tmp_path):
    """Many slabs at a realistic N: the slabs share ONE set of per-op scratch (digit slices, split
    partials, gather sums) sized for the largest slab and are staged inside the two slab buffers, so
    device memory does not grow with the number of slabs -- the case this mode exists for (bed
    larger than HBM).  Results equal the resident operator's to the FP64 order of the slab sum."""
    import ctypes
    from flashpca_b200 import Data, SVDWideOnline
    from flashpca_b200.synth import SynthSpec
    monkeypatch.delenv("FPB_PATH", raising=False)
    s = SynthSpec(60001, 2600, seed=3, missing_rate=0.002)
    stem = str(tmp_path / "syn")
    s.write_plink(stem)
    d = Data()
    d.read_pheno(stem + ".fam", 6)
    d.geno_filename = stem + ".bed"
    d.get_size()
    fr, tot = ctypes.c_uint64(), ctypes.c_uint64()

    def used():
        assert native_lib.fpb_device_memory(0, ctypes.byref(fr), ctypes.byref(tot)) == 0
        return tot.value - fr.value

    base = used()
    res = SVDWideOnline(d, 0, 3)
    x = np.random.default_rng(8).standard_normal(s.n)
    y_res = res.perform_op(x)
    res.close()
    few = SVDWideOnline(d, 0, 3, snps_per_slab=1300)     # 2 slabs
    mem_few = used() - base
    y_few = few.perform_op(x)
    few.close()
    many = SVDWideOnline(d, 0, 3, snps_per_slab=100)     # 26 slabs
    mem_many = used() - base
    y_many = many.perform_op(x)
    assert np.array_equal(many.perform_op(x), y_many)    # bit-reproducible
    assert _relerr(y_many, y_res) <= 1e-13 and _relerr(y_few, y_res) <= 1e-13
    got = many.pca(5, 11, 500, 1e-6)
    assert got["nconv"] == 5
    many.close()
    # 13x the slabs, 13x smaller slab buffers: the footprint must shrink, not grow
    assert mem_many < mem_few, (mem_many, mem_few)


def test_streaming_pca_matches_resident(native_lib, monkeypatch):
    monkeypatch.delenv("FPB_PATH", raising=False)
    _, payload, n, p = load_fixture("hapmap3")
    _, op = _streaming_op("hapmap3", 4000)
    got = op.pca(10, 21, 500, 1e-6)
    res = _mk(payload, n, p).pca(10, 21, 500, 1e-6)
    assert got["nconv"] == 10
    assert np.abs(got["values"] / res["values"] - 1).max() < 1e-9
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, 10)
    assert np.abs(got["values"] / p / ref["d"] - 1).max() < 1e-6


def test_streaming_preloaded_meansd_and_shard(native_lib, path):
    """Preloaded (mean, sd) are split per slab; a streamed SNP shard equals the
    resident shard."""
    from flashpca_b200 import Data, SVDWideOnline
    _, payload, n, p = load_fixture("data_chr1")
    base = O.COracle(payload, n, p)
    base.perform_op(np.ones(n), 0)
    maf = base.meansd()[:, 0] / 2.0
    pre = np.asfortranarray(np.stack([2 * maf, 2 * maf * (1 - maf)], axis=1))
    stem = FIXTURES["data_chr1"]
    d = Data()
    d.read_pheno(stem + ".fam", 6)
    d.geno_filename = stem + ".bed"
    d.get_size()
    d.prepare()
    x = np.random.default_rng(9).standard_normal(n)
    op = SVDWideOnline(d, 0, 3, snps_per_slab=250, meansd=pre)
    orc = O.COracle(payload, n, p, meansd=pre)
    assert _relerr(op.perform_op(x), orc.perform_op(x, 0)) <= OP_RTOL
    assert np.array_equal(op.meansd(), pre)
    a = SVDWideOnline(d, 0, 3, snp_begin=200, snp_count=700, snps_per_slab=256)
    b = SVDWideOnline(d, 0, 3, snp_begin=200, snp_count=700)
    assert a.p == b.p == 700
    assert _relerr(a.perform_op(x), b.perform_op(x)) <= 1e-13
    assert np.array_equal(a.meansd(), b.meansd())


def test_two_vector_kernels_match_single_vector_path(native_lib, monkeypatch):
    """Without the tcgen05 block path (FPB_UMMA=0) the block variants take their columns two at a
    time (k_imma_gemv_tma_2v / _t_2v, odd column counts leave one for the single-vector kernels);
    FPB_PAIR=0 loops over columns.  Same integer sums: results agree to the FP64 recombination."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    monkeypatch.delenv("FPB_GEMV", raising=False)
    monkeypatch.setenv("FPB_UMMA", "0")
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(40003, 2501, seed=11, missing_rate=0.003)
    rng = np.random.default_rng(5)
    m = rng.standard_normal((s.n, 3))
    v = rng.standard_normal((s.p, 5))
    m[:, 1] *= 1e-40                      # very different scales per column: one step per lane
    op = s.create_operator()
    y, t, z = op.perform_op_mat(m), op.crossprod2(m), op.prod3(v)
    assert np.array_equal(op.perform_op_mat(m), y)          # bit-reproducible
    for j in range(3):
        assert _relerr(y[:, j], op.perform_op(m[:, j])) <= 1e-13
        assert _relerr(t[:, j], op.crossprod(m[:, j])) <= 1e-13
    for j in range(5):
        assert _relerr(z[:, j], op.prod(v[:, j])) <= 1e-13
    sub = s.create_operator(j0=0, j1=64)
    orc = O.COracle(s.packed_bed(0, 64), s.n, 64)
    assert _relerr(sub.perform_op_mat(m), orc.perform_op(m, 0)) <= OP_RTOL


@pytest.mark.parametrize("name", ["data_chr1", "hapmap3"])
@pytest.mark.parametrize("k", [3, 4, 7, 8, 20])
def test_tcgen05_block_ops_vs_oracle(native_lib, monkeypatch, name, k):
    """Block variants with k >= 3 columns (perform_op_mat / crossprod2 / prod3,
    svdwide.cpp:71-118, 157-188, 312-343) run on the tcgen05 kernels (fpb_umma.cuh), 8 or 4 columns
    per pass; k = 20 is the loadings / --check / --project shape of a 20-dimensional PCA."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    monkeypatch.delenv("FPB_GEMV", raising=False)
    monkeypatch.delenv("FPB_UMMA", raising=False)
    _, payload, n, p = load_fixture(name)
    op = _mk(payload, n, p)
    orc = O.COracle(payload, n, p)
    rng = np.random.default_rng(100 + k)
    m = rng.standard_normal((n, k)) * np.exp(rng.uniform(-30, 30, size=k))   # one scale per column
    w = rng.standard_normal((p, k)) * np.exp(rng.uniform(-30, 30, size=k))
    y = op.perform_op_mat(m)
    y_ref = orc.perform_op(m, 0)
    t, t_ref = op.crossprod2(m), orc.crossprod(m, 0)
    z, z_ref = op.prod3(w), orc.prod(w, 0)
    for j in range(k):
        assert _relerr(y[:, j], y_ref[:, j]) <= OP_RTOL
        assert _relerr(t[:, j], t_ref[:, j]) <= OP_RTOL
        assert _relerr(z[:, j], z_ref[:, j]) <= OP_RTOL
    assert np.array_equal(op.perform_op_mat(m), y)          # bit-reproducible
    op.close()


def test_tcgen05_block_ops_match_single_vector_path(native_lib, monkeypatch):
    """Same integer sums as the mma.sync kernels: per column the block result equals the
    single-vector op to the FP64 recombination order; ragged N and P, missing genotypes, dead
    columns, a zero and a NaN vector in the block."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    monkeypatch.delenv("FPB_GEMV", raising=False)
    monkeypatch.delenv("FPB_UMMA", raising=False)
    from flashpca_b200.synth import SynthSpec
    s = SynthSpec(40003, 2501, seed=11, missing_rate=0.003)
    rng = np.random.default_rng(5)
    m = rng.standard_normal((s.n, 11))
    v = rng.standard_normal((s.p, 6))
    m[:, 1] *= 1e-40
    m[:, 4] = 0.0
    op = s.create_operator()
    y, t, z = op.perform_op_mat(m), op.crossprod2(m), op.prod3(v)
    assert np.array_equal(y[:, 4], np.zeros(s.n)) and np.array_equal(t[:, 4], np.zeros(s.p))
    for j in range(11):
        assert _relerr(y[:, j], op.perform_op(m[:, j])) <= 1e-13
        assert _relerr(t[:, j], op.crossprod(m[:, j])) <= 1e-13
    for j in range(6):
        assert _relerr(z[:, j], op.prod(v[:, j])) <= 1e-13
    m[7, 2] = np.nan
    y = op.perform_op_mat(m)
    assert np.isnan(y[:, 2]).all() and np.isfinite(np.delete(y, 2, axis=1)).all()
    sub = s.create_operator(j0=0, j1=64)
    orc = O.COracle(s.packed_bed(0, 64), s.n, 64)
    assert _relerr(sub.perform_op_mat(m[:, :2].repeat(2, axis=1)), orc.perform_op(m[:, :2].repeat(2, axis=1), 0)) <= OP_RTOL


@pytest.mark.parametrize("n,p", [(1, 5), (3, 1), (129, 517), (513, 130), (2049, 1025), (70001, 300)])
def test_tcgen05_block_ops_ragged(native_lib, monkeypatch, n, p):
    """Edge shapes through the tcgen05 block kernels: fewer rows than one 128-row box, one
    individual, N just past a 512-individual stage, several column splits (N > 65536)."""
    monkeypatch.delenv("FPB_PATH", raising=False)
    monkeypatch.delenv("FPB_UMMA", raising=False)
    rng = np.random.default_rng(n * 31 + p)
    codes = rng.integers(0, 4, size=(n, p), dtype=np.uint8)
    if n > 2:
        codes[:, 0] = 3          # monomorphic column (PLINK 11 = dosage 0 everywhere)
    payload = _pack(codes)
    op = _mk(payload, n, p)
    orc = O.COracle(payload, n, p)
    m = rng.standard_normal((n, 5))
    w = rng.standard_normal((p, 5))
    assert _relerr(op.perform_op_mat(m), orc.perform_op(m, 0)) <= OP_RTOL
    assert _relerr(op.crossprod2(m), orc.crossprod(m, 0)) <= OP_RTOL
    assert _relerr(op.prod3(w), orc.prod(w, 0)) <= OP_RTOL
    op.close()


def test_device_memory_query(native_lib):
    import ctypes
    fr, tot = ctypes.c_uint64(), ctypes.c_uint64()
    assert native_lib.fpb_device_memory(0, ctypes.byref(fr), ctypes.byref(tot)) == 0
    assert 0 < fr.value <= tot.value and tot.value > (8 << 30)
    assert native_lib.fpb_device_memory(99, ctypes.byref(fr), ctypes.byref(tot)) != 0
