"""GPU tests of the C++ front end (flashpca_b200/host -> `flashpca` binary) and
of the multi-GPU path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import FIXTURES, ROOT, load_fixture
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli():
    from flashpca_b200 import build
    build.build_lib()
    path = build.build_cli()
    assert os.path.exists(path)
    return path


def _read_table(path, skip_header=True, first_numeric_col=2):
    rows = open(path).read().splitlines()
    hdr = rows[0].split("\t") if skip_header else None
    body = rows[1:] if skip_header else rows
    ids = [r.split("\t")[:first_numeric_col] for r in body]
    vals = np.array([[float(v) for v in r.split("\t")[first_numeric_col:]] for r in body])
    return hdr, ids, vals


def test_cli_pca_outputs_match_dense(cli, tmp_path):
    """HapMap3/test_pca.R:40-43 command line (BASELINE config 0): --ndim 10 --tol 1e-6
    --outload --outmeansd --precision 20, outputs vs dense eigh, FID/IID order exact."""
    stem = FIXTURES["hapmap3"]
    out = subprocess.run([cli, "--bfile", stem, "--ndim", "10", "--tol", "1e-6", "--outload",
                          "loadings.txt", "--outmeansd", "meansd.txt", "--precision", "20",
                          "--notime", "-v"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PCA begin" in out.stdout and "PCA done" in out.stdout and "Goodbye!" in out.stdout
    assert "blocksize: 14389 (" in out.stdout          # flashpca.cpp:690 log line
    _, payload, n, p = load_fixture("hapmap3")
    x, msd = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, 10)
    ev = np.loadtxt(tmp_path / "eigenvalues.txt")
    assert np.abs(ev / ref["d"] - 1).max() < 1e-6
    pve = np.loadtxt(tmp_path / "pve.txt")
    assert np.abs(pve / ref["pve"] - 1).max() < 1e-6
    hdr, ids, u = _read_table(tmp_path / "eigenvectors.txt")
    assert hdr == ["FID", "IID"] + ["U%d" % (i + 1) for i in range(10)]
    fid, iid = O.read_fam_ids(stem + ".fam")
    assert [r[0] for r in ids] == fid and [r[1] for r in ids] == iid   # bit-exact order
    assert np.abs(O.sign_align(u, ref["U"]) - ref["U"]).max() < 5e-6    # tol 1e-6 solve
    hdr, _, pcs = _read_table(tmp_path / "pcs.txt")
    assert hdr[2] == "PC1"
    assert np.abs(O.sign_align(pcs, ref["Px"]) - ref["Px"]).max() < 5e-6 * np.abs(ref["Px"]).max()
    hdr, sid, ms = _read_table(tmp_path / "meansd.txt")
    assert hdr == ["SNP", "RefAllele", "Mean", "SD"]
    assert np.array_equal(ms, msd)                    # --precision 20 round-trips doubles
    bim = [ln.split() for ln in open(stem + ".bim").read().splitlines()]
    assert [r[0] for r in sid] == [b[1] for b in bim] and [r[1] for r in sid] == [b[4] for b in bim]
    hdr, _, v = _read_table(tmp_path / "loadings.txt")
    vref = O.dense_loadings(x, ref["U"], ref["d"], ref["div"])
    assert np.abs(O.sign_align(v, vref) - vref).max() < 5e-6 * np.abs(vref).max()

    # --check on the files just written (README.md:194-207): mse < 1e-8
    chk = subprocess.run([cli, "--bfile", stem, "--check", "--outvec", "eigenvectors.txt",
                          "--outval", "eigenvalues.txt", "--notime"], cwd=tmp_path,
                         capture_output=True, text=True, timeout=300)
    assert chk.returncode == 0, chk.stdout + chk.stderr
    mse = float(chk.stdout.split("Mean squared error: ")[1].split(",")[0])
    assert mse < 1e-8

    # --project with the saved loadings + mean/sd reproduces the PCs (HapMap3/test_pca.R:46-67)
    prj = subprocess.run([cli, "--bfile", stem, "--project", "--inload", "loadings.txt",
                          "--inmeansd", "meansd.txt", "--outproj", "proj.txt", "--notime",
                          "--precision", "20"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=300)
    assert prj.returncode == 0, prj.stdout + prj.stderr
    _, _, proj = _read_table(tmp_path / "proj.txt")
    assert np.abs(proj - pcs).max() < 1e-5 * np.abs(pcs).max()


def test_cli_batch_mode_matches_online(cli, tmp_path):
    """--batch (read_bed + standardise + SVDWide, flashpca.cpp:597-601,
    randompca.cpp:121-166) and the default online mode give the same PCA."""
    stem = FIXTURES["data_chr1"]
    for tag, extra in (("online", []), ("batch", ["--batch"])):
        out = subprocess.run([cli, "--bfile", stem, "--ndim", "5", "--tol", "1e-9", "--notime",
                              "--precision", "17", "--suffix", "_%s.txt" % tag] + extra,
                             cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
    a = np.loadtxt(tmp_path / "eigenvalues_online.txt")
    b = np.loadtxt(tmp_path / "eigenvalues_batch.txt")
    assert np.abs(a / b - 1).max() < 1e-9
    _, _, ua = _read_table(tmp_path / "eigenvectors_online.txt")
    _, _, ub = _read_table(tmp_path / "eigenvectors_batch.txt")
    assert np.abs(O.sign_align(ua, ub) - ub).max() < 1e-7
    _, payload, n, p = load_fixture("data_chr1")
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    assert np.abs(b / O.dense_pca(x, 5)["d"] - 1).max() < 1e-6


def test_cli_default_precision_and_errors(cli, tmp_path):
    stem = FIXTURES["data_chr1"]
    out = subprocess.run([cli, "--bfile", stem, "--ndim", "3", "--notime", "--suffix", ".tsv"],
                         cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = open(tmp_path / "eigenvalues.tsv").read().split()
    assert len(lines) == 3 and all(len(v.replace(".", "").lstrip("0")) <= 7 for v in lines)
    # max_dim guard (flashpca.cpp:623-633)
    bad = subprocess.run([cli, "--bfile", stem, "--ndim", "500"], cwd=tmp_path,
                         capture_output=True, text=True)
    assert bad.returncode != 0 and "but only 478allowed" in bad.stderr
    bad = subprocess.run([cli, "--bfile", stem, "--standx", "sd"], cwd=tmp_path,
                         capture_output=True, text=True)
    assert bad.returncode != 0 and "unknown standardization method" in bad.stderr
    bad = subprocess.run([cli, "--bfile", "/nonexistent/x"], cwd=tmp_path, capture_output=True,
                         text=True)
    assert bad.returncode != 0 and "Error reading file" in bad.stderr
    bad = subprocess.run([cli, "--bfile", stem, "--memory", "5", "--blocksize", "10"],
                         cwd=tmp_path, capture_output=True, text=True)
    assert bad.returncode != 0 and "cannot specify both --memory and --blocksize" in bad.stderr


def test_cli_block_solver_extension(cli, tmp_path):
    """FPB_SOLVER=block: the command line with the block Krylov solver writes the same outputs."""
    stem = FIXTURES["hapmap3"]
    env = dict(os.environ, FPB_SOLVER="block")
    out = subprocess.run([cli, "--bfile", stem, "--ndim", "10", "--notime", "--precision", "12", "-v",
                          "--suffix", ".blk"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Block Krylov solver:" in out.stdout
    _, payload, n, p = load_fixture("hapmap3")
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    ref = O.dense_pca(x, 10)
    ev = np.loadtxt(tmp_path / "eigenvalues.blk")
    assert np.abs(ev / ref["d"] - 1).max() < 1e-6
    _, _, u = _read_table(tmp_path / "eigenvectors.blk")
    assert np.abs(O.sign_align(u, ref["U"]) - ref["U"]).max() < 5e-6


def test_cli_ndim_beyond_63(cli, tmp_path):
    """`--ndim 100` (ncv = 201) and the reference's own maximum on this fileset, 478
    (flashpca.cpp:623-633): accepted by the guard AND by the solver."""
    stem = FIXTURES["data_chr1"]
    _, payload, n, p = load_fixture("data_chr1")
    x, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    for ndim in (100, 478):
        out = subprocess.run([cli, "--bfile", stem, "--ndim", str(ndim), "--notime", "--precision",
                              "12", "--suffix", ".%d" % ndim], cwd=tmp_path, capture_output=True,
                             text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        ev = np.loadtxt(tmp_path / ("eigenvalues.%d" % ndim))
        assert ev.shape == (ndim,)
        assert np.abs(ev / O.dense_pca(x, ndim)["d"] - 1).max() < 1e-6


def test_eigen_typed_operator_surface_runs(tmp_path):
    """The Eigen::MatrixXd overloads of host/svdwide.hpp (the maintainer's swap-the-header path,
    INTEGRATION.md section 1) return exactly what the Matrix-typed calls return."""
    from test_host import build_eigen_adaptor_check
    exe = build_eigen_adaptor_check(tmp_path)
    out = subprocess.run([exe, FIXTURES["data_chr1"]], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "max abs difference: 0" in out.stdout


@pytest.mark.parametrize("peer", ["1", "0"])
def test_nccl_sharded_two_gpus(peer):
    """SNP-sharded op and solve on 2 GPUs, one process each: with the peer-memory shard sum
    (csrc/fpb_peer.cuh over CUDA IPC; FPB_PEER=1, the default) and with ncclAllReduce (FPB_PEER=0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(ROOT, "tests", "_nccl_shard_worker.py")
    env = dict(os.environ, FPB_PEER=peer, FPB_PEER_TIMEOUT_S="20")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          "29617", script], capture_output=True, text=True, timeout=180, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "NCCL_SHARD_OK" in out.stdout
    assert ("shard sum: peer" if peer == "1" else "shard sum: nccl") in out.stdout


@pytest.mark.parametrize("mode,stand,divisor", [("plink", 3, 2), ("plink", 2, 1), ("matrix", 3, 2),
                                                ("matrix", 1, 2), ("matrix", 0, 0)])
def test_flashpcar_entry_points(mode, stand, divisor, tmp_path):
    """flashpca_plink_internal / flashpca_internal (flashpcaR/src/flashpca.cpp:17-197)
    as plain C++ (host/flashpcar.hpp): same parameter lists and named result fields;
    checked like flashpcaR/tests/testthat/test_pca.R:45-105 against the dense
    eigendecomposition of the oracle-standardised matrix."""
    import json
    from flashpca_b200 import build
    build.build_lib()
    exe = build.build_rapi_check()
    stem = FIXTURES["data_chr1"]
    ndim = 6
    out = subprocess.run([exe, mode, stem, str(ndim), str(stand), str(divisor), "1", "1"],
                         capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0, out.stderr
    res = json.loads(out.stdout)
    _, payload, n, p = load_fixture("data_chr1")
    x = O.dosage_matrix(O.dense_codes(payload, n, p))
    s_ref, msd = O.standardise_matrix(x, stand)
    div = {0: 1.0, 1: n - 1.0, 2: float(p)}[divisor]
    w, v = np.linalg.eigh(s_ref @ s_ref.T / div)
    d_ref, u_ref = w[::-1][:ndim], v[:, ::-1][:, :ndim]
    d = np.array(res["values"])
    assert np.abs(d / d_ref - 1).max() < 1e-6
    mat = lambda m: np.array(m["data"]).reshape(m["ncol"], m["nrow"]).T
    u = O.sign_align(mat(res["vectors"]), u_ref)
    assert np.abs(u - u_ref).max() < 1e-6
    px = mat(res["projection"])
    assert np.abs(np.abs(px) - np.abs(u_ref * np.sqrt(d_ref))).max() < 1e-6 * np.abs(px).max()
    trace = np.sum(s_ref * s_ref) / div
    assert np.abs(np.array(res["pve"]) / (d_ref / trace) - 1).max() < 1e-6
    vload = mat(res["loadings"])
    vref = s_ref.T @ u_ref / np.sqrt(d_ref) / np.sqrt(div)
    assert np.abs(O.sign_align(vload, vref) - vref).max() < 1e-6 * np.abs(vref).max()
    if stand != 0:      # return_scale: center / scale = per-SNP mean / sd (flashpca.cpp:46-52)
        assert np.allclose(res["center"], msd[:, 0], rtol=1e-12, atol=1e-13)
        assert np.allclose(res["scale"], msd[:, 1], rtol=1e-12, atol=1e-13)
    else:
        assert res["center"] == [] and res["scale"] == []
    if mode == "plink":  # rownames "FID:IID" in fam order (flashpca.cpp:147-155)
        fid, iid = O.read_fam_ids(stem + ".fam")
        assert res["rownames"] == [a + ":" + b for a, b in zip(fid, iid)]
    else:
        assert res["rownames"] == []


def test_cli_streaming_mode_matches_resident(cli, tmp_path):
    """FPB_STREAM_SLAB_SNPS forces the out-of-HBM mode of the C++ front end (the
    role of upstream's --memory block size): same output files."""
    stem = FIXTURES["data_chr1"]
    outs = {}
    for tag, env in (("res", {}), ("str", {"FPB_STREAM_SLAB_SNPS": "200"})):
        wd = tmp_path / tag
        wd.mkdir()
        e = dict(os.environ)
        e.pop("FPB_STREAM_SLAB_SNPS", None)
        e.update(env)
        out = subprocess.run([cli, "--bfile", stem, "--ndim", "5", "--tol", "1e-8", "--precision",
                              "15", "--notime", "-v"], cwd=wd, capture_output=True, text=True, env=e)
        assert out.returncode == 0, out.stdout + out.stderr
        if tag == "str":
            assert "streamed from host memory" in out.stdout
        outs[tag] = np.loadtxt(wd / "eigenvalues.txt")
    assert np.abs(outs["str"] / outs["res"] - 1).max() < 1e-9


def _script_rmse(a, b):
    """HapMap3/test_pca.R:156-160: sign-invariant per-column criterion of the reference script."""
    r = [min(np.mean(a[:, m] - b[:, m]) ** 2, np.mean(a[:, m] + b[:, m]) ** 2)
         for m in range(a.shape[1])]
    return float(np.sqrt(np.sum(r)))


def test_reference_end_to_end_script(cli, tmp_path):
    """HapMap3/test_pca.R:40-246 with numpy's dense SVD in the place of R's svd()/RSpectra:
    PCA of HM3_thinned_autosomal_overlap (--ndim 10 --tol 1e-6 --outload --outmeansd --precision 20),
    --project onto the same data, onto the 1000 Genomes set (N = 1092, no missing genotypes) and with
    --inmaf, then --check; FID/IID and SNP/allele order exact, everything else at err.tol = 1e-6."""
    import re
    from conftest import GOLDEN
    hm3 = os.path.join(GOLDEN, "hm3_overlap", "HM3_thinned_autosomal_overlap")
    kg1 = os.path.join(GOLDEN, "kg1_overlap",
                       "1kg.ref.phase1_release_v3.20101123_thinned_autosomal_overlap")
    k, tol, err_tol = 10, 1e-6, 1e-6

    def run(*args):
        out = subprocess.run([cli, *args], cwd=tmp_path, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        return out.stdout

    # ---- the script's R side, in numpy (scale2, svd)
    n1 = O.count_lines(hm3 + ".fam")
    pay1, _, p = O.read_bed_payload(hm3 + ".bed", n1)
    d1 = O.dosage_matrix(O.dense_codes(pay1, n1, p))
    maf = np.nanmean(d1, axis=0) / 2
    center, scale = 2 * maf, np.sqrt(2 * maf * (1 - maf))
    x = (d1 - center) / scale
    x[np.isnan(x)] = 0.0
    x /= np.sqrt(p)
    u, s, vt = np.linalg.svd(x, full_matrices=False)
    u, s, v = u[:, :k], s[:k], vt[:k].T
    fam = [ln.split() for ln in open(hm3 + ".fam")]
    bim = [ln.split() for ln in open(hm3 + ".bim")]

    # ---- the script's command lines
    run("--bfile", hm3, "--ndim", str(k), "--tol", str(tol), "--outload", "loadings.txt",
        "--outmeansd", "meansd.txt", "--precision", "20")
    run("--bfile", hm3, "--project", "--inmeansd", "meansd.txt", "--outproj", "projections.txt",
        "--inload", "loadings.txt", "-v", "--precision", "20")
    run("--bfile", kg1, "--project", "--inmeansd", "meansd.txt", "--outproj", "projections.1kg.txt",
        "--inload", "loadings.txt", "-v", "--precision", "20")
    chk = run("--bfile", hm3, "--check", "--outval", "eigenvalues.txt", "--outvec",
              "eigenvectors.txt", "-v", "--precision", "20", "--notime")

    hdr, ids, evec = _read_table(tmp_path / "eigenvectors.txt")
    assert [r[0] for r in ids] == [f[0] for f in fam] and [r[1] for r in ids] == [f[1] for f in fam]
    ev = np.loadtxt(tmp_path / "eigenvalues.txt")
    _, ids, pcs = _read_table(tmp_path / "pcs.txt")
    assert [r[0] for r in ids] == [f[0] for f in fam] and [r[1] for r in ids] == [f[1] for f in fam]
    pve = np.loadtxt(tmp_path / "pve.txt")
    _, sid, msd = _read_table(tmp_path / "meansd.txt")
    assert [r[0] for r in sid] == [b[1] for b in bim] and [r[1] for r in sid] == [b[4] for b in bim]
    _, sid, load = _read_table(tmp_path / "loadings.txt")
    assert [r[0] for r in sid] == [b[1] for b in bim] and [r[1] for r in sid] == [b[4] for b in bim]
    _, ids, proj = _read_table(tmp_path / "projections.txt")
    assert [r[0] for r in ids] == [f[0] for f in fam] and [r[1] for r in ids] == [f[1] for f in fam]
    _, ids1kg, proj1kg = _read_table(tmp_path / "projections.1kg.txt")
    fam2 = [ln.split() for ln in open(kg1 + ".fam")]
    assert [r[0] for r in ids1kg] == [f[0] for f in fam2] and [r[1] for r in ids1kg] == [f[1] for f in fam2]

    # scaling (test_pca.R:123-139), eigenvalues (:143-150), pve (:184-192)
    assert np.sqrt(np.mean((msd[:, 0] - center) ** 2)) < err_tol
    assert np.sqrt(np.mean((msd[:, 1] - scale) ** 2)) < err_tol
    assert np.sqrt(np.mean((ev - s ** 2) ** 2)) < err_tol
    assert np.sqrt(np.mean((pve - s ** 2 / (x ** 2).sum()) ** 2)) < err_tol
    # eigenvectors, PCs, loadings, projection = PCs (:154-228): the script's criterion and a
    # stricter element-wise one (tol 1e-6 solve: vector error ~ residual / gap)
    xv = x @ v
    for got, want, scale_ in ((evec, u, 1.0), (pcs, xv, np.abs(xv).max()), (load, v, 1.0),
                              (proj, xv, np.abs(xv).max())):
        assert _script_rmse(got, want) < err_tol
        assert np.abs(O.sign_align(got, want) - want).max() < 2e-5 * scale_
    # 1KG projected onto the HM3 PCA (:111-114, 232-245)
    n2 = O.count_lines(kg1 + ".fam")
    pay2, _, p2 = O.read_bed_payload(kg1 + ".bed", n2)
    assert (n2, p2) == (1092, p)
    d2 = O.dosage_matrix(O.dense_codes(pay2, n2, p2))
    assert not np.isnan(d2).any()
    want1kg = ((d2 - center) / scale) @ v / np.sqrt(p)
    assert _script_rmse(proj1kg, want1kg) < err_tol
    assert np.abs(O.sign_align(proj1kg, want1kg) - want1kg).max() < 2e-5 * np.abs(want1kg).max()
    # --check (:70-76, 107-110, 249-250): printed sums of squared errors vs the dense formula
    rows = [re.split(r",| ", ln) for ln in chk.splitlines() if "eval" in ln]
    assert len(rows) == k
    sse_obs = np.array([float(r[6]) for r in rows])
    assert np.allclose([float(r[1]) for r in rows], ev, rtol=1e-5)   # printed with 6 significant digits
    sse_exp = ((x @ (x.T @ evec) - evec * ev[None, :]) ** 2).sum(axis=0)
    assert ((sse_obs - sse_exp) ** 2 < err_tol).all()
    assert np.sqrt(sse_obs.sum() / (n1 * k)) < 1e-4       # README.md:207: mse < 1e-8

    # --inmaf (:62-67).  The script writes a 2-column maf.txt, which read_MAF (data.cpp:419-496)
    # rejects: it wants PLINK .frq rows (CHR SNP A1 A2 MAF NCHROBS), and maf2meansd
    # (randompca.cpp:745-751) uses sd = 2 p (1 - p) without the square root.
    with open(tmp_path / "maf.txt", "w") as f:
        f.write("SNP MAF\n")
        for b, m in zip(bim, maf):
            f.write("%s %.20g\n" % (b[1], m))
    bad = subprocess.run([cli, "--bfile", hm3, "--project", "--inmaf", "maf.txt", "--outproj",
                          "projections.maf.txt", "--inload", "loadings.txt"], cwd=tmp_path,
                         capture_output=True, text=True)
    assert bad.returncode != 0 and "inconsistent number of columns" in bad.stderr
    with open(tmp_path / "hm3.frq", "w") as f:
        f.write(" CHR SNP A1 A2 MAF NCHROBS\n")
        for b, m in zip(bim, maf):
            f.write(" %s %s %s %s %.20g %d\n" % (b[0], b[1], b[4], b[5], m, 2 * n1))
    run("--bfile", hm3, "--project", "--inmaf", "hm3.frq", "--outproj", "projections.maf.txt",
        "--inload", "loadings.txt", "--precision", "20")
    _, _, projmaf = _read_table(tmp_path / "projections.maf.txt")
    xm = (d1 - 2 * maf) / (2 * maf * (1 - maf))
    xm[np.isnan(xm)] = 0.0
    wantmaf = xm @ load / np.sqrt(p)
    assert np.abs(projmaf - wantmaf).max() < 1e-9 * np.abs(wantmaf).max()
