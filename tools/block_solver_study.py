"""How many passes over the genotype matrix would a block solver need?  
CPU study on oracle-sized matrices: block Krylov (block Lanczos with full re-orthogonalisation,
Rayleigh-Ritz on the accumulated block Krylov space) against the single-vector IRLM restatement
of the oracle.  A "pass" = one application of X X' to a block of b vectors (one read of the
matrix per half on the GPU, shared by the b vectors)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from flashpca_b200.synth import SynthSpec  # noqa: E402


def block_krylov(A_mul, n, k, b, tol, max_blocks=60, seed=1):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, b)))
    basis = [Q]
    AV = []
    for it in range(1, max_blocks + 1):
        W = A_mul(basis[-1])            # one pass over the matrix for b vectors
        AV.append(W)
        V = np.concatenate(basis, axis=1)
        H = V.T @ np.concatenate(AV, axis=1)
        H = (H + H.T) / 2
        w, S = np.linalg.eigh(H)
        w, S = w[::-1][:k], S[:, ::-1][:, :k]
        # residuals of the Ritz pairs: ||A u - theta u|| with A V known
        U = V @ S
        R = np.concatenate(AV, axis=1) @ S - U * w
        res = np.linalg.norm(R, axis=0)
        if V.shape[1] >= k and np.all(res < tol * np.maximum(np.abs(w), 2.2e-16 ** (2 / 3))):
            return w, it
        # next block: orthogonalise W against everything (twice)
        for _ in range(2):
            W = W - V @ (V.T @ W)
        Qn, _ = np.linalg.qr(W)
        basis.append(Qn)
    return w, max_blocks


def study(name, x, k=20):
    n, p = x.shape
    ref = O.dense_pca(x, k)
    A_mul = lambda M: x @ (x.T @ M)
    print("%s: N=%d P=%d k=%d" % (name, n, p, k))
    res = O.spectra_irlm(lambda v: A_mul(v), n, k, 2 * k + 1, 500, 1e-6)
    err = np.abs(res["values"] / p / ref["d"] - 1).max()
    print("  single-vector IRLM (Spectra schedule, tol 1e-6): %d passes, eigenvalue rel err %.1e"
          % (res["nops"], err))
    for b in (2, 4, 8, 16, 24):
        w, passes = block_krylov(A_mul, n, k, b, 1e-6)
        err = np.abs(w / p / ref["d"] - 1).max()
        print("  block Krylov b=%2d: %3d block passes (%4d vector-products), eigenvalue rel err %.1e"
              % (b, passes, passes * b, err))


if __name__ == "__main__":
    s = SynthSpec(1501, 6000, seed=5, npop=25, fst=0.05)
    xs, _ = O.dense_standardise(np.ascontiguousarray(s.codes().T))
    study("synthetic Balding-Nichols (25 populations)", xs)
    stem = os.path.join(ROOT, "tests", "golden", "hapmap3", "data")
    n = O.count_lines(stem + ".fam")
    payload, _, p = O.read_bed_payload(stem + ".bed", n)
    xh, _ = O.dense_standardise(O.dense_codes(payload, n, p))
    study("HapMap3 fixture", xh)
