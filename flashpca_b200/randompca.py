"""Host mirror of upstream class RandomPCA for the PCA path
(randompca.h:54-108, randompca.cpp:168-218, 627-820)."""
from __future__ import annotations

import numpy as np

from ._lib import FpbError
from .svdwide import SVDWideOnline

DIVISOR_NONE, DIVISOR_N1, DIVISOR_P = 0, 1, 2


class RandomPCA:
    def __init__(self):
        self.stand_method_x = 3
        self.divisor = DIVISOR_P
        self.verbose = False
        self.U = self.V = self.Px = self.d = self.pve = self.X_meansd = None
        self.trace = 0.0
        self.err = None
        self.mse = self.rmse = 0.0
        self.nops = 0

    def _div(self, n, p):
        return {DIVISOR_NONE: 1.0, DIVISOR_N1: n - 1.0, DIVISOR_P: float(p)}[self.divisor]

    def pca_fast(self, dat, block_size: int, ndim: int, maxiter: int, tol: float, seed: int = 1,
                 do_loadings: bool = False, device: int = 0, op: SVDWideOnline | None = None):
        """randompca.cpp:168-218.  `seed` and `block_size` are accepted and, as
        upstream, do not influence the solve."""
        own = op is None
        if own:
            op = SVDWideOnline(dat, block_size, self.stand_method_x, self.verbose, device=device)
        try:
            n, p = op.n, op.p
            res = op.pca(ndim, 2 * ndim + 1, maxiter, tol)
            self.nops = res["nops"]
            if res["nconv"] < ndim:  # randompca.cpp:212-217
                raise FpbError("Spectra eigen-decomposition was not successful, status: 1")
            div = self._div(n, p)
            self.U = res["vectors"]
            self.d = res["values"] / div
            if do_loadings:  # randompca.cpp:191-204
                v = op.crossprod2(self.U)
                self.V = v / np.sqrt(self.d)[None, :] / np.sqrt(div)
            self.trace = op.trace / div
            self.pve = self.d / self.trace
            self.Px = self.U * np.sqrt(self.d)[None, :]
            self.X_meansd = op.meansd()
        finally:
            if own:
                op.close()
        return self

    def pca_fast_matrix(self, x, ndim: int, maxiter: int, tol: float, seed: int = 1,
                        do_loadings: bool = False, device: int = 0):
        """randompca.cpp:121-166 (the MatrixXd overload; flashpca_internal,
        flashpcaR/src/flashpca.cpp:17-93; `flashpca --batch`)."""
        from .svdwide import SVDWide
        op = SVDWide(x, self.stand_method_x, self.verbose, device)
        try:
            n, p = op.n, op.p
            res = op.pca(ndim, 2 * ndim + 1, maxiter, tol)
            self.nops = res["nops"]
            if res["nconv"] < ndim:
                raise FpbError("Spectra eigen-decomposition was not successful, status: 1")
            div = self._div(n, p)
            self.U = res["vectors"]
            self.d = res["values"] / div
            if do_loadings:  # randompca.cpp:149-153
                self.V = op.crossprod2(self.U) / np.sqrt(self.d)[None, :] / np.sqrt(div)
            self.trace = op.trace / div
            self.pve = self.d / self.trace
            self.Px = self.U * np.sqrt(self.d)[None, :]
            self.X_meansd = op.meansd()
        finally:
            op.close()
        return self

    def check(self, dat, block_size: int, evec: np.ndarray, evals: np.ndarray, device: int = 0,
              op: SVDWideOnline | None = None):
        """randompca.cpp:663-703: err_j = || X X' u_j / div - u_j d_j ||^2."""
        own = op is None
        if own:
            op = SVDWideOnline(dat, block_size, self.stand_method_x, self.verbose, device=device)
        try:
            evec = np.asarray(evec, dtype=np.float64)
            evals = np.asarray(evals, dtype=np.float64).ravel()
            if evec.shape[0] != op.n:
                raise FpbError("Eigenvector dimension doesn't match data")
            if evec.shape[1] != evals.size:
                raise FpbError("Eigenvector dimension doesn't match the number of eigenvalues")
            div = self._div(op.n, op.p)
            xxu = op.perform_op_mat(evec) / div
            resid = xxu - evec * evals[None, :]
            self.err = np.sum(resid * resid, axis=0)
            self.mse = float(self.err.sum() / (op.n * evals.size))
            self.rmse = float(np.sqrt(self.mse))
        finally:
            if own:
                op.close()
        return self

    def project(self, dat, block_size: int, loadings: np.ndarray, device: int = 0,
                op: SVDWideOnline | None = None):
        """randompca.cpp:798-820: Px = X V / sqrt(div) with dat.X_meansd preloaded."""
        own = op is None
        if own:
            op = SVDWideOnline(dat, block_size, self.stand_method_x, self.verbose, device=device)
        try:
            v = np.asarray(loadings, dtype=np.float64)
            if v.shape[0] != op.p:
                raise FpbError("The number of SNPs in the loadings doesn't match the data")
            self.V = v
            self.Px = op.prod3(v) / np.sqrt(self._div(op.n, op.p))
        finally:
            if own:
                op.close()
        return self
