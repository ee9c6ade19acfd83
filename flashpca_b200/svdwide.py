"""Host mirror of upstream class SVDWideOnline (svdwide.h:32-107): the
matrix-free operator Spectra calls back into.  Same method names and argument
meaning; every body forwards to the C ABI (include/flashpca_b200.h)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import FpbError, check


def _f64(a, rows):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.shape[0] != rows:
        raise FpbError("dimension mismatch: expected %d rows, got %d" % (rows, a.shape[0]))
    return np.asfortranarray(a)


class SVDWide:
    """svdwide.h:9-30 + standardise() (util.cpp:24-192): the in-memory operator of
    RandomPCA::pca_fast(MatrixXd&, ...).  `mat` is an N x P dosage matrix with NaN
    for missing; it is copied to HBM and standardised there."""

    def __init__(self, mat, stand_method: int = 3, verbose: bool = False, device: int = 0):
        self.lib = _lib.load()
        x = np.asfortranarray(np.asarray(mat, dtype=np.float64))
        if x.ndim != 2:
            raise FpbError("X must be a matrix")
        self.n, self.p = x.shape
        self.h = ctypes.c_void_p()
        check(self.lib.fpb_create_dense(ctypes.byref(self.h), x.ctypes.data, self.n, self.p,
                                        int(stand_method), device))
        self.nops = 1

    close = lambda self: SVDWideOnline.close(self)
    __del__ = lambda self: SVDWideOnline.__del__(self)
    rows = lambda self: self.n
    cols = lambda self: self.n
    trace = property(lambda self: SVDWideOnline.trace.fget(self))
    meansd = lambda self: SVDWideOnline.meansd(self)
    _call = lambda self, *a: SVDWideOnline._call(self, *a)
    perform_op = lambda self, x, y=None: SVDWideOnline.perform_op(self, x, y)
    crossprod2 = lambda self, x: SVDWideOnline.crossprod2(self, x)
    prod3 = lambda self, x: SVDWideOnline.prod3(self, x)
    pca = lambda self, *a, **kw: SVDWideOnline.pca(self, *a, **kw)
    _vector_buffer = lambda self, *a: SVDWideOnline._vector_buffer(self, *a)

    def standardised(self) -> np.ndarray:
        out = np.zeros((self.n, self.p), order="F")
        check(self.lib.fpb_get_dense(self.h, out.ctypes.data), self.h)
        return out


class SVDWideOnline:
    """svdwide.h:32-107.  `dat` is a flashpca_b200.Data (or None when the packed
    genotypes are given directly)."""

    def __init__(self, dat=None, block_size: int = 0, stand_method: int = 3,
                 verbose: bool = False, *, payload: np.ndarray | None = None,
                 n: int | None = None, nsnps: int | None = None, device: int = 0,
                 snp_begin: int = 0, snp_count: int = 0, meansd: np.ndarray | None = None,
                 snps_per_slab: int = 0, _handle=None):
        self.lib = _lib.load()
        self.verbose = verbose
        self.block_size = block_size  # accepted for interface parity; HBM-resident, unused
        self.stand_method = stand_method
        self.nops = 1
        self.h = ctypes.c_void_p()
        msd = None
        if meansd is None and dat is not None and dat.use_preloaded_maf:
            meansd = dat.X_meansd
        if meansd is not None:
            self._msd = np.asfortranarray(meansd, dtype=np.float64)
            msd = self._msd.ctypes.data
        if _handle is not None:
            self.h = _handle
        elif payload is not None:
            payload = np.ascontiguousarray(payload, dtype=np.uint8)
            npb = (n + 3) // 4
            if payload.size < npb * nsnps:
                raise FpbError("packed genotype buffer too small")
            check(self.lib.fpb_create(ctypes.byref(self.h), payload.ctypes.data, n, nsnps,
                                      stand_method, msd, device))
        elif snps_per_slab:
            # out-of-HBM mode: genotypes in pinned host memory, streamed slab by slab per op
            check(self.lib.fpb_create_streaming(ctypes.byref(self.h), dat.geno_filename.encode(),
                                                dat.N, snp_begin, snp_count, snps_per_slab,
                                                stand_method, msd, device))
        else:
            check(self.lib.fpb_create_from_file(ctypes.byref(self.h),
                                                dat.geno_filename.encode(), dat.N, snp_begin,
                                                snp_count, stand_method, msd, device))
        self.n = int(self.lib.fpb_rows(self.h))
        self.p = int(self.lib.fpb_nsnps(self.h))
        if dat is not None and not dat.use_preloaded_maf and snp_count == 0 and snp_begin == 0:
            dat.X_meansd = self.meansd()

    def close(self):
        if getattr(self, "h", None):
            self.lib.fpb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- svdwide.h:77-78
    def rows(self) -> int:
        return self.n

    def cols(self) -> int:
        return self.n

    @property
    def trace(self) -> float:
        """svdwide.h:37: sum of squares of the standardised matrix (known after
        staging; upstream fills it during the first perform_op)."""
        t = ctypes.c_double()
        check(self.lib.fpb_get_trace(self.h, ctypes.byref(t)), self.h)
        return t.value

    def meansd(self) -> np.ndarray:
        out = np.zeros((self.p, 2), dtype=np.float64, order="F")
        check(self.lib.fpb_get_meansd(self.h, out.ctypes.data), self.h)
        return out

    def path_info(self) -> int:
        """Compute path chosen at staging: bit mask of _lib.PATH_*."""
        return int(self.lib.fpb_path_info(self.h))

    def bed_payload(self) -> np.ndarray:
        out = np.zeros(((self.n + 3) // 4) * self.p, dtype=np.uint8)
        check(self.lib.fpb_get_bed(self.h, out.ctypes.data), self.h)
        return out

    def _call(self, fn_multi, x, rows_in, rows_out):
        vec = np.ndim(x) == 1
        xin = _f64(x, rows_in)
        k = xin.shape[1]
        y = np.empty((rows_out, k), dtype=np.float64, order="F")
        check(fn_multi(self.h, xin.ctypes.data, k, y.ctypes.data), self.h)
        self.nops += 1
        return y[:, 0].copy() if vec else y

    def perform_op(self, x_in, y_out=None):
        """svdwide.cpp:21-68: y = X X' x (un-normalised)."""
        y = self._call(self.lib.fpb_perform_op_multi, x_in, self.n, self.n)
        if y_out is not None:
            y_out[...] = y
            return y_out
        return y

    def perform_op_mat(self, x):
        """svdwide.cpp:71-118."""
        return self._call(self.lib.fpb_perform_op_multi, x, self.n, self.n)

    perform_op_multi = perform_op_mat  # svdwide.cpp:229-275

    def crossprod(self, x_in, y_out=None):
        """svdwide.cpp:122-153: y = X' x."""
        y = self._call(self.lib.fpb_crossprod_multi, x_in, self.n, self.p)
        if y_out is not None:
            y_out[...] = y
            return y_out
        return y

    def crossprod2(self, x):
        """svdwide.cpp:157-188."""
        return self._call(self.lib.fpb_crossprod_multi, x, self.n, self.p)

    def prod(self, x_in, y_out=None):
        """svdwide.cpp:193-226: y = X x."""
        y = self._call(self.lib.fpb_prod_multi, x_in, self.p, self.n)
        if y_out is not None:
            y_out[...] = y
            return y_out
        return y

    def prod3(self, x):
        """svdwide.cpp:312-343."""
        return self._call(self.lib.fpb_prod_multi, x, self.p, self.n)

    def prod2(self, x):
        """svdwide.cpp:278-309: Y = x' X  (k x nsnps)."""
        return np.asfortranarray(self.crossprod2(x).T)

    # -- whole solve (Lanczos basis resident in HBM)
    def _vector_buffer(self, nev, want_vectors, out_vectors):
        if not want_vectors:
            return None
        if out_vectors is None:
            return np.zeros((self.n, nev), order="F")
        if (out_vectors.shape != (self.n, nev) or out_vectors.dtype != np.float64
                or not out_vectors.flags.f_contiguous):
            raise FpbError("out_vectors must be an N x nev column-major float64 array")
        return out_vectors

    def pca(self, nev: int, ncv: int, maxiter: int, tol: float, want_vectors: bool = True,
            out_vectors: np.ndarray | None = None):
        """want_vectors=False skips the eigenvector download (SNP-sharded runs: every rank holds
        the same vectors, one rank needs them on the host).  out_vectors: caller-owned N x nev
        column-major buffer for the eigenvectors (a buffer that is reused between calls spares the
        page faults of 8 N nev fresh bytes: 5 ms of a 60 ms solve at N = 500,000, k = 20)."""
        evals = np.zeros(nev)
        evecs = self._vector_buffer(nev, want_vectors, out_vectors)
        nconv, nops, niter = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        check(self.lib.fpb_pca(self.h, nev, ncv, maxiter, float(tol), evals.ctypes.data,
                               evecs.ctypes.data if want_vectors else None, ctypes.byref(nconv),
                               ctypes.byref(nops), ctypes.byref(niter)), self.h)
        return dict(values=evals, vectors=evecs, nconv=nconv.value, nops=nops.value,
                    niter=niter.value)

    def pca_block(self, nev: int, tol: float, block: int = 8, max_passes: int = 40,
                  want_vectors: bool = True, out_vectors: np.ndarray | None = None):
        """Block Krylov solve (extension, include/flashpca_b200.h fpb_pca_block): same result fields
        as pca(); `npasses` = operator passes of `block` columns."""
        evals = np.zeros(nev)
        evecs = self._vector_buffer(nev, want_vectors, out_vectors)
        nconv, npasses = ctypes.c_uint32(), ctypes.c_uint32()
        check(self.lib.fpb_pca_block(self.h, nev, block, max_passes, float(tol), evals.ctypes.data,
                                     evecs.ctypes.data if want_vectors else None, ctypes.byref(nconv),
                                     ctypes.byref(npasses)), self.h)
        return dict(values=evals, vectors=evecs, nconv=nconv.value, npasses=npasses.value,
                    nops=npasses.value * block)

    def pca_residual(self, nev: int, div: float) -> np.ndarray:
        """randompca.cpp:663-703 on the solver's own eigenpairs, computed on the device:
        err_j = ||X X' u_j / div - u_j d_j||^2 (collective when SNP-sharded)."""
        err = np.zeros(nev)
        check(self.lib.fpb_pca_residual(self.h, float(div), err.ctypes.data, nev), self.h)
        return err

    def pca_phase_seconds(self) -> dict:
        buf = np.zeros(4)
        self.lib.fpb_pca_phase_times(self.h, buf.ctypes.data)
        return dict(iterate=buf[0], assemble=buf[1], download=buf[2], total=buf[3])

    def op_times_ms(self) -> np.ndarray:
        buf = np.zeros(4096, dtype=np.float32)
        m = self.lib.fpb_pca_op_times(self.h, buf.ctypes.data, buf.size)
        return buf[:m].copy()
