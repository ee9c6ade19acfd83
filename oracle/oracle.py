"""CPU ORACLE (test infrastructure, NOT product code).

numpy / ctypes side of the oracle for FlashPCA2's blocked partial
eigendecomposition path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module.  The product package ``flashpca_b200`` never does.

Three layers, each citing the upstream file:line it restates:

* ``dense_*``   -- float64 numpy restatement of decode + standardise + the dense
  ``eigen(tcrossprod(S)/div)`` criterion the reference's own tests use
  (flashpcaR/tests/testthat/test_pca.R:45-70, HapMap3/test_pca.R:121-246).
* ``COracle``   -- ctypes wrapper over ``liboracle.so`` (flashpca_oracle.c), the
  blocked operator family of svdwide.cpp over data.cpp's read_snp_block.
* ``spectra_irlm`` -- restatement of Spectra 0.8.1 ``SymEigsSolver<double,
  LARGEST_ALGE, Op>`` (third-party, pinned by the reference's Dockerfile:19-20,
  source NOT under /root/reference; call sites randompca.cpp:132-148,174-190).
  Its iteration trajectory is unpinned upstream; parity is judged on converged
  quantities only.

Parity pinning: the reference has no golden vectors for this path; the dense
criterion above is what its own tests check, and the SURVEY.md section 8c
constants (generated with an independent numpy script) are asserted in
tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

STANDARDISE_BINOM = 2   # util.h:36
STANDARDISE_BINOM2 = 3  # util.h:37
DIVISOR_NONE, DIVISOR_N1, DIVISOR_P = 0, 1, 2  # randompca.h:50-52
VAR_TOL = 1e-9  # util.h:33


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "flashpca_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB


# --------------------------------------------------------------------------
# PLINK file helpers (data.cpp:150-176, 408-413, 589-672)
# --------------------------------------------------------------------------

def count_lines(path: str) -> int:
    """Number of newline-terminated lines: data.cpp:523-532 drops a final
    unterminated line (``if(!in.eof()) lines.push_back``)."""
    with open(path, "rb") as f:
        return f.read().count(b"\n")


def read_bed_payload(path: str, n: int) -> tuple[np.ndarray, int, int]:
    """data.cpp:150-176: no magic check; len = filesize-3; np = ceil(N/4);
    nsnps = len // np.  Returns (payload bytes, np, nsnps)."""
    raw = np.fromfile(path, dtype=np.uint8)
    payload = raw[3:]
    npb = (n + 3) // 4
    nsnps = payload.size // npb
    return np.ascontiguousarray(payload[: nsnps * npb]), npb, nsnps


def read_fam_ids(path: str) -> tuple[list[str], list[str]]:
    """data.cpp:639-672: FID = token 0, IID = token 1."""
    fid, iid = [], []
    with open(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")[:-1]
    for ln in lines:
        tok = ln.split()
        fid.append(tok[0].decode())
        iid.append(tok[1].decode())
    return fid, iid


# --------------------------------------------------------------------------
# Dense numpy oracle
# --------------------------------------------------------------------------

def dense_codes(payload: np.ndarray, n: int, p: int) -> np.ndarray:
    """Raw 2-bit PLINK codes as an (N, P) uint8 matrix
    (data.cpp:128-148: code = (byte >> 2q) & 3, q = individual within byte)."""
    npb = (n + 3) // 4
    b = payload[: p * npb].reshape(p, npb)
    codes = np.empty((p, npb * 4), dtype=np.uint8)
    for q in range(4):
        codes[:, q::4] = (b >> (2 * q)) & 3
    return np.ascontiguousarray(codes[:, :n].T)


def dense_standardise(codes: np.ndarray, stand_method: int = STANDARDISE_BINOM2,
                      meansd: np.ndarray | None = None):
    """data.cpp:257-333: per-SNP mean over non-missing dosages, binomial sd,
    lookup by raw code (3->0, 2->1, 0->2 copies of the minor allele, 1->missing
    -> 0 after standardisation); sd <= VAR_TOL -> all-zero column.
    Returns (X float64 (N,P), meansd (P,2))."""
    dosage_of_code = np.array([2.0, np.nan, 1.0, 0.0])
    d = dosage_of_code[codes]
    if meansd is None:
        mean = np.nanmean(d, axis=0)
        pf = mean / 2.0
        if stand_method == STANDARDISE_BINOM:
            sd = np.sqrt(pf * (1 - pf))
        elif stand_method == STANDARDISE_BINOM2:
            sd = np.sqrt(2.0 * pf * (1 - pf))
        else:
            raise ValueError("unknown standardisation method: %d" % stand_method)
        meansd = np.stack([mean, sd], axis=1)
    mean, sd = meansd[:, 0], meansd[:, 1]
    ok = sd > VAR_TOL
    with np.errstate(divide="ignore", invalid="ignore"):
        x = (d - mean[None, :]) / sd[None, :]
    x[np.isnan(d)] = 0.0
    x[:, ~ok] = 0.0
    return x, meansd


def standardise_matrix(x: np.ndarray, method: int):
    """util.cpp:24-192 standardise(X, method): methods 0 none, 1 sd, 2 binom,
    3 binom2, 4 center (util.h:34-38); NaN = missing.  Returns (X', meansd)."""
    x = np.array(x, dtype=np.float64, order="F")
    n, p = x.shape
    na = np.isnan(x)
    cnt = (~na).sum(axis=0).astype(np.float64)
    if method in (0, 4):
        mean = np.nansum(x, axis=0) / cnt
        sd = np.ones(p)
        if method == 0:
            out = np.where(na, mean[None, :], x)
        else:
            out = np.where(na, 0.0, x - mean[None, :])
    elif method == 1:
        k = 1.0
        s = np.nansum(x - k, axis=0)
        sq = np.nansum((x - k) ** 2, axis=0)
        var = (sq - s * s / cnt) / (cnt - 1)
        mean = (s + k * cnt) / cnt
        sd = np.sqrt(var)
        with np.errstate(divide="ignore", invalid="ignore"):
            z = (x - mean[None, :]) / sd[None, :]
        out = np.where(na, 0.0, np.where(sd[None, :] > VAR_TOL, z, mean[None, :]))
    elif method in (2, 3):
        mean = np.nansum(x, axis=0) / cnt
        r = mean / 2.0
        sd = np.sqrt((1.0 if method == 2 else 2.0) * r * (1.0 - r))
        with np.errstate(divide="ignore", invalid="ignore"):
            z = (x - mean[None, :]) / sd[None, :]
        out = np.where(na, 0.0, np.where(sd[None, :] > VAR_TOL, z, mean[None, :]))
    else:
        raise ValueError("unknown standardization method")
    return np.asfortranarray(out), np.stack([mean, sd], axis=1)


def dosage_matrix(codes: np.ndarray) -> np.ndarray:
    """decode_plink (data.cpp:65-126) as doubles with NaN for missing: what
    flashpcaR hands to flashpca_internal for a genotype matrix."""
    return np.array([2.0, np.nan, 1.0, 0.0])[codes]


def dense_pca(x: np.ndarray, ndim: int, divisor: int = DIVISOR_P):
    """randompca.cpp:180-210 applied to a dense eigh of X X' (the criterion of
    test_pca.R:47,70).  Returns dict(d, U, Px, pve, trace)."""
    n, p = x.shape
    div = {DIVISOR_NONE: 1.0, DIVISOR_N1: n - 1.0, DIVISOR_P: float(p)}[divisor]
    g = x @ x.T
    w, v = np.linalg.eigh(g)
    idx = np.argsort(w)[::-1][:ndim]
    lam = w[idx]
    u = v[:, idx]
    d = lam / div
    trace = float(np.sum(x * x)) / div
    return dict(d=d, U=u, Px=u * np.sqrt(d)[None, :], pve=d / trace, trace=trace, div=div)


def dense_loadings(x: np.ndarray, u: np.ndarray, d: np.ndarray, div: float) -> np.ndarray:
    """randompca.cpp:191-204: V[:,j] = X' u_j / sqrt(d_j) / sqrt(div)."""
    return (x.T @ u) / np.sqrt(d)[None, :] / np.sqrt(div)


def sign_align(a: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """Column-wise sign alignment (HapMap3/test_pca.R:156-160 is sign-invariant)."""
    s = np.sign(np.sum(a * ref, axis=0))
    s[s == 0] = 1.0
    return a * s[None, :]


# --------------------------------------------------------------------------
# C oracle wrapper
# --------------------------------------------------------------------------

class COracle:
    """ctypes wrapper over flashpca_oracle.c (blocked operator family)."""

    def __init__(self, payload: np.ndarray, n: int, p: int,
                 stand_method: int = STANDARDISE_BINOM2,
                 meansd: np.ndarray | None = None, threads: int | None = None):
        build()
        self.lib = ctypes.CDLL(_LIB)
        L = self.lib
        dp = ctypes.POINTER(ctypes.c_double)
        L.fo_create.restype = ctypes.c_void_p
        L.fo_create.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_ulonglong,
                                ctypes.c_int, ctypes.c_void_p]
        L.fo_destroy.argtypes = [ctypes.c_void_p]
        for name in ("fo_perform_op_multi", "fo_crossprod_multi", "fo_prod_multi"):
            getattr(L, name).argtypes = [ctypes.c_void_p, dp, dp, ctypes.c_uint, ctypes.c_uint]
            getattr(L, name).restype = ctypes.c_int
        L.fo_get_trace.restype = ctypes.c_double
        L.fo_get_trace.argtypes = [ctypes.c_void_p]
        L.fo_get_meansd.argtypes = [ctypes.c_void_p, dp]
        L.fo_get_lookup.argtypes = [ctypes.c_void_p, dp]
        L.fo_block_size_from_memory.restype = ctypes.c_uint
        L.fo_block_size_from_memory.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong,
                                                ctypes.c_uint, ctypes.c_int, ctypes.c_int]
        L.fo_num_threads.restype = ctypes.c_int
        if threads is not None:
            L.fo_set_num_threads(int(threads))
        self.n, self.p = int(n), int(p)
        self.payload = np.ascontiguousarray(payload, dtype=np.uint8)
        assert self.payload.size >= ((self.n + 3) // 4) * self.p
        msd = None
        if meansd is not None:
            self._msd = np.asfortranarray(meansd, dtype=np.float64)
            msd = self._msd.ctypes.data
        self.h = L.fo_create(self.payload.ctypes.data, self.n, self.p, int(stand_method), msd)

    def __del__(self):
        h = getattr(self, "h", None)
        if h:
            self.lib.fo_destroy(h)
            self.h = None

    @property
    def threads(self) -> int:
        return int(self.lib.fo_num_threads())

    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def _call(self, fn, xin, rows_in, rows_out, block_size):
        xin = np.asarray(xin, dtype=np.float64)
        vec = xin.ndim == 1
        xin = np.asfortranarray(xin.reshape(rows_in, -1))
        k = xin.shape[1]
        y = np.zeros((rows_out, k), dtype=np.float64, order="F")
        rc = fn(self.h, self._dp(xin), self._dp(y), k, int(block_size))
        if rc != 0:
            raise RuntimeError("unknown standardisation method")
        return y[:, 0].copy() if vec else y

    def perform_op(self, x, block_size=0):
        return self._call(self.lib.fo_perform_op_multi, x, self.n, self.n, block_size)

    def crossprod(self, x, block_size=0):
        return self._call(self.lib.fo_crossprod_multi, x, self.n, self.p, block_size)

    def prod(self, v, block_size=0):
        return self._call(self.lib.fo_prod_multi, v, self.p, self.n, block_size)

    @property
    def trace(self) -> float:
        return float(self.lib.fo_get_trace(self.h))

    def meansd(self) -> np.ndarray:
        out = np.zeros((self.p, 2), dtype=np.float64, order="F")
        self.lib.fo_get_meansd(self.h, self._dp(out))
        return out

    def lookup(self) -> np.ndarray:
        out = np.zeros((4, self.p), dtype=np.float64, order="F")
        self.lib.fo_get_lookup(self.h, self._dp(out))
        return out

    def block_size_from_memory(self, ndim, do_loadings=False, memory_mb=2048) -> int:
        return int(self.lib.fo_block_size_from_memory(self.n, self.p, ndim,
                                                      int(do_loadings), memory_mb))


# --------------------------------------------------------------------------
# Spectra 0.8.1 SymEigsSolver restatement (third-party; see module docstring)
# --------------------------------------------------------------------------

def synth_packed_bed(spec, j0: int, j1: int) -> np.ndarray:
    """SNP columns [j0, j1) of a flashpca_b200.synth.SynthSpec matrix, generated on the host cores
    by liboracle.so (same counter-based hash as synth.py / the device generator).  Input
    preparation for the CPU baseline legs of bench.py."""
    lib = ctypes.CDLL(build())
    npb = (spec.n + 3) // 4
    out = np.empty((j1 - j0) * npb, dtype=np.uint8)
    thr = np.ascontiguousarray(spec.thresholds, dtype=np.uint32)
    pop = np.ascontiguousarray(spec.pop, dtype=np.uint8)
    lib.fo_synth_bed.restype = None
    lib.fo_synth_bed.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_ulonglong,
                                 ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_uint, ctypes.c_ulonglong]
    lib.fo_synth_bed(out.ctypes.data, spec.n, j0, j1, spec.p, pop.ctypes.data, thr.ctypes.data,
                     spec.miss_thr, spec.seed)
    return out


def simple_random_vec(n: int, seed: int = 0) -> np.ndarray:
    """Spectra ``SimpleRandom<double>``: Park-Miller LCG a=16807, m=2^31-1,
    seed 0 -> 1, value/m - 0.5."""
    m, a = 2147483647, 16807
    r = (seed & m) if seed else 1
    out = np.empty(n)
    for i in range(n):
        r = (a * r) % m
        out[i] = r / m - 0.5
    return out


def _tridiag_shift_qr(h: np.ndarray, mu: float):
    hs = h - mu * np.eye(h.shape[0])
    q, r = np.linalg.qr(hs)
    return q, r @ q + mu * np.eye(h.shape[0])


def spectra_irlm(op, n: int, nev: int, ncv: int, maxit: int = 500, tol: float = 1e-6,
                 v0: np.ndarray | None = None):
    """Implicitly restarted Lanczos as Spectra 0.8.1 runs it for flashpca
    (randompca.cpp:174-178: nev=ndim, ncv=2*ndim+1, LARGEST_ALGE).
    ``op(x) -> A x``.  Returns dict(values, vectors, nops, niter, nconv)."""
    eps = np.finfo(np.float64).eps
    near0 = np.finfo(np.float64).tiny * 10
    eps23 = eps ** (2.0 / 3.0)
    V = np.zeros((n, ncv))
    H = np.zeros((ncv, ncv))
    nops = 0

    # init(): start residual from SimpleRandom(0)
    r0 = simple_random_vec(n, 0) if v0 is None else np.asarray(v0, dtype=np.float64)
    v = r0 / np.linalg.norm(r0)
    w = op(v)
    nops += 1
    H[0, 0] = v @ w
    f = w - v * H[0, 0]
    V[:, 0] = v
    if np.max(np.abs(f)) < eps:
        f[:] = 0

    def factorize_from(from_k, to_m, fk):
        nonlocal f, nops
        if to_m <= from_k:
            return
        f = fk.copy()
        beta = np.linalg.norm(f)
        H[:, from_k:] = 0
        H[from_k:, :from_k] = 0
        for i in range(from_k, to_m):
            restart = False
            if beta < near0:
                f = simple_random_vec(n, 2 * i)
                Vi = V[:, :i]
                f = f - Vi @ (Vi.T @ f)
                beta = np.linalg.norm(f)
                restart = True
            V[:, i] = f / beta
            H[i, i - 1] = 0.0 if restart else beta
            w = op(V[:, i])
            nops += 1
            Hii = V[:, i] @ w
            H[i - 1, i] = H[i, i - 1]
            H[i, i] = Hii
            if restart:
                f = w - Hii * V[:, i]
            else:
                f = w - H[i, i - 1] * V[:, i - 1] - Hii * V[:, i]
            beta = np.linalg.norm(f)
            Vi = V[:, : i + 1]
            Vf = Vi.T @ f
            count = 0
            while count < 5 and np.max(np.abs(Vf)) > eps * beta:
                if beta < near0:
                    f[:] = 0
                    beta = 0.0
                    break
                f = f - Vi @ Vf
                H[i - 1, i] += Vf[i - 1]
                H[i, i - 1] = H[i - 1, i]
                H[i, i] += Vf[i]
                beta = np.linalg.norm(f)
                Vf = Vi.T @ f
                count += 1

    def retrieve_ritzpair():
        evals, evecs = np.linalg.eigh(H)
        ind = np.argsort(evals)[::-1]  # LARGEST_ALGE
        return evals[ind], evecs[ncv - 1, ind], evecs[:, ind[:nev]]

    factorize_from(1, ncv, f)
    ritz_val, ritz_est, ritz_vec = retrieve_ritzpair()
    nconv, it = 0, 0
    conv = np.zeros(nev, dtype=bool)
    for it in range(maxit):
        thresh = tol * np.maximum(eps23, np.abs(ritz_val[:nev]))
        resid = np.abs(ritz_est[:nev]) * np.linalg.norm(f)
        conv = resid < thresh
        nconv = int(conv.sum())
        if nconv >= nev:
            break
        nev_new = nev + int(np.sum(np.abs(ritz_est[nev:]) < near0))
        nev_new += min(nconv, (ncv - nev_new) // 2)
        if nev_new == 1 and ncv >= 6:
            nev_new = ncv // 2
        elif nev_new == 1 and ncv > 2:
            nev_new = 2
        nev_new = min(nev_new, ncv - 1)
        k = nev_new
        Q = np.eye(ncv)
        for i in range(k, ncv):
            q, hn = _tridiag_shift_qr(H, ritz_val[i])
            Q = Q @ q
            H[:, :] = hn
        Vs = V @ Q[:, : k + 1]
        V[:, : k + 1] = Vs
        fk = f * Q[ncv - 1, k - 1] + V[:, k] * H[k, k - 1]
        factorize_from(k, ncv, fk)
        ritz_val, ritz_est, ritz_vec = retrieve_ritzpair()
    order = np.argsort(ritz_val[:nev])[::-1]
    vals = ritz_val[:nev][order]
    vecs = (V @ ritz_vec)[:, order]
    conv = conv[order]
    return dict(values=vals[conv], vectors=vecs[:, conv], nops=nops, niter=it + 1,
                nconv=nconv)


def oracle_pca(payload, n, p, ndim, stand_method=STANDARDISE_BINOM2, divisor=DIVISOR_P,
               maxiter=500, tol=1e-6, block_size=0, do_loadings=False):
    """randompca.cpp:168-218 on the C oracle operator + the IRLM restatement."""
    orc = COracle(payload, n, p, stand_method)
    res = spectra_irlm(lambda x: orc.perform_op(x, block_size), n, ndim, 2 * ndim + 1,
                       maxiter, tol)
    if res["nconv"] < ndim:
        raise RuntimeError("Spectra eigen-decomposition was not successful")
    div = {DIVISOR_NONE: 1.0, DIVISOR_N1: n - 1.0, DIVISOR_P: float(p)}[divisor]
    d = res["values"] / div
    u = res["vectors"]
    out = dict(d=d, U=u, Px=u * np.sqrt(d)[None, :], trace=orc.trace / div,
               meansd=orc.meansd(), nops=res["nops"])
    out["pve"] = d / out["trace"]
    if do_loadings:
        out["V"] = orc.crossprod(u, block_size) / np.sqrt(d)[None, :] / np.sqrt(div)
    return out
