"""Synthetic PLINK genotypes: Balding-Nichols populations,
counter-based integer hashing so the same matrix can be produced
  * on the host with numpy (`packed_bed`, small sizes: tests, oracle input), and
  * directly in HBM by the native library (`SynthSpec.create_operator`,
    fpb_create_synthetic; bench-sized matrices never exist on the host),
bit for bit.  Ancestral MAF p_j ~ U(0.05, 0.5); K populations with F_ST:
p_kj ~ Beta(p_j (1-F)/F, (1-p_j)(1-F)/F); g_ij ~ Binomial(2, p_kj) via two
32-bit uniforms compared against floor(p_kj * 2^32); code map 2->00, 1->10,
0->11, missing->01 (data.cpp:41-45)."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


@dataclass
class SynthSpec:
    n: int
    p: int
    seed: int = 20240601
    npop: int = 25
    fst: float = 0.02
    missing_rate: float = 0.0015

    def __post_init__(self):
        rng = np.random.default_rng(self.seed)
        anc = rng.uniform(0.05, 0.5, size=self.p)
        f = self.fst
        a = anc * (1 - f) / f
        b = (1 - anc) * (1 - f) / f
        pk = rng.beta(a[None, :], b[None, :], size=(self.npop, self.p))
        pk = np.clip(pk, 1e-4, 1 - 1e-4)
        self.thresholds = np.ascontiguousarray(np.floor(pk * 4294967296.0).astype(np.uint32))
        self.pop = np.ascontiguousarray(rng.integers(0, self.npop, size=self.n).astype(np.uint8))
        self.miss_thr = int(self.missing_rate * 4294967296.0)

    # ---- host generation (numpy), SNP range [j0, j1)
    def codes(self, j0: int = 0, j1: int | None = None) -> np.ndarray:
        """(j1-j0, n) uint8 raw PLINK codes."""
        j1 = self.p if j1 is None else j1
        jj = np.arange(j0, j1, dtype=np.uint64)[:, None]
        ii = np.arange(self.n, dtype=np.uint64)[None, :]
        with np.errstate(over="ignore"):
            key = jj * np.uint64(0x100000001B3) + ii
            h = _mix64(np.uint64(self.seed) ^ _mix64(key))
            h2 = _mix64(h ^ np.uint64(0xD6E8FEB86659FD93))
        u1 = (h & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        u2 = (h >> np.uint64(32)).astype(np.uint32)
        um = (h2 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        t = self.thresholds[:, j0:j1][self.pop, :].T  # (snps, n)
        g = (u1 < t).astype(np.uint8) + (u2 < t).astype(np.uint8)
        code = np.where(g == 2, 0, np.where(g == 1, 2, 3)).astype(np.uint8)
        code[um < np.uint32(self.miss_thr)] = 1
        return code

    def packed_bed(self, j0: int = 0, j1: int | None = None) -> np.ndarray:
        """Packed payload (no 3-byte header), SNP-major, pad bits 0."""
        j1 = self.p if j1 is None else j1
        npb = (self.n + 3) // 4
        out = np.empty((j1 - j0, npb), dtype=np.uint8)
        step = max(1, int(2e7) // max(1, self.n))  # bound numpy temporaries
        for a in range(j0, j1, step):
            b = min(j1, a + step)
            code = self.codes(a, b)
            padded = np.zeros((b - a, npb * 4), dtype=np.uint8)
            padded[:, : self.n] = code
            q = padded.reshape(b - a, npb, 4)
            out[a - j0: b - j0] = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4)
                                   | (q[:, :, 3] << 6))
        return out.reshape(-1)

    def write_plink(self, stem: str) -> None:
        """bed/bim/fam fileset: bim `chr rs<j> 0 <j> A C`, fam `F<i> I<i> 0 0 0 -9`."""
        with open(stem + ".bed", "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            step = max(1, (64 << 20) // max(1, (self.n + 3) // 4))
            for j0 in range(0, self.p, step):
                f.write(self.packed_bed(j0, min(self.p, j0 + step)).tobytes())
        with open(stem + ".bim", "w") as f:
            for j in range(self.p):
                f.write("1\trs%d\t0\t%d\tA\tC\n" % (j + 1, j + 1))
        with open(stem + ".fam", "w") as f:
            for i in range(self.n):
                f.write("F%d I%d 0 0 0 -9\n" % (i + 1, i + 1))

    # ---- device generation
    def create_operator(self, stand_method: int = 3, device: int = 0, j0: int = 0,
                        j1: int | None = None):
        """SVDWideOnline over SNPs [j0, j1) generated in HBM by fpb_create_synthetic."""
        from . import _lib
        from .svdwide import SVDWideOnline
        j1 = self.p if j1 is None else j1
        lib = _lib.load()
        thr = np.ascontiguousarray(self.thresholds[:, j0:j1])
        h = ctypes.c_void_p()
        _lib.check(lib.fpb_create_synthetic(ctypes.byref(h), self.n, j1 - j0, j0,
                                            self.pop.ctypes.data, thr.ctypes.data, self.npop,
                                            self.miss_thr, self.seed, stand_method, device))
        return SVDWideOnline(_handle=h, stand_method=stand_method)
