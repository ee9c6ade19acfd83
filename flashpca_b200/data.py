"""Host mirror of upstream class Data (data.h:81-100, data.cpp) for the PCA path:
PLINK fam/bim parsing and bed sizing.  Genotypes are never decoded on the host;
they are staged to HBM by the native library."""
from __future__ import annotations

import os

import numpy as np

from ._lib import FpbError


def _lines(path: str, what: str) -> list[bytes]:
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        raise FpbError("%s %s" % (what, path)) from e
    # data.cpp:523-532 / 600-606: a final line without '\n' is dropped
    return data.split(b"\n")[:-1]


class Data:
    """data.h:81-100.  Attributes keep upstream names."""

    def __init__(self):
        self.N = 0
        self.nsnps = 0
        self.np = 0
        self.len = 0
        self.geno_filename = ""
        self.stand_method_x = 3
        self.verbose = False
        self.use_preloaded_maf = False
        self.X_meansd = None
        self.fam_ids: list[str] = []
        self.indiv_ids: list[str] = []
        self.snp_ids: list[str] = []
        self.ref_alleles: list[str] = []
        self.alt_alleles: list[str] = []
        self.bp: list[int] = []

    def read_pheno(self, filename: str, firstcol: int) -> None:
        """data.cpp:408-413 via read_text (:504-586): sets N; every field from
        `firstcol` (1-based) on must parse as a number."""
        lines = _lines(filename, "Error reading file")
        nf = None
        for i, ln in enumerate(lines):
            tok = ln.split()
            fields = tok[firstcol - 1:]
            if nf is None:
                nf = len(fields)
            elif nf != len(fields):
                raise FpbError("Error reading file '%s': inconsistent number of columns" % filename)
            for t in fields:
                try:
                    float(t)
                except ValueError:
                    raise FpbError("Error reading file '%s', line %d: '%s' cannot be parsed as a "
                                   "number" % (filename, i + 1, t.decode()))
        self.N = len(lines)

    def read_plink_fam(self, filename: str) -> None:
        """data.cpp:639-672: FID = token 0, IID = token 1."""
        for ln in _lines(filename, "[Data::read_plink_fam] Error reading file"):
            tok = ln.split()
            self.fam_ids.append(tok[0].decode())
            self.indiv_ids.append(tok[1].decode())

    def read_plink_bim(self, filename: str) -> None:
        """data.cpp:589-637: SNP = token 1, ref = token 4, alt = token 5, bp = token 3."""
        for i, ln in enumerate(_lines(filename, "Error reading file")):
            tok = ln.split()
            self.snp_ids.append(tok[1].decode())
            self.ref_alleles.append(tok[4].decode())
            self.alt_alleles.append(tok[5].decode())
            try:
                self.bp.append(int(tok[3]))
            except ValueError:
                raise FpbError("Error reading file '%s', line %d: '%s' cannot be parsed as a "
                               "number" % (filename, i + 1, tok[3].decode()))

    def get_size(self) -> None:
        """data.cpp:150-176: len = filesize - 3 (no magic check), np = ceil(N/4),
        nsnps = len / np."""
        try:
            size = os.path.getsize(self.geno_filename)
        except OSError as e:
            raise FpbError("[Data::read_bed] Error reading file %s, error %s"
                           % (self.geno_filename, e.strerror)) from e
        self.len = size - 3
        self.np = (self.N + 3) // 4
        self.nsnps = self.len // self.np

    def prepare(self) -> None:
        """data.cpp:179-206: the stream/buffers live in the native library; only
        the zero-initialised X_meansd is kept here."""
        if not os.path.exists(self.geno_filename):
            raise FpbError("[Data::read_bed] Error reading file %s" % self.geno_filename)
        if not self.use_preloaded_maf:
            self.X_meansd = np.zeros((self.nsnps, 2), order="F")
