"""flashpca_b200 -- B200-native implementation of FlashPCA2's blocked partial
eigendecomposition hot path (RandomPCA::pca_fast -> Spectra IRLM ->
SVDWideOnline::perform_op -> Data::read_snp_block).

The product is the CUDA library behind include/flashpca_b200.h plus the C++ host
layer in flashpca_b200/host (the `flashpca` command line).  This Python package
is the thin host-side mirror of the same operator interface used by the tests
and bench.py; it binds the C ABI with ctypes and has no CPU fallback.
"""
from .data import Data
from .randompca import RandomPCA
from .svdwide import SVDWide, SVDWideOnline

STANDARDISE_BINOM = 2   # util.h:36
STANDARDISE_BINOM2 = 3  # util.h:37
DIVISOR_NONE, DIVISOR_N1, DIVISOR_P = 0, 1, 2  # randompca.h:50-52

__all__ = ["Data", "SVDWide", "SVDWideOnline", "RandomPCA", "STANDARDISE_BINOM", "STANDARDISE_BINOM2",
           "DIVISOR_NONE", "DIVISOR_N1", "DIVISOR_P"]
