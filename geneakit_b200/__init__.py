"""geneakit_b200: B200-native dense kinship (gen.phi / gen.phiMean / gen.gc) behind the
reference's own API.  Everything computes in libgeneakit_b200.so (hand-written sm_100a CUDA);
importing this package fails loudly when that library has not been built."""
from . import cgeneakit                                   # noqa: F401
from .create import genealogy                             # noqa: F401
from .identify import pro, founder                        # noqa: F401
from .compute import phi, phiMean, phiOver, phiCI, f, gc, phi_device   # noqa: F401
from .synth import synthetic_pedigree                     # noqa: F401

__all__ = ["genealogy", "pro", "founder", "phi", "phiMean", "phiOver", "phiCI", "f", "gc", "phi_device",
           "synthetic_pedigree", "cgeneakit"]
