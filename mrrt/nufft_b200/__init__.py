"""mrrt.nufft_b200 -- B200-native (sm_100a) NUFFT hot path, drop-in for the
``NufftBase`` operator of mritools/mrrt.nufft (table and sparse modes, 1-3 D, single and
double precision).  Host code is Python; all transforms run in pre-built CUDA loaded
through a C ABI (include/b200nufft.h).  No CPU fallback."""
from ._kernels import BeattyKernel, KaiserBesselKernel, kaiser_bessel, kaiser_bessel_ft
from ._nufft import NufftBase, nufft_adj, nufft_forward
from ._sense import SenseNufft
from ._toeplitz import ToeplitzNorm
from ._sharded import CoilShardedNufft, SampleShardedNufft, shard_range
from ._slab import SlabShardedNufft

__all__ = [
    "NufftBase",
    "nufft_forward",
    "nufft_adj",
    "BeattyKernel",
    "KaiserBesselKernel",
    "kaiser_bessel",
    "kaiser_bessel_ft",
    "SampleShardedNufft",
    "SlabShardedNufft",
    "CoilShardedNufft",
    "shard_range",
    "SenseNufft",
    "ToeplitzNorm",
]
__version__ = "0.1.0"
