"""sabreur_b200 -- B200-native demultiplexing hot path behind sabreur's interface.

The product is the C-ABI shared library (include/sabreur_gpu.h, built from sabreur_b200/csrc/ for
sm_100a).  This package is its thin host-side mirror for Python callers, tests and the benchmark:

    demux.Demux            context + streaming of a byte stream through pinned slots
    demux.se_demux/pe_demux the reference's entry points (src/demux.rs:13-20, 71-79), same argument meaning
    synth                  synthetic workloads of BASELINE.json's shapes

There is no CPU fallback anywhere in this package.
"""
from . import _ffi  # noqa: F401

__all__ = ["_ffi"]
__version__ = "0.1.0"
