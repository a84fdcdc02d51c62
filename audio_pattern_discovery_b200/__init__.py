"""audio_pattern_discovery_b200 -- B200-native all-pairs banded weighted DTW.

A drop-in for the hot path of dkohlsdorf/audio_pattern_discovery
(src/alignments.rs + the glue around it): hand-written CUDA for sm_100a behind the
C ABI of include/apd.h, with this package as the host-side mirror of the
reference's Rust interface.  No CPU fallback exists anywhere in the package.
"""
from .alignments import (APD_MODE_FAST, APD_MODE_STRICT, Alignment, AlignmentParams,  # noqa: F401
                         AlignmentWorkers, ApdError, Context, visible_devices)
from .discovery import Discovery  # noqa: F401
from .spectrogram import NDSequence  # noqa: F401
from .clustering import AgglomerativeClustering, ClusteringOperation, Merge  # noqa: F401,E402
