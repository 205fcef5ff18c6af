"""B200-native SIFT engine: Python binding of the C ABI in include/sift_b200.h.

The product is libsift_b200.so (hand-written sm_100a CUDA behind `extern "C"`); this package is
only the ctypes binding used by the tests and the benchmark, mirroring the reference's two entry
points (src/sift.hh:65-75).  There is no CPU path: importing works anywhere, but creating a
context without the built library or without a B200 raises.
"""
from .api import (  # noqa: F401
    KP_DTYPE,
    SiftError,
    SiftParams,
    SiftContext,
    detect_keypoints_and_descriptors,
    match_keypoints,
    library_path,
    load_library,
    declared_symbols,
    comm_unique_id,
    comm_attach_all,
    collection_match_all,
    partition_pairs_native,
    detect_batch,
    pinned_array,
)
