"""inplacemsdradixsort_b200 -- B200-native MSD radix sort of 64-bit key + 64-bit rid pairs,
a drop-in for the hot path of the reference's msb_64 (sort(), include/msb_64.h).

The product is the CUDA library built from csrc/ (C ABI in include/msb64_b200.h);
this package only binds it.  There is no CPU implementation: importing works
without a GPU (so that the build and the symbol table can be checked), every call
that would sort fails loudly when the library or a device is missing.
"""
from .msb64 import (  # noqa: F401
    DeviceArray, EXPORTS, MSB64_MAX_PAIRS, Msb64Error, check, device_count, fill,
    free_pinned, get_range_schedule, get_schedule, last_level_times, last_stats, launch_count, library_path, load_library,
    mafree, mamalloc, pinned, set_schedule, sort, sort_device, sort_pairs, sort_tensors,
    workspace_bytes,
)

__all__ = [
    "DeviceArray", "EXPORTS", "MSB64_MAX_PAIRS", "Msb64Error", "check", "device_count",
    "fill", "free_pinned", "get_range_schedule", "get_schedule", "last_level_times", "last_stats", "launch_count", "library_path",
    "load_library", "mafree", "mamalloc", "pinned", "set_schedule", "sort", "sort_device",
    "sort_pairs", "sort_tensors", "workspace_bytes",
]
