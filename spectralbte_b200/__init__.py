"""spectralbte_b200: B200 (sm_100a) collision hot path of SpectralBTE behind the reference's own
C entry points. CUDA kernels + C ABI in csrc/ (libsbte_b200.so), host-side mirror in api.py."""
from .api import (K2_AUTO, K2_BATCH, K2_GENERIC, K2_STREAM, K2_STREAM_DEEP, Collisions, DeviceArray,  # noqa: F401
                  Slab, velocity_grids, weights_filename)
