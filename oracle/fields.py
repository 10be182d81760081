"""TEST INFRASTRUCTURE -- exactly reproducible pseudo-random test fields (integer hash, no libm,
no RNG-stream dependence) shared by oracle/make_golden.py and the tests, so that the seeded
inputs of a golden vector need not be stored next to it."""
import numpy as np


def hash_field(n, k, salt=0):
    """(n, k) float64 array with values in [-1, 1), a pure function of (row, column, salt)."""
    i = np.arange(n, dtype=np.uint64)[:, None]
    j = np.arange(k, dtype=np.uint64)[None, :]
    h = (i * np.uint64(2654435761) + j * np.uint64(40503) + np.uint64(salt) * np.uint64(97)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    return (h & np.uint64(0xFFFFF)).astype(np.float64) / 524288.0 - 1.0
