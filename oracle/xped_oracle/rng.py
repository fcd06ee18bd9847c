"""TEST INFRASTRUCTURE ONLY -- counter-based uniform[-1,1) fill, identical to the device kernel
`fill_random_kernel` (xped_b200/csrc/kernels.cu): splitmix64 keyed by (seed, element index).
Stands in for Tensor::setRandom (src/Core/Tensor.cpp:33-44, uniform[-1,1] from std::mt19937) so that the CPU
oracle and the GPU fill bit-identically."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return z ^ (z >> np.uint64(31))


def uniform_fill(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = _splitmix64(np.array([seed], dtype=np.uint64))[0]
        idx = np.arange(n, dtype=np.uint64)
        r = _splitmix64(key ^ (idx * np.uint64(0xD1342543DE82EF95)))
    return (r >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0
