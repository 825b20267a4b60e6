"""computeOrbDescriptor's `(float)cos(angle)` / `(float)sin(angle)` (reference Features/orbextractor.cpp:45-46) are libm's cosf / sinf
(float overloads under `using namespace std`), so the steering uses glibc's single-precision algorithm, not a rounded double cosine.
csrc/glibc_sincosf.h restates that algorithm for host and device:
  * CPU: the host restatement equals this machine's libm on EVERY float in [0, 6.2832] (the descriptor's whole input domain: fastAtan2
    degrees in [0, 360] times pi/180), 1.09e9 inputs, in parallel over the host cores;
  * GPU: the device restatement equals the host restatement on 48 M inputs spread over the domain.
"""
import struct
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
import pytest


def _bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def test_host_restatement_is_libm_on_the_whole_domain(ob):
    hi = _bits(6.2832)
    workers = max(1, min(16, os.cpu_count() or 1))
    edges = np.linspace(0, hi + 1, 8 * workers + 1).astype(np.int64)
    with ThreadPoolExecutor(workers) as ex:           # ctypes releases the GIL
        diffs = list(ex.map(lambda k: ob.selftest_sincosf(int(edges[k]), int(edges[k + 1] - 1))[0], range(len(edges) - 1)))
    assert sum(diffs) == 0, f"{sum(diffs)} of {hi + 1} floats differ from libm sinf / cosf"


def test_cosf_is_not_the_rounded_double_cosine(ob):
    """The distinction is real: on a window of the domain thousands of inputs differ in the last bit."""
    lo = _bits(1.0)
    _, v = ob.selftest_sincosf(lo, lo + 199_999, want_values=True)
    x = np.arange(lo, lo + 200_000, dtype=np.uint32).view(np.float32).astype(np.float64)
    n = int(np.sum(v[:, 0] != np.sin(x).astype(np.float32)) + np.sum(v[:, 1] != np.cos(x).astype(np.float32)))
    assert n > 1000


@pytest.mark.gpu
def test_device_restatement_equals_host(ob):
    ctx = ob.Context(max_frames=1)
    try:
        hi = _bits(6.2832)
        rng = np.random.default_rng(0)
        starts = [0, _bits(2.0 ** -12) - 1000, _bits(0.78539816) - 500_000, hi - 999_999] + [int(s) for s in rng.integers(0, hi - 2_000_000, 44)]
        for s in starts:
            e = min(s + 999_999, hi)
            dev = ctx.selftest_sincosf_device(s, e)
            nd, host = ob.selftest_sincosf(s, e, want_values=True)
            assert nd == 0 and dev.tobytes() == host.tobytes(), f"window at {s:#x}"
    finally:
        ctx.close()
