"""CPU: oracle.c's restatement of libstdc++'s std::sort (se_std_sort: introsort + final insertion sort, heap-sort fallback)
against the REAL std::sort on the element type and comparator of PrecisionCollisionDetection::solve
(agario/utils/collision_detection.hpp:29-31), through the compiled reference harness (oracle/ref_harness.cpp ref_std_sort_pairs).
std::sort is not stable: what is compared is the final position of every (id, y) pair, ties included, since the order of cells with
equal y decides where the scan of a strip stops (quirk Q7).  Skipped where oracle/_ref has not been built."""
import ctypes as C

import numpy as np
import pytest

from _helpers import oracle_lib, ref_lib


def _both(ys):
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    n = len(ys)
    out = []
    for fn in (oracle_lib().oracle_std_sort_pairs, ref_lib().ref_std_sort_pairs):
        ids = np.arange(n, dtype=np.int32)
        y = ys.copy()
        fn(ids.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int(n))
        out.append((ids, y))
    return out


def _killer(n):
    """median-of-three killer (Musser 1997) for even n: drives the partitioning quadratic, i.e. into the depth-limit fallback"""
    k = n // 2
    a = [0] * n
    for i in range(1, k + 1):
        if i % 2 == 1:
            a[i - 1] = i
            a[i] = k + i
        a[k + i - 1] = 2 * i
    return np.array(a, dtype=np.float32)


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built (no reference tree)")
def test_restated_std_sort_equals_libstdcxx():
    rng = np.random.default_rng(20261018)
    cases = []
    for n in list(range(0, 70)) + [95, 96, 128, 129, 200, 255, 256, 257, 400, 1000]:
        for levels in (1, 2, 3, 5, 17, 10 ** 6):  # number of distinct keys: from "all equal" to "no ties"
            for _ in range(6):
                cases.append(rng.integers(0, levels, size=n).astype(np.float32))
        cases.append(np.arange(n, dtype=np.float32))            # sorted
        cases.append(np.arange(n, dtype=np.float32)[::-1])      # reversed
        cases.append(np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(np.float32))  # organ pipe
        cases.append((np.arange(n) % 4).astype(np.float32))     # sawtooth of ties
        if n % 2 == 0 and n >= 4:
            cases.append(_killer(n))
    # long adversarial inputs: the heap-sort fallback must be reached (2 * lg(n) partitioning levels are not enough)
    for n in (2048, 4096):
        cases.append(_killer(n))
        cases.append(np.floor(_killer(n) / 7))
    ties_moved = 0
    heap_calls = C.c_int.in_dll(oracle_lib(), "oracle_std_sort_heap_calls")
    heap_calls.value = 0
    for ys in cases:
        (oi, oy), (ri, ry) = _both(ys)
        assert np.array_equal(oy, ry)
        assert np.array_equal(oi, ri), f"n={len(ys)}: tie order differs at {np.argwhere(oi != ri)[:5].ravel().tolist()}"
        assert np.all(np.diff(oy) >= 0)
        # (how often the unstable sort really reorders equal keys: the reason this test exists)
        stable = np.argsort(ys, kind="stable").astype(np.int32)
        ties_moved += int(not np.array_equal(stable, oi))
    assert heap_calls.value > 0, "the depth-limit fallback (heap sort) was never reached"
    assert ties_moved > 100, "the inputs never exercised the instability of std::sort"
