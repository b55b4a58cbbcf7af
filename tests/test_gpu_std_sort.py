"""GPU: the device's restatement of libstdc++'s std::sort (k_step strip_std_sort, through agarcl_selftest_std_sort) against the
oracle's (oracle.c se_std_sort), which tests/test_std_sort.py pins against the real std::sort: the same final place for every
index, equal keys included -- in k_step the order of cells with equal y decides where the scan of a collision strip stops."""
import ctypes as C

import numpy as np
import pytest

from _helpers import oracle_lib
from agarcl_b200 import _lib

pytestmark = pytest.mark.gpu


def _killer(n):
    k = n // 2
    a = [0] * n
    for i in range(1, k + 1):
        if i % 2 == 1:
            a[i - 1] = i
            a[i] = k + i
        a[k + i - 1] = 2 * i
    return np.array(a, dtype=np.float32)


def test_device_std_sort_equals_oracle_restatement():
    rng = np.random.default_rng(7)
    cases = []
    for n in list(range(0, 40)) + [63, 64, 65, 100, 128, 200, 255, 256]:
        for levels in (1, 2, 3, 5, 17, 10 ** 6):
            for _ in range(3):
                cases.append(rng.integers(0, levels, size=n).astype(np.float32))
        cases.append(np.arange(n, dtype=np.float32)[::-1].copy())
        cases.append((np.arange(n) % 4).astype(np.float32))
        if n % 2 == 0 and n >= 4:
            cases.append(_killer(n))
    for n in (2048, 4096):  # the heap-sort fallback
        cases.append(_killer(n))
        cases.append(np.floor(_killer(n) / 7).astype(np.float32))
    heap_calls = C.c_int.in_dll(oracle_lib(), "oracle_std_sort_heap_calls")
    heap_calls.value = 0
    moved = 0
    for ys in cases:
        n = len(ys)
        ids = np.arange(n, dtype=np.int32)
        y = ys.copy()
        oracle_lib().oracle_std_sort_pairs(ids.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int(n))
        out = np.zeros(max(n, 1), dtype=np.uint16)
        _lib.check(_lib.lib().agarcl_selftest_std_sort(ys.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(out[:n].astype(np.int32), ids), f"n={n}: device and oracle leave equal keys in different places"
        moved += int(not np.array_equal(np.argsort(ys, kind="stable").astype(np.int32), ids))
    assert heap_calls.value > 0 and moved > 50
