"""CPU: Engine::disrupt (agario/engine/Engine.hpp:1263-1294) is the one place the path calls libm -- atanf in
Velocity::direction, cosf / sinf in Velocity(angle, speed) (agario/core/types.hpp:158-174).  glibc's float functions
are not correctly rounded, but they are deterministic algorithms; the CUDA path (device_math.cuh g_atanf / g_sincosf)
and the oracle's trig_mode 1 restate them operation for operation.  Here: the restatement equals the libm of this box
bit for bit on a dense sample of arguments (oracle/trig_check.c runs all 2^32), and so do the fragments' velocities."""
import ctypes as C

import numpy as np

from _helpers import fptr, oracle_lib


def _both(x, which):
    lib = oracle_lib()
    x = np.ascontiguousarray(x, np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    lib.oracle_trig_array(fptr(x), fptr(a), x.size, which, 0)
    lib.oracle_trig_array(fptr(x), fptr(b), x.size, which, 1)
    return a, b


def _same(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def test_restated_atanf_sinf_cosf_equal_libm_bit_for_bit():
    rng = np.random.default_rng(0)
    every = np.arange(0, 1 << 32, 1531, dtype=np.uint64).astype(np.uint32).view(np.float32)  # 2.8 M bit patterns over all of fp32
    ratios = (rng.uniform(-300, 300, 400000) / rng.uniform(-300, 300, 400000)).astype(np.float32)  # what direction() feeds atanf
    edges = np.array([0.4375, 0.6875, 1.1875, 2.4375, 2.0 ** 25, 2.0 ** -29, np.inf, -np.inf, np.nan, 0.0, -0.0], np.float32)
    edges = np.concatenate([np.nextafter(edges, np.float32(-np.inf)), edges, np.nextafter(edges, np.float32(np.inf)), -edges])
    a, b = _both(np.concatenate([every, ratios, edges]), 0)
    assert _same(a, b).all(), f"atanf: {int((~_same(a, b)).sum())} mismatches"
    # disrupt's angles: theta + theta + 2 pi c / n with |theta| <= 3 pi / 2, i.e. well inside the fast reduction (|y| < 120)
    angles = np.concatenate([every[np.abs(every) < 120.0], rng.uniform(-16, 16, 1000000).astype(np.float32),
                             np.array([0.0, -0.0, 2.0 ** -12, np.pi / 4, np.nan, np.inf], np.float32)])
    for which in (1, 2):
        a, b = _both(angles, which)
        ok = _same(a, b) | ~(np.abs(angles) < 120.0) & np.isnan(b)  # (the restatement has no reduce_large: NaN from 120 on)
        assert ok.all(), f"{'sinf' if which == 1 else 'cosf'}: {int((~ok).sum())} mismatches"


def test_disrupt_velocities_equal_reference_libm():
    lib = oracle_lib()
    lib.oracle_disrupt_velocity.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rng = np.random.default_rng(1)
    for _ in range(20000):
        vx, vy = rng.uniform(-300, 300, size=2).astype(np.float32)
        if rng.random() < 0.05:
            vy = np.float32(0.0)  # dx/dy = +-inf: atan = +-pi/2
        if rng.random() < 0.02:
            vx = vy = np.float32(0.0)  # 0/0 = NaN: NaN split velocity in the reference
        num = int(rng.integers(1, 14))
        c = int(rng.integers(0, num))
        a = (C.c_float(), C.c_float())
        b = (C.c_float(), C.c_float())
        lib.oracle_disrupt_velocity(C.c_float(vx), C.c_float(vy), c, num, 0, C.byref(a[0]), C.byref(a[1]))
        lib.oracle_disrupt_velocity(C.c_float(vx), C.c_float(vy), c, num, 1, C.byref(b[0]), C.byref(b[1]))
        for u, v in zip(a, b):
            assert np.float32(u.value).view(np.uint32) == np.float32(v.value).view(np.uint32) or (np.isnan(u.value) and np.isnan(v.value)), (vx, vy, c, num)
