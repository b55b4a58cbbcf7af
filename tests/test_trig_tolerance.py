"""CPU: the ONE floating-point tolerance of the path.  Engine::disrupt (Engine.hpp:1263-1294) derives the
fragments' split velocities from glibc atanf/cosf/sinf; the CUDA path (and oracle trig_mode 1) uses a portable
IEEE-double algorithm instead.  Stated tolerance: |delta split-velocity component| <= 2.5e-4 world units/s
(speed is max_speed(25) = 73.0 u/s, i.e. a relative 3.4e-6: a few fp32 ulps of the angle)."""
import ctypes as C

import numpy as np

from _helpers import oracle_lib

TOL = 2.5e-4


def test_disrupt_velocity_tolerance():
    lib = oracle_lib()
    lib.oracle_disrupt_velocity.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rng = np.random.default_rng(0)
    worst = 0.0
    exact = total = 0
    for _ in range(20000):
        vx, vy = rng.uniform(-300, 300, size=2).astype(np.float32)
        if rng.random() < 0.05:
            vy = np.float32(0.0)  # dx/dy = +-inf: atan = +-pi/2
        num = int(rng.integers(1, 14))
        c = int(rng.integers(0, num))
        a = (C.c_float(), C.c_float())
        b = (C.c_float(), C.c_float())
        lib.oracle_disrupt_velocity(C.c_float(vx), C.c_float(vy), c, num, 0, C.byref(a[0]), C.byref(a[1]))
        lib.oracle_disrupt_velocity(C.c_float(vx), C.c_float(vy), c, num, 1, C.byref(b[0]), C.byref(b[1]))
        d = max(abs(a[0].value - b[0].value), abs(a[1].value - b[1].value))
        worst = max(worst, d)
        exact += d == 0.0
        total += 1
    print(f"disrupt split velocity: libm vs portable trig: max |delta| = {worst:.3e} u/s, bit-identical in {100 * exact / total:.1f} % of draws")
    assert worst <= TOL


def test_zero_velocity_is_nan_in_both():
    # a cell with zero velocity: dx/dy = 0/0 = NaN -> NaN split velocity in the reference; the portable path must agree
    lib = oracle_lib()
    lib.oracle_disrupt_velocity.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for mode in (0, 1):
        x, y = C.c_float(), C.c_float()
        lib.oracle_disrupt_velocity(C.c_float(0.0), C.c_float(0.0), 1, 3, mode, C.byref(x), C.byref(y))
        assert np.isnan(x.value) and np.isnan(y.value)
