// device_math.cuh — scalar device helpers with the exact arithmetic of the reference value types.
//
// Contract: every numWrapper<float> expression of the reference is one fp32
// operation per C++ operator, no FMA contraction (this library is compiled with -fmad=false), IEEE
// division and square root (nvcc defaults), std::min/std::max written as the ternaries libstdc++
// uses (they propagate NaN differently from fminf/fmaxf).  Functions of an integer mass that go
// through double pow() in the reference (Engine.hpp:1296-1302) are host-built lookup tables.
#pragma once
#include <cstdint>

#include "../../include/agarcl_b200.h"

#define AG_FULL 0xffffffffu
#define AG_PI 3.14159265358979323846

namespace ag {

struct Luts {
  const float* radius;       // radius_conversion(m), core/utils.hpp:8-11
  const float* max_speed;    // Engine::max_speed(m), Engine.hpp:1300-1302
  const float* split_speed;  // Engine::split_speed(m), Engine.hpp:1296-1298
  float anti_team[AGARCL_VET_CAP + 1];  // (float)pow(1.1, n-1), Engine.hpp:567
};

__device__ __forceinline__ float fmin_std(float a, float b) { return (b < a) ? b : a; }  // std::min
__device__ __forceinline__ float fmax_std(float a, float b) { return (a < b) ? b : a; }  // std::max
// agario::clamp, core/utils.hpp:18-21
__device__ __forceinline__ float clamp_std(float x, float lo, float hi) { return fmax_std(fmin_std(x, hi), lo); }
// Engine::check_boundary_collisions on one axis, Engine.hpp:695-698 (NaN -> 0, quirk Q19)
__device__ __forceinline__ float bound_axis(float x, float r, float W) { return fmax_std(0.0f, clamp_std(x, r, W - r)); }

// the beyond-the-table paths are cold: kept out of line so that they do not dilute the hot code in the I-cache
static __device__ __noinline__ float radius_exact(uint32_t m) { return (float)sqrt((double)m / 1.0 / AG_PI); }
static __device__ __noinline__ float max_speed_exact(uint32_t m) { return (float)(300.0 / pow((double)m, 0.439)); }
static __device__ __noinline__ float split_speed_exact(uint32_t m) {
  double v = 3.0 * pow((double)(float)(300.0 / pow((double)m, 0.439)), 1.2);
  v = (130.0 < v) ? 130.0 : v;
  v = (v < 20.0) ? 20.0 : v;
  return (float)v;
}
__device__ __forceinline__ float radius_of(const Luts& T, uint32_t m) {
  return m < AGARCL_LUT_SIZE ? T.radius[m] : radius_exact(m);
}
// beyond the table the double pow of the device is used (<= 2 ulp in double, then rounded): callers flag it
__device__ __forceinline__ float max_speed_of(const Luts& T, uint32_t m, uint32_t& flags) {
  if (m < AGARCL_LUT_SIZE) return T.max_speed[m];
  flags |= AGARCL_FLAG_MASS_LUT;
  return max_speed_exact(m);
}
__device__ __forceinline__ float split_speed_of(const Luts& T, uint32_t m, uint32_t& flags) {
  if (m < AGARCL_LUT_SIZE) return T.split_speed[m];
  flags |= AGARCL_FLAG_MASS_LUT;
  return split_speed_exact(m);
}

// ---------------------------------------------------------------------------------------------
// L2 residency hints.  One env-step streams ~2 GB of observation through the 126 MB L2 while the game
// state that is actually touched is ~15 KB per instance (~61 MB for 4096 instances).  State accesses
// carry an evict_last policy and observation writes an evict_first one, so the state is still in L2
// when the next step (and the next tick) reads it instead of coming back from HBM behind the writes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_evict_last() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_evict_first() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_keep(const float4* a) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(l2_evict_last()));
  return v;
}
__device__ __forceinline__ int4 ldg_keep(const int4* a) {
  int4 v;
  asm volatile("ld.global.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(l2_evict_last()));
  return v;
}
__device__ __forceinline__ float2 ldg_keep(const float2* a) {
  float2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(a), "l"(l2_evict_last()));
  return v;
}
__device__ __forceinline__ uint32_t ldg_keep(const uint32_t* a) {
  uint32_t v;
  asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(l2_evict_last()));
  return v;
}
__device__ __forceinline__ void stg_keep(float4* a, float4 v) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(l2_evict_last()) : "memory");
}
__device__ __forceinline__ void stg_keep(int4* a, int4 v) {
  asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(l2_evict_last()) : "memory");
}
// scatter onto the observation: reductions / stores that leave L2 first
__device__ __forceinline__ void red_add_stream(int32_t* a, int v) {
  asm volatile("red.global.add.L2::cache_hint.s32 [%0], %1, %2;" :: "l"(a), "r"(v), "l"(l2_evict_first()) : "memory");
}

// static_cast<int>(float) as x86-64 cvttss2si does it: NaN / out of range -> INT_MIN (quirk Q20)
__device__ __forceinline__ int to_int_x86(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int)0x80000000;
  return (int)v;
}

// Coordinate::norm_sqr of a difference, core/types.hpp:83-87
__device__ __forceinline__ float sqr_dist(float ax, float ay, float bx, float by) {
  float dx = fabsf(ax - bx), dy = fabsf(ay - by);
  return dx * dx + dy * dy;
}
// Ball::collides_with, Ball.hpp:31-34
__device__ __forceinline__ bool collides(float ax, float ay, float ar, float bx, float by, float br) {
  float r = fmax_std(ar, br);
  return r * r >= sqr_dist(ax, ay, bx, by);
}
// Ball::can_eat, Ball.hpp:45-47 (double compare: 1.1 is inexact, integer forms differ)
__device__ __forceinline__ bool can_eat_mass(uint32_t m, uint32_t other) { return (double)m > (double)other * 1.1; }
// Cell::can_eat(const Cell&), Entities.hpp:143-151
__device__ __forceinline__ bool cell_can_eat_cell(uint32_t m, uint32_t other) { return m > 25u && can_eat_mass(m, other); }
// smallest integer mass m with (double)m > other*1.1
__device__ __forceinline__ uint32_t min_eater_mass(uint32_t other) { return (uint32_t)floor((double)other * 1.1) + 1u; }
// Cell::set_mass, Entities.hpp:171-177
__device__ __forceinline__ uint32_t floor_mass(uint32_t m) { return m > AGARCL_CELL_MIN_SIZE ? m : AGARCL_CELL_MIN_SIZE; }
__device__ __forceinline__ float vmag(float dx, float dy) { return sqrtf(dx * dx + dy * dy); }  // Velocity::magnitude

// Velocity::decelerate, core/types.hpp:208-223
__device__ __forceinline__ void decelerate(float& dx, float& dy, float decel, float dt) {
  float mag = vmag(dx, dy);
  float xr = dx / mag, yr = dy / mag;
  float ddx = xr * decel;
  if (fabsf(ddx * dt) <= fabsf(dx)) dx -= ddx * dt; else dx = 0.0f;
  float ddy = yr * decel;
  if (fabsf(ddy * dt) <= fabsf(dy)) dy -= ddy * dt; else dy = 0.0f;
}

// ---------------------------------------------------------------------------------------------
// Trigonometry of Engine::disrupt (Engine.hpp:1279-1283; Velocity::direction and Velocity(angle, speed),
// core/types.hpp:158-174).  The reference calls libm's atanf / cosf / sinf.  They are not correctly rounded, so
// "any accurate implementation" does not reproduce them; but they are deterministic algorithms, restated here
// operation for operation:
//   atanf  -- the fdlibm single-precision routine (glibc sysdeps/ieee754/flt-32/s_atanf.c): argument reduction to
//             one of four intervals, an 11-term odd/even polynomial, all in fp32
//   sinf / cosf -- the "optimized routines" implementation glibc has shipped since 2.28 (sysdeps/ieee754/flt-32/
//             s_sinf.c, s_cosf.c, sincosf.h): reduction by pi/2 and a degree-7/8 polynomial, all in fp64, one rounding
// atanf is plain fp32 arithmetic (-fmad=false: every a*b+c is two IEEE operations).  For sinf / cosf glibc selects its
// FMA build at run time on every x86-64 CPU that has the instruction (sysdeps/x86_64/fpu/multiarch): the polynomial
// and the reduction below fuse exactly where that build does (explicit fma()); the non-FMA build differs from it for
// 34 of the 2^32 fp32 arguments.  oracle/trig_check.c compares the restatement with the libm of the build box over
// ALL 2^32 arguments: no difference.
// The oracle's trig_mode 1 is the same restatement in C; its trig_mode 0 calls libm itself.
// ---------------------------------------------------------------------------------------------
// (cold code: out of line, scalar constants only -- nothing of it may cost the hot paths registers or stack)
static __device__ __noinline__ float g_atanf(float x) {
  const float hi3 = 1.5707962513e+00f, lo3 = 7.5497894159e-08f;
  const int32_t hx = __float_as_int(x), ix = hx & 0x7fffffff;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;  // NaN
    return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
  }
  float hi = 0.0f, lo = 0.0f;  // atanhi[id], atanlo[id]
  bool reduced = true;
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    reduced = false;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {  // |x| < 1.1875
      if (ix < 0x3f300000) { hi = 4.6364760399e-01f; lo = 5.0121582440e-09f; x = (2.0f * x - 1.0f) / (2.0f + x); }  // 7/16 <= |x| < 11/16
      else { hi = 7.8539812565e-01f; lo = 3.7748947079e-08f; x = (x - 1.0f) / (x + 1.0f); }                        // 11/16 <= |x| < 19/16
    } else {
      if (ix < 0x401c0000) { hi = 9.8279368877e-01f; lo = 3.4473217170e-08f; x = (x - 1.5f) / (1.0f + 1.5f * x); }  // |x| < 2.4375
      else { hi = hi3; lo = lo3; x = -1.0f / x; }
    }
  }
  const float z = x * x, w = z * z;
  const float s1 = z * (3.3333334327e-01f + w * (1.4285714924e-01f + w * (9.0908870101e-02f + w * (6.6610731184e-02f + w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
  const float s2 = w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f + w * (-5.8335702866e-02f + w * -3.6531571299e-02f))));
  if (!reduced) return x - x * (s1 + s2);
  const float r = hi - ((x * (s1 + s2) - lo) - x);
  return hx < 0 ? -r : r;
}
// sincosf.h: sinf_poly with the table row `neg` (row 1 = negated cosine coefficients) and quadrant parity n
static __device__ __forceinline__ float g_sincos_poly(double x, double x2, bool neg, int n) {
  if ((n & 1) == 0) {
    const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
    const double x3 = x * x2;
    const double s1 = fma(x2, s3c, s2c);
    const double x7 = x3 * x2;
    const double s = fma(x3, s1c, x);
    return (float)fma(x7, s1, s);
  }
  const double sg = neg ? -1.0 : 1.0;
  const double c0 = sg * 0x1p0, c1c = sg * -0x1.ffffffd0c621cp-2, c2c = sg * 0x1.55553e1068f19p-5, c3c = sg * -0x1.6c087e89a359dp-10,
               c4c = sg * 0x1.99343027bf8c3p-16;
  const double x4 = x2 * x2;
  const double c2 = fma(x2, c4c, c3c);
  const double c1 = fma(x2, c1c, c0);
  const double x6 = x4 * x2;
  const double c = fma(x4, c2c, c1);
  return (float)fma(x6, c2, c);
}
// s_sinf.c / s_cosf.c; |y| >= 120 (reduce_large there) cannot come out of Engine::disrupt, whose angles stay below 5 pi: NaN
static __device__ __noinline__ float g_sincosf(float y, int is_cos) {
  const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ffu;  // abstop12
  double x = (double)y;
  if (top < ((__float_as_uint(0x1.921fb6p-1f) >> 20) & 0x7ffu)) {  // |y| < pi/4
    if (top < ((__float_as_uint(0x1p-12f) >> 20) & 0x7ffu)) return is_cos ? 1.0f : y;
    return g_sincos_poly(x, x * x, false, is_cos);
  }
  if (top < ((__float_as_uint(120.0f) >> 20) & 0x7ffu)) {
    const double r = x * 0x1.45F306DC9C883p+23;  // 2/pi * 2^24: the quadrant ends up in bits 24..31
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = fma(-(double)n, 0x1.921FB54442D18p0, x);
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign[] = {1, -1, -1, 1}
    return g_sincos_poly(x * sgn, x * x, (n & 2) != 0, n ^ is_cos);
  }
  return __int_as_float(0x7fc00000);
}
// Velocity::direction, core/types.hpp:167-174 (quirk Q9: atan(dx/dy), then +-pi in double)
__device__ __forceinline__ float vel_direction(float dx, float dy) {
  float angle = g_atanf(dx / dy);
  if (dx < 0) {
    if (dy > 0) angle = (float)((double)angle + AG_PI);
    else angle = (float)((double)angle - AG_PI);
  }
  return angle;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based per-instance RNG for spawn points.
// draw k of instance g = word (k & 3) of philox(counter = (k >> 2, 0, g, 0), key = seed).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float philox_uniform(uint32_t seed_lo, uint32_t seed_hi, uint32_t instance, uint32_t k) {
  uint4 r = philox4x32_10(make_uint4(k >> 2, 0u, instance, 0u), make_uint2(seed_lo, seed_hi));
  uint32_t w = (k & 3u) == 0 ? r.x : (k & 3u) == 1 ? r.y : (k & 3u) == 2 ? r.z : r.w;
  return (float)(w >> 8) * (1.0f / 16777216.0f);  // [0,1), 24 bits like generate_canonical<float,24>
}

}  // namespace ag
