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
// Portable trigonometry for Engine::disrupt (Engine.hpp:1279-1283; Velocity::direction and
// Velocity(angle, speed), core/types.hpp:158-174).  The reference calls glibc atanf/cosf/sinf, which
// are not correctly rounded (measured: ~1 % of results differ from the correctly rounded value by
// one ulp), so they cannot be matched bit-for-bit on a GPU.  This is the "stated fp32 tolerance"
// site of the path.  Both the device and oracle.c (trig_mode 1) use the SAME fixed algorithm in IEEE
// double arithmetic without contraction, so GPU == oracle bit-for-bit, and both are within 1 ulp
// (fp32) of the reference.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double p_atan(double q) {
  if (q != q) return q;
  bool neg = q < 0.0;
  double t = neg ? -q : q;
  bool inv = t > 1.0;
  if (inv) t = 1.0 / t;  // 1/inf = 0
  bool shift = t > 0.4142135623730950488;
  if (shift) t = (t - 1.0) / (t + 1.0);
  double z = t * t, s = 0.0;
  for (int k = 24; k >= 0; --k) s = s * z + ((k & 1) ? -1.0 : 1.0) / (double)(2 * k + 1);
  double r = t * s;
  if (shift) r = 0.78539816339744830962 + r;
  if (inv) r = 1.57079632679489661923 - r;
  return neg ? -r : r;
}
__device__ __forceinline__ void p_sincos(double x, double* sn, double* cs) {
  if (!(x > -1.0e6 && x < 1.0e6)) { *sn = __longlong_as_double(0x7ff8000000000000LL); *cs = *sn; return; }
  double k = rint(x * 0.63661977236758134308);
  double r = (x - k * 1.57079632673412561417e+00) - k * 6.07710050650619224932e-11;  // Cody-Waite, hi has 33 bits
  double z = r * r;
  double ps = 0.0, pc = 0.0;
  // Taylor: sin r = r * sum (-z)^j / (2j+1)!,  cos r = sum (-z)^j / (2j)!
  const double fs[10] = {1.0, 6.0, 120.0, 5040.0, 362880.0, 39916800.0, 6227020800.0, 1307674368000.0,
                         355687428096000.0, 121645100408832000.0};
  const double fc[10] = {1.0, 2.0, 24.0, 720.0, 40320.0, 3628800.0, 479001600.0, 87178291200.0,
                         20922789888000.0, 6402373705728000.0};
  for (int j = 9; j >= 0; --j) {
    double sg = (j & 1) ? -1.0 : 1.0;
    ps = ps * z + sg / fs[j];
    pc = pc * z + sg / fc[j];
  }
  double s0 = r * ps, c0 = pc;
  long long q = (long long)k;
  int m = (int)(((q % 4) + 4) % 4);
  double s1 = (m == 0) ? s0 : (m == 1) ? c0 : (m == 2) ? -s0 : -c0;
  double c1 = (m == 0) ? c0 : (m == 1) ? -s0 : (m == 2) ? -c0 : s0;
  *sn = s1;
  *cs = c1;
}
// Velocity::direction, core/types.hpp:167-174 (quirk Q9: atan(dx/dy), then +-pi in double)
__device__ __forceinline__ float vel_direction(float dx, float dy) {
  float angle = (float)p_atan((double)(dx / dy));
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
