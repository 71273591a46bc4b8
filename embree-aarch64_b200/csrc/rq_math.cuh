// FP32 ray/triangle tests in the reference's exact operation order.
//
// Numerics spec followed (reference, read-only):
//   dot(a,b)   = madd(a.x,b.x, madd(a.y,b.y, a.z*b.z))            common/math/vec3.h:204
//   cross(a,b) = (msub(a.y,b.z, a.z*b.y), msub(a.z,b.x, a.x*b.z), msub(a.x,b.y, a.y*b.x))
//                                                                 common/math/vec3.h:209, math.h:367-373
//   madd/msub are single-rounding FMAs on the AVX2/AVX-512 targets   common/simd/vfloat8_avx.h:404-410
//   Moeller-Trumbore on precomputed edges    kernels/geometry/triangle_intersector_moeller.h:62-103,290-327
//   Pluecker (watertight, RTC_SCENE_FLAG_ROBUST)  kernels/geometry/triangle_intersector_pluecker.h:61-108
//   stable_triangle_normal                          common/math/vec3.h:210-222
// Every multiply/add below is written with an explicit rounding intrinsic so that neither nvcc
// (-fmad) nor a host compiler (-ffp-contract) can re-associate or fuse differently from the
// reference: which side of a shared edge wins depends on these roundings.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#  define RQ_HD __host__ __device__ __forceinline__
#else
#  define RQ_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#  define rq_fma(a, b, c) __fmaf_rn((a), (b), (c))
#  define rq_mul(a, b)    __fmul_rn((a), (b))
#  define rq_add(a, b)    __fadd_rn((a), (b))
#  define rq_sub(a, b)    __fsub_rn((a), (b))
#  define rq_div(a, b)    __fdiv_rn((a), (b))
#  define rq_rcp(a)       __frcp_rn((a))
#else
   // host build (unit tests only): compile with -ffp-contract=off
#  define rq_fma(a, b, c) fmaf((a), (b), (c))
#  define rq_mul(a, b)    ((a) * (b))
#  define rq_add(a, b)    ((a) + (b))
#  define rq_sub(a, b)    ((a) - (b))
#  define rq_div(a, b)    ((a) / (b))
#  define rq_rcp(a)       (1.0f / (a))
#endif

struct RQVec3 { float x, y, z; };

RQ_HD RQVec3 rq_v3(float x, float y, float z) { RQVec3 r; r.x = x; r.y = y; r.z = z; return r; }
RQ_HD RQVec3 rq_vsub(RQVec3 a, RQVec3 b) { return rq_v3(rq_sub(a.x, b.x), rq_sub(a.y, b.y), rq_sub(a.z, b.z)); }
RQ_HD RQVec3 rq_vadd(RQVec3 a, RQVec3 b) { return rq_v3(rq_add(a.x, b.x), rq_add(a.y, b.y), rq_add(a.z, b.z)); }
RQ_HD float rq_dot(RQVec3 a, RQVec3 b) { return rq_fma(a.x, b.x, rq_fma(a.y, b.y, rq_mul(a.z, b.z))); }
RQ_HD float rq_msub(float a, float b, float c) { return rq_fma(a, b, -c); }
RQ_HD RQVec3 rq_cross(RQVec3 a, RQVec3 b) {
  return rq_v3(rq_msub(a.y, b.z, rq_mul(a.z, b.y)),
               rq_msub(a.z, b.x, rq_mul(a.x, b.z)),
               rq_msub(a.x, b.y, rq_mul(a.y, b.x)));
}
RQ_HD float rq_bits2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
RQ_HD uint32_t rq_f2bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { uint32_t u; float f; } c; c.f = f; return c.u;
#endif
}
RQ_HD float rq_xorsign(float a, uint32_t sgn) { return rq_bits2f(rq_f2bits(a) ^ sgn); }

struct RQTriHit {
  float t, u, v;
  RQVec3 Ng;
};

// Default scene flags: Moeller-Trumbore with edges e1 = v0-v1, e2 = v2-v0 and Ng = cross(e2,e1).
// Accept rule (moeller.h:85-97):  den != 0, U >= 0, V >= 0, U+V <= |den|,  |den|*tnear < T <= |den|*tfar.
// u,v,t = {U,V,T} * rcp(|den|) as in the reference (moeller.h:30-36); its rcp is an estimate + one Newton step (1-2 ulp), here the
// correctly rounded reciprocal.  (Three IEEE divisions instead cost 4.2 % of all warp instructions of the closest-hit kernel at
// 1.5-4 active lanes: profiles/r01z_ncu_source.txt.)
RQ_HD bool rq_moeller(RQVec3 O, RQVec3 D, float tnear, float tfar,
                      RQVec3 v0, RQVec3 v1, RQVec3 v2, RQTriHit& hit) {
  const RQVec3 e1 = rq_vsub(v0, v1);
  const RQVec3 e2 = rq_vsub(v2, v0);
  const RQVec3 Ng = rq_cross(e2, e1);
  const RQVec3 C = rq_vsub(v0, O);
  const RQVec3 R = rq_cross(C, D);
  const float den = rq_dot(Ng, D);
  const float absDen = fabsf(den);
  const uint32_t sgn = rq_f2bits(den) & 0x80000000u;
  const float U = rq_xorsign(rq_dot(R, e2), sgn);
  const float V = rq_xorsign(rq_dot(R, e1), sgn);
  if (!((den != 0.0f) & (U >= 0.0f) & (V >= 0.0f) & (rq_add(U, V) <= absDen))) return false;
  const float T = rq_xorsign(rq_dot(Ng, C), sgn);
  if (!((rq_mul(absDen, tnear) < T) & (T <= rq_mul(absDen, tfar)))) return false;
  const float rcpAbsDen = rq_rcp(absDen);
  hit.t = rq_mul(T, rcpAbsDen);
  hit.u = rq_mul(U, rcpAbsDen);
  hit.v = rq_mul(V, rcpAbsDen);
  hit.Ng = Ng;
  return true;
}

// common/math/vec3.h:210-222
RQ_HD RQVec3 rq_stable_normal(RQVec3 a, RQVec3 b, RQVec3 c) {
  const float ab_x = rq_mul(a.z, b.y), ab_y = rq_mul(a.x, b.z), ab_z = rq_mul(a.y, b.x);
  const float bc_x = rq_mul(b.z, c.y), bc_y = rq_mul(b.x, c.z), bc_z = rq_mul(b.y, c.x);
  const RQVec3 cab = rq_v3(rq_msub(a.y, b.z, ab_x), rq_msub(a.z, b.x, ab_y), rq_msub(a.x, b.y, ab_z));
  const RQVec3 cbc = rq_v3(rq_msub(b.y, c.z, bc_x), rq_msub(b.z, c.x, bc_y), rq_msub(b.x, c.y, bc_z));
  return rq_v3(fabsf(ab_x) < fabsf(bc_x) ? cab.x : cbc.x,
               fabsf(ab_y) < fabsf(bc_y) ? cab.y : cbc.y,
               fabsf(ab_z) < fabsf(bc_z) ? cab.z : cbc.z);
}

// RTC_SCENE_FLAG_ROBUST: Pluecker test on vertices translated to the ray origin.
// Accept rule (pluecker.h:84-104): min(U,V,W) >= -eps or max(U,V,W) <= eps with eps = ulp*|U+V+W|,
// den != 0, tnear <= t <= tfar with t = T/den;  u = U/UVW, v = V/UVW (0 when |UVW| < 1e-18).
RQ_HD bool rq_pluecker(RQVec3 O, RQVec3 D, float tnear, float tfar,
                       RQVec3 tv0, RQVec3 tv1, RQVec3 tv2, RQTriHit& hit) {
  const RQVec3 v0 = rq_vsub(tv0, O), v1 = rq_vsub(tv1, O), v2 = rq_vsub(tv2, O);
  const RQVec3 e0 = rq_vsub(v2, v0), e1 = rq_vsub(v0, v1), e2 = rq_vsub(v1, v2);
  const float U = rq_dot(rq_cross(e0, rq_vadd(v2, v0)), D);
  const float V = rq_dot(rq_cross(e1, rq_vadd(v0, v1)), D);
  const float W = rq_dot(rq_cross(e2, rq_vadd(v1, v2)), D);
  const float UVW = rq_add(rq_add(U, V), W);
  const float eps = rq_mul(1.1920928955078125e-07f, fabsf(UVW));
  const float mn = fminf(fminf(U, V), W), mx = fmaxf(fmaxf(U, V), W);
  if (!((mn >= -eps) | (mx <= eps))) return false;
  const RQVec3 Ng = rq_stable_normal(e0, e1, e2);
  const float dn = rq_dot(Ng, D);
  const float den = rq_add(dn, dn);
  const float tn = rq_dot(v0, Ng);
  const float T = rq_add(tn, tn);
  const float t = rq_div(T, den);
  if (!((tnear <= t) & (t <= tfar) & (den != 0.0f))) return false;
  const bool tiny = fabsf(UVW) < 1e-18f;
  hit.t = t;
  hit.u = tiny ? 0.0f : rq_div(U, UVW);
  hit.v = tiny ? 0.0f : rq_div(V, UVW);
  hit.Ng = Ng;
  return true;
}
