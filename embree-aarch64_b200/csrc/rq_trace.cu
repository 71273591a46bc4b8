// Ray-stream traversal kernels for sm_100a: closest-hit (rtcIntersect1M) and any-hit (rtcOccluded1M).
//
// Replaces the reference's CPU traversal stack for triangle scenes
//   RayStreamFilter::filterAOS          kernels/bvh/bvh_intersector_stream_filters.cpp:11-152
//   BVHNIntersectorKHybrid<8,K,...>     kernels/bvh/bvh_intersector_hybrid.cpp:36-374
//   BVHNIntersector1<8,...>             kernels/bvh/bvh_intersector1.cpp:30-202
//   intersectNode<8,8> (slab test)      kernels/bvh/node_intersector1.h:527-578, robust :621-636
//   Intersect1EpilogM / Occluded1EpilogM kernels/geometry/intersector_epilog.h:215-294,378-444
//   RayStreamAOS get/setHitByOffset     kernels/common/ray.h:1098-1185
// with one design for the GPU: one ray per thread, a while-while loop over 128-byte quantised
// 8-wide nodes (five 16-byte ld.global.nc per node, three per triangle), child ordering by ray
// octant instead of a distance sort, a node-group stack (one 8-byte entry per visited level) and
// the reference's FP32 triangle tests (rq_math.cuh).
//
// Ray semantics preserved: a ray is inactive unless tnear <= tfar (NaN = inactive)
// and is then left untouched; box culling uses max(tnear,0) / max(tfar,0) (bvh_intersector1.cpp:64)
// while the triangle depth test uses the ray's own tnear and current tfar (moeller.h:93-94); a
// miss writes nothing; a closest hit writes tfar, Ng, u, v, primID, geomID, instID[0]; an
// occlusion hit writes tfar = -inf only; mask/time/id/flags are ignored.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include "rq_device.h"
#include "rq_math.cuh"

namespace {

struct TraceParams {
  const uint4* nodes;        // RQNode array viewed as 8 x uint4 per node
  const float4* tris;        // RQTri array viewed as 3 x float4 per triangle
  char* rays;
  size_t stride;
  uint32_t numRays;
  uint32_t instID0;
  uint32_t streamSemantics;
  uint32_t split;            // 1 = one triangle per loop iteration (T/N split), 0 = whole leaf list at once
  uint32_t refillBelow;      // idle lanes fetch new rays when fewer than this many lanes are traversing
  unsigned int* workCounter; // global ray cursor, zero at launch
  RQTraceCounters* counters;
};

__device__ __forceinline__ float rcpSafe(float d) {           // common/math/vec3fa.h:172-177
  const float a = fabsf(d) < 1e-18f ? copysignf(1e-18f, d) : d;
  return 1.0f / a;
}

// byte j of w as float WITHOUT the int->float converter: I2F.U8 issues on the quarter-rate XU pipe and
// 48 of them per node made that pipe the limiter (ncu: xu 54-71 % busy).  PRMT drops the byte into
// the mantissa of 2^23 (0x4B000000 | q == 8388608 + q exactly), one FADD removes the bias.
__device__ __forceinline__ float byteToFloat(uint32_t w, uint32_t j) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | j)) - 8388608.0f;
}
// byte j of w as the float 1 + q*2^-15 (bits 0x3F80qq00): a single PRMT, the bias is folded into the FMA
__device__ __forceinline__ float byteToUnit(uint32_t w, uint32_t j) {
  return __uint_as_float(__byte_perm(w, 0x3F800000u, 0x7604u | (j << 4)));
}

// far-plane inflation (1 + 2^-22): absorbs the rounding of the slab arithmetic so a box is never
// culled while the ray still touches it (role of the reference's robust slab test, 1 -/+ 3 ulp)
#define RQ_FAR_INFLATE 1.00000023841857910156f

// Persistent kernel: every warp keeps pulling rays from a global counter.  A lane that finishes its
// ray goes idle; when fewer than P.refillBelow lanes of the warp are still traversing, the idle
// lanes fetch new rays (one atomicAdd per warp per refill), so SIMD lanes stay busy although ray
// lifetimes differ by an order of magnitude (miss after 4 nodes vs hit after 40).
template <bool OCCLUDED, bool ROBUST, bool COUNT, bool ALIGNED, bool SPLIT, int STACK>
__global__ void __launch_bounds__(128)
k_trace(const TraceParams P) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;

  // ---- per-lane ray state ----
  bool active = false;
  uint32_t rid = 0;
  float ox = 0.f, oy = 0.f, oz = 0.f, tnear = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, tfar = 0.f;
  float idx_ = 0.f, idy_ = 0.f, idz_ = 0.f, tnearBox = 0.f;
  uint32_t octinv = 0;
  uint2 stack[STACK];
  int sp = 0;
  uint2 ng = make_uint2(0u, 0u);
  uint32_t tmask = 0u, triBase = 0u;                            // pending leaf triangles of the current node
  bool found = false;
  float hu = 0.f, hv = 0.f; RQVec3 hNg = rq_v3(0.f, 0.f, 0.f); uint32_t hPrim = 0, hGeom = 0;
  unsigned long long cntRays = 0, cntNodes = 0, cntTris = 0, cntHits = 0, cntEmpty = 0, cntHitNodes = 0; unsigned cntStack = 0, rayNodes = 0;
  bool exhausted = false;                                       // warp-uniform: the global counter ran past numRays

  for (;;) {
    // ================= refill: idle lanes take the next rays of the stream =================
    const unsigned idle = __ballot_sync(FULL, !active);
    if (idle) {
      if (!exhausted) {
        const int n = __popc(idle);
        const int leader = __ffs(idle) - 1;
        unsigned base = 0;
        if ((int)lane == leader) base = atomicAdd(P.workCounter, (unsigned)n);
        base = __shfl_sync(FULL, base, leader);
        if (base + (unsigned)n >= P.numRays) exhausted = true;
        if (!active) {
          rid = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
          if (rid < P.numRays) {
            const char* rp = P.rays + (size_t)rid * P.stride;
            if (ALIGNED) {
              const float4 a = *(const float4*)rp, b = *(const float4*)(rp + 16);
              ox = a.x; oy = a.y; oz = a.z; tnear = a.w; dx = b.x; dy = b.y; dz = b.z;
              tfar = *(const float*)(rp + 32);
            } else {
              const float* f = (const float*)rp;
              ox = f[0]; oy = f[1]; oz = f[2]; tnear = f[3]; dx = f[4]; dy = f[5]; dz = f[6]; tfar = f[8];
            }
            bool ok = tnear <= tfar;                            // entry rules (NaN => inactive, ray untouched)
            if (OCCLUDED) {
              ok = ok && !(tfar < 0.0f);                        // already occluded (stream_filters.cpp:78, intersector1.cpp:132)
              if (P.streamSemantics) ok = ok && (tnear >= 0.0f);  // bvh_intersector_stream.cpp:303-305
            }
            if (ok) {
              active = true; found = false; sp = 0; tmask = 0u;
              idx_ = rcpSafe(dx); idy_ = rcpSafe(dy); idz_ = rcpSafe(dz);
              tnearBox = fmaxf(tnear, 0.0f);
              octinv = 7u - ((dx < 0.f ? 1u : 0u) | (dy < 0.f ? 2u : 0u) | (dz < 0.f ? 4u : 0u));
              ng = make_uint2(0u, 0x80000000u);                 // virtual parent: one inner hit -> node 0
              if (COUNT) { cntRays++; rayNodes = 0; }
            }
          }
        }
      }
      if (__ballot_sync(FULL, active) == 0u) {
        if (exhausted) break;
        continue;
      }
    }

    // ================= traverse until too few lanes are busy =================
    // Every iteration has two warp-wide phases.  T: lanes with pending leaf triangles test ONE
    // triangle.  N: lanes without pending triangles pop / descend ONE node.  A lane never starts a
    // node before its triangles are done (tfar must shrink first), but lanes no longer wait for the
    // lane with the longest triangle list: that list is spread over several iterations while the
    // other lanes keep descending.
    for (;;) {
      const RQVec3 O = rq_v3(ox, oy, oz), D = rq_v3(dx, dy, dz);
      // ---------------- T phase ----------------
      const bool hasTri = active && (tmask != 0u);
      if (__any_sync(FULL, hasTri)) {
        // SPLIT: one triangle per iteration (incoherent streams); otherwise the whole list now
        // (coherent streams: neighbouring lanes have lists of similar length)
        while (SPLIT ? hasTri : (active && tmask != 0u)) {
          const uint32_t b = 31u - (uint32_t)__clz((int)tmask);
          tmask &= ~(1u << b);
          const float4* tp = P.tris + (size_t)(triBase + b) * 3;
          const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
          if (COUNT) cntTris++;
          const RQVec3 v0 = rq_v3(t0.x, t0.y, t0.z), v1 = rq_v3(t0.w, t1.x, t1.y), v2 = rq_v3(t1.z, t1.w, t2.x);
          RQTriHit h;
          const bool ok = ROBUST ? rq_pluecker(O, D, tnear, tfar, v0, v1, v2, h)
                                 : rq_moeller(O, D, tnear, tfar, v0, v1, v2, h);
          if (ok) {
            found = true;
            if (OCCLUDED) { tmask = 0u; ng.y = 0u; sp = 0; }    // any hit ends the ray (finishes in the N phase)
            else {
              tfar = h.t; hu = h.u; hv = h.v; hNg = h.Ng;
              hPrim = __float_as_uint(t2.y); hGeom = __float_as_uint(t2.z);
            }
          }
          if (SPLIT) break;
        }
      }
      // ---------------- N phase ----------------
      if (active && tmask == 0u) {
        if (!(ng.y & 0xFF000000u)) {                            // node group exhausted: pop, or the ray is finished
          if (sp > STACK) sp = STACK;                           // entries beyond the stack were dropped (cannot happen: STACK >= depth)
          if (sp == 0) {
            active = false;
            if (found) {
              char* rp = P.rays + (size_t)rid * P.stride;
              if (COUNT) { cntHits++; cntHitNodes += rayNodes; }
              if (OCCLUDED) {
                *(float*)(rp + 32) = -INFINITY;
              } else {
                *(float*)(rp + 32) = tfar;
                if (ALIGNED) {
                  *(float4*)(rp + 48) = make_float4(hNg.x, hNg.y, hNg.z, hu);
                  *(float4*)(rp + 64) = make_float4(hv, __uint_as_float(hPrim), __uint_as_float(hGeom), __uint_as_float(P.instID0));
                } else {
                  float* f = (float*)(rp + 48);
                  f[0] = hNg.x; f[1] = hNg.y; f[2] = hNg.z; f[3] = hu; f[4] = hv;
                  ((uint32_t*)f)[5] = hPrim; ((uint32_t*)f)[6] = hGeom; ((uint32_t*)f)[7] = P.instID0;
                }
              }
            }
          } else {
            ng = stack[--sp];
          }
        }
        if (active) {
          // ---- descend: take the nearest pending inner child (highest bit) ----
          const uint32_t bit = 31u - (uint32_t)__clz((int)ng.y);
          ng.y &= ~(1u << bit);
          if (ng.y & 0xFF000000u) {
            if (sp < STACK) stack[sp] = ng;
            sp++;
            if (COUNT) cntStack = max(cntStack, (unsigned)sp);
          }
          const uint32_t slot = (bit - 24u) ^ octinv;
          const uint32_t rel = __popc(ng.y & 0xFFu & ((1u << slot) - 1u));
          const uint4* np = P.nodes + (size_t)(ng.x + rel) * 8;
          const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
          if (COUNT) { cntNodes++; rayNodes++; }

          // Slab test of the 8 quantised child boxes.  Plane q of an axis lies at t = q*a + b with
          // a = 2^e * idir (t per grid step) and b = (p - org) * idir.  The byte q is turned into the
          // float v = 1 + q*2^-15 by ONE PRMT (bits 0x3F80qq00), so t = fma(v, A, B) with A = a*2^15,
          // B = b - A: two instructions per plane and no int->float converter.  B is rounded once
          // per node; the margin eps >= that rounding keeps the test conservative (near planes
          // earlier, far planes later), and far planes are additionally inflated by 1 + 2^-22.
          const float tfarBox = fmaxf(tfar, 0.0f);
          const float Ax = __uint_as_float((n0.w & 0xFFu) << 23) * idx_ * 32768.0f;
          const float Ay = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * idy_ * 32768.0f;
          const float Az = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * idz_ * 32768.0f;
          const float bx = (__uint_as_float(n0.x) - ox) * idx_;
          const float by = (__uint_as_float(n0.y) - oy) * idy_;
          const float bz = (__uint_as_float(n0.z) - oz) * idz_;
          const float ex = (fabsf(Ax) + fabsf(bx)) * 2.384185791015625e-07f;
          const float ey = (fabsf(Ay) + fabsf(by)) * 2.384185791015625e-07f;
          const float ez = (fabsf(Az) + fabsf(bz)) * 2.384185791015625e-07f;
          const float Bx = (bx - Ax) - ex, By = (by - Ay) - ey, Bz = (bz - Az) - ez;
          const float Axf = Ax * RQ_FAR_INFLATE, Ayf = Ay * RQ_FAR_INFLATE, Azf = Az * RQ_FAR_INFLATE;
          const float Bxf = (bx - Ax) * RQ_FAR_INFLATE + ex, Byf = (by - Ay) * RQ_FAR_INFLATE + ey, Bzf = (bz - Az) * RQ_FAR_INFLATE + ez;
          // a grid step too large for the 2^15 pre-scale (absurd extents x axis-parallel ray): enter every child
          const bool overflow = !(fabsf(Ax) < 1e37f && fabsf(Ay) < 1e37f && fabsf(Az) < 1e37f);
          // near/far quantised planes per axis by ray direction sign (two words = 8 slots each)
          const uint32_t qlx0 = n2.x, qlx1 = n2.y, qly0 = n2.z, qly1 = n2.w, qlz0 = n3.x, qlz1 = n3.y;
          const uint32_t qhx0 = n3.z, qhx1 = n3.w, qhy0 = n4.x, qhy1 = n4.y, qhz0 = n4.z, qhz1 = n4.w;
          const bool nx = dx < 0.f, ny = dy < 0.f, nz = dz < 0.f;
          const uint32_t nearX[2] = {nx ? qhx0 : qlx0, nx ? qhx1 : qlx1}, farX[2] = {nx ? qlx0 : qhx0, nx ? qlx1 : qhx1};
          const uint32_t nearY[2] = {ny ? qhy0 : qly0, ny ? qhy1 : qly1}, farY[2] = {ny ? qly0 : qhy0, ny ? qly1 : qhy1};
          const uint32_t nearZ[2] = {nz ? qhz0 : qlz0, nz ? qhz1 : qlz1}, farZ[2] = {nz ? qlz0 : qhz0, nz ? qlz1 : qhz1};
          const uint32_t metaW[2] = {n1.z, n1.w};
          const uint32_t octinv4 = octinv * 0x01010101u;

          uint32_t hitmask = 0;
          #pragma unroll
          for (int h = 0; h < 2; h++) {
            // per-byte bookkeeping for 4 children at once: inner children get their bit at
            // 24 + (slot ^ octinv), leaf children at their triangle offset
            const uint32_t meta4 = metaW[h];
            const uint32_t inner4 = ((meta4 & (meta4 << 1)) & 0x10101010u) >> 4;          // 0x01 in bytes with index >= 24
            const uint32_t index4 = (meta4 ^ (octinv4 & (inner4 * 0xFFu))) & 0x1F1F1F1Fu;
            const uint32_t bits4 = (meta4 >> 5) & 0x07070707u;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
              const float tminx = fmaf(byteToUnit(nearX[h], j), Ax, Bx);
              const float tminy = fmaf(byteToUnit(nearY[h], j), Ay, By);
              const float tminz = fmaf(byteToUnit(nearZ[h], j), Az, Bz);
              const float tmaxx = fmaf(byteToUnit(farX[h], j), Axf, Bxf);
              const float tmaxy = fmaf(byteToUnit(farY[h], j), Ayf, Byf);
              const float tmaxz = fmaf(byteToUnit(farZ[h], j), Azf, Bzf);
              const float tmin = fmaxf(fmaxf(tminx, tminy), fmaxf(tminz, tnearBox));
              const float tmax = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, tfarBox));
              const uint32_t contrib = ((bits4 >> (8 * j)) & 0xFFu) << ((index4 >> (8 * j)) & 0xFFu);
              hitmask |= ((tmin <= tmax) | overflow) ? contrib : 0u;
            }
          }
          if (COUNT && hitmask == 0u) cntEmpty++;
          ng = make_uint2(n1.x, (hitmask & 0xFF000000u) | (n0.w >> 24));
          tmask = hitmask & 0x00FFFFFFu;
          triBase = n1.y;
        }
      }
      const unsigned act = __ballot_sync(FULL, active);
      if (act == 0u) break;
      if (!exhausted && (unsigned)__popc(act) < P.refillBelow) break;
    }
  }

  if (COUNT) {
    atomicAdd(&P.counters->rays, cntRays);
    atomicAdd(&P.counters->nodes, cntNodes);
    atomicAdd(&P.counters->tris, cntTris);
    atomicAdd(&P.counters->hits, cntHits);
    atomicAdd(&P.counters->emptyNodes, cntEmpty);
    atomicAdd(&P.counters->hitNodes, cntHitNodes);
    atomicMax(&P.counters->stackMax, (unsigned long long)cntStack);
  }
}

cudaError_t launchOne(void (*kern)(const TraceParams), const TraceParams& P, cudaStream_t s) {
  // persistent grid: as many CTAs as fit on the device at once (multiple of the SM count), never
  // more than the stream needs
  static int numSMs = 0;
  if (!numSMs) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev); }
  int perSM = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, 128, 0);
  if (e != cudaSuccess) return e;
  if (perSM < 1) perSM = 1;
  const unsigned need = (P.numRays + 127u) / 128u;
  const unsigned grid = need < (unsigned)(numSMs * perSM) ? need : (unsigned)(numSMs * perSM);
  kern<<<grid, 128, 0, s>>>(P);
  rqCountLaunch(1);
  return cudaGetLastError();
}

template <bool OCC, bool ROBUST, bool COUNT, bool ALIGNED>
cudaError_t launchStack(const TraceParams& P, uint32_t depth, cudaStream_t s) {
  if (depth <= 32) {
    if (P.split) return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, true, 32>, P, s);
    return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, false, 32>, P, s);
  }
  if (depth <= 96) return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, false, 96>, P, s);
  return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, false, 208>, P, s);
}

}  // namespace

// The host passes the node / triangle offsets from its copy of the image header, so launching
// needs no device read.
static int launchTrace(bool occ, const RQTraceArgs* a, cudaStream_t s) {
  if (a->numRays == 0) return 0;
  TraceParams P;
  P.nodes = (const uint4*)((const char*)a->image + a->nodesOffset);
  P.tris = (const float4*)((const char*)a->image + a->trisOffset);
  P.rays = (char*)a->rays; P.stride = a->stride; P.numRays = a->numRays; P.instID0 = a->instID0;
  P.streamSemantics = a->streamSemantics; P.counters = a->counters;
  P.workCounter = a->workCounter;
  P.refillBelow = a->refillBelow ? a->refillBelow : 26u;
  P.split = a->split;
  if (!P.workCounter) return (int)cudaErrorInvalidValue;
  {
    cudaError_t ez = cudaMemsetAsync(P.workCounter, 0, sizeof(unsigned int), s);   // stream ordered with the launch
    if (ez != cudaSuccess) return (int)ez;
  }
  const bool aligned = (((uintptr_t)a->rays | (uintptr_t)a->stride) & 15u) == 0;
  const bool count = a->counters != nullptr;
  const bool robust = a->robust != 0;
  cudaError_t e;
#define RQ_DISPATCH(OCC, ROB, CNT, ALN) e = launchStack<OCC, ROB, CNT, ALN>(P, a->depth, s)
#define RQ_D3(OCC, ROB, CNT) do { if (aligned) RQ_DISPATCH(OCC, ROB, CNT, true); else RQ_DISPATCH(OCC, ROB, CNT, false); } while (0)
#define RQ_D2(OCC, ROB) do { if (count) RQ_D3(OCC, ROB, true); else RQ_D3(OCC, ROB, false); } while (0)
#define RQ_D1(OCC) do { if (robust) RQ_D2(OCC, true); else RQ_D2(OCC, false); } while (0)
  if (occ) RQ_D1(true); else RQ_D1(false);
  return (int)e;
}

int rqLaunchIntersect(const RQTraceArgs* a, rqStream stream) { return launchTrace(false, a, (cudaStream_t)stream); }
int rqLaunchOccluded(const RQTraceArgs* a, rqStream stream) { return launchTrace(true, a, (cudaStream_t)stream); }

// ------------------------------------------------------------------------------------------------
// layout adapters: SoA / pointer-SoA  <->  dense AoS RTCRayHit scratch (80-byte records)
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
k_gather_soa(const RQSoAView v, const int* __restrict__ valid, uint32_t n, float* __restrict__ aos) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* r = aos + (size_t)i * 20;
  const bool ok = valid ? valid[i] != 0 : true;
  r[0] = v.org_x[i]; r[1] = v.org_y[i]; r[2] = v.org_z[i];
  r[3] = ok ? v.tnear[i] : INFINITY;                            // invalid lane = inactive ray
  r[4] = v.dir_x[i]; r[5] = v.dir_y[i]; r[6] = v.dir_z[i]; r[7] = 0.f;
  r[8] = ok ? v.tfar[i] : -INFINITY;
  ((uint32_t*)r)[9] = 0; ((uint32_t*)r)[10] = 0; ((uint32_t*)r)[11] = 0;
  ((uint32_t*)r)[18] = RQ_INVALID;                              // geomID marks "hit written"
}
__global__ void __launch_bounds__(256)
k_scatter_soa(const RQSoAView v, uint32_t n, const float* __restrict__ aos, int occluded) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = aos + (size_t)i * 20;
  if (occluded) {
    if (r[8] == -INFINITY && r[3] != INFINITY) v.tfar[i] = -INFINITY;   // only newly occluded, valid lanes
    return;
  }
  if (((const uint32_t*)r)[18] == RQ_INVALID) return;           // miss or inactive: nothing is written
  v.tfar[i] = r[8];
  v.Ng_x[i] = r[12]; v.Ng_y[i] = r[13]; v.Ng_z[i] = r[14]; v.u[i] = r[15]; v.v[i] = r[16];
  v.primID[i] = ((const uint32_t*)r)[17]; v.geomID[i] = ((const uint32_t*)r)[18];
  if (v.instID0) v.instID0[i] = ((const uint32_t*)r)[19];
}
}  // namespace

int rqGatherSoA(const RQSoAView* v, const int* valid, uint32_t n, void* aos, rqStream stream) {
  if (!n) return 0;
  k_gather_soa<<<(n + 255u) / 256u, 256, 0, (cudaStream_t)stream>>>(*v, valid, n, (float*)aos);
  rqCountLaunch(1);
  return (int)cudaGetLastError();
}
int rqScatterSoA(const RQSoAView* v, uint32_t n, const void* aos, int occluded, rqStream stream) {
  if (!n) return 0;
  k_scatter_soa<<<(n + 255u) / 256u, 256, 0, (cudaStream_t)stream>>>(*v, n, (const float*)aos, occluded);
  rqCountLaunch(1);
  return (int)cudaGetLastError();
}
