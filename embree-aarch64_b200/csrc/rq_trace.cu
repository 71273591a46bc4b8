// Ray-stream traversal kernels for sm_100a: closest-hit (rtcIntersect1M) and any-hit (rtcOccluded1M).
//
// Replaces the reference's CPU traversal stack for triangle scenes
//   RayStreamFilter::filterAOS          kernels/bvh/bvh_intersector_stream_filters.cpp:11-152
//   BVHNIntersectorKHybrid<8,K,...>     kernels/bvh/bvh_intersector_hybrid.cpp:36-374
//   BVHNIntersector1<8,...>             kernels/bvh/bvh_intersector1.cpp:30-202
//   intersectNode<8,8> (slab test)      kernels/bvh/node_intersector1.h:527-578, robust :621-636
//   Intersect1EpilogM / Occluded1EpilogM kernels/geometry/intersector_epilog.h:215-294,378-444
//   RayStreamAOS get/setHitByOffset     kernels/common/ray.h:1098-1185
// with one design for the GPU: one ray per thread, a persistent while-while loop over 128-byte
// quantised 8-wide nodes (three 32-byte ld.global.nc per node, 32 + 16 bytes per triangle), child
// ordering by ray octant instead of a distance sort, a node-group stack (one 8-byte entry per
// visited level, shared memory first) and the reference's FP32 triangle tests (rq_math.cuh).
// Variants: INST (single-level instancing), LIST (compact hit output for host-staged streams).
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
  const char* nodes;         // RQNode array, 128 bytes per node (traversal reads bytes 0..95 as 3 x 32 B)
  const char* tris;          // RQTri array, 48 bytes per triangle (32 + 16 B, order by index parity: rq_types.h)
  char* rays;                // records are read here ...
  char* out;                 // ... and hit fields written here (same layout; == rays unless the stream is staged, see rtcore_api.cpp)
  size_t stride;
  uint32_t numRays;
  uint32_t instID0;
  uint32_t streamSemantics;
  uint32_t split;            // 1 = one triangle per loop iteration (T/N split), 0 = whole leaf list at once
  uint32_t tVote;            // split only: 0 = T then N every iteration; K >= 1 = one phase per iteration, T when >= K lanes wait for it
  uint32_t sdepth;           // stack entries per thread kept in shared memory (the rest spills to local memory)
  uint32_t refillBelow;      // idle lanes fetch new rays when fewer than this many lanes are traversing
  unsigned int* workCounter; // global ray cursor, zero at launch
  RQTraceCounters* counters;
  const RQInstance* instances; // INST kernels only: table indexed by the instance records of the top-level BVH
  char* hitList;               // LIST kernels only: compact output, one record per ray that hit (48 B closest, 4 B occluded) ...
  unsigned int* hitCount;      // ... appended through this counter (zero at launch); nothing is written to `out`
  uint32_t packed;             // LIST kernels only: rays are dense 32-byte records {org.xyz, tnear, dir.xyz, tfar}
  const float4* verts;         // COMPACT kernels only: vertex pool of the image ...
  const uint32_t* meta;        // ... and the per-triangle geomID | quad-flag words
};

__device__ __forceinline__ float rcpSafe(float d) {           // common/math/vec3fa.h:172-177
  const float a = fabsf(d) < 1e-18f ? copysignf(1e-18f, d) : d;
  return __frcp_rn(a);                                          // = 1.0f / a (both correctly rounded), shorter instruction sequence
}

// Byte j of w as the float 1 + q*2^-16 (bits 0x3F800000 + q*128) with ONE integer dot product:
// IDP.4A issues on the FMA-heavy pipe.  History (ncu, profiles/): I2F.U8 kept the quarter-rate XU
// pipe 54-71 % busy; PRMT (bits 0x3F80qq00) moved the work to the half-rate ALU pipe, which then
// became the limiter (alu 65-74 % of peak, fma 30 %, stall math_pipe_throttle); the dot product
// moves it to the pipe that has room.  The bias is folded into the FMA of the slab test.
__device__ __forceinline__ float byteToUnit(uint32_t w, uint32_t j) {
  return __uint_as_float(__dp4a(w, 0x80u << (8u * j), 0x3F800000u));
}

// 256-bit read-only global load (LDG.E.256.CONSTANT, new on sm_100).  The L1 data pipe spends one
// wavefront per 128-byte line a load instruction touches, whatever its width; with divergent rays
// every lane touches its own line, so an 80-byte node read as 5 x 16 B cost 5 wavefronts per lane
// and made l1tex__data_pipe_lsu_wavefronts the limiter (82 % of peak, profiles/r01b_ncu_trace.txt).
// Three 32-byte loads per node and 32 + 16 bytes per triangle cut that by 40 %.
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
#ifdef RQ_LD128                                                  // A/B switch: two 16-byte loads instead
  const uint4 a = __ldg((const uint4*)p), b = __ldg((const uint4*)p + 1);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  return;
#endif
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

// far-plane inflation (1 + 2^-22): absorbs the rounding of the slab arithmetic so a box is never
// culled while the ray still touches it (role of the reference's robust slab test, 1 -/+ 3 ulp)
#define RQ_FAR_INFLATE 1.00000023841857910156f

// Persistent kernel: every warp keeps pulling rays from a global counter.  A lane that finishes its
// ray goes idle; when fewer than P.refillBelow lanes of the warp are still traversing, the idle
// lanes fetch new rays (one atomicAdd per warp per refill), so SIMD lanes stay busy although ray
// lifetimes differ by an order of magnitude (miss after 4 nodes vs hit after 40).
//
// INST = the scene contains RTC_GEOMETRY_TYPE_INSTANCE geometries (single level; reference:
// kernels/geometry/instance_intersector.cpp:48-105).  A triangle record whose pad word has bit 31
// set is an instance: the lane saves its pending top-level state on the traversal stack, maps the
// ray into the instance's space (org' = xfmPoint(world2local, org), dir' = xfmVector(world2local,
// dir); tnear / tfar carry over unchanged because dir is not renormalised), traverses the
// instanced scene's BVH through per-lane node / triangle base pointers and, once its part of the
// stack is empty again, restores the world-space ray.  Hits report the instanced scene's
// geomID / primID, Ng in instance space and instID[0] = geomID of the instance.
//
// LIST = compact output for host-staged streams: instead of scattering hit fields into the ray
// records, a ray that hit appends one record {rid, tfar, -, - | Ng.xyz, u | v, primID, geomID, instID}
// (closest) or its index (occluded) to a list; the host downloads only that list and scatters it
// into the caller's buffer, so the PCIe link carries ~13 instead of 80 bytes per ray outbound.
// COMPACT = the image uses indexed 16-byte triangle records (RTC_SCENE_FLAG_COMPACT; reference: Triangle4i, trianglei.h): the
// three vertices are fetched through the indices (one more dependent load per test), geomID and the quad flag only for the final hit.
template <bool OCCLUDED, bool ROBUST, bool COUNT, bool ALIGNED, bool SPLIT, int SPILL, bool INST = false, bool LIST = false, bool COMPACT = false>
#ifndef RQ_MIN_CTAS
#define RQ_MIN_CTAS 8   /* 64 registers: 8 CTAs = 32 warps per SM; (128,1) let ptxas take 95 registers and cost 20 % (profiles/r01k_ab.log) */
#endif
// The split closest-hit loop (incoherent streams) is bound by load latency at 32 warps per SM (issue slots 71 % busy, long-scoreboard
// stalls first): 9 CTAs per SM = 56 registers, still without spills, +4.0 % on both scenes; 10 CTAs (48 registers) spill and lose
// 10 %; the whole-list loops (occlusion, coherent) lose 1-2 % at 9 and stay at 8 (same-box A/B: profiles/r02l_ab_ctas.log).
#ifndef RQ_MIN_CTAS_SPLIT
#define RQ_MIN_CTAS_SPLIT 9
#endif
__global__ void __launch_bounds__(128, INST ? 5 : ((SPLIT && !OCCLUDED) ? RQ_MIN_CTAS_SPLIT : RQ_MIN_CTAS))
k_trace(const TraceParams P) {
  // Traversal stack: one 8-byte node-group entry per tree level.  The first P.sdepth levels live
  // in shared memory, entry-major ([level][thread]) so that lanes with different stack depths still
  // hit 32 different banks: a push/pop is 2 wavefronts per warp.  In local memory the same access
  // touched one line per lane and the stack made up 27 % of all L1 sectors
  // (profiles/r01c: 192 M local of 717 M sectors).  Levels beyond P.sdepth spill to local memory.
  extern __shared__ uint2 s_stack[];
  constexpr bool FOLD = !(SPLIT && !OCCLUDED) && !INST;       // see the slab test
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;

  // Per-node bookkeeping by table instead of per-child shifts (LDS runs on the idle LSU pipe):
  //   s_perm[o][m]: slot mask m -> priority mask, slot k moves to bit k ^ o (o = 7 - ray octant)
  //   s_exp3[m]   : slot mask m -> triangle mask, slot k covers bits 3k..3k+2
  __shared__ uint8_t s_perm[8 * 256];
  __shared__ uint32_t s_exp3[256];
  for (unsigned i = threadIdx.x; i < 8u * 256u; i += blockDim.x) {
    const unsigned o = i >> 8, m = i & 255u;
    unsigned r = 0;
    for (unsigned k = 0; k < 8; k++) if (m & (1u << k)) r |= 1u << (k ^ o);
    s_perm[i] = (uint8_t)r;
  }
  for (unsigned m = threadIdx.x; m < 256u; m += blockDim.x) {
    unsigned r = 0;
    for (unsigned k = 0; k < 8; k++) if (m & (1u << k)) r |= 7u << (3u * k);
    s_exp3[m] = r;
  }
  __syncthreads();

  // ---- per-lane ray state ----
  bool active = false;
  uint32_t rid = 0;
  float ox = 0.f, oy = 0.f, oz = 0.f, tnear = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, tfar = 0.f;
  float idx_ = 0.f, idy_ = 0.f, idz_ = 0.f, tnearBox = 0.f;
  uint32_t octinv = 0;
  uint2 spill[SPILL ? SPILL : 1];
  const uint32_t sdepth = P.sdepth;
  int sp = 0;
  uint2 ng = make_uint2(0u, 0u);
  uint32_t tmask = 0u, triBase = 0u, tvalid = 0u;               // pending leaf triangles of the current node
  bool found = false, pend = false;                              // pend: the lane's finished ray has a result to write
  float hu = 0.f, hv = 0.f; RQVec3 hNg = rq_v3(0.f, 0.f, 0.f); uint32_t hPrim = 0, hGeom = 0;
  unsigned long long cntRays = 0, cntNodes = 0, cntTris = 0, cntHits = 0, cntEmpty = 0, cntHitNodes = 0, cntLate = 0; unsigned cntStack = 0, rayNodes = 0;
  bool exhausted = false;                                       // warp-uniform: the global counter ran past numRays
  // INST only: BVH arrays the lane currently traverses, the saved world-space ray, the instance it is in
  const char* cnodes = P.nodes; const char* ctris = P.tris;
  float wox = 0.f, woy = 0.f, woz = 0.f, wdx = 0.f, wdy = 0.f, wdz = 0.f;
  uint32_t curInst = RQ_INVALID, hInst = P.instID0; int spBase = 0;

  for (;;) {
    // ================= results of the rays that finished since the last refill =================
    // Written here and not where the ray ends: there, one or two lanes of the warp would execute the stores (7.5 % of the
    // closest-hit kernel's warp instructions ran at <= 4 active lanes, profiles/r01final_ncu_source.txt); here all lanes that
    // went idle in the meantime do it together.  An idle lane keeps its ray id and hit registers until it is refilled.
    {
      const unsigned fm = __ballot_sync(FULL, pend);
      if (pend) {
        pend = false;
        if (COMPACT && !OCCLUDED) {                             // the hit's triangle index -> geomID, quad half (quad_intersector_moeller.h:28-37)
          const uint32_t m = __ldg(P.meta + hGeom);
          if (m & RQ_META_FLIPUV) { hu = 1.0f - fminf(hu, 1.0f); hv = 1.0f - fminf(hv, 1.0f); }
          hGeom = m & ~RQ_META_FLIPUV;
        }
        if (LIST) {
          // warp-aggregated append: the lanes flushing now share one atomic
          const int fl = __ffs(fm) - 1;
          unsigned fbase = 0;
          if ((int)lane == fl) fbase = atomicAdd(P.hitCount, (unsigned)__popc(fm));
          fbase = __shfl_sync(fm, fbase, fl);
          const unsigned fslot = fbase + (unsigned)__popc(fm & ((1u << lane) - 1u));
          if (OCCLUDED) {
            ((uint32_t*)P.hitList)[fslot] = rid;
          } else {
            float4* d = (float4*)(P.hitList + (size_t)fslot * 48);
            d[0] = make_float4(__uint_as_float(rid), tfar, 0.f, 0.f);
            d[1] = make_float4(hNg.x, hNg.y, hNg.z, hu);
            d[2] = make_float4(hv, __uint_as_float(hPrim), __uint_as_float(hGeom), __uint_as_float(INST ? hInst : P.instID0));
          }
        } else {
          char* rp = P.out + (size_t)rid * P.stride;
          if (COUNT) { cntHits++; cntHitNodes += rayNodes; }
          if (OCCLUDED) {
            *(float*)(rp + 32) = -INFINITY;
          } else {
            *(float*)(rp + 32) = tfar;
            if (ALIGNED) {
              *(float4*)(rp + 48) = make_float4(hNg.x, hNg.y, hNg.z, hu);
              *(float4*)(rp + 64) = make_float4(hv, __uint_as_float(hPrim), __uint_as_float(hGeom), __uint_as_float(INST ? hInst : P.instID0));
            } else {
              float* f = (float*)(rp + 48);
              f[0] = hNg.x; f[1] = hNg.y; f[2] = hNg.z; f[3] = hu; f[4] = hv;
              ((uint32_t*)f)[5] = hPrim; ((uint32_t*)f)[6] = hGeom; ((uint32_t*)f)[7] = INST ? hInst : P.instID0;
            }
          }
        }
      }
    }
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
            if (LIST && P.packed) {                             // host-packed upload: 32 bytes per ray
              const float4 a = __ldg((const float4*)rp), b = __ldg((const float4*)(rp + 16));
              ox = a.x; oy = a.y; oz = a.z; tnear = a.w; dx = b.x; dy = b.y; dz = b.z; tfar = b.w;
            } else if (ALIGNED) {
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
              if (INST) { cnodes = P.nodes; ctris = P.tris; curInst = RQ_INVALID; hInst = P.instID0; }
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
      const unsigned tb = __ballot_sync(FULL, hasTri);
      // Phase vote (SPLIT with P.tVote = K >= 1): an iteration runs ONE phase.  Lanes that reach
      // triangles wait until K lanes do (or nobody has a node to visit), so the expensive N phase
      // (~230 instructions) no longer runs with the T lanes switched off and the T phase
      // (~120 instructions) runs at least K lanes wide.
      bool runN = true;
      if (SPLIT && P.tVote) {
        const unsigned nb = __ballot_sync(FULL, active && tmask == 0u);
        runN = !(tb != 0u && ((unsigned)__popc(tb) >= P.tVote || nb == 0u));
      }
      if (tb != 0u && !(SPLIT && P.tVote && runN)) {
        // SPLIT: one triangle per iteration (incoherent streams); otherwise the whole list now
        // (coherent streams: neighbouring lanes have lists of similar length)
        while (SPLIT ? hasTri : (active && tmask != 0u)) {
          const uint32_t b = 31u - (uint32_t)__clz((int)tmask);
          tmask &= ~(1u << b);
          const uint32_t ti = triBase + __popc(tvalid & ((1u << b) - 1u));
          uint32_t tw[8]; float4 t2;
          if (COMPACT) {
            const uint4 rec = __ldg((const uint4*)(P.tris + (size_t)ti * 16));
            const float4 p0 = __ldg(P.verts + rec.x), p1 = __ldg(P.verts + rec.y), p2 = __ldg(P.verts + rec.z);
            tw[0] = __float_as_uint(p0.x); tw[1] = __float_as_uint(p0.y); tw[2] = __float_as_uint(p0.z);
            tw[3] = __float_as_uint(p1.x); tw[4] = __float_as_uint(p1.y); tw[5] = __float_as_uint(p1.z);
            tw[6] = __float_as_uint(p2.x); tw[7] = __float_as_uint(p2.y);
            t2 = make_float4(p2.z, __uint_as_float(rec.w), __uint_as_float(ti), 0.f);   // "geomID" = the triangle's index until the ray is finished
          } else {
            const char* tp = (INST ? ctris : P.tris) + (size_t)ti * 48;
            const uint32_t odd = ti & 1u;                       // odd records store their last 16 bytes first (32-byte alignment of the wide load)
            ldg256(tp + (odd ? 16 : 0), tw);
            t2 = __ldg((const float4*)(tp + (odd ? 0 : 32)));
          }
          if (COUNT) cntTris++;
          if (INST && (__float_as_uint(t2.w) & 0x80000000u)) {
            // ---- instance record: enter the instanced scene ----
            const uint4* ip = (const uint4*)(P.instances + (__float_as_uint(t2.w) & 0x7FFFFFFFu));
            const uint4 i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2), i3 = __ldg(ip + 3), i4 = __ldg(ip + 4);
            // pending top-level state -> three stack entries (capacity: the host adds them to the depth bound)
            const uint2 sv[3] = {ng, make_uint2(tmask, triBase), make_uint2(tvalid, 0u)};
            #pragma unroll
            for (int e = 0; e < 3; e++) {
              if ((uint32_t)sp < sdepth) s_stack[(uint32_t)sp * 128u + threadIdx.x] = sv[e];
              else if (SPILL && sp < (int)sdepth + SPILL) spill[(uint32_t)sp - sdepth] = sv[e];
              sp++;
            }
            spBase = sp;
            wox = ox; woy = oy; woz = oz; wdx = dx; wdy = dy; wdz = dz;
            // world2local, column major: vx = (i0.x,i0.y,i0.z) vy = (i0.w,i1.x,i1.y) vz = (i1.z,i1.w,i2.x) p = (i2.y,i2.z,i2.w)
            const float vxx = __uint_as_float(i0.x), vxy = __uint_as_float(i0.y), vxz = __uint_as_float(i0.z);
            const float vyx = __uint_as_float(i0.w), vyy = __uint_as_float(i1.x), vyz = __uint_as_float(i1.y);
            const float vzx = __uint_as_float(i1.z), vzy = __uint_as_float(i1.w), vzz = __uint_as_float(i2.x);
            const float px = __uint_as_float(i2.y), py = __uint_as_float(i2.z), pz = __uint_as_float(i2.w);
            // xfmPoint / xfmVector with the reference's FMA nesting (common/math/affinespace.h:102, linearspace3.h:156)
            ox = __fmaf_rn(wox, vxx, __fmaf_rn(woy, vyx, __fmaf_rn(woz, vzx, px)));
            oy = __fmaf_rn(wox, vxy, __fmaf_rn(woy, vyy, __fmaf_rn(woz, vzy, py)));
            oz = __fmaf_rn(wox, vxz, __fmaf_rn(woy, vyz, __fmaf_rn(woz, vzz, pz)));
            dx = __fmaf_rn(wdx, vxx, __fmaf_rn(wdy, vyx, __fmul_rn(wdz, vzx)));
            dy = __fmaf_rn(wdx, vxy, __fmaf_rn(wdy, vyy, __fmul_rn(wdz, vzy)));
            dz = __fmaf_rn(wdx, vxz, __fmaf_rn(wdy, vyz, __fmul_rn(wdz, vzz)));
            idx_ = rcpSafe(dx); idy_ = rcpSafe(dy); idz_ = rcpSafe(dz);
            octinv = 7u - ((dx < 0.f ? 1u : 0u) | (dy < 0.f ? 2u : 0u) | (dz < 0.f ? 4u : 0u));
            cnodes = (const char*)(((unsigned long long)i3.y << 32) | i3.x);
            ctris = (const char*)(((unsigned long long)i3.w << 32) | i3.z);
            curInst = i4.x;
            ng = make_uint2(0u, 0x80000000u); tmask = 0u;           // virtual parent of the instanced root
            break;
          }
          const RQVec3 v0 = rq_v3(__uint_as_float(tw[0]), __uint_as_float(tw[1]), __uint_as_float(tw[2]));
          const RQVec3 v1 = rq_v3(__uint_as_float(tw[3]), __uint_as_float(tw[4]), __uint_as_float(tw[5]));
          const RQVec3 v2 = rq_v3(__uint_as_float(tw[6]), __uint_as_float(tw[7]), t2.x);
          RQTriHit h;
          const bool ok = ROBUST ? rq_pluecker(O, D, tnear, tfar, v0, v1, v2, h)
                                 : rq_moeller(O, D, tnear, tfar, v0, v1, v2, h);
          if (ok) {
            found = true;
            if (OCCLUDED) { tmask = 0u; ng.y = 0u; sp = 0; if (INST) curInst = RQ_INVALID; }    // any hit ends the ray (finishes in the N phase)
            else {
              tfar = h.t; hu = h.u; hv = h.v; hNg = h.Ng;
              if (__float_as_uint(t2.w) & RQ_PAD_FLIPUV) {       // second triangle of a quad (quad_intersector_moeller.h:28-37)
                hu = 1.0f - fminf(hu, 1.0f); hv = 1.0f - fminf(hv, 1.0f);
              }
              hPrim = __float_as_uint(t2.y); hGeom = __float_as_uint(t2.z);
              if (INST) hInst = (curInst != RQ_INVALID) ? curInst : P.instID0;
            }
          }
          if (SPLIT) break;
        }
      }
      // ---------------- N phase ----------------
      if (runN && active && tmask == 0u) {
        bool leftInstance = false;
        if (INST && !(ng.y & 0xFF000000u) && curInst != RQ_INVALID && sp == spBase) {
          // ---- the instanced scene is done: back to world space and to the saved top-level state ----
          uint2 sv[3];
          #pragma unroll
          for (int e = 2; e >= 0; e--) {
            --sp;
            if ((uint32_t)sp < sdepth) sv[e] = s_stack[(uint32_t)sp * 128u + threadIdx.x];
            else if (SPILL) sv[e] = spill[(uint32_t)sp - sdepth];
            else sv[e] = make_uint2(0u, 0u);
          }
          ng = sv[0]; tmask = sv[1].x; triBase = sv[1].y; tvalid = sv[2].x;
          ox = wox; oy = woy; oz = woz; dx = wdx; dy = wdy; dz = wdz;
          idx_ = rcpSafe(dx); idy_ = rcpSafe(dy); idz_ = rcpSafe(dz);
          octinv = 7u - ((dx < 0.f ? 1u : 0u) | (dy < 0.f ? 2u : 0u) | (dz < 0.f ? 4u : 0u));
          cnodes = P.nodes; ctris = P.tris; curInst = RQ_INVALID;
          leftInstance = true;                                  // pending triangles / nodes of the top level continue next iteration
        } else
        if (!(ng.y & 0xFF000000u)) {                            // node group exhausted: pop, or the ray is finished
          if (sp > (int)sdepth + SPILL) sp = (int)sdepth + SPILL;  // entries beyond the stack were dropped (cannot happen: capacity >= depth)
          if (sp == 0) {
            active = false;
            pend = found;                                       // the result is written at the next refill point, together with the other lanes that finished
          } else {
            --sp;
            if ((uint32_t)sp < sdepth) ng = s_stack[(uint32_t)sp * 128u + threadIdx.x];
            else if (SPILL) ng = spill[(uint32_t)sp - sdepth];
          }
        }
        if (active && !(INST && leftInstance)) {
          // ---- descend: take the nearest pending inner child (highest bit) ----
          const uint32_t bit = 31u - (uint32_t)__clz((int)ng.y);
          ng.y &= ~(1u << bit);
          if (ng.y & 0xFF000000u) {
            if ((uint32_t)sp < sdepth) s_stack[(uint32_t)sp * 128u + threadIdx.x] = ng;
            else if (SPILL && sp < (int)sdepth + SPILL) spill[(uint32_t)sp - sdepth] = ng;
            sp++;
            if (COUNT) cntStack = max(cntStack, (unsigned)sp);
          }
          const uint32_t slot = (bit - 24u) ^ octinv;
          const uint32_t rel = __popc(ng.y & 0xFFu & ((1u << slot) - 1u));
          const char* np = (INST ? cnodes : P.nodes) + (size_t)(ng.x + rel) * 128;
          uint32_t na[8], nb[8], nc[8];
          ldg256(np, na); ldg256(np + 32, nb); ldg256(np + 64, nc);
          const uint4 n0 = make_uint4(na[0], na[1], na[2], na[3]), n1 = make_uint4(na[4], na[5], na[6], na[7]);
          const uint4 n2 = make_uint4(nb[0], nb[1], nb[2], nb[3]), n3 = make_uint4(nb[4], nb[5], nb[6], nb[7]);
          const uint4 n4 = make_uint4(nc[0], nc[1], nc[2], nc[3]);
          if (COUNT) {
            cntNodes++; rayNodes++;
            // how many fetches would a distance test at pop time have saved?  (exact bounds from the cold part)
            const float* cb = (const float*)(np + 80);
            const float t0x = ((dx < 0.f ? cb[3] : cb[0]) - ox) * idx_, t0y = ((dy < 0.f ? cb[4] : cb[1]) - oy) * idy_,
                        t0z = ((dz < 0.f ? cb[5] : cb[2]) - oz) * idz_;
            if (fmaxf(fmaxf(t0x, t0y), t0z) > tfar) cntLate++;
          }

          // Slab test of the 8 quantised child boxes.  Plane q of an axis lies at t = q*a + b with
          // a = 2^e * idir (t per grid step) and b = (p - org) * idir.  The byte q becomes the float
          // v = 1 + q*2^-16 with one IDP.4A, so t = fma(v, A, B) with A = a*2^16, B = b - A: two
          // FMA-pipe instructions per plane.  B is rounded once per node; the margin eps >= that
          // rounding keeps the test conservative (near planes earlier, far planes later), and far
          // planes are additionally inflated by 1 + 2^-22.
          const float tfarBox = fmaxf(tfar, 0.0f);
          const float Ax = __uint_as_float((n0.w & 0xFFu) << 23) * idx_ * 65536.0f;
          const float Ay = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * idy_ * 65536.0f;
          const float Az = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * idz_ * 65536.0f;
          const float bx = (__uint_as_float(n0.x) - ox) * idx_;
          const float by = (__uint_as_float(n0.y) - oy) * idy_;
          const float bz = (__uint_as_float(n0.z) - oz) * idz_;
          const float ex = (fabsf(Ax) + fabsf(bx)) * 2.384185791015625e-07f;
          const float ey = (fabsf(Ay) + fabsf(by)) * 2.384185791015625e-07f;
          const float ez = (fabsf(Az) + fabsf(bz)) * 2.384185791015625e-07f;
          const float Bx = (bx - Ax) - ex, By = (by - Ay) - ey, Bz = (bz - Az) - ez;
          const float Axf = Ax * RQ_FAR_INFLATE, Ayf = Ay * RQ_FAR_INFLATE, Azf = Az * RQ_FAR_INFLATE;
          const float Bxf = FOLD ? fmaf(bx - Ax, RQ_FAR_INFLATE, ex) : (bx - Ax) * RQ_FAR_INFLATE + ex;
          const float Byf = FOLD ? fmaf(by - Ay, RQ_FAR_INFLATE, ey) : (by - Ay) * RQ_FAR_INFLATE + ey;
          const float Bzf = FOLD ? fmaf(bz - Az, RQ_FAR_INFLATE, ez) : (bz - Az) * RQ_FAR_INFLATE + ez;
          // a grid step too large for the 2^16 pre-scale (absurd extents x axis-parallel ray): enter every child
          const bool overflow = !(fabsf(Ax) < 1e37f && fabsf(Ay) < 1e37f && fabsf(Az) < 1e37f);
          // near/far quantised planes per axis by ray direction sign (two words = 8 slots each)
          const uint32_t qlx0 = n2.x, qlx1 = n2.y, qly0 = n2.z, qly1 = n2.w, qlz0 = n3.x, qlz1 = n3.y;
          const uint32_t qhx0 = n3.z, qhx1 = n3.w, qhy0 = n4.x, qhy1 = n4.y, qhz0 = n4.z, qhz1 = n4.w;
          const bool nx = dx < 0.f, ny = dy < 0.f, nz = dz < 0.f;
          const uint32_t nearX[2] = {nx ? qhx0 : qlx0, nx ? qhx1 : qlx1}, farX[2] = {nx ? qlx0 : qhx0, nx ? qlx1 : qhx1};
          const uint32_t nearY[2] = {ny ? qhy0 : qly0, ny ? qhy1 : qly1}, farY[2] = {ny ? qly0 : qhy0, ny ? qly1 : qhy1};
          const uint32_t nearZ[2] = {nz ? qhz0 : qlz0, nz ? qhz1 : qlz1}, farZ[2] = {nz ? qlz0 : qhz0, nz ? qlz1 : qhz1};

          // FOLD (the 64-register loops: occlusion / coherent): m and M are clamped with the ray's own interval and one difference
          // decides -- same-box A/B +1.0 ... +1.5 % occlusion rate, answers identical; in the 56-register split closest-hit kernel
          // the same change spills and costs 7 % (profiles/r02p_ab_fold.log), so that kernel keeps the three differences.
          // Child k is hit iff  m <= M, m <= tfar, M >= tnear  with m = max of its three near
          // distances, M = min of its three far distances.  The three differences are formed on
          // the FMA pipe; the OR of their sign bits is the miss flag, shifted into `miss` by one
          // funnel shift: 2 FMNMX3 + LOP3 + SHF on the ALU pipe per child (was 15).
          uint32_t miss = 0;
          #pragma unroll
          for (int k = 7; k >= 0; k--) {
            const int h = k >> 2, j = k & 3;
            const float tminx = fmaf(byteToUnit(nearX[h], j), Ax, Bx);
            const float tminy = fmaf(byteToUnit(nearY[h], j), Ay, By);
            const float tminz = fmaf(byteToUnit(nearZ[h], j), Az, Bz);
            const float tmaxx = fmaf(byteToUnit(farX[h], j), Axf, Bxf);
            const float tmaxy = fmaf(byteToUnit(farY[h], j), Ayf, Byf);
            const float tmaxz = fmaf(byteToUnit(farZ[h], j), Azf, Bzf);
            uint32_t sgn;
            if (FOLD) {                                         // the ray's own interval joins the min / max: 6 instead of 7 ALU instructions per child
              const float m = fmaxf(fmaxf(fmaxf(tminx, tminy), tminz), tnearBox);
              const float M = fminf(fminf(fminf(tmaxx, tmaxy), tmaxz), tfarBox);
              sgn = __float_as_uint(M - m);
            } else {
              const float m = fmaxf(fmaxf(tminx, tminy), tminz);
              const float M = fminf(fminf(tmaxx, tmaxy), tmaxz);
              sgn = __float_as_uint(M - m) | __float_as_uint(tfarBox - m) | __float_as_uint(M - tnearBox);
            }
            miss = __funnelshift_l(sgn, miss, 1);
          }
          const uint32_t masks = n1.z;
          const uint32_t hit8 = overflow ? 0xFFu : (~miss & 0xFFu);
          const uint32_t imask = masks >> 24;
          const uint32_t prio = s_perm[octinv * 256u + (hit8 & imask)];
          if (COUNT && (hit8 & imask) == 0u && (s_exp3[hit8] & masks & 0x00FFFFFFu) == 0u) cntEmpty++;
          ng = make_uint2(n1.x, (prio << 24) | imask);
          tvalid = masks & 0x00FFFFFFu;
          tmask = s_exp3[hit8] & tvalid;
          triBase = n1.y;
#ifdef RQ_PREFETCH
          // Experiment (off): the record this lane will need in its next iteration is known now -- ask for it.  The first use of a
          // node / triangle record is where 10.7 % / 10.4 % of the stall samples sit (profiles/r01final_ncu_source.txt), but the
          // kernel is issue-bound: the extra address arithmetic + CCTL.E.PF1 cost 3-6 % closest-hit and 4-29 % occlusion
          // throughput on configs[1] / configs[2] (profiles/r01pf_sweep_prefetch.log).
          if (tmask == 0u) {
            if (ng.y & 0xFF000000u) {
              const uint32_t nbit = 31u - (uint32_t)__clz((int)ng.y);
              const uint32_t nslot = (nbit - 24u) ^ octinv;
              const uint32_t nrel = __popc(ng.y & 0xFFu & ((1u << nslot) - 1u));
              asm volatile("prefetch.global.L1 [%0];" :: "l"((INST ? cnodes : P.nodes) + (size_t)(ng.x + nrel) * 128));
            }
          }
#if RQ_PREFETCH >= 2
          else {
            const uint32_t pb = 31u - (uint32_t)__clz((int)tmask);
            const uint32_t pti = triBase + __popc(tvalid & ((1u << pb) - 1u));
            asm volatile("prefetch.global.L1 [%0];" :: "l"((INST ? ctris : P.tris) + (size_t)pti * 48));
          }
#endif
#endif
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
    atomicAdd(&P.counters->lateNodes, cntLate);
    atomicMax(&P.counters->stackMax, (unsigned long long)cntStack);
  }
}

cudaError_t launchOne(void (*kern)(const TraceParams), const TraceParams& P, cudaStream_t s) {
  // persistent grid: as many CTAs as fit on the device at once (multiple of the SM count), never
  // more than the stream needs
  static int numSMs = 0;
  if (!numSMs) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev); }
  const size_t smem = (size_t)P.sdepth * 128u * sizeof(uint2);
  int perSM = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, 128, smem);
  if (e != cudaSuccess) return e;
  if (perSM < 1) perSM = 1;
  const unsigned need = (P.numRays + 127u) / 128u;
  const unsigned grid = need < (unsigned)(numSMs * perSM) ? need : (unsigned)(numSMs * perSM);
  kern<<<grid, 128, smem, s>>>(P);
  rqCountLaunch(1);
  return cudaGetLastError();
}

// Stack capacity = tree depth (one node-group entry per level).  P.sdepth levels live in shared
// memory (device option stack_smem, 0..16), the rest in a local-memory array of SPILL entries.
template <bool OCC, bool ROBUST, bool COUNT, bool ALIGNED>
cudaError_t launchStack(TraceParams& P, uint32_t depth, cudaStream_t s) {
  if (P.sdepth > depth) P.sdepth = depth;
  const uint32_t spill = depth - P.sdepth;
  if (P.hitList) {
    // compact-output variants (host-staged streams of flat scenes): shared levels + at most 32 local entries, no counters
    if (spill > 32 || P.instances) return cudaErrorInvalidValue;
    if (spill == 0) {
      if (P.split) return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, true, 0, false, true>, P, s);
      return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, false, 0, false, true>, P, s);
    }
    if (P.split) return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, true, 32, false, true>, P, s);
    return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, false, 32, false, true>, P, s);
  }
  if (P.verts) {
    // compact (indexed) images: flat scenes only, in-place output, no counters; deep trees fall into the widest local stack
    if (P.instances || P.hitList) return cudaErrorInvalidValue;
    if (spill == 0) {
      if (P.split) return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, true, 0, false, false, true>, P, s);
      return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, false, 0, false, false, true>, P, s);
    }
    if (spill <= 32) {
      if (P.split) return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, true, 32, false, false, true>, P, s);
      return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, false, 32, false, false, true>, P, s);
    }
    return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, false, 208, false, false, true>, P, s);
  }
  if (P.instances) {
    // instanced scenes: one stack configuration (shared levels + 32 local entries), no counters
    if (spill > 32) return cudaErrorInvalidValue;
    if (P.split) return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, true, 32, true>, P, s);
    return launchOne(k_trace<OCC, ROBUST, false, ALIGNED, false, 32, true>, P, s);
  }
  if (spill == 0) {
    if (P.split) return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, true, 0>, P, s);
    return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, false, 0>, P, s);
  }
  if (spill <= 32) {
    if (P.split) return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, true, 32>, P, s);
    return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, false, 32>, P, s);
  }
  if (spill <= 96) return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, false, 96>, P, s);
  return launchOne(k_trace<OCC, ROBUST, COUNT, ALIGNED, false, 208>, P, s);
}

}  // namespace

// The host passes the node / triangle offsets from its copy of the image header, so launching
// needs no device read.
static int launchTrace(bool occ, const RQTraceArgs* a, cudaStream_t s) {
  if (a->numRays == 0) return 0;
  TraceParams P;
  P.nodes = (const char*)a->image + a->nodesOffset;
  P.tris = (const char*)a->image + a->trisOffset;
  P.rays = (char*)a->rays; P.out = a->out ? (char*)a->out : (char*)a->rays; P.stride = a->stride; P.numRays = a->numRays; P.instID0 = a->instID0;
  P.streamSemantics = a->streamSemantics; P.counters = a->counters;
  P.workCounter = a->workCounter;
  P.refillBelow = a->refillBelow ? a->refillBelow : 26u;
  P.split = a->split; P.tVote = a->tVote; P.sdepth = a->stackSmem;
  P.instances = (const RQInstance*)a->instances;
  P.verts = a->compact ? (const float4*)((const char*)a->image + a->vertsOffset) : nullptr;
  P.meta = a->compact ? (const uint32_t*)((const char*)a->image + a->metaOffset) : nullptr;
  P.hitList = (char*)a->hitList; P.hitCount = a->hitCount; P.packed = a->hitList ? a->packed : 0u;
  if (P.verts && (P.hitList || P.instances || a->counters)) return (int)cudaErrorInvalidValue;   // the host never asks for these combinations
  if (P.hitList) {
    if (!P.hitCount) return (int)cudaErrorInvalidValue;
    cudaError_t ec = cudaMemsetAsync(P.hitCount, 0, sizeof(unsigned int), s);
    if (ec != cudaSuccess) return (int)ec;
  }
  if (!P.workCounter) return (int)cudaErrorInvalidValue;
  {
    cudaError_t ez = cudaMemsetAsync(P.workCounter, 0, sizeof(unsigned int), s);   // stream ordered with the launch
    if (ez != cudaSuccess) return (int)ez;
  }
  const bool aligned = (((uintptr_t)a->rays | (uintptr_t)P.out | (uintptr_t)a->stride) & 15u) == 0;
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
// layout adapters: SoA packets / pointer-SoA / array-of-pointers  <->  dense AoS RTCRayHit scratch (80-byte records)
// (reference: filterSOA / filterSOP / filterAOP, kernels/bvh/bvh_intersector_stream_filters.cpp:155-592)
// ------------------------------------------------------------------------------------------------
namespace {
template <typename T>
__device__ __forceinline__ T* soaAt(T* field, const RQSoAView& v, uint32_t i) {
  const uint32_t m = i / v.N, j = i - m * v.N;
  return (T*)((char*)field + (size_t)m * v.packetStride) + j;
}
__global__ void __launch_bounds__(256)
k_gather_soa(const RQSoAView v, const int* __restrict__ valid, uint32_t n, float* __restrict__ aos) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* r = aos + (size_t)i * 20;
  const bool ok = valid ? valid[i] != 0 : true;
  r[0] = *soaAt(v.org_x, v, i); r[1] = *soaAt(v.org_y, v, i); r[2] = *soaAt(v.org_z, v, i);
  r[3] = ok ? *soaAt(v.tnear, v, i) : INFINITY;                 // invalid lane = inactive ray
  r[4] = *soaAt(v.dir_x, v, i); r[5] = *soaAt(v.dir_y, v, i); r[6] = *soaAt(v.dir_z, v, i); r[7] = 0.f;
  r[8] = ok ? *soaAt(v.tfar, v, i) : -INFINITY;
  ((uint32_t*)r)[9] = 0; ((uint32_t*)r)[10] = 0; ((uint32_t*)r)[11] = 0;
  ((uint32_t*)r)[18] = RQ_INVALID;                              // geomID marks "hit written"
}
__global__ void __launch_bounds__(256)
k_scatter_soa(const RQSoAView v, uint32_t n, const float* __restrict__ aos, int occluded) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = aos + (size_t)i * 20;
  if (occluded) {
    if (r[8] == -INFINITY && r[3] != INFINITY) *soaAt(v.tfar, v, i) = -INFINITY;   // only newly occluded, valid lanes
    return;
  }
  if (((const uint32_t*)r)[18] == RQ_INVALID) return;           // miss or inactive: nothing is written
  *soaAt(v.tfar, v, i) = r[8];
  *soaAt(v.Ng_x, v, i) = r[12]; *soaAt(v.Ng_y, v, i) = r[13]; *soaAt(v.Ng_z, v, i) = r[14];
  *soaAt(v.u, v, i) = r[15]; *soaAt(v.v, v, i) = r[16];
  *soaAt(v.primID, v, i) = ((const uint32_t*)r)[17]; *soaAt(v.geomID, v, i) = ((const uint32_t*)r)[18];
  if (v.instID0) *soaAt(v.instID0, v, i) = ((const uint32_t*)r)[19];
}
// one thread per record, 4-byte words (the caller's records are only guaranteed 4-byte aligned)
__global__ void __launch_bounds__(256)
k_gather_aop(const void* const* __restrict__ ptrs, uint32_t n, int recBytes, uint32_t* __restrict__ aos) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* s = (const uint32_t*)ptrs[i];
  uint32_t* d = aos + (size_t)i * 20;
  const int words = recBytes / 4;
  for (int k = 0; k < words; k++) d[k] = s[k];
  for (int k = words; k < 20; k++) d[k] = 0u;
  d[18] = RQ_INVALID;                                           // geomID marks "hit written" (the caller's own value is never read back)
}
__global__ void __launch_bounds__(256)
k_scatter_aop(void* const* __restrict__ ptrs, uint32_t n, const uint32_t* __restrict__ aos, int occluded) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* r = aos + (size_t)i * 20;
  uint32_t* d = (uint32_t*)ptrs[i];
  if (occluded) { if (r[8] == 0xFF800000u) d[8] = 0xFF800000u; return; }   // tfar = -inf
  if (r[18] == RQ_INVALID) return;
  d[8] = r[8];
  for (int k = 12; k < 20; k++) d[k] = r[k];
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
int rqGatherAoP(const void* const* ptrs, uint32_t n, int recBytes, void* aos, rqStream stream) {
  if (!n) return 0;
  k_gather_aop<<<(n + 255u) / 256u, 256, 0, (cudaStream_t)stream>>>(ptrs, n, recBytes, (uint32_t*)aos);
  rqCountLaunch(1);
  return (int)cudaGetLastError();
}
int rqScatterAoP(void* const* ptrs, uint32_t n, const void* aos, int occluded, rqStream stream) {
  if (!n) return 0;
  k_scatter_aop<<<(n + 255u) / 256u, 256, 0, (cudaStream_t)stream>>>(ptrs, n, (const uint32_t*)aos, occluded);
  rqCountLaunch(1);
  return (int)cudaGetLastError();
}
