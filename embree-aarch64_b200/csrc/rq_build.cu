// GPU BVH builder for triangle scenes (sm_100a).
//
// Replaces, on the device, the reference's CPU build path
//   createPrimRefArray            kernels/builders/primrefgen.cpp:35-57, scene_triangle_mesh.h:131-153,208-221
//   BVHNBuilderSAH<8,Triangle4>   kernels/bvh/bvh_builder_sah.cpp:85-191
//   GeneralBVHBuilder::recurse    kernels/builders/bvh_builder_sah.h:222-319 (greedy 8-wide widening by SAH)
//   BVHBuilderMorton              kernels/builders/bvh_builder_morton.h:70-104,300-433 (Morton front end)
//   radix_sort_u32                common/algorithms/parallel_sort.h
//   BVHNStatistics                kernels/bvh/bvh_statistics.cpp:41-160 (SAH figure)
// with a pipeline designed for the GPU rather than translated:
//   1. k_setup_prims   one thread per triangle: validity filter (index range, |v| < 1.844e18),
//                      48-byte triangle record, scene + centroid bounds (warp shuffles + atomics)
//   2. k_morton        63-bit Morton code of the box centre (21 bits per axis)
//   3. radix sort      LSD, 8 bits per pass, stable warp-match ranking (hand written, no CUB)
//   4. k_hierarchy     binary radix tree over the sorted codes (one thread per inner node)
//   5. k_refit_dp      bottom-up: exact bounds + SAH dynamic programme that decides, per binary
//                      node and per forest size 1..7, how to cut the binary tree into 8-wide nodes
//   6. k_emit          top-down, one level per launch: materialise 128-byte quantised 8-wide
//                      nodes, octant-ordered child slots, leaf triangles copied into leaf order
// All traffic is streaming/coalesced except the unavoidable gathers (vertex fetch, leaf copy).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdio.h>
#include <float.h>
#include <vector>
#include "rq_device.h"

#define RQ_FLT_LARGE 1.844E18f      // common/math/constants.h:34 (vertex validity bound)

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = (int)e_; goto fail; } } while (0)

static unsigned long long g_launches = 0;
unsigned long long rqLaunchCount(void) { return g_launches; }
void rqCountLaunch(unsigned n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

namespace {

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {          // order-preserving float -> uint
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
__device__ __forceinline__ float warpMin(float v) {
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warpMax(float v) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__host__ __device__ __forceinline__ float halfArea(float dx, float dy, float dz) {
  return dx * (dy + dz) + dy * dz;                            // common/math/vec3.h halfArea
}

struct Bounds12 {                                             // 12 ordered-uint slots in global memory (+ statistics for the pre-split)
  uint32_t sceneLo[3], sceneHi[3], centLo[3], centHi[3];
  float    extSum;                                            // sum over valid primitives of the largest box extent
  uint32_t numValid;
};

// ----------------------------------------------------------------------------------------------
// 1. primitive setup
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_setup_prims(const RQGeomDesc* __restrict__ geoms, int numGeoms, uint32_t N,
              RQTri* __restrict__ trisIn, Bounds12* bounds, uint32_t* invalidCount, uint32_t* __restrict__ idxOut /* compact layout: 3 pool indices per triangle, else NULL */) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool valid = false;
  if (g < N) {
    int a = 0, b = numGeoms - 1;                              // last mesh with primBase <= g
    while (a < b) { int m = (a + b + 1) >> 1; if (geoms[m].primBase <= g) a = m; else b = m - 1; }
    const RQGeomDesc G = geoms[a];
    const uint32_t local = g - G.primBase;
    RQTri t;
    t.primID = local; t.geomID = G.geomID; t.pad = RQ_PAD_INVALID;
    for (int k = 0; k < 3; k++) { t.v0[k] = 0.f; t.v1[k] = 0.f; t.v2[k] = 0.f; }
    uint32_t i0 = RQ_INVALID, i1 = RQ_INVALID, i2 = RQ_INVALID;
    if (G.type == 1u) {
      // instance primitive: v0 = lower, v1 = upper corner of its world bounds (v2 = lower again), so every
      // later stage that takes min/max over the three "vertices" sees exactly that box
      valid = true;
      for (int k = 0; k < 3; k++) {
        t.v0[k] = G.lo[k]; t.v1[k] = G.hi[k]; t.v2[k] = G.lo[k];
        valid &= (G.lo[k] > -RQ_FLT_LARGE) & (G.hi[k] < RQ_FLT_LARGE) & (G.lo[k] <= G.hi[k]);   // scene_instance.h:92-95 (isvalid)
      }
      if (valid) {
        t.pad = RQ_PAD_INSTANCE | G.instIndex;
        for (int k = 0; k < 3; k++) { lo[k] = G.lo[k]; hi[k] = G.hi[k]; clo[k] = chi[k] = 0.5f * lo[k] + 0.5f * hi[k]; }
      }
    } else if (G.type == 2u) {
      // quad q = two triangles (v0,v1,v3) and (v2,v3,v1); the quad is dropped as a whole unless all four vertices are valid
      // (scene_quad_mesh.h:131-154)
      const uint32_t* ip = (const uint32_t*)(G.indices + (size_t)(local >> 1) * G.indexStride);
      const uint32_t q0 = ip[0], q1 = ip[1], q2 = ip[2], q3 = ip[3];
      t.primID = local >> 1;
      if (q0 < G.numVerts && q1 < G.numVerts && q2 < G.numVerts && q3 < G.numVerts) {
        const uint32_t other = (local & 1u) ? q0 : q2;         // the vertex this half does not use must be valid too
        const float* po = (const float*)(G.vertices + (size_t)other * G.vertexStride);
        bool ov = true;
        for (int k = 0; k < 3; k++) ov &= (po[k] > -RQ_FLT_LARGE) & (po[k] < RQ_FLT_LARGE);
        if (ov) { if (local & 1u) { i0 = q2; i1 = q3; i2 = q1; } else { i0 = q0; i1 = q1; i2 = q3; } }
      }
    } else {
      const uint32_t* ip = (const uint32_t*)(G.indices + (size_t)local * G.indexStride);
      i0 = ip[0]; i1 = ip[1]; i2 = ip[2];
    }
    if (G.type != 1u && i0 < G.numVerts && i1 < G.numVerts && i2 < G.numVerts) {
      const float* p0 = (const float*)(G.vertices + (size_t)i0 * G.vertexStride);
      const float* p1 = (const float*)(G.vertices + (size_t)i1 * G.vertexStride);
      const float* p2 = (const float*)(G.vertices + (size_t)i2 * G.vertexStride);
      valid = true;
      for (int k = 0; k < 3; k++) {
        t.v0[k] = p0[k]; t.v1[k] = p1[k]; t.v2[k] = p2[k];
        valid &= (t.v0[k] > -RQ_FLT_LARGE) & (t.v0[k] < RQ_FLT_LARGE);   // NaN fails both
        valid &= (t.v1[k] > -RQ_FLT_LARGE) & (t.v1[k] < RQ_FLT_LARGE);
        valid &= (t.v2[k] > -RQ_FLT_LARGE) & (t.v2[k] < RQ_FLT_LARGE);
      }
      if (valid) {
        t.pad = (G.type == 2u && (local & 1u)) ? RQ_PAD_FLIPUV : 0u;
        for (int k = 0; k < 3; k++) {
          lo[k] = fminf(fminf(t.v0[k], t.v1[k]), t.v2[k]);
          hi[k] = fmaxf(fmaxf(t.v0[k], t.v1[k]), t.v2[k]);
          clo[k] = chi[k] = 0.5f * lo[k] + 0.5f * hi[k];
        }
      }
    }
    if (idxOut) {
      const bool in = G.type != 1u && i0 < G.numVerts && i1 < G.numVerts && i2 < G.numVerts;
      idxOut[3 * (size_t)g + 0] = in ? G.vertBase + i0 : 0u; idxOut[3 * (size_t)g + 1] = in ? G.vertBase + i1 : 0u; idxOut[3 * (size_t)g + 2] = in ? G.vertBase + i2 : 0u;
    }
    // three 16-byte stores
    float4* dst = (float4*)(trisIn + g);
    dst[0] = make_float4(t.v0[0], t.v0[1], t.v0[2], t.v1[0]);
    dst[1] = make_float4(t.v1[1], t.v1[2], t.v2[0], t.v2[1]);
    dst[2] = make_float4(t.v2[2], __uint_as_float(t.primID), __uint_as_float(t.geomID), __uint_as_float(t.pad));
  }
  // block reduction: warp shuffles, shared memory across the 8 warps, then one atomic per block and slot
  // (one atomic set per WARP meant 312 K x 12 same-address L2 atomics for 10 M triangles: 0.6 ms of the phase)
  __shared__ float red[8][12];
  __shared__ unsigned redInvalid[8];
  __shared__ float redExt[8];
  __shared__ unsigned redValid[8];
  const unsigned nInvalid = __popc(__ballot_sync(0xffffffffu, (g < N) && !valid));
  const unsigned nValid = __popc(__ballot_sync(0xffffffffu, valid));
  float extMaxAxis = valid ? fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]) : 0.f;   // before the warp reductions overwrite lo / hi
  for (int o = 16; o; o >>= 1) extMaxAxis += __shfl_xor_sync(0xffffffffu, extMaxAxis, o);
  for (int k = 0; k < 3; k++) {
    lo[k] = warpMin(lo[k]); hi[k] = warpMax(hi[k]); clo[k] = warpMin(clo[k]); chi[k] = warpMax(chi[k]);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    for (int k = 0; k < 3; k++) { red[warp][k] = lo[k]; red[warp][3 + k] = hi[k]; red[warp][6 + k] = clo[k]; red[warp][9 + k] = chi[k]; }
    redInvalid[warp] = nInvalid; redExt[warp] = extMaxAxis; redValid[warp] = nValid;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    const bool isMin = threadIdx.x < 3 || (threadIdx.x >= 6 && threadIdx.x < 9);
    float v = red[0][threadIdx.x];
    for (int w = 1; w < 8; w++) v = isMin ? fminf(v, red[w][threadIdx.x]) : fmaxf(v, red[w][threadIdx.x]);
    uint32_t* dst = threadIdx.x < 3 ? &bounds->sceneLo[threadIdx.x] : threadIdx.x < 6 ? &bounds->sceneHi[threadIdx.x - 3]
                  : threadIdx.x < 9 ? &bounds->centLo[threadIdx.x - 6] : &bounds->centHi[threadIdx.x - 9];
    // an all-invalid block holds +/-FLT_MAX sentinels: they never win against a real bound
    if (isMin) { if (v < FLT_MAX) atomicMin(dst, f2ord(v)); } else { if (v > -FLT_MAX) atomicMax(dst, f2ord(v)); }
  }
  if (threadIdx.x == 12) {
    unsigned tot = 0;
    for (int w = 0; w < 8; w++) tot += redInvalid[w];
    if (tot) atomicAdd(invalidCount, tot);
  }
  if (threadIdx.x == 13) {
    float e = 0.f; unsigned v = 0;
    for (int w = 0; w < 8; w++) { e += redExt[w]; v += redValid[w]; }
    if (v) { atomicAdd(&bounds->extSum, e); atomicAdd(&bounds->numValid, v); }
  }
}

// vertex pool of a compact image: every mesh's vertices (arbitrary stride) as one float4 array, meshes in descriptor order
__global__ void __launch_bounds__(256)
k_copy_verts(const RQGeomDesc* __restrict__ geoms, int numGeoms, uint32_t totalVerts, float4* __restrict__ pool) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= totalVerts) return;
  int a = 0, b = numGeoms - 1;                                // last mesh with vertBase <= v (instances hold no vertices: numVerts = 0)
  while (a < b) { int m = (a + b + 1) >> 1; if (geoms[m].vertBase <= v) a = m; else b = m - 1; }
  const RQGeomDesc& G = geoms[a];
  const uint32_t local = v - G.vertBase;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (G.vertices != nullptr && local < G.numVerts) {
    const float* p = (const float*)(G.vertices + (size_t)local * G.vertexStride);
    o = make_float4(p[0], p[1], p[2], 0.f);
  }
  pool[v] = o;
}

// ----------------------------------------------------------------------------------------------
// 1b. Pre-split of large triangles (RTC_BUILD_QUALITY_HIGH).  Reference: the spatial-split builders the reference selects at
//     HIGH quality (BVHNBuilderFastSpatialSAH, kernels/bvh/bvh_builder_sah_spatial.cpp; primrefgen_presplit.h splits the
//     primitives with the highest priority on a grid before the build; splitter.h TriangleSplitter clips a triangle at a plane).
//     Here: a triangle whose box is more than PRESPLIT_FACTOR x longer than the average primitive box AND leaves more than
//     PRESPLIT_EMPTY of that box's surface empty (the reference's priority: area(box) - projected triangle area) is cut along a uniform grid
//     (cell = PRESPLIT_CELL x the average extent, doubled until the triangle spans at most 64 cells); each non-empty piece
//     becomes its own primitive REFERENCE -- a copy of the 48-byte record plus the tight box of the clipped polygon -- so a
//     long triangle no longer drags one huge box through the hierarchy.  The leaves then hold the triangle once per reference;
//     a ray may test it twice, answers do not change.
// ----------------------------------------------------------------------------------------------
constexpr float PRESPLIT_FACTOR = 8.0f, PRESPLIT_CELL = 4.0f, PRESPLIT_EMPTY = 0.6f;
constexpr int TLS_THREADS_PRE = 256;                          // tiles of 1024 values, like the treelet compaction (k_treelet_scan scans the tile sums)
constexpr int PRESPLIT_MAX_CELLS = 64;

// Sutherland-Hodgman against one axis-aligned half space (inclusive: a vertex on the plane belongs to both sides)
__device__ int clipAxis(const float (*in)[3], int n, int axis, float pos, bool keepAbove, float (*out)[3]) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    const float* a = in[i]; const float* b = in[(i + 1) % n];
    const bool ia = keepAbove ? a[axis] >= pos : a[axis] <= pos, ib = keepAbove ? b[axis] >= pos : b[axis] <= pos;
    if (ia) { out[m][0] = a[0]; out[m][1] = a[1]; out[m][2] = a[2]; m++; }
    if (ia != ib) {
      const float tpar = (pos - a[axis]) / (b[axis] - a[axis]);
      for (int c = 0; c < 3; c++) out[m][c] = c == axis ? pos : a[c] + tpar * (b[c] - a[c]);
      m++;
    }
  }
  return m;
}

// Calls emit(lo, hi) for every non-empty piece of the triangle; returns the number of pieces (>= 1).
template <typename Emit>
__device__ int presplitTriangle(const float v[3][3], const float sceneLo[3], float avgExt, Emit emit) {
  float lo[3], hi[3];
  for (int c = 0; c < 3; c++) { lo[c] = fminf(fminf(v[0][c], v[1][c]), v[2][c]); hi[c] = fmaxf(fmaxf(v[0][c], v[1][c]), v[2][c]); }
  const float emax = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
  if (!(avgExt > 0.f) || !(emax > PRESPLIT_FACTOR * avgExt)) { emit(lo, hi); return 1; }
  {
    // The reference's split priority (primrefgen_presplit.h:45-56): how much of the box's surface the triangle does NOT account
    // for, area(box) - projected primitive area (priminfo.h:13-19: |d.x| + |d.y| + |d.z|, d = cross of two edges).  An axis-aligned
    // floor triangle fills half of its box's surface and gains nothing from being cut; a diagonal sliver leaves almost all of it empty.
    const float e0[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]}, e1[3] = {v[2][0] - v[0][0], v[2][1] - v[0][1], v[2][2] - v[0][2]};
    const float areaPrim = fabsf(e0[1] * e1[2] - e0[2] * e1[1]) + fabsf(e0[2] * e1[0] - e0[0] * e1[2]) + fabsf(e0[0] * e1[1] - e0[1] * e1[0]);
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    const float areaBox = 2.0f * (dx * dy + dy * dz + dz * dx);
    if (!(areaBox - areaPrim > PRESPLIT_EMPTY * areaBox)) { emit(lo, hi); return 1; }
  }
  float L = PRESPLIT_CELL * avgExt;
  int i0[3], i1[3];
  for (;;) {
    long long cells = 1;
    for (int c = 0; c < 3; c++) {
      i0[c] = (int)floorf((lo[c] - sceneLo[c]) / L); i1[c] = (int)floorf((hi[c] - sceneLo[c]) / L);
      cells *= (long long)(i1[c] - i0[c] + 1);
    }
    if (cells <= PRESPLIT_MAX_CELLS) break;
    L *= 2.0f;
  }
  int pieces = 0;
  float P0[10][3], P1[10][3];
  for (int iz = i0[2]; iz <= i1[2]; iz++) for (int iy = i0[1]; iy <= i1[1]; iy++) for (int ix = i0[0]; ix <= i1[0]; ix++) {
    const int ic[3] = {ix, iy, iz};
    float cl[3], ch[3];
    for (int c = 0; c < 3; c++) { cl[c] = sceneLo[c] + (float)ic[c] * L; ch[c] = sceneLo[c] + (float)(ic[c] + 1) * L; }
    int n = 3;
    for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) P0[k][c] = v[k][c];
    for (int c = 0; c < 3 && n > 0; c++) {
      n = clipAxis(P0, n, c, cl[c], true, P1);
      if (n > 0) n = clipAxis(P1, n, c, ch[c], false, P0);
    }
    if (n <= 0) continue;
    float plo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, phi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = 0; k < n; k++) for (int c = 0; c < 3; c++) { plo[c] = fminf(plo[c], P0[k][c]); phi[c] = fmaxf(phi[c], P0[k][c]); }
    // conservative: pad by more than the rounding of the clip arithmetic, never beyond the triangle's own box
    for (int c = 0; c < 3; c++) {
      const float d = 2e-6f * (fabsf(plo[c]) + fabsf(phi[c]) + (hi[c] - lo[c]));
      plo[c] = fmaxf(plo[c] - d, lo[c]); phi[c] = fminf(phi[c] + d, hi[c]);
    }
    emit(plo, phi);
    pieces++;
  }
  if (pieces == 0) { emit(lo, hi); pieces = 1; }                  // cannot happen (every vertex lies in a visited cell); keep the triangle anyway
  return pieces;
}

__device__ __forceinline__ bool loadTriForSplit(const RQTri* __restrict__ trisIn, uint32_t g, float v[3][3]) {
  const float4* src = (const float4*)(trisIn + g);
  const float4 a = src[0], b = src[1], c = src[2];
  v[0][0] = a.x; v[0][1] = a.y; v[0][2] = a.z; v[1][0] = a.w; v[1][1] = b.x; v[1][2] = b.y; v[2][0] = b.z; v[2][1] = b.w; v[2][2] = c.x;
  const uint32_t pad = __float_as_uint(c.w);
  return pad != RQ_PAD_INVALID && !(pad & RQ_PAD_INSTANCE);      // invalid primitives and instances are never split
}

__global__ void __launch_bounds__(256)
k_presplit_count(const RQTri* __restrict__ trisIn, uint32_t N, const Bounds12* __restrict__ bounds, uint32_t* __restrict__ pieces) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  float v[3][3];
  uint32_t cnt = 1u;
  if (loadTriForSplit(trisIn, g, v)) {
    const float sceneLo[3] = {ord2f(bounds->sceneLo[0]), ord2f(bounds->sceneLo[1]), ord2f(bounds->sceneLo[2])};
    const float avg = bounds->numValid ? bounds->extSum / (float)bounds->numValid : 0.f;
    cnt = (uint32_t)presplitTriangle(v, sceneLo, avg, [](const float*, const float*) {});
  }
  pieces[g] = cnt;
}

// exclusive scan of `pieces` in tiles of TLS_TILE values: tile sums here, k_treelet_scan over the sums, offsets applied in k_presplit_write
__global__ void __launch_bounds__(TLS_THREADS_PRE)
k_tile_sums(const uint32_t* __restrict__ vals, uint32_t n, uint32_t* __restrict__ tileSum) {
  __shared__ uint32_t wsum[TLS_THREADS_PRE / 32];
  const uint32_t base = blockIdx.x * (TLS_THREADS_PRE * 4) + threadIdx.x * 4;
  uint32_t c = 0;
  #pragma unroll
  for (int i = 0; i < 4; i++) if (base + i < n) c += vals[base + i];
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < TLS_THREADS_PRE / 32; w++) t += wsum[w]; tileSum[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(TLS_THREADS_PRE)
k_presplit_write(const RQTri* __restrict__ trisIn, uint32_t N, const Bounds12* __restrict__ bounds, const uint32_t* __restrict__ pieces,
                 const uint32_t* __restrict__ tileOffset, RQTri* __restrict__ trisOut, float4* __restrict__ refLo, float4* __restrict__ refHi,
                 const uint32_t* __restrict__ idxIn, uint32_t* __restrict__ idxOut) {
  __shared__ uint32_t wsum[TLS_THREADS_PRE / 32];
  const uint32_t base = blockIdx.x * (TLS_THREADS_PRE * 4) + threadIdx.x * 4;
  uint32_t c = 0;
  #pragma unroll
  for (int i = 0; i < 4; i++) if (base + i < N) c += pieces[base + i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = c;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t off = tileOffset[blockIdx.x] + incl - c;
  for (int w = 0; w < warp; w++) off += wsum[w];
  const float sceneLo[3] = {ord2f(bounds->sceneLo[0]), ord2f(bounds->sceneLo[1]), ord2f(bounds->sceneLo[2])};
  const float avg = bounds->numValid ? bounds->extSum / (float)bounds->numValid : 0.f;
  for (int i = 0; i < 4; i++) {
    const uint32_t g = base + i;
    if (g >= N) break;
    const float4* src = (const float4*)(trisIn + g);
    const float4 a = src[0], b = src[1], cc = src[2];
    float v[3][3];
    const bool splittable = loadTriForSplit(trisIn, g, v);
    uint32_t o = off;
    auto put = [&](const float* lo, const float* hi) {
      float4* dst = (float4*)(trisOut + o);
      dst[0] = a; dst[1] = b; dst[2] = cc;
      refLo[o] = make_float4(lo[0], lo[1], lo[2], 0.f); refHi[o] = make_float4(hi[0], hi[1], hi[2], 0.f);
      if (idxOut) { idxOut[3 * (size_t)o] = idxIn[3 * (size_t)g]; idxOut[3 * (size_t)o + 1] = idxIn[3 * (size_t)g + 1]; idxOut[3 * (size_t)o + 2] = idxIn[3 * (size_t)g + 2]; }
      o++;
    };
    if (splittable) presplitTriangle(v, sceneLo, avg, put);
    else {                                                        // kept as it is: its box is what its three "vertices" span (instances: lower / upper corner)
      float lo[3], hi[3];
      for (int k = 0; k < 3; k++) { lo[k] = fminf(fminf(v[0][k], v[1][k]), v[2][k]); hi[k] = fmaxf(fmaxf(v[0][k], v[1][k]), v[2][k]); }
      put(lo, hi);
    }
    off += pieces[g];
  }
}

// ----------------------------------------------------------------------------------------------
// 2. Morton codes
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t expand21(uint32_t v) {    // spread 21 bits to every third bit
  uint64_t x = v & 0x1FFFFFull;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8)  & 0x100F00F00F00F00Full;
  x = (x | x << 4)  & 0x10C30C30C30C30C3ull;
  x = (x | x << 2)  & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(256)
k_morton(const RQTri* __restrict__ trisIn, uint32_t N, const Bounds12* __restrict__ bounds,
         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int cubic, uint64_t keyMask,
         const float4* __restrict__ refLo, const float4* __restrict__ refHi /* pre-split references: their own boxes (codes then span the SCENE bounds), else NULL */) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const float4* src = (const float4*)(trisIn + g);
  const float4 a = src[0], b = src[1], c = src[2];
  uint64_t key = ~0ull;                                       // invalid primitives sort to the end
  if (__float_as_uint(c.w) != RQ_PAD_INVALID) {
    const float v0[3] = {a.x, a.y, a.z}, v1[3] = {a.w, b.x, b.y}, v2[3] = {b.z, b.w, c.x};
    uint32_t q[3];
    // cubic: one scale for all axes (the largest centroid extent), so a Morton cell is a cube.  With per-axis scales a flat
    // scene gets cells that are as flat as the scene, and the bits of the short axis split neighbouring triangles by height
    // before the long axes have separated them (noisy terrain: 10 M-triangle scene, SAH 28.9 per-axis vs cubic).
    float extMax = 0.f;
    const uint32_t* bLo = refLo ? bounds->sceneLo : bounds->centLo;   // centres of clipped references can leave the centroid bounds of the whole triangles
    const uint32_t* bHi = refLo ? bounds->sceneHi : bounds->centHi;
    for (int k = 0; k < 3; k++) extMax = fmaxf(extMax, ord2f(bHi[k]) - ord2f(bLo[k]));
    for (int k = 0; k < 3; k++) {
      float lo = fminf(fminf(v0[k], v1[k]), v2[k]), hi = fmaxf(fmaxf(v0[k], v1[k]), v2[k]);
      if (refLo) { lo = k == 0 ? refLo[g].x : k == 1 ? refLo[g].y : refLo[g].z; hi = k == 0 ? refHi[g].x : k == 1 ? refHi[g].y : refHi[g].z; }
      const float cen = 0.5f * lo + 0.5f * hi;
      const float cl = ord2f(bLo[k]), ch = ord2f(bHi[k]);
      const float ext = cubic ? extMax : ch - cl;
      float x = ext > 0.f ? (cen - cl) / ext : 0.f;
      x = fminf(fmaxf(x, 0.f), 1.f);
      q[k] = min(2097151u, (uint32_t)(x * 2097152.0f));
    }
    key = ((expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2])) & keyMask;
  }
  keys[g] = key;
  vals[g] = g;
}

// ----------------------------------------------------------------------------------------------
// 3. LSD radix sort, 8-bit digits.  Tile = 8 warps x 8 items x 32 lanes = 2048 keys, laid out
//    warp-major (warp w owns keys [w*256, w*256+256) of the tile, item-major inside) so that the
//    per-warp running digit counters give a stable rank.
// ----------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const uint64_t* __restrict__ keys, uint32_t N, int shift, uint32_t numTiles,
            uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * SORT_TILE;
  #pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t idx = base + i * SORT_THREADS + threadIdx.x;
    if (idx < N) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * numTiles + blockIdx.x] = h[threadIdx.x];   // digit-major
}

// one block per digit: exclusive scan of that digit's row, row total to digitTotal
__global__ void __launch_bounds__(256)
k_sort_scan_rows(uint32_t* __restrict__ hist, uint32_t numTiles, uint32_t* __restrict__ digitTotal) {
  __shared__ uint32_t part[256];
  uint32_t* row = hist + (size_t)blockIdx.x * numTiles;
  const uint32_t chunk = (numTiles + 255u) / 256u;
  const uint32_t b = threadIdx.x * chunk, e = min(b + chunk, numTiles);
  uint32_t s = 0;
  for (uint32_t i = b; i < e; i++) s += row[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {                          // Hillis-Steele inclusive scan
    uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;                        // exclusive prefix of this chunk
  for (uint32_t i = b; i < e; i++) { const uint32_t v = row[i]; row[i] = run; run += v; }
  if (threadIdx.x == 255) digitTotal[blockIdx.x] = part[255];
}

__global__ void __launch_bounds__(256)
k_sort_scan_digits(uint32_t* __restrict__ digitTotal) {          // 256 totals -> exclusive bases
  __shared__ uint32_t part[256];
  const uint32_t s = digitTotal[threadIdx.x];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  digitTotal[threadIdx.x] = part[threadIdx.x] - s;
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const uint64_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
               uint64_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t N, int shift,
               uint32_t numTiles, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ digitBase) {
  __shared__ uint32_t cnt[SORT_THREADS / 32][256];
  __shared__ uint32_t gbase[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (SORT_THREADS / 32) * 256; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
  gbase[threadIdx.x] = digitBase[threadIdx.x] + hist[threadIdx.x * numTiles + blockIdx.x];
  __syncthreads();
  const uint32_t wbase = blockIdx.x * SORT_TILE + warp * (32 * SORT_ITEMS);
  uint64_t key[SORT_ITEMS]; uint32_t val[SORT_ITEMS], rank[SORT_ITEMS];
  #pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t idx = wbase + i * 32 + lane;
    const bool ok = idx < N;
    key[i] = ok ? keysIn[idx] : 0; val[i] = ok ? valsIn[idx] : 0; rank[i] = 0;
    const unsigned vmask = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
      const unsigned peers = __match_any_sync(vmask, d);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(peers); }
      old = __shfl_sync(peers, old, leader);
      rank[i] = old + __popc(peers & ((1u << lane) - 1u));
    }
    __syncwarp();
  }
  __syncthreads();
  {                                                             // exclusive prefix over the 8 warps
    uint32_t run = 0;
    #pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) { const uint32_t t = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += t; }
  }
  __syncthreads();
  #pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t idx = wbase + i * 32 + lane;
    if (idx < N) {
      const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
      const uint32_t pos = gbase[d] + cnt[warp][d] + rank[i];
      keysOut[pos] = key[i]; valsOut[pos] = val[i];
    }
  }
}

// NOTE: k_sort_hist tiles keys as base + i*256 + thread (order inside a tile is irrelevant for a
// histogram); k_sort_scatter uses the same tile range [blockIdx*2048, +2048).

// ----------------------------------------------------------------------------------------------
// 4. binary radix tree (Karras 2012) over sorted 63-bit codes, ties broken by position.
//    Node ids: inner i in [0, n-1), leaf k -> (n-1)+k.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256)
k_hierarchy(const uint64_t* __restrict__ keys, int n, uint32_t* __restrict__ left, uint32_t* __restrict__ right,
            uint32_t* __restrict__ parent, uint32_t* __restrict__ rangeFirst, uint32_t* __restrict__ rangeLast = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const uint32_t L = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
  const uint32_t R = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
  left[i] = L; right[i] = R; rangeFirst[i] = (uint32_t)lo;
  if (rangeLast) rangeLast[i] = (uint32_t)hi;
  parent[L] = (uint32_t)i; parent[R] = (uint32_t)i;
  if (i == 0) parent[0] = RQ_INVALID;
}

// ----------------------------------------------------------------------------------------------
// 5. bottom-up refit + SAH collapse programme.
//    For binary node x and forest size i (1..7):  C(x,i) = cheapest way to represent x's subtree
//    as at most i roots of the 8-wide tree.
//      leaf(x)      = A(x) * T(x) * costTri            if T(x) <= maxLeafTris
//      split(x,j)   = min_k C(left,k) + C(right,j-k)
//      C(x,1)       = min(leaf(x), split(x,8) + A(x)*costNode)
//      C(x,i)       = min(split(x,i), C(x,i-1))
//    dec word: bit0 = "C(x,1) is an inner node"; bits [3(j-1), 3(j-1)+2] for j=2..8 = k chosen
//    for split(x,j) (0 = fall back to C(x,j-1)).
// ----------------------------------------------------------------------------------------------
struct B2 {                    // binary-tree arrays (2n-1 nodes unless noted)
  float4* lo;                  // xyz = lower, w = half area
  float4* hi;                  // xyz = upper, w = triangle count (uint bits)
  float*  cost;                // 8 floats per node, [0..6] = C(x,1..7)
  uint32_t* dec;
  uint32_t* left;              // n-1
  uint32_t* right;             // n-1
  uint32_t* parent;            // 2n-1
  uint32_t* rangeFirst;        // n-1
  uint32_t* flag;              // n-1 arrival counters
  const float4* refLo;         // pre-split (RTC_BUILD_QUALITY_HIGH): box of primitive reference i, indexed like trisIn; NULL otherwise
  const float4* refHi;
};

// Leaf k of the binary tree (node id n-1+k): bounds of its triangle and the trivial collapse programme.
__device__ __forceinline__ void initLeaf(const B2& t, uint32_t node, const RQTri* __restrict__ trisIn, uint32_t tri, float costTri) {
  const float4* src = (const float4*)(trisIn + tri);
  const float4 a = src[0], b = src[1], c = src[2];
  float lx = fminf(fminf(a.x, a.w), b.z), ly = fminf(fminf(a.y, b.x), b.w), lz = fminf(fminf(a.z, b.y), c.x);
  float hx = fmaxf(fmaxf(a.x, a.w), b.z), hy = fmaxf(fmaxf(a.y, b.x), b.w), hz = fmaxf(fmaxf(a.z, b.y), c.x);
  if (t.refLo) { const float4 rl = t.refLo[tri], rh = t.refHi[tri]; lx = rl.x; ly = rl.y; lz = rl.z; hx = rh.x; hy = rh.y; hz = rh.z; }
  const float A = halfArea(hx - lx, hy - ly, hz - lz);
  t.lo[node] = make_float4(lx, ly, lz, A);
  t.hi[node] = make_float4(hx, hy, hz, __uint_as_float(1u));
  const float cl = A * costTri;
  float4* cp = (float4*)(t.cost + (size_t)node * 8);
  cp[0] = make_float4(cl, cl, cl, cl);
  cp[1] = make_float4(cl, cl, cl, 0.f);
  t.dec[node] = 0;
}

// Bounds, triangle count and the collapse programme of binary node `cur` from its finished children L, R.
__device__ __forceinline__ void combineNode(const B2& t, uint32_t cur, uint32_t L, uint32_t R,
                                            float costNode, float costTri, int maxLeafTris) {
    // data produced by other threads / SMs in this launch: read through L2 (ld.cg), never L1
    const float4 llo = __ldcg(t.lo + L), lhi = __ldcg(t.hi + L), rlo = __ldcg(t.lo + R), rhi = __ldcg(t.hi + R);
    const float4 l0 = __ldcg((const float4*)(t.cost + (size_t)L * 8)), l1 = __ldcg((const float4*)(t.cost + (size_t)L * 8) + 1);
    const float4 r0 = __ldcg((const float4*)(t.cost + (size_t)R * 8)), r1 = __ldcg((const float4*)(t.cost + (size_t)R * 8) + 1);
    const float cl[8] = {0.f, l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z};
    const float cr[8] = {0.f, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z};
    const float lx = fminf(llo.x, rlo.x), ly = fminf(llo.y, rlo.y), lz = fminf(llo.z, rlo.z);
    const float hx = fmaxf(lhi.x, rhi.x), hy = fmaxf(lhi.y, rhi.y), hz = fmaxf(lhi.z, rhi.z);
    const float A = halfArea(hx - lx, hy - ly, hz - lz);
    const uint32_t T = __float_as_uint(lhi.w) + __float_as_uint(rhi.w);
    float dist[9]; uint32_t kb[9];
    #pragma unroll
    for (int j = 2; j <= 8; j++) {
      float best = FLT_MAX; uint32_t bk = 1;
      #pragma unroll
      for (int kk = 1; kk <= 7; kk++) {
        if (kk < j && j - kk <= 7) {
          const float v = cl[kk] + cr[j - kk];
          if (v < best) { best = v; bk = kk; }
        }
      }
      dist[j] = best; kb[j] = bk;
    }
    const float leafCost = (T <= (uint32_t)maxLeafTris) ? A * (float)T * costTri : FLT_MAX;
    const float innerCost = dist[8] + A * costNode;
    float c[8];
    uint32_t dec = (kb[8] << 21);
    if (T > (uint32_t)maxLeafTris || innerCost < leafCost) { c[1] = innerCost; dec |= 1u; } else c[1] = leafCost;
    #pragma unroll
    for (int i = 2; i <= 7; i++) {
      if (dist[i] < c[i - 1]) { c[i] = dist[i]; dec |= kb[i] << (3 * (i - 1)); }
      else c[i] = c[i - 1];
    }
    t.lo[cur] = make_float4(lx, ly, lz, A);
    t.hi[cur] = make_float4(hx, hy, hz, __uint_as_float(T));
    float4* cp = (float4*)(t.cost + (size_t)cur * 8);
    cp[0] = make_float4(c[1], c[2], c[3], c[4]);
    cp[1] = make_float4(c[5], c[6], c[7], 0.f);
    t.dec[cur] = dec;
}

__global__ void __launch_bounds__(256)
k_refit_dp(B2 t, int n, const RQTri* __restrict__ trisIn, const uint32_t* __restrict__ vals,
           float costNode, float costTri, int maxLeafTris, int initLeaves = 1) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t node = (uint32_t)(n - 1 + k);
  if (initLeaves) initLeaf(t, node, trisIn, vals[k], costTri);
  __threadfence();
  uint32_t cur = t.parent[node];
  while (cur != RQ_INVALID) {
    if (atomicAdd(&t.flag[cur], 1u) == 0u) return;             // first arrival: sibling not ready yet
    __threadfence();
    combineNode(t, cur, t.left[cur], t.right[cur], costNode, costTri, maxLeafTris);
    __threadfence();
    cur = t.parent[cur];
  }
}

// ----------------------------------------------------------------------------------------------
// 5b. PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) as the alternative to the
//     radix tree: clusters sit in Morton order; every iteration each cluster looks R positions to
//     either side for the neighbour with the smallest merged surface area, mutual nearest
//     neighbours merge, the array is compacted (order preserved).  Because a merge only ever joins
//     finished clusters, the bounds and the collapse programme of the new node are computed right
//     there -- no separate refit pass.  Inner node ids are handed out downwards from n-2, so the
//     last merge (the root) is node 0 as in the radix tree.
// ----------------------------------------------------------------------------------------------
constexpr int PLOC_THREADS = 256;
constexpr int PLOC_MAX_RADIUS = 32;

__global__ void __launch_bounds__(256)
k_ploc_init(B2 t, int n, const RQTri* __restrict__ trisIn, const uint32_t* __restrict__ vals, float costTri,
            uint32_t* __restrict__ cid) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t node = (uint32_t)(n - 1 + k);
  initLeaf(t, node, trisIn, vals[k], costTri);
  cid[k] = node;
}

__global__ void __launch_bounds__(PLOC_THREADS)
k_ploc_nn(B2 t, const uint32_t* __restrict__ cid, const uint32_t* __restrict__ mPtr, int radius, uint32_t* __restrict__ nn) {
  __shared__ float sb[PLOC_THREADS + 2 * PLOC_MAX_RADIUS][6];
  const uint32_t m = *mPtr;                                     // cluster count of this iteration lives on the device (no host round trip)
  if (blockIdx.x * PLOC_THREADS >= m) return;                   // the grid is sized for an upper bound
  const int b0 = (int)(blockIdx.x * PLOC_THREADS) - radius;
  const int span = PLOC_THREADS + 2 * radius;
  for (int s = threadIdx.x; s < span; s += PLOC_THREADS) {
    const int g = b0 + s;
    if (g >= 0 && g < (int)m) {
      const uint32_t c = cid[g];
      const float4 lo = t.lo[c], hi = t.hi[c];
      sb[s][0] = lo.x; sb[s][1] = lo.y; sb[s][2] = lo.z; sb[s][3] = hi.x; sb[s][4] = hi.y; sb[s][5] = hi.z;
    }
  }
  __syncthreads();
  const int i = (int)(blockIdx.x * PLOC_THREADS + threadIdx.x);
  if (i >= (int)m) return;
  const int si = (int)threadIdx.x + radius;
  const float lx = sb[si][0], ly = sb[si][1], lz = sb[si][2], hx = sb[si][3], hy = sb[si][4], hz = sb[si][5];
  float best = FLT_MAX; int bj = -1;
  const int j0 = max(i - radius, 0), j1 = min(i + radius, (int)m - 1);
  for (int j = j0; j <= j1; j++) {
    if (j == i) continue;
    const int sj = j - b0;
    const float d = halfArea(fmaxf(hx, sb[sj][3]) - fminf(lx, sb[sj][0]), fmaxf(hy, sb[sj][4]) - fminf(ly, sb[sj][1]),
                             fmaxf(hz, sb[sj][5]) - fminf(lz, sb[sj][2]));
    // ties go to the "buddy" position i ^ 1, else to the lower position.  With the lower position alone, a run of identical boxes
    // (70 000 copies of one triangle: tests/test_gpu_scale.py::test_degenerate_inputs_build_and_answer) produced ONE mutual pair
    // per iteration -- everybody pointed at the start of the run -- and the build gave up at its iteration bound; with the buddy rule
    // positions 2k and 2k+1 choose each other and such a run halves every iteration.
    if (d < best || (d == best && j == (i ^ 1))) { best = d; bj = j; }
  }
  nn[i] = (uint32_t)bj;
}

// merge mutual pairs (the lower position keeps the new cluster), flag survivors, count them per block
__global__ void __launch_bounds__(PLOC_THREADS)
k_ploc_merge(B2 t, uint32_t* __restrict__ cid, const uint32_t* __restrict__ mPtr, const uint32_t* __restrict__ nn, uint32_t* __restrict__ keep,
             uint32_t* blockCount, uint32_t* nextInner, float costNode, float costTri, int maxLeafTris,
             uint32_t* mNext, uint32_t* iterations, uint32_t* blocksDone) {
  __shared__ uint32_t part[PLOC_THREADS];
  __shared__ bool isLast;
  const uint32_t m = *mPtr;
  if (blockIdx.x * PLOC_THREADS >= m) return;                   // block-uniform: nobody reaches the barrier below
  const uint32_t i = blockIdx.x * PLOC_THREADS + threadIdx.x;
  bool alive = false;
  if (i < m) {
    const uint32_t j = nn[i];
    const bool mutual = j < m && nn[j] == i;
    alive = !(mutual && j < i);
    if (mutual && i < j) {
      const uint32_t id = atomicSub(nextInner, 1u);            // n-2, n-3, ... 0
      const uint32_t L = cid[i], R = cid[j];
      t.left[id] = L; t.right[id] = R; t.parent[L] = id; t.parent[R] = id;
      combineNode(t, id, L, R, costNode, costTri, maxLeafTris);
      cid[i] = id;
    }
    keep[i] = alive ? 1u : 0u;
  }
  const unsigned cnt = __syncthreads_count(alive);
  // The block that finishes last turns the per-block survivor counts into exclusive offsets and publishes the next cluster count
  // (threadfence reduction): no separate one-block scan launch per iteration.
  const uint32_t numBlocks = (m + PLOC_THREADS - 1) / PLOC_THREADS;
  if (threadIdx.x == 0) {
    blockCount[blockIdx.x] = cnt;
    __threadfence();
    isLast = (atomicAdd(blocksDone, 1u) == numBlocks - 1u);
  }
  __syncthreads();
  if (!isLast) return;
  __threadfence();
  const uint32_t chunk = (numBlocks + PLOC_THREADS - 1) / PLOC_THREADS;
  const uint32_t b0 = threadIdx.x * chunk, e0 = min(b0 + chunk, numBlocks);
  uint32_t sum = 0;
  for (uint32_t k = b0; k < e0; k++) sum += __ldcg(blockCount + k);
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < PLOC_THREADS; o <<= 1) {                  // Hillis-Steele inclusive scan of the 256 partial sums
    const uint32_t v = (int)threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - sum;
  for (uint32_t k = b0; k < e0; k++) { const uint32_t v = __ldcg(blockCount + k); blockCount[k] = run; run += v; }
  if (threadIdx.x == PLOC_THREADS - 1) {
    *mNext = part[PLOC_THREADS - 1];
    if (m > 1) atomicAdd(iterations, 1u);
    *blocksDone = 0u;                                           // ready for the next iteration (stream ordered)
  }
}

__global__ void __launch_bounds__(PLOC_THREADS)
k_ploc_compact(const uint32_t* __restrict__ cidIn, const uint32_t* __restrict__ mPtr, const uint32_t* __restrict__ keep,
               const uint32_t* __restrict__ blockOffset, uint32_t* __restrict__ cidOut) {
  __shared__ uint32_t warpSum[PLOC_THREADS / 32];
  const uint32_t m = *mPtr;
  if (blockIdx.x * PLOC_THREADS >= m) return;
  const uint32_t i = blockIdx.x * PLOC_THREADS + threadIdx.x;
  const bool alive = i < m && keep[i] != 0u;
  const unsigned bal = __ballot_sync(0xffffffffu, alive);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warpSum[warp] = __popc(bal);
  __syncthreads();
  uint32_t off = blockOffset[blockIdx.x];
  for (int w = 0; w < warp; w++) off += warpSum[w];
  if (alive) cidOut[off + __popc(bal & ((1u << lane) - 1u))] = cidIn[i];
}

// The tail of the clustering -- and all of it when PLOC only has to join a few ten thousand treelet roots -- runs in ONE block:
// below ~64 K clusters an iteration is a handful of microseconds of work, and three launches plus a host read-back every few
// iterations cost more than the work itself (round 1: 163 launches to build a 3 K-triangle scene).  Same algorithm, same
// tie rules, block barriers instead of kernel boundaries; the cluster array ping-pongs between cidA and cidB.
constexpr int PLOC_TAIL_THREADS = 1024;
constexpr uint32_t PLOC_TAIL_MAX = 1u << 12;                    // one SM against 148: above ~4 K clusters the grid version wins (35 K treelet roots: 2.0 ms single block, 0.45 ms grids)
__global__ void __launch_bounds__(PLOC_TAIL_THREADS)
k_ploc_tail(B2 t, uint32_t* __restrict__ cidA, uint32_t* __restrict__ cidB, uint32_t* __restrict__ nn, const uint32_t* __restrict__ mPtr,
            uint32_t* nextInner, uint32_t* iterations, int radius, float costNode, float costTri, int maxLeafTris) {
  __shared__ uint32_t warpCnt[PLOC_TAIL_THREADS / 32];
  __shared__ uint32_t s_run;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t m = *mPtr;
  uint32_t* cin = cidA; uint32_t* cout = cidB;
  uint32_t its = 0;
  // the boxes of the current clusters, by position (SoA, 6 x PLOC_TAIL_MAX floats of dynamic shared memory): every iteration loads
  // them once from L2 and the neighbour search (2 x radius candidates per cluster) runs out of shared memory -- the search used to
  // fetch both float4 of every candidate through L2, ~30 latency-bound iterations = 0.32 ms of the 10 M-triangle build
  extern __shared__ float s_box[];
  while (m > 1u) {
    for (uint32_t i = tid; i < m; i += PLOC_TAIL_THREADS) {
      const float4 lo = __ldcg(t.lo + cin[i]), hi = __ldcg(t.hi + cin[i]);
      s_box[i] = lo.x; s_box[PLOC_TAIL_MAX + i] = lo.y; s_box[2 * PLOC_TAIL_MAX + i] = lo.z;
      s_box[3 * PLOC_TAIL_MAX + i] = hi.x; s_box[4 * PLOC_TAIL_MAX + i] = hi.y; s_box[5 * PLOC_TAIL_MAX + i] = hi.z;
    }
    __syncthreads();
    for (uint32_t i = tid; i < m; i += PLOC_TAIL_THREADS) {
      const float lx = s_box[i], ly = s_box[PLOC_TAIL_MAX + i], lz = s_box[2 * PLOC_TAIL_MAX + i];
      const float hx = s_box[3 * PLOC_TAIL_MAX + i], hy = s_box[4 * PLOC_TAIL_MAX + i], hz = s_box[5 * PLOC_TAIL_MAX + i];
      float best = FLT_MAX; int bj = -1;
      const int j0 = max((int)i - radius, 0), j1 = min((int)i + radius, (int)m - 1);
      for (int j = j0; j <= j1; j++) {
        if (j == (int)i) continue;
        const float d = halfArea(fmaxf(hx, s_box[3 * PLOC_TAIL_MAX + j]) - fminf(lx, s_box[j]), fmaxf(hy, s_box[4 * PLOC_TAIL_MAX + j]) - fminf(ly, s_box[PLOC_TAIL_MAX + j]),
                                 fmaxf(hz, s_box[5 * PLOC_TAIL_MAX + j]) - fminf(lz, s_box[2 * PLOC_TAIL_MAX + j]));
        if (d < best || (d == best && j == (int)(i ^ 1u))) { best = d; bj = j; }   // tie rule of k_ploc_nn
      }
      nn[i] = (uint32_t)bj;
    }
    __syncthreads();
    if (tid == 0) s_run = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += PLOC_TAIL_THREADS) {      // ordered compaction, one chunk of 1024 positions at a time
      const uint32_t i = base + tid;
      bool alive = false; uint32_t id = 0u;
      if (i < m) {
        const uint32_t j = nn[i];
        const bool mutual = j < m && nn[j] == i;
        alive = !(mutual && j < i);
        id = cin[i];
        if (mutual && i < j) {
          id = atomicSub(nextInner, 1u);
          const uint32_t L = cin[i], R = cin[j];
          t.left[id] = L; t.right[id] = R; t.parent[L] = id; t.parent[R] = id;
          combineNode(t, id, L, R, costNode, costTri, maxLeafTris);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) warpCnt[warp] = __popc(bal);
      __syncthreads();
      uint32_t off = s_run;
      for (int w = 0; w < warp; w++) off += warpCnt[w];
      if (alive) cout[off + __popc(bal & ((1u << lane) - 1u))] = id;
      __syncthreads();
      if (tid == 0) { uint32_t tot = 0; for (int w = 0; w < PLOC_TAIL_THREADS / 32; w++) tot += warpCnt[w]; s_run += tot; }
      __syncthreads();
    }
    const uint32_t next = s_run;
    __syncthreads();
    if (next >= m) break;                                            // cannot happen (the globally closest pair is always mutual); never spin
    m = next; its++;
    uint32_t* tmp = cin; cin = cout; cout = tmp;
  }
  if (tid == 0) atomicAdd(iterations, its);
}

// ----------------------------------------------------------------------------------------------
// 5c. Binned-SAH treelets (builder 2): the stage that replaces the reference's top-down binned-SAH recursion
//     (kernels/builders/bvh_builder_sah.h:222-319, heuristic_binning.h:210-256 bin, :336-392 best,
//     heuristic_binning_array_aligned.h:60-123 find / split) on the GPU.
//
//     The Morton order only decides which triangles are built TOGETHER: the radix tree over the sorted codes is cut into
//     "treelets" -- maximal subtrees of at most K triangles, i.e. the triangles of one Morton cell, contiguous in the
//     sorted array.  Inside a treelet the binary hierarchy is rebuilt top-down by the surface-area heuristic, by ONE WARP
//     working in shared memory:
//       * nodes of more than TL_SMALL triangles: 16 centroid bins on each of the 3 axes (shared-memory atomics on
//         order-preserving uint keys), prefix / suffix box scans over the bins with width-16 shuffles (lanes 0..15 scan
//         left to right, lanes 16..31 the mirrored bins, so one scan yields both sides), SAH = A_L n_L + A_R n_R for all
//         45 candidate planes at once, warp arg-min, stable partition of the treelet's index permutation;
//       * subtrees of at most TL_SMALL triangles: one THREAD each.  The stable partitions above keep every slice in Morton
//         order, so the slice is split where the highest differing bit of its first and last code flips (the radix-tree rule
//         on the sub-sequence) -- a few dozen instructions per node.  (Round-2 history: an exact sweep SAH per thread here --
//         insertion sort along each axis, all m-1 positions -- was 38 % of the kernel's 3.5 G warp instructions at 7 active
//         lanes, profiles/r02d_ncu_treelet.txt, for 1 % of traversal speed; the last levels end up inside one 8-wide node
//         whose leaf slots the collapse programme chooses anyway.)
//     Above the treelets the existing PLOC stage clusters the treelet roots (a few ten thousand boxes) and the SAH dynamic
//     programme then cuts the binary tree into 8-wide nodes as before.
//     Node ids: the T-1 nodes above the treelets take ids [0, T-1) (PLOC hands them out downwards, the root ends up as 0);
//     treelet t, which starts at sorted position a_t and holds n_t triangles, owns ids T-1 + a_t - t + [0, n_t - 1), and the
//     subtree over a range of m triangles rooted at local index j uses [j, j+m-1): left child j+1, right child j+m_left.
// ----------------------------------------------------------------------------------------------
constexpr int TL_BINS = 16;
constexpr int TL_SMALL = 32;
constexpr int TL_SWEEP = 16;                                   // RTC_BUILD_QUALITY_HIGH: nodes of at most this many triangles get an exact sweep SAH in the thread phase
constexpr uint32_t TL_ORD_PINF = 0xFF800000u;                 // f2ord(+inf)
constexpr uint32_t TL_ORD_NINF = 0x007FFFFFu;                 // f2ord(-inf)

// which sorted positions start a treelet, and how large it is (0 = not a start)
__global__ void __launch_bounds__(256)
k_treelet_mark(int n, const uint32_t* __restrict__ rparent, const uint32_t* __restrict__ rfirst, const uint32_t* __restrict__ rlast,
               uint32_t K, uint32_t* __restrict__ sizeAt) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= 2 * n - 1) return;
  const uint32_t par = rparent[x];
  const uint32_t psize = (par == RQ_INVALID) ? 0xFFFFFFFFu : rlast[par] - rfirst[par] + 1u;
  if (x < n - 1) {
    const uint32_t size = rlast[x] - rfirst[x] + 1u;
    if (size <= K && psize > K) sizeAt[rfirst[x]] = size;
  } else if (psize > K) {
    sizeAt[x - (n - 1)] = 1u;                                   // a single triangle directly below a large node
  }
}

constexpr int TLS_THREADS = 256, TLS_ITEMS = 4, TLS_TILE = TLS_THREADS * TLS_ITEMS;
__global__ void __launch_bounds__(TLS_THREADS)
k_treelet_count(const uint32_t* __restrict__ sizeAt, uint32_t n, uint32_t* __restrict__ blockCount) {
  const uint32_t base = blockIdx.x * TLS_TILE + threadIdx.x * TLS_ITEMS;
  int c = 0;
  #pragma unroll
  for (int i = 0; i < TLS_ITEMS; i++) if (base + i < n && sizeAt[base + i] != 0u) c++;
  __shared__ uint32_t wsum[TLS_THREADS / 32];
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = (uint32_t)c;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < TLS_THREADS / 32; w++) t += wsum[w]; blockCount[blockIdx.x] = t; }
}
// one block: per-block counts -> exclusive offsets (in place), total -> *total
__global__ void __launch_bounds__(1024)
k_treelet_scan(uint32_t* __restrict__ blockCount, uint32_t numBlocks, uint32_t* __restrict__ total) {
  __shared__ uint32_t part[1024];
  const uint32_t chunk = (numBlocks + 1023u) / 1024u;
  const uint32_t b0 = min(threadIdx.x * chunk, numBlocks), e0 = min(b0 + chunk, numBlocks);
  uint32_t sum = 0;
  for (uint32_t k = b0; k < e0; k++) sum += blockCount[k];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t v = (int)threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - sum;
  for (uint32_t k = b0; k < e0; k++) { const uint32_t v = blockCount[k]; blockCount[k] = run; run += v; }
  if (threadIdx.x == 1023) *total = part[1023];
}
__global__ void __launch_bounds__(TLS_THREADS)
k_treelet_write(const uint32_t* __restrict__ sizeAt, uint32_t n, const uint32_t* __restrict__ blockOffset, uint32_t* __restrict__ treeletStart) {
  __shared__ uint32_t wsum[TLS_THREADS / 32];
  const uint32_t base = blockIdx.x * TLS_TILE + threadIdx.x * TLS_ITEMS;
  bool f[TLS_ITEMS]; uint32_t c = 0;
  #pragma unroll
  for (int i = 0; i < TLS_ITEMS; i++) { f[i] = base + i < n && sizeAt[base + i] != 0u; c += f[i] ? 1u : 0u; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = c;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t off = blockOffset[blockIdx.x] + incl - c;
  for (int w = 0; w < warp; w++) off += wsum[w];
  #pragma unroll
  for (int i = 0; i < TLS_ITEMS; i++) if (f[i]) treeletStart[off++] = base + i;
}
// the clusters PLOC starts from: the root of every treelet, in Morton order
__global__ void __launch_bounds__(256)
k_treelet_roots(int n, const uint32_t* __restrict__ treeletStart, const uint32_t* __restrict__ sizeAt, uint32_t T, uint32_t* __restrict__ cid) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t a = treeletStart[t];
  cid[t] = sizeAt[a] == 1u ? (uint32_t)(n - 1) + a : (T - 1u) + (a - t);
}

template <int K, int SWEEP>
struct TreeletSmem {                                           // one per warp
  float    box[6][K];                                          // lo.xyz, hi.xyz of the treelet's triangles (SoA: lane-strided access is conflict free)
  uint32_t key[K];                                             // upper 32 bits of the Morton code (what the triangles were sorted by)
  uint16_t perm[K], perm2[K];                                  // the treelet's triangles, partitioned in place node by node
  uint32_t bins[3][TL_BINS][7];                                // per axis and bin: ordered-uint lo.xyz, hi.xyz, count
  uint32_t stack[16][2];                                       // pending large nodes: begin | end << 16, local node index
  uint32_t small[K / 2][2];                                    // subtrees left to the thread phase, same encoding
  float    keys[SWEEP ? SWEEP : 1][32];                        // thread phase (SWEEP > 0): sort keys of the lane's current node
  float    suffix[SWEEP ? SWEEP : 1][32];                      // thread phase (SWEEP > 0): right-side areas of the lane's current node
};

__device__ __forceinline__ float boxArea6(const float lo[3], const float hi[3]) {
  return halfArea(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]);
}

// one lane splits ONE node of a small subtree: the slice perm[b, e) (2 <= e - b <= TL_SMALL) with local node index j.  Radix-tree
// split of the Morton-ordered slice (Karras 2012 on the sub-sequence; equal codes are halved by position).  Returns the size of the
// left child; the caller queues the children.
template <int K, int SWEEP>
__device__ __forceinline__ uint32_t treeletSmallNode(TreeletSmem<K, SWEEP>& S, const B2& t, uint32_t b, uint32_t e, uint32_t j, uint32_t base, uint32_t leaf0, int lane) {
  {
    const uint32_t m = e - b;
    uint32_t mL = m >> 1;
    if (SWEEP > 0 && m > 2u && m <= (uint32_t)SWEEP) {
      // RTC_BUILD_QUALITY_HIGH only.  The last levels decide which two or three triangles share a leaf slot: exact sweep SAH,
      // all three axes, every split position (the slice is insertion-sorted along the axis, suffix areas in shared scratch).
      // Worth +8 % occlusion and +1.5 % closest-hit Mrays/s on the 10 M-triangle scene, but one thread per subtree runs at ~7
      // active lanes: +6 ... +15 ms of build time (profiles/r02b_ab_c3.jsonl, r02e_ab.jsonl, r02f_ab.jsonl) -- the default
      // quality (MEDIUM) keeps the radix rule below TL_SMALL triangles.
      float best = INFINITY; int bestAxis = -1; int split = (int)mL;
      const int mi = (int)m;
      for (int a = 0; a < 3; a++) {
        for (int i = 0; i < mi; i++) { const uint32_t p = S.perm[b + i]; S.keys[i][lane] = S.box[a][p] + S.box[3 + a][p]; }
        for (int i = 1; i < mi; i++) {
          const float k = S.keys[i][lane]; const uint16_t p = S.perm[b + i];
          int q = i - 1;
          while (q >= 0 && (S.keys[q][lane] > k || (S.keys[q][lane] == k && S.perm[b + q] > p))) {
            S.keys[q + 1][lane] = S.keys[q][lane]; S.perm[b + q + 1] = S.perm[b + q]; q--;
          }
          S.keys[q + 1][lane] = k; S.perm[b + q + 1] = p;
        }
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = mi - 1; i >= 1; i--) {
          const uint32_t p = S.perm[b + i];
          for (int c = 0; c < 3; c++) { lo[c] = fminf(lo[c], S.box[c][p]); hi[c] = fmaxf(hi[c], S.box[3 + c][p]); }
          S.suffix[i][lane] = boxArea6(lo, hi);
        }
        for (int c = 0; c < 3; c++) { lo[c] = INFINITY; hi[c] = -INFINITY; }
        for (int i = 1; i < mi; i++) {
          const uint32_t p = S.perm[b + i - 1];
          for (int c = 0; c < 3; c++) { lo[c] = fminf(lo[c], S.box[c][p]); hi[c] = fmaxf(hi[c], S.box[3 + c][p]); }
          const float cost = boxArea6(lo, hi) * (float)i + S.suffix[i][lane] * (float)(mi - i);
          if (cost < best) { best = cost; bestAxis = a; split = i; }
        }
      }
      if (bestAxis >= 0 && bestAxis != 2) {                     // the slice is sorted along z now: restore the winning order
        const int a = bestAxis;
        for (int i = 0; i < mi; i++) { const uint32_t p = S.perm[b + i]; S.keys[i][lane] = S.box[a][p] + S.box[3 + a][p]; }
        for (int i = 1; i < mi; i++) {
          const float k = S.keys[i][lane]; const uint16_t p = S.perm[b + i];
          int q = i - 1;
          while (q >= 0 && (S.keys[q][lane] > k || (S.keys[q][lane] == k && S.perm[b + q] > p))) {
            S.keys[q + 1][lane] = S.keys[q][lane]; S.perm[b + q + 1] = S.perm[b + q]; q--;
          }
          S.keys[q + 1][lane] = k; S.perm[b + q + 1] = p;
        }
      }
      mL = (uint32_t)split;
    } else {
    const uint32_t kf = S.key[S.perm[b]], kl = S.key[S.perm[e - 1u]];
    if (m > 2u && kf != kl) {
      const int prefix = __clz((int)(kf ^ kl));                 // the slice is sorted: find the first code that differs from kf in the top differing bit
      uint32_t lo = 0u, step = m;                               // largest lo with clz(kf ^ key[b + lo]) > prefix
      do {
        step = (step + 1u) >> 1;
        const uint32_t q = lo + step;
        if (q < m && __clz((int)(kf ^ S.key[S.perm[b + q]])) > prefix) lo = q;
      } while (step > 1u);
      mL = lo + 1u;
    }
    }
    const uint32_t mR = m - mL;
    const uint32_t jL = j + 1u, jR = j + mL;
    const uint32_t g = base + j;
    const uint32_t refL = mL == 1u ? leaf0 + S.perm[b] : base + jL;
    const uint32_t refR = mR == 1u ? leaf0 + S.perm[b + mL] : base + jR;
    t.left[g] = refL; t.right[g] = refR; t.parent[refL] = g; t.parent[refR] = g;
    return mL;
  }
}

template <int K, int SWEEP>
__global__ void __launch_bounds__(128)
k_treelet_build(B2 t, int n, const uint32_t* __restrict__ treeletStart, const uint32_t* __restrict__ sizeAt, uint32_t T,
                const uint64_t* __restrict__ keys) {
  extern __shared__ __align__(16) unsigned char tlSmemRaw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  TreeletSmem<K, SWEEP>& S = reinterpret_cast<TreeletSmem<K, SWEEP>*>(tlSmemRaw)[warp];
  const uint32_t tIdx = blockIdx.x * 4u + (uint32_t)warp;
  if (tIdx >= T) return;                                        // warps are independent: no block-wide barrier below
  const uint32_t a0 = treeletStart[tIdx], m0 = sizeAt[a0];
  if (m0 < 2u || m0 > (uint32_t)K) return;                      // a single triangle is its own cluster
  const uint32_t base = (T - 1u) + (a0 - tIdx);                 // global id of this treelet's local node 0 (its root)
  const uint32_t leaf0 = (uint32_t)(n - 1) + a0;                // global id of this treelet's first triangle (binary-tree leaf)
  for (uint32_t i = lane; i < m0; i += 32u) {
    const float4 lo = t.lo[leaf0 + i], hi = t.hi[leaf0 + i];
    S.box[0][i] = lo.x; S.box[1][i] = lo.y; S.box[2][i] = lo.z; S.box[3][i] = hi.x; S.box[4][i] = hi.y; S.box[5][i] = hi.z;
    S.perm[i] = (uint16_t)i;
    S.key[i] = (uint32_t)(keys[a0 + i] >> 32);
  }
  int sp = 0, nsmall = 0;                                       // warp-uniform
  if (m0 > (uint32_t)TL_SMALL) { if (lane == 0) { S.stack[0][0] = m0 << 16; S.stack[0][1] = 0u; } sp = 1; }
  else { if (lane == 0) S.small[0][0] = m0 << 10; nsmall = 1; }   // small-subtree entries: begin | end << 10 | local node index << 20
  __syncwarp();

  // ---------------- warp-cooperative phase: nodes of more than TL_SMALL triangles ----------------
  while (sp > 0) {
    --sp;
    const uint32_t be = S.stack[sp][0], j = S.stack[sp][1];
    const uint32_t b = be & 0xFFFFu, e = be >> 16, m = e - b;
    // centroid bounds (of lo + hi = twice the centroid, like the reference's PrimRef::center2)
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = b + lane; i < e; i += 32u) {
      const uint32_t p = S.perm[i];
      #pragma unroll
      for (int a = 0; a < 3; a++) { const float c = S.box[a][p] + S.box[3 + a][p]; cmin[a] = fminf(cmin[a], c); cmax[a] = fmaxf(cmax[a], c); }
    }
    #pragma unroll
    for (int a = 0; a < 3; a++) { cmin[a] = warpMin(cmin[a]); cmax[a] = warpMax(cmax[a]); }
    float scale[3];
    #pragma unroll
    for (int a = 0; a < 3; a++) scale[a] = cmax[a] > cmin[a] ? ((float)TL_BINS * 0.999f) / (cmax[a] - cmin[a]) : 0.f;
    // ---- bin (heuristic_binning.h:210-256) ----
    for (int w = lane; w < 3 * TL_BINS * 7; w += 32) {
      const int f = w % 7;
      (&S.bins[0][0][0])[w] = f < 3 ? TL_ORD_PINF : (f < 6 ? TL_ORD_NINF : 0u);
    }
    __syncwarp();
    // The triangles of a node are still in Morton order, so 32 consecutive ones fall into the same two or three bins and
    // their shared-memory atomics serialise 16-32 ways (round-2 history: 1.7 ms per tree level for 10 M triangles; grouping
    // equal bins with match.any + redux.sync was slower still -- the hardware runs one redux per group, profiles/r02d_ncu_treelet.txt).
    // Instead every lane walks its OWN contiguous share of the node, so that at any moment the 32 lanes touch triangles
    // from all over the node: the bins they hit are as spread out as the node's triangles are.
    {
      const uint32_t per = (m + 31u) >> 5;
      for (uint32_t it = 0; it < per; it++) {
        const uint32_t o = (uint32_t)lane * per + it;
        if (o < m) {
          const uint32_t p = S.perm[b + o];
          const float lx = S.box[0][p], ly = S.box[1][p], lz = S.box[2][p], hx = S.box[3][p], hy = S.box[4][p], hz = S.box[5][p];
          const uint32_t olx = f2ord(lx), oly = f2ord(ly), olz = f2ord(lz), ohx = f2ord(hx), ohy = f2ord(hy), ohz = f2ord(hz);
          const float c[3] = {lx + hx, ly + hy, lz + hz};
          #pragma unroll
          for (int a = 0; a < 3; a++) {
            const int bin = min(TL_BINS - 1, (int)((c[a] - cmin[a]) * scale[a]));
            uint32_t* B = S.bins[a][bin];
            atomicMin(&B[0], olx); atomicMin(&B[1], oly); atomicMin(&B[2], olz);
            atomicMax(&B[3], ohx); atomicMax(&B[4], ohy); atomicMax(&B[5], ohz);
            atomicAdd(&B[6], 1u);
          }
        }
      }
    }
    __syncwarp();
    // ---- best (heuristic_binning.h:336-392): lanes 0..15 accumulate bins 0..idx, lanes 16..31 bins 15..15-idx ----
    const int half = lane >> 4, idx = lane & 15;
    const int myBin = half ? (TL_BINS - 1 - idx) : idx;
    float bestCost = INFINITY; uint32_t bestKey = 0xFFFFFFFFu, bestLeft = 0u;
    #pragma unroll
    for (int a = 0; a < 3; a++) {
      const uint32_t* B = S.bins[a][myBin];
      float lo[3] = {ord2f(B[0]), ord2f(B[1]), ord2f(B[2])}, hi[3] = {ord2f(B[3]), ord2f(B[4]), ord2f(B[5])};
      uint32_t cnt = B[6];
      #pragma unroll
      for (int d = 1; d < TL_BINS; d <<= 1) {
        float vlo[3], vhi[3];
        #pragma unroll
        for (int c = 0; c < 3; c++) { vlo[c] = __shfl_up_sync(FULL, lo[c], d, 16); vhi[c] = __shfl_up_sync(FULL, hi[c], d, 16); }
        const uint32_t vc = __shfl_up_sync(FULL, cnt, d, 16);
        if (idx >= d) {
          #pragma unroll
          for (int c = 0; c < 3; c++) { lo[c] = fminf(lo[c], vlo[c]); hi[c] = fmaxf(hi[c], vhi[c]); }
          cnt += vc;
        }
      }
      const float area = cnt ? boxArea6(lo, hi) : 0.f;
      // plane idx (between bin idx and idx+1): left = lanes' own prefix, right = the mirrored prefix of lane 30 - idx
      const int src = (30 - idx) & 31;
      const float areaR = __shfl_sync(FULL, area, src);
      const uint32_t cntR = __shfl_sync(FULL, cnt, src);
      if (half == 0 && idx < TL_BINS - 1 && cnt != 0u && cntR != 0u) {
        const float cost = area * (float)cnt + areaR * (float)cntR;
        if (cost < bestCost) { bestCost = cost; bestKey = ((uint32_t)a << 8) | (uint32_t)idx; bestLeft = cnt; }
      }
    }
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float oc = __shfl_xor_sync(FULL, bestCost, o);
      const uint32_t ok = __shfl_xor_sync(FULL, bestKey, o), ol = __shfl_xor_sync(FULL, bestLeft, o);
      if (oc < bestCost || (oc == bestCost && ok < bestKey)) { bestCost = oc; bestKey = ok; bestLeft = ol; }
    }
    uint32_t mL;
    if (bestKey == 0xFFFFFFFFu) {
      mL = m >> 1;                                              // all centroids in one bin on every axis: object median, order kept
    } else {
      // ---- split (heuristic_binning_array_aligned.h:79-123): stable partition of the index slice through perm2 ----
      const int a = (int)(bestKey >> 8), pos = (int)(bestKey & 255u);
      mL = bestLeft;
      uint32_t offL = 0u, offR = 0u;
      for (uint32_t i0 = b; i0 < e; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const bool ok = i < e;
        uint32_t p = 0u; bool left = false;
        if (ok) {
          p = S.perm[i];
          const float c = S.box[a][p] + S.box[3 + a][p];
          left = min(TL_BINS - 1, (int)((c - cmin[a]) * scale[a])) <= pos;
        }
        const unsigned lm = __ballot_sync(FULL, ok && left), rm = __ballot_sync(FULL, ok && !left);
        const unsigned below = (1u << lane) - 1u;
        if (ok) {
          if (left) S.perm2[b + offL + __popc(lm & below)] = (uint16_t)p;
          else S.perm2[b + mL + offR + __popc(rm & below)] = (uint16_t)p;
        }
        offL += __popc(lm); offR += __popc(rm);
      }
      __syncwarp();
      for (uint32_t i = b + lane; i < e; i += 32u) S.perm[i] = S.perm2[i];
      __syncwarp();
    }
    const uint32_t mR = m - mL, jL = j + 1u, jR = j + mL;
    if (lane == 0) {
      const uint32_t g = base + j;
      const uint32_t refL = mL == 1u ? leaf0 + S.perm[b] : base + jL;
      const uint32_t refR = mR == 1u ? leaf0 + S.perm[b + mL] : base + jR;
      t.left[g] = refL; t.right[g] = refR; t.parent[refL] = g; t.parent[refR] = g;
    }
    // children: large ones back on the stack (the larger first, so the smaller is split next and the stack stays
    // logarithmic), small ones to the thread phase
    const uint32_t cb[2] = {b, b + mL}, ce[2] = {b + mL, e}, cj[2] = {jL, jR};
    const int first = mL >= mR ? 0 : 1;
    #pragma unroll
    for (int q = 0; q < 2; q++) {
      const int c = q == 0 ? first : 1 - first;
      const uint32_t cm = ce[c] - cb[c];
      if (cm < 2u) continue;
      // (every stacked range is at least as large as all ranges above it together, so at most log2(K / TL_SMALL) + 1 <= 6 are pending)
      if (cm > (uint32_t)TL_SMALL) { if (lane == 0) { S.stack[sp][0] = cb[c] | (ce[c] << 16); S.stack[sp][1] = cj[c]; } sp++; }
      else { if (lane == 0) (&S.small[0][0])[nsmall] = cb[c] | (ce[c] << 10) | (cj[c] << 20); nsmall++; }
    }
    __syncwarp();
  }

  // ---------------- thread phase: the small subtrees, level by level, one lane per NODE ----------------
  // (one lane per SUBTREE ran 2-5 lanes wide -- a 133-triangle treelet has 6-8 small subtrees of very different sizes -- and was a
  // quarter of this kernel, profiles/r02m_ncu_build.txt; the nodes of one level of all small subtrees are independent, so the warp
  // takes them 32 at a time from a queue in shared memory and pushes their children for the next round.)
  {
    uint32_t* qa = &S.small[0][0]; uint32_t* qb = qa + K / 2;   // ranges of >= 2 triangles are disjoint: at most K/2 per level
    int na = nsmall;
    const unsigned below = (1u << lane) - 1u;
    __syncwarp();
    while (na > 0) {
      int nb = 0;
      for (int i0 = 0; i0 < na; i0 += 32) {
        const int i = i0 + lane;
        uint32_t cL = 0u, cR = 0u;
        if (i < na) {
          const uint32_t w = qa[i];
          const uint32_t b = w & 1023u, e = (w >> 10) & 1023u, j = w >> 20;
          const uint32_t mL = treeletSmallNode<K, SWEEP>(S, t, b, e, j, base, leaf0, lane);
          if (mL >= 2u) cL = b | ((b + mL) << 10) | ((j + 1u) << 20);
          if (e - b - mL >= 2u) cR = (b + mL) | (e << 10) | ((j + mL) << 20);
        }
        const unsigned bl = __ballot_sync(FULL, cL != 0u), br = __ballot_sync(FULL, cR != 0u);
        if (cL) qb[nb + __popc(bl & below)] = cL;
        if (cR) qb[nb + __popc(bl) + __popc(br & below)] = cR;
        nb += __popc(bl) + __popc(br);
      }
      __syncwarp();
      uint32_t* x = qa; qa = qb; qb = x; na = nb;
    }
  }
}

// Bounds + SAH collapse programme inside the treelets, one WARP per treelet, level by level.
// The leaf-to-root climb with arrival counters (k_refit_dp) loses half of its lanes at every level and pays an atomic, two fences and an
// L2 round trip per node: 1.4 ms for the 10 M-triangle scene.  Doing the combine in the lane that built a small subtree (inside
// k_treelet_build) was no better: that phase runs 4.4 lanes wide at 20 warps per SM (profiles/r02m_ncu_build.txt).  Here the warp first
// lists the treelet's inner nodes in breadth-first order (shared memory only), then sweeps the levels bottom-up: all nodes of a level
// are independent, so the 32 lanes combine 32 nodes at a time and the L2 latency is paid once per level, not once per node.
constexpr int TDP_WARPS = 8;
template <int K>
__global__ void __launch_bounds__(TDP_WARPS * 32)
k_treelet_dp(B2 t, int n, const uint32_t* __restrict__ treeletStart, const uint32_t* __restrict__ sizeAt, uint32_t T,
             float costNode, float costTri, int maxLeafTris) {
  __shared__ uint32_t s_left[TDP_WARPS][K], s_right[TDP_WARPS][K];
  __shared__ uint16_t s_bfs[TDP_WARPS][K], s_lvl[TDP_WARPS][K];   // nodes in breadth-first order; first position of every level
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const uint32_t tIdx = blockIdx.x * TDP_WARPS + (uint32_t)warp;
  if (tIdx >= T) return;                                        // warps are independent: no block-wide barrier below
  const uint32_t a0 = treeletStart[tIdx], m0 = sizeAt[a0];
  if (m0 < 2u || m0 > (uint32_t)K) return;
  const uint32_t base = (T - 1u) + (a0 - tIdx);                 // global id of the treelet's root; its m0 - 1 inner nodes are base + [0, m0 - 1)
  const uint32_t firstLeaf = (uint32_t)(n - 1);
  uint32_t* L = s_left[warp]; uint32_t* R = s_right[warp]; uint16_t* bfs = s_bfs[warp]; uint16_t* lvl = s_lvl[warp];
  for (uint32_t i = lane; i < m0 - 1u; i += 32u) { L[i] = t.left[base + i]; R[i] = t.right[base + i]; }
  if (lane == 0) { bfs[0] = 0; lvl[0] = 0; }
  __syncwarp();
  // ---- breadth-first order ----
  uint32_t s = 0, e = 1, nl = 1;                                 // current level = bfs[s, e); nl levels known so far   (all warp-uniform)
  while (s < e) {
    uint32_t out = e;
    for (uint32_t i0 = s; i0 < e; i0 += 32u) {
      const uint32_t i = i0 + lane;
      uint32_t cl = RQ_INVALID, cr = RQ_INVALID;
      if (i < e) { const uint32_t x = bfs[i]; cl = L[x]; cr = R[x]; }
      const bool il = cl < firstLeaf, ir = cr < firstLeaf;       // inner children (of this treelet: ids base + ...)
      const unsigned bl = __ballot_sync(FULL, il), br = __ballot_sync(FULL, ir);
      const unsigned below = (1u << lane) - 1u;
      if (il) bfs[out + __popc(bl & below)] = (uint16_t)(cl - base);
      if (ir) bfs[out + __popc(bl) + __popc(br & below)] = (uint16_t)(cr - base);
      out += __popc(bl) + __popc(br);
    }
    __syncwarp();
    s = e; e = out;
    if (s < e) { if (lane == 0) lvl[nl] = (uint16_t)s; nl++; }
  }
  __syncwarp();
  // ---- bottom-up over the levels ----
  for (uint32_t l = nl; l-- > 0; ) {
    const uint32_t ls = lvl[l], le = (l + 1u < nl) ? (uint32_t)lvl[l + 1u] : (m0 - 1u);
    for (uint32_t i = ls + lane; i < le; i += 32u) {
      const uint32_t x = bfs[i];
      combineNode(t, base + x, L[x], R[x], costNode, costTri, maxLeafTris);
    }
    __threadfence_block();
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
k_leaf_init(B2 t, int n, const RQTri* __restrict__ trisIn, const uint32_t* __restrict__ vals, float costTri) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  initLeaf(t, (uint32_t)(n - 1 + k), trisIn, vals[k], costTri);
}

// ----------------------------------------------------------------------------------------------
// 6. top-down emission of 8-wide nodes, one tree level per launch.
// ----------------------------------------------------------------------------------------------
struct EmitCounters {
  uint32_t nodeCount;          // 8-wide nodes allocated so far
  uint32_t triCount;           // leaf triangles allocated so far
  uint32_t nextCount;          // entries written to the next-level queue
  uint32_t leafSlots;
  double sahInnerQ, sahLeafQ;  // reference SAH terms on de-quantised boxes
  double sahInnerX, sahLeafX;  // same on exact boxes
  double sahLeafTrisQ;         // leaf term weighted by triangles instead of blocks: sum A(leaf) * numTris (de-quantised boxes)
};

__device__ __forceinline__ uint8_t expForExtent(float ext) {
  // smallest biased exponent eb with ext <= 255 * 2^(eb-127); 2^(eb-127) must be a normal float
  if (!(ext > 0.f)) return 1;
  int ex; frexpf(ext * (1.0f / 255.0f), &ex);                  // ext/255 = m * 2^ex, m in [0.5,1)
  int eb = ex + 127;                                           // 2^ex >= ext/255
  if (eb < 1) eb = 1;
  if (eb > 254) eb = 254;
  return (uint8_t)eb;
}

// One tree level per launch, but the host does not wait for it: the level's entry count lives on the device (`lv`, written by
// k_emit_advance after the previous level), the grid is a fixed persistent one (grid-stride over the queue), and the host reads the
// per-level counts back once per batch of levels.  (One launch + read-back + synchronise per level cost 12 round trips = 0.3 ms of
// the 10 M-triangle build.)
struct EmitLevel { uint32_t count, nodeEnd; };                // queue entries of a level; nodes allocated once that level has been emitted

// EIGHT LANES PER 8-WIDE NODE.  Round-2 history (profiles/r02m_ncu_build.txt): with one thread per node everything below lived in
// dynamically indexed local arrays (880 bytes of stack per thread; at 1024 threads per SM that is 0.9 MB per SM and spills to DRAM:
// 0.73 GB written by a launch whose nodes and triangles are 0.24 GB), the greedy slot assignment alone was ~2500 instructions per
// thread, and four counters in one 32-byte sector took an atomic per node.  Now lane k of a group first holds the k-th budget slot of
// the collapse programme (the <= 8 children are found by splitting budget intervals in parallel), then its child: box, scores, slot,
// quantised bytes all stay in registers; the 128-byte node is assembled in shared memory and leaves as one coalesced line; the
// allocation counters are bumped once per BLOCK (32 nodes).
constexpr int EMIT_THREADS = 256, EMIT_GROUPS = EMIT_THREADS / 8;

__global__ void __launch_bounds__(EMIT_THREADS)
k_emit(B2 t, int n, const uint2* __restrict__ queue, const EmitLevel* __restrict__ lv, uint2* __restrict__ nextQueue,
       EmitCounters* ctr, RQNode* __restrict__ nodes, RQTri* __restrict__ trisOut,
       const RQTri* __restrict__ trisIn, const uint32_t* __restrict__ vals, uint32_t level,
       const uint32_t* __restrict__ parentOf /* wide parent per queue entry */, uint32_t* __restrict__ nextParentOf,
       const uint32_t* __restrict__ idxIn, RQTriC* __restrict__ trisC, uint32_t* __restrict__ metaOut) {
  const uint32_t count = lv[level].count;
  if (count == 0u) return;                                      // a level launched past the bottom of the tree
  static_assert(EMIT_GROUPS == 32, "the block-level scan below uses one warp, one lane per group");
  __shared__ uint2 s_x[EMIT_GROUPS][8];                         // gather: inbox of every budget slot
  __shared__ uint32_t s_cnt[3][EMIT_GROUPS], s_base[3][EMIT_GROUPS];
  __shared__ __align__(16) uint32_t s_node[EMIT_GROUPS][32];    // the node records being assembled
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u, sub = threadIdx.x & 7u, grp = threadIdx.x >> 3, gsh = lane & 24u;
  const uint32_t firstLeaf = (uint32_t)(n - 1);
  double sInnerQ = 0.0, sLeafQ = 0.0, sInnerX = 0.0, sLeafX = 0.0, sLeafTrisQ = 0.0;

  for (uint32_t q0 = blockIdx.x * EMIT_GROUPS; q0 < count; q0 += gridDim.x * EMIT_GROUPS) {   // block-uniform trip count
    const uint32_t q = q0 + grp;
    const bool valid = q < count;
    uint32_t b = 0u, w = 0u;
    if (valid) { const uint2 e = queue[q]; b = e.x; w = e.y; }

    // ---- gather the <= 8 children by replaying the DP decisions: the item whose budget interval starts at k lives in lane k ----
    uint32_t m = RQ_INVALID; int bud = 0, state = 0;             // state: 0 = empty, 1 = pending, 2 = child found
    bool isInner = false, force = false;
    if (valid && sub == 0u) { m = b; bud = 8; state = 1; force = b < firstLeaf; }   // the node itself always splits 8 ways (b >= firstLeaf: a single-triangle scene)
    for (int round = 0; round < 16; round++) {
      s_x[grp][sub] = make_uint2(RQ_INVALID, 0u);
      __syncwarp();
      if (state == 1) {
        if (m >= firstLeaf) { state = 2; isInner = false; }
        else {
          const uint32_t dec = t.dec[m];
          int i = bud;
          if (!force) while (i > 1 && ((dec >> (3 * (i - 1))) & 7u) == 0u) i--;
          if (i == 1) { state = 2; isInner = (dec & 1u) != 0u; }
          else {
            const int kk = (int)((dec >> (3 * (i - 1))) & 7u);
            s_x[grp][sub + kk] = make_uint2(t.right[m], (uint32_t)(i - kk));
            m = t.left[m]; bud = kk;
          }
        }
        force = false;
      }
      __syncwarp();
      const uint2 in = s_x[grp][sub];
      if (in.x != RQ_INVALID) { m = in.x; bud = (int)in.y; state = 1; }
      if (!__any_sync(FULL, state == 1)) break;
    }
    const bool has = state == 2;
    const bool leaf = has && !isInner;
    const unsigned hb = __ballot_sync(FULL, has), ib = __ballot_sync(FULL, has && isInner), lb = __ballot_sync(FULL, leaf);
    const int nc = __popc((hb >> gsh) & 0xFFu);

    // ---- node box and this lane's child box ----
    float4 nlo = make_float4(0.f, 0.f, 0.f, 0.f), nhi = nlo, lo = nlo, hi = nlo;
    if (valid) { nlo = t.lo[b]; nhi = t.hi[b]; }
    if (has) { lo = t.lo[m]; hi = t.hi[m]; }
    const uint32_t ctris = has ? __float_as_uint(hi.w) : 0u;

    // ---- slot assignment: slot s (bit a set = "towards +axis a") takes the child whose centre lies furthest in that diagonal
    //      direction; greedy maximum over the 8x8 score table (ties: first child, then first slot) ----
    int mySlot = -1;
    {
      const float ncx = 0.5f * (nlo.x + nhi.x), ncy = 0.5f * (nlo.y + nhi.y), ncz = 0.5f * (nlo.z + nhi.z);
      const float dx = 0.5f * (lo.x + hi.x) - ncx, dy = 0.5f * (lo.y + hi.y) - ncy, dz = 0.5f * (lo.z + hi.z) - ncz;
      float sc[8];
      #pragma unroll
      for (int sl = 0; sl < 8; sl++) sc[sl] = ((sl & 1) ? dx : -dx) + ((sl & 2) ? dy : -dy) + ((sl & 4) ? dz : -dz);
      // Every round: the largest score still available (a plain max: scores of taken slots and of placed children are -FLT_MAX),
      // then the first child (lowest lane) and, within it, the first slot that attain it -- what the sequential double loop with
      // its strict comparison picks.  ~40 instructions per round instead of ~110 for an arg-max carried through the shuffles.
      uint32_t freeS = 0xFFu;
      if (!has) {
        #pragma unroll
        for (int sl = 0; sl < 8; sl++) sc[sl] = -FLT_MAX;
      }
      for (int r = 0; r < 8; r++) {
        if (__all_sync(FULL, r >= nc)) break;
        float best = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), fmaxf(fmaxf(sc[4], sc[5]), fmaxf(sc[6], sc[7])));
        #pragma unroll
        for (int o = 1; o < 8; o <<= 1) best = fmaxf(best, __shfl_xor_sync(FULL, best, o));
        uint32_t eq = 0u;
        #pragma unroll
        for (int sl = 0; sl < 8; sl++) eq |= (sc[sl] == best) ? (1u << sl) : 0u;
        const bool none = !(best > -FLT_MAX);                   // no comparable score left (NaN boxes cannot occur; be safe): first child, first slot
        if (none) eq = (has && mySlot < 0) ? freeS : 0u;
        const unsigned cand = (__ballot_sync(FULL, eq != 0u) >> gsh) & 0xFFu;
        const int kl = __ffs(cand) - 1;                          // group-uniform; >= 0 while r < nc
        const int ks = __shfl_sync(FULL, __ffs(eq) - 1, (int)gsh + (kl < 0 ? 0 : kl));
        if (r < nc && kl >= 0) {
          const bool mine = (int)sub == kl;
          if (mine) mySlot = ks;
          freeS &= ~(1u << ks);
          #pragma unroll
          for (int sl = 0; sl < 8; sl++) if (mine || sl == ks) sc[sl] = -FLT_MAX;
        }
      }
    }

    // ---- allocate children / triangles: one atomic per counter per block ----
    const uint32_t numInner = __popc((ib >> gsh) & 0xFFu), numLeaves = __popc((lb >> gsh) & 0xFFu);
    uint32_t numLeafTris = leaf ? ctris : 0u;
    #pragma unroll
    for (int o = 1; o < 8; o <<= 1) numLeafTris += __shfl_xor_sync(FULL, numLeafTris, o);
    if (sub == 0u) { s_cnt[0][grp] = numInner; s_cnt[1][grp] = numLeafTris; s_cnt[2][grp] = numLeaves; }
    __syncthreads();
    if (threadIdx.x < 32u) {
      const uint32_t a = s_cnt[0][lane], c = s_cnt[1][lane], d = s_cnt[2][lane];
      uint32_t si = a, st = c, sl = d;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(FULL, si, o), y = __shfl_up_sync(FULL, st, o), z = __shfl_up_sync(FULL, sl, o);
        if ((int)lane >= o) { si += x; st += y; sl += z; }
      }
      uint32_t bi = 0u, bt = 0u, bq = 0u;
      if (lane == 31u) {
        if (si) { bi = atomicAdd(&ctr->nodeCount, si); bq = atomicAdd(&ctr->nextCount, si); }
        if (st) bt = atomicAdd(&ctr->triCount, st);
        if (sl) atomicAdd(&ctr->leafSlots, sl);
      }
      bi = __shfl_sync(FULL, bi, 31); bt = __shfl_sync(FULL, bt, 31); bq = __shfl_sync(FULL, bq, 31);
      s_base[0][lane] = bi + si - a; s_base[1][lane] = bt + st - c; s_base[2][lane] = bq + si - a;
    }
    __syncthreads();
    const uint32_t childBase = numInner ? s_base[0][grp] : 0u, triBase = numLeafTris ? s_base[1][grp] : 0u, qBase = s_base[2][grp];

    // ---- rank of this child among the inner children / first triangle of its leaf slot, both in SLOT order ----
    uint32_t innerRank = 0u, triOff = 0u;
    {
      const uint32_t packed = has ? ((uint32_t)mySlot | (isInner ? 0x10u : 0u) | (ctris << 8)) : 0xFu;
      #pragma unroll
      for (int k = 0; k < 8; k++) {
        const uint32_t pk = __shfl_sync(FULL, packed, (int)gsh + k);
        if (has && (pk & 0xFu) < (uint32_t)mySlot) { if (pk & 0x10u) innerRank++; else triOff += pk >> 8; }
      }
    }

    // ---- quantisation grid of the node (every lane of the group computes the same) and this child's bytes ----
    const float np[3] = {nlo.x, nlo.y, nlo.z};
    const float ext[3] = {__fsub_ru(nhi.x, nlo.x), __fsub_ru(nhi.y, nlo.y), __fsub_ru(nhi.z, nlo.z)};
    const float clo[3] = {lo.x, lo.y, lo.z}, chi[3] = {hi.x, hi.y, hi.z};
    uint32_t ebits = 0u, qlo[3] = {0u, 0u, 0u}, qhi[3] = {0u, 0u, 0u};
    float dq[3];
    #pragma unroll
    for (int a = 0; a < 3; a++) {
      uint8_t eb = expForExtent(ext[a]);
      // make sure the far face is representable: ceil(ext / 2^e) <= 255
      while (eb < 254 && __fmul_ru(ext[a], __uint_as_float((uint32_t)(254 - eb) << 23)) > 255.0f) eb++;
      ebits |= (uint32_t)eb << (8 * a);
      const float step = __uint_as_float((uint32_t)eb << 23);            // 2^(eb-127)
      const float inv = __uint_as_float((uint32_t)(254 - eb) << 23);      // 2^(127-eb)
      // floor / ceil with directed rounding: decoded box always contains the exact one
      float fl = floorf(__fmul_rd(__fsub_rd(clo[a], np[a]), inv));
      float fh = ceilf(__fmul_ru(__fsub_ru(chi[a], np[a]), inv));
      fl = fminf(fmaxf(fl, 0.f), 255.f); fh = fminf(fmaxf(fh, 0.f), 255.f);
      qlo[a] = (uint32_t)fl; qhi[a] = (uint32_t)fh;
      dq[a] = (fh - fl) * step;
    }
    uint32_t masks = 0u;
    if (has) {
      const double Aq = (double)halfArea(dq[0], dq[1], dq[2]);
      const double Ax = (double)halfArea(chi[0] - clo[0], chi[1] - clo[1], chi[2] - clo[2]);
      if (isInner) {
        masks = 1u << (24 + mySlot);
        sInnerQ += Aq; sInnerX += Ax;
      } else {
        masks = ((1u << ctris) - 1u) << (3 * mySlot);          // 1..3 triangles, stored in slot order
        const double blocks = (double)((ctris + 3u) / 4u);
        sLeafQ += Aq * blocks; sLeafX += Ax * blocks; sLeafTrisQ += Aq * (double)ctris;
      }
    }
    #pragma unroll
    for (int o = 1; o < 8; o <<= 1) masks |= __shfl_xor_sync(FULL, masks, o);
    if (level == 0u && valid && sub == 0u) { sInnerQ += (double)nlo.w; sInnerX += (double)nlo.w; }   // the root's own box

    // ---- assemble the 128-byte record in shared memory: lane k writes words 4k .. 4k+3, then the children drop their bytes in ----
    {
      uint4 v;
      if (sub == 0u) v = make_uint4(__float_as_uint(nlo.x), __float_as_uint(nlo.y), __float_as_uint(nlo.z), ebits);
      else if (sub == 1u) v = make_uint4(childBase, triBase, masks, 0u);
      else if (sub == 2u) v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);                 // qlo of empty slots = 255 ...
      else if (sub == 3u) v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);                                     // ... qhi = 0
      else if (sub == 4u) v = make_uint4(0u, 0u, 0u, 0u);
      else if (sub == 5u) v = make_uint4(__float_as_uint(nlo.x), __float_as_uint(nlo.y), __float_as_uint(nlo.z), __float_as_uint(nhi.x));
      else if (sub == 6u) v = make_uint4(__float_as_uint(nhi.y), __float_as_uint(nhi.z), (parentOf && valid) ? parentOf[q] : RQ_INVALID, __float_as_uint(nhi.w));
      else v = make_uint4(level, 0u, 0u, 0u);
      ((uint4*)s_node[grp])[sub] = v;
      __syncwarp();
      if (has) {
        uint8_t* nb = (uint8_t*)s_node[grp];
        #pragma unroll
        for (int a = 0; a < 3; a++) { nb[32 + a * 8 + mySlot] = (uint8_t)qlo[a]; nb[56 + a * 8 + mySlot] = (uint8_t)qhi[a]; }
      }
      __syncwarp();
      if (valid) ((uint4*)(nodes + w))[sub] = ((const uint4*)s_node[grp])[sub];
    }

    // ---- inner child: an entry of the next level's queue ----
    if (has && isInner) {
      nextQueue[qBase + innerRank] = make_uint2(m, childBase + innerRank);
      nextParentOf[qBase + innerRank] = w;
    }
    // ---- leaf slot: its <= 3 triangles = the leaves below binary node m (tiny depth-first walk; works for the radix tree and for
    //      PLOC, whose subtrees are not contiguous in Morton order) ----
    if (leaf) {
      uint32_t walk[4]; int wsp = 0; uint32_t j = 0u;
      walk[wsp++] = m;
      while (wsp > 0 && j < ctris) {
        const uint32_t x = walk[--wsp];
        if (x >= firstLeaf) {
          const uint32_t tin = vals[x - firstLeaf];
          const float4* src = (const float4*)(trisIn + tin);
          const uint32_t dsti = triBase + triOff + j;
          if (trisC) {                                          // compact layout: indices + primID, geomID / quad flag on the side
            const float4 c = src[2];
            RQTriC r; r.v0 = idxIn[3 * (size_t)tin]; r.v1 = idxIn[3 * (size_t)tin + 1]; r.v2 = idxIn[3 * (size_t)tin + 2]; r.primID = __float_as_uint(c.y);
            trisC[dsti] = r;
            metaOut[dsti] = (__float_as_uint(c.z) & ~RQ_META_FLIPUV) | ((__float_as_uint(c.w) & RQ_PAD_FLIPUV) ? RQ_META_FLIPUV : 0u);
          } else {
            float4* dst = (float4*)(trisOut + dsti);
            const float4 a0 = src[0], a1 = src[1], a2 = src[2];
            if (dsti & 1u) { dst[0] = a2; dst[1] = a0; dst[2] = a1; }   // odd records: last 16 bytes first (rq_types.h)
            else { dst[0] = a0; dst[1] = a1; dst[2] = a2; }
          }
          j++;
        } else if (wsp < 3) { walk[wsp++] = t.right[x]; walk[wsp++] = t.left[x]; }
      }
    }
  }

  double s[5] = {sInnerQ, sLeafQ, sInnerX, sLeafX, sLeafTrisQ};
  #pragma unroll
  for (int k = 0; k < 5; k++)
    for (int o = 16; o; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ctr->sahInnerQ, s[0]); atomicAdd(&ctr->sahLeafQ, s[1]);
    atomicAdd(&ctr->sahInnerX, s[2]); atomicAdd(&ctr->sahLeafX, s[3]); atomicAdd(&ctr->sahLeafTrisQ, s[4]);
  }
}

// between two levels: the counts of the level just emitted become the next level's work description
__global__ void k_emit_advance(EmitCounters* ctr, EmitLevel* lv, uint32_t level) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    lv[level].nodeEnd = ctr->nodeCount;
    lv[level + 1].count = ctr->nextCount;
    ctr->nextCount = 0u;
  }
}

__global__ void k_empty_root(RQNode* nodes) {                 // scene without valid triangles
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    RQNode N;
    memset(&N, 0, sizeof(N));
    N.e[0] = N.e[1] = N.e[2] = 1; N.parent = RQ_INVALID;
    for (int s = 0; s < 8; s++) for (int a = 0; a < 3; a++) { N.qlo[a][s] = 255; N.qhi[a][s] = 0; }
    nodes[0] = N;
  }
}

// Build scratch comes from the device's stream-ordered pool (cudaMallocAsync): after the first
// commit the pool hands the same pages back, so a rebuild pays no cudaMalloc/cudaFree (which cost
// more than the build kernels: 13 of them took 5-6 ms of a 9 ms commit of 1 M triangles).
struct ScratchScope {
  cudaStream_t stream;
  explicit ScratchScope(cudaStream_t s) : stream(s) {
    int dev = 0; cudaMemPool_t pool = nullptr;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long thr = 0;
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      // a 50 M-triangle build needs ~22 GB of scratch: with a 16 GB threshold the pool gave pages back at every synchronisation
      // and the next phase re-allocated them (80 ms spent in the hierarchy phase, profiles/r01s_build_c5.jsonl)
      if (thr < (64ull << 30)) { thr = 64ull << 30; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    }
    cudaGetLastError();
  }
};
static thread_local cudaStream_t t_scratchStream = nullptr;
static thread_local rqAllocMonitorFn t_monFn = nullptr;
static thread_local void* t_monUser = nullptr;
inline bool monitorAlloc(size_t bytes) { return !t_monFn || t_monFn(t_monUser, (long long)bytes, false); }
inline void monitorFree(size_t bytes) { if (t_monFn) t_monFn(t_monUser, -(long long)bytes, true); }
template <typename T>
struct DevBuf {
  T* p = nullptr; cudaStream_t s = nullptr; size_t bytes = 0;
  cudaError_t alloc(size_t n) {
    s = t_scratchStream;
    const size_t want = (n ? n : 1) * sizeof(T);
    if (!monitorAlloc(want)) return cudaErrorMemoryAllocation;   // vetoed by the application's memory monitor
    const cudaError_t e = cudaMallocAsync((void**)&p, want, s);
    if (e == cudaSuccess) bytes = want; else { p = nullptr; monitorFree(want); }
    return e;
  }
  ~DevBuf() { if (p) { cudaFreeAsync(p, s); monitorFree(bytes); } }
  void swap(DevBuf& o) { std::swap(p, o.p); std::swap(s, o.s); std::swap(bytes, o.bytes); }
};

inline unsigned blocksFor(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void rqSetAllocMonitor(rqAllocMonitorFn fn, void* user) { t_monFn = fn; t_monUser = user; }

// Images come from the stream-ordered pool as well (a re-commit then reuses the pages of the image it replaces instead of paying
// cudaMalloc / cudaFree: 5 ms of a 10 ms commit of 1 M triangles, profiles/r01t_bench.json).  Freeing keeps cudaFree's guarantee --
// nothing on the device still reads the image -- by synchronising the device first.
void rqFreeImage(RQDeviceImage* img) {
  if (img && img->base) {
    cudaDeviceSynchronize();
    if (cudaFreeAsync(img->base, cudaStreamPerThread) != cudaSuccess) { cudaGetLastError(); cudaFree(img->base); }
    img->base = nullptr;
  }
}
int rqAllocImage(void** p, size_t bytes, rqStream stream) { return (int)cudaMallocAsync(p, bytes, (cudaStream_t)stream); }

int rqBuildBVH(const RQGeomDesc* geoms, int numGeoms, uint32_t sceneFlags, const RQBuildParams* params,
               rqStream stream_, RQDeviceImage* out, RQBuildStats* stats) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ScratchScope scratch(stream);
  t_scratchStream = stream;
  int err = 0;
  RQBuildParams P = {1.0f, 1.0f, 3, 0, 2, 8, 1, 256, 0, 0};
  if (params) P = *params;
  if (P.maxLeafTris < 1) P.maxLeafTris = 1;
  if (P.maxLeafTris > 3) P.maxLeafTris = 3;

  uint64_t total = 0;
  std::vector<RQGeomDesc> hg(geoms, geoms + numGeoms);
  uint64_t totalVerts = 0; bool hasInstances = false;
  for (auto& g : hg) {
    g.primBase = (uint32_t)total; total += g.numTris;
    g.vertBase = (uint32_t)totalVerts;
    if (g.type == 1u) hasInstances = true; else totalVerts += g.numVerts;
  }
  if (total >= 0x7FFFFFF0ull || totalVerts >= 0xFFFFFFF0ull) return (int)cudaErrorInvalidValue;
  uint32_t N = (uint32_t)total;                                 // primitive references (grows when large triangles are pre-split)
  const uint32_t Nin = N;
  // RTC_SCENE_FLAG_COMPACT (= 2): indexed leaves + a vertex pool inside the image.  Instance primitives have no room in a
  // 16-byte record, so a scene with instances keeps the 48-byte layout (the flag is a memory hint, as in the reference).
  const bool compact = (sceneFlags & 2u) != 0u && !hasInstances;
  const uint32_t numVerts = (compact && total > 0) ? (uint32_t)totalVerts : 0u;

  cudaEvent_t ev[7]; for (auto& e : ev) e = nullptr;
  DevBuf<RQGeomDesc> dGeoms; DevBuf<RQTri> trisIn, trisOut; DevBuf<uint64_t> keys0, keys1;
  DevBuf<uint32_t> vals0, vals1, hist, digitTotal, left, right, parent, rangeFirst, flag, dec, qParent0, qParent1;
  DevBuf<uint32_t> cid0, cid1, nnBuf, blockCount, plocCtr; uint32_t plocIters = 0;
  DevBuf<uint32_t> rparent, rfirst, rlast, sizeAt, tlStart, tlBlock, tlTotal; uint32_t numTreelets = 0;
  DevBuf<float4> blo, bhi; DevBuf<float> cost; DevBuf<uint2> queue0, queue1; DevBuf<RQNode> nodes;
  DevBuf<Bounds12> dBounds; DevBuf<uint32_t> dInvalid; DevBuf<EmitCounters> dCtr; DevBuf<EmitLevel> dLevels;
  DevBuf<uint32_t> idx3, metaOut; DevBuf<RQTriC> trisC; DevBuf<float4> vpool;
  DevBuf<uint32_t> pieces, tileSum, preTotal, idxSplit; DevBuf<RQTri> trisSplit; DevBuf<float4> refLo, refHi;
  bool useRefBoxes = false; uint32_t numSplitRefs = 0;
  Bounds12 hb; uint32_t hInvalid = 0; EmitCounters hc;
  uint32_t n = 0, depth = 0, numNodes = 1, numTris = 0;
  std::vector<uint32_t> levelEnd(1, 1u);                        // level 0 = the root = node range [0,1)
  RQImageHeader H;
  void* image = nullptr; size_t imageAccounted = 0;
  memset(&hc, 0, sizeof(hc));
  memset(&H, 0, sizeof(H));

  for (auto& e : ev) CK(cudaEventCreate(&e));
  CK(cudaEventRecord(ev[0], stream));
  CK(dBounds.alloc(1)); CK(dInvalid.alloc(1)); CK(dCtr.alloc(1));
  for (int k = 0; k < 3; k++) { hb.sceneLo[k] = hb.centLo[k] = 0xFFFFFFFFu; hb.sceneHi[k] = hb.centHi[k] = 0u; }
  hb.extSum = 0.f; hb.numValid = 0u;
  CK(cudaMemcpyAsync(dBounds.p, &hb, sizeof(hb), cudaMemcpyHostToDevice, stream));
  CK(cudaMemsetAsync(dInvalid.p, 0, 4, stream));

  if (N > 0) {
    CK(dGeoms.alloc(numGeoms));
    CK(cudaMemcpyAsync(dGeoms.p, hg.data(), sizeof(RQGeomDesc) * numGeoms, cudaMemcpyHostToDevice, stream));
    CK(trisIn.alloc(N));
    if (compact) CK(idx3.alloc(3 * (size_t)N));
    k_setup_prims<<<blocksFor(N, 256), 256, 0, stream>>>(dGeoms.p, numGeoms, N, trisIn.p, dBounds.p, dInvalid.p, compact ? idx3.p : nullptr);
    if (P.presplit) {
      // ---- RTC_BUILD_QUALITY_HIGH: large triangles become several references with clipped boxes (section 1b) ----
      const uint32_t tiles = blocksFor(N, TLS_THREADS_PRE * 4);
      CK(pieces.alloc(N)); CK(tileSum.alloc(tiles + 1)); CK(preTotal.alloc(1));
      k_presplit_count<<<blocksFor(N, 256), 256, 0, stream>>>(trisIn.p, N, dBounds.p, pieces.p);
      k_tile_sums<<<tiles, TLS_THREADS_PRE, 0, stream>>>(pieces.p, N, tileSum.p);
      k_treelet_scan<<<1, 1024, 0, stream>>>(tileSum.p, tiles, preTotal.p);
      rqCountLaunch(3);
      uint32_t N2 = 0;
      CK(cudaMemcpyAsync(&N2, preTotal.p, 4, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      CK(cudaGetLastError());
      if (N2 < N || (uint64_t)N2 >= 0x7FFFFFF0ull) { err = (int)cudaErrorUnknown; goto fail; }
      CK(refLo.alloc(N2)); CK(refHi.alloc(N2));
      if (N2 > N) {
        CK(trisSplit.alloc(N2));
        if (compact) CK(idxSplit.alloc(3 * (size_t)N2));
        k_presplit_write<<<tiles, TLS_THREADS_PRE, 0, stream>>>(trisIn.p, N, dBounds.p, pieces.p, tileSum.p, trisSplit.p, refLo.p, refHi.p,
                                                                compact ? idx3.p : nullptr, compact ? idxSplit.p : nullptr);
        rqCountLaunch(1);
        CK(cudaGetLastError());
        trisIn.swap(trisSplit);
        if (compact) idx3.swap(idxSplit);
        numSplitRefs = N2 - N;
        N = N2;
        useRefBoxes = true;
      }
    }
    CK(keys0.alloc(N)); CK(keys1.alloc(N)); CK(vals0.alloc(N)); CK(vals1.alloc(N));
    // The treelet builder only needs the Morton order down to cells of a few hundred triangles (everything below is rebuilt by SAH):
    // it sorts on the upper 32 bits of the code (10-11 bits per axis), i.e. half the radix passes; ties keep input order.
    k_morton<<<blocksFor(N, 256), 256, 0, stream>>>(trisIn.p, N, dBounds.p, keys0.p, vals0.p, P.mortonCubic,
                                                    P.builder == 2 ? 0xFFFFFFFF00000000ull : ~0ull,
                                                    useRefBoxes ? refLo.p : nullptr, useRefBoxes ? refHi.p : nullptr);
    rqCountLaunch(2);
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ev[1], stream));

  if (N > 0) {                                                  // ---- radix sort ----
    const uint32_t numTiles = blocksFor(N, SORT_TILE);
    CK(hist.alloc((size_t)256 * numTiles)); CK(digitTotal.alloc(256));
    uint64_t *kin = keys0.p, *kout = keys1.p; uint32_t *vin = vals0.p, *vout = vals1.p;
    for (int pass = (P.builder == 2 ? 4 : 0); pass < 8; pass++) {
      const int shift = pass * 8;
      k_sort_hist<<<numTiles, SORT_THREADS, 0, stream>>>(kin, N, shift, numTiles, hist.p);
      k_sort_scan_rows<<<256, 256, 0, stream>>>(hist.p, numTiles, digitTotal.p);
      k_sort_scan_digits<<<1, 256, 0, stream>>>(digitTotal.p);
      k_sort_scatter<<<numTiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, N, shift, numTiles, hist.p, digitTotal.p);
      rqCountLaunch(4);
      std::swap(kin, kout); std::swap(vin, vout);
    }
    CK(cudaGetLastError());
    // 8 (or 4) passes: result is back in keys0/vals0
  }
  CK(cudaEventRecord(ev[2], stream));
  CK(cudaMemcpyAsync(&hInvalid, dInvalid.p, 4, cudaMemcpyDeviceToHost, stream));
  CK(cudaMemcpyAsync(&hb, dBounds.p, sizeof(hb), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  n = N - hInvalid;

  {                                                             // ---- hierarchy + refit ----
    B2 t; memset(&t, 0, sizeof(t));
    const size_t n2 = n ? 2 * (size_t)n - 1 : 1;
    CK(nodes.alloc(n ? (size_t)n : 1));
    if (n > 0) {
      CK(blo.alloc(n2)); CK(bhi.alloc(n2)); CK(cost.alloc(n2 * 8)); CK(dec.alloc(n2));
      CK(left.alloc(n)); CK(right.alloc(n)); CK(parent.alloc(n2)); CK(rangeFirst.alloc(n)); CK(flag.alloc(n));
      if (compact) { CK(trisC.alloc(n)); CK(metaOut.alloc(n)); } else CK(trisOut.alloc(n));
      CK(queue0.alloc(n)); CK(queue1.alloc(n)); CK(qParent0.alloc(n)); CK(qParent1.alloc(n));
      t.lo = blo.p; t.hi = bhi.p; t.cost = cost.p; t.dec = dec.p; t.left = left.p; t.right = right.p;
      t.parent = parent.p; t.rangeFirst = rangeFirst.p; t.flag = flag.p;
      t.refLo = useRefBoxes ? refLo.p : nullptr; t.refHi = useRefBoxes ? refHi.p : nullptr;
      CK(cudaMemsetAsync(flag.p, 0, sizeof(uint32_t) * n, stream));
      CK(cudaMemsetAsync(parent.p, 0xFF, sizeof(uint32_t) * n2, stream));
      if ((P.builder == 1 || P.builder == 2) && n > 1) {
        const int radius = P.plocRadius < 1 ? 1 : (P.plocRadius > PLOC_MAX_RADIUS ? PLOC_MAX_RADIUS : P.plocRadius);
        CK(cid0.alloc(n)); CK(cid1.alloc(n)); CK(nnBuf.alloc(n)); CK(blockCount.alloc(blocksFor(n, PLOC_THREADS) + 1)); CK(plocCtr.alloc(5));
        uint32_t m0 = n;                                          // clusters PLOC starts from
        if (P.builder == 2) {
          // ---- binned-SAH treelets: radix tree -> treelet cut -> one warp per treelet -> bottom-up bounds + collapse programme ----
          const uint32_t K = P.treeletSize >= 512 ? 512u : 256u;
          const uint32_t numTiles = blocksFor(n, TLS_TILE);
          CK(rparent.alloc(n2)); CK(rfirst.alloc(n)); CK(rlast.alloc(n)); CK(sizeAt.alloc(n)); CK(tlStart.alloc(n)); CK(tlBlock.alloc(numTiles + 1)); CK(tlTotal.alloc(1));
          CK(cudaMemsetAsync(rparent.p, 0xFF, sizeof(uint32_t) * n2, stream));
          CK(cudaMemsetAsync(sizeAt.p, 0, sizeof(uint32_t) * n, stream));
          // the radix tree is only consulted for its ranges; its child arrays land in buffers PLOC overwrites later
          k_hierarchy<<<blocksFor(n - 1, 256), 256, 0, stream>>>(keys0.p, (int)n, cid0.p, cid1.p, rparent.p, rfirst.p, rlast.p);
          k_treelet_mark<<<blocksFor(2 * (size_t)n - 1, 256), 256, 0, stream>>>((int)n, rparent.p, rfirst.p, rlast.p, K, sizeAt.p);
          k_treelet_count<<<numTiles, TLS_THREADS, 0, stream>>>(sizeAt.p, n, tlBlock.p);
          k_treelet_scan<<<1, 1024, 0, stream>>>(tlBlock.p, numTiles, tlTotal.p);
          k_treelet_write<<<numTiles, TLS_THREADS, 0, stream>>>(sizeAt.p, n, tlBlock.p, tlStart.p);
          k_leaf_init<<<blocksFor(n, 256), 256, 0, stream>>>(t, (int)n, trisIn.p, vals0.p, P.costTri);
          rqCountLaunch(6);
          uint32_t T = 0;
          CK(cudaMemcpyAsync(&T, tlTotal.p, 4, cudaMemcpyDeviceToHost, stream));
          CK(cudaStreamSynchronize(stream));
          CK(cudaGetLastError());
          if (T == 0 || T > n) { err = (int)cudaErrorUnknown; goto fail; }
#define RQ_TREELET_LAUNCH(KK, SW) do {                                                                                          \
            const size_t smem = 4 * sizeof(TreeletSmem<KK, SW>);                                                                      \
            CK(cudaFuncSetAttribute(k_treelet_build<KK, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
            k_treelet_build<KK, SW><<<blocksFor(T, 4), 128, smem, stream>>>(t, (int)n, tlStart.p, sizeAt.p, T, keys0.p);              \
          } while (0)
          if (K == 512u) { if (P.sweepBottom) RQ_TREELET_LAUNCH(512, TL_SWEEP); else RQ_TREELET_LAUNCH(512, 0); }
          else { if (P.sweepBottom) RQ_TREELET_LAUNCH(256, TL_SWEEP); else RQ_TREELET_LAUNCH(256, 0); }
#undef RQ_TREELET_LAUNCH
          CK(cudaEventRecord(ev[3], stream));
          if (K == 512u) k_treelet_dp<512><<<blocksFor(T, TDP_WARPS), TDP_WARPS * 32, 0, stream>>>(t, (int)n, tlStart.p, sizeAt.p, T, P.costNode, P.costTri, P.maxLeafTris);
          else k_treelet_dp<256><<<blocksFor(T, TDP_WARPS), TDP_WARPS * 32, 0, stream>>>(t, (int)n, tlStart.p, sizeAt.p, T, P.costNode, P.costTri, P.maxLeafTris);
          k_treelet_roots<<<blocksFor(T, 256), 256, 0, stream>>>((int)n, tlStart.p, sizeAt.p, T, cid0.p);
          rqCountLaunch(3);
          CK(cudaGetLastError());
          m0 = T; numTreelets = T;
        } else {
          k_ploc_init<<<blocksFor(n, 256), 256, 0, stream>>>(t, (int)n, trisIn.p, vals0.p, P.costTri, cid0.p);
          rqCountLaunch(1);
          CK(cudaEventRecord(ev[3], stream));
        }
        // ---- PLOC: iterate nearest-neighbour search / merge / compaction until one cluster is left ----
        // counters on the device: [0] next inner node id, [1]/[2] cluster count of the current / next iteration (ping-pong),
        // [3] iterations that merged something, [4] blocks of the running merge kernel that are done.  The host reads the count back only every few iterations (to shrink the
        // grids and to detect the end); in between the kernels are launched for the last known upper bound.
        const uint32_t initCtr[5] = {m0 >= 2u ? m0 - 2u : 0u, m0, m0, 0u, 0u};
        CK(cudaMemcpyAsync(plocCtr.p, initCtr, sizeof(initCtr), cudaMemcpyHostToDevice, stream));
        uint32_t m = m0; uint32_t *cin = cid0.p, *cout = cid1.p;
        uint32_t it = 0;
        uint32_t plocCap = 4096;                                  // iteration bound of the grid-level loop (RQ_B200_PLOC_CAP: test hook for the stall fallback)
        if (const char* ev = getenv("RQ_B200_PLOC_CAP")) plocCap = (uint32_t)atoi(ev);
        while (m > PLOC_TAIL_MAX) {                                 // large cluster counts: one grid per step
          const unsigned nb = blocksFor(m, PLOC_THREADS);
          const uint32_t window = m > (1u << 20) ? 2u : (m > (1u << 16) ? 4u : 8u);
          for (uint32_t w = 0; w < window; w++, it++) {
            uint32_t* mCur = plocCtr.p + 1 + (it & 1u);
            uint32_t* mNext = plocCtr.p + 1 + ((it + 1u) & 1u);
            k_ploc_nn<<<nb, PLOC_THREADS, 0, stream>>>(t, cin, mCur, radius, nnBuf.p);
            k_ploc_merge<<<nb, PLOC_THREADS, 0, stream>>>(t, cin, mCur, nnBuf.p, flag.p, blockCount.p, plocCtr.p, P.costNode, P.costTri, P.maxLeafTris,
                                                          mNext, plocCtr.p + 3, plocCtr.p + 4);
            k_ploc_compact<<<nb, PLOC_THREADS, 0, stream>>>(cin, mCur, flag.p, blockCount.p, cout);
            rqCountLaunch(3);
            std::swap(cin, cout);
          }
          uint32_t next = 0;
          CK(cudaMemcpyAsync(&next, plocCtr.p + 1 + (it & 1u), 4, cudaMemcpyDeviceToHost, stream));
          CK(cudaStreamSynchronize(stream));
          if (next >= m || next == 0) { err = (int)cudaErrorUnknown; goto fail; }   // every iteration merges at least the globally closest pair
          m = next;
          if (it > plocCap) { err = RQ_BUILD_STALLED; goto fail; }
        }
        if (m > 1) {                                                // the tail (or everything, for treelet roots): one block, no host round trips
          constexpr size_t tailSmem = 6 * (size_t)PLOC_TAIL_MAX * sizeof(float);
          CK(cudaFuncSetAttribute(k_ploc_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tailSmem));
          k_ploc_tail<<<1, PLOC_TAIL_THREADS, tailSmem, stream>>>(t, cin, cout, nnBuf.p, plocCtr.p + 1 + (it & 1u), plocCtr.p, plocCtr.p + 3, radius,
                                                           P.costNode, P.costTri, P.maxLeafTris);
          rqCountLaunch(1);
        }
        CK(cudaMemcpyAsync(&plocIters, plocCtr.p + 3, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaGetLastError());
      } else {
        if (n > 1) { k_hierarchy<<<blocksFor(n - 1, 256), 256, 0, stream>>>(keys0.p, (int)n, left.p, right.p, parent.p, rangeFirst.p); rqCountLaunch(1); }
        CK(cudaEventRecord(ev[3], stream));
        k_refit_dp<<<blocksFor(n, 256), 256, 0, stream>>>(t, (int)n, trisIn.p, vals0.p, P.costNode, P.costTri, P.maxLeafTris);
        rqCountLaunch(1);
        CK(cudaGetLastError());
      }
      CK(cudaEventRecord(ev[4], stream));

      // ---- emission, level by level ----
      hc.nodeCount = 1;
      CK(cudaMemcpyAsync(dCtr.p, &hc, sizeof(hc), cudaMemcpyHostToDevice, stream));
      const uint2 rootEntry = make_uint2(n > 1 ? 0u : 0u /* n==1: node 0 is the leaf */, 0u);
      CK(cudaMemcpyAsync(queue0.p, &rootEntry, sizeof(uint2), cudaMemcpyHostToDevice, stream));
      uint2 *qin = queue0.p, *qout = queue1.p; uint32_t *pin = qParent0.p, *pout = qParent1.p;
      {
        constexpr uint32_t MAXL = 208, BATCH = 4;                 // levels are launched BATCH at a time without waiting for their counts
        static int emitGrid = 0;
        if (!emitGrid) { int dv = 0, sms = 0; cudaGetDevice(&dv); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dv); emitGrid = (sms > 0 ? sms : 148) * 4; }
        CK(dLevels.alloc(MAXL + 1));
        std::vector<EmitLevel> hl(MAXL + 1);
        const EmitLevel first = {1u, 0u};
        CK(cudaMemcpyAsync(dLevels.p, &first, sizeof(first), cudaMemcpyHostToDevice, stream));
        uint32_t launched = 0, bound = 1; bool done = false;
        while (!done) {
          for (uint32_t k = 0; k < BATCH; k++, launched++) {
            const uint32_t grid = std::min<uint32_t>(blocksFor(bound, EMIT_GROUPS), (uint32_t)emitGrid);
            k_emit<<<grid, EMIT_THREADS, 0, stream>>>(t, (int)n, qin, dLevels.p, qout, dCtr.p, nodes.p, trisOut.p,
                                            trisIn.p, vals0.p, launched, launched ? pin : nullptr, pout,
                                            compact ? idx3.p : nullptr, compact ? trisC.p : nullptr, compact ? metaOut.p : nullptr);
            k_emit_advance<<<1, 32, 0, stream>>>(dCtr.p, dLevels.p, launched);
            rqCountLaunch(2);
            std::swap(qin, qout); std::swap(pin, pout);
            bound = bound > n / 8u ? n : bound * 8u;              // a level holds at most 8x the entries of the one above (and never more than n)
          }
          CK(cudaMemcpyAsync(hl.data(), dLevels.p, sizeof(EmitLevel) * (launched + 1), cudaMemcpyDeviceToHost, stream));
          CK(cudaStreamSynchronize(stream));
          CK(cudaGetLastError());
          while (depth < launched && hl[depth].count > 0) {        // levels that had entries (an empty level launched nothing)
            if (depth) levelEnd.push_back(hl[depth - 1].nodeEnd);  // nodes allocated while level depth-1 was emitted = level depth
            depth++;
          }
          if (depth < launched) done = true;
          else bound = std::max<uint32_t>(hl[launched].count, 1u);
          if (!done && hl[launched].count == 0) done = true;
          if (launched + BATCH > MAXL) { if (!done) { err = (int)cudaErrorUnknown; goto fail; } }
        }
        CK(cudaMemcpyAsync(&hc, dCtr.p, sizeof(hc), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
      }
      numNodes = hc.nodeCount; numTris = hc.triCount;
    } else {
      CK(cudaEventRecord(ev[3], stream)); CK(cudaEventRecord(ev[4], stream));
      k_empty_root<<<1, 32, 0, stream>>>(nodes.p); rqCountLaunch(1);
      depth = 1; numNodes = 1; numTris = 0;
    }
    CK(cudaEventRecord(ev[5], stream));
  }

  {                                                             // ---- final image ----
    H.magic = RQ_IMAGE_MAGIC; H.numNodes = numNodes; H.numTris = numTris; H.depth = depth; H.flags = sceneFlags;
    for (int k = 0; k < 3; k++) {
      H.lo[k] = n ? ord2f(hb.sceneLo[k]) : INFINITY; H.hi[k] = n ? ord2f(hb.sceneHi[k]) : -INFINITY;
    }
    H.lo[3] = H.hi[3] = 0.f;
    H.nodesOffset = 128;
    H.trisOffset = H.nodesOffset + (uint64_t)numNodes * sizeof(RQNode);
    if (compact) {
      H.layout = 1u; H.numVerts = numVerts;
      H.metaOffset = (H.trisOffset + (uint64_t)numTris * sizeof(RQTriC) + 127ull) & ~127ull;
      H.vertsOffset = (H.metaOffset + (uint64_t)numTris * 4ull + 127ull) & ~127ull;
      H.totalBytes = (H.vertsOffset + (uint64_t)numVerts * 16ull + 127ull) & ~127ull;
    } else {
      H.totalBytes = (H.trisOffset + (uint64_t)numTris * sizeof(RQTri) + 127ull) & ~127ull;
    }
    const double rootA = n ? (double)halfArea(H.hi[0] - H.lo[0], H.hi[1] - H.lo[1], H.hi[2] - H.lo[2]) : 0.0;
    H.sah = rootA > 0 ? (hc.sahInnerQ + hc.sahLeafQ) / rootA : 0.0;
    if (!monitorAlloc(H.totalBytes)) { err = (int)cudaErrorMemoryAllocation; goto fail; }
    imageAccounted = (size_t)H.totalBytes;
    CK(cudaMallocAsync(&image, H.totalBytes, stream));
    {
      // every section is written in full below; only the alignment gaps behind the sections need zeroing (the whole-image memset was 0.1 ms for 0.66 GB)
      auto zeroGap = [&](uint64_t from, uint64_t to) -> cudaError_t {
        return to > from ? cudaMemsetAsync((char*)image + from, 0, (size_t)(to - from), stream) : cudaSuccess;
      };
      if (compact) {
        CK(zeroGap(H.trisOffset + (uint64_t)numTris * sizeof(RQTriC), H.metaOffset));
        CK(zeroGap(H.metaOffset + (uint64_t)numTris * 4ull, H.vertsOffset));
        CK(zeroGap(H.vertsOffset + (uint64_t)numVerts * 16ull, H.totalBytes));
      } else {
        CK(zeroGap(H.trisOffset + (uint64_t)numTris * sizeof(RQTri), H.totalBytes));
      }
    }
    CK(cudaMemcpyAsync(image, &H, sizeof(H), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync((char*)image + H.nodesOffset, nodes.p, (size_t)numNodes * sizeof(RQNode), cudaMemcpyDeviceToDevice, stream));
    if (compact) {
      if (numTris) {
        CK(cudaMemcpyAsync((char*)image + H.trisOffset, trisC.p, (size_t)numTris * sizeof(RQTriC), cudaMemcpyDeviceToDevice, stream));
        CK(cudaMemcpyAsync((char*)image + H.metaOffset, metaOut.p, (size_t)numTris * 4, cudaMemcpyDeviceToDevice, stream));
      }
      if (numVerts) {                                            // the pool is written straight into the image
        k_copy_verts<<<blocksFor(numVerts, 256), 256, 0, stream>>>(dGeoms.p, numGeoms, numVerts, (float4*)((char*)image + H.vertsOffset));
        rqCountLaunch(1);
        CK(cudaGetLastError());
      }
    } else if (numTris) CK(cudaMemcpyAsync((char*)image + H.trisOffset, trisOut.p, (size_t)numTris * sizeof(RQTri), cudaMemcpyDeviceToDevice, stream));
    CK(cudaEventRecord(ev[6], stream));
    CK(cudaStreamSynchronize(stream));
    if (stats) {
      memset(stats, 0, sizeof(*stats));
      stats->numPrimsIn = Nin; stats->numPrimsValid = n; stats->numSplitRefs = numSplitRefs; stats->numNodes = numNodes; stats->numTris = numTris;
      stats->depth = depth; stats->numLeaves = hc.leafSlots; stats->sah = H.sah;
      stats->sahExact = rootA > 0 ? (hc.sahInnerX + hc.sahLeafX) / rootA : 0.0;
      stats->sahInner = rootA > 0 ? hc.sahInnerQ / rootA : 0.0;
      stats->sahLeafTris = rootA > 0 ? hc.sahLeafTrisQ / rootA : 0.0;
      stats->bytes = H.totalBytes;
      stats->builderIterations = plocIters;
      stats->numTreelets = numTreelets;
      cudaEventElapsedTime(&stats->msTotal, ev[0], ev[6]);
      cudaEventElapsedTime(&stats->msPrims, ev[0], ev[1]);
      cudaEventElapsedTime(&stats->msSort, ev[1], ev[2]);
      cudaEventElapsedTime(&stats->msHierarchy, ev[2], ev[3]);
      cudaEventElapsedTime(&stats->msRefit, ev[3], ev[4]);
      cudaEventElapsedTime(&stats->msEmit, ev[4], ev[6]);
    }
    out->base = image; out->header = H; image = nullptr; imageAccounted = 0;
    out->numLevels = 0;
    if (levelEnd.size() == depth && depth <= RQ_MAX_LEVELS && levelEnd.back() == numNodes) {
      out->numLevels = depth;
      for (uint32_t l = 0; l < depth; l++) out->levelEnd[l] = levelEnd[l];
    }
  }
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  return 0;

fail:
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  if (image) cudaFreeAsync(image, stream);
  if (imageAccounted) monitorFree(imageAccounted);
  cudaGetLastError();
  return err ? err : (int)cudaErrorUnknown;
}

// ================================================================================================
// Refit: same topology, new vertex positions (RTC_BUILD_QUALITY_REFIT).
//   k_refit_tris   one thread per triangle record of the image: (geomID, primID) -> index buffer ->
//                  three vertices, validity rule of the build (invalid => NaN vertices: never hit)
//   k_refit_nodes  one thread per node of one tree level (deepest level first): child boxes from
//                  the leaf triangles / from the children's exact bounds written by the previous
//                  launch, new quantisation grid, same directed rounding as k_emit
// Traffic: 48 B read+written per triangle plus the vertex gathers, 128 B read+written per node.
// ================================================================================================
namespace {

__global__ void __launch_bounds__(256)
k_refit_tris(const RQGeomDesc* __restrict__ geomsByID, uint32_t numSlots, RQTri* __restrict__ tris, uint32_t numTris) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numTris) return;
  float4* rec = (float4*)(tris + i);
  const uint32_t odd = i & 1u;                                  // odd records: last 16 bytes first (rq_types.h)
  const float4 c = rec[odd ? 0 : 2];
  const uint32_t primID = __float_as_uint(c.y), geomID = __float_as_uint(c.z);
  const uint32_t pad = __float_as_uint(c.w);
  if (pad & RQ_PAD_INSTANCE) return;                            // instance records keep their box (a moved instance forces a rebuild)
  if (geomID >= numSlots) return;
  const RQGeomDesc G = geomsByID[geomID];
  if (G.indices == nullptr || G.vertices == nullptr) return;
  uint32_t i0 = RQ_INVALID, i1 = RQ_INVALID, i2 = RQ_INVALID;
  if (G.type == 2u) {
    if (2u * primID >= G.numTris) return;
    const uint32_t* ip = (const uint32_t*)(G.indices + (size_t)primID * G.indexStride);
    const uint32_t q0 = ip[0], q1 = ip[1], q2 = ip[2], q3 = ip[3];
    if (q0 < G.numVerts && q1 < G.numVerts && q2 < G.numVerts && q3 < G.numVerts) {
      const uint32_t other = (pad & RQ_PAD_FLIPUV) ? q0 : q2;
      const float* po = (const float*)(G.vertices + (size_t)other * G.vertexStride);
      bool ov = true;
      for (int k = 0; k < 3; k++) ov &= (po[k] > -RQ_FLT_LARGE) & (po[k] < RQ_FLT_LARGE);
      if (ov) { if (pad & RQ_PAD_FLIPUV) { i0 = q2; i1 = q3; i2 = q1; } else { i0 = q0; i1 = q1; i2 = q3; } }
    }
  } else {
    if (primID >= G.numTris) return;
    const uint32_t* ip = (const uint32_t*)(G.indices + (size_t)primID * G.indexStride);
    i0 = ip[0]; i1 = ip[1]; i2 = ip[2];
  }
  const float qnan = __uint_as_float(0x7FC00000u);
  float v[9] = {qnan, qnan, qnan, qnan, qnan, qnan, qnan, qnan, qnan};
  if (i0 < G.numVerts && i1 < G.numVerts && i2 < G.numVerts) {
    const float* p0 = (const float*)(G.vertices + (size_t)i0 * G.vertexStride);
    const float* p1 = (const float*)(G.vertices + (size_t)i1 * G.vertexStride);
    const float* p2 = (const float*)(G.vertices + (size_t)i2 * G.vertexStride);
    bool valid = true;
    for (int k = 0; k < 3; k++) {
      v[k] = p0[k]; v[3 + k] = p1[k]; v[6 + k] = p2[k];
    }
    for (int k = 0; k < 9; k++) valid &= (v[k] > -RQ_FLT_LARGE) & (v[k] < RQ_FLT_LARGE);
    if (!valid) for (int k = 0; k < 9; k++) v[k] = qnan;
  }
  const float4 a = make_float4(v[0], v[1], v[2], v[3]), b = make_float4(v[4], v[5], v[6], v[7]);
  const float4 c2 = make_float4(v[8], c.y, c.z, c.w);
  if (odd) { rec[0] = c2; rec[1] = a; rec[2] = b; } else { rec[0] = a; rec[1] = b; rec[2] = c2; }
}

struct RefitSums { double sahInnerQ, sahLeafQ, sahInnerX, sahLeafX, sahLeafTrisQ; };

template <bool COMPACT>
__global__ void __launch_bounds__(128)
k_refit_nodes(RQNode* __restrict__ nodes, const RQTri* __restrict__ tris, uint32_t first, uint32_t count, RefitSums* sums,
              const RQTriC* __restrict__ trisC, const float4* __restrict__ pool) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (q < count) {
    RQNode N;
    {
      const uint4* s4 = (const uint4*)(nodes + first + q); uint4* d4 = (uint4*)&N;
      #pragma unroll
      for (int i = 0; i < 8; i++) d4[i] = s4[i];
    }
    const uint32_t imask = N.masks >> 24, tvalid = N.masks & 0x00FFFFFFu;
    float clo[8][3], chi[8][3]; bool present[8], empty[8];
    float nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = 0; k < 8; k++) {
      present[k] = false; empty[k] = true;
      for (int a = 0; a < 3; a++) { clo[k][a] = FLT_MAX; chi[k][a] = -FLT_MAX; }
      if (imask & (1u << k)) {
        present[k] = true;
        const RQNode* ch = nodes + N.childBase + __popc(imask & ((1u << k) - 1u));
        for (int a = 0; a < 3; a++) { clo[k][a] = ch->lo[a]; chi[k][a] = ch->hi[a]; }
      } else {
        const uint32_t bits = (tvalid >> (3 * k)) & 7u;
        if (bits) {
          present[k] = true;
          for (uint32_t j = 0; j < 3; j++) {
            if (!(bits & (1u << j))) continue;
            const uint32_t b = 3u * k + j;
            const uint32_t ti = N.triBase + __popc(tvalid & ((1u << b) - 1u));
            float4 ta, tb, tc;
            if (COMPACT) {                                      // same nine floats, gathered through the indices
              const RQTriC r = trisC[ti];
              const float4 p0 = pool[r.v0], p1 = pool[r.v1], p2 = pool[r.v2];
              ta = make_float4(p0.x, p0.y, p0.z, p1.x); tb = make_float4(p1.y, p1.z, p2.x, p2.y); tc = make_float4(p2.z, 0.f, 0.f, 0.f);
            } else {
              const float4* src = (const float4*)(tris + ti);
              if (ti & 1u) { tc = src[0]; ta = src[1]; tb = src[2]; } else { ta = src[0]; tb = src[1]; tc = src[2]; }
            }
            // fminf / fmaxf drop NaN operands: a triangle invalidated by k_refit_tris (or with non-finite pool vertices) adds nothing
            clo[k][0] = fminf(clo[k][0], fminf(fminf(ta.x, ta.w), tb.z)); chi[k][0] = fmaxf(chi[k][0], fmaxf(fmaxf(ta.x, ta.w), tb.z));
            clo[k][1] = fminf(clo[k][1], fminf(fminf(ta.y, tb.x), tb.w)); chi[k][1] = fmaxf(chi[k][1], fmaxf(fmaxf(ta.y, tb.x), tb.w));
            clo[k][2] = fminf(clo[k][2], fminf(fminf(ta.z, tb.y), tc.x)); chi[k][2] = fmaxf(chi[k][2], fmaxf(fmaxf(ta.z, tb.y), tc.x));
          }
        }
      }
      if (present[k]) {
        empty[k] = !(clo[k][0] <= chi[k][0] && clo[k][1] <= chi[k][1] && clo[k][2] <= chi[k][2]);
        if (!empty[k]) for (int a = 0; a < 3; a++) { nlo[a] = fminf(nlo[a], clo[k][a]); nhi[a] = fmaxf(nhi[a], chi[k][a]); }
      }
    }
    const bool nodeEmpty = !(nlo[0] <= nhi[0] && nlo[1] <= nhi[1] && nlo[2] <= nhi[2]);
    float inv[3], step[3];
    for (int a = 0; a < 3; a++) {
      N.p[a] = nodeEmpty ? 0.f : nlo[a];
      const float ext = nodeEmpty ? 0.f : __fsub_ru(nhi[a], nlo[a]);
      uint8_t eb = expForExtent(ext);
      while (eb < 254 && __fmul_ru(ext, __uint_as_float((uint32_t)(254 - eb) << 23)) > 255.0f) eb++;
      N.e[a] = eb;
      step[a] = __uint_as_float((uint32_t)eb << 23);
      inv[a] = __uint_as_float((uint32_t)(254 - eb) << 23);
    }
    for (int k = 0; k < 8; k++) {
      for (int a = 0; a < 3; a++) { N.qlo[a][k] = 255; N.qhi[a][k] = 0; }   // absent / empty slot: never entered
      if (!present[k] || empty[k]) continue;
      float dq[3];
      for (int a = 0; a < 3; a++) {
        float fl = floorf(__fmul_rd(__fsub_rd(clo[k][a], N.p[a]), inv[a]));
        float fh = ceilf(__fmul_ru(__fsub_ru(chi[k][a], N.p[a]), inv[a]));
        fl = fminf(fmaxf(fl, 0.f), 255.f); fh = fminf(fmaxf(fh, 0.f), 255.f);
        N.qlo[a][k] = (uint8_t)fl; N.qhi[a][k] = (uint8_t)fh;
        dq[a] = (fh - fl) * step[a];
      }
      const double Aq = (double)halfArea(dq[0], dq[1], dq[2]);
      const double Ax = (double)halfArea(chi[k][0] - clo[k][0], chi[k][1] - clo[k][1], chi[k][2] - clo[k][2]);
      if (imask & (1u << k)) { s[0] += Aq; s[2] += Ax; } else { s[1] += Aq; s[3] += Ax; s[4] += Aq * (double)__popc((tvalid >> (3 * k)) & 7u); }
    }
    // an empty node keeps an inverted exact box, so its parent skips it as well
    for (int a = 0; a < 3; a++) { N.lo[a] = nodeEmpty ? FLT_MAX : nlo[a]; N.hi[a] = nodeEmpty ? -FLT_MAX : nhi[a]; }
    if (first + q == 0 && !nodeEmpty) {
      const double A = (double)halfArea(nhi[0] - nlo[0], nhi[1] - nlo[1], nhi[2] - nlo[2]);
      s[0] += A; s[2] += A;                                     // the root's own box
    }
    {
      const uint4* s4 = (const uint4*)&N; uint4* d4 = (uint4*)(nodes + first + q);
      #pragma unroll
      for (int i = 0; i < 8; i++) d4[i] = s4[i];
    }
  }
  #pragma unroll
  for (int k = 0; k < 5; k++)
    for (int o = 16; o; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  if ((threadIdx.x & 31) == 0 && (s[0] != 0.0 || s[1] != 0.0)) {
    atomicAdd(&sums->sahInnerQ, s[0]); atomicAdd(&sums->sahLeafQ, s[1]);
    atomicAdd(&sums->sahInnerX, s[2]); atomicAdd(&sums->sahLeafX, s[3]); atomicAdd(&sums->sahLeafTrisQ, s[4]);
  }
}

}  // namespace

int rqRefitBVH(const RQGeomDesc* geomsByID, int numSlots, RQDeviceImage* img, rqStream stream_, RQBuildStats* stats) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!img || !img->base || img->numLevels == 0 || img->numLevels > RQ_MAX_LEVELS) return (int)cudaErrorInvalidValue;
  ScratchScope scratch(stream);
  t_scratchStream = stream;
  int err = 0;
  RQImageHeader& H = img->header;
  RQNode* nodes = (RQNode*)((char*)img->base + H.nodesOffset);
  RQTri* tris = (RQTri*)((char*)img->base + H.trisOffset);
  DevBuf<RQGeomDesc> dGeoms; DevBuf<RefitSums> dSums;
  RefitSums hs; float rootBox[6];
  cudaEvent_t ev[3]; for (auto& e : ev) e = nullptr;
  memset(&hs, 0, sizeof(hs));

  for (auto& e : ev) CK(cudaEventCreate(&e));
  CK(cudaEventRecord(ev[0], stream));
  CK(dGeoms.alloc(numSlots > 0 ? numSlots : 1)); CK(dSums.alloc(1));
  if (numSlots > 0) CK(cudaMemcpyAsync(dGeoms.p, geomsByID, sizeof(RQGeomDesc) * numSlots, cudaMemcpyHostToDevice, stream));
  CK(cudaMemsetAsync(dSums.p, 0, sizeof(RefitSums), stream));
  if (H.layout == 1u) {
    // compact image: the records hold indices, only the vertex pool changes (meshes in geomID order, as at build time)
    std::vector<RQGeomDesc> hg(geomsByID, geomsByID + (numSlots > 0 ? numSlots : 0));
    uint64_t tv = 0;
    for (auto& g : hg) { g.vertBase = (uint32_t)tv; if (g.type != 1u && g.indices != nullptr) tv += g.numVerts; else g.numVerts = 0; }
    if (tv != H.numVerts) { err = (int)cudaErrorInvalidValue; goto fail; }      // not the meshes this image was built from
    if (numSlots > 0) CK(cudaMemcpyAsync(dGeoms.p, hg.data(), sizeof(RQGeomDesc) * numSlots, cudaMemcpyHostToDevice, stream));
    if (H.numVerts) {
      k_copy_verts<<<blocksFor(H.numVerts, 256), 256, 0, stream>>>(dGeoms.p, numSlots, H.numVerts, (float4*)((char*)img->base + H.vertsOffset));
      rqCountLaunch(1);
      CK(cudaGetLastError());
    }
  } else if (H.numTris) {
    k_refit_tris<<<blocksFor(H.numTris, 256), 256, 0, stream>>>(dGeoms.p, (uint32_t)(numSlots > 0 ? numSlots : 0), tris, H.numTris);
    rqCountLaunch(1);
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ev[1], stream));
  for (int l = (int)img->numLevels - 1; l >= 0; l--) {
    const uint32_t first = l ? img->levelEnd[l - 1] : 0u, count = img->levelEnd[l] - first;
    if (!count) continue;
    if (H.layout == 1u)
      k_refit_nodes<true><<<blocksFor(count, 128), 128, 0, stream>>>(nodes, tris, first, count, dSums.p, (const RQTriC*)((char*)img->base + H.trisOffset),
                                                                     (const float4*)((char*)img->base + H.vertsOffset));
    else
      k_refit_nodes<false><<<blocksFor(count, 128), 128, 0, stream>>>(nodes, tris, first, count, dSums.p, nullptr, nullptr);
    rqCountLaunch(1);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&hs, dSums.p, sizeof(hs), cudaMemcpyDeviceToHost, stream));
  CK(cudaMemcpyAsync(rootBox, (const char*)nodes + 80, sizeof(rootBox), cudaMemcpyDeviceToHost, stream));   // RQNode::lo, hi of node 0
  CK(cudaStreamSynchronize(stream));
  {
    const bool any = rootBox[0] <= rootBox[3] && rootBox[1] <= rootBox[4] && rootBox[2] <= rootBox[5];
    for (int k = 0; k < 3; k++) { H.lo[k] = any ? rootBox[k] : INFINITY; H.hi[k] = any ? rootBox[3 + k] : -INFINITY; }
    const double rootA = any ? (double)halfArea(H.hi[0] - H.lo[0], H.hi[1] - H.lo[1], H.hi[2] - H.lo[2]) : 0.0;
    H.sah = rootA > 0 ? (hs.sahInnerQ + hs.sahLeafQ) / rootA : 0.0;
    CK(cudaMemcpyAsync(img->base, &H, sizeof(H), cudaMemcpyHostToDevice, stream));
    CK(cudaEventRecord(ev[2], stream));
    CK(cudaStreamSynchronize(stream));
    if (stats) {
      stats->sah = H.sah;
      stats->sahExact = rootA > 0 ? (hs.sahInnerX + hs.sahLeafX) / rootA : 0.0;
      stats->sahInner = rootA > 0 ? hs.sahInnerQ / rootA : 0.0;
      stats->sahLeafTris = rootA > 0 ? hs.sahLeafTrisQ / rootA : 0.0;
      stats->msSort = stats->msHierarchy = stats->msEmit = 0.f;
      cudaEventElapsedTime(&stats->msTotal, ev[0], ev[2]);
      cudaEventElapsedTime(&stats->msPrims, ev[0], ev[1]);
      cudaEventElapsedTime(&stats->msRefit, ev[1], ev[2]);
      stats->builderIterations = 0;
      stats->refitCount++;
    }
  }
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  return 0;

fail:
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  cudaGetLastError();
  return err ? err : (int)cudaErrorUnknown;
}

// ================================================================================================
// Validation of an adopted image (rtcxSetSceneImage / rtcxLoadSceneImage): see rq_device.h.
// ================================================================================================
namespace {
__global__ void __launch_bounds__(256)
k_validate_image(const RQNode* __restrict__ nodes, const RQTri* __restrict__ tris, uint32_t numNodes, uint32_t numTris, uint32_t depth,
                 unsigned int* violations, uint32_t compact, uint32_t numVerts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  if (i < numNodes) {
    const RQNode* N = nodes + i;
    const uint32_t imask = N->masks >> 24, tvalid = N->masks & 0x00FFFFFFu;
    const uint32_t ni = __popc(imask), nt = __popc(tvalid);
    if (N->level >= depth) bad = true;
    if (ni) {
      // children live behind their parent (the builder emits level by level) and one level deeper
      if (N->childBase <= i || (uint64_t)N->childBase + ni > numNodes) bad = true;
      else for (uint32_t k = 0; k < ni; k++) if (nodes[N->childBase + k].level != N->level + 1u) bad = true;
    }
    if (nt && (uint64_t)N->triBase + nt > numTris) bad = true;
    for (uint32_t k = 0; k < 8; k++) {                          // a slot is inner or leaf, never both; leaf bits are contiguous from bit 0 of the slot
      const uint32_t tb = (tvalid >> (3 * k)) & 7u;
      if (((imask >> k) & 1u) && tb) bad = true;
      if (tb != 0u && tb != 1u && tb != 3u && tb != 7u) bad = true;
    }
  }
  if (i < numTris && compact) {
    const RQTriC r = ((const RQTriC*)tris)[i];
    if (r.v0 >= numVerts || r.v1 >= numVerts || r.v2 >= numVerts) bad = true;
  } else if (i < numTris) {
    const uint4* rec = (const uint4*)(tris + i);
    const uint32_t pad = (i & 1u) ? rec[0].w : rec[2].w;        // odd records are stored rotated (rq_types.h)
    if (pad & RQ_PAD_INSTANCE) bad = true;                      // instance records refer to other scenes' device memory
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAdd(violations, 1u);
}
}  // namespace

int rqValidateImage(const void* image, const RQImageHeader* H, rqStream stream_, unsigned int* violations) {
  cudaStream_t stream = (cudaStream_t)stream_;
  unsigned int* d = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&d, sizeof(unsigned int), stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(d, 0, sizeof(unsigned int), stream);
  const uint32_t n = H->numNodes > H->numTris ? H->numNodes : H->numTris;
  if (e == cudaSuccess && n) {
    k_validate_image<<<blocksFor(n, 256), 256, 0, stream>>>((const RQNode*)((const char*)image + H->nodesOffset),
                                                            (const RQTri*)((const char*)image + H->trisOffset), H->numNodes, H->numTris, H->depth, d,
                                                            H->layout, H->numVerts);
    rqCountLaunch(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(violations, d, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFreeAsync(d, stream);
  return (int)e;
}
