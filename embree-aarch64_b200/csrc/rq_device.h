// Internal interface between the host API (api_*.cpp) and the CUDA side (rq_build.cu, rq_trace.cu).
// Plain functions, plain pointers; every function returns a cudaError_t-compatible int (0 = ok).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "rq_types.h"

typedef struct CUstream_st* rqStream;

struct RQBuildParams {
  float costNode;        // SAH cost of visiting one 8-wide node           (default 1.0)
  float costTri;         // SAH cost of testing one triangle               (default 0.3)
  int   maxLeafTris;     // triangles per leaf slot, 1..3                  (default 3)
  int   verbose;
  int   builder;         // binary-tree front end: 0 = radix tree over Morton codes (LBVH), 1 = PLOC over the triangles,
                         // 2 = binned-SAH treelets (top-down SAH inside Morton cells of <= treeletSize triangles, PLOC above them)
  int   plocRadius;      // PLOC search radius in Morton-order positions, 1..32          (default 8)
  int   mortonCubic;     // 1 = Morton cells are cubes (one scale for all axes), 0 = per-axis normalisation
  int   treeletSize;     // builder 2: largest treelet, 256 or 512 triangles
  int   presplit;        // 1 = large triangles are pre-split into several references with clipped boxes (RTC_BUILD_QUALITY_HIGH)
  int   sweepBottom;     // builder 2: 1 = exact sweep SAH for nodes of <= 16 triangles (slow, best leaves: RTC_BUILD_QUALITY_HIGH), 0 = radix rule
};

// A committed BVH living in device memory: one allocation, header first.
#define RQ_MAX_LEVELS 208
struct RQDeviceImage {
  void*         base;       // device pointer to the image (RQImageHeader at offset 0)
  RQImageHeader header;     // host copy of the header
  // Nodes are emitted one tree level per launch, so level L occupies the contiguous node range
  // [L ? levelEnd[L-1] : 0, levelEnd[L]); the refit walks these ranges bottom-up.  numLevels = 0
  // for an image adopted from elsewhere (rtcxSetSceneImage): such an image cannot be refitted.
  uint32_t      numLevels;
  uint32_t      levelEnd[RQ_MAX_LEVELS];
};

// Builds the BVH for the given meshes on `stream` (geoms is a HOST array whose index/vertex
// pointers must be device-readable).  On success *out owns a new device allocation.
// RQ_BUILD_STALLED: the PLOC stage did not converge within its iteration bound (adversarial input: e.g. boxes whose nearest
// neighbour is always the one to their left merge ONE pair per iteration); nothing was allocated, the caller may retry with the
// radix-tree front end (params->builder = 0), which has no such bound.
#define RQ_BUILD_STALLED (-1001)
int rqBuildBVH(const RQGeomDesc* geoms, int numGeoms, uint32_t sceneFlags, const RQBuildParams* params,
               rqStream stream, RQDeviceImage* out, RQBuildStats* stats);
void rqFreeImage(RQDeviceImage* img);
int rqAllocImage(void** p, size_t bytes, rqStream stream);   // memory rqFreeImage can release (stream-ordered pool)

// Refit (RTC_BUILD_QUALITY_REFIT, reference: kernels/bvh/bvh_refit.cpp, bvh_builder_twolevel.h:95-140):
// the topology of `img` is kept; every triangle record re-reads its three vertices through the
// index buffer of its mesh (geomsByID is a HOST array indexed by geomID, device-readable pointers,
// numTris = 0 for absent slots) and all node boxes are recomputed and re-quantised bottom-up, one
// launch per tree level.  Updates img->header (bounds, SAH) on host and device.
int rqRefitBVH(const RQGeomDesc* geomsByID, int numSlots, RQDeviceImage* img, rqStream stream, RQBuildStats* stats);

// Ray-stream kernels.  `rays` is device-readable AoS memory: RTCRayHit (intersect) or RTCRay
// (occluded) records `stride` bytes apart.  instID0 is written to hit.instID[0] on a hit.
struct RQTraceArgs {
  const void* image;       // device image base
  uint64_t    nodesOffset; // header.nodesOffset / header.trisOffset
  uint64_t    trisOffset;
  uint32_t    depth;       // header.depth
  uint32_t    robust;      // 1 = Pluecker (RTC_SCENE_FLAG_ROBUST), 0 = Moeller-Trumbore
  void*       rays;
  void*       out;         // where hit fields are written (same record layout); NULL = in place
  uint32_t    numRays;
  size_t      stride;
  uint32_t    instID0;
  uint32_t    streamSemantics;  // occluded only: 1 = stream entry rules (M>1), 0 = single-ray rules
  RQTraceCounters* counters;    // device pointer or NULL (NULL = fast kernel)
  unsigned int* workCounter;    // device scratch word, exclusive to this launch until it completes (zeroed by the launcher)
  uint32_t    refillBelow;      // persistent-kernel refill threshold in lanes, 0 = default
  uint32_t    split;            // traversal loop shape: 1 = one triangle per iteration, 0 = whole leaf list per node
  uint32_t    stackSmem;        // traversal stack levels kept in shared memory (0 = all in local memory), the rest spills to local
  uint32_t    tVote;            // split only: 0 = both phases every iteration, K = triangle phase when >= K lanes wait for it
  uint32_t    packed;           // 1 = `rays` holds dense 32-byte records {org.xyz, tnear, dir.xyz, tfar} (stride 32); needs hitList
  void*       hitList;          // compact output (flat scenes only): device buffer for one record per hit ray, 48 B (closest) or
  unsigned int* hitCount;       //   4 B (occluded), appended through this device counter (zeroed by the launcher); `out` is then unused
  uint32_t    compact;          // 1 = the image uses the compact (indexed) triangle layout: metaOffset / vertsOffset are valid
  uint64_t    metaOffset, vertsOffset;
  const void* instances;        // device array of RQInstance when the scene holds instance geometries, else NULL;
                                // `depth` must then cover top-level depth + 3 + the deepest instanced BVH
};
int rqLaunchIntersect(const RQTraceArgs* a, rqStream stream);
int rqLaunchOccluded(const RQTraceArgs* a, rqStream stream);

// Layout adapters for the non-AoS entry points (reference: RayStreamFilter::filterAOP / filterSOA / filterSOP,
// kernels/bvh/bvh_intersector_stream_filters.cpp:155-592): gather every ray of the call into a dense AoS scratch buffer of
// 80-byte RTCRayHit records, trace that with the stream kernels, scatter the results back.  All pointers are device readable.
struct RQSoAView {              // ray i = lane (i % N) of packet (i / N); its field lives at field + (i / N) * packetStride bytes + (i % N) * 4
  float *org_x, *org_y, *org_z, *tnear, *dir_x, *dir_y, *dir_z, *time, *tfar;
  unsigned int *mask, *id, *flags;
  float *Ng_x, *Ng_y, *Ng_z, *u, *v;
  unsigned int *primID, *geomID, *instID0;
  uint32_t N;                   // packet width (RTCRayHitNp / one packet: N = number of rays, packetStride = 0)
  size_t   packetStride;        // bytes between consecutive packets (rtcIntersectNM byteStride)
};
int rqGatherSoA(const RQSoAView* v, const int* valid, uint32_t n, void* aosRayHit, rqStream stream);
int rqScatterSoA(const RQSoAView* v, uint32_t n, const void* aosRayHit, int occluded, rqStream stream);
// array-of-pointers streams (rtcIntersect1Mp / rtcOccluded1Mp): ptrs[i] addresses an RTCRayHit (80 B) or an RTCRay (48 B)
int rqGatherAoP(const void* const* ptrs, uint32_t n, int recBytes, void* aosRayHit, rqStream stream);
int rqScatterAoP(void* const* ptrs, uint32_t n, const void* aosRayHit, int occluded, rqStream stream);

// Consistency check of an image that did not come out of rqBuildBVH (rtcxSetSceneImage / rtcxLoadSceneImage): every child and
// triangle reference in range, child levels = parent level + 1, deepest level < header depth, no instance records.  *violations
// receives the number of offending records (0 = usable).
int rqValidateImage(const void* image, const RQImageHeader* header, rqStream stream, unsigned int* violations);

// Memory monitor hook (reference: RTCMemoryMonitorFunction, kernels/common/device.cpp:318-327).  While a monitor is set for the
// calling thread, every device allocation of rqBuildBVH / rqRefitBVH (scratch and the image) first asks fn(user, +bytes, false) --
// false vetoes it and the build fails with cudaErrorMemoryAllocation -- and reports fn(user, -bytes, true) when the memory is
// released.  The image of a successful build stays accounted: whoever frees it reports -header.totalBytes.
typedef bool (*rqAllocMonitorFn)(void* user, long long bytes, bool post);
void rqSetAllocMonitor(rqAllocMonitorFn fn, void* user);

// Number of kernel launches issued by this library since load (bench.py's gpu_launches claim).
unsigned long long rqLaunchCount(void);
void rqCountLaunch(unsigned n);
