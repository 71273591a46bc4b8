/* b200-rayquery: B200-specific additions to the rtcore C ABI.
 *
 * Nothing here exists in the reference; these entry points expose what a GPU engine has that a
 * CPU library does not: the CUDA stream queries are enqueued on, the flat device image of a
 * committed BVH (for NVLink replication to other GPUs), build statistics comparable with the
 * reference's BENCHMARK_BUILD line (kernels/bvh/bvh.cpp:173-178) and STAT counters
 * (kernels/common/stat.h:61-79), and a launch counter for benchmarking.
 * Plain C, plain pointers and sizes; no CUDA or torch types in any signature.
 */
#ifndef B200_RAYQUERY_EXT_H
#define B200_RAYQUERY_EXT_H

#include "embree3/rtcore.h"

struct RTCXBuildStats {
  unsigned int numPrimsIn;      /* triangles submitted to the build                               */
  unsigned int numPrimsValid;   /* after dropping invalid ones (scene_triangle_mesh.h:131-153)    */
  unsigned int numNodes;        /* 128-byte 8-wide nodes                                          */
  unsigned int numTris;         /* 48-byte triangle records                                       */
  unsigned int depth;           /* levels of 8-wide nodes                                         */
  unsigned int numLeaves;       /* leaf slots (1..3 triangles each)                               */
  double sah;                   /* SAH cost, reference formula (bvh_statistics.h:36-38,99-101) on the
                                   de-quantised boxes traversal actually tests                    */
  double sahExact;              /* same formula on the exact fp32 child boxes                     */
  float msTotal, msPrims, msSort, msHierarchy, msRefit, msEmit;   /* device time per build phase  */
  unsigned long long bytes;     /* size of the device image                                       */
  unsigned int builderIterations; /* PLOC merge iterations (0 for the radix-tree front end)       */
  unsigned int refitCount;      /* refits since the last full build (RTC_BUILD_QUALITY_REFIT path) */
  double sahInner;              /* inner-node term of `sah`; sah - sahInner = leaf term with one block per leaf slot,
                                   the weighting of BVHNStatistics (a reference leaf block holds <= 4 triangles, a slot <= 3) */
  double sahLeafTris;           /* leaf term weighted by triangles instead of blocks: sum A(slot) * numTris / A(root);
                                   sahInner + sahLeafTris / 4 is the figure to hold against the reference's blocks of four */
  unsigned int numTreelets;     /* binned-SAH treelets of the last full build (gpu_builder=sah), 0 for the other front ends */
  float msBroadcast;            /* device option gpus=N: wall time of the NVLink replication of the image to the peer GPUs */
  unsigned int numSplitRefs;    /* RTC_BUILD_QUALITY_HIGH: extra primitive references created by pre-splitting large triangles
                                   (numTris = numPrimsValid + numSplitRefs: a split triangle is stored once per reference) */
  unsigned int pad;
};

struct RTCXTraceCounters {
  unsigned long long rays;      /* active rays traced                                             */
  unsigned long long nodes;     /* node records fetched  (x128 B: algorithmic node bytes)         */
  unsigned long long tris;      /* triangle records fetched (x48 B)                               */
  unsigned long long hits;      /* rays that found a hit / are occluded                           */
  unsigned long long stackMax;  /* deepest traversal stack                                        */
  unsigned long long emptyNodes;/* node records fetched whose children were all missed or culled  */
  unsigned long long hitNodes;  /* node records fetched by rays that report a hit                 */
  unsigned long long lateNodes; /* node records fetched although their own box lies beyond the current tfar */
};

/* Stream (a cudaStream_t passed as void*) on which builds and device-resident queries are
 * enqueued.  NULL selects the device's own stream.  With the device option "async=1" calls on
 * device-resident ray buffers return without synchronising (stream ordered). */
RTC_API void rtcxSetDeviceStream(RTCDevice device, void* cudaStream);
RTC_API void rtcxSynchronizeDevice(RTCDevice device);
RTC_API int  rtcxGetDeviceOrdinal(RTCDevice device);
/* GPUs this device object drives (device option "gpus=N": GPUs ordinal .. ordinal+N-1 of one box; a commit builds on the first
 * and replicates the image over NVLink, host-resident streams are sharded contiguously across all of them). */
RTC_API int  rtcxGetDeviceGpuCount(RTCDevice device);

/* Statistics of the last commit of `scene`; returns 0 on success. */
RTC_API int rtcxGetSceneBuildStats(RTCScene scene, struct RTCXBuildStats* stats_o);

/* The committed BVH as one flat, offset-based device allocation (header + nodes + triangles).
 * rtcxGetSceneImage returns the DEVICE pointer and its size; rtcxSetSceneImage makes `scene`
 * (on any device / any process) adopt a byte copy of such an image from device-readable memory
 * and marks it committed -- this is how a BVH built on one GPU is replicated after an NVLink
 * broadcast.  */
RTC_API const void* rtcxGetSceneImage(RTCScene scene, size_t* bytes_o);
RTC_API void rtcxSetSceneImage(RTCScene scene, const void* deviceImage, size_t bytes);
/* Copies the image into caller memory (device or host, `bytes` must equal the image size): the
 * send buffer of the broadcast. */
RTC_API void rtcxCopySceneImage(RTCScene scene, void* dst, size_t bytes);

/* The image as a file: save the committed BVH of `scene`, or make `scene` adopt a saved one
 * (validated like rtcxSetSceneImage).  Return 0 on success, -1 on error (see rtcGetDeviceError). */
RTC_API int rtcxSaveSceneImage(RTCScene scene, const char* path);
RTC_API int rtcxLoadSceneImage(RTCScene scene, const char* path);

/* Same as rtcIntersect1M / rtcOccluded1M, run with the instrumented kernel variant; the counters
 * are the measured numerator of the traversal roofline. */
RTC_API void rtcxIntersect1MCounted(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit* rayhit,
                                    unsigned int M, size_t byteStride, struct RTCXTraceCounters* counters_o);
RTC_API void rtcxOccluded1MCounted(RTCScene scene, struct RTCIntersectContext* context, struct RTCRay* ray,
                                   unsigned int M, size_t byteStride, struct RTCXTraceCounters* counters_o);

/* Kernels launched by this library since it was loaded. */
RTC_API unsigned long long rtcxGetLaunchCount(void);
/* Bytes the library has copied host->device / device->host for host-resident ray streams on `device` since it was created. */
RTC_API void rtcxGetTransferBytes(RTCDevice device, unsigned long long* h2d_o, unsigned long long* d2h_o);

#endif
