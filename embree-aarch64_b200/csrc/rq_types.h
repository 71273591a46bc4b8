// Device-resident data layout of the committed scene (plain PODs shared by host C++ and CUDA).
//
// What these replace in the reference (read-only, /root/reference):
//   RQNode  <- AABBNode_t<NodeRef,8> (256 B, kernels/bvh/bvh_node_aabb.h:199-206) and
//              QuantizedNode_t (136 B, kernels/bvh/bvh_node_qaabb.h:120-205)
//   RQTri   <- TriangleM<4> / TriangleMv<4> leaf blocks (176 B per 4 triangles,
//              kernels/geometry/triangle.h:13-161, trianglev.h)
//   the flat, offset-based image replaces the pointer-based BVHN<8> (kernels/bvh/bvh.h:41-231)
#pragma once
#include <stdint.h>

#define RQ_INVALID 0xFFFFFFFFu

// One 8-wide node = one 128-byte, 128-byte-aligned line.
// Traversal reads bytes 0..95 as three 32-byte loads (LDG.E.256); bytes 0..79 are the hot fields:
//   child box k, axis a:  lo = p[a] + qlo[a][k] * 2^(e[a]-127),  hi = p[a] + qhi[a][k] * 2^(e[a]-127)
//   (floor/ceil quantised: the decoded box always contains the exact child box)
// masks: bits 24..31 = imask, bit 24+k set <=> slot k holds an inner node;
//        bits 0..23  = tvalid, slot k owns bits 3k..3k+2, one bit per triangle of a leaf slot
//        (1..3 triangles, low bits first); an empty slot has neither.
// inner child k lives at node index childBase + popcount(imask & ((1<<k)-1));
// the triangle of tvalid bit b lives at triangle index triBase + popcount(tvalid & ((1<<b)-1)).
// The remaining 48 bytes are "cold": exact bounds and bookkeeping for statistics / refit.
struct alignas(128) RQNode {
  float    p[3];            //  0  origin of the quantisation grid (= exact lower corner)
  uint8_t  e[3];            // 12  biased exponents of the grid step per axis
  uint8_t  pad0;            // 15
  uint32_t childBase;       // 16
  uint32_t triBase;         // 20
  uint32_t masks;           // 24  imask << 24 | tvalid
  uint32_t pad1;            // 28
  uint8_t  qlo[3][8];       // 32  qlo[axis][slot]
  uint8_t  qhi[3][8];       // 56
  // ---- cold part (never touched by traversal) ----
  float    lo[3];           // 80  exact fp32 bounds of this node
  float    hi[3];           // 92
  uint32_t parent;          // 104 parent node index, RQ_INVALID for the root
  uint32_t numTris;         // 108 triangles below this node
  uint32_t level;           // 112 depth (root = 0)
  uint32_t pad[3];          // 116
};
static_assert(sizeof(RQNode) == 128, "RQNode must be one 128-byte line");

// One triangle = 48 bytes, read as one 32-byte and one 16-byte load.  In the device image the
// records of ODD index are stored rotated -- bytes 32..47 (v2.z, primID, geomID, pad) first, then
// bytes 0..31 -- so that the 32-byte part of every record is 32-byte aligned (48*i + 16 for odd i).
// The struct below is the canonical (even / builder-input) order.  Vertices are stored verbatim (fp32 bits of the
// user's vertex buffer) so that e1 = v0-v1, e2 = v2-v0, Ng = cross(e2,e1) reproduce the reference's
// precomputed TriangleM<4>::e1/e2 bit for bit (triangle.h:45-51) and the watertight test can use
// v0,v1,v2 directly (trianglev.h).
struct alignas(16) RQTri {
  float    v0[3];
  float    v1[3];
  float    v2[3];
  uint32_t primID;
  uint32_t geomID;
  uint32_t pad;
};
static_assert(sizeof(RQTri) == 48, "RQTri must be 48 bytes");

struct alignas(16) RQTriC {   // compact (indexed) triangle record, see RQImageHeader::layout
  uint32_t v0, v1, v2;       // vertex pool indices
  uint32_t primID;
};
static_assert(sizeof(RQTriC) == 16, "RQTriC must be 16 bytes");
#define RQ_META_FLIPUV 0x80000000u

// One instance of another committed scene (RTC_GEOMETRY_TYPE_INSTANCE, single level), 80 bytes =
// five 16-byte loads.  A triangle record of the top-level BVH whose `pad` word is
// RQ_PAD_INSTANCE | i stands for instances[i]; its v0 / v1 hold the world-space bounds.
#define RQ_PAD_INVALID  1u
#define RQ_PAD_INSTANCE 0x80000000u
// second triangle (v2, v3, v1) of a quad (RTC_GEOMETRY_TYPE_QUAD): a hit reports u = 1 - u, v = 1 - v and the quad's primID
// (kernels/geometry/quad_intersector_moeller.h:122-144, quad_intersector_pluecker.h:179-180)
#define RQ_PAD_FLIPUV   0x40000000u
struct alignas(16) RQInstance {
  float    w2l[12];         // world -> instance space, column major: vx, vy, vz, p (AffineSpace3fa)
  uint64_t nodes;           // device address of the instanced scene's RQNode array
  uint64_t tris;            // device address of its RQTri array
  uint32_t geomID;          // geomID of the instance geometry in the top-level scene -> hit.instID[0]
  uint32_t depth;           // levels of the instanced BVH
  uint32_t pad[2];
};
static_assert(sizeof(RQInstance) == 80, "RQInstance is five 16-byte words");

// One triangle mesh as the builder sees it (device-readable pointers, arbitrary 4-byte-multiple strides).
struct RQGeomDesc {
  const uint8_t* indices;    // RTC_FORMAT_UINT3 records
  const uint8_t* vertices;   // RTC_FORMAT_FLOAT3 records
  uint32_t indexStride;
  uint32_t vertexStride;
  uint32_t numTris;
  uint32_t numVerts;
  uint32_t primBase;         // first global primitive number of this mesh
  uint32_t geomID;
  uint32_t type;             // 0 = triangle mesh; 1 = instance: one primitive with the bounds below (numTris = 1);
                             // 2 = quad mesh: RTC_FORMAT_UINT4 index records, numTris = 2 x quads (triangle 2q = (v0,v1,v3), 2q+1 = (v2,v3,v1))
  uint32_t instIndex;        // index into the scene's RQInstance table
  float    lo[3], hi[3];     // instance only: world-space bounds = xfmBounds(local2world, bounds of the instanced scene)
  uint32_t vertBase;         // first vertex of this mesh in the vertex pool of a compact image (filled in by the builder)
  uint32_t pad;
};

// Flat image of a committed BVH: header + nodes + triangles, all offset based, so a byte copy
// (cudaMemcpyPeer / NCCL broadcast) yields a usable replica on another GPU.
struct RQImageHeader {
  uint64_t magic;            // 'RQB200v2'
  uint32_t numNodes;
  uint32_t numTris;
  uint32_t depth;            // levels of 8-wide nodes (stack bound)
  uint32_t flags;            // RTCSceneFlags the scene was committed with
  float    lo[4];
  float    hi[4];
  uint64_t nodesOffset;      // byte offsets from the start of the image (128-aligned)
  uint64_t trisOffset;
  uint64_t totalBytes;
  double   sah;              // SAH cost, reference formula (bvh_statistics.h:36-38,99-101)
  // RTC_SCENE_FLAG_COMPACT layout (layout == 1; reference: Triangle4i, kernels/geometry/trianglei.h): the triangle section holds
  // 16-byte RQTriC records (three indices into the image's own vertex pool + primID), followed by one 4-byte word per triangle
  // (geomID, bit 31 = second half of a quad) that is only read for the final hit, and the vertex pool (one float4 per vertex of
  // every attached mesh, in geomID order).  ~28 instead of 48 bytes per triangle; one more dependent load per triangle test.
  uint64_t metaOffset;
  uint64_t vertsOffset;
  uint32_t numVerts;
  uint32_t layout;           // 0 = 48-byte RQTri records, 1 = compact
  uint64_t pad[2];
};
static_assert(sizeof(RQImageHeader) == 128, "header is one line");
#define RQ_IMAGE_MAGIC 0x3276303032425152ull   /* 'RQB200v2' */

struct RQBuildStats {
  uint32_t numPrimsIn;       // triangles submitted
  uint32_t numPrimsValid;    // after dropping out-of-range indices / non-finite vertices
  uint32_t numNodes;
  uint32_t numTris;
  uint32_t depth;
  uint32_t numLeaves;        // leaf slots
  double   sah;              // reference formula on de-quantised boxes
  double   sahExact;         // same with the exact fp32 child boxes
  float    msTotal, msPrims, msSort, msHierarchy, msRefit, msEmit;
  uint64_t bytes;
  uint32_t builderIterations; // PLOC merge iterations (0 for the radix tree)
  uint32_t refitCount;        // refits applied to this BVH since its last full build (0 = freshly built)
  double   sahInner;          // inner-node term of `sah` alone (sah - sahInner = leaf term, one block per leaf slot)
  double   sahLeafTris;       // leaf term weighted by triangles: sum A(leaf slot) * numTris / A(root)
  uint32_t numTreelets;       // binned-SAH treelets of the last full build (0 for the other front ends)
  float    msBroadcast;       // gpus=N: wall time of replicating the image to the peer GPUs after the last commit
  uint32_t numSplitRefs;      // extra primitive references created by the pre-split of large triangles (RTC_BUILD_QUALITY_HIGH)
  uint32_t pad;
};

// Per-call traversal counters (instrumented kernel variant only).
struct RQTraceCounters {
  unsigned long long rays;        // active rays traced
  unsigned long long nodes;       // 8-wide node records fetched
  unsigned long long tris;        // triangle records fetched
  unsigned long long hits;        // rays that report a hit / are occluded
  unsigned long long stackMax;    // deepest traversal stack seen
  unsigned long long emptyNodes;  // node records fetched whose children were all missed / culled
  unsigned long long hitNodes;    // node records fetched by rays that end up reporting a hit
  unsigned long long lateNodes;   // node records fetched although the node's own box already lies beyond the ray's current tfar
};
