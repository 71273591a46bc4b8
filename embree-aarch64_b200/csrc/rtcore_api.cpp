// Host side of the drop-in boundary: the rtc* C entry points of include/embree3/rtcore.h and the
// small object model behind them.  Mirrors, for the triangle / ray-stream path only, the reference's
//   kernels/common/rtcore.cpp      (entry points, RTC_CATCH error convention: rtcore.h:28-54)
//   kernels/common/device.cpp      (config string, first-error-wins slot + callback: :258-316)
//   kernels/common/scene.cpp       (bind/detach with lowest-free geomID: :595-653, commit: :655-913)
//   kernels/common/geometry.cpp    (MODIFIED/COMMITTED state: :80-107)
//   kernels/common/scene_triangle_mesh.cpp (buffer rules: :35-80)
// All computation happens in CUDA (rq_build.cu, rq_trace.cu); there is no CPU fallback: without a
// usable CUDA device rtcNewDevice fails.
#define RTC_EXPORT_API
#include "../../include/rq_b200.h"
#include "rq_device.h"

#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <thread>
#include <functional>
#include <chrono>
#if defined(__SSE2__)
#include <immintrin.h>
#endif
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct rtc_error : public std::exception {
  RTCError code; std::string msg;
  rtc_error(RTCError c, const std::string& m) : code(c), msg(m) {}
  const char* what() const noexcept override { return msg.c_str(); }
};
[[noreturn]] void fail(RTCError c, const char* m) { throw rtc_error(c, m); }

void cudaCheck(int e, const char* what) {
  if (e == 0) return;
  const cudaError_t ce = (cudaError_t)e;
  cudaGetLastError();
  std::string m = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(ce);
  throw rtc_error(ce == cudaErrorMemoryAllocation ? RTC_ERROR_OUT_OF_MEMORY : RTC_ERROR_UNKNOWN, m);
}

thread_local RTCError g_threadError = RTC_ERROR_NONE;
std::mutex g_apiMutex;                                   // device create/retain/release (rtcore.cpp:20-36)

struct RefCounted {
  std::atomic<long> refs{1};
  virtual ~RefCounted() {}
  void retain() { refs.fetch_add(1); }
  void release() { if (refs.fetch_sub(1) == 1) delete this; }
};

// Small persistent pool of host threads (created on first use, lives as long as the device): packs rays for the upload and
// scatters downloaded hit lists.  The reference spends all host threads on traversal itself (TBB / internal tasking).
struct HostPool {
  std::vector<std::thread> th; std::mutex m; std::condition_variable cv; std::deque<std::function<void()>> q; bool quit = false;
  explicit HostPool(int n) {
    for (int i = 0; i < n; i++)
      th.emplace_back([this] {
        for (;;) {
          std::function<void()> fn;
          { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return quit || !q.empty(); }); if (q.empty()) return; fn = std::move(q.front()); q.pop_front(); }
          fn();
        }
      });
  }
  ~HostPool() { { std::lock_guard<std::mutex> lk(m); quit = true; } cv.notify_all(); for (auto& t : th) if (t.joinable()) t.join(); }
  void submit(std::function<void()> fn) { { std::lock_guard<std::mutex> lk(m); q.push_back(std::move(fn)); } cv.notify_one(); }
  size_t size() const { return th.size(); }
};

// --------------------------------------------------------------------------------------------
struct Device : RefCounted {
  int ordinal = 0;
  bool hasGpu = false;
  int verbose = 0, benchmark = 0, async = 0;
  size_t chunkRays = 1u << 20;
  RQBuildParams build{1.0f, 1.0f, 3, 0, 2, 8, 1, 256, 0, 0};   // binned-SAH treelets + PLOC above them by default (gpu_builder=sah); see DESIGN.md 4.1 for the A/B against PLOC / radix tree
  cudaStream_t ownStream = nullptr, userStream = nullptr;
  std::mutex errMutex;
  RTCError error = RTC_ERROR_NONE;
  RTCErrorFunction errFn = nullptr; void* errPtr = nullptr;
  RTCMemoryMonitorFunction memFn = nullptr; void* memPtr = nullptr;
  // staging for host-resident ray streams: a small ring of (stream, device buffer) pairs
  static const int kRing = 6;
  std::mutex stageMutex;
  cudaStream_t ringStream[kRing] = {nullptr, nullptr, nullptr, nullptr};
  void* ringBuf[kRing] = {nullptr, nullptr, nullptr, nullptr};
  size_t ringCap[kRing] = {0, 0, 0, 0};
  // compact hit download (d2h=3): per ring slot a device hit list + counter and their page-locked host mirrors
  void* listDev[kRing] = {nullptr, nullptr, nullptr, nullptr};
  void* listHost[kRing] = {nullptr, nullptr, nullptr, nullptr};
  size_t listCap[kRing] = {0, 0, 0, 0};
  unsigned int* countHost = nullptr;      // kRing page-locked words
  unsigned int* countDev = nullptr;       // kRing device words, 32 bytes apart
  RQTraceCounters* dCounters = nullptr;
  unsigned int* dWork = nullptr;          // ray cursors of the persistent kernels: one per ring stream + one for the device stream
  std::mutex launchMutex;                 // (cursor reset + launch) pairs on the device stream are enqueued atomically
  // traversal schedule per query kind, tuned on B200 (DESIGN.md "traversal schedule"): incoherent
  // closest-hit streams refill idle lanes early and take one triangle per iteration; coherent
  // streams and occlusion streams keep a warp's rays together (late refill, whole leaf lists)
  int refillClosest = 26, refillCoherent = 4, refillOccluded = 4;
  int splitClosest = 1, splitCoherent = 0, splitOccluded = 0;
  int tVote = 0;
  int stackSmem = 16;                     // stack levels kept in shared memory, deeper levels spill to local memory (same-box A/B: +4.4 % closest, +1 % occluded vs local only)
  // page-locked host streams: 0 = staged both ways (H2D, kernel, D2H); 1 = rays staged by DMA, hit fields written by
  // the kernel straight into the caller's buffer over PCIe (no D2H copy); 2 = traced in place over PCIe (no copies at all).
  // Same-box A/B on B200 / PCIe Gen5 (profiles/r01g_ab.log): 0 -> 664 Mrays/s, 1 -> 606, 2 -> 603: SM-issued PCIe
  // transactions are 32-64 bytes and lose to the copy engines' large TLPs, so staging stays the default.
  int zeroCopy = 0;
  // hit download of staged host streams: 0 = the whole span back (one linear copy), 1 = only bytes [32, record size) of every
  // record (tfar .. hit; a strided 2-D copy, the ray part never changes), 2 = as 1 but tfar alone for occlusion streams,
  // 3 = compact: the kernel appends one record per hit ray to a list, only the list is downloaded and a host thread scatters
  // it into the caller's buffer (default).  Same-box A/B (profiles/r01o_ab_d2h_rows.log, r01p2_ab_compact_pool.log):
  // 0 -> 662-668 Mrays/s end to end, 1 -> 463, 2 -> 446 (the copy engines handle 48-byte rows badly), 3 -> 804
  int d2hMode = 3;
  // the compact path's stage hand-offs (two host callbacks + pool wake-ups per chunk) cost ~2 ms per call: below ~4 M rays the plain
  // path with whole-span copies wins (configs[0], 1 M rays: 4.5 ms compact vs 2.35 ms plain; profiles/r01y_small_probe.log)
  unsigned compactMinRays = 4u << 20;
  int scatterThreads = 8;                 // d2h=3: host threads that scatter hit lists (and pack rays); 0 = all hardware threads; at most 16
  // d2h=3: 0 = upload whole records (default), 1 = hybrid (pack while the pool has room, see traceStreamCompact), 2 = pack every chunk.
  // Same-box probe on configs[1] (profiles/r01r_e2e_probe.log, 16 host threads): 0 -> 777-804 Mrays/s, 1 -> 788-826, 2 -> 733-776:
  // packing reads the caller's 80-byte records at ~50 GB/s, the same rate the copy engine uploads them, and together with the
  // hit scatter the host memory system (~150 GB/s) becomes the bound; so the default leaves the host cores alone.
  int packRays = 0;
  int packPageable = 1;                   // pageable caller memory is always packed by the host pool (see traceStreamCompact)
  int packDepth = 2;                      // hybrid: chunks allowed in the pack stage at once
  void* packHost[kRing] = {nullptr, nullptr, nullptr, nullptr};
  size_t packCap[kRing] = {0, 0, 0, 0};
  std::atomic<unsigned long long> h2dBytes{0}, d2hBytes{0};   // bytes moved over PCIe by staged queries (rtcxGetTransferBytes)
  // staging contexts for short host streams: one per concurrent caller, recycled through a free list
  struct SmallStage { cudaStream_t stream = nullptr; void* buf = nullptr; size_t cap = 0; unsigned int* work = nullptr; };
  std::vector<SmallStage*> smallFree, smallAll; std::mutex smallMutex;
  SmallStage* acquireSmallStage() {
    {
      std::lock_guard<std::mutex> l(smallMutex);
      if (!smallFree.empty()) { SmallStage* s = smallFree.back(); smallFree.pop_back(); return s; }
    }
    SmallStage* s = new SmallStage();
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc((void**)&s->work, 32) != cudaSuccess) {
      if (s->stream) cudaStreamDestroy(s->stream);
      delete s; cudaGetLastError();
      throw rtc_error(RTC_ERROR_OUT_OF_MEMORY, "cannot create a staging context");
    }
    std::lock_guard<std::mutex> l(smallMutex);
    smallAll.push_back(s);
    return s;
  }
  void releaseSmallStage(SmallStage* s) { std::lock_guard<std::mutex> l(smallMutex); smallFree.push_back(s); }
  HostPool* pool = nullptr; std::mutex poolMutex;
  HostPool& hostPool() {
    std::lock_guard<std::mutex> l(poolMutex);
    if (!pool) {
      int n = scatterThreads > 0 ? scatterThreads : (int)std::thread::hardware_concurrency();
      if (packRays && scatterThreads == 8) n = (int)std::thread::hardware_concurrency();   // packing wants every core
      pool = new HostPool(std::max(1, std::min(n, 16)));
    }
    return *pool;
  }
  // staged upload of pageable geometry buffers at commit (stagedUpload below): page-locked chunks + the event of their last DMA
  static const int kGeoRing = 3; static const size_t kGeoChunk = (size_t)8 << 20;
  void* geoStage[kGeoRing] = {nullptr, nullptr, nullptr}; cudaEvent_t geoEvent[kGeoRing] = {nullptr, nullptr, nullptr};
  bool geoBusy[kGeoRing] = {false, false, false};   // the slot's last DMA may still be in flight (its event is waited for before the slot is refilled)
  size_t geoNext = 0;                               // ring position of the next chunk (continues across uploads)
  std::mutex geoMutex;
  int stageGeometry = 32;                 // stage_geometry=<MB>: smallest pageable buffer that takes the staged route (0 = never: plain cudaMemcpyAsync from the caller's pages)
  int refitEnabled = 1;                   // refit=0: RTC_BUILD_QUALITY_REFIT geometries are rebuilt like any other
  // Multi-GPU ("gpus=N"): this device object drives GPU `ordinal`; every further GPU is a peer device object of its
  // own (own streams, staging rings, host pool).  A commit builds here and replicates the flat image to the peers over
  // NVLink (cudaMemcpyPeerAsync); a host-resident stream is cut into one contiguous shard per GPU, traced concurrently, and
  // every GPU lands its hits in the caller's buffer -- broadcast and gather are the only inter-GPU steps.
  int numGpus = 1;
  std::vector<Device*> peers;             // owned
  unsigned shardMinRays = 1u << 20;       // shorter host streams stay on the primary GPU

  cudaStream_t stream() const { return userStream ? userStream : ownStream; }
  void bind() const { if (hasGpu) cudaSetDevice(ordinal); }

  ~Device() override {
    for (Device* p : peers) p->release();
    delete pool;
    if (hasGpu) {
      cudaSetDevice(ordinal);
      for (int i = 0; i < kRing; i++) { if (ringBuf[i]) cudaFree(ringBuf[i]); if (ringStream[i]) cudaStreamDestroy(ringStream[i]); }
      for (int i = 0; i < kRing; i++) {
        if (listDev[i]) cudaFree(listDev[i]);
        if (listHost[i]) cudaFreeHost(listHost[i]);
        if (packHost[i]) cudaFreeHost(packHost[i]);
      }
      for (SmallStage* s : smallAll) { if (s->buf) cudaFree(s->buf); if (s->work) cudaFree(s->work); if (s->stream) cudaStreamDestroy(s->stream); delete s; }
      for (int i = 0; i < kGeoRing; i++) {
        if (geoBusy[i] && geoEvent[i]) cudaEventSynchronize(geoEvent[i]);   // a commit that failed half way may have left a DMA behind
        if (geoStage[i]) cudaFreeHost(geoStage[i]);
        if (geoEvent[i]) cudaEventDestroy(geoEvent[i]);
      }
      if (countHost) cudaFreeHost(countHost);
      if (countDev) cudaFree(countDev);
      if (dCounters) cudaFree(dCounters);
      if (dWork) cudaFree(dWork);
      if (ownStream) cudaStreamDestroy(ownStream);
    }
  }
};

void processError(Device* dev, RTCError code, const char* msg) {
  if (dev) {
    if (dev->verbose >= 1) fprintf(stderr, "b200-rayquery error %d: %s\n", (int)code, msg);
    RTCErrorFunction fn; void* p;
    {
      std::lock_guard<std::mutex> l(dev->errMutex);
      if (dev->error == RTC_ERROR_NONE) dev->error = code;      // first error wins (device.cpp:258-271)
      fn = dev->errFn; p = dev->errPtr;
    }
    if (fn) fn(p, code, msg);
  } else {
    if (g_threadError == RTC_ERROR_NONE) g_threadError = code;
  }
}

#define RTC_TRY try {
#define RTC_CATCH(dev)                                                                              \
  } catch (const rtc_error& e) { processError((dev), e.code, e.what());                             \
  } catch (const std::bad_alloc&) { processError((dev), RTC_ERROR_OUT_OF_MEMORY, "out of memory");  \
  } catch (const std::exception& e) { processError((dev), RTC_ERROR_UNKNOWN, e.what());             \
  } catch (...) { processError((dev), RTC_ERROR_UNKNOWN, "unknown exception caught"); }
#define VERIFY_HANDLE(h) do { if ((h) == nullptr) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid argument"); } while (0)

// key=value[,key=value...] -- same grammar as the reference's device config (state.cpp:256-449);
// unknown keys are ignored like the reference ignores keys of disabled features.
void parseConfig(Device* d, const char* cfg, bool* allowNoGpu) {
  if (!cfg) return;
  std::string s(cfg);
  size_t i = 0;
  while (i < s.size()) {
    size_t j = s.find_first_of(", \t\n", i);
    if (j == std::string::npos) j = s.size();
    std::string tok = s.substr(i, j - i);
    i = j + 1;
    size_t eq = tok.find('=');
    if (tok.empty() || eq == std::string::npos) continue;
    const std::string k = tok.substr(0, eq), v = tok.substr(eq + 1);
    if (k == "verbose") d->verbose = atoi(v.c_str());
    else if (k == "benchmark") d->benchmark = atoi(v.c_str());
    else if (k == "gpu") d->ordinal = atoi(v.c_str());
    else if (k == "gpus") d->numGpus = std::max(1, atoi(v.c_str()));
    else if (k == "shard_min_rays") d->shardMinRays = (unsigned)std::max(1ll, atoll(v.c_str()));
    else if (k == "async") d->async = atoi(v.c_str());
    else if (k == "chunk_rays") d->chunkRays = (size_t)std::max(1024ll, atoll(v.c_str()));
    else if (k == "cost_node") d->build.costNode = (float)atof(v.c_str());
    else if (k == "cost_tri") d->build.costTri = (float)atof(v.c_str());
    else if (k == "leaf_tris") d->build.maxLeafTris = atoi(v.c_str());
    else if (k == "gpu_builder") d->build.builder = (v == "sah") ? 2 : (v == "ploc") ? 1 : 0;
    else if (k == "morton_cubic") d->build.mortonCubic = atoi(v.c_str()) != 0;
    else if (k == "treelet") d->build.treeletSize = atoi(v.c_str()) >= 512 ? 512 : 256;
    else if (k == "sweep_bottom") d->build.sweepBottom = atoi(v.c_str()) != 0;
    else if (k == "presplit") d->build.presplit = atoi(v.c_str()) != 0;
    else if (k == "ploc_radius") d->build.plocRadius = atoi(v.c_str());
    else if (k == "split_closest") d->splitClosest = atoi(v.c_str());
    else if (k == "stack_smem") d->stackSmem = std::max(0, std::min(16, atoi(v.c_str())));
    else if (k == "zerocopy") d->zeroCopy = atoi(v.c_str());
    else if (k == "refit") d->refitEnabled = atoi(v.c_str());
    else if (k == "d2h") d->d2hMode = atoi(v.c_str());
    else if (k == "compact_min_rays") d->compactMinRays = (unsigned)std::max(0ll, atoll(v.c_str()));
    else if (k == "scatter_threads" || k == "host_threads") d->scatterThreads = atoi(v.c_str());
    else if (k == "stage_geometry") d->stageGeometry = atoi(v.c_str());
    else if (k == "pack_rays") d->packRays = atoi(v.c_str());
    else if (k == "pack_pageable") d->packPageable = atoi(v.c_str());
    else if (k == "pack_depth") d->packDepth = std::max(1, atoi(v.c_str()));
    else if (k == "tvote") d->tVote = std::max(0, std::min(32, atoi(v.c_str())));
    else if (k == "split_occluded") d->splitOccluded = atoi(v.c_str());
    else if (k == "refill") d->refillClosest = std::max(1, std::min(32, atoi(v.c_str())));
    else if (k == "refill_coherent") d->refillCoherent = std::max(1, std::min(32, atoi(v.c_str())));
    else if (k == "refill_occluded") d->refillOccluded = std::max(1, std::min(32, atoi(v.c_str())));
    else if (k == "split_coherent") d->splitCoherent = atoi(v.c_str());
    else if (k == "allow_no_gpu") *allowNoGpu = atoi(v.c_str()) != 0;
  }
  d->build.verbose = d->verbose;
}

// --------------------------------------------------------------------------------------------
struct Buffer : RefCounted {
  Device* dev; char* ptr; size_t bytes; bool shared;
  Buffer(Device* d, size_t n, void* sharedPtr) : dev(d), ptr((char*)sharedPtr), bytes(n), shared(sharedPtr != nullptr) {
    dev->retain();
    if (!shared) {
      if (dev->memFn && !dev->memFn(dev->memPtr, (ssize_t)bytes, false)) { dev->release(); fail(RTC_ERROR_OUT_OF_MEMORY, "memory monitor forced termination"); }
      if (posix_memalign((void**)&ptr, 64, bytes ? bytes : 64) != 0) { dev->release(); throw std::bad_alloc(); }
    }
  }
  ~Buffer() override {
    if (!shared) { free(ptr); if (dev->memFn) dev->memFn(dev->memPtr, -(ssize_t)bytes, true); }
    dev->release();
  }
};

struct BufferView {
  Buffer* buf = nullptr; size_t offset = 0, stride = 0; unsigned count = 0; RTCFormat format = RTC_FORMAT_UNDEFINED;
  void set(Buffer* b, size_t off, size_t st, unsigned n, RTCFormat f) {
    b->retain(); if (buf) buf->release();
    buf = b; offset = off; stride = st; count = n; format = f;
  }
  void clear() { if (buf) buf->release(); buf = nullptr; }
  const char* data() const { return buf ? buf->ptr + offset : nullptr; }
};

struct Scene;
struct Geometry : RefCounted {
  Device* dev; RTCGeometryType type;
  // RTC_GEOMETRY_TYPE_INSTANCE (kernels/common/scene_instance.h): the instanced scene and local2world,
  // column major (vx, vy, vz, p) like AffineSpace3fa
  Scene* instanced = nullptr;
  float l2w[12] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
  BufferView vertices, indices;
  bool committed = false, enabled = true;
  unsigned modCounter = 0, mask = 0xFFFFFFFFu;
  unsigned topoCounter = 0;                              // bumped by every change a refit cannot absorb (anything but new vertex positions)
  void* userPtr = nullptr;
  RTCBuildQuality quality = RTC_BUILD_QUALITY_MEDIUM;
  Geometry(Device* d, RTCGeometryType t) : dev(d), type(t) { dev->retain(); }
  ~Geometry() override;
  void update() { ++modCounter; ++topoCounter; committed = false; }
  void updateVertices() { ++modCounter; committed = false; }   // rtcUpdateGeometryBuffer(RTC_BUFFER_TYPE_VERTEX)
};

struct Scene : RefCounted {
  Device* dev;
  std::mutex geomMutex, buildMutex;
  std::vector<Geometry*> geoms;                          // index = geomID
  std::set<unsigned> freeIDs; unsigned nextID = 0;       // lowest-free-first (common/sys/alloc.h:100-125)
  std::vector<unsigned> seenMod, seenTopo;
  RTCSceneFlags flags = RTC_SCENE_FLAG_NONE; RTCBuildQuality quality = RTC_BUILD_QUALITY_MEDIUM;
  bool modified = true, everCommitted = false;
  RQDeviceImage image{nullptr, {}};
  bool accounted = false;                                // the image was reported to the memory monitor (built here, not adopted / replicated)
  RQInstance* dInstances = nullptr; unsigned numInstances = 0; unsigned traceDepth = 0;   // instance table of a scene with instance geometries
  unsigned long long epoch = 0;                          // bumped by every commit / image adoption
  std::vector<std::pair<Scene*, unsigned long long>> instancedEpochs;   // distinct instanced scenes (each retained) and the epoch their device pointers were taken at
  void clearInstanced() { for (auto& e : instancedEpochs) e.first->release(); instancedEpochs.clear(); }
  RQBuildStats stats{};
  std::vector<Scene*> peerScenes;                        // gpus=N: one replica scene per peer device (same index as Device::peers), owned
  bool replicated = false;                               // the replicas hold the image of the current commit
  RTCProgressMonitorFunction progress = nullptr; void* progressPtr = nullptr;
  explicit Scene(Device* d) : dev(d) { dev->retain(); memset(&stats, 0, sizeof(stats)); }
  ~Scene() override {
    for (Scene* p : peerScenes) if (p) p->release();
    for (Geometry* g : geoms) if (g) g->release();
    if (image.base) {
      dev->bind();
      const size_t bytes = (size_t)image.header.totalBytes;
      rqFreeImage(&image);
      if (accounted && dev->memFn) dev->memFn(dev->memPtr, -(ssize_t)bytes, true);
    }
    if (dInstances) { dev->bind(); cudaFree(dInstances); }
    clearInstanced();
    dev->release();
  }
};
Geometry::~Geometry() { vertices.clear(); indices.clear(); if (instanced) instanced->release(); dev->release(); }

bool isDevicePointer(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
// page-locked host memory (cudaHostAlloc / cudaHostRegister): returns the address the GPU can use for it, else NULL
void* mappedHostPointer(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// The application's memory monitor (rtcSetDeviceMemoryMonitorFunction) also sees the DEVICE memory of a commit: the uploaded
// geometry buffers, the build scratch and the BVH image (reference: every allocation of a build goes through
// Device::memoryMonitor, kernels/common/device.cpp:318-327; verify.cpp:4564-4634 vetoes a random one and checks the balance).
bool deviceMonitor(void* user, long long bytes, bool post) {
  Device* dev = (Device*)user;
  return !dev->memFn || dev->memFn(dev->memPtr, (ssize_t)bytes, post);
}
struct MonitorScope {
  explicit MonitorScope(Device* dev) { rqSetAllocMonitor(dev->memFn ? deviceMonitor : nullptr, dev); }
  ~MonitorScope() { rqSetAllocMonitor(nullptr, nullptr); }
};
void freeImage(Device* dev, RQDeviceImage* img) {
  if (!img->base) return;
  const size_t bytes = (size_t)img->header.totalBytes;
  rqFreeImage(img);
  if (dev->memFn) dev->memFn(dev->memPtr, -(ssize_t)bytes, true);
}

// Pageable geometry buffers (what a drop-in application attaches: malloc'ed vertex / index arrays).  cudaMemcpyAsync from pageable
// memory bounces through the driver's own staging buffer at ~12 GB/s and made the H2D of the 180 MB of the 10 M-triangle scene
// two thirds of rtcCommitScene's wall time (16 of 22 ms; the build kernels take 5.6).  Instead the library's host threads copy
// 8 MB chunks into a ring of page-locked buffers and the copy engine takes them from there, chunk k+1 being filled while chunk k is
// in flight.  The hand-offs to the thread pool cost ~1.2 ms per buffer, so the route pays from ~20 MB on (a 6 MB vertex buffer:
// 1.2 ms plain, 2.6 ms staged; 180 MB: 16 ms plain, 6 ms staged -- profiles/r02final_ab_stage_small.log, r02r_build_c3_*.jsonl);
// the default threshold is 32 MB (device option stage_geometry=<MB>, 0 = never).
void stagedUpload(Device* dev, char* dst, const char* src, size_t bytes, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(dev->geoMutex);
  for (int i = 0; i < Device::kGeoRing; i++) {
    if (!dev->geoStage[i]) {
      cudaCheck(cudaHostAlloc(&dev->geoStage[i], Device::kGeoChunk, cudaHostAllocDefault), "geometry staging buffer");
      cudaCheck(cudaEventCreateWithFlags(&dev->geoEvent[i], cudaEventDisableTiming), "geometry staging event");
    }
  }
  HostPool& pool = dev->hostPool();
  const size_t parts = std::max<size_t>(1, std::min<size_t>(pool.size(), 8));
  size_t k = 0;
  const auto tStart = std::chrono::steady_clock::now();
  double msCopy = 0.0, msRing = 0.0;
  for (size_t off = 0; off < bytes; off += Device::kGeoChunk, k++) {
    const int slot = (int)(dev->geoNext++ % Device::kGeoRing);
    const size_t n = std::min(Device::kGeoChunk, bytes - off);
    const auto t0 = std::chrono::steady_clock::now();
    if (dev->geoBusy[slot]) { cudaCheck(cudaEventSynchronize(dev->geoEvent[slot]), "geometry upload (ring)"); dev->geoBusy[slot] = false; }
    const auto t1 = std::chrono::steady_clock::now();
    std::mutex m; std::condition_variable cv; size_t left = parts;
    const size_t per = ((n + parts - 1) / parts + 63) & ~(size_t)63;
    for (size_t t = 0; t < parts; t++) {
      const size_t b = std::min(n, t * per), e = std::min(n, b + per);
      pool.submit([&, b, e, slot, off] {
        if (e > b) memcpy((char*)dev->geoStage[slot] + b, src + off + b, e - b);
        std::lock_guard<std::mutex> l(m);
        if (--left == 0) cv.notify_one();
      });
    }
    { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return left == 0; }); }
    const auto t2 = std::chrono::steady_clock::now();
    msRing += std::chrono::duration<double, std::milli>(t1 - t0).count(); msCopy += std::chrono::duration<double, std::milli>(t2 - t1).count();
    cudaCheck(cudaMemcpyAsync(dst + off, dev->geoStage[slot], n, cudaMemcpyHostToDevice, s), "geometry upload (copy)");
    cudaCheck(cudaEventRecord(dev->geoEvent[slot], s), "geometry upload (event)");
    dev->geoBusy[slot] = true;
  }
  // no drain here: the DMAs of the last chunks stay in flight behind this call (the commit synchronises its stream before it returns,
  // and a slot is refilled only after its event) -- waiting for them cost ~1 ms per buffer
  if (dev->verbose >= 2) {
    const auto tEnd = std::chrono::steady_clock::now();
    fprintf(stderr, "  staged upload %.1f MB: %.3f ms on the calling thread (host copies %.3f, ring waits %.3f)\n", bytes / 1e6,
            std::chrono::duration<double, std::milli>(tEnd - tStart).count(), msCopy, msRing);
  }
}

struct TempDev {                                         // device copies of host geometry buffers, freed after the build
  std::vector<void*> ptrs; std::vector<size_t> sizes; cudaStream_t stream = nullptr; Device* dev = nullptr;   // stream-ordered pool: no cudaMalloc/cudaFree stalls on re-commit
  ~TempDev() { for (size_t i = 0; i < ptrs.size(); i++) { cudaFreeAsync(ptrs[i], stream); if (dev && dev->memFn) dev->memFn(dev->memPtr, -(ssize_t)sizes[i], true); } }
  const uint8_t* upload(const char* src, size_t bytes, cudaStream_t s) {
    void* d = nullptr;
    stream = s;
    const size_t want = bytes ? bytes : 16;
    if (dev && dev->memFn && !dev->memFn(dev->memPtr, (ssize_t)want, false)) fail(RTC_ERROR_OUT_OF_MEMORY, "memory monitor forced termination");
    const int e = cudaMallocAsync(&d, want, s);
    if (e) { if (dev && dev->memFn) dev->memFn(dev->memPtr, -(ssize_t)want, true); cudaCheck(e, "geometry upload (alloc)"); }
    ptrs.push_back(d); sizes.push_back(want);
    if (dev && dev->stageGeometry > 0 && bytes >= ((size_t)dev->stageGeometry << 20) && !mappedHostPointer(src)) stagedUpload(dev, (char*)d, src, bytes, s);
    else if (bytes) cudaCheck(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, s), "geometry upload (copy)");
    return (const uint8_t*)d;
  }
};

// gpus=N: copy the committed image of `sc` to every peer GPU (collective 1 of 2: the NVLink broadcast of SURVEY 8e, here a
// fan-out of cudaMemcpyPeerAsync on the peers' own streams so the copies run concurrently through the switch).
void replicateScene(Scene* sc) {
  Device* dev = sc->dev;
  sc->replicated = false;
  sc->stats.msBroadcast = 0.f;
  if (dev->peers.empty()) return;
  if (sc->numInstances) return;                          // an instanced image refers to other scenes' device memory: queries stay on the primary GPU
  const RQImageHeader& H = sc->image.header;
  const size_t bytes = (size_t)H.totalBytes;
  if (sc->peerScenes.size() != dev->peers.size()) {
    for (Scene* p : sc->peerScenes) if (p) p->release();
    sc->peerScenes.clear();
    for (Device* pd : dev->peers) sc->peerScenes.push_back(new Scene(pd));
  }
  std::vector<void*> dst(dev->peers.size(), nullptr);
  auto t0 = std::chrono::steady_clock::now();
  try {
    for (size_t g = 0; g < dev->peers.size(); g++) {                // allocations first (the first commit grows the peers' pools: not transfer time)
      Device* pd = dev->peers[g];
      pd->bind();
      cudaCheck(rqAllocImage(&dst[g], bytes, (rqStream)pd->stream()), "replica alloc");
      cudaCheck(cudaStreamSynchronize(pd->stream()), "replica alloc");
    }
    t0 = std::chrono::steady_clock::now();
    for (size_t g = 0; g < dev->peers.size(); g++) {
      Device* pd = dev->peers[g];
      pd->bind();
      cudaCheck(cudaMemcpyPeerAsync(dst[g], pd->ordinal, sc->image.base, dev->ordinal, bytes, pd->stream()), "replica copy");
    }
    for (size_t g = 0; g < dev->peers.size(); g++) {
      Device* pd = dev->peers[g];
      pd->bind();
      cudaCheck(cudaStreamSynchronize(pd->stream()), "replica copy");
      Scene* ps = sc->peerScenes[g];
      if (ps->image.base) rqFreeImage(&ps->image);
      ps->image.base = dst[g]; dst[g] = nullptr; ps->image.header = H; ps->image.numLevels = 0;
      ps->flags = sc->flags; ps->stats = sc->stats; ps->modified = false; ps->everCommitted = true; ps->epoch++;
    }
  } catch (...) {
    for (size_t g = 0; g < dst.size(); g++) if (dst[g]) { dev->peers[g]->bind(); RQDeviceImage tmp; memset(&tmp, 0, sizeof(tmp)); tmp.base = dst[g]; rqFreeImage(&tmp); }
    dev->bind();
    throw;
  }
  dev->bind();
  sc->stats.msBroadcast = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  sc->replicated = true;
}

void commitScene(Scene* sc) {
  Device* dev = sc->dev;
  std::unique_lock<std::mutex> lock(sc->buildMutex);     // one committer at a time; joiners simply wait
  // Refit instead of rebuild (reference: RTC_BUILD_QUALITY_REFIT meshes of the two-level builder,
  // bvh_builder_twolevel.h:95-140 -> bvh_refit.cpp): the scene itself is unchanged (no attach /
  // detach / flags / quality), and every geometry touched since the last commit only got new vertex
  // positions (rtcUpdateGeometryBuffer on the vertex buffer) and has build quality REFIT.
  bool refit = dev->refitEnabled && !sc->modified && sc->everCommitted && sc->image.base && sc->image.numLevels > 0;
  {
    std::lock_guard<std::mutex> gl(sc->geomMutex);
    bool changed = sc->modified || !sc->everCommitted;
    if (sc->seenMod.size() != sc->geoms.size()) { sc->seenMod.resize(sc->geoms.size(), 0xFFFFFFFFu); changed = true; refit = false; }
    if (sc->seenTopo.size() != sc->geoms.size()) { sc->seenTopo.resize(sc->geoms.size(), 0xFFFFFFFFu); refit = false; }
    for (size_t i = 0; i < sc->geoms.size(); i++) {
      Geometry* g = sc->geoms[i];
      if (!g) continue;
      if (g->enabled && !g->committed) fail(RTC_ERROR_INVALID_OPERATION, "geometry not committed");   // geometry.cpp:103-107
      if (g->modCounter != sc->seenMod[i]) {
        changed = true;
        if (g->topoCounter != sc->seenTopo[i] || g->quality != RTC_BUILD_QUALITY_REFIT) refit = false;
      }
    }
    for (const auto& e : sc->instancedEpochs) if (e.first->epoch != e.second) { changed = true; refit = false; }   // an instanced scene was re-committed
    if (!changed) return;
  }
  if (!dev->hasGpu) fail(RTC_ERROR_UNKNOWN, "no CUDA device: cannot build (there is no CPU fallback)");
  dev->bind();
  cudaStream_t s = dev->stream();
  if (sc->progress && !sc->progress(sc->progressPtr, 0.0)) fail(RTC_ERROR_CANCELLED, "progress monitor forced termination");

  std::vector<RQGeomDesc> descs;
  std::vector<RQInstance> insts;
  unsigned instDepth = 0;
  TempDev tmp; tmp.dev = dev;
  MonitorScope monitor(dev);
  // the counters this commit is about to absorb: written back to the scene only after the build / refit has succeeded, so a
  // failed commit (out of memory, cancelled, invalid instance) is retried by the next rtcCommitScene instead of being skipped
  std::vector<unsigned> newMod, newTopo;
  {
    std::lock_guard<std::mutex> gl(sc->geomMutex);
    newMod = sc->seenMod; newTopo = sc->seenTopo;
    for (size_t i = 0; i < sc->geoms.size(); i++) {
      Geometry* g = sc->geoms[i];
      if (!g) continue;
      newMod[i] = g->modCounter; newTopo[i] = g->topoCounter;
      if (g->enabled && g->type == RTC_GEOMETRY_TYPE_INSTANCE) {
        // one primitive whose box is xfmBounds(local2world, bounds of the instanced scene) (scene_instance.h:61-66);
        // the instanced scene must be committed first, as in the reference
        Scene* in = g->instanced;
        if (!in) continue;
        if (in == sc) fail(RTC_ERROR_INVALID_OPERATION, "a scene cannot instantiate itself");
        if (!in->everCommitted || in->modified || !in->image.base) fail(RTC_ERROR_INVALID_OPERATION, "instanced scene not committed");
        if (in->numInstances) fail(RTC_ERROR_INVALID_OPERATION, "multi-level instancing is not supported (RTC_MAX_INSTANCE_LEVEL_COUNT = 1)");
        if (((in->flags ^ sc->flags) & RTC_SCENE_FLAG_ROBUST) != 0)
          fail(RTC_ERROR_INVALID_OPERATION, "an instanced scene must use the same RTC_SCENE_FLAG_ROBUST setting as the scene instantiating it");
        refit = false;
        RQInstance I; memset(&I, 0, sizeof(I));
        RQGeomDesc d; memset(&d, 0, sizeof(d));
        const float* m = g->l2w;
        // world2local = rcp(local2world) = (adjoint / det, -(l^-1 * p)) (linearspace3.h:44-50, affinespace.h rcp), evaluated
        // like the reference's lowest-ISA object does (scene_instance.cpp:131 is built for SSE2: every product and sum rounds
        // separately, dot = (x + y) + z, adjoint * rcp(det)).  Answers inside an instance move by |translation| * few ulp when
        // this is evaluated differently, so the same order is kept here.
        {
          volatile float t0, t1;                                // volatile: no FMA contraction whatever the host flags
          auto mulsub = [&](float a, float b, float c, float d) { t0 = a * b; t1 = c * d; return (float)(t0 - t1); };
          const float vx[3] = {m[0], m[1], m[2]}, vy[3] = {m[3], m[4], m[5]}, vz[3] = {m[6], m[7], m[8]}, p[3] = {m[9], m[10], m[11]};
          const float c0[3] = {mulsub(vy[1], vz[2], vy[2], vz[1]), mulsub(vy[2], vz[0], vy[0], vz[2]), mulsub(vy[0], vz[1], vy[1], vz[0])};   // cross(vy, vz)
          const float c1[3] = {mulsub(vz[1], vx[2], vz[2], vx[1]), mulsub(vz[2], vx[0], vz[0], vx[2]), mulsub(vz[0], vx[1], vz[1], vx[0])};   // cross(vz, vx)
          const float c2[3] = {mulsub(vx[1], vy[2], vx[2], vy[1]), mulsub(vx[2], vy[0], vx[0], vy[2]), mulsub(vx[0], vy[1], vx[1], vy[0])};   // cross(vx, vy)
          volatile float d0 = vx[0] * c0[0], d1 = vx[1] * c0[1], d2 = vx[2] * c0[2];
          volatile float ds = d0 + d1;
          const float det = ds + d2;
          const float rd = 1.0f / det;
          volatile float il[9];                                 // column k of the inverse = (c0[k], c1[k], c2[k]) * rcp(det)
          for (int k = 0; k < 3; k++) { il[3 * k + 0] = c0[k] * rd; il[3 * k + 1] = c1[k] * rd; il[3 * k + 2] = c2[k] * rd; }
          for (int k = 0; k < 9; k++) I.w2l[k] = il[k];
          for (int r = 0; r < 3; r++) {
            volatile float a = p[2] * il[6 + r], b2 = p[1] * il[3 + r], c = p[0] * il[r];
            volatile float s1 = a + b2;
            I.w2l[9 + r] = -(s1 + c);
          }
        }
        const RQImageHeader& IH = in->image.header;
        I.nodes = (uint64_t)((char*)in->image.base + IH.nodesOffset);
        I.tris = (uint64_t)((char*)in->image.base + IH.trisOffset);
        I.geomID = (uint32_t)i; I.depth = IH.depth;
        instDepth = std::max(instDepth, (unsigned)IH.depth);
        // xfmBounds: the 8 corners through xfmPoint with the reference's FMA nesting (affinespace.h:102-118)
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int c = 0; c < 8; c++) {
          const float x = (c & 4) ? IH.hi[0] : IH.lo[0], y = (c & 2) ? IH.hi[1] : IH.lo[1], z = (c & 1) ? IH.hi[2] : IH.lo[2];
          for (int r = 0; r < 3; r++) {
            const float v = fmaf(x, m[r], fmaf(y, m[3 + r], fmaf(z, m[6 + r], m[9 + r])));
            lo[r] = fminf(lo[r], v); hi[r] = fmaxf(hi[r], v);      // NaN corners are ignored; an empty instanced scene yields an inverted box
          }
        }
        d.type = 1; d.numTris = 1; d.geomID = (uint32_t)i; d.instIndex = (uint32_t)insts.size();
        for (int r = 0; r < 3; r++) { d.lo[r] = lo[r]; d.hi[r] = hi[r]; }
        insts.push_back(I);
        descs.push_back(d);
        continue;
      }
      if (!g->enabled || g->indices.count == 0) continue;
      RQGeomDesc d; memset(&d, 0, sizeof(d));
      const char* ip = g->indices.data(); const char* vp = g->vertices.data();
      const bool quad = g->type == RTC_GEOMETRY_TYPE_QUAD;   // two triangles per quad: (v0,v1,v3), (v2,v3,v1)
      const size_t ibytes = (size_t)(g->indices.count - 1) * g->indices.stride + (quad ? 16 : 12);
      const size_t vbytes = g->vertices.count ? (size_t)(g->vertices.count - 1) * g->vertices.stride + 12 : 0;
      d.indices = isDevicePointer(ip) ? (const uint8_t*)ip : tmp.upload(ip, ibytes, s);
      d.vertices = (vp && isDevicePointer(vp)) ? (const uint8_t*)vp : tmp.upload(vp, vbytes, s);
      d.indexStride = (uint32_t)g->indices.stride; d.vertexStride = (uint32_t)g->vertices.stride;
      d.numTris = quad ? 2u * g->indices.count : g->indices.count; d.numVerts = g->vertices.count; d.geomID = (uint32_t)i;
      d.type = quad ? 2u : 0u;
      descs.push_back(d);
    }
  }
  if (refit) {
    std::vector<RQGeomDesc> byID(sc->geoms.size());
    memset(byID.data(), 0, sizeof(RQGeomDesc) * byID.size());
    for (const RQGeomDesc& d : descs) byID[d.geomID] = d;
    cudaCheck(rqRefitBVH(byID.data(), (int)byID.size(), &sc->image, (rqStream)s, &sc->stats), "BVH refit");
  } else {
    RQDeviceImage img; memset(&img, 0, sizeof(img));
    RQBuildStats st;
    RQBuildParams bp = dev->build;
    // scene build quality: LOW = the fast radix-tree front end (role of the reference's Morton builder for
    // RTC_BUILD_QUALITY_LOW, scene.cpp:118-124), HIGH = large triangles pre-split into clipped references (role of the
    // reference's spatial-split builder, bvh_builder_sah_spatial.cpp), 512-triangle treelets with an exact sweep at the
    // bottom, PLOC search radius >= 16
    if (!insts.empty()) bp.maxLeafTris = 1;                  // every instance gets its own child box: entering one costs a ray transform + a root fetch
    if (sc->quality == RTC_BUILD_QUALITY_LOW) bp.builder = 0;
    else if (sc->quality == RTC_BUILD_QUALITY_HIGH) { bp.plocRadius = std::max(bp.plocRadius, 16); bp.treeletSize = 512; bp.sweepBottom = 1; bp.presplit = 1; if (bp.builder == 0) bp.builder = 2; }
    int be = rqBuildBVH(descs.data(), (int)descs.size(), (uint32_t)sc->flags, &bp, (rqStream)s, &img, &st);
    if (be == RQ_BUILD_STALLED) {                            // PLOC made no progress on this input: the radix tree always terminates
      if (dev->verbose >= 1) fprintf(stderr, "b200-rayquery: PLOC stage stalled, rebuilding with the radix-tree front end\n");
      bp.builder = 0;
      be = rqBuildBVH(descs.data(), (int)descs.size(), (uint32_t)sc->flags, &bp, (rqStream)s, &img, &st);
    }
    cudaCheck(be, "BVH build");
    if (sc->image.base) { if (sc->accounted) freeImage(dev, &sc->image); else rqFreeImage(&sc->image); }
    sc->image = img; sc->stats = st; sc->accounted = true;
    // instance table (traversal reads it through TraceParams::instances)
    if (sc->dInstances) { cudaFree(sc->dInstances); sc->dInstances = nullptr; }
    sc->numInstances = (unsigned)insts.size();
    sc->clearInstanced();
    {
      std::lock_guard<std::mutex> gl(sc->geomMutex);
      for (Geometry* g : sc->geoms)
        if (g && g->enabled && g->type == RTC_GEOMETRY_TYPE_INSTANCE && g->instanced) {
          bool seen = false;
          for (auto& e : sc->instancedEpochs) seen |= (e.first == g->instanced);
          if (!seen) { g->instanced->retain(); sc->instancedEpochs.push_back({g->instanced, g->instanced->epoch}); }
        }
    }
    sc->traceDepth = img.header.depth;
    if (!insts.empty()) {
      cudaCheck(cudaMalloc((void**)&sc->dInstances, insts.size() * sizeof(RQInstance)), "instance table");
      cudaCheck(cudaMemcpyAsync(sc->dInstances, insts.data(), insts.size() * sizeof(RQInstance), cudaMemcpyHostToDevice, s), "instance table");
      cudaCheck(cudaStreamSynchronize(s), "instance table");
      sc->traceDepth = img.header.depth + 3 + instDepth;       // the lane parks 3 entries of top-level state while inside an instance
    }
  }
  replicateScene(sc);                                    // gpus=N: the peers' replicas follow every build and refit
  {
    std::lock_guard<std::mutex> gl(sc->geomMutex);
    for (size_t i = 0; i < newMod.size() && i < sc->seenMod.size(); i++) { sc->seenMod[i] = newMod[i]; sc->seenTopo[i] = newTopo[i]; }
  }
  sc->modified = false; sc->everCommitted = true; sc->epoch++;
  if (sc->progress) sc->progress(sc->progressPtr, 1.0);
  if (dev->benchmark || dev->verbose >= 2) {
    // same fields as the reference's line (bvh.cpp:173-178): seconds, prims/s, SAH, bytes
    const RQBuildStats& st = sc->stats;
    printf("BENCHMARK_BUILD %g %g %g %llu BVH8q<triangle>.b200\n", st.msTotal * 1e-3, st.numPrimsValid / (st.msTotal * 1e-3 + 1e-12),
           st.sah, (unsigned long long)st.bytes);
    fflush(stdout);
  }
}

// --------------------------------------------------------------------------------------------
// ray-stream dispatch
// --------------------------------------------------------------------------------------------
void checkQuery(Scene* sc, RTCIntersectContext* ctx) {
  if (!sc->everCommitted || sc->modified) fail(RTC_ERROR_INVALID_OPERATION, "scene not committed");   // scene.cpp:13,30
  if (ctx && ctx->filter) fail(RTC_ERROR_INVALID_OPERATION, "filter callbacks cannot run on the GPU");
  // the instance table holds device pointers into the instanced scenes' images: a re-commit there invalidates them
  for (const auto& e : sc->instancedEpochs)
    if (e.first->epoch != e.second || e.first->modified)
      fail(RTC_ERROR_INVALID_OPERATION, "an instanced scene changed: commit the instantiating scene again");
}

void fillArgs(Scene* sc, RTCIntersectContext* ctx, RQTraceArgs& a, bool stream) {
  memset(&a, 0, sizeof(a));
  a.image = sc->image.base; a.nodesOffset = sc->image.header.nodesOffset; a.trisOffset = sc->image.header.trisOffset;
  a.compact = sc->image.header.layout == 1u ? 1u : 0u; a.metaOffset = sc->image.header.metaOffset; a.vertsOffset = sc->image.header.vertsOffset;
  a.depth = sc->numInstances ? sc->traceDepth : sc->image.header.depth;
  a.instances = sc->numInstances ? sc->dInstances : nullptr;
  a.robust = (sc->flags & RTC_SCENE_FLAG_ROBUST) ? 1u : 0u;
  a.instID0 = ctx ? ctx->instID[0] : RTC_INVALID_GEOMETRY_ID;
  a.streamSemantics = stream ? 1u : 0u;
}

// Host-staged stream, compact both ways (device option d2h=3, the default).
//   in : a pool of host threads packs the 32 bytes of every record the kernels read (org, tnear, dir, tfar) into a page-locked
//        staging buffer (streaming stores), one linear H2D per 1 M-ray chunk moves 32 instead of 80 / 48 bytes per ray
//        (pack_rays=0: the whole span is uploaded as it is);
//   out: the kernel appends one 48-byte (closest) / 4-byte (occluded) record per ray that hit to a list; only the list is
//        downloaded and the pool scatters it into the caller's records while later chunks are in flight.
// Why: the link is the bound of the end-to-end path.  Whole spans both ways: 47-49 GB/s per direction (tools/pcie_probe.py),
// 662-668 Mrays/s on configs[1]; list download: 804 Mrays/s (profiles/r01p2_ab_compact_pool.log); the host packs 80 -> 32 bytes
// at 84 (8 threads) - 117 GB/s (16 threads) of source bytes on the B200 box's Xeon (tools/hostpack_probe.c).
// The caller's memory is only touched by CPU loads / stores here, so it need not be page-locked.
inline void packRay32(const char* s, char* d) {
#if defined(__SSE4_1__) || defined(__AVX__)
  const __m128 a = _mm_loadu_ps((const float*)s);
  __m128 c = _mm_loadu_ps((const float*)(s + 16));
  c = _mm_insert_ps(c, _mm_load_ss((const float*)(s + 32)), 0x30);   // dir.xyz, tfar
  _mm_stream_ps((float*)d, a); _mm_stream_ps((float*)(d + 16), c);
#elif defined(__SSE2__)
  float t[8]; memcpy(t, s, 28); memcpy(t + 7, s + 32, 4);
  _mm_stream_ps((float*)d, _mm_loadu_ps(t)); _mm_stream_ps((float*)(d + 16), _mm_loadu_ps(t + 4));
#else
  memcpy(d, s, 28); memcpy(d + 28, s + 32, 4);
#endif
}

// The call is event driven -- no thread ever waits for the GPU except the caller waiting for a free ring slot:
//   caller      : per 1 M-ray chunk: take a free slot, hand the chunk's pack slices to the pool
//   last slice  : enqueues H2D + kernel + count download on the slot's stream, then a host callback
//   callback 1  : (CUDA thread, no CUDA calls allowed) asks the pool to fetch the list: count is known now
//   pool        : enqueues the list download + callback 2
//   callback 2  : hands the scatter slices to the pool; the last slice frees the slot
// (An event-polling progress thread in place of the two callbacks was tried: same timings, and it starves a one-thread pool.)
struct CompactCall;
struct CompactChunk {
  CompactCall* call; int slot; char* h; unsigned n; unsigned count; unsigned packLeft; unsigned scatterLeft; unsigned packParts; bool packed;
  double tIssue = 0, tEnq = 0, tCount = 0, tList = 0, tDone = 0; // host clock, ms since the call started (printed at verbose >= 3)
};
struct CompactCall {
  Device* dev; RQTraceArgs a; size_t stride; bool occluded, pack; size_t recBytes, recList; unsigned T;
  std::mutex m; std::condition_variable cv;                     // guards everything below and the chunks' counters
  bool slotBusy[Device::kRing] = {false, false, false, false};
  int pending = 0; unsigned chunksDone = 0; int error = 0;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double now() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
  int packing = 0;                                              // chunks whose pack slices are still queued or running
  void submit(std::function<void()> fn) {
    { std::lock_guard<std::mutex> lk(m); pending++; }
    dev->hostPool().submit([this, fn] { fn(); std::lock_guard<std::mutex> lk(m); if (--pending == 0) cv.notify_all(); });
  }
  void fail(int e) { std::lock_guard<std::mutex> lk(m); if (!error) error = e ? e : (int)cudaErrorUnknown; cudaGetLastError(); }
  void chunkDone(CompactChunk* c) { c->tDone = now(); std::lock_guard<std::mutex> lk(m); slotBusy[c->slot] = false; chunksDone++; cv.notify_all(); }

  void packSlice(CompactChunk* c, unsigned p) {
    const size_t b0 = (size_t)c->n * p / c->packParts, e0 = (size_t)c->n * (p + 1) / c->packParts;
    char* d = (char*)dev->packHost[c->slot];
    for (size_t i = b0; i < e0; i++) packRay32(c->h + i * stride, d + i * 32);
#if defined(__SSE2__)
    _mm_sfence();
#endif
    bool last;
    { std::lock_guard<std::mutex> lk(m); last = (--c->packLeft == 0); if (last) { packing--; cv.notify_all(); } }
    if (last) enqueue(c);
  }
  static void CUDART_CB onCount(void* p) { CompactChunk* c = (CompactChunk*)p; c->tCount = c->call->now(); c->call->submit([c] { c->call->fetchList(c); }); }
  static void CUDART_CB onList(void* p) { CompactChunk* c = (CompactChunk*)p; c->tList = c->call->now(); c->call->startScatter(c); }

  void enqueue(CompactChunk* c) {                               // any thread
    cudaSetDevice(dev->ordinal);
    const int r = c->slot;
    cudaStream_t s = dev->ringStream[r];
    RQTraceArgs x = a;
    int e = 0;
    if (c->packed) {
      e = cudaMemcpyAsync(dev->ringBuf[r], dev->packHost[r], (size_t)c->n * 32, cudaMemcpyHostToDevice, s);
      dev->h2dBytes += (unsigned long long)c->n * 32;
      x.stride = 32; x.packed = 1;
    } else {
      const size_t span = (size_t)(c->n - 1) * stride + recBytes;
      e = cudaMemcpyAsync(dev->ringBuf[r], c->h, span, cudaMemcpyHostToDevice, s);
      dev->h2dBytes += span;
      x.stride = stride; x.packed = 0;
    }
    x.rays = dev->ringBuf[r]; x.out = nullptr; x.numRays = c->n;
    x.workCounter = dev->dWork + 8 * r;
    x.hitList = dev->listDev[r]; x.hitCount = dev->countDev + 8 * r;
    if (!e) e = occluded ? rqLaunchOccluded(&x, (rqStream)s) : rqLaunchIntersect(&x, (rqStream)s);
    if (!e) e = cudaMemcpyAsync(&dev->countHost[r], x.hitCount, sizeof(unsigned), cudaMemcpyDeviceToHost, s);
    if (!e) e = cudaLaunchHostFunc(s, onCount, c);
    c->tEnq = now();
    if (e) { fail(e); cudaStreamSynchronize(s); cudaGetLastError(); chunkDone(c); }   // whatever was queued on the slot is finished before it is reused
  }
  void fetchList(CompactChunk* c) {                             // pool thread, the kernel of this chunk has finished
    cudaSetDevice(dev->ordinal);
    const int r = c->slot;
    cudaStream_t s = dev->ringStream[r];
    c->count = std::min(dev->countHost[r], c->n);
    int e = 0;
    if (c->count) {
      e = cudaMemcpyAsync(dev->listHost[r], dev->listDev[r], (size_t)c->count * recList, cudaMemcpyDeviceToHost, s);
      dev->d2hBytes += (unsigned long long)c->count * recList;
    }
    dev->d2hBytes += sizeof(unsigned);
    if (!e) e = cudaLaunchHostFunc(s, onList, c);
    if (e) { fail(e); cudaStreamSynchronize(s); cudaGetLastError(); chunkDone(c); }
  }
  void startScatter(CompactChunk* c) {                          // CUDA callback thread: only queues work
    const unsigned per = std::max(8192u, (c->count + T - 1) / T);
    const unsigned parts = std::max(1u, (c->count + per - 1) / per);
    { std::lock_guard<std::mutex> lk(m); c->scatterLeft = parts; }
    for (unsigned p = 0; p < parts; p++) submit([this, c, per, p] { scatterSlice(c, p * per, std::min(c->count, (p + 1) * per)); });
  }
  void scatterSlice(CompactChunk* c, unsigned begin, unsigned end) {
    if (occluded) {
      const uint32_t* ids = (const uint32_t*)dev->listHost[c->slot];
      for (unsigned k = begin; k < end; k++) {
        if (k + 16 < end && ids[k + 16] < c->n) __builtin_prefetch(c->h + (size_t)ids[k + 16] * stride + 32, 1);
        if (ids[k] < c->n) *(float*)(c->h + (size_t)ids[k] * stride + 32) = -INFINITY;
      }
    } else {
      const char* recs = (const char*)dev->listHost[c->slot];
      for (unsigned k = begin; k < end; k++) {
        if (k + 16 < end) {
          uint32_t nid; memcpy(&nid, recs + (size_t)(k + 16) * 48, 4);
          if (nid < c->n) { char* nd = c->h + (size_t)nid * stride; __builtin_prefetch(nd + 32, 1); __builtin_prefetch(nd + 79, 1); }
        }
        const char* rec = recs + (size_t)k * 48;
        uint32_t rid; memcpy(&rid, rec, 4);
        if (rid >= c->n) continue;
        char* dst = c->h + (size_t)rid * stride;
        memcpy(dst + 32, rec + 4, 4);                           // tfar
        memcpy(dst + 48, rec + 16, 32);                         // Ng, u, v, primID, geomID, instID[0]
      }
    }
    bool last;
    { std::lock_guard<std::mutex> lk(m); last = (--c->scatterLeft == 0); }
    if (last) chunkDone(c);
  }
};

void traceStreamCompact(Device* dev, RQTraceArgs a, char* rays, unsigned M, size_t stride, bool occluded, size_t recBytes) {
  std::lock_guard<std::mutex> l(dev->stageMutex);
  // 1 M-ray chunks for long streams; a short stream is cut into ~8 pieces so that upload, kernel, list download and scatter
  // of neighbouring pieces still overlap (configs[0], 1 M rays: one chunk = 239 Mrays/s end to end, profiles/r01x_c1.json)
  size_t chunk = dev->chunkRays;
  if ((size_t)M < 8 * chunk) chunk = std::max<size_t>(65536, (((size_t)M + 7) / 8 + 32767) & ~(size_t)32767);
  if (chunk > dev->chunkRays) chunk = dev->chunkRays;
  if (!dev->countHost) cudaCheck(cudaMallocHost((void**)&dev->countHost, sizeof(unsigned) * Device::kRing), "hit counters");
  if (!dev->countDev) cudaCheck(cudaMalloc((void**)&dev->countDev, 32 * Device::kRing), "hit counters");
  for (int r = 0; r < Device::kRing; r++)
    if (!dev->ringStream[r]) cudaCheck(cudaStreamCreateWithFlags(&dev->ringStream[r], cudaStreamNonBlocking), "stream");
  CompactCall call;
  // Pageable caller memory (what a drop-in application hands over: malloc'ed rays): cudaMemcpyAsync would bounce it through the
  // driver's own small staging buffer at ~12 GB/s, synchronously (measured 154 Mrays/s end to end against 693 from page-locked
  // memory, profiles/r02a_bench.json).  The host pool packs it instead -- CPU loads from the caller's pages, streaming stores of the
  // 32 bytes per ray the kernels read into page-locked staging -- so the link carries 32 instead of 80 / 48 bytes per ray.
  const bool pageable = dev->packPageable && mappedHostPointer(rays) == nullptr;
  call.dev = dev; call.a = a; call.stride = stride; call.occluded = occluded; call.pack = dev->packRays != 0 || pageable;
  const bool packAll = dev->packRays >= 2 || pageable;
  call.recBytes = recBytes; call.recList = occluded ? 4 : 48; call.T = (unsigned)dev->hostPool().size();
  const unsigned numChunks = (unsigned)((M + chunk - 1) / chunk);
  std::vector<CompactChunk> chunks(numChunks);
  using clk = std::chrono::steady_clock;
  const auto tCall = clk::now();
  double tSlot = 0;
  unsigned issued = 0;
  auto finish = [&] {                                           // every issued chunk done and no closure of this call left
    std::unique_lock<std::mutex> lk(call.m);
    call.cv.wait(lk, [&] { return call.chunksDone == issued && call.pending == 0; });
  };
  try {
    unsigned done = 0;
    for (unsigned i = 0; i < numChunks; i++) {
      const unsigned n = (unsigned)std::min<size_t>(chunk, M - done);
      const size_t span = (size_t)(n - 1) * stride + recBytes;
      const int r = (int)(i % Device::kRing);
      const auto t0 = clk::now();
      { std::unique_lock<std::mutex> lk(call.m); call.cv.wait(lk, [&] { return !call.slotBusy[r]; }); }
      tSlot += std::chrono::duration<double, std::milli>(clk::now() - t0).count();
      cudaStream_t s = dev->ringStream[r];                      // the slot is idle: its buffers may be re-allocated
      if (dev->ringCap[r] < span) {
        cudaCheck(cudaStreamSynchronize(s), "staging");
        if (dev->ringBuf[r]) cudaFree(dev->ringBuf[r]);
        dev->ringBuf[r] = nullptr; dev->ringCap[r] = 0;
        cudaCheck(cudaMalloc(&dev->ringBuf[r], span + 256), "staging alloc");
        dev->ringCap[r] = span;
      }
      if (dev->listCap[r] < (size_t)n * 48) {
        cudaCheck(cudaStreamSynchronize(s), "staging");
        if (dev->listDev[r]) cudaFree(dev->listDev[r]);
        if (dev->listHost[r]) cudaFreeHost(dev->listHost[r]);
        dev->listDev[r] = nullptr; dev->listHost[r] = nullptr; dev->listCap[r] = 0;
        cudaCheck(cudaMalloc(&dev->listDev[r], (size_t)n * 48), "hit list alloc");
        cudaCheck(cudaMallocHost(&dev->listHost[r], (size_t)n * 48), "hit list alloc");
        dev->listCap[r] = (size_t)n * 48;
      }
      if (call.pack && dev->packCap[r] < (size_t)n * 32) {
        cudaCheck(cudaStreamSynchronize(s), "staging");
        if (dev->packHost[r]) cudaFreeHost(dev->packHost[r]);
        dev->packHost[r] = nullptr; dev->packCap[r] = 0;
        cudaCheck(cudaMallocHost(&dev->packHost[r], (size_t)n * 32 + 64), "pack staging alloc");
        dev->packCap[r] = (size_t)n * 32;
      }
      CompactChunk* c = &chunks[i];
      c->call = &call; c->slot = r; c->h = rays + (size_t)done * stride; c->n = n; c->count = 0; c->scatterLeft = 0;
      // Hybrid upload: the copy engine moves whole records at ~55 GB/s without any CPU work, the pool packs at ~50 GB/s of source
      // bytes on this box (page-locked 4 KB pages) and then uploads 2.5x fewer bytes -- the two run side by side.  A chunk is
      // packed while fewer than packDepth chunks are in the pack stage, otherwise it goes to the copy engine as it is.
      {
        std::lock_guard<std::mutex> lk(call.m);
        c->packed = call.pack && (packAll || call.packing < dev->packDepth);
        if (c->packed) call.packing++;
        call.slotBusy[r] = true;
      }
      c->packParts = c->packed ? std::max(1u, std::min(2 * call.T, (n + 16383u) / 16384u)) : 1u;
      c->packLeft = c->packParts;
      c->tIssue = call.now();
      issued++;
      if (c->packed) { for (unsigned p = 0; p < c->packParts; p++) call.submit([&call, c, p] { call.packSlice(c, p); }); }
      else call.submit([&call, c] { call.enqueue(c); });
      done += n;
    }
  } catch (...) {
    finish();
    for (int r = 0; r < Device::kRing; r++) if (dev->ringStream[r]) cudaStreamSynchronize(dev->ringStream[r]);
    cudaGetLastError();
    throw;
  }
  finish();
  if (dev->verbose >= 2)
    fprintf(stderr, "b200-rayquery staged %s stream: %u rays in %u chunks, %.2f ms total, caller waited %.2f ms for ring slots; pool %u threads, pack %d\n",
            occluded ? "occlusion" : "closest-hit", M, numChunks, std::chrono::duration<double, std::milli>(clk::now() - tCall).count(), tSlot,
            call.T, (int)call.pack);
  if (dev->verbose >= 3)
    for (unsigned i = 0; i < issued; i++)
      fprintf(stderr, "  chunk %2u slot %d rays %7u hits %7u: issued %.3f, enqueued %.3f, count back %.3f, list back %.3f, scattered %.3f ms\n", i, chunks[i].slot,
              chunks[i].n, chunks[i].count, chunks[i].tIssue, chunks[i].tEnq, chunks[i].tCount, chunks[i].tList, chunks[i].tDone);
  if (call.error) {
    for (int r = 0; r < Device::kRing; r++) if (dev->ringStream[r]) cudaStreamSynchronize(dev->ringStream[r]);
    cudaGetLastError();
    cudaCheck(call.error, "staged trace");
  }
}

// Trace M records of `stride` bytes at `rays`; occluded selects the any-hit kernel; recBytes is
// 80 (RTCRayHit) or 48 (RTCRay).  Device-resident memory is traced in place; host memory is staged
// through a ring of device buffers so copies of one chunk overlap the kernel of another.
void traceStreamOn(Scene* sc, RTCIntersectContext* ctx, void* rays, unsigned M, size_t stride, bool occluded,
                   size_t recBytes, int streamRule, RQTraceCounters* countersOut = nullptr);

// streamRule: entry rules of occlusion rays -- 1 = stream filter (AoS / AoP streams with M > 1: rays with tnear < 0 are skipped,
// bvh_intersector_stream.cpp:303-305), 0 = single ray / packet (tnear is clamped to 0 instead, bvh_intersector1.cpp:132,
// bvh_intersector_hybrid.cpp:153,403), -1 = decide by M like rtcOccluded1M does.
void traceStream(Scene* sc, RTCIntersectContext* ctx, void* rays, unsigned M, size_t stride, bool occluded,
                 size_t recBytes, RQTraceCounters* countersOut, int streamRule = -1) {
  if (M == 0) return;
  Device* dev = sc->dev;
  if (!dev->hasGpu) fail(RTC_ERROR_UNKNOWN, "no CUDA device");
  // the kernels read records through 4-byte (16-byte when possible) words: a misaligned address would fault and poison the context
  if (((uintptr_t)rays & 3u) || (stride & 3u)) fail(RTC_ERROR_INVALID_ARGUMENT, "ray not aligned to 4 bytes");
  if (M > 1 && stride < recBytes) fail(RTC_ERROR_INVALID_OPERATION, "byteStride too small");   // overlapping records: hit writes would race with ray reads
  if (streamRule < 0) streamRule = M > 1 ? 1 : 0;
  if (sc->replicated && !countersOut) {
    cudaPointerAttributes pa;
    const bool known = cudaPointerGetAttributes(&pa, rays) == cudaSuccess;
    if (!known) cudaGetLastError();
    if (known && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged)) {
      // device-resident stream: traced where it lives (the replica of that GPU)
      for (size_t g = 0; g < dev->peers.size(); g++)
        if (pa.type == cudaMemoryTypeDevice && dev->peers[g]->ordinal == pa.device) {
          traceStream(sc->peerScenes[g], ctx, rays, M, stride, occluded, recBytes, nullptr, streamRule);
          dev->bind();
          return;
        }
    } else if (M >= dev->shardMinRays) {
      // host-resident stream: one contiguous shard per GPU (keeps whatever coherence the stream has), all shards in flight at
      // once, each GPU stages its shard through its own ring and lands its hits in the caller's records (collective 2 of 2)
      const unsigned G = 1u + (unsigned)dev->peers.size();
      std::vector<std::thread> th;
      std::vector<std::exception_ptr> errs(G);
      auto shard = [&](unsigned g) {
        const unsigned b = (unsigned)((unsigned long long)M * g / G), e = (unsigned)((unsigned long long)M * (g + 1) / G);
        if (e <= b) return;
        Scene* target = g == 0 ? sc : sc->peerScenes[g - 1];
        try {
          traceStreamOn(target, ctx, (char*)rays + (size_t)b * stride, e - b, stride, occluded, recBytes, streamRule);
        } catch (...) { errs[g] = std::current_exception(); }
      };
      for (unsigned g = 1; g < G; g++) th.emplace_back(shard, g);
      shard(0);
      for (auto& t : th) t.join();
      dev->bind();
      for (unsigned g = 0; g < G; g++) if (errs[g]) std::rethrow_exception(errs[g]);
      return;
    }
  }
  traceStreamOn(sc, ctx, rays, M, stride, occluded, recBytes, streamRule, countersOut);
}

// one stream on the GPU of sc->dev
void traceStreamOn(Scene* sc, RTCIntersectContext* ctx, void* rays, unsigned M, size_t stride, bool occluded,
                   size_t recBytes, int streamRule, RQTraceCounters* countersOut) {
  Device* dev = sc->dev;
  dev->bind();
  RQTraceArgs a; fillArgs(sc, ctx, a, streamRule != 0);
  RQTraceCounters* dC = nullptr;
  if (countersOut && a.compact) fail(RTC_ERROR_INVALID_OPERATION, "traversal counters are not available for RTC_SCENE_FLAG_COMPACT scenes");
  if (countersOut) {
    std::lock_guard<std::mutex> l(dev->stageMutex);
    if (!dev->dCounters) cudaCheck(cudaMalloc((void**)&dev->dCounters, sizeof(RQTraceCounters)), "counters");
    dC = dev->dCounters;
    cudaCheck(cudaMemsetAsync(dC, 0, sizeof(RQTraceCounters), dev->stream()), "counters");
    cudaCheck(cudaStreamSynchronize(dev->stream()), "counters");
  }
  a.counters = dC;
  const bool coherent = ctx && (ctx->flags & RTC_INTERSECT_CONTEXT_FLAG_COHERENT);   // a hint only, as in the reference
  a.refillBelow = (unsigned)(occluded ? dev->refillOccluded : coherent ? dev->refillCoherent : dev->refillClosest);
  a.split = (occluded ? dev->splitOccluded : coherent ? dev->splitCoherent : dev->splitClosest) ? 1u : 0u;
  a.tVote = (unsigned)dev->tVote;
  a.stackSmem = (unsigned)dev->stackSmem;
  void* mapped = nullptr;
  const bool onDevice = isDevicePointer(rays);
  if (!onDevice && dev->zeroCopy && M >= 4096) mapped = mappedHostPointer(rays);
  if (onDevice || (mapped && dev->zeroCopy == 2)) {
    // Device-resident stream: traced in place.  Page-locked host stream: also traced in place -- the
    // persistent kernel reads each ray once and writes only the hit fields of rays that hit, so the
    // PCIe link carries ~64 of 80 bytes per ray inbound and ~10 bytes per ray outbound instead of
    // two full copies of the stream; the call still returns only when the results are in the buffer.
    cudaStream_t s = dev->stream();
    a.rays = mapped ? mapped : rays; a.numRays = M; a.stride = stride;
    a.workCounter = dev->dWork + 8 * Device::kRing;
    {
      std::lock_guard<std::mutex> ll(dev->launchMutex);
      cudaCheck(occluded ? rqLaunchOccluded(&a, (rqStream)s) : rqLaunchIntersect(&a, (rqStream)s), "trace launch");
    }
    if (!dev->async || countersOut || mapped) cudaCheck(cudaStreamSynchronize(s), "trace");
  } else if (dev->d2hMode == 3 && !mapped && !countersOut && !sc->numInstances && !a.compact && M >= 65536 && stride >= recBytes &&
             // page-locked streams below ~4 M rays are faster through whole-span copies; pageable ones never are (the driver bounces them)
             (M >= dev->compactMinRays || (M >= (1u << 18) && dev->packPageable && mappedHostPointer(rays) == nullptr)) &&
             a.depth <= 32 + (unsigned)dev->stackSmem) {
    traceStreamCompact(dev, a, (char*)rays, M, stride, occluded, recBytes);
  } else if (!mapped && !countersOut && M < 65536) {
    // Short host streams (a tile of rays per call, many calling threads -- how the reference's tutorials drive rtcIntersect1M,
    // viewer_stream_device.cpp:287-341): every call borrows its own (stream, buffer, cursor) from a free list, so concurrent
    // callers overlap on the GPU instead of queueing behind one staging mutex.
    Device::SmallStage* st = dev->acquireSmallStage();
    try {
      const size_t span = (size_t)(M - 1) * stride + recBytes;
      if (st->cap < span) {
        if (st->buf) cudaFree(st->buf);
        st->buf = nullptr; st->cap = 0;
        const size_t want = std::max<size_t>(span + 256, 1u << 20);
        cudaCheck(cudaMalloc(&st->buf, want), "staging alloc");
        st->cap = want - 256;
      }
      cudaCheck(cudaMemcpyAsync(st->buf, rays, span, cudaMemcpyHostToDevice, st->stream), "ray upload");
      a.rays = st->buf; a.out = nullptr; a.numRays = M; a.stride = stride; a.workCounter = st->work;
      cudaCheck(occluded ? rqLaunchOccluded(&a, (rqStream)st->stream) : rqLaunchIntersect(&a, (rqStream)st->stream), "trace launch");
      cudaCheck(cudaMemcpyAsync(rays, st->buf, span, cudaMemcpyDeviceToHost, st->stream), "hit download");
      cudaCheck(cudaStreamSynchronize(st->stream), "trace");
      dev->h2dBytes += span; dev->d2hBytes += span;
    } catch (...) { cudaStreamSynchronize(st->stream); cudaGetLastError(); dev->releaseSmallStage(st); throw; }
    dev->releaseSmallStage(st);
  } else {
    std::lock_guard<std::mutex> l(dev->stageMutex);     // host-staged calls of one device are serialised
    // 1 M-ray chunks for long streams, ~8 pieces for shorter ones so that the copies of neighbouring pieces overlap
    // (configs[0], 1 M rays: 3.3 ms as one chunk, 2.35 ms in 128-256 K pieces; profiles/r01y_small_probe.log)
    size_t chunk = dev->chunkRays;
    if ((size_t)M < 8 * chunk) chunk = std::min(chunk, std::max<size_t>(65536, (((size_t)M + 7) / 8 + 32767) & ~(size_t)32767));
    unsigned done = 0; int slot = 0;
    try {
    while (done < M) {
      const unsigned n = (unsigned)std::min<size_t>(chunk, M - done);
      const size_t span = (size_t)(n - 1) * stride + recBytes;
      const int r = slot % Device::kRing; slot++;
      if (!dev->ringStream[r]) cudaCheck(cudaStreamCreateWithFlags(&dev->ringStream[r], cudaStreamNonBlocking), "stream");
      if (dev->ringCap[r] < span) {
        cudaCheck(cudaStreamSynchronize(dev->ringStream[r]), "staging");
        if (dev->ringBuf[r]) cudaFree(dev->ringBuf[r]);
        dev->ringBuf[r] = nullptr; dev->ringCap[r] = 0;
        cudaCheck(cudaMalloc(&dev->ringBuf[r], span + 256), "staging alloc");
        dev->ringCap[r] = span;
      }
      char* h = (char*)rays + (size_t)done * stride;
      cudaStream_t s = dev->ringStream[r];
      cudaCheck(cudaMemcpyAsync(dev->ringBuf[r], h, span, cudaMemcpyHostToDevice, s), "ray upload");
      dev->h2dBytes += span;
      a.rays = dev->ringBuf[r]; a.numRays = n; a.stride = stride;
      // page-locked caller memory: the kernel writes tfar / the hit of rays that hit straight into it
      // (posted PCIe writes, ~10 bytes per ray on average) instead of copying the whole span back
      a.out = mapped ? (char*)mapped + (size_t)done * stride : nullptr;
      a.workCounter = dev->dWork + 8 * r;
      cudaCheck(occluded ? rqLaunchOccluded(&a, (rqStream)s) : rqLaunchIntersect(&a, (rqStream)s), "trace launch");
      if (!mapped) {
        if ((dev->d2hMode == 1 || dev->d2hMode == 2) && n > 1 && stride >= recBytes) {
          const size_t width = (occluded && dev->d2hMode == 2) ? 4 : recBytes - 32;
          cudaCheck(cudaMemcpy2DAsync(h + 32, stride, (char*)dev->ringBuf[r] + 32, stride, width, n, cudaMemcpyDeviceToHost, s), "hit download");
          dev->d2hBytes += (unsigned long long)width * n;
        } else {
          cudaCheck(cudaMemcpyAsync(h, dev->ringBuf[r], span, cudaMemcpyDeviceToHost, s), "hit download");
          dev->d2hBytes += span;
        }
      }
      done += n;
    }
    } catch (...) {
      // earlier pieces may still be copying into the caller's buffer or reading the ring slots: drain them before the error leaves the call
      for (int r = 0; r < Device::kRing; r++) if (dev->ringStream[r]) cudaStreamSynchronize(dev->ringStream[r]);
      cudaGetLastError();
      throw;
    }
    for (int r = 0; r < Device::kRing; r++) if (dev->ringStream[r]) cudaCheck(cudaStreamSynchronize(dev->ringStream[r]), "trace");
  }
  if (countersOut) {
    cudaCheck(cudaDeviceSynchronize(), "counters");
    cudaCheck(cudaMemcpy(countersOut, dC, sizeof(RQTraceCounters), cudaMemcpyDeviceToHost), "counters");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Adapters for the non-AoS entry points (rtcore_ray.h:52-251; reference: RayStreamFilter::filterAOP / filterSOA / filterSOP,
// bvh_intersector_stream_filters.cpp:155-592).  Every flavour becomes ONE gather of all its rays into a dense AoS scratch
// stream, ONE trace of that stream, ONE scatter of the hits:
//   * layouts living in device memory: gather / scatter kernels (rq_trace.cu), scratch from the stream-ordered pool, the
//     trace runs in place on the scratch -- nothing crosses PCIe;
//   * layouts living in host memory: the caller thread gathers into a host scratch stream that goes through the staged
//     host path of traceStream like any rtcIntersect1M stream.
// ------------------------------------------------------------------------------------------------------------------
RQSoAView viewOfNp(const RTCRayNp* r, const RTCHitNp* h, unsigned N) {
  RQSoAView v; memset(&v, 0, sizeof(v));
  v.org_x = r->org_x; v.org_y = r->org_y; v.org_z = r->org_z; v.tnear = r->tnear; v.dir_x = r->dir_x; v.dir_y = r->dir_y; v.dir_z = r->dir_z;
  v.time = r->time; v.tfar = r->tfar; v.mask = r->mask; v.id = r->id; v.flags = r->flags;
  if (h) { v.Ng_x = h->Ng_x; v.Ng_y = h->Ng_y; v.Ng_z = h->Ng_z; v.u = h->u; v.v = h->v; v.primID = h->primID; v.geomID = h->geomID; v.instID0 = h->instID[0]; }
  v.N = N ? N : 1u; v.packetStride = 0;
  return v;
}
// SoA block(s) of runtime width N (RTCRayN / RTCRayHitN, also RTCRay4/8/16): field k of a packet starts at word k*N
RQSoAView viewOfN(void* base, unsigned N, size_t packetStride, bool withHit) {
  RQSoAView v; memset(&v, 0, sizeof(v));
  float* f = (float*)base; unsigned* u = (unsigned*)base;
  v.org_x = f; v.org_y = f + N; v.org_z = f + 2 * N; v.tnear = f + 3 * N; v.dir_x = f + 4 * N; v.dir_y = f + 5 * N; v.dir_z = f + 6 * N;
  v.time = f + 7 * N; v.tfar = f + 8 * N; v.mask = u + 9 * N; v.id = u + 10 * N; v.flags = u + 11 * N;
  if (withHit) {
    float* h = f + 12 * N; unsigned* uh = (unsigned*)h;
    v.Ng_x = h; v.Ng_y = h + N; v.Ng_z = h + 2 * N; v.u = h + 3 * N; v.v = h + 4 * N; v.primID = uh + 5 * N; v.geomID = uh + 6 * N; v.instID0 = uh + 7 * N;
  }
  v.N = N ? N : 1u; v.packetStride = packetStride;
  return v;
}
template <typename T> inline T* soaAtHost(T* field, const RQSoAView& v, unsigned i) {
  const unsigned m = i / v.N, j = i - m * v.N;
  return (T*)((char*)field + (size_t)m * v.packetStride) + j;
}

struct PoolScratch {                                     // device scratch from the stream-ordered pool, released in stream order
  void* p = nullptr; cudaStream_t s = nullptr;
  void alloc(size_t bytes, cudaStream_t st) { s = st; cudaCheck(cudaMallocAsync(&p, bytes ? bytes : 16, st), "scratch alloc"); }
  ~PoolScratch() { if (p) cudaFreeAsync(p, s); }
};

void traceSoA(Scene* sc, RTCIntersectContext* ctx, const RQSoAView& v, const int* valid, unsigned total, bool occluded, int streamRule) {
  if (total == 0) return;
  Device* dev = sc->dev;
  if (!dev->hasGpu) fail(RTC_ERROR_UNKNOWN, "no CUDA device");
  dev->bind();
  const size_t rec = occluded ? sizeof(RTCRay) : sizeof(RTCRayHit);
  if (isDevicePointer(v.org_x)) {
    cudaStream_t s = dev->stream();
    PoolScratch aos, dvalid;
    aos.alloc((size_t)total * sizeof(RTCRayHit), s);
    const int* dv = valid;
    if (valid && !isDevicePointer(valid)) {              // packet entry points: the (short) valid mask usually lives with the caller
      dvalid.alloc((size_t)total * sizeof(int), s);
      cudaCheck(cudaMemcpyAsync(dvalid.p, valid, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, s), "valid mask upload");
      dv = (const int*)dvalid.p;
    }
    cudaCheck(rqGatherSoA(&v, dv, total, aos.p, (rqStream)s), "SoA gather");
    traceStream(sc, ctx, aos.p, total, sizeof(RTCRayHit), occluded, rec, nullptr, streamRule);
    cudaCheck(rqScatterSoA(&v, total, aos.p, occluded ? 1 : 0, (rqStream)s), "SoA scatter");
    if (!dev->async) cudaCheck(cudaStreamSynchronize(s), "trace");
    return;
  }
  std::vector<RTCRayHit> tmp(total);
  for (unsigned i = 0; i < total; i++) {
    RTCRayHit& r = tmp[i];
    const bool ok = valid ? valid[i] != 0 : true;
    r.ray.org_x = *soaAtHost(v.org_x, v, i); r.ray.org_y = *soaAtHost(v.org_y, v, i); r.ray.org_z = *soaAtHost(v.org_z, v, i);
    r.ray.dir_x = *soaAtHost(v.dir_x, v, i); r.ray.dir_y = *soaAtHost(v.dir_y, v, i); r.ray.dir_z = *soaAtHost(v.dir_z, v, i);
    r.ray.tnear = ok ? *soaAtHost(v.tnear, v, i) : INFINITY; r.ray.tfar = ok ? *soaAtHost(v.tfar, v, i) : -INFINITY;   // inactive lane
    r.ray.time = 0.f; r.ray.mask = r.ray.id = r.ray.flags = 0;
    r.hit.geomID = RTC_INVALID_GEOMETRY_ID;
  }
  traceStream(sc, ctx, tmp.data(), total, sizeof(RTCRayHit), occluded, rec, nullptr, streamRule);
  for (unsigned i = 0; i < total; i++) {
    const RTCRayHit& r = tmp[i];
    const bool ok = valid ? valid[i] != 0 : true;
    if (!ok) continue;
    if (occluded) { if (r.ray.tfar == -INFINITY) *soaAtHost(v.tfar, v, i) = -INFINITY; continue; }
    if (r.hit.geomID == RTC_INVALID_GEOMETRY_ID) continue;
    *soaAtHost(v.tfar, v, i) = r.ray.tfar;
    *soaAtHost(v.Ng_x, v, i) = r.hit.Ng_x; *soaAtHost(v.Ng_y, v, i) = r.hit.Ng_y; *soaAtHost(v.Ng_z, v, i) = r.hit.Ng_z;
    *soaAtHost(v.u, v, i) = r.hit.u; *soaAtHost(v.v, v, i) = r.hit.v;
    *soaAtHost(v.primID, v, i) = r.hit.primID; *soaAtHost(v.geomID, v, i) = r.hit.geomID;
    if (v.instID0) *soaAtHost(v.instID0, v, i) = r.hit.instID[0];
  }
}

// array-of-pointers streams (rtcIntersect1Mp / rtcOccluded1Mp)
void traceAoP(Scene* sc, RTCIntersectContext* ctx, void** ptrs, unsigned M, bool occluded) {
  if (M == 0) return;
  Device* dev = sc->dev;
  if (!dev->hasGpu) fail(RTC_ERROR_UNKNOWN, "no CUDA device");
  dev->bind();
  const size_t rec = occluded ? sizeof(RTCRay) : sizeof(RTCRayHit);
  if (isDevicePointer(ptrs)) {                           // a device array of device pointers
    cudaStream_t s = dev->stream();
    PoolScratch aos;
    aos.alloc((size_t)M * sizeof(RTCRayHit), s);
    cudaCheck(rqGatherAoP((const void* const*)ptrs, M, (int)rec, aos.p, (rqStream)s), "AoP gather");
    traceStream(sc, ctx, aos.p, M, sizeof(RTCRayHit), occluded, rec, nullptr, -1);
    cudaCheck(rqScatterAoP((void* const*)ptrs, M, aos.p, occluded ? 1 : 0, (rqStream)s), "AoP scatter");
    if (!dev->async) cudaCheck(cudaStreamSynchronize(s), "trace");
    return;
  }
  std::vector<RTCRayHit> tmp(M);
  for (unsigned i = 0; i < M; i++) { memcpy(&tmp[i], ptrs[i], sizeof(RTCRay)); tmp[i].hit.geomID = RTC_INVALID_GEOMETRY_ID; }
  traceStream(sc, ctx, tmp.data(), M, sizeof(RTCRayHit), occluded, rec, nullptr, -1);
  for (unsigned i = 0; i < M; i++) {
    if (occluded) { if (tmp[i].ray.tfar == -INFINITY) ((RTCRay*)ptrs[i])->tfar = -INFINITY; continue; }
    if (tmp[i].hit.geomID == RTC_INVALID_GEOMETRY_ID) continue;
    RTCRayHit* d = (RTCRayHit*)ptrs[i];
    d->ray.tfar = tmp[i].ray.tfar; d->hit = tmp[i].hit;
  }
}

inline Device* devOf(Scene* s) { return s ? s->dev : nullptr; }
inline Device* devOf(Geometry* g) { return g ? g->dev : nullptr; }
inline Device* devOf(Buffer* b) { return b ? b->dev : nullptr; }

void unsupported(Device* dev, const char* what) {
  processError(dev, RTC_ERROR_INVALID_OPERATION, (std::string(what) + " is not supported by the B200 ray-query device").c_str());
}

}  // namespace

// ================================================================================================
// device
// ================================================================================================
RTC_API RTCDevice rtcNewDevice(const char* config) {
  std::lock_guard<std::mutex> l(g_apiMutex);
  Device* d = nullptr;
  RTC_TRY
    d = new Device();
    bool allowNoGpu = false;
    // The reference reads extra configuration from .embree3 files (device.cpp:69-77); the equivalent for an application that
    // cannot be recompiled is the environment: B200RQ_CONFIG is parsed after (and overrides) the string the application passes.
    std::string cfgAll = config ? config : "";
    if (const char* envCfg = getenv("B200RQ_CONFIG")) { cfgAll += ","; cfgAll += envCfg; }
    config = cfgAll.c_str();
    parseConfig(d, config, &allowNoGpu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      if (!allowNoGpu) { delete d; d = nullptr; fail(RTC_ERROR_UNKNOWN, "no CUDA device available (this engine has no CPU fallback)"); }
    } else {
      if (d->ordinal < 0 || d->ordinal >= n) { delete d; d = nullptr; fail(RTC_ERROR_INVALID_ARGUMENT, "gpu ordinal out of range"); }
      cudaDeviceProp p;
      cudaCheck(cudaGetDeviceProperties(&p, d->ordinal), "device query");
      if (p.major < 10) { delete d; d = nullptr; fail(RTC_ERROR_UNSUPPORTED_CPU, "kernels are built for sm_100a only"); }
      d->hasGpu = true;
      d->bind();
      cudaCheck(cudaStreamCreateWithFlags(&d->ownStream, cudaStreamNonBlocking), "stream");
      cudaCheck(cudaMalloc((void**)&d->dWork, 32 * (Device::kRing + 1)), "work counters");
      if (d->verbose >= 1)
        printf("b200-rayquery %s on GPU %d: %s, %d SMs, %.1f GB\n", RTC_VERSION_STRING, d->ordinal, p.name, p.multiProcessorCount, p.totalGlobalMem / 1e9);
      if (d->numGpus > 1) {
        // gpus=N: GPUs ordinal .. ordinal+N-1; the host threads of the box are shared between them
        if (d->ordinal + d->numGpus > n) { delete d; d = nullptr; fail(RTC_ERROR_INVALID_ARGUMENT, "gpus=N exceeds the number of CUDA devices"); }
        const int hw = (int)std::thread::hardware_concurrency();
        const int perGpu = std::max(2, std::min(8, hw / d->numGpus));
        if (d->scatterThreads == 8) d->scatterThreads = perGpu;
        for (int g = 1; g < d->numGpus; g++) {
          Device* pd = new Device();
          d->peers.push_back(pd);
          bool dummy = false;
          parseConfig(pd, config, &dummy);
          pd->numGpus = 1; pd->ordinal = d->ordinal + g; pd->scatterThreads = d->scatterThreads; pd->hasGpu = true;
          pd->bind();
          cudaCheck(cudaStreamCreateWithFlags(&pd->ownStream, cudaStreamNonBlocking), "stream");
          cudaCheck(cudaMalloc((void**)&pd->dWork, 32 * (Device::kRing + 1)), "work counters");
          int can = 0;
          if (cudaDeviceCanAccessPeer(&can, pd->ordinal, d->ordinal) == cudaSuccess && can) { if (cudaDeviceEnablePeerAccess(d->ordinal, 0) != cudaSuccess) cudaGetLastError(); }
          d->bind();
          can = 0;
          if (cudaDeviceCanAccessPeer(&can, d->ordinal, pd->ordinal) == cudaSuccess && can) { if (cudaDeviceEnablePeerAccess(pd->ordinal, 0) != cudaSuccess) cudaGetLastError(); }
        }
        // The images live in the stream-ordered pools (cudaMallocAsync), and pool memory is NOT covered by cudaDeviceEnablePeerAccess:
        // without an explicit access grant cudaMemcpyPeerAsync between two pool allocations is staged through host memory (the
        // replication of the 0.66 GB image took 18-22 ms = PCIe speed on NV18 boxes, profiles/r02r_cabi_gpus.jsonl).  Every GPU of the
        // group may read and write every other GPU's pool.
        for (int a = 0; a < d->numGpus; a++) {
          cudaMemPool_t pool = nullptr;
          if (cudaDeviceGetDefaultMemPool(&pool, d->ordinal + a) != cudaSuccess) { cudaGetLastError(); continue; }
          for (int b = 0; b < d->numGpus; b++) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d->ordinal + b, d->ordinal + a) != cudaSuccess || !can) { cudaGetLastError(); continue; }
            cudaMemAccessDesc desc; memset(&desc, 0, sizeof(desc));
            desc.location.type = cudaMemLocationTypeDevice; desc.location.id = d->ordinal + b; desc.flags = cudaMemAccessFlagsProtReadWrite;
            if (cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) cudaGetLastError();
          }
        }
        d->bind();
      }
    }
    return (RTCDevice)d;
  RTC_CATCH(nullptr)
  return nullptr;
}

RTC_API void rtcRetainDevice(RTCDevice h) {
  std::lock_guard<std::mutex> l(g_apiMutex);
  RTC_TRY VERIFY_HANDLE(h); ((Device*)h)->retain(); RTC_CATCH((Device*)h)
}
RTC_API void rtcReleaseDevice(RTCDevice h) {
  std::lock_guard<std::mutex> l(g_apiMutex);
  RTC_TRY VERIFY_HANDLE(h); ((Device*)h)->release(); RTC_CATCH(nullptr)
}

RTC_API ssize_t rtcGetDeviceProperty(RTCDevice h, enum RTCDeviceProperty prop) {
  Device* d = (Device*)h;
  RTC_TRY
    VERIFY_HANDLE(h);
    switch (prop) {
      case RTC_DEVICE_PROPERTY_VERSION: return RTC_VERSION;
      case RTC_DEVICE_PROPERTY_VERSION_MAJOR: return RTC_VERSION_MAJOR;
      case RTC_DEVICE_PROPERTY_VERSION_MINOR: return RTC_VERSION_MINOR;
      case RTC_DEVICE_PROPERTY_VERSION_PATCH: return RTC_VERSION_PATCH;
      case RTC_DEVICE_PROPERTY_NATIVE_RAY4_SUPPORTED: case RTC_DEVICE_PROPERTY_NATIVE_RAY8_SUPPORTED:
      case RTC_DEVICE_PROPERTY_NATIVE_RAY16_SUPPORTED: return 0;        // packets are repacked, streams are native
      case RTC_DEVICE_PROPERTY_RAY_STREAM_SUPPORTED: return 1;
      case RTC_DEVICE_PROPERTY_BACKFACE_CULLING_CURVES_ENABLED: return 0;
      case RTC_DEVICE_PROPERTY_RAY_MASK_SUPPORTED: return 0;
      case RTC_DEVICE_PROPERTY_BACKFACE_CULLING_ENABLED: return 0;
      case RTC_DEVICE_PROPERTY_FILTER_FUNCTION_SUPPORTED: return 0;
      case RTC_DEVICE_PROPERTY_IGNORE_INVALID_RAYS_ENABLED: return 0;
      case RTC_DEVICE_PROPERTY_COMPACT_POLYS_ENABLED: return 0;
      case RTC_DEVICE_PROPERTY_TRIANGLE_GEOMETRY_SUPPORTED: case RTC_DEVICE_PROPERTY_QUAD_GEOMETRY_SUPPORTED: return 1;
      case RTC_DEVICE_PROPERTY_SUBDIVISION_GEOMETRY_SUPPORTED:
      case RTC_DEVICE_PROPERTY_CURVE_GEOMETRY_SUPPORTED: case RTC_DEVICE_PROPERTY_USER_GEOMETRY_SUPPORTED:
      case RTC_DEVICE_PROPERTY_POINT_GEOMETRY_SUPPORTED: return 0;
      case RTC_DEVICE_PROPERTY_TASKING_SYSTEM: return 0;
      case RTC_DEVICE_PROPERTY_JOIN_COMMIT_SUPPORTED: return 1;
      case RTC_DEVICE_PROPERTY_PARALLEL_COMMIT_SUPPORTED: return 0;
      default: fail(RTC_ERROR_INVALID_ARGUMENT, "unknown readable property");
    }
  RTC_CATCH(d)
  return 0;
}
RTC_API void rtcSetDeviceProperty(RTCDevice h, const enum RTCDeviceProperty, ssize_t) {
  Device* d = (Device*)h;
  RTC_TRY VERIFY_HANDLE(h); fail(RTC_ERROR_INVALID_ARGUMENT, "unknown writable property"); RTC_CATCH(d)
}
RTC_API enum RTCError rtcGetDeviceError(RTCDevice h) {
  Device* d = (Device*)h;
  if (!d) { RTCError e = g_threadError; g_threadError = RTC_ERROR_NONE; return e; }
  std::lock_guard<std::mutex> l(d->errMutex);
  RTCError e = d->error; d->error = RTC_ERROR_NONE; return e;
}
RTC_API void rtcSetDeviceErrorFunction(RTCDevice h, RTCErrorFunction fn, void* p) {
  Device* d = (Device*)h;
  RTC_TRY VERIFY_HANDLE(h); { std::lock_guard<std::mutex> l(d->errMutex); d->errFn = fn; d->errPtr = p; } RTC_CATCH(d)
}
RTC_API void rtcSetDeviceMemoryMonitorFunction(RTCDevice h, RTCMemoryMonitorFunction fn, void* p) {
  Device* d = (Device*)h;
  RTC_TRY VERIFY_HANDLE(h); d->memFn = fn; d->memPtr = p; RTC_CATCH(d)
}

// ================================================================================================
// buffers
// ================================================================================================
RTC_API RTCBuffer rtcNewBuffer(RTCDevice h, size_t bytes) {
  Device* d = (Device*)h;
  RTC_TRY VERIFY_HANDLE(h); return (RTCBuffer) new Buffer(d, bytes, nullptr); RTC_CATCH(d)
  return nullptr;
}
RTC_API RTCBuffer rtcNewSharedBuffer(RTCDevice h, void* ptr, size_t bytes) {
  Device* d = (Device*)h;
  RTC_TRY VERIFY_HANDLE(h); return (RTCBuffer) new Buffer(d, bytes, ptr); RTC_CATCH(d)   // a NULL pointer is not rejected by the reference either (rtcore.cpp:128-141); it then owns its memory here
  return nullptr;
}
RTC_API void* rtcGetBufferData(RTCBuffer h) {
  Buffer* b = (Buffer*)h;
  RTC_TRY VERIFY_HANDLE(h); return b->ptr; RTC_CATCH(devOf(b))
  return nullptr;
}
RTC_API void rtcRetainBuffer(RTCBuffer h) { Buffer* b = (Buffer*)h; RTC_TRY VERIFY_HANDLE(h); b->retain(); RTC_CATCH(devOf(b)) }
RTC_API void rtcReleaseBuffer(RTCBuffer h) { Buffer* b = (Buffer*)h; Device* d = devOf(b); RTC_TRY VERIFY_HANDLE(h); b->release(); RTC_CATCH(d) }

// ================================================================================================
// geometry
// ================================================================================================
RTC_API RTCGeometry rtcNewGeometry(RTCDevice h, enum RTCGeometryType type) {
  Device* d = (Device*)h;
  RTC_TRY
    VERIFY_HANDLE(h);
    if (type != RTC_GEOMETRY_TYPE_TRIANGLE && type != RTC_GEOMETRY_TYPE_QUAD && type != RTC_GEOMETRY_TYPE_INSTANCE)
      fail(RTC_ERROR_INVALID_OPERATION, "only RTC_GEOMETRY_TYPE_TRIANGLE, _QUAD and _INSTANCE are supported by the B200 ray-query device");
    return (RTCGeometry) new Geometry(d, type);
  RTC_CATCH(d)
  return nullptr;
}
RTC_API void rtcRetainGeometry(RTCGeometry h) { Geometry* g = (Geometry*)h; RTC_TRY VERIFY_HANDLE(h); g->retain(); RTC_CATCH(devOf(g)) }
RTC_API void rtcReleaseGeometry(RTCGeometry h) { Geometry* g = (Geometry*)h; Device* d = devOf(g); RTC_TRY VERIFY_HANDLE(h); g->release(); RTC_CATCH(d) }
RTC_API void rtcCommitGeometry(RTCGeometry h) { Geometry* g = (Geometry*)h; RTC_TRY VERIFY_HANDLE(h); ++g->modCounter; g->committed = true; RTC_CATCH(devOf(g)) }
RTC_API void rtcEnableGeometry(RTCGeometry h) { Geometry* g = (Geometry*)h; RTC_TRY VERIFY_HANDLE(h); if (!g->enabled) { g->enabled = true; ++g->modCounter; ++g->topoCounter; } RTC_CATCH(devOf(g)) }
RTC_API void rtcDisableGeometry(RTCGeometry h) { Geometry* g = (Geometry*)h; RTC_TRY VERIFY_HANDLE(h); if (g->enabled) { g->enabled = false; ++g->modCounter; ++g->topoCounter; } RTC_CATCH(devOf(g)) }
RTC_API void rtcSetGeometryTimeStepCount(RTCGeometry h, unsigned int n) {
  Geometry* g = (Geometry*)h;
  RTC_TRY VERIFY_HANDLE(h);
    if (n > RTC_MAX_TIME_STEP_COUNT) fail(RTC_ERROR_INVALID_ARGUMENT, "number of time steps is out of range");   // rtcore.cpp:1335-1336
    if (n > 1) fail(RTC_ERROR_INVALID_OPERATION, "motion blur is not supported by the B200 ray-query device");
  RTC_CATCH(devOf(g))
}
RTC_API void rtcSetGeometryMask(RTCGeometry h, unsigned int mask) { Geometry* g = (Geometry*)h; RTC_TRY VERIFY_HANDLE(h); g->mask = mask; RTC_CATCH(devOf(g)) }
RTC_API void rtcSetGeometryBuildQuality(RTCGeometry h, enum RTCBuildQuality q) {
  Geometry* g = (Geometry*)h;
  RTC_TRY VERIFY_HANDLE(h);
    if (q != RTC_BUILD_QUALITY_LOW && q != RTC_BUILD_QUALITY_MEDIUM && q != RTC_BUILD_QUALITY_HIGH && q != RTC_BUILD_QUALITY_REFIT)
      throw std::runtime_error("invalid build quality");   // the reference throws a plain runtime_error here: RTC_ERROR_UNKNOWN (rtcore.cpp:233,1386)
    g->quality = q; g->update();
  RTC_CATCH(devOf(g))
}

namespace {
void setBuffer(Geometry* g, RTCBufferType type, unsigned slot, RTCFormat format, Buffer* buf, size_t offset, size_t stride, unsigned num) {
  if (g->type != RTC_GEOMETRY_TYPE_TRIANGLE && g->type != RTC_GEOMETRY_TYPE_QUAD) fail(RTC_ERROR_INVALID_OPERATION, "operation not supported for this geometry");
  if ((((size_t)buf->ptr + offset) & 3) || (stride & 3)) fail(RTC_ERROR_INVALID_OPERATION, "data must be 4 bytes aligned");
  if (type == RTC_BUFFER_TYPE_VERTEX) {
    if (format != RTC_FORMAT_FLOAT3) fail(RTC_ERROR_INVALID_OPERATION, "invalid vertex buffer format");
    if (stride * (size_t)num > 16ull * 1024 * 1024 * 1024) fail(RTC_ERROR_INVALID_OPERATION, "vertex buffer can be at most 16GB large");
    if (slot != 0) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid vertex buffer slot");
    g->vertices.set(buf, offset, stride, num, format);
  } else if (type == RTC_BUFFER_TYPE_INDEX) {
    if (slot != 0) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid buffer slot");
    if (format != (g->type == RTC_GEOMETRY_TYPE_QUAD ? RTC_FORMAT_UINT4 : RTC_FORMAT_UINT3))
      fail(RTC_ERROR_INVALID_OPERATION, "invalid index buffer format");   // scene_triangle_mesh.cpp:62, scene_quad_mesh.cpp:62
    g->indices.set(buf, offset, stride, num, format);
  } else if (type == RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE) {
    fail(RTC_ERROR_INVALID_OPERATION, "vertex attributes are not supported by the B200 ray-query device");
  } else {
    fail(RTC_ERROR_INVALID_ARGUMENT, "unknown buffer type");
  }
  g->update();
}
}  // namespace

RTC_API void rtcSetGeometryBuffer(RTCGeometry h, enum RTCBufferType type, unsigned int slot, enum RTCFormat format, RTCBuffer hb,
                                  size_t byteOffset, size_t byteStride, size_t itemCount) {
  Geometry* g = (Geometry*)h; Buffer* b = (Buffer*)hb;
  RTC_TRY
    VERIFY_HANDLE(h); VERIFY_HANDLE(hb);
    if (g->dev != b->dev) fail(RTC_ERROR_INVALID_ARGUMENT, "inputs are from different devices");
    if (itemCount > 0xFFFFFFFFu) fail(RTC_ERROR_INVALID_ARGUMENT, "buffer too large");
    setBuffer(g, type, slot, format, b, byteOffset, byteStride, (unsigned)itemCount);
  RTC_CATCH(devOf(g))
}
RTC_API void rtcSetSharedGeometryBuffer(RTCGeometry h, enum RTCBufferType type, unsigned int slot, enum RTCFormat format, const void* ptr,
                                        size_t byteOffset, size_t byteStride, size_t itemCount) {
  Geometry* g = (Geometry*)h;
  RTC_TRY
    VERIFY_HANDLE(h);
    if (itemCount > 0xFFFFFFFFu) fail(RTC_ERROR_INVALID_ARGUMENT, "buffer too large");
    Buffer* b = new Buffer(g->dev, itemCount * byteStride, (char*)ptr + byteOffset);
    try { setBuffer(g, type, slot, format, b, 0, byteStride, (unsigned)itemCount); } catch (...) { b->release(); throw; }
    b->release();
  RTC_CATCH(devOf(g))
}
RTC_API void* rtcSetNewGeometryBuffer(RTCGeometry h, enum RTCBufferType type, unsigned int slot, enum RTCFormat format,
                                      size_t byteStride, size_t itemCount) {
  Geometry* g = (Geometry*)h;
  RTC_TRY
    VERIFY_HANDLE(h);
    if (itemCount > 0xFFFFFFFFu) fail(RTC_ERROR_INVALID_ARGUMENT, "buffer too large");
    size_t bytes = itemCount * byteStride;
    if (type == RTC_BUFFER_TYPE_VERTEX || type == RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE) bytes += (16 - (byteStride % 16)) % 16;
    Buffer* b = new Buffer(g->dev, bytes, nullptr);
    try { setBuffer(g, type, slot, format, b, 0, byteStride, (unsigned)itemCount); } catch (...) { b->release(); throw; }
    void* p = b->ptr;
    b->release();
    return p;
  RTC_CATCH(devOf(g))
  return nullptr;
}
RTC_API void* rtcGetGeometryBufferData(RTCGeometry h, enum RTCBufferType type, unsigned int slot) {
  Geometry* g = (Geometry*)h;
  RTC_TRY
    VERIFY_HANDLE(h);
    if (type == RTC_BUFFER_TYPE_INDEX && slot == 0) return (void*)g->indices.data();
    if (type == RTC_BUFFER_TYPE_VERTEX && slot == 0) return (void*)g->vertices.data();
    fail(RTC_ERROR_INVALID_ARGUMENT, "unknown buffer type");
  RTC_CATCH(devOf(g))
  return nullptr;
}
RTC_API void rtcUpdateGeometryBuffer(RTCGeometry h, enum RTCBufferType type, unsigned int slot) {
  Geometry* g = (Geometry*)h;
  RTC_TRY
    VERIFY_HANDLE(h);
    // scene_triangle_mesh.cpp:109-135: index slot 0, vertex slot < number of time steps (1 here), anything else is an invalid argument
    if (type == RTC_BUFFER_TYPE_INDEX) { if (slot != 0) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid buffer slot"); g->update(); }
    else if (type == RTC_BUFFER_TYPE_VERTEX) { if (slot != 0) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid buffer slot"); g->updateVertices(); }
    else if (type == RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid buffer slot");   // no attribute buffers can be bound
    else fail(RTC_ERROR_INVALID_ARGUMENT, "unknown buffer type");
  RTC_CATCH(devOf(g))
}
RTC_API void rtcSetGeometryUserData(RTCGeometry h, void* p) { Geometry* g = (Geometry*)h; RTC_TRY VERIFY_HANDLE(h); g->userPtr = p; RTC_CATCH(devOf(g)) }
RTC_API void* rtcGetGeometryUserData(RTCGeometry h) { Geometry* g = (Geometry*)h; RTC_TRY VERIFY_HANDLE(h); return g->userPtr; RTC_CATCH(devOf(g)) return nullptr; }
RTC_API void rtcSetGeometryIntersectFilterFunction(RTCGeometry h, RTCFilterFunctionN f) {
  Geometry* g = (Geometry*)h;
  RTC_TRY VERIFY_HANDLE(h); if (f) fail(RTC_ERROR_INVALID_OPERATION, "filter callbacks cannot run on the GPU"); RTC_CATCH(devOf(g))
}
RTC_API void rtcSetGeometryOccludedFilterFunction(RTCGeometry h, RTCFilterFunctionN f) {
  Geometry* g = (Geometry*)h;
  RTC_TRY VERIFY_HANDLE(h); if (f) fail(RTC_ERROR_INVALID_OPERATION, "filter callbacks cannot run on the GPU"); RTC_CATCH(devOf(g))
}

// ---- instances (kernels/common/rtcore.cpp:1008-1125, scene_instance.cpp) ----
RTC_API void rtcSetGeometryInstancedScene(RTCGeometry h, RTCScene hs) {
  Geometry* g = (Geometry*)h; Scene* sc = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(h); VERIFY_HANDLE(hs);
    if (g->type != RTC_GEOMETRY_TYPE_INSTANCE) fail(RTC_ERROR_INVALID_OPERATION, "operation not supported for this geometry");
    if (g->dev != sc->dev) fail(RTC_ERROR_INVALID_ARGUMENT, "inputs are from different devices");
    sc->retain();
    if (g->instanced) g->instanced->release();
    g->instanced = sc;
    g->update();
  RTC_CATCH(devOf(g))
}
RTC_API void rtcSetGeometryTransform(RTCGeometry h, unsigned int timeStep, enum RTCFormat format, const void* xfm) {
  Geometry* g = (Geometry*)h;
  RTC_TRY
    VERIFY_HANDLE(h); VERIFY_HANDLE(xfm);
    if (g->type != RTC_GEOMETRY_TYPE_INSTANCE) fail(RTC_ERROR_INVALID_OPERATION, "operation not supported for this geometry");
    if (timeStep != 0) fail(RTC_ERROR_INVALID_OPERATION, "motion blur is not supported by the B200 ray-query device");
    const float* m = (const float*)xfm;
    float* o = g->l2w;                                        // vx, vy, vz, p  (loadTransform, rtcore.cpp:1008-1039)
    switch (format) {
      case RTC_FORMAT_FLOAT3X4_ROW_MAJOR:
        o[0] = m[0]; o[1] = m[4]; o[2] = m[8];  o[3] = m[1]; o[4] = m[5]; o[5] = m[9];
        o[6] = m[2]; o[7] = m[6]; o[8] = m[10]; o[9] = m[3]; o[10] = m[7]; o[11] = m[11];
        break;
      case RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR:
        for (int i = 0; i < 12; i++) o[i] = m[i];
        break;
      case RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR:
        for (int c = 0; c < 4; c++) for (int r = 0; r < 3; r++) o[3 * c + r] = m[4 * c + r];
        break;
      default: fail(RTC_ERROR_INVALID_OPERATION, "invalid matrix format");
    }
    g->update();
  RTC_CATCH(devOf(g))
}
RTC_API void rtcGetGeometryTransform(RTCGeometry h, float, enum RTCFormat format, void* xfm) {
  Geometry* g = (Geometry*)h;
  RTC_TRY
    VERIFY_HANDLE(h); VERIFY_HANDLE(xfm);
    float* m = (float*)xfm;
    const float* o = g->l2w;                                  // identity for non-instance geometries, as in the reference (geometry.h getTransform)
    switch (format) {                                         // storeTransform, rtcore.cpp:1041-1069
      case RTC_FORMAT_FLOAT3X4_ROW_MAJOR:
        m[0] = o[0]; m[1] = o[3]; m[2] = o[6]; m[3] = o[9];
        m[4] = o[1]; m[5] = o[4]; m[6] = o[7]; m[7] = o[10];
        m[8] = o[2]; m[9] = o[5]; m[10] = o[8]; m[11] = o[11];
        break;
      case RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR:
        for (int i = 0; i < 12; i++) m[i] = o[i];
        break;
      case RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR:
        for (int c = 0; c < 4; c++) { for (int r = 0; r < 3; r++) m[4 * c + r] = o[3 * c + r]; m[4 * c + 3] = (c == 3) ? 1.f : 0.f; }
        break;
      default: fail(RTC_ERROR_INVALID_OPERATION, "invalid matrix format");
    }
  RTC_CATCH(devOf(g))
}

// ================================================================================================
// scene
// ================================================================================================
RTC_API RTCScene rtcNewScene(RTCDevice h) {
  Device* d = (Device*)h;
  RTC_TRY VERIFY_HANDLE(h); return (RTCScene) new Scene(d); RTC_CATCH(d)
  return nullptr;
}
RTC_API RTCDevice rtcGetSceneDevice(RTCScene h) {
  Scene* s = (Scene*)h;
  RTC_TRY VERIFY_HANDLE(h); s->dev->retain(); return (RTCDevice)s->dev; RTC_CATCH(devOf(s))   // extra reference (rtcore.cpp:204)
  return nullptr;
}
RTC_API void rtcRetainScene(RTCScene h) { Scene* s = (Scene*)h; RTC_TRY VERIFY_HANDLE(h); s->retain(); RTC_CATCH(devOf(s)) }
RTC_API void rtcReleaseScene(RTCScene h) { Scene* s = (Scene*)h; Device* d = devOf(s); RTC_TRY VERIFY_HANDLE(h); s->release(); RTC_CATCH(d) }

namespace {
unsigned bindGeometry(Scene* s, unsigned geomID, Geometry* g) {
  std::lock_guard<std::mutex> l(s->geomMutex);
  if (geomID == RTC_INVALID_GEOMETRY_ID) {
    if (!s->freeIDs.empty()) { geomID = *s->freeIDs.begin(); s->freeIDs.erase(s->freeIDs.begin()); }
    else geomID = s->nextID++;
  } else {
    if (geomID < s->geoms.size() && s->geoms[geomID]) fail(RTC_ERROR_INVALID_OPERATION, "invalid geometry ID provided");
    if (geomID >= s->nextID) { for (unsigned i = s->nextID; i < geomID; i++) s->freeIDs.insert(i); s->nextID = geomID + 1; }
    else s->freeIDs.erase(geomID);
  }
  if (geomID >= s->geoms.size()) s->geoms.resize(geomID + 1, nullptr);
  g->retain();
  s->geoms[geomID] = g;
  if (geomID < s->seenMod.size()) s->seenMod[geomID] = 0xFFFFFFFFu;
  if (g->enabled) s->modified = true;
  return geomID;
}
}  // namespace

RTC_API unsigned int rtcAttachGeometry(RTCScene hs, RTCGeometry hg) {
  Scene* s = (Scene*)hs; Geometry* g = (Geometry*)hg;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(hg);
    if (s->dev != g->dev) fail(RTC_ERROR_INVALID_ARGUMENT, "inputs are from different devices");
    return bindGeometry(s, RTC_INVALID_GEOMETRY_ID, g);
  RTC_CATCH(devOf(s))
  return (unsigned)-1;
}
RTC_API void rtcAttachGeometryByID(RTCScene hs, RTCGeometry hg, unsigned int geomID) {
  Scene* s = (Scene*)hs; Geometry* g = (Geometry*)hg;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(hg);
    if (geomID == RTC_INVALID_GEOMETRY_ID) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid argument");
    if (s->dev != g->dev) fail(RTC_ERROR_INVALID_ARGUMENT, "inputs are from different devices");
    bindGeometry(s, geomID, g);
  RTC_CATCH(devOf(s))
}
RTC_API void rtcDetachGeometry(RTCScene hs, unsigned int geomID) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs);
    if (geomID == RTC_INVALID_GEOMETRY_ID) fail(RTC_ERROR_INVALID_ARGUMENT, "invalid argument");
    Geometry* g = nullptr;
    {
      std::lock_guard<std::mutex> l(s->geomMutex);
      if (geomID >= s->geoms.size()) fail(RTC_ERROR_INVALID_OPERATION, "invalid geometry ID");
      g = s->geoms[geomID];
      if (!g) fail(RTC_ERROR_INVALID_OPERATION, "invalid geometry");
      if (g->enabled) s->modified = true;
      s->geoms[geomID] = nullptr;
      s->freeIDs.insert(geomID);
    }
    g->release();
  RTC_CATCH(devOf(s))
}
RTC_API RTCGeometry rtcGetGeometry(RTCScene hs, unsigned int geomID) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs);
    std::lock_guard<std::mutex> l(s->geomMutex);
    if (geomID >= s->geoms.size()) return nullptr;
    return (RTCGeometry)s->geoms[geomID];
  RTC_CATCH(devOf(s))
  return nullptr;
}
RTC_API void rtcCommitScene(RTCScene hs) { Scene* s = (Scene*)hs; RTC_TRY VERIFY_HANDLE(hs); commitScene(s); RTC_CATCH(devOf(s)) }
RTC_API void rtcJoinCommitScene(RTCScene hs) { Scene* s = (Scene*)hs; RTC_TRY VERIFY_HANDLE(hs); commitScene(s); RTC_CATCH(devOf(s)) }
RTC_API void rtcSetSceneProgressMonitorFunction(RTCScene hs, RTCProgressMonitorFunction f, void* p) {
  Scene* s = (Scene*)hs;
  RTC_TRY VERIFY_HANDLE(hs); s->progress = f; s->progressPtr = p; RTC_CATCH(devOf(s))
}
RTC_API void rtcSetSceneBuildQuality(RTCScene hs, enum RTCBuildQuality q) {
  Scene* s = (Scene*)hs;
  RTC_TRY VERIFY_HANDLE(hs);
    if (q != RTC_BUILD_QUALITY_LOW && q != RTC_BUILD_QUALITY_MEDIUM && q != RTC_BUILD_QUALITY_HIGH)
      throw std::runtime_error("invalid build quality");   // the reference throws a plain runtime_error here: RTC_ERROR_UNKNOWN (rtcore.cpp:233,1386)
    if (q != s->quality) { s->quality = q; s->modified = true; }
  RTC_CATCH(devOf(s))
}
RTC_API void rtcSetSceneFlags(RTCScene hs, enum RTCSceneFlags f) {
  Scene* s = (Scene*)hs;
  RTC_TRY VERIFY_HANDLE(hs); if (f != s->flags) { s->flags = f; s->modified = true; } RTC_CATCH(devOf(s))
}
RTC_API enum RTCSceneFlags rtcGetSceneFlags(RTCScene hs) {
  Scene* s = (Scene*)hs;
  RTC_TRY VERIFY_HANDLE(hs); return s->flags; RTC_CATCH(devOf(s))
  return RTC_SCENE_FLAG_NONE;
}
RTC_API void rtcGetSceneBounds(RTCScene hs, struct RTCBounds* b) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(b);
    if (!s->everCommitted || s->modified) fail(RTC_ERROR_INVALID_OPERATION, "scene not committed");
    const RQImageHeader& H = s->image.header;
    b->lower_x = H.lo[0]; b->lower_y = H.lo[1]; b->lower_z = H.lo[2]; b->align0 = 0;
    b->upper_x = H.hi[0]; b->upper_y = H.hi[1]; b->upper_z = H.hi[2]; b->align1 = 0;
  RTC_CATCH(devOf(s))
}
RTC_API void rtcGetSceneLinearBounds(RTCScene hs, struct RTCLinearBounds* b) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(b);
    rtcGetSceneBounds(hs, &b->bounds0); b->bounds1 = b->bounds0;
  RTC_CATCH(devOf(s))
}

// ================================================================================================
// queries
// ================================================================================================
RTC_API void rtcIntersect1M(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRayHit* rayhit, unsigned int M, size_t byteStride) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); traceStream(s, ctx, rayhit, M, byteStride, false, sizeof(RTCRayHit), nullptr); RTC_CATCH(devOf(s))
}
RTC_API void rtcOccluded1M(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRay* ray, unsigned int M, size_t byteStride) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); traceStream(s, ctx, ray, M, byteStride, true, sizeof(RTCRay), nullptr); RTC_CATCH(devOf(s))
}
RTC_API void rtcIntersect1(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRayHit* rayhit) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); traceStream(s, ctx, rayhit, 1, sizeof(RTCRayHit), false, sizeof(RTCRayHit), nullptr); RTC_CATCH(devOf(s))
}
RTC_API void rtcOccluded1(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRay* ray) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); traceStream(s, ctx, ray, 1, sizeof(RTCRay), true, sizeof(RTCRay), nullptr); RTC_CATCH(devOf(s))
}
// Packets of 4 / 8 / 16 rays: the lanes follow the packet kernels' entry rules (tnear clamped to 0, not rejected:
// bvh_intersector_hybrid.cpp:153,403), so the stream rule is off.
#define B200RQ_PACKET_ENTRY(W)                                                                                              \
  RTC_API void rtcIntersect##W(const int* valid, RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRayHit##W* rh) {    \
    Scene* s = (Scene*)hs;                                                                                                  \
    RTC_TRY checkQuery(s, ctx); traceSoA(s, ctx, viewOfN(rh, W, 0, true), valid, W, false, 0); RTC_CATCH(devOf(s))           \
  }                                                                                                                         \
  RTC_API void rtcOccluded##W(const int* valid, RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRay##W* r) {         \
    Scene* s = (Scene*)hs;                                                                                                  \
    RTC_TRY checkQuery(s, ctx); traceSoA(s, ctx, viewOfN(r, W, 0, false), valid, W, true, 0); RTC_CATCH(devOf(s))            \
  }
B200RQ_PACKET_ENTRY(4)
B200RQ_PACKET_ENTRY(8)
B200RQ_PACKET_ENTRY(16)

// rtcore.cpp:626-731,877-905: M == 1 is the single-ray path, otherwise filterAOP (stream rules)
RTC_API void rtcIntersect1Mp(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRayHit** rh, unsigned int M) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); traceAoP(s, ctx, (void**)rh, M, false); RTC_CATCH(devOf(s))
}
RTC_API void rtcOccluded1Mp(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRay** r, unsigned int M) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); traceAoP(s, ctx, (void**)r, M, true); RTC_CATCH(devOf(s))
}
// rtcore.cpp:733-790,906-945: N == 1 is an AoS stream of M records, otherwise filterSOA: M packets of N lanes.  Occlusion rays take
// the filter's octant-sorting branch (stream_filters.cpp:333-404 -> occludedN), i.e. the stream entry rules
RTC_API void rtcIntersectNM(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRayHitN* rh, unsigned int N, unsigned int M, size_t byteStride) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    checkQuery(s, ctx);
    if (N == 1) traceStream(s, ctx, rh, M, byteStride, false, sizeof(RTCRayHit), nullptr);
    else if ((unsigned long long)N * M > 0xFFFFFFFFull) fail(RTC_ERROR_INVALID_ARGUMENT, "too many rays in one call");
    else traceSoA(s, ctx, viewOfN(rh, N, byteStride, true), nullptr, N * M, false, 0);
  RTC_CATCH(devOf(s))
}
RTC_API void rtcOccludedNM(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRayN* r, unsigned int N, unsigned int M, size_t byteStride) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    checkQuery(s, ctx);
    if (N == 1) traceStream(s, ctx, r, M, byteStride, true, sizeof(RTCRay), nullptr);
    else if ((unsigned long long)N * M > 0xFFFFFFFFull) fail(RTC_ERROR_INVALID_ARGUMENT, "too many rays in one call");
    else traceSoA(s, ctx, viewOfN(r, N, byteStride, false), nullptr, N * M, true, 1);
  RTC_CATCH(devOf(s))
}
// rtcore.cpp:792-846,947-974: filterSOP (occlusion: octant-sorting branch, stream_filters.cpp:497-577 -> stream entry rules)
RTC_API void rtcIntersectNp(RTCScene hs, struct RTCIntersectContext* ctx, const struct RTCRayHitNp* rh, unsigned int N) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); VERIFY_HANDLE(rh); traceSoA(s, ctx, viewOfNp(&rh->ray, &rh->hit, N), nullptr, N, false, 0); RTC_CATCH(devOf(s))
}
RTC_API void rtcOccludedNp(RTCScene hs, struct RTCIntersectContext* ctx, const struct RTCRayNp* r, unsigned int N) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); VERIFY_HANDLE(r); traceSoA(s, ctx, viewOfNp(r, nullptr, N), nullptr, N, true, 1); RTC_CATCH(devOf(s))
}

// ================================================================================================
// B200 extensions (include/rq_b200.h)
// ================================================================================================
RTC_API void rtcxSetDeviceStream(RTCDevice h, void* stream) { Device* d = (Device*)h; RTC_TRY VERIFY_HANDLE(h); d->userStream = (cudaStream_t)stream; RTC_CATCH(d) }
RTC_API void rtcxSynchronizeDevice(RTCDevice h) {
  Device* d = (Device*)h;
  RTC_TRY VERIFY_HANDLE(h); if (d->hasGpu) { d->bind(); cudaCheck(cudaStreamSynchronize(d->stream()), "synchronize"); } RTC_CATCH(d)
}
RTC_API int rtcxGetDeviceOrdinal(RTCDevice h) { Device* d = (Device*)h; return d && d->hasGpu ? d->ordinal : -1; }
RTC_API int rtcxGetDeviceGpuCount(RTCDevice h) { Device* d = (Device*)h; return d && d->hasGpu ? 1 + (int)d->peers.size() : 0; }
RTC_API int rtcxGetSceneBuildStats(RTCScene hs, struct RTCXBuildStats* o) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(o);
    if (!s->everCommitted) fail(RTC_ERROR_INVALID_OPERATION, "scene not committed");
    static_assert(sizeof(RTCXBuildStats) == sizeof(RQBuildStats), "stats layouts must agree");
    memcpy(o, &s->stats, sizeof(*o));
    return 0;
  RTC_CATCH(devOf(s))
  return -1;
}
RTC_API const void* rtcxGetSceneImage(RTCScene hs, size_t* bytes) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs);
    if (!s->everCommitted || s->modified) fail(RTC_ERROR_INVALID_OPERATION, "scene not committed");
    if (s->numInstances) fail(RTC_ERROR_INVALID_OPERATION, "the image of a scene with instances refers to other scenes and cannot be exported");
    if (bytes) *bytes = (size_t)s->image.header.totalBytes;
    return s->image.base;
  RTC_CATCH(devOf(s))
  return nullptr;
}
RTC_API void rtcxCopySceneImage(RTCScene hs, void* dst, size_t bytes) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(dst);
    if (!s->everCommitted || s->modified) fail(RTC_ERROR_INVALID_OPERATION, "scene not committed");
    if (s->numInstances) fail(RTC_ERROR_INVALID_OPERATION, "the image of a scene with instances refers to other scenes and cannot be exported");
    if (bytes != s->image.header.totalBytes) fail(RTC_ERROR_INVALID_ARGUMENT, "size does not match the image");
    s->dev->bind();
    cudaCheck(cudaMemcpyAsync(dst, s->image.base, bytes, cudaMemcpyDefault, s->dev->stream()), "image copy");
    cudaCheck(cudaStreamSynchronize(s->dev->stream()), "image copy");
  RTC_CATCH(devOf(s))
}
namespace {
// adopt a byte copy of a BVH image from device- or host-readable memory; throws on a bad image
void adoptImage(Scene* s, const void* src, size_t bytes) {
  Device* dev = s->dev;
  if (!dev->hasGpu) fail(RTC_ERROR_UNKNOWN, "no CUDA device");
  if (bytes < sizeof(RQImageHeader)) fail(RTC_ERROR_INVALID_ARGUMENT, "image too small");
  dev->bind();
  std::unique_lock<std::mutex> lock(s->buildMutex);
  RQImageHeader H;
  cudaCheck(cudaMemcpy(&H, src, sizeof(H), cudaMemcpyDefault), "image header");
  bool okHeader = H.magic == RQ_IMAGE_MAGIC && H.totalBytes == bytes && H.nodesOffset == 128 && H.layout <= 1u &&
                  H.trisOffset == H.nodesOffset + (uint64_t)H.numNodes * sizeof(RQNode);
  if (okHeader && H.layout == 1u)
    okHeader = H.metaOffset >= H.trisOffset + (uint64_t)H.numTris * sizeof(RQTriC) && H.vertsOffset >= H.metaOffset + (uint64_t)H.numTris * 4ull &&
               H.vertsOffset + (uint64_t)H.numVerts * 16ull <= H.totalBytes && (H.metaOffset & 15u) == 0 && (H.vertsOffset & 15u) == 0;
  else if (okHeader)
    okHeader = H.trisOffset + (uint64_t)H.numTris * sizeof(RQTri) <= H.totalBytes;
  if (!okHeader) fail(RTC_ERROR_INVALID_ARGUMENT, "not a BVH image");
  if (H.depth == 0 || H.depth > RQ_MAX_LEVELS || H.numNodes == 0) fail(RTC_ERROR_INVALID_ARGUMENT, "not a BVH image (depth / node count out of range)");
  void* p = nullptr;
  cudaCheck(rqAllocImage(&p, bytes, (rqStream)dev->stream()), "image alloc");
  int e = cudaMemcpyAsync(p, src, bytes, cudaMemcpyDefault, dev->stream());
  // a truncated or corrupt file with a plausible header must not reach the traversal kernels: out-of-range child / triangle
  // references would be context-fatal illegal addresses, a too small depth silently wrong answers
  unsigned int violations = 0;
  if (!e) e = rqValidateImage(p, &H, (rqStream)dev->stream(), &violations);
  if (!e) e = cudaStreamSynchronize(dev->stream());
  if (e || violations) {
    RQDeviceImage tmpImg; memset(&tmpImg, 0, sizeof(tmpImg)); tmpImg.base = p; rqFreeImage(&tmpImg);
    cudaCheck(e, "image copy");
    fail(RTC_ERROR_INVALID_ARGUMENT, "corrupt BVH image (references out of range)");
  }
  if (s->image.base) { if (s->accounted) freeImage(dev, &s->image); else rqFreeImage(&s->image); }
  s->accounted = false;
  if (s->dInstances) { cudaFree(s->dInstances); s->dInstances = nullptr; }
  s->numInstances = 0; s->clearInstanced(); s->epoch++;
  s->image.base = p; s->image.header = H; s->image.numLevels = 0;   // adopted image: level ranges unknown, never refitted
  s->flags = (RTCSceneFlags)H.flags;
  memset(&s->stats, 0, sizeof(s->stats));
  s->stats.numNodes = H.numNodes; s->stats.numTris = H.numTris; s->stats.depth = H.depth; s->stats.sah = H.sah; s->stats.bytes = H.totalBytes;
  s->stats.numPrimsValid = H.numTris;
  s->modified = false; s->everCommitted = true;
  { std::lock_guard<std::mutex> gl(s->geomMutex); s->seenMod.assign(s->geoms.size(), 0); for (size_t i = 0; i < s->geoms.size(); i++) if (s->geoms[i]) s->seenMod[i] = s->geoms[i]->modCounter; }
}
}  // namespace
RTC_API void rtcxSetSceneImage(RTCScene hs, const void* src, size_t bytes) {
  Scene* s = (Scene*)hs;
  RTC_TRY VERIFY_HANDLE(hs); VERIFY_HANDLE(src); adoptImage(s, src, bytes); RTC_CATCH(devOf(s))
}
// The flat image is position independent, so a file holding its bytes is a loadable BVH
// (the reference has no serialisation).
RTC_API int rtcxSaveSceneImage(RTCScene hs, const char* path) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(path);
    if (!s->everCommitted || s->modified) fail(RTC_ERROR_INVALID_OPERATION, "scene not committed");
    if (s->numInstances) fail(RTC_ERROR_INVALID_OPERATION, "the image of a scene with instances refers to other scenes and cannot be exported");
    const size_t bytes = (size_t)s->image.header.totalBytes;
    std::vector<char> host(bytes);
    s->dev->bind();
    cudaCheck(cudaMemcpyAsync(host.data(), s->image.base, bytes, cudaMemcpyDeviceToHost, s->dev->stream()), "image download");
    cudaCheck(cudaStreamSynchronize(s->dev->stream()), "image download");
    FILE* f = fopen(path, "wb");
    if (!f) fail(RTC_ERROR_INVALID_ARGUMENT, "cannot open file for writing");
    const size_t w = fwrite(host.data(), 1, bytes, f);
    const int c = fclose(f);
    if (w != bytes || c != 0) fail(RTC_ERROR_UNKNOWN, "short write");
    return 0;
  RTC_CATCH(devOf(s))
  return -1;
}
RTC_API int rtcxLoadSceneImage(RTCScene hs, const char* path) {
  Scene* s = (Scene*)hs;
  RTC_TRY
    VERIFY_HANDLE(hs); VERIFY_HANDLE(path);
    FILE* f = fopen(path, "rb");
    if (!f) fail(RTC_ERROR_INVALID_ARGUMENT, "cannot open file for reading");
    std::vector<char> host;
    try {
      fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
      if (n < (long)sizeof(RQImageHeader)) { fclose(f); f = nullptr; fail(RTC_ERROR_INVALID_ARGUMENT, "not a BVH image"); }
      host.resize((size_t)n);
      const size_t r = fread(host.data(), 1, (size_t)n, f);
      fclose(f); f = nullptr;
      if (r != (size_t)n) fail(RTC_ERROR_UNKNOWN, "short read");
    } catch (...) { if (f) fclose(f); throw; }
    adoptImage(s, host.data(), host.size());                   // validates the header, uploads, marks committed
    return 0;
  RTC_CATCH(devOf(s))
  return -1;
}
RTC_API void rtcxIntersect1MCounted(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRayHit* rh, unsigned int M, size_t stride,
                                    struct RTCXTraceCounters* out) {
  Scene* s = (Scene*)hs;
  static_assert(sizeof(RTCXTraceCounters) == sizeof(RQTraceCounters), "counter layouts must agree");
  RTC_TRY checkQuery(s, ctx); VERIFY_HANDLE(out); traceStream(s, ctx, rh, M, stride, false, sizeof(RTCRayHit), (RQTraceCounters*)out); RTC_CATCH(devOf(s))
}
RTC_API void rtcxOccluded1MCounted(RTCScene hs, struct RTCIntersectContext* ctx, struct RTCRay* r, unsigned int M, size_t stride,
                                   struct RTCXTraceCounters* out) {
  Scene* s = (Scene*)hs;
  RTC_TRY checkQuery(s, ctx); VERIFY_HANDLE(out); traceStream(s, ctx, r, M, stride, true, sizeof(RTCRay), (RQTraceCounters*)out); RTC_CATCH(devOf(s))
}
RTC_API unsigned long long rtcxGetLaunchCount(void) { return rqLaunchCount(); }
RTC_API void rtcxGetTransferBytes(RTCDevice h, unsigned long long* h2d, unsigned long long* d2h) {
  Device* d = (Device*)h;
  RTC_TRY
    VERIFY_HANDLE(h);
    unsigned long long a = d->h2dBytes.load(), b = d->d2hBytes.load();
    for (Device* p : d->peers) { a += p->h2dBytes.load(); b += p->d2hBytes.load(); }
    if (h2d) *h2d = a;
    if (d2h) *d2h = b;
  RTC_CATCH(d)
}

// ================================================================================================
// Entry points outside the hot path: exported so existing programs link, raise INVALID_OPERATION.
// (reference: kernels/common/rtcore.cpp -- curves, subdivision, instancing, user geometry,
//  point queries, collision, interpolation, BVH builder API)
// ================================================================================================
#define B200RQ_STUB(ret, name, args, devexpr, retval) \
  RTC_API ret name args { unsupported((devexpr), #name); return retval; }
#define B200RQ_NODEV ((Device*)nullptr)
B200RQ_STUB(void, rtcSetGeometryTimeRange, (RTCGeometry g, float, float), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryVertexAttributeCount, (RTCGeometry g, unsigned int), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryMaxRadiusScale, (RTCGeometry g, float), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryPointQueryFunction, (RTCGeometry g, RTCPointQueryFunction), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryUserPrimitiveCount, (RTCGeometry g, unsigned int), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryBoundsFunction, (RTCGeometry g, void*, void*), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryIntersectFunction, (RTCGeometry g, void*), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryOccludedFunction, (RTCGeometry g, void*), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcFilterIntersection, (const void*, const void*), B200RQ_NODEV, )
B200RQ_STUB(void, rtcFilterOcclusion, (const void*, const void*), B200RQ_NODEV, )
B200RQ_STUB(void, rtcSetGeometryTransformQuaternion, (RTCGeometry g, unsigned int, const void*), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryTessellationRate, (RTCGeometry g, float), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryTopologyCount, (RTCGeometry g, unsigned int), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometrySubdivisionMode, (RTCGeometry g, unsigned int, enum RTCSubdivisionMode), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryVertexAttributeTopology, (RTCGeometry g, unsigned int, unsigned int), devOf((Geometry*)g), )
B200RQ_STUB(void, rtcSetGeometryDisplacementFunction, (RTCGeometry g, void*), devOf((Geometry*)g), )
B200RQ_STUB(unsigned int, rtcGetGeometryFirstHalfEdge, (RTCGeometry g, unsigned int), devOf((Geometry*)g), 0)
B200RQ_STUB(unsigned int, rtcGetGeometryFace, (RTCGeometry g, unsigned int), devOf((Geometry*)g), 0)
B200RQ_STUB(unsigned int, rtcGetGeometryNextHalfEdge, (RTCGeometry g, unsigned int), devOf((Geometry*)g), 0)
B200RQ_STUB(unsigned int, rtcGetGeometryPreviousHalfEdge, (RTCGeometry g, unsigned int), devOf((Geometry*)g), 0)
B200RQ_STUB(unsigned int, rtcGetGeometryOppositeHalfEdge, (RTCGeometry g, unsigned int, unsigned int), devOf((Geometry*)g), 0)
B200RQ_STUB(void, rtcInterpolate, (const void*), B200RQ_NODEV, )
B200RQ_STUB(void, rtcInterpolateN, (const void*), B200RQ_NODEV, )
B200RQ_STUB(bool, rtcPointQuery, (RTCScene s, void*, void*, RTCPointQueryFunction, void*), devOf((Scene*)s), false)
B200RQ_STUB(bool, rtcPointQuery4, (const int*, RTCScene s, void*, void*, RTCPointQueryFunction, void**), devOf((Scene*)s), false)
B200RQ_STUB(bool, rtcPointQuery8, (const int*, RTCScene s, void*, void*, RTCPointQueryFunction, void**), devOf((Scene*)s), false)
B200RQ_STUB(bool, rtcPointQuery16, (const int*, RTCScene s, void*, void*, RTCPointQueryFunction, void**), devOf((Scene*)s), false)
B200RQ_STUB(void, rtcCollide, (RTCScene s, RTCScene, void*, void*), devOf((Scene*)s), )
B200RQ_STUB(void*, rtcNewBVH, (RTCDevice d), (Device*)d, nullptr)
B200RQ_STUB(void*, rtcBuildBVH, (const void*), B200RQ_NODEV, nullptr)
B200RQ_STUB(void*, rtcThreadLocalAlloc, (void*, size_t, size_t), B200RQ_NODEV, nullptr)
B200RQ_STUB(void, rtcRetainBVH, (void*), B200RQ_NODEV, )
B200RQ_STUB(void, rtcReleaseBVH, (void*), B200RQ_NODEV, )
B200RQ_STUB(void, rtcMakeStaticBVH, (void*), B200RQ_NODEV, )   /* kernels/common/rtcore_builder.cpp:409 (exported, not in the public header) */
