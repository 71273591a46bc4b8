/* b200-rayquery: public C API of the B200-native ray-query engine.
 *
 * This single header is OUR OWN statement of the slice of Embree 3.12.1's `rtcore` C API that the
 * hot path needs (triangle-mesh scenes, BVH commit, closest-hit / any-hit ray streams).  It is
 * binary compatible with the reference headers -- same struct layouts, same enumerator values,
 * same function signatures -- so that an application compiled against the reference's
 * <embree3/rtcore.h> links against libembree3.so from this repository unchanged.
 * Layout / value sources (reference, read-only):
 *     RTCRay/RTCHit/RTCRayHit  ..... include/embree3/rtcore_ray.h:11-49   (48 B + 32 B, 16-B aligned)
 *     RTCRay{4,8,16}, Np, N  ....... include/embree3/rtcore_ray.h:52-251
 *     RTCFormat, RTCBuildQuality, RTCBounds, RTCIntersectContext .. rtcore_common.h:44-224
 *     RTCDeviceProperty, RTCError .. include/embree3/rtcore_device.h:14-85
 *     RTCBufferType  ............... include/embree3/rtcore_buffer.h:11-49
 *     RTCSceneFlags + scene calls .. include/embree3/rtcore_scene.h:17-141
 *     RTCGeometryType + geometry ... include/embree3/rtcore_geometry.h:18-199
 * Everything is in one file (the per-topic headers of the reference are thin forwarders here).
 *
 * Entry points outside the hot path (curves, subdivision, instancing, point queries, collision,
 * filter callbacks, interpolation) are still exported so existing programs link; they raise
 * RTC_ERROR_INVALID_OPERATION on the device.
 */
#ifndef B200_RAYQUERY_RTCORE_H
#define B200_RAYQUERY_RTCORE_H

#include <stddef.h>
#include <stdbool.h>
#include <sys/types.h>

#define RTC_VERSION_MAJOR 3
#define RTC_VERSION_MINOR 12
#define RTC_VERSION_PATCH 1
#define RTC_VERSION 31201
#define RTC_VERSION_STRING "3.12.1-b200"
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1
#define RTC_MIN_WIDTH 0
#define RTC_MAX_TIME_STEP_COUNT 129
#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)

#define RTC_NAMESPACE_BEGIN
#define RTC_NAMESPACE_END
#define RTC_NAMESPACE_USE

#ifdef __cplusplus
#  define RTC_API_EXTERN_C extern "C"
#else
#  define RTC_API_EXTERN_C
#endif
#if defined(RTC_EXPORT_API)
#  define RTC_API RTC_API_EXTERN_C __attribute__((visibility("default")))
#else
#  define RTC_API RTC_API_EXTERN_C
#endif
#define RTC_ALIGN(n) __attribute__((aligned(n)))
#define RTC_FORCEINLINE inline __attribute__((always_inline))
#ifndef RTC_DEPRECATED
#  define RTC_DEPRECATED __attribute__((deprecated))
#endif

/* ------------------------------------------------------------------ handles */
typedef struct RTCDeviceTy*   RTCDevice;
typedef struct RTCBufferTy*   RTCBuffer;
typedef struct RTCSceneTy*    RTCScene;
typedef struct RTCGeometryTy* RTCGeometry;

/* -------------------------------------------------------------------- enums */
enum RTCFormat {
  RTC_FORMAT_UNDEFINED = 0,
  RTC_FORMAT_UCHAR  = 0x1001, RTC_FORMAT_UCHAR2,  RTC_FORMAT_UCHAR3,  RTC_FORMAT_UCHAR4,
  RTC_FORMAT_CHAR   = 0x2001, RTC_FORMAT_CHAR2,   RTC_FORMAT_CHAR3,   RTC_FORMAT_CHAR4,
  RTC_FORMAT_USHORT = 0x3001, RTC_FORMAT_USHORT2, RTC_FORMAT_USHORT3, RTC_FORMAT_USHORT4,
  RTC_FORMAT_SHORT  = 0x4001, RTC_FORMAT_SHORT2,  RTC_FORMAT_SHORT3,  RTC_FORMAT_SHORT4,
  RTC_FORMAT_UINT   = 0x5001, RTC_FORMAT_UINT2,   RTC_FORMAT_UINT3,   RTC_FORMAT_UINT4,
  RTC_FORMAT_INT    = 0x6001, RTC_FORMAT_INT2,    RTC_FORMAT_INT3,    RTC_FORMAT_INT4,
  RTC_FORMAT_ULLONG = 0x7001, RTC_FORMAT_ULLONG2, RTC_FORMAT_ULLONG3, RTC_FORMAT_ULLONG4,
  RTC_FORMAT_LLONG  = 0x8001, RTC_FORMAT_LLONG2,  RTC_FORMAT_LLONG3,  RTC_FORMAT_LLONG4,
  RTC_FORMAT_FLOAT  = 0x9001, RTC_FORMAT_FLOAT2,  RTC_FORMAT_FLOAT3,  RTC_FORMAT_FLOAT4,
  RTC_FORMAT_FLOAT5,  RTC_FORMAT_FLOAT6,  RTC_FORMAT_FLOAT7,  RTC_FORMAT_FLOAT8,
  RTC_FORMAT_FLOAT9,  RTC_FORMAT_FLOAT10, RTC_FORMAT_FLOAT11, RTC_FORMAT_FLOAT12,
  RTC_FORMAT_FLOAT13, RTC_FORMAT_FLOAT14, RTC_FORMAT_FLOAT15, RTC_FORMAT_FLOAT16,
  RTC_FORMAT_FLOAT2X2_ROW_MAJOR = 0x9122, RTC_FORMAT_FLOAT2X3_ROW_MAJOR = 0x9123,
  RTC_FORMAT_FLOAT2X4_ROW_MAJOR = 0x9124, RTC_FORMAT_FLOAT3X2_ROW_MAJOR = 0x9132,
  RTC_FORMAT_FLOAT3X3_ROW_MAJOR = 0x9133, RTC_FORMAT_FLOAT3X4_ROW_MAJOR = 0x9134,
  RTC_FORMAT_FLOAT4X2_ROW_MAJOR = 0x9142, RTC_FORMAT_FLOAT4X3_ROW_MAJOR = 0x9143,
  RTC_FORMAT_FLOAT4X4_ROW_MAJOR = 0x9144,
  RTC_FORMAT_FLOAT2X2_COLUMN_MAJOR = 0x9222, RTC_FORMAT_FLOAT2X3_COLUMN_MAJOR = 0x9223,
  RTC_FORMAT_FLOAT2X4_COLUMN_MAJOR = 0x9224, RTC_FORMAT_FLOAT3X2_COLUMN_MAJOR = 0x9232,
  RTC_FORMAT_FLOAT3X3_COLUMN_MAJOR = 0x9233, RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR = 0x9234,
  RTC_FORMAT_FLOAT4X2_COLUMN_MAJOR = 0x9242, RTC_FORMAT_FLOAT4X3_COLUMN_MAJOR = 0x9243,
  RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR = 0x9244,
  RTC_FORMAT_GRID = 0xA001
};

enum RTCBuildQuality {
  RTC_BUILD_QUALITY_LOW = 0, RTC_BUILD_QUALITY_MEDIUM = 1,
  RTC_BUILD_QUALITY_HIGH = 2, RTC_BUILD_QUALITY_REFIT = 3
};

enum RTCIntersectContextFlags {
  RTC_INTERSECT_CONTEXT_FLAG_NONE = 0,
  RTC_INTERSECT_CONTEXT_FLAG_INCOHERENT = 0,   /* default: rays of a stream are unrelated   */
  RTC_INTERSECT_CONTEXT_FLAG_COHERENT = 1      /* hint only: neighbouring rays are similar */
};

enum RTCDeviceProperty {
  RTC_DEVICE_PROPERTY_VERSION = 0, RTC_DEVICE_PROPERTY_VERSION_MAJOR = 1,
  RTC_DEVICE_PROPERTY_VERSION_MINOR = 2, RTC_DEVICE_PROPERTY_VERSION_PATCH = 3,
  RTC_DEVICE_PROPERTY_NATIVE_RAY4_SUPPORTED = 32, RTC_DEVICE_PROPERTY_NATIVE_RAY8_SUPPORTED = 33,
  RTC_DEVICE_PROPERTY_NATIVE_RAY16_SUPPORTED = 34, RTC_DEVICE_PROPERTY_RAY_STREAM_SUPPORTED = 35,
  RTC_DEVICE_PROPERTY_BACKFACE_CULLING_CURVES_ENABLED = 63,
  RTC_DEVICE_PROPERTY_RAY_MASK_SUPPORTED = 64, RTC_DEVICE_PROPERTY_BACKFACE_CULLING_ENABLED = 65,
  RTC_DEVICE_PROPERTY_FILTER_FUNCTION_SUPPORTED = 66,
  RTC_DEVICE_PROPERTY_IGNORE_INVALID_RAYS_ENABLED = 67,
  RTC_DEVICE_PROPERTY_COMPACT_POLYS_ENABLED = 68,
  RTC_DEVICE_PROPERTY_TRIANGLE_GEOMETRY_SUPPORTED = 96, RTC_DEVICE_PROPERTY_QUAD_GEOMETRY_SUPPORTED = 97,
  RTC_DEVICE_PROPERTY_SUBDIVISION_GEOMETRY_SUPPORTED = 98,
  RTC_DEVICE_PROPERTY_CURVE_GEOMETRY_SUPPORTED = 99, RTC_DEVICE_PROPERTY_USER_GEOMETRY_SUPPORTED = 100,
  RTC_DEVICE_PROPERTY_POINT_GEOMETRY_SUPPORTED = 101,
  RTC_DEVICE_PROPERTY_TASKING_SYSTEM = 128, RTC_DEVICE_PROPERTY_JOIN_COMMIT_SUPPORTED = 129,
  RTC_DEVICE_PROPERTY_PARALLEL_COMMIT_SUPPORTED = 130
};

enum RTCError {
  RTC_ERROR_NONE = 0, RTC_ERROR_UNKNOWN = 1, RTC_ERROR_INVALID_ARGUMENT = 2,
  RTC_ERROR_INVALID_OPERATION = 3, RTC_ERROR_OUT_OF_MEMORY = 4,
  RTC_ERROR_UNSUPPORTED_CPU = 5, RTC_ERROR_CANCELLED = 6
};

enum RTCBufferType {
  RTC_BUFFER_TYPE_INDEX = 0, RTC_BUFFER_TYPE_VERTEX = 1, RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE = 2,
  RTC_BUFFER_TYPE_NORMAL = 3, RTC_BUFFER_TYPE_TANGENT = 4, RTC_BUFFER_TYPE_NORMAL_DERIVATIVE = 5,
  RTC_BUFFER_TYPE_GRID = 8, RTC_BUFFER_TYPE_FACE = 16, RTC_BUFFER_TYPE_LEVEL = 17,
  RTC_BUFFER_TYPE_EDGE_CREASE_INDEX = 18, RTC_BUFFER_TYPE_EDGE_CREASE_WEIGHT = 19,
  RTC_BUFFER_TYPE_VERTEX_CREASE_INDEX = 20, RTC_BUFFER_TYPE_VERTEX_CREASE_WEIGHT = 21,
  RTC_BUFFER_TYPE_HOLE = 22, RTC_BUFFER_TYPE_FLAGS = 32
};

enum RTCSceneFlags {
  RTC_SCENE_FLAG_NONE = 0, RTC_SCENE_FLAG_DYNAMIC = 1, RTC_SCENE_FLAG_COMPACT = 2,
  RTC_SCENE_FLAG_ROBUST = 4,                   /* selects the watertight (Pluecker) triangle test */
  RTC_SCENE_FLAG_CONTEXT_FILTER_FUNCTION = 8
};

enum RTCGeometryType {
  RTC_GEOMETRY_TYPE_TRIANGLE = 0,              /* the only type this engine builds */
  RTC_GEOMETRY_TYPE_QUAD = 1, RTC_GEOMETRY_TYPE_GRID = 2, RTC_GEOMETRY_TYPE_SUBDIVISION = 8,
  RTC_GEOMETRY_TYPE_CONE_LINEAR_CURVE = 15, RTC_GEOMETRY_TYPE_ROUND_LINEAR_CURVE = 16,
  RTC_GEOMETRY_TYPE_FLAT_LINEAR_CURVE = 17, RTC_GEOMETRY_TYPE_ROUND_BEZIER_CURVE = 24,
  RTC_GEOMETRY_TYPE_FLAT_BEZIER_CURVE = 25, RTC_GEOMETRY_TYPE_NORMAL_ORIENTED_BEZIER_CURVE = 26,
  RTC_GEOMETRY_TYPE_ROUND_BSPLINE_CURVE = 32, RTC_GEOMETRY_TYPE_FLAT_BSPLINE_CURVE = 33,
  RTC_GEOMETRY_TYPE_NORMAL_ORIENTED_BSPLINE_CURVE = 34, RTC_GEOMETRY_TYPE_ROUND_HERMITE_CURVE = 40,
  RTC_GEOMETRY_TYPE_FLAT_HERMITE_CURVE = 41, RTC_GEOMETRY_TYPE_NORMAL_ORIENTED_HERMITE_CURVE = 42,
  RTC_GEOMETRY_TYPE_SPHERE_POINT = 50, RTC_GEOMETRY_TYPE_DISC_POINT = 51,
  RTC_GEOMETRY_TYPE_ORIENTED_DISC_POINT = 52, RTC_GEOMETRY_TYPE_ROUND_CATMULL_ROM_CURVE = 58,
  RTC_GEOMETRY_TYPE_FLAT_CATMULL_ROM_CURVE = 59, RTC_GEOMETRY_TYPE_NORMAL_ORIENTED_CATMULL_ROM_CURVE = 60,
  RTC_GEOMETRY_TYPE_USER = 120, RTC_GEOMETRY_TYPE_INSTANCE = 121
};

enum RTCSubdivisionMode {
  RTC_SUBDIVISION_MODE_NO_BOUNDARY = 0, RTC_SUBDIVISION_MODE_SMOOTH_BOUNDARY = 1,
  RTC_SUBDIVISION_MODE_PIN_CORNERS = 2, RTC_SUBDIVISION_MODE_PIN_BOUNDARY = 3,
  RTC_SUBDIVISION_MODE_PIN_ALL = 4
};

/* ----------------------------------------------------- single ray / hit (AoS) */
struct RTC_ALIGN(16) RTCRay {
  float org_x, org_y, org_z, tnear;            /* origin, start of the parametric interval      */
  float dir_x, dir_y, dir_z, time;             /* direction (need not be normalised), time unused */
  float tfar;                                  /* in: end of interval; out: hit distance / -inf */
  unsigned int mask, id, flags;                /* carried through untouched                      */
};
struct RTC_ALIGN(16) RTCHit {
  float Ng_x, Ng_y, Ng_z;                      /* unnormalised geometric normal cross(v1-v0,v2-v0) */
  float u, v;                                  /* barycentrics: P = v0 + u(v1-v0) + v(v2-v0)     */
  unsigned int primID, geomID;
  unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT];
};
struct RTCRayHit { struct RTCRay ray; struct RTCHit hit; };

/* ------------------------------------------------------- fixed-width packets */
#define B200RQ_DECL_PACKET(W, A)                                                             \
  struct RTC_ALIGN(A) RTCRay##W {                                                            \
    float org_x[W], org_y[W], org_z[W], tnear[W], dir_x[W], dir_y[W], dir_z[W], time[W],     \
          tfar[W];                                                                           \
    unsigned int mask[W], id[W], flags[W]; };                                                \
  struct RTC_ALIGN(A) RTCHit##W {                                                            \
    float Ng_x[W], Ng_y[W], Ng_z[W], u[W], v[W];                                             \
    unsigned int primID[W], geomID[W], instID[RTC_MAX_INSTANCE_LEVEL_COUNT][W]; };           \
  struct RTCRayHit##W { struct RTCRay##W ray; struct RTCHit##W hit; };
B200RQ_DECL_PACKET(4, 16)
B200RQ_DECL_PACKET(8, 32)
B200RQ_DECL_PACKET(16, 64)

/* ------------------------------------- pointer-SoA stream and runtime-N SoA  */
struct RTCRayNp {
  float *org_x, *org_y, *org_z, *tnear, *dir_x, *dir_y, *dir_z, *time, *tfar;
  unsigned int *mask, *id, *flags;
};
struct RTCHitNp {
  float *Ng_x, *Ng_y, *Ng_z, *u, *v;
  unsigned int *primID, *geomID, *instID[RTC_MAX_INSTANCE_LEVEL_COUNT];
};
struct RTCRayHitNp { struct RTCRayNp ray; struct RTCHitNp hit; };
struct RTCRayN; struct RTCHitN; struct RTCRayHitN;   /* N floats per field, field-major */

#ifdef __cplusplus
/* field k of a runtime-N SoA block lives at word k*N + i (ray: 12 fields, hit: 7 + instID levels) */
#define B200RQ_RAYN_F(name, k) RTC_FORCEINLINE float& RTCRayN_##name(RTCRayN* p, unsigned int N, unsigned int i) { return ((float*)p)[(k)*N+i]; }
#define B200RQ_RAYN_U(name, k) RTC_FORCEINLINE unsigned int& RTCRayN_##name(RTCRayN* p, unsigned int N, unsigned int i) { return ((unsigned int*)p)[(k)*N+i]; }
#define B200RQ_HITN_F(name, k) RTC_FORCEINLINE float& RTCHitN_##name(RTCHitN* p, unsigned int N, unsigned int i) { return ((float*)p)[(k)*N+i]; }
#define B200RQ_HITN_U(name, k) RTC_FORCEINLINE unsigned int& RTCHitN_##name(RTCHitN* p, unsigned int N, unsigned int i) { return ((unsigned int*)p)[(k)*N+i]; }
B200RQ_RAYN_F(org_x,0) B200RQ_RAYN_F(org_y,1) B200RQ_RAYN_F(org_z,2) B200RQ_RAYN_F(tnear,3)
B200RQ_RAYN_F(dir_x,4) B200RQ_RAYN_F(dir_y,5) B200RQ_RAYN_F(dir_z,6) B200RQ_RAYN_F(time,7)
B200RQ_RAYN_F(tfar,8)  B200RQ_RAYN_U(mask,9)  B200RQ_RAYN_U(id,10)   B200RQ_RAYN_U(flags,11)
B200RQ_HITN_F(Ng_x,0)  B200RQ_HITN_F(Ng_y,1)  B200RQ_HITN_F(Ng_z,2)  B200RQ_HITN_F(u,3) B200RQ_HITN_F(v,4)
B200RQ_HITN_U(primID,5) B200RQ_HITN_U(geomID,6)
RTC_FORCEINLINE unsigned int& RTCHitN_instID(RTCHitN* p, unsigned int N, unsigned int i, unsigned int l) { return ((unsigned int*)p)[7*N+i+N*l]; }
RTC_FORCEINLINE RTCRayN* RTCRayHitN_RayN(RTCRayHitN* p, unsigned int N) { return (RTCRayN*)&((float*)p)[0]; }
RTC_FORCEINLINE RTCHitN* RTCRayHitN_HitN(RTCRayHitN* p, unsigned int N) { return (RTCHitN*)&((float*)p)[12*N]; }
#endif

/* ----------------------------------------------------- bounds, query context */
struct RTC_ALIGN(16) RTCBounds {
  float lower_x, lower_y, lower_z, align0;
  float upper_x, upper_y, upper_z, align1;
};
struct RTC_ALIGN(16) RTCLinearBounds { struct RTCBounds bounds0, bounds1; };

struct RTCFilterFunctionNArguments {
  int* valid; void* geometryUserPtr; struct RTCIntersectContext* context;
  struct RTCRayN* ray; struct RTCHitN* hit; unsigned int N;
};
typedef void (*RTCFilterFunctionN)(const struct RTCFilterFunctionNArguments* args);

struct RTCIntersectContext {
  enum RTCIntersectContextFlags flags;
  RTCFilterFunctionN filter;                   /* must be NULL: callbacks cannot run on the GPU   */
  unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT];   /* copied into hit.instID on every hit    */
};
RTC_FORCEINLINE void rtcInitIntersectContext(struct RTCIntersectContext* c) {
  c->flags = RTC_INTERSECT_CONTEXT_FLAG_INCOHERENT; c->filter = NULL;
  c->instID[0] = RTC_INVALID_GEOMETRY_ID;
}

struct RTC_ALIGN(16) RTCPointQuery { float x, y, z, time, radius; };
struct RTC_ALIGN(16) RTCPointQueryContext {
  float world2inst[RTC_MAX_INSTANCE_LEVEL_COUNT][16];
  float inst2world[RTC_MAX_INSTANCE_LEVEL_COUNT][16];
  unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT];
  unsigned int instStackSize;
};
struct RTC_ALIGN(16) RTCPointQueryFunctionArguments {
  struct RTCPointQuery* query; void* userPtr; unsigned int primID, geomID;
  struct RTCPointQueryContext* context; float similarityScale;
};
typedef bool (*RTCPointQueryFunction)(struct RTCPointQueryFunctionArguments* args);

typedef void (*RTCErrorFunction)(void* userPtr, enum RTCError code, const char* str);
typedef bool (*RTCMemoryMonitorFunction)(void* ptr, ssize_t bytes, bool post);
typedef bool (*RTCProgressMonitorFunction)(void* ptr, double n);

/* ---------------------------------------------------------------- device API */
RTC_API RTCDevice rtcNewDevice(const char* config);
RTC_API void rtcRetainDevice(RTCDevice device);
RTC_API void rtcReleaseDevice(RTCDevice device);
RTC_API ssize_t rtcGetDeviceProperty(RTCDevice device, enum RTCDeviceProperty prop);
RTC_API void rtcSetDeviceProperty(RTCDevice device, const enum RTCDeviceProperty prop, ssize_t value);
RTC_API enum RTCError rtcGetDeviceError(RTCDevice device);
RTC_API void rtcSetDeviceErrorFunction(RTCDevice device, RTCErrorFunction error, void* userPtr);
RTC_API void rtcSetDeviceMemoryMonitorFunction(RTCDevice device, RTCMemoryMonitorFunction memoryMonitor, void* userPtr);

/* ---------------------------------------------------------------- buffer API */
RTC_API RTCBuffer rtcNewBuffer(RTCDevice device, size_t byteSize);
RTC_API RTCBuffer rtcNewSharedBuffer(RTCDevice device, void* ptr, size_t byteSize);
RTC_API void* rtcGetBufferData(RTCBuffer buffer);
RTC_API void rtcRetainBuffer(RTCBuffer buffer);
RTC_API void rtcReleaseBuffer(RTCBuffer buffer);

/* -------------------------------------------------------------- geometry API */
RTC_API RTCGeometry rtcNewGeometry(RTCDevice device, enum RTCGeometryType type);
RTC_API void rtcRetainGeometry(RTCGeometry geometry);
RTC_API void rtcReleaseGeometry(RTCGeometry geometry);
RTC_API void rtcCommitGeometry(RTCGeometry geometry);
RTC_API void rtcEnableGeometry(RTCGeometry geometry);
RTC_API void rtcDisableGeometry(RTCGeometry geometry);
RTC_API void rtcSetGeometryTimeStepCount(RTCGeometry geometry, unsigned int timeStepCount);
RTC_API void rtcSetGeometryMask(RTCGeometry geometry, unsigned int mask);
RTC_API void rtcSetGeometryBuildQuality(RTCGeometry geometry, enum RTCBuildQuality quality);
RTC_API void rtcSetGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot, enum RTCFormat format, RTCBuffer buffer, size_t byteOffset, size_t byteStride, size_t itemCount);
RTC_API void rtcSetSharedGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot, enum RTCFormat format, const void* ptr, size_t byteOffset, size_t byteStride, size_t itemCount);
RTC_API void* rtcSetNewGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot, enum RTCFormat format, size_t byteStride, size_t itemCount);
RTC_API void* rtcGetGeometryBufferData(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot);
RTC_API void rtcUpdateGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot);
/* instances (rtcore_geometry.h:224-233): single level, time step 0 only */
RTC_API void rtcSetGeometryInstancedScene(RTCGeometry geometry, RTCScene scene);
RTC_API void rtcSetGeometryTransform(RTCGeometry geometry, unsigned int timeStep, enum RTCFormat format, const void* xfm);
RTC_API void rtcGetGeometryTransform(RTCGeometry geometry, float time, enum RTCFormat format, void* xfm);
RTC_API void rtcSetGeometryUserData(RTCGeometry geometry, void* ptr);
RTC_API void* rtcGetGeometryUserData(RTCGeometry geometry);
RTC_API void rtcSetGeometryIntersectFilterFunction(RTCGeometry geometry, RTCFilterFunctionN filter);
RTC_API void rtcSetGeometryOccludedFilterFunction(RTCGeometry geometry, RTCFilterFunctionN filter);

/* ----------------------------------------------------------------- scene API */
RTC_API RTCScene rtcNewScene(RTCDevice device);
RTC_API RTCDevice rtcGetSceneDevice(RTCScene scene);
RTC_API void rtcRetainScene(RTCScene scene);
RTC_API void rtcReleaseScene(RTCScene scene);
RTC_API unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry);
RTC_API void rtcAttachGeometryByID(RTCScene scene, RTCGeometry geometry, unsigned int geomID);
RTC_API void rtcDetachGeometry(RTCScene scene, unsigned int geomID);
RTC_API RTCGeometry rtcGetGeometry(RTCScene scene, unsigned int geomID);
RTC_API void rtcCommitScene(RTCScene scene);
RTC_API void rtcJoinCommitScene(RTCScene scene);
RTC_API void rtcSetSceneProgressMonitorFunction(RTCScene scene, RTCProgressMonitorFunction progress, void* ptr);
RTC_API void rtcSetSceneBuildQuality(RTCScene scene, enum RTCBuildQuality quality);
RTC_API void rtcSetSceneFlags(RTCScene scene, enum RTCSceneFlags flags);
RTC_API enum RTCSceneFlags rtcGetSceneFlags(RTCScene scene);
RTC_API void rtcGetSceneBounds(RTCScene scene, struct RTCBounds* bounds_o);
RTC_API void rtcGetSceneLinearBounds(RTCScene scene, struct RTCLinearBounds* bounds_o);

/* ---- ray queries.  The stream forms (1M) are the hot path; all others funnel into them ---- */
RTC_API void rtcIntersect1(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit* rayhit);
RTC_API void rtcIntersect4(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit4* rayhit);
RTC_API void rtcIntersect8(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit8* rayhit);
RTC_API void rtcIntersect16(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit16* rayhit);
RTC_API void rtcIntersect1M(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit* rayhit, unsigned int M, size_t byteStride);
RTC_API void rtcIntersect1Mp(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit** rayhit, unsigned int M);
RTC_API void rtcIntersectNM(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHitN* rayhit, unsigned int N, unsigned int M, size_t byteStride);
RTC_API void rtcIntersectNp(RTCScene scene, struct RTCIntersectContext* context, const struct RTCRayHitNp* rayhit, unsigned int N);
RTC_API void rtcOccluded1(RTCScene scene, struct RTCIntersectContext* context, struct RTCRay* ray);
RTC_API void rtcOccluded4(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRay4* ray);
RTC_API void rtcOccluded8(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRay8* ray);
RTC_API void rtcOccluded16(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRay16* ray);
RTC_API void rtcOccluded1M(RTCScene scene, struct RTCIntersectContext* context, struct RTCRay* ray, unsigned int M, size_t byteStride);
RTC_API void rtcOccluded1Mp(RTCScene scene, struct RTCIntersectContext* context, struct RTCRay** ray, unsigned int M);
RTC_API void rtcOccludedNM(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayN* ray, unsigned int N, unsigned int M, size_t byteStride);
RTC_API void rtcOccludedNp(RTCScene scene, struct RTCIntersectContext* context, const struct RTCRayNp* ray, unsigned int N);

#ifdef __cplusplus
inline RTCSceneFlags operator|(RTCSceneFlags a, RTCSceneFlags b) { return (RTCSceneFlags)((size_t)a | (size_t)b); }
#endif

#endif /* B200_RAYQUERY_RTCORE_H */
