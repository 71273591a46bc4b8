/* A plain C99 application of the rtcore API, compiled against include/embree3 and linked with -lembree3 exactly like a program
 * written for the reference library would be (tests/test_dropin_c.py builds and runs it).  It exercises the drop-in path end to
 * end from C -- shared and library-owned buffers, commit, rtcIntersect1M / rtcOccluded1M on malloc'ed ray arrays -- and ports the
 * idea of the reference's MemoryMonitorTest (tutorials/verify/verify.cpp:4564-4634): run once counting memory-monitor
 * invocations, then veto single invocations and require (a) an RTC_ERROR_OUT_OF_MEMORY, (b) no crash, (c) the monitored byte
 * balance back at zero once everything is released, (d) a clean commit afterwards.  With this engine the monitor also sees the
 * DEVICE memory of a commit (uploaded buffers, build scratch, BVH image).
 * Prints one line per check: "ok <name> ..." or "FAIL <name> ..."; exit code = number of failures. */
#include <embree3/rtcore.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int g_errors, g_lastError;
static long long g_balance, g_peak, g_calls, g_veto = -1;

static void onError(void* user, enum RTCError code, const char* str) { (void)user; (void)str; g_errors++; g_lastError = (int)code; }
static bool onMemory(void* user, ssize_t bytes, bool post) {
  (void)user; (void)post;
  const long long call = g_calls++;
  if (bytes > 0 && call == g_veto) return false;                 /* veto this allocation */
  g_balance += (long long)bytes;
  if (g_balance > g_peak) g_peak = g_balance;
  return true;
}

static RTCScene makeScene(RTCDevice dev, int n, float** keepVerts) {
  RTCScene scene = rtcNewScene(dev);
  /* geometry 0: (n x n)-cell plane at y = 0 in library-owned buffers */
  RTCGeometry g = rtcNewGeometry(dev, RTC_GEOMETRY_TYPE_TRIANGLE);
  float* v = (float*)rtcSetNewGeometryBuffer(g, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, 3 * sizeof(float), (size_t)(n + 1) * (n + 1));
  unsigned* t = (unsigned*)rtcSetNewGeometryBuffer(g, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT3, 3 * sizeof(unsigned), (size_t)2 * n * n);
  if (v && t) {
    for (int y = 0; y <= n; y++) for (int x = 0; x <= n; x++) { float* p = v + 3 * (y * (n + 1) + x); p[0] = -1.f + 2.f * x / n; p[1] = 0.f; p[2] = -1.f + 2.f * y / n; }
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) {
      unsigned a = y * (n + 1) + x, b = a + 1, c = a + n + 1, d = c + 1; unsigned* q = t + 6 * (y * n + x);
      q[0] = a; q[1] = b; q[2] = c; q[3] = d; q[4] = c; q[5] = b;
    }
  }
  rtcCommitGeometry(g); rtcAttachGeometry(scene, g); rtcReleaseGeometry(g);
  /* geometry 1: one big triangle above the plane, in a shared (application-owned) buffer with 16 bytes of tail padding */
  static const unsigned tri[3] = {0, 1, 2};
  float* sv = (float*)malloc(sizeof(float) * (9 + 4));
  const float init[9] = {-0.5f, 1.f, -0.5f, 0.5f, 1.f, -0.5f, 0.f, 1.f, 0.5f};
  memcpy(sv, init, sizeof(init));
  *keepVerts = sv;
  g = rtcNewGeometry(dev, RTC_GEOMETRY_TYPE_TRIANGLE);
  rtcSetSharedGeometryBuffer(g, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, sv, 0, 3 * sizeof(float), 3);
  rtcSetSharedGeometryBuffer(g, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT3, tri, 0, 3 * sizeof(unsigned), 1);
  rtcCommitGeometry(g); rtcAttachGeometry(scene, g); rtcReleaseGeometry(g);
  return scene;
}

static int check(int cond, const char* name, long long a, long long b) {
  printf("%s %s %lld %lld\n", cond ? "ok" : "FAIL", name, a, b);
  return cond ? 0 : 1;
}

int main(int argc, char** argv) {
  int fails = 0;
  const char* cfg = argc > 1 ? argv[1] : "";
  RTCDevice dev = rtcNewDevice(cfg);
  if (!dev) { printf("FAIL device %d 0\n", (int)rtcGetDeviceError(NULL)); return 1; }
  rtcSetDeviceErrorFunction(dev, onError, NULL);
  rtcSetDeviceMemoryMonitorFunction(dev, onMemory, NULL);

  /* ---- the stream path on malloc'ed rays ---- */
  float* keep = NULL;
  RTCScene scene = makeScene(dev, 96, &keep);
  rtcCommitScene(scene);
  fails += check(g_errors == 0, "commit", g_errors, g_lastError);
  const unsigned M = 300000;
  struct RTCRayHit* rh = (struct RTCRayHit*)malloc(sizeof(struct RTCRayHit) * M);
  struct RTCRay* sh = (struct RTCRay*)malloc(sizeof(struct RTCRay) * M);
  unsigned expectTri = 0;
  for (unsigned i = 0; i < M; i++) {
    const float x = -1.2f + 2.4f * (float)(i % 600) / 599.f, z = -1.2f + 2.4f * (float)(i / 600) / 499.f;
    struct RTCRayHit* r = rh + i;
    memset(r, 0, sizeof(*r));
    r->ray.org_x = x; r->ray.org_y = 2.f; r->ray.org_z = z; r->ray.dir_y = -1.f; r->ray.tfar = INFINITY; r->ray.mask = 0xFFFFFFFFu;
    r->hit.geomID = RTC_INVALID_GEOMETRY_ID; r->hit.primID = RTC_INVALID_GEOMETRY_ID; r->hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
    sh[i] = r->ray;
    (void)expectTri;
  }
  struct RTCIntersectContext ctx;
  rtcInitIntersectContext(&ctx);
  rtcIntersect1M(scene, &ctx, rh, M, sizeof(struct RTCRayHit));
  rtcOccluded1M(scene, &ctx, sh, M, sizeof(struct RTCRay));
  unsigned hitPlane = 0, hitTri = 0, miss = 0, bad = 0, occl = 0;
  for (unsigned i = 0; i < M; i++) {
    const struct RTCRayHit* r = rh + i;
    const float ax = fabsf(r->ray.org_x), az = fabsf(r->ray.org_z), am = ax > az ? ax : az;
    const int inside = am < 0.998f, outside = am > 1.002f;         /* rays within 0.002 of the plane's border may go either way */
    if (r->hit.geomID == 1) { hitTri++; if (fabsf(r->ray.tfar - 1.f) > 1e-5f) bad++; }
    else if (r->hit.geomID == 0) { hitPlane++; if (fabsf(r->ray.tfar - 2.f) > 1e-5f || outside || r->hit.primID >= 2u * 96u * 96u) bad++; }
    else { miss++; if (inside || r->ray.tfar != INFINITY) bad++; }
    if (sh[i].tfar == -INFINITY) occl++;
    if ((sh[i].tfar == -INFINITY) != (r->hit.geomID != RTC_INVALID_GEOMETRY_ID)) bad++;
  }
  fails += check(bad == 0 && hitTri > 1000 && hitPlane > 100000 && miss > 10000, "stream", hitPlane + hitTri, miss);
  fails += check(occl == hitPlane + hitTri, "occluded", occl, hitPlane + hitTri);
  fails += check(g_errors == 0 && rtcGetDeviceError(dev) == RTC_ERROR_NONE, "noerror", g_errors, g_lastError);
  free(rh); free(sh);
  rtcReleaseScene(scene); free(keep);
  fails += check(g_balance == 0, "balance_after_release", g_balance, g_peak);

  /* ---- memory monitor as fault injection ---- */
  g_calls = 0; g_peak = 0;
  scene = makeScene(dev, 64, &keep);
  rtcCommitScene(scene);
  rtcReleaseScene(scene); free(keep);
  const long long total = g_calls;
  fails += check(total > 8 && g_balance == 0 && g_errors == 0, "monitor_count", total, g_peak);
  int vetoed = 0;
  for (long long k = 0; k < total; k += (total > 40 ? total / 20 : 1)) {
    g_calls = 0; g_errors = 0; g_lastError = 0; g_veto = k;
    scene = makeScene(dev, 64, &keep);
    rtcCommitScene(scene);
    const int err = (int)rtcGetDeviceError(dev);
    g_veto = -1;
    if (g_errors > 0) vetoed++;
    if (g_errors > 0 && (g_lastError != RTC_ERROR_OUT_OF_MEMORY && err != RTC_ERROR_OUT_OF_MEMORY)) fails += check(0, "veto_error_code", k, err);
    rtcCommitScene(scene);                                        /* the failed commit is retried, not skipped */
    struct RTCRayHit one; memset(&one, 0, sizeof(one));
    one.ray.org_y = 2.f; one.ray.dir_y = -1.f; one.ray.tfar = INFINITY; one.hit.geomID = RTC_INVALID_GEOMETRY_ID;
    rtcIntersect1(scene, &ctx, &one);
    if (one.hit.geomID != 1 || rtcGetDeviceError(dev) != RTC_ERROR_NONE) fails += check(0, "recommit_after_veto", k, one.hit.geomID);
    rtcReleaseScene(scene); free(keep);
    if (g_balance != 0) { fails += check(0, "veto_balance", k, g_balance); g_balance = 0; }
  }
  fails += check(vetoed > 0, "vetoes_seen", vetoed, total);
  rtcReleaseDevice(dev);
  printf("%s done %d 0\n", fails ? "FAIL" : "ok", fails);
  return fails;
}
