/* Host memory probe for a possible "pack rays on the CPU before the H2D copy" stage: T threads gather the 32 useful bytes
 * (org, tnear, dir, tfar) of every 80-byte RTCRayHit record into a dense array.  Prints GB/s of source bytes consumed.
 * gcc -O3 -march=native -pthread tools/hostpack_probe.c -o /tmp/hostpack_probe && /tmp/hostpack_probe */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <immintrin.h>

static char* src; static char* dst; static size_t N; static int T;
static void* work(void* arg) {
  const size_t t = (size_t)arg, b = N * t / T, e = N * (t + 1) / T;
  for (size_t i = b; i < e; i++) {
    const char* s = src + i * 80; char* d = dst + i * 32;
    __m128 a = _mm_loadu_ps((const float*)s), c = _mm_loadu_ps((const float*)(s + 16));
    float tf = *(const float*)(s + 32);
    c = _mm_insert_ps(c, _mm_set_ss(tf), 0x30);            /* dir.xyz, tfar */
    _mm_stream_ps((float*)d, a); _mm_stream_ps((float*)(d + 16), c);
  }
  return 0;
}
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int main(void) {
  N = 16u << 20;
  src = aligned_alloc(64, N * 80); dst = aligned_alloc(64, N * 32);
  memset(src, 1, N * 80); memset(dst, 0, N * 32);
  const int ts[] = {1, 2, 4, 8, 12, 16};
  for (int k = 0; k < 6; k++) {
    T = ts[k]; double best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
      pthread_t th[16]; const double t0 = now();
      for (long t = 0; t < T; t++) pthread_create(&th[t], 0, work, (void*)t);
      for (int t = 0; t < T; t++) pthread_join(th[t], 0);
      const double dt = now() - t0; if (dt < best) best = dt;
    }
    printf("threads %2d: %.1f GB/s source (%.2f ms per 16.7M rays)\n", T, N * 80 / best / 1e9, best * 1e3);
  }
  return 0;
}
