/* CPU-baseline driver for the reference arm of bench.py.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE: nothing in the product links or calls this file.
 *
 * The reference library processes a stream call on the calling thread; parallelism over rays is the
 * application's job (tutorials/verify/verify.cpp:4671-4679 runs a parallel_for over tiles).  This driver
 * is that application: one pthread per host core, each pulling 4096-ray chunks from a shared cursor and
 * calling rtcIntersect1M / rtcOccluded1M of whatever rtcore library the function pointer comes from
 * (bench.py passes the address resolved from oracle/_ref/libembree3_ref.so).  No Python, no GIL, no
 * per-chunk interpreter overhead inside the timed region.  Threads are created before the clock starts
 * and released together by a flag; the clock stops when the last thread has drained the cursor.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>

struct Ctx { int flags; void* filter; unsigned instID[1]; };   /* RTCIntersectContext, rtcore_common.h:178-190 */
typedef void (*stream_fn)(void* scene, struct Ctx* ctx, void* rays, unsigned M, size_t stride);

struct Job {
  stream_fn fn; void* scene; char* rays; size_t n, stride; unsigned chunk; int flags;
  atomic_size_t cursor; atomic_int go;
};

static void* worker(void* p) {
  struct Job* j = (struct Job*)p;
  struct Ctx ctx; ctx.flags = j->flags; ctx.filter = NULL; ctx.instID[0] = 0xFFFFFFFFu;
  while (!atomic_load_explicit(&j->go, memory_order_acquire)) sched_yield();
  for (;;) {
    const size_t b = atomic_fetch_add(&j->cursor, (size_t)j->chunk);
    if (b >= j->n) break;
    const size_t m = j->n - b < j->chunk ? j->n - b : j->chunk;
    j->fn(j->scene, &ctx, j->rays + b * j->stride, (unsigned)m, j->stride);
  }
  return NULL;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

/* Traces n records of `stride` bytes with `threads` host threads in `chunk`-ray calls; returns seconds (< 0 on error). */
__attribute__((visibility("default")))
double cb_trace_stream(void* fn, void* scene, int coherent, void* rays, size_t n, size_t stride, unsigned chunk, int threads) {
  if (!fn || !scene || !rays || threads < 1 || chunk < 1) return -1.0;
  struct Job j;
  j.fn = (stream_fn)fn; j.scene = scene; j.rays = (char*)rays; j.n = n; j.stride = stride; j.chunk = chunk; j.flags = coherent ? 1 : 0;
  atomic_init(&j.cursor, 0); atomic_init(&j.go, 0);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  if (!th) return -1.0;
  int made = 0;
  for (; made < threads; made++) if (pthread_create(&th[made], NULL, worker, &j)) break;   /* fewer threads than asked for still drain the cursor */
  if (made == 0) { free(th); return -1.0; }
  const double t0 = now_s();
  atomic_store_explicit(&j.go, 1, memory_order_release);
  for (int k = 0; k < made; k++) pthread_join(th[k], NULL);
  const double t1 = now_s();
  free(th);
  return t1 - t0;
}
