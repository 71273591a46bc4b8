/* rq_oracle.c -- CPU restatement (plain C99, single-threaded) of the reference's algorithm for the
 * hot path: triangle BVH8 build with binned SAH, single-ray traversal, Moeller-Trumbore / Pluecker
 * triangle tests, the AoS ray-stream entry/exit rules, and the SAH statistic.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker used by tests/, by
 * __graft_entry__.smoke() and (as a fallback when oracle/_ref is absent) by bench.py's CPU
 * baseline.  Nothing under embree-aarch64_b200/ includes, links, loads or calls it.
 * Parity status: PINNED -- tests/test_oracle.py checks it against (a) the reference's own
 * known-answer tests (TriangleHitTest analytic expectations, tutorials/verify/verify.cpp:2339-2426),
 * (b) golden vectors produced by the real reference library (tests/golden/, generator
 * tests/golden/make_golden.py, library built by oracle/build_ref.py), (c) oracle/_ref live when present.
 * The quad-mesh and single-level-instancing restatements further down are pinned the same way
 * (tests/golden/quads_*.npz, inst_*.npz; generators make_golden_quads.py, make_golden_instances.py).
 *
 * Each function cites the reference source it restates (paths relative to /root/reference).
 * Differences that are deliberate and inside the stated tolerances:
 *   - t,u,v are divided exactly; the reference multiplies by rcp() = estimate + 1 Newton step
 *     (common/simd/vfloat8_avx.h:280-303), 1-2 ulp away.
 *   - rays are processed one at a time (the reference's packet/hybrid kernels drop to the same
 *     single-ray code for incoherent rays: kernels/bvh/bvh_intersector_hybrid.cpp:237-255).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define RQO_API __attribute__((visibility("default")))
#define INVALID_ID 0xFFFFFFFFu
#define FLT_LARGE_ 1.844E18f                 /* common/math/constants.h:34 */
#define MIN_RCP_INPUT 1E-18f                 /* common/math/constants.h:31 */
#define ULP_ 1.1920928955078125e-07f         /* std::numeric_limits<float>::epsilon(), constants.h:136 */

typedef struct { float x, y, z; } v3;
typedef struct { v3 lo, hi; } box3;

/* ---- common/math/vec3.h:204,209 and math.h:304-322,367-373 (FMA forms of the AVX2 targets) ---- */
static inline float madd(float a, float b, float c) { return fmaf(a, b, c); }
static inline float msub(float a, float b, float c) { return fmaf(a, b, -c); }
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float dot3(v3 a, v3 b) { return madd(a.x, b.x, madd(a.y, b.y, a.z * b.z)); }
static inline v3 cross3(v3 a, v3 b) {
  return V(msub(a.y, b.z, a.z * b.y), msub(a.z, b.x, a.x * b.z), msub(a.x, b.y, a.y * b.x));
}
static inline float xorsign(float a, uint32_t s) { union { float f; uint32_t u; } c; c.f = a; c.u ^= s; return c.f; }
static inline uint32_t signmsk(float a) { union { float f; uint32_t u; } c; c.f = a; return c.u & 0x80000000u; }
static inline float half_area(v3 d) { return madd(d.x, (d.y + d.z), d.y * d.z); }       /* vec3.h halfArea */
static inline box3 box_empty(void) { box3 b = {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; return b; }
static inline void box_extend(box3* b, const box3* o) {
  b->lo.x = fminf(b->lo.x, o->lo.x); b->lo.y = fminf(b->lo.y, o->lo.y); b->lo.z = fminf(b->lo.z, o->lo.z);
  b->hi.x = fmaxf(b->hi.x, o->hi.x); b->hi.y = fmaxf(b->hi.y, o->hi.y); b->hi.z = fmaxf(b->hi.z, o->hi.z);
}
static inline float box_half_area(const box3* b) { return half_area(vsub(b->hi, b->lo)); }
static inline float comp(v3 a, int d) { return d == 0 ? a.x : d == 1 ? a.y : a.z; }

/* ------------------------------------------------------------------------------------------ */
typedef struct { v3 v0, v1, v2; uint32_t primID, geomID, flip; } tri_t;   /* flip: second triangle (v2,v3,v1) of a quad */
typedef struct { box3 b; uint32_t tri; } primref_t;                 /* kernels/common/primref.h:11-105 */

typedef struct node_s {
  int nchild;                  /* 0 = leaf */
  box3 cbox[8];
  struct node_s* child[8];
  uint32_t first, count;       /* leaf: range in the ordered triangle array */
} node_t;

typedef struct {
  tri_t* tris; uint32_t ntris;                 /* valid triangles, later in leaf order */
  node_t* root; box3 bounds;
  int robust;
  double sah_inner, sah_leaf; uint64_t nnodes, nleaves, nblocks;
} scene_t;

/* kernels/common/scene_triangle_mesh.h:131-153: index range + |v| < FLT_LARGE (NaN fails) */
static int vertex_valid(v3 p) {
  return p.x > -FLT_LARGE_ && p.x < FLT_LARGE_ && p.y > -FLT_LARGE_ && p.y < FLT_LARGE_ && p.z > -FLT_LARGE_ && p.z < FLT_LARGE_;
}

/* ---- binned SAH, kernels/builders/heuristic_binning.h:15-114 (mapping), :210-256 (bin), :336-392 (best) ---- */
#define BINS 32
typedef struct { int dim, pos; float sah; size_t num; v3 ofs, scale; } split_t;

static split_t find_split(const primref_t* pr, uint32_t begin, uint32_t end, const box3* centBounds) {
  split_t s; s.dim = -1; s.pos = 0; s.sah = INFINITY;
  const size_t N = end - begin;
  size_t num = (size_t)(4.0f + 0.05f * (float)N); if (num > BINS) num = BINS;
  s.num = num;
  const v3 size = vsub(centBounds->hi, centBounds->lo);
  const float eps = 1E-34f;
  float diag[3] = {fmaxf(eps, size.x), fmaxf(eps, size.y), fmaxf(eps, size.z)}, scale[3];
  for (int d = 0; d < 3; d++) scale[d] = diag[d] > eps ? (0.99f * (float)num) / diag[d] : 0.0f;
  s.ofs = centBounds->lo; s.scale = V(scale[0], scale[1], scale[2]);
  static box3 bb[BINS][3]; static uint32_t cnt[BINS][3];
  for (size_t i = 0; i < num; i++) for (int d = 0; d < 3; d++) { bb[i][d] = box_empty(); cnt[i][d] = 0; }
  for (uint32_t i = begin; i < end; i++) {
    const v3 c2 = vadd(pr[i].b.lo, pr[i].b.hi);                      /* center2 = lower+upper */
    const int bx = (int)floorf((c2.x - s.ofs.x) * scale[0]), by = (int)floorf((c2.y - s.ofs.y) * scale[1]),
              bz = (int)floorf((c2.z - s.ofs.z) * scale[2]);
    const int b[3] = {bx < 0 ? 0 : bx >= (int)num ? (int)num - 1 : bx, by < 0 ? 0 : by >= (int)num ? (int)num - 1 : by,
                      bz < 0 ? 0 : bz >= (int)num ? (int)num - 1 : bz};
    for (int d = 0; d < 3; d++) { box_extend(&bb[b[d]][d], &pr[i].b); cnt[b[d]][d]++; }
  }
  float rArea[BINS][3]; uint32_t rCnt[BINS][3];
  for (int d = 0; d < 3; d++) {
    box3 bx = box_empty(); uint32_t c = 0;
    for (size_t i = num - 1; i > 0; i--) { c += cnt[i][d]; rCnt[i][d] = c; box_extend(&bx, &bb[i][d]); rArea[i][d] = box_half_area(&bx); }
    bx = box_empty(); c = 0;
    float best = INFINITY; int bestPos = 0;
    for (size_t i = 1; i < num; i++) {
      c += cnt[i - 1][d]; box_extend(&bx, &bb[i - 1][d]);
      const float lA = box_half_area(&bx), rA = rArea[i][d];
      const uint32_t lC = (c + 3) >> 2, rC = (rCnt[i][d] + 3) >> 2;        /* blocks of 4: bvh_builder_sah.cpp:565 */
      const float sah = madd(lA, (float)lC, rA * (float)rC);
      if (sah < best) { best = sah; bestPos = (int)i; }
    }
    if (scale[d] == 0.0f) continue;                                    /* mapping.invalid(dim) */
    if (best < s.sah && bestPos != 0) { s.dim = d; s.pos = bestPos; s.sah = best; }
  }
  return s;
}

static int cmp_primref(const void* a, const void* b) {                 /* deterministic_order: by (primID, geomID) via tri index */
  const uint32_t x = ((const primref_t*)a)->tri, y = ((const primref_t*)b)->tri;
  return x < y ? -1 : x > y;
}

typedef struct { uint32_t begin, end; box3 geom, cent; int depth; } rec_t;

static void rec_bounds(const primref_t* pr, rec_t* r) {
  r->geom = box_empty(); r->cent = box_empty();
  for (uint32_t i = r->begin; i < r->end; i++) {
    box_extend(&r->geom, &pr[i].b);
    const v3 c2 = vadd(pr[i].b.lo, pr[i].b.hi); box3 cb = {c2, c2}; box_extend(&r->cent, &cb);
  }
}

/* heuristic_binning_array_aligned.h:79-123 (split by bin position) and :125-160 (splitFallback = median) */
static uint32_t do_split(primref_t* pr, const rec_t* r, const split_t* s) {
  if (s->dim < 0) return (r->begin + r->end) / 2;
  uint32_t i = r->begin, j = r->end;
  const float ofs = comp(s->ofs, s->dim), sc = comp(s->scale, s->dim);
  while (i < j) {
    const v3 c2 = vadd(pr[i].b.lo, pr[i].b.hi);
    if ((int)floorf((comp(c2, s->dim) - ofs) * sc) < s->pos) i++;
    else { j--; primref_t t = pr[i]; pr[i] = pr[j]; pr[j] = t; }
  }
  if (i == r->begin || i == r->end) return (r->begin + r->end) / 2;
  return i;
}

/* kernels/builders/bvh_builder_sah.h:155-220 createLargeLeaf, :222-319 recurse;
 * settings kernels/bvh/bvh_builder_sah.cpp:565-566: branching 8, blocks of 4, minLeaf 4, maxLeaf 28, travCost 1, intCost 1 */
#define MIN_LEAF 4
#define MAX_LEAF 28
static node_t* build_rec(scene_t* sc, primref_t* pr, rec_t cur, tri_t* ordered, const tri_t* src, uint32_t* cursor);

static node_t* make_leaf(scene_t* sc, primref_t* pr, rec_t cur, tri_t* ordered, const tri_t* src, uint32_t* cursor) {
  const uint32_t n = cur.end - cur.begin;
  if (n <= MAX_LEAF) {
    node_t* leaf = (node_t*)calloc(1, sizeof(node_t));
    qsort(pr + cur.begin, n, sizeof(primref_t), cmp_primref);
    leaf->first = *cursor; leaf->count = n;
    for (uint32_t i = 0; i < n; i++) ordered[(*cursor)++] = src[pr[cur.begin + i].tri];
    return leaf;
  }
  /* too many primitives for one leaf: median splits into up to 8 children */
  rec_t ch[8]; int nc = 1; ch[0] = cur;
  while (nc < 8) {
    int best = -1; uint32_t bs = MAX_LEAF;
    for (int i = 0; i < nc; i++) if (ch[i].end - ch[i].begin > bs) { bs = ch[i].end - ch[i].begin; best = i; }
    if (best < 0) break;
    rec_t l = ch[best], r = ch[best]; const uint32_t mid = (l.begin + l.end) / 2;
    l.end = mid; r.begin = mid; rec_bounds(pr, &l); rec_bounds(pr, &r);
    ch[best] = l; ch[nc++] = r;
  }
  node_t* nd = (node_t*)calloc(1, sizeof(node_t));
  nd->nchild = nc;
  for (int i = 0; i < nc; i++) { nd->cbox[i] = ch[i].geom; ch[i].depth = cur.depth + 1; nd->child[i] = make_leaf(sc, pr, ch[i], ordered, src, cursor); }
  return nd;
}

static node_t* build_rec(scene_t* sc, primref_t* pr, rec_t cur, tri_t* ordered, const tri_t* src, uint32_t* cursor) {
  const uint32_t n = cur.end - cur.begin;
  split_t sp = find_split(pr, cur.begin, cur.end, &cur.cent);
  const float leafSAH = box_half_area(&cur.geom) * (float)((n + 3) >> 2);
  const float splitSAH = 1.0f * box_half_area(&cur.geom) + 1.0f * sp.sah;
  if (n <= MIN_LEAF || cur.depth + 8 >= 40 || (n <= MAX_LEAF && leafSAH <= splitSAH))
    return make_leaf(sc, pr, cur, ordered, src, cursor);
  rec_t ch[8]; int nc = 2;
  {
    const uint32_t mid = do_split(pr, &cur, &sp);
    ch[0] = cur; ch[0].end = mid; ch[1] = cur; ch[1].begin = mid;
    ch[0].depth = ch[1].depth = cur.depth + 1; rec_bounds(pr, &ch[0]); rec_bounds(pr, &ch[1]);
  }
  while (nc < 8) {                                               /* widen: split the child with the largest half area */
    float bestA = -INFINITY; int best = -1;
    for (int i = 0; i < nc; i++) {
      if (ch[i].end - ch[i].begin <= MIN_LEAF) continue;
      const float a = box_half_area(&ch[i].geom);
      if (a > bestA) { bestA = a; best = i; }
    }
    if (best < 0) break;
    split_t s2 = find_split(pr, ch[best].begin, ch[best].end, &ch[best].cent);
    const uint32_t mid = do_split(pr, &ch[best], &s2);
    rec_t l = ch[best], r = ch[best]; l.end = mid; r.begin = mid; rec_bounds(pr, &l); rec_bounds(pr, &r);
    ch[best] = l; ch[nc++] = r;
  }
  for (int i = 1; i < nc; i++) {                                  /* children sorted by size, descending (:293) */
    rec_t k = ch[i]; int j = i - 1;
    while (j >= 0 && (ch[j].end - ch[j].begin) < (k.end - k.begin)) { ch[j + 1] = ch[j]; j--; }
    ch[j + 1] = k;
  }
  node_t* nd = (node_t*)calloc(1, sizeof(node_t));
  nd->nchild = nc;
  for (int i = 0; i < nc; i++) { nd->cbox[i] = ch[i].geom; nd->child[i] = build_rec(sc, pr, ch[i], ordered, src, cursor); }
  return nd;
}

/* kernels/bvh/bvh_statistics.cpp:41-160, bvh_statistics.h:36-38,99-101:
 * sah = [ sum_inner halfArea(node box) + sum_leaf halfArea(leaf box) * numBlocks ] / halfArea(root box) */
static void stat_rec(scene_t* sc, const node_t* nd, const box3* b) {
  if (nd->nchild == 0) { sc->nleaves++; const uint32_t blocks = (nd->count + 3) / 4; sc->nblocks += blocks; sc->sah_leaf += (double)box_half_area(b) * blocks; return; }
  sc->nnodes++; sc->sah_inner += (double)box_half_area(b);
  for (int i = 0; i < nd->nchild; i++) stat_rec(sc, nd->child[i], &nd->cbox[i]);
}
static void free_rec(node_t* nd) { if (!nd) return; for (int i = 0; i < nd->nchild; i++) free_rec(nd->child[i]); free(nd); }

/* quads != 0: RTC_GEOMETRY_TYPE_QUAD, `indices` holds numTris records of FOUR vertex indices; a quad is intersected as the
 * triangles (v0,v1,v3) and (v2,v3,v1), the second reporting u = 1 - u, v = 1 - v (kernels/geometry/quad_intersector_moeller.h:122-144,
 * quad_intersector_pluecker.h:179-180; hit finalisation :28-37 / :34-43), and is dropped as a whole unless all four vertices are
 * valid (kernels/common/scene_quad_mesh.h:131-154) */
typedef struct { const void* indices; const void* vertices; uint32_t indexStride, vertexStride, numTris, numVerts, geomID, quads; } rqo_mesh;

RQO_API void* rqo_build(const rqo_mesh* meshes, int nmeshes, int robust) {
  scene_t* sc = (scene_t*)calloc(1, sizeof(scene_t));
  sc->robust = robust; sc->bounds = box_empty();
  uint64_t total = 0;
  for (int m = 0; m < nmeshes; m++) total += (uint64_t)meshes[m].numTris * (meshes[m].quads ? 2 : 1);
  tri_t* src = (tri_t*)malloc(sizeof(tri_t) * (total ? total : 1));
  primref_t* pr = (primref_t*)malloc(sizeof(primref_t) * (total ? total : 1));
  uint32_t n = 0;
  for (int m = 0; m < nmeshes; m++) {                            /* builders/primrefgen.cpp:35-57 */
    const rqo_mesh* M = &meshes[m];
    for (uint32_t i = 0; i < M->numTris && M->quads; i++) {
      const uint32_t* ix = (const uint32_t*)((const char*)M->indices + (size_t)i * M->indexStride);
      if (ix[0] >= M->numVerts || ix[1] >= M->numVerts || ix[2] >= M->numVerts || ix[3] >= M->numVerts) continue;
      v3 q[4]; int okq = 1;
      for (int k = 0; k < 4; k++) {
        const float* p = (const float*)((const char*)M->vertices + (size_t)ix[k] * M->vertexStride);
        q[k] = V(p[0], p[1], p[2]); okq &= vertex_valid(q[k]);
      }
      if (!okq) continue;
      for (int half = 0; half < 2; half++) {
        tri_t t; t.primID = i; t.geomID = M->geomID; t.flip = (uint32_t)half;
        if (half == 0) { t.v0 = q[0]; t.v1 = q[1]; t.v2 = q[3]; } else { t.v0 = q[2]; t.v1 = q[3]; t.v2 = q[1]; }
        box3 bx; bx.lo = V(fminf(fminf(t.v0.x, t.v1.x), t.v2.x), fminf(fminf(t.v0.y, t.v1.y), t.v2.y), fminf(fminf(t.v0.z, t.v1.z), t.v2.z));
        bx.hi = V(fmaxf(fmaxf(t.v0.x, t.v1.x), t.v2.x), fmaxf(fmaxf(t.v0.y, t.v1.y), t.v2.y), fmaxf(fmaxf(t.v0.z, t.v1.z), t.v2.z));
        src[n] = t; pr[n].b = bx; pr[n].tri = n; n++;
      }
    }
    for (uint32_t i = 0; i < M->numTris && !M->quads; i++) {
      const uint32_t* ix = (const uint32_t*)((const char*)M->indices + (size_t)i * M->indexStride);
      if (ix[0] >= M->numVerts || ix[1] >= M->numVerts || ix[2] >= M->numVerts) continue;
      const float* a = (const float*)((const char*)M->vertices + (size_t)ix[0] * M->vertexStride);
      const float* b = (const float*)((const char*)M->vertices + (size_t)ix[1] * M->vertexStride);
      const float* c = (const float*)((const char*)M->vertices + (size_t)ix[2] * M->vertexStride);
      tri_t t; t.v0 = V(a[0], a[1], a[2]); t.v1 = V(b[0], b[1], b[2]); t.v2 = V(c[0], c[1], c[2]); t.primID = i; t.geomID = M->geomID; t.flip = 0;
      if (!vertex_valid(t.v0) || !vertex_valid(t.v1) || !vertex_valid(t.v2)) continue;
      box3 bx; bx.lo = V(fminf(fminf(t.v0.x, t.v1.x), t.v2.x), fminf(fminf(t.v0.y, t.v1.y), t.v2.y), fminf(fminf(t.v0.z, t.v1.z), t.v2.z));
      bx.hi = V(fmaxf(fmaxf(t.v0.x, t.v1.x), t.v2.x), fmaxf(fmaxf(t.v0.y, t.v1.y), t.v2.y), fmaxf(fmaxf(t.v0.z, t.v1.z), t.v2.z));
      src[n] = t; pr[n].b = bx; pr[n].tri = n; n++;
    }
  }
  sc->ntris = n;
  sc->tris = (tri_t*)malloc(sizeof(tri_t) * (n ? n : 1));
  if (n) {
    rec_t root; root.begin = 0; root.end = n; root.depth = 1; rec_bounds(pr, &root);
    sc->bounds = root.geom;
    uint32_t cursor = 0;
    sc->root = build_rec(sc, pr, root, sc->tris, src, &cursor);
    stat_rec(sc, sc->root, &sc->bounds);
  }
  free(src); free(pr);
  return sc;
}
RQO_API void rqo_free(void* h) { scene_t* sc = (scene_t*)h; if (!sc) return; free_rec(sc->root); free(sc->tris); free(sc); }
RQO_API double rqo_sah(void* h) { scene_t* sc = (scene_t*)h; const double A = box_half_area(&sc->bounds); return A > 0 ? (sc->sah_inner + sc->sah_leaf) / A : 0.0; }
RQO_API void rqo_stats(void* h, uint64_t out[4]) { scene_t* sc = (scene_t*)h; out[0] = sc->ntris; out[1] = sc->nnodes; out[2] = sc->nleaves; out[3] = sc->nblocks; }
RQO_API void rqo_bounds(void* h, float out[6]) { scene_t* sc = (scene_t*)h; out[0] = sc->bounds.lo.x; out[1] = sc->bounds.lo.y; out[2] = sc->bounds.lo.z; out[3] = sc->bounds.hi.x; out[4] = sc->bounds.hi.y; out[5] = sc->bounds.hi.z; }

/* ------------------------------------------------------------------------------------------ */
typedef struct { float t, u, v; v3 Ng; } hit_t;

/* kernels/geometry/triangle_intersector_moeller.h:62-103 (+ :30-36 finalize) */
RQO_API int rqo_moeller(const float O_[3], const float D_[3], float tnear, float tfar, const float a[3], const float b[3], const float c[3], float out[6]) {
  const v3 O = V(O_[0], O_[1], O_[2]), D = V(D_[0], D_[1], D_[2]);
  const v3 v0 = V(a[0], a[1], a[2]), v1 = V(b[0], b[1], b[2]), v2 = V(c[0], c[1], c[2]);
  const v3 e1 = vsub(v0, v1), e2 = vsub(v2, v0);                   /* triangle.h:45-51 */
  const v3 Ng = cross3(e2, e1);
  const v3 C = vsub(v0, O);
  const v3 R = cross3(C, D);
  const float den = dot3(Ng, D);
  const float absDen = fabsf(den);
  const uint32_t sgn = signmsk(den);
  const float U = xorsign(dot3(R, e2), sgn);
  const float Vv = xorsign(dot3(R, e1), sgn);
  if (!((den != 0.0f) & (U >= 0.0f) & (Vv >= 0.0f) & (U + Vv <= absDen))) return 0;
  const float T = xorsign(dot3(Ng, C), sgn);
  if (!((absDen * tnear < T) & (T <= absDen * tfar))) return 0;
  out[0] = T / absDen; out[1] = U / absDen; out[2] = Vv / absDen; out[3] = Ng.x; out[4] = Ng.y; out[5] = Ng.z;
  return 1;
}

/* common/math/vec3.h:210-222 */
static v3 stable_triangle_normal(v3 a, v3 b, v3 c) {
  const float ab_x = a.z * b.y, ab_y = a.x * b.z, ab_z = a.y * b.x;
  const float bc_x = b.z * c.y, bc_y = b.x * c.z, bc_z = b.y * c.x;
  const v3 cab = V(msub(a.y, b.z, ab_x), msub(a.z, b.x, ab_y), msub(a.x, b.y, ab_z));
  const v3 cbc = V(msub(b.y, c.z, bc_x), msub(b.z, c.x, bc_y), msub(b.x, c.y, bc_z));
  return V(fabsf(ab_x) < fabsf(bc_x) ? cab.x : cbc.x, fabsf(ab_y) < fabsf(bc_y) ? cab.y : cbc.y, fabsf(ab_z) < fabsf(bc_z) ? cab.z : cbc.z);
}

/* kernels/geometry/triangle_intersector_pluecker.h:61-108 (+ :25-32 finalize) */
RQO_API int rqo_pluecker(const float O_[3], const float D_[3], float tnear, float tfar, const float a[3], const float b[3], const float c[3], float out[6]) {
  const v3 O = V(O_[0], O_[1], O_[2]), D = V(D_[0], D_[1], D_[2]);
  const v3 v0 = vsub(V(a[0], a[1], a[2]), O), v1 = vsub(V(b[0], b[1], b[2]), O), v2 = vsub(V(c[0], c[1], c[2]), O);
  const v3 e0 = vsub(v2, v0), e1 = vsub(v0, v1), e2 = vsub(v1, v2);
  const float U = dot3(cross3(e0, vadd(v2, v0)), D);
  const float Vv = dot3(cross3(e1, vadd(v0, v1)), D);
  const float W = dot3(cross3(e2, vadd(v1, v2)), D);
  const float UVW = U + Vv + W;
  const float eps = ULP_ * fabsf(UVW);
  if (!((fminf(fminf(U, Vv), W) >= -eps) | (fmaxf(fmaxf(U, Vv), W) <= eps))) return 0;
  const v3 Ng = stable_triangle_normal(e0, e1, e2);
  const float dn = dot3(Ng, D), den = dn + dn;
  const float tn = dot3(v0, Ng), T = tn + tn;
  const float t = T / den;
  if (!((tnear <= t) & (t <= tfar) & (den != 0.0f))) return 0;
  const int tiny = fabsf(UVW) < MIN_RCP_INPUT;
  out[0] = t; out[1] = tiny ? 0.0f : U / UVW; out[2] = tiny ? 0.0f : Vv / UVW; out[3] = Ng.x; out[4] = Ng.y; out[5] = Ng.z;
  return 1;
}

typedef struct { float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar; uint32_t mask, id, flags; } ray_t;
typedef struct { float Ng_x, Ng_y, Ng_z, u, v; uint32_t primID, geomID, instID; } rhit_t;

static inline float rcp_safe(float d) { return 1.0f / (fabsf(d) < MIN_RCP_INPUT ? MIN_RCP_INPUT : d); }   /* vec3fa.h:172-177 */

/* kernels/bvh/bvh_intersector1.cpp:30-119 (closest hit) and :121-202 (any hit);
 * slab test kernels/bvh/node_intersector1.h:527-578: t = fmsub(plane, rdir, org*rdir), hit if max(tNear*,tnear) <= min(tFar*,tfar);
 * RTC_SCENE_FLAG_ROBUST selects intersectNodeRobust (:621-636): t = (plane - org) * rdir_{near,far} with
 * rdir_near = (1-3ulp)*rdir, rdir_far = (1+3ulp)*rdir (TravRayBase<N,Nx,true>, :121-135);
 * closest-hit child order: nearest first (bvh_traverser1.h:519-643); cull popped nodes with dist > tfar (:92-96) */
/* work counters of the restated reference traversal (what EMBREE_STAT_COUNTERS counts per ray, kernels/common/stat.h:61-79):
 * [0] rays traced, [1] inner nodes visited (8-wide slab tests), [2] leaves visited, [3] Triangle4 blocks tested, [4] triangles tested */
static uint64_t g_cnt[5];
RQO_API void rqo_trace_counters(uint64_t out[5], int reset) { for (int i = 0; i < 5; i++) { out[i] = g_cnt[i]; if (reset) g_cnt[i] = 0; } }

static int trace_one(const scene_t* sc, ray_t* ray, rhit_t* hit, int occluded, uint32_t instID) {
  if (!sc->root) return 0;
  g_cnt[0]++;
  const v3 O = V(ray->org_x, ray->org_y, ray->org_z), D = V(ray->dir_x, ray->dir_y, ray->dir_z);
  const float Of[3] = {O.x, O.y, O.z}, Df[3] = {D.x, D.y, D.z};
  const v3 rdir = V(rcp_safe(D.x), rcp_safe(D.y), rcp_safe(D.z));
  const v3 ordir = V(O.x * rdir.x, O.y * rdir.y, O.z * rdir.z);
  const float tnearBox = fmaxf(ray->tnear, 0.0f);
  struct { const node_t* n; float d; } stack[256]; int sp = 0;
  stack[sp].n = sc->root; stack[sp++].d = -INFINITY;
  int found = 0;
  while (sp) {
    sp--;
    const node_t* nd = stack[sp].n;
    if (!occluded && stack[sp].d > ray->tfar) continue;
    if (nd->nchild == 0) {
      g_cnt[2]++; g_cnt[3] += (nd->count + 3) / 4; g_cnt[4] += nd->count;
      for (uint32_t i = 0; i < nd->count; i++) {
        const tri_t* t = &sc->tris[nd->first + i];
        float o[6];
        const int ok = sc->robust ? rqo_pluecker(Of, Df, ray->tnear, ray->tfar, &t->v0.x, &t->v1.x, &t->v2.x, o)
                                  : rqo_moeller(Of, Df, ray->tnear, ray->tfar, &t->v0.x, &t->v1.x, &t->v2.x, o);
        if (!ok) continue;
        found = 1;
        if (occluded) return 1;
        /* epilog, kernels/geometry/intersector_epilog.h:280-290 */
        ray->tfar = o[0]; hit->u = o[1]; hit->v = o[2]; hit->Ng_x = o[3]; hit->Ng_y = o[4]; hit->Ng_z = o[5];
        if (t->flip) { hit->u = 1.0f - fminf(hit->u, 1.0f); hit->v = 1.0f - fminf(hit->v, 1.0f); }   /* quad_intersector_moeller.h:28-37 */
        hit->primID = t->primID; hit->geomID = t->geomID; hit->instID = instID;
      }
      continue;
    }
    const float tfarBox = fmaxf(ray->tfar, 0.0f);
    g_cnt[1]++;
    int idx[8]; float dist[8]; int nh = 0;
    for (int i = 0; i < nd->nchild; i++) {
      const box3* b = &nd->cbox[i];
      const float nx = rdir.x >= 0 ? b->lo.x : b->hi.x, fx = rdir.x >= 0 ? b->hi.x : b->lo.x;
      const float ny = rdir.y >= 0 ? b->lo.y : b->hi.y, fy = rdir.y >= 0 ? b->hi.y : b->lo.y;
      const float nz = rdir.z >= 0 ? b->lo.z : b->hi.z, fz = rdir.z >= 0 ? b->hi.z : b->lo.z;
      float tN, tF;
      if (sc->robust) {
        const float dn = 1.0f - 3.0f * ULP_, up = 1.0f + 3.0f * ULP_;
        tN = fmaxf(fmaxf((nx - O.x) * (dn * rdir.x), (ny - O.y) * (dn * rdir.y)), fmaxf((nz - O.z) * (dn * rdir.z), tnearBox));
        tF = fminf(fminf((fx - O.x) * (up * rdir.x), (fy - O.y) * (up * rdir.y)), fminf((fz - O.z) * (up * rdir.z), tfarBox));
      } else {
        tN = fmaxf(fmaxf(msub(nx, rdir.x, ordir.x), msub(ny, rdir.y, ordir.y)), fmaxf(msub(nz, rdir.z, ordir.z), tnearBox));
        tF = fminf(fminf(msub(fx, rdir.x, ordir.x), msub(fy, rdir.y, ordir.y)), fminf(msub(fz, rdir.z, ordir.z), tfarBox));
      }
      if (tN <= tF) { idx[nh] = i; dist[nh] = tN; nh++; }
    }
    for (int i = 1; i < nh; i++) {                                /* far children first on the stack => nearest popped first */
      const int k = idx[i]; const float d = dist[i]; int j = i - 1;
      while (j >= 0 && dist[j] < d) { idx[j + 1] = idx[j]; dist[j + 1] = dist[j]; j--; }
      idx[j + 1] = k; dist[j + 1] = d;
    }
    for (int i = 0; i < nh && sp < 256; i++) { stack[sp].n = nd->child[idx[i]]; stack[sp++].d = dist[i]; }
  }
  return found;
}

/* rtcIntersect1M: kernels/common/rtcore.cpp:595-624; stream filter kernels/bvh/bvh_intersector_stream_filters.cpp:135-151;
 * hit scatter kernels/common/ray.h:1119-1160 (only lanes with geomID != -1 are written) */
RQO_API void rqo_intersect1M(void* h, void* rayhit, uint32_t M, size_t stride, uint32_t instID) {
  const scene_t* sc = (const scene_t*)h;
  for (uint32_t i = 0; i < M; i++) {
    ray_t* r = (ray_t*)((char*)rayhit + (size_t)i * stride);
    rhit_t* ht = (rhit_t*)((char*)r + 48);
    if (!(r->tnear <= r->tfar)) continue;                          /* inactive: untouched */
    ray_t tmp = *r; rhit_t th; memset(&th, 0, sizeof(th));
    if (trace_one(sc, &tmp, &th, 0, instID)) { r->tfar = tmp.tfar; *ht = th; }
  }
}

/* rtcOccluded1M: rtcore.cpp:848-875; stream rules stream_filters.cpp:78 (skip tnear>tfar or tfar<0) and
 * bvh_intersector_stream.cpp:303-305 (tnear<0 invalid) for M>1; single-ray rules bvh_intersector1.cpp:132 for M==1;
 * result: tfar = -inf and nothing else (ray.h:1163-1185) */
RQO_API void rqo_occluded1M(void* h, void* ray, uint32_t M, size_t stride) {
  const scene_t* sc = (const scene_t*)h;
  for (uint32_t i = 0; i < M; i++) {
    ray_t* r = (ray_t*)((char*)ray + (size_t)i * stride);
    if (!(r->tnear <= r->tfar) || r->tfar < 0.0f) continue;
    if (M > 1 && !(r->tnear >= 0.0f)) continue;
    ray_t tmp = *r; rhit_t th;
    if (trace_one(sc, &tmp, &th, 1, INVALID_ID)) r->tfar = -INFINITY;
  }
}

/* ------------------------------------------------------------------------------------------
 * Single-level instancing (RTC_GEOMETRY_TYPE_INSTANCE).  Follows
 *   kernels/geometry/instance_intersector.cpp:48-105   ray -> instance space, trace the instanced scene, restore
 *   kernels/common/scene_instance.h:61-66,140-143      bounds = xfmBounds(local2world, scene bounds); world2local = rcp(local2world)
 *   common/math/affinespace.h:102-118, linearspace3.h:44-50,155-156   xfmPoint / xfmVector / inverse = adjoint / det
 * The top level is restated as a plain loop over the instances (closest hit does not depend on the visiting order);
 * hits carry the instanced scene's geomID / primID, Ng in instance space and instID[0] = geomID of the instance.
 * ------------------------------------------------------------------------------------------ */
typedef struct { void* scene; float l2w[12]; uint32_t geomID; } rqo_instance;       /* l2w column major: vx, vy, vz, p */
typedef struct { const scene_t* scene; float w2l[12]; uint32_t geomID; box3 wb; } inst_t;
typedef struct { const scene_t* base; inst_t* inst; int ninst; box3 bounds; } top_t;

static v3 xfm_point(const float* m, v3 p) {
  return V(madd(p.x, m[0], madd(p.y, m[3], madd(p.z, m[6], m[9]))), madd(p.x, m[1], madd(p.y, m[4], madd(p.z, m[7], m[10]))),
           madd(p.x, m[2], madd(p.y, m[5], madd(p.z, m[8], m[11]))));
}
static v3 xfm_vector(const float* m, v3 v) {
  return V(madd(v.x, m[0], madd(v.y, m[3], v.z * m[6])), madd(v.x, m[1], madd(v.y, m[4], v.z * m[7])), madd(v.x, m[2], madd(v.y, m[5], v.z * m[8])));
}

/* sensitivity probe (tests only): 0 = the reference's float inverse as its SSE2 object evaluates it (default), 1 = inverse evaluated in
 * double and rounded once, 2 = float with fused multiply-adds.  Answers on instanced geometry differ between them by about
 * |translation| * few ulp in instance space (u/v of small triangles move by up to 1e-4); the reference's own rcp(det) is an
 * approximation + one Newton step (common/math/math.h:58-75), which no portable code reproduces bit for bit. */
static int g_w2l_mode = 0;
RQO_API void rqo_set_w2l_mode(int m) { g_w2l_mode = m; }

RQO_API void* rqo_build_top(void* base, const rqo_instance* insts, int n) {
  top_t* t = (top_t*)calloc(1, sizeof(top_t));
  t->base = (const scene_t*)base; t->ninst = 0; t->inst = (inst_t*)calloc(n > 0 ? n : 1, sizeof(inst_t));
  t->bounds = box_empty();
  if (t->base && t->base->root) box_extend(&t->bounds, &t->base->bounds);
  for (int i = 0; i < n; i++) {
    const float* m = insts[i].l2w;
    const scene_t* sc = (const scene_t*)insts[i].scene;
    inst_t I; I.scene = sc; I.geomID = insts[i].geomID;
    const v3 vx = V(m[0], m[1], m[2]), vy = V(m[3], m[4], m[5]), vz = V(m[6], m[7], m[8]), p = V(m[9], m[10], m[11]);
    const v3 c0 = cross3(vy, vz), c1 = cross3(vz, vx), c2 = cross3(vx, vy);          /* rows of the adjoint's transpose */
    const float det = dot3(vx, c0);
    const float il[9] = {c0.x / det, c1.x / det, c2.x / det, c0.y / det, c1.y / det, c2.y / det, c0.z / det, c1.z / det, c2.z / det};
    for (int k = 0; k < 9; k++) I.w2l[k] = il[k];
    const v3 ip = xfm_vector(il, p);
    I.w2l[9] = -ip.x; I.w2l[10] = -ip.y; I.w2l[11] = -ip.z;
    if (g_w2l_mode == 0) {
      /* the reference evaluates rcp(local2world) in its lowest-ISA object (scene_instance.cpp:131 built for SSE2): products and sums
       * round separately (no FMA), dot = (x + y) + z, adjoint * rcp(det) */
      const float m0x = vy.y * vz.z - vy.z * vz.y, m0y = vy.z * vz.x - vy.x * vz.z, m0z = vy.x * vz.y - vy.y * vz.x;
      const float m1x = vz.y * vx.z - vz.z * vx.y, m1y = vz.z * vx.x - vz.x * vx.z, m1z = vz.x * vx.y - vz.y * vx.x;
      const float m2x = vx.y * vy.z - vx.z * vy.y, m2y = vx.z * vy.x - vx.x * vy.z, m2z = vx.x * vy.y - vx.y * vy.x;
      const float dt = (vx.x * m0x + vx.y * m0y) + vx.z * m0z;
      const float rd = 1.0f / dt;
      const float jl[9] = {m0x * rd, m1x * rd, m2x * rd, m0y * rd, m1y * rd, m2y * rd, m0z * rd, m1z * rd, m2z * rd};
      for (int k = 0; k < 9; k++) I.w2l[k] = jl[k];
      for (int r = 0; r < 3; r++) I.w2l[9 + r] = -((p.z * jl[6 + r] + p.y * jl[3 + r]) + p.x * jl[r]);
    }
    if (g_w2l_mode == 1) {
      const double dvx[3] = {m[0], m[1], m[2]}, dvy[3] = {m[3], m[4], m[5]}, dvz[3] = {m[6], m[7], m[8]}, dp[3] = {m[9], m[10], m[11]};
      const double d0[3] = {dvy[1] * dvz[2] - dvy[2] * dvz[1], dvy[2] * dvz[0] - dvy[0] * dvz[2], dvy[0] * dvz[1] - dvy[1] * dvz[0]};
      const double d1[3] = {dvz[1] * dvx[2] - dvz[2] * dvx[1], dvz[2] * dvx[0] - dvz[0] * dvx[2], dvz[0] * dvx[1] - dvz[1] * dvx[0]};
      const double d2[3] = {dvx[1] * dvy[2] - dvx[2] * dvy[1], dvx[2] * dvy[0] - dvx[0] * dvy[2], dvx[0] * dvy[1] - dvx[1] * dvy[0]};
      const double dd = dvx[0] * d0[0] + dvx[1] * d0[1] + dvx[2] * d0[2];
      double dl[9];
      for (int k = 0; k < 3; k++) { dl[3 * k + 0] = d0[k] / dd; dl[3 * k + 1] = d1[k] / dd; dl[3 * k + 2] = d2[k] / dd; }
      for (int k = 0; k < 9; k++) I.w2l[k] = (float)dl[k];
      for (int r = 0; r < 3; r++) I.w2l[9 + r] = (float)(-(dl[r] * dp[0] + dl[3 + r] * dp[1] + dl[6 + r] * dp[2]));
    }
    I.wb = box_empty();
    if (sc && sc->root) {
      for (int c = 0; c < 8; c++) {
        const v3 q = xfm_point(m, V((c & 4) ? sc->bounds.hi.x : sc->bounds.lo.x, (c & 2) ? sc->bounds.hi.y : sc->bounds.lo.y, (c & 1) ? sc->bounds.hi.z : sc->bounds.lo.z));
        box3 b; b.lo = q; b.hi = q; box_extend(&I.wb, &b);
      }
      if (vertex_valid(I.wb.lo) && vertex_valid(I.wb.hi)) { box_extend(&t->bounds, &I.wb); t->inst[t->ninst++] = I; }
    }
  }
  return t;
}
RQO_API void rqo_free_top(void* h) { top_t* t = (top_t*)h; if (!t) return; free(t->inst); free(t); }
RQO_API void rqo_top_bounds(void* h, float out[6]) { top_t* t = (top_t*)h; out[0] = t->bounds.lo.x; out[1] = t->bounds.lo.y; out[2] = t->bounds.lo.z; out[3] = t->bounds.hi.x; out[4] = t->bounds.hi.y; out[5] = t->bounds.hi.z; }

static int trace_top(const top_t* t, ray_t* ray, rhit_t* hit, int occluded, uint32_t ctxInst) {
  int found = 0;
  if (t->base && t->base->root) { found |= trace_one(t->base, ray, hit, occluded, ctxInst); if (occluded && found) return 1; }
  for (int i = 0; i < t->ninst; i++) {
    const inst_t* I = &t->inst[i];
    ray_t lr = *ray;                                                /* tnear, tfar carry over (instance_intersector.cpp:62-63) */
    const v3 o = xfm_point(I->w2l, V(ray->org_x, ray->org_y, ray->org_z)), d = xfm_vector(I->w2l, V(ray->dir_x, ray->dir_y, ray->dir_z));
    lr.org_x = o.x; lr.org_y = o.y; lr.org_z = o.z; lr.dir_x = d.x; lr.dir_y = d.y; lr.dir_z = d.z;
    if (trace_one(I->scene, &lr, hit, occluded, I->geomID)) {
      found = 1;
      if (occluded) return 1;
      ray->tfar = lr.tfar;
    }
  }
  return found;
}
RQO_API void rqo_top_intersect1M(void* h, void* rayhit, uint32_t M, size_t stride, uint32_t instID) {
  const top_t* t = (const top_t*)h;
  for (uint32_t i = 0; i < M; i++) {
    ray_t* r = (ray_t*)((char*)rayhit + (size_t)i * stride);
    rhit_t* ht = (rhit_t*)((char*)r + 48);
    if (!(r->tnear <= r->tfar)) continue;
    ray_t tmp = *r; rhit_t th; memset(&th, 0, sizeof(th));
    if (trace_top(t, &tmp, &th, 0, instID)) { r->tfar = tmp.tfar; *ht = th; }
  }
}
RQO_API void rqo_top_occluded1M(void* h, void* ray, uint32_t M, size_t stride) {
  const top_t* t = (const top_t*)h;
  for (uint32_t i = 0; i < M; i++) {
    ray_t* r = (ray_t*)((char*)ray + (size_t)i * stride);
    if (!(r->tnear <= r->tfar) || r->tfar < 0.0f) continue;
    if (M > 1 && !(r->tnear >= 0.0f)) continue;
    ray_t tmp = *r; rhit_t th;
    if (trace_top(t, &tmp, &th, 1, INVALID_ID)) r->tfar = -INFINITY;
  }
}
