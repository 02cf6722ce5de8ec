/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference hot path
 *     /root/reference/src/defect_projection.py::ray_tracing (lines 527-563)
 * and its helpers.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or as the timed CPU baseline.  The shipped package never imports it.
 *
 * PARITY UNPINNED at the ray-caster boundary: the reference hands the rays to
 * open3d==0.18.0 RaycastingScene.cast_rays (Intel Embree 3.x, CPU, float32);
 * neither library is in /root/reference nor installable offline, and the
 * reference has no tests or golden vectors.  What IS pinned (tests/golden/,
 * made by tests/golden/make_golden.py from the reference's own Python):
 *   - heatmap_to_points  (src/defect_projection.py:165-179)
 *   - compute_rays       (src/defect_projection.py:196-223)
 *   - the hit-point formula and ordering of intersect_rays_with_mesh (:258-264)
 *   - create_intersection_pcd colour mapping (:268-294)
 * For the closest-hit arithmetic itself the oracle states two semantics:
 *   (A) orc_cast_*_f32 : the float32 watertight test (Woop/Benthin/Wald 2013)
 *       with every operation individually rounded (no FMA), closest hit = min t,
 *       ties in t broken by the smaller face id.  This is the exact arithmetic
 *       contract of the CUDA traversal kernel: the GPU must match it bit for bit
 *       on EVERY ray (face id and t).
 *   (B) orc_cast_f64   : Moeller-Trumbore in float64 on the same float32-rounded
 *       inputs the reference feeds Embree (vertices posed in float64 then cast
 *       to float32, :549-550/:245; rays float32 from origin 0, :247-251), plus a
 *       tie classifier.  Rays whose f64 hit lies within tau of a triangle edge,
 *       that graze, or that have a second surface within tau_t are "ties": any
 *       float32 ray caster (Embree included) may resolve them differently.
 *       On non-tie rays face ids must be bit-exact; hit points within
 *       1e-5 * bbox diagonal (north_star).
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* H1  heatmap_to_points  (src/defect_projection.py:165-179)                  */
/*     y,x = np.where(heatmap > thr); row-major order; strict '>'             */
/* ------------------------------------------------------------------------- */
ORC_API int64_t orc_heatmap_to_points_f64(const double *heat, int64_t H, int64_t W, double thr,
                                          int64_t *xs, int64_t *ys, double *I)
{
    int64_t n = 0;
    for (int64_t y = 0; y < H; ++y)
        for (int64_t x = 0; x < W; ++x) {
            double v = heat[y * W + x];
            if (v > thr) {
                if (xs) { xs[n] = x; ys[n] = y; I[n] = v; }
                ++n;
            }
        }
    return n;
}

ORC_API int64_t orc_heatmap_to_points_f32(const float *heat, int64_t H, int64_t W, float thr,
                                          int64_t *xs, int64_t *ys, float *I)
{
    int64_t n = 0;
    for (int64_t y = 0; y < H; ++y)
        for (int64_t x = 0; x < W; ++x) {
            float v = heat[y * W + x];
            if (v > thr) {
                if (xs) { xs[n] = x; ys[n] = y; I[n] = v; }
                ++n;
            }
        }
    return n;
}

/* ------------------------------------------------------------------------- */
/* H2  compute_rays  (src/defect_projection.py:196-223), float64              */
/*     d = (xn, yn, 1) / sqrt(xn*xn + yn*yn + 1), pixel centre = integer      */
/* ------------------------------------------------------------------------- */
static inline void ray_dir_f64(double x, double y, double fx, double fy, double cx, double cy,
                               double d[3])
{
    double xn = (x - cx) / fx;
    double yn = (y - cy) / fy;
    double n = sqrt((xn * xn + yn * yn) + 1.0);
    d[0] = xn / n;
    d[1] = yn / n;
    d[2] = 1.0 / n;
}

ORC_API void orc_compute_rays(const int64_t *xs, const int64_t *ys, int64_t n, double fx, double fy,
                              double cx, double cy, double *rays)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        ray_dir_f64((double)xs[i], (double)ys[i], fx, fy, cx, cy, rays + 3 * i);
}

/* Object-frame rays (rigid mode of the CUDA path): o = tinv, d = Rinv * d_cam,
 * float64 arithmetic, each op rounded, then cast to float32.  xf = 16 doubles:
 * fx fy cx cy | Rinv (row-major 3x3) | tinv (3).  rays6 = [n][6] float32. */
ORC_API void orc_rays_object_frame(const int64_t *xs, const int64_t *ys, int64_t n,
                                   const double *xf, float *rays6)
{
    const double *Ri = xf + 4, *ti = xf + 13;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double d[3];
        ray_dir_f64((double)xs[i], (double)ys[i], xf[0], xf[1], xf[2], xf[3], d);
        float *r = rays6 + 6 * i;
        r[0] = (float)ti[0]; r[1] = (float)ti[1]; r[2] = (float)ti[2];
        for (int k = 0; k < 3; ++k)
            r[3 + k] = (float)((Ri[3 * k + 0] * d[0] + Ri[3 * k + 1] * d[1]) + Ri[3 * k + 2] * d[2]);
    }
}

/* ------------------------------------------------------------------------- */
/* H3  mesh re-posing (src/defect_projection.py:549-550) then the float32 cast */
/*     of TriangleMesh.from_legacy (:245): v' = T*(v,1) in float64 -> float32  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_pose_vertices(const double *V, int64_t nV, const double *T, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nV; ++i) {
        double x = V[3 * i], y = V[3 * i + 1], z = V[3 * i + 2];
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] =
                (float)(((T[4 * k + 0] * x + T[4 * k + 1] * y) + T[4 * k + 2] * z) + T[4 * k + 3]);
    }
}

/* ------------------------------------------------------------------------- */
/* (A) float32 watertight ray/triangle test -- the arithmetic contract         */
/* ------------------------------------------------------------------------- */
typedef struct {
    float o[3];
    int kx, ky, kz;
    float Sx, Sy, Sz;
} wt_ray;

static inline void wt_setup(const float *r6, wt_ray *w)
{
    w->o[0] = r6[0]; w->o[1] = r6[1]; w->o[2] = r6[2];
    float ax = fabsf(r6[3]), ay = fabsf(r6[4]), az = fabsf(r6[5]);
    int kz = 0;
    float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    const float *d = r6 + 3;
    if (d[kz] < 0.0f) { int t = kx; kx = ky; ky = t; }
    w->kx = kx; w->ky = ky; w->kz = kz;
    w->Sx = d[kx] / d[kz];
    w->Sy = d[ky] / d[kz];
    w->Sz = 1.0f / d[kz];
}

/* returns 1 and *t if the ray hits the triangle with t >= 0 */
static inline int wt_test(const wt_ray *w, const float *v0, const float *v1, const float *v2, float *tout)
{
    const int kx = w->kx, ky = w->ky, kz = w->kz;
    float A[3], B[3], C[3];
    for (int k = 0; k < 3; ++k) {
        A[k] = v0[k] - w->o[k];
        B[k] = v1[k] - w->o[k];
        C[k] = v2[k] - w->o[k];
    }
    float p;
    p = w->Sx * A[kz]; const float Ax = A[kx] - p;
    p = w->Sy * A[kz]; const float Ay = A[ky] - p;
    p = w->Sx * B[kz]; const float Bx = B[kx] - p;
    p = w->Sy * B[kz]; const float By = B[ky] - p;
    p = w->Sx * C[kz]; const float Cx = C[kx] - p;
    p = w->Sy * C[kz]; const float Cy = C[ky] - p;
    float q1, q2;
    q1 = Cx * By; q2 = Cy * Bx; float U = q1 - q2;
    q1 = Ax * Cy; q2 = Ay * Cx; float V = q1 - q2;
    q1 = Bx * Ay; q2 = By * Ax; float W = q1 - q2;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double d1, d2;
        d1 = (double)Cx * (double)By; d2 = (double)Cy * (double)Bx; U = (float)(d1 - d2);
        d1 = (double)Ax * (double)Cy; d2 = (double)Ay * (double)Cx; V = (float)(d1 - d2);
        d1 = (double)Bx * (double)Ay; d2 = (double)By * (double)Ax; W = (float)(d1 - d2);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
    float det = U + V;
    det = det + W;
    if (det == 0.0f) return 0;
    const float Az = w->Sz * A[kz];
    const float Bz = w->Sz * B[kz];
    const float Cz = w->Sz * C[kz];
    float T = U * Az;
    q1 = V * Bz; T = T + q1;
    q1 = W * Cz; T = T + q1;
    const float t = T / det;
    if (!(t >= 0.0f)) return 0;          /* also rejects NaN */
    *tout = t;
    return 1;
}

ORC_API int orc_tri_test_f32(const float *r6, const float *v0, const float *v1, const float *v2, float *t)
{
    wt_ray w;
    wt_setup(r6, &w);
    return wt_test(&w, v0, v1, v2, t);
}

/* brute force: every ray against every triangle.  O(n*nF). */
ORC_API void orc_cast_brute_f32(const float *V, const int32_t *F, int64_t nF, const float *rays6,
                                int64_t n, float *tout, int32_t *fout)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        wt_ray w;
        wt_setup(rays6 + 6 * i, &w);
        float best = INFINITY;
        int32_t bf = -1;
        for (int64_t f = 0; f < nF; ++f) {
            float t;
            if (wt_test(&w, V + 3 * (int64_t)F[3 * f], V + 3 * (int64_t)F[3 * f + 1],
                        V + 3 * (int64_t)F[3 * f + 2], &t)) {
                if (t < best) { best = t; bf = (int32_t)f; }  /* ascending f => smaller id wins ties */
            }
        }
        tout[i] = best;
        fout[i] = bf;
    }
}

/* ------------------------------------------------------------------------- */
/* CPU BVH (binned SAH, binary) used by the oracle casters and the CPU baseline */
/* ------------------------------------------------------------------------- */
typedef struct {
    float lo[3], hi[3];
    int32_t left;      /* internal: index of left child (right = left+1); leaf: first prim */
    int32_t count;     /* 0 => internal, >0 => leaf with 'count' prims */
} bnode;

typedef struct {
    int64_t nF, nV;
    const float *V;        /* borrowed or owned copy */
    const int32_t *F;
    float *Vown;
    int32_t *Fown;
    int32_t *prim;         /* leaf order -> face id */
    bnode *nodes;
    int64_t nnodes;
    float *clo, *chi;      /* per-face boxes */
    float pad;             /* absolute box padding */
    double scale;          /* max |coordinate| */
} orc_bvh;

#define NBINS 16
#define LEAF_MAX 4

static void face_box(const orc_bvh *b, int64_t f, float lo[3], float hi[3])
{
    for (int k = 0; k < 3; ++k) {
        float a = b->V[3 * (int64_t)b->F[3 * f] + k];
        float c = b->V[3 * (int64_t)b->F[3 * f + 1] + k];
        float d = b->V[3 * (int64_t)b->F[3 * f + 2] + k];
        lo[k] = fminf(a, fminf(c, d));
        hi[k] = fmaxf(a, fmaxf(c, d));
    }
}

static inline float box_area(const float lo[3], const float hi[3])
{
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

typedef struct {
    orc_bvh *b;
    int64_t next;          /* node allocator (atomic) */
} build_ctx;

static void build_rec(build_ctx *c, int64_t node, int64_t first, int64_t count, int depth)
{
    orc_bvh *b = c->b;
    bnode *nd = &b->nodes[node];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = first; i < first + count; ++i) {
        int64_t f = b->prim[i];
        for (int k = 0; k < 3; ++k) {
            float l = b->clo[3 * f + k], h = b->chi[3 * f + k];
            lo[k] = fminf(lo[k], l);
            hi[k] = fmaxf(hi[k], h);
            float cc = 0.5f * (l + h);
            clo[k] = fminf(clo[k], cc);
            chi[k] = fmaxf(chi[k], cc);
        }
    }
    for (int k = 0; k < 3; ++k) { nd->lo[k] = lo[k] - b->pad; nd->hi[k] = hi[k] + b->pad; }
    if (count <= LEAF_MAX) {
        nd->left = (int32_t)first;
        nd->count = (int32_t)count;
        return;
    }
    /* binned SAH on the widest centroid axis, falling back to the others */
    int best_axis = -1, best_bin = -1;
    float best_cost = INFINITY;
    for (int axis = 0; axis < 3; ++axis) {
        float ext = chi[axis] - clo[axis];
        if (!(ext > 0.0f)) continue;
        int cnt[NBINS] = {0};
        float blo[NBINS][3], bhi[NBINS][3];
        for (int j = 0; j < NBINS; ++j)
            for (int k = 0; k < 3; ++k) { blo[j][k] = INFINITY; bhi[j][k] = -INFINITY; }
        float sc = (float)NBINS / ext;
        for (int64_t i = first; i < first + count; ++i) {
            int64_t f = b->prim[i];
            float cc = 0.5f * (b->clo[3 * f + axis] + b->chi[3 * f + axis]);
            int bin = (int)((cc - clo[axis]) * sc);
            if (bin >= NBINS) bin = NBINS - 1;
            if (bin < 0) bin = 0;
            cnt[bin]++;
            for (int k = 0; k < 3; ++k) {
                blo[bin][k] = fminf(blo[bin][k], b->clo[3 * f + k]);
                bhi[bin][k] = fmaxf(bhi[bin][k], b->chi[3 * f + k]);
            }
        }
        float rarea[NBINS];
        int rcnt[NBINS];
        float rl[3] = {INFINITY, INFINITY, INFINITY}, rh[3] = {-INFINITY, -INFINITY, -INFINITY};
        int rc = 0;
        for (int j = NBINS - 1; j > 0; --j) {
            for (int k = 0; k < 3; ++k) { rl[k] = fminf(rl[k], blo[j][k]); rh[k] = fmaxf(rh[k], bhi[j][k]); }
            rc += cnt[j];
            rarea[j] = rc ? box_area(rl, rh) : 0.0f;
            rcnt[j] = rc;
        }
        float ll[3] = {INFINITY, INFINITY, INFINITY}, lh[3] = {-INFINITY, -INFINITY, -INFINITY};
        int lc = 0;
        for (int j = 0; j < NBINS - 1; ++j) {
            for (int k = 0; k < 3; ++k) { ll[k] = fminf(ll[k], blo[j][k]); lh[k] = fmaxf(lh[k], bhi[j][k]); }
            lc += cnt[j];
            if (lc == 0 || rcnt[j + 1] == 0) continue;
            float cost = box_area(ll, lh) * (float)lc + rarea[j + 1] * (float)rcnt[j + 1];
            if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = j; }
        }
    }
    int64_t mid;
    if (best_axis < 0) {
        mid = first + count / 2;      /* all centroids coincide: split by index */
    } else {
        float ext = chi[best_axis] - clo[best_axis];
        float sc = (float)NBINS / ext;
        int64_t i = first, j = first + count - 1;
        while (i <= j) {
            int64_t f = b->prim[i];
            float cc = 0.5f * (b->clo[3 * f + best_axis] + b->chi[3 * f + best_axis]);
            int bin = (int)((cc - clo[best_axis]) * sc);
            if (bin >= NBINS) bin = NBINS - 1;
            if (bin < 0) bin = 0;
            if (bin <= best_bin) ++i;
            else { int32_t t = b->prim[i]; b->prim[i] = b->prim[j]; b->prim[j] = t; --j; }
        }
        mid = i;
        if (mid == first || mid == first + count) mid = first + count / 2;
    }
    int64_t l;
#pragma omp atomic capture
    { l = c->next; c->next += 2; }
    nd->left = (int32_t)l;
    nd->count = 0;
    int64_t lcount = mid - first;
    if (count > 20000 && depth < 12) {
#pragma omp task default(shared)
        build_rec(c, l, first, lcount, depth + 1);
#pragma omp task default(shared)
        build_rec(c, l + 1, mid, count - lcount, depth + 1);
#pragma omp taskwait
    } else {
        build_rec(c, l, first, lcount, depth + 1);
        build_rec(c, l + 1, mid, count - lcount, depth + 1);
    }
}

ORC_API void orc_bvh_free(orc_bvh *b)
{
    if (!b) return;
    free(b->Vown); free(b->Fown); free(b->prim); free(b->nodes); free(b->clo); free(b->chi);
    free(b);
}

/* copies V and F so the caller's arrays may go away */
ORC_API orc_bvh *orc_bvh_build(const float *V, int64_t nV, const int32_t *F, int64_t nF)
{
    orc_bvh *b = (orc_bvh *)calloc(1, sizeof(orc_bvh));
    b->nF = nF; b->nV = nV;
    b->Vown = (float *)malloc(sizeof(float) * 3 * (size_t)(nV > 0 ? nV : 1));
    b->Fown = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(nF > 0 ? nF : 1));
    memcpy(b->Vown, V, sizeof(float) * 3 * (size_t)nV);
    memcpy(b->Fown, F, sizeof(int32_t) * 3 * (size_t)nF);
    b->V = b->Vown; b->F = b->Fown;
    b->prim = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nF > 0 ? nF : 1));
    b->clo = (float *)malloc(sizeof(float) * 3 * (size_t)(nF > 0 ? nF : 1));
    b->chi = (float *)malloc(sizeof(float) * 3 * (size_t)(nF > 0 ? nF : 1));
    b->nodes = (bnode *)malloc(sizeof(bnode) * (size_t)(2 * nF + 2));
    double s = 0.0;
    for (int64_t i = 0; i < 3 * nV; ++i) { double a = fabs((double)V[i]); if (a > s) s = a; }
    b->scale = s;
    b->pad = (float)(s * ldexp(1.0, -18));
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < nF; ++f) {
        b->prim[f] = (int32_t)f;
        face_box(b, f, b->clo + 3 * f, b->chi + 3 * f);
    }
    if (nF == 0) {
        b->nnodes = 1;
        for (int k = 0; k < 3; ++k) { b->nodes[0].lo[k] = INFINITY; b->nodes[0].hi[k] = -INFINITY; }
        b->nodes[0].left = 0; b->nodes[0].count = 0;
        return b;
    }
    build_ctx c;
    c.b = b;
    c.next = 1;
#pragma omp parallel
    {
#pragma omp single
        build_rec(&c, 0, 0, nF, 0);
    }
    b->nnodes = c.next;
    return b;
}

ORC_API int64_t orc_bvh_num_nodes(const orc_bvh *b) { return b->nnodes; }

#define T_SLACK 1.0001f   /* culling bound = tbest * T_SLACK (same constant as the CUDA kernel) */

/* slab test in float64 against the padded float32 box; returns entry distance or -1 */
static inline int slab_f64(const bnode *nd, const double o[3], const double id[3], double tmax, double *tent)
{
    double t0 = 0.0, t1 = tmax;
    for (int k = 0; k < 3; ++k) {
        double a = ((double)nd->lo[k] - o[k]) * id[k];
        double c = ((double)nd->hi[k] - o[k]) * id[k];
        if (a != a || c != c) continue;         /* 0*inf: origin on the slab plane, axis-parallel */
        double n = a < c ? a : c, f = a < c ? c : a;
        if (n > t0) t0 = n;
        if (f < t1) t1 = f;
    }
    *tent = t0;
    return t0 <= t1;
}

static void cast_one_f32(const orc_bvh *b, const float *r6, float *tout, int32_t *fout)
{
    wt_ray w;
    wt_setup(r6, &w);
    double o[3] = {r6[0], r6[1], r6[2]}, id[3];
    for (int k = 0; k < 3; ++k) id[k] = 1.0 / (double)r6[3 + k];
    float best = INFINITY;
    int32_t bf = -1;
    int32_t stack[128];
    int sp = 0;
    if (b->nF > 0) stack[sp++] = 0;
    while (sp) {
        const bnode *nd = &b->nodes[stack[--sp]];
        double te;
        double bound = (best == INFINITY) ? INFINITY : (double)(best * T_SLACK);
        if (!slab_f64(nd, o, id, bound, &te)) continue;
        if (nd->count) {
            for (int32_t i = nd->left; i < nd->left + nd->count; ++i) {
                int64_t f = b->prim[i];
                float t;
                if (wt_test(&w, b->V + 3 * (int64_t)b->F[3 * f], b->V + 3 * (int64_t)b->F[3 * f + 1],
                            b->V + 3 * (int64_t)b->F[3 * f + 2], &t)) {
                    if (t < best || (t == best && (int32_t)f < bf)) { best = t; bf = (int32_t)f; }
                }
            }
        } else {
            /* near child first */
            const bnode *l = &b->nodes[nd->left], *r = l + 1;
            double tl, tr;
            int hl = slab_f64(l, o, id, bound, &tl), hr = slab_f64(r, o, id, bound, &tr);
            if (hl && hr) {
                if (tl <= tr) { stack[sp++] = nd->left + 1; stack[sp++] = nd->left; }
                else { stack[sp++] = nd->left; stack[sp++] = nd->left + 1; }
            } else if (hl) stack[sp++] = nd->left;
            else if (hr) stack[sp++] = nd->left + 1;
        }
    }
    *tout = best;
    *fout = bf;
}

ORC_API void orc_cast_bvh_f32(const orc_bvh *b, const float *rays6, int64_t n, float *tout, int32_t *fout)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) cast_one_f32(b, rays6 + 6 * i, tout + i, fout + i);
}

/* ------------------------------------------------------------------------- */
/* (B) float64 truth + tie classification                                     */
/* ------------------------------------------------------------------------- */
typedef struct {
    int hit;          /* exact hit: u,v,w >= 0, t >= 0 */
    int fat;          /* hit of the triangle grown by tau (in the plane normal to the ray) */
    double t;
    double margin;    /* signed distance (mm) from the hit to the nearest edge, + inside */
    double cosang;    /* |n . d| */
} mt_res;

static inline void cross3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static inline void mt_f64(const double o[3], const double d[3], const float *p0, const float *p1,
                          const float *p2, double tau, mt_res *r)
{
    double v0[3] = {p0[0], p0[1], p0[2]}, e1[3], e2[3], e3[3];
    for (int k = 0; k < 3; ++k) { e1[k] = (double)p1[k] - v0[k]; e2[k] = (double)p2[k] - v0[k]; e3[k] = (double)p2[k] - (double)p1[k]; }
    double P[3], Q[3], T[3], N[3];
    cross3(d, e2, P);
    double det = dot3(e1, P);
    r->hit = r->fat = 0;
    r->t = INFINITY;
    r->margin = -INFINITY;
    cross3(e1, e2, N);
    double nn = sqrt(dot3(N, N)), dn = sqrt(dot3(d, d));
    r->cosang = (nn > 0.0 && dn > 0.0) ? fabs(det) / (nn * dn) : 0.0;
    if (det == 0.0) return;
    double inv = 1.0 / det;
    for (int k = 0; k < 3; ++k) T[k] = o[k] - v0[k];
    double u = dot3(T, P) * inv;
    cross3(T, e1, Q);
    double v = dot3(d, Q) * inv;
    double t = dot3(e2, Q) * inv;
    double w = 1.0 - u - v;
    /* barycentric b_i -> distance to the opposite edge measured in the plane normal to the ray:
       b_i * (|det|/|d|) / |edge_i|   (|det|/|d| = twice the projected area) */
    double pa = fabs(det) / dn;
    double l1 = sqrt(dot3(e1, e1)), l2 = sqrt(dot3(e2, e2)), l3 = sqrt(dot3(e3, e3));
    double mu = (l2 > 0.0) ? u * pa / l2 : -INFINITY;   /* u = 0 on edge v0-v2 */
    double mv = (l1 > 0.0) ? v * pa / l1 : -INFINITY;   /* v = 0 on edge v0-v1 */
    double mw = (l3 > 0.0) ? w * pa / l3 : -INFINITY;   /* w = 0 on edge v1-v2 */
    double m = fmin(mu, fmin(mv, mw));
    r->margin = m;
    r->t = t;
    r->hit = (u >= 0.0 && v >= 0.0 && w >= 0.0 && t >= 0.0);
    r->fat = (m >= -tau && t >= -tau);
}

/* evaluate one (ray, face) pair in float64: used to check a GPU-claimed face on tie rays */
ORC_API void orc_eval_face_f64(const float *V, const int32_t *F, const float *rays6, const int32_t *face,
                               int64_t n, double tau, double *t, double *margin, double *cosang)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        t[i] = INFINITY; margin[i] = -INFINITY; cosang[i] = 0.0;
        if (face[i] < 0) continue;
        const float *r6 = rays6 + 6 * i;
        double o[3] = {r6[0], r6[1], r6[2]}, d[3] = {r6[3], r6[4], r6[5]};
        int64_t f = face[i];
        mt_res r;
        mt_f64(o, d, V + 3 * (int64_t)F[3 * f], V + 3 * (int64_t)F[3 * f + 1], V + 3 * (int64_t)F[3 * f + 2], tau, &r);
        t[i] = r.t; margin[i] = r.margin; cosang[i] = r.cosang;
    }
}

static inline int slab_pad_f64(const bnode *nd, const double o[3], const double id[3], double pad)
{
    double t0 = -pad, t1 = INFINITY;
    for (int k = 0; k < 3; ++k) {
        double a = ((double)nd->lo[k] - pad - o[k]) * id[k];
        double c = ((double)nd->hi[k] + pad - o[k]) * id[k];
        if (a != a || c != c) continue;
        double n = a < c ? a : c, f = a < c ? c : a;
        if (n > t0) t0 = n;
        if (f < t1) t1 = f;
    }
    return t0 <= t1;
}

/*
 * tie flag bits: 1 = nearest hit within tau of an edge, 2 = another face (fat) at t <= t1 + tau_t,
 *                4 = no exact hit but a fat one exists, 8 = grazing (|cos| < graze)
 * brute != 0 => test every face (validation of the BVH path on small cases).
 */
ORC_API void orc_cast_f64(const orc_bvh *b, const float *rays6, int64_t n, double tau, double tau_t,
                          double graze, int brute, double *tout, int32_t *fout, uint8_t *tie)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        const float *r6 = rays6 + 6 * i;
        double o[3] = {r6[0], r6[1], r6[2]}, d[3] = {r6[3], r6[4], r6[5]}, id[3];
        for (int k = 0; k < 3; ++k) id[k] = 1.0 / d[k];
        double t1 = INFINITY, m1 = 0.0, c1 = 1.0;
        int32_t f1 = -1;
        /* two smallest-t fat hits with distinct faces */
        double ma = INFINITY, mb = INFINITY;
        int32_t fa = -1;
#define VISIT(f_)                                                                                     \
        do {                                                                                          \
            const int64_t f = (int64_t)(f_);                                                          \
            mt_res r;                                                                                 \
            mt_f64(o, d, b->V + 3 * (int64_t)b->F[3 * f], b->V + 3 * (int64_t)b->F[3 * f + 1],          \
                   b->V + 3 * (int64_t)b->F[3 * f + 2], tau, &r);                                      \
            if (r.hit && (r.t < t1 || (r.t == t1 && (int32_t)f < f1))) {                              \
                t1 = r.t; f1 = (int32_t)f; m1 = r.margin; c1 = r.cosang;                              \
            }                                                                                         \
            if (r.fat) {                                                                              \
                if (r.t < ma) { mb = ma; ma = r.t; fa = (int32_t)f; }                                 \
                else if (r.t < mb) { mb = r.t; }                                                      \
            }                                                                                         \
        } while (0)
        if (brute) {
            for (int64_t ff = 0; ff < b->nF; ++ff) VISIT(ff);
        } else {
            int32_t stack[256];
            int sp = 0;
            if (b->nF > 0) stack[sp++] = 0;
            while (sp) {
                const bnode *nd = &b->nodes[stack[--sp]];
                if (!slab_pad_f64(nd, o, id, 4.0 * tau)) continue;
                if (nd->count) {
                    for (int32_t j = nd->left; j < nd->left + nd->count; ++j) VISIT(b->prim[j]);
                } else {
                    stack[sp++] = nd->left;
                    stack[sp++] = nd->left + 1;
                }
            }
        }
#undef VISIT
        uint8_t tf = 0;
        if (f1 >= 0) {
            if (m1 < tau) tf |= 1;
            double other = (fa != f1) ? ma : mb;
            if (other <= t1 + tau_t) tf |= 2;
            if (c1 < graze) tf |= 8;
        } else if (ma < INFINITY) {
            tf |= 4;
        }
        tout[i] = t1;
        fout[i] = f1;
        tie[i] = tf;
    }
}

/* ------------------------------------------------------------------------- */
/* H6 / H7  per-face hit histogram, per-face and per-vertex max intensity      */
/*          (extensions named by north_star; semantics fixed in SURVEY.md 8a)  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_accumulate(const int32_t *face, const float *I, int64_t n, const int32_t *F,
                            int32_t *hist, float *fmax, float *vmax)
{
    for (int64_t i = 0; i < n; ++i) {
        int32_t f = face[i];
        if (f < 0) continue;
        float v = I[i] > 0.0f ? I[i] : 0.0f;
        if (hist) hist[f] += 1;
        if (fmax && v > fmax[f]) fmax[f] = v;
        if (vmax)
            for (int k = 0; k < 3; ++k) {
                int32_t vi = F[3 * (int64_t)f + k];
                if (v > vmax[vi]) vmax[vi] = v;
            }
    }
}

/* ------------------------------------------------------------------------- */
/* One whole frame on the CPU, all host threads: the timed CPU baseline.       */
/* threshold -> rays (camera frame, origin 0) -> closest hit -> accumulate.    */
/* The mesh must already be posed into the camera frame (as the reference      */
/* does, :549-550).  Returns the number of rays.                               */
/* ------------------------------------------------------------------------- */
ORC_API int64_t orc_project_frame_f32(const orc_bvh *b, const float *heat, int64_t H, int64_t W, float thr,
                                      const double *K4 /* fx fy cx cy */, float *tout, int32_t *fout,
                                      int32_t *hist, float *fmax, float *vmax, int64_t *nhits)
{
    int64_t HW = H * W;
    int64_t *xs = (int64_t *)malloc(sizeof(int64_t) * (size_t)HW);
    int64_t *ys = (int64_t *)malloc(sizeof(int64_t) * (size_t)HW);
    float *I = (float *)malloc(sizeof(float) * (size_t)HW);
    int64_t n = orc_heatmap_to_points_f32(heat, H, W, thr, xs, ys, I);
    int64_t hits = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : hits)
    for (int64_t i = 0; i < n; ++i) {
        double d[3];
        ray_dir_f64((double)xs[i], (double)ys[i], K4[0], K4[1], K4[2], K4[3], d);
        float r6[6] = {0.0f, 0.0f, 0.0f, (float)d[0], (float)d[1], (float)d[2]};
        cast_one_f32(b, r6, tout + i, fout + i);
        hits += (fout[i] >= 0);
    }
    if (hist || fmax || vmax) orc_accumulate(fout, I, n, b->F, hist, fmax, vmax);
    if (nhits) *nhits = hits;
    free(xs); free(ys); free(I);
    return n;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thread count of the following parallel regions.  A launcher may export OMP_NUM_THREADS=1 to every rank
 * (torch.distributed.run does); the CPU arm of bench.py sets the count it measured from its affinity mask instead. */
ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* DataReader.get_heatmap  (datareader.py:639-675)  -- SURVEY.md 8f #3         */
/*   heatmap = data - np.min(data); heatmap = heatmap / np.max(heatmap)  :658   */
/*   vis = cv2.resize(heatmap, (o, o), INTER_LINEAR), o = min(H, W)       :664  */
/*   full = zeros(H, W); full[y0:y0+o, x0:x0+o] = vis                     :669  */
/* cv2 (opencv-python 4.13 wheel, IPP) is a third-party dependency; its         */
/* CV_32F / CV_64F linear resize was pinned against the real cv2 in the build   */
/* container (tests/golden/make_golden.py section E):                           */
/*   v = fma(j + .5, src/dst, -.5) [double]; s = floor(v); f = v - s;           */
/*   s < 0 -> s = 0, f = 0;  s >= src-1 -> s = src-1, f = 0;  f cast to T       */
/*   row = fma(n[s+1] - n[s], fx, n[s]);  out = fma(row1 - row0, fy, row0)      */
/* ------------------------------------------------------------------------- */
static void lin_coef(int64_t j, int64_t src, int64_t dst, int64_t *s0, int64_t *s1, double *f)
{
    double scale = (double)src / (double)dst;
    double v = fma((double)j + 0.5, scale, -0.5);
    double fl = floor(v);
    double fr = v - fl;
    int64_t s = (int64_t)fl;
    if (s < 0) { s = 0; fr = 0.0; }
    if (s >= src - 1) { s = src - 1; fr = 0.0; }
    *s0 = s;
    *s1 = s + 1 < src ? s + 1 : src - 1;
    *f = fr;
}

#define ORC_PREPARE(NAME, T, FMA)                                                                          \
    ORC_API void NAME(const T *data, int64_t sh, int64_t sw, int64_t H, int64_t W, double *out)            \
    {                                                                                                      \
        T mn = data[0], mx = data[0];                                                                      \
        int nan = 0;                                                                                       \
        for (int64_t i = 0; i < sh * sw; ++i) {                                                            \
            T v = data[i];                                                                                 \
            if (v != v) nan = 1;                                                                           \
            if (v < mn) mn = v;                                                                            \
            if (v > mx) mx = v;                                                                            \
        }                                                                                                  \
        if (nan) { mn = (T)NAN; mx = (T)NAN; }                                                             \
        T range = (T)(mx - mn);                                                                            \
        int64_t o = H < W ? H : W, y0 = (H - o) / 2, x0 = (W - o) / 2;                                      \
        for (int64_t e = 0; e < H * W; ++e) out[e] = 0.0;                                                  \
        T *row = (T *)malloc(sizeof(T) * (size_t)(sh * (o > 0 ? o : 1)));                                  \
        for (int64_t y = 0; y < sh; ++y)                                                                   \
            for (int64_t j = 0; j < o; ++j) {                                                              \
                int64_t s0, s1; double fd;                                                                 \
                lin_coef(j, sw, o, &s0, &s1, &fd);                                                         \
                T f = (T)fd;                                                                               \
                T a = (T)((T)(data[y * sw + s0] - mn) / range), b = (T)((T)(data[y * sw + s1] - mn) / range); \
                row[y * o + j] = FMA((T)(b - a), f, a);                                                    \
            }                                                                                              \
        for (int64_t i = 0; i < o; ++i) {                                                                  \
            int64_t s0, s1; double fd;                                                                     \
            lin_coef(i, sh, o, &s0, &s1, &fd);                                                             \
            T f = (T)fd;                                                                                   \
            for (int64_t j = 0; j < o; ++j) {                                                              \
                T r0 = row[s0 * o + j], r1 = row[s1 * o + j];                                              \
                out[(y0 + i) * W + x0 + j] = (double)FMA((T)(r1 - r0), f, r0);                             \
            }                                                                                              \
        }                                                                                                  \
        free(row);                                                                                         \
    }

ORC_PREPARE(orc_prepare_heatmap_f64, double, fma)
ORC_PREPARE(orc_prepare_heatmap_f32, float, fmaf)
