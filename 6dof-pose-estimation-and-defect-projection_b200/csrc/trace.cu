// Subsystems (2), (4), (5): pixel -> ray generation, closest-hit traversal of the compressed
// 8-wide BVH, and accumulation of per-face / per-vertex results.
//
// Replaces, for one launch over a whole batch of frames,
//   compute_rays              /root/reference/src/defect_projection.py:196-223
//   RaycastingScene.cast_rays /root/reference/src/defect_projection.py:253-259  (Open3D/Embree)
//   the hit-point formula     /root/reference/src/defect_projection.py:261-263
//
// Arithmetic contract (identical to oracle/oracle.c semantic (A), checked bit for bit):
//   * ray directions are computed in float64 exactly as the reference does
//     (d = (xn, yn, 1)/sqrt(xn^2+yn^2+1), xn = (x-cx)/fx, no half-pixel offset), moved to the
//     object frame in float64 (d_obj = Rinv*d, o = tinv) and only then rounded to float32
//     -- the same float32 tensor the reference hands to cast_rays (:251);
//   * the ray/triangle test is the watertight test of Woop, Benthin, Wald (JCGT 2013) in
//     float32 with every operation individually rounded (no FMA contraction), double
//     precision fallback when an edge function is exactly zero;
//   * closest hit = minimum t, ties in t go to the smaller face id; t >= 0.
// Box tests are free to use FMA: they only need to be conservative, which the build-time
// padding of the leaf boxes plus the per-ray plane padding below guarantee.
#include "dp_internal.cuh"

namespace dp {

namespace {

constexpr int TR_THREADS = 128;

struct RayCtx {
    // watertight test constants
    float ox, oy, oz;     // origin permuted: (kx, ky, kz)
    float Sx, Sy, Sz;
    int kx, ky, kz;
};

__device__ __forceinline__ float sel3(float x, float y, float z, int k)
{
    return k == 0 ? x : (k == 1 ? y : z);
}

// returns true and t when the ray hits with t >= 0.  Mirrors wt_test() of oracle/oracle.c.
__device__ __forceinline__ bool tri_test(const RayCtx &r, const float4 &p0, const float4 &p1, const float4 &p2,
                                         float &tout)
{
    const float Akx = __fsub_rn(sel3(p0.x, p0.y, p0.z, r.kx), r.ox);
    const float Aky = __fsub_rn(sel3(p0.x, p0.y, p0.z, r.ky), r.oy);
    const float Akz = __fsub_rn(sel3(p0.x, p0.y, p0.z, r.kz), r.oz);
    const float Bkx = __fsub_rn(sel3(p1.x, p1.y, p1.z, r.kx), r.ox);
    const float Bky = __fsub_rn(sel3(p1.x, p1.y, p1.z, r.ky), r.oy);
    const float Bkz = __fsub_rn(sel3(p1.x, p1.y, p1.z, r.kz), r.oz);
    const float Ckx = __fsub_rn(sel3(p2.x, p2.y, p2.z, r.kx), r.ox);
    const float Cky = __fsub_rn(sel3(p2.x, p2.y, p2.z, r.ky), r.oy);
    const float Ckz = __fsub_rn(sel3(p2.x, p2.y, p2.z, r.kz), r.oz);

    const float Ax = __fsub_rn(Akx, __fmul_rn(r.Sx, Akz));
    const float Ay = __fsub_rn(Aky, __fmul_rn(r.Sy, Akz));
    const float Bx = __fsub_rn(Bkx, __fmul_rn(r.Sx, Bkz));
    const float By = __fsub_rn(Bky, __fmul_rn(r.Sy, Bkz));
    const float Cx = __fsub_rn(Ckx, __fmul_rn(r.Sx, Ckz));
    const float Cy = __fsub_rn(Cky, __fmul_rn(r.Sy, Ckz));

    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
        V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
        W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return false;
    const float Az = __fmul_rn(r.Sz, Akz);
    const float Bz = __fmul_rn(r.Sz, Bkz);
    const float Cz = __fmul_rn(r.Sz, Ckz);
    float T = __fmul_rn(U, Az);
    T = __fadd_rn(T, __fmul_rn(V, Bz));
    T = __fadd_rn(T, __fmul_rn(W, Cz));
    const float t = __fdiv_rn(T, det);
    if (!(t >= 0.0f)) return false;
    tout = t;
    return true;
}

__device__ __forceinline__ float safe_inv(float d)
{
    return fabsf(d) > 1e-18f ? __fdiv_rn(1.0f, d) : copysignf(1e18f, d);
}

__device__ __forceinline__ float byte_f(unsigned w, int i) { return (float)((w >> (8 * i)) & 0xffu); }

// Closest hit of one ray.  stack: shared-memory column of this thread (stride TR_THREADS).
template <bool STATS>
__device__ __forceinline__ void traverse(const WideNode *__restrict__ nodes, const TriRec *__restrict__ tris,
                                         float pad_abs, float ox, float oy, float oz, float dx, float dy,
                                         float dz, uint2 *stack, float &best_t, int &best_f, unsigned &n_nodes,
                                         unsigned &n_tris)
{
    RayCtx r;
    {
        const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
        int kz = 0;
        float m = ax;
        if (ay > m) { kz = 1; m = ay; }
        if (az > m) { kz = 2; }
        int kx = kz + 1; if (kx == 3) kx = 0;
        int ky = kx + 1; if (ky == 3) ky = 0;
        const float dkz = sel3(dx, dy, dz, kz);
        if (dkz < 0.0f) { const int t = kx; kx = ky; ky = t; }
        r.kx = kx; r.ky = ky; r.kz = kz;
        r.Sx = __fdiv_rn(sel3(dx, dy, dz, kx), dkz);
        r.Sy = __fdiv_rn(sel3(dx, dy, dz, ky), dkz);
        r.Sz = __fdiv_rn(1.0f, dkz);
        r.ox = sel3(ox, oy, oz, kx);
        r.oy = sel3(ox, oy, oz, ky);
        r.oz = sel3(ox, oy, oz, kz);
    }
    const float ix = safe_inv(dx), iy = safe_inv(dy), iz = safe_inv(dz);
    // every slab plane is moved outwards by pad_abs (in space), i.e. pad_abs*|1/d| in t
    const float px = pad_abs * fabsf(ix), py = pad_abs * fabsf(iy), pz = pad_abs * fabsf(iz);
    const unsigned oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
    const unsigned octinv = 7u ^ oct;

    uint2 lstack[STACK_LOCAL];
    int sp = 0;
    float best = best_t;
    int bf = best_f;
    float tlimit = best * T_SLACK;

    uint2 ng = make_uint2(0u, 0x80000000u);
    for (;;) {
        // ---- take the nearest-ordered inner child of the current node group
        const unsigned hits = ng.y;
        const int bit = 31 - __clz(hits);
        ng.y &= ~(1u << bit);
        if (ng.y & 0xff000000u) {
            if (sp < STACK_SMEM) stack[sp * TR_THREADS] = ng; else lstack[sp - STACK_SMEM] = ng;
            ++sp;
        }
        const unsigned slot = (unsigned)(bit - 24) ^ octinv;
        const unsigned rel = __popc(hits & 0xffu & ((1u << slot) - 1u));
        const uint4 *np = reinterpret_cast<const uint4 *>(nodes + (ng.x + rel));
        const uint4 w0 = __ldg(np), w1 = __ldg(np + 1), w2 = __ldg(np + 2), w3 = __ldg(np + 3), w4 = __ldg(np + 4);
        if (STATS) ++n_nodes;

        const float adjx = __uint_as_float((w0.w & 0xffu) << 23) * ix;
        const float adjy = __uint_as_float(((w0.w >> 8) & 0xffu) << 23) * iy;
        const float adjz = __uint_as_float(((w0.w >> 16) & 0xffu) << 23) * iz;
        const float orgx = (__uint_as_float(w0.x) - ox) * ix;
        const float orgy = (__uint_as_float(w0.y) - oy) * iy;
        const float orgz = (__uint_as_float(w0.z) - oz) * iz;
        const float nox = orgx - px, fox = orgx + px;
        const float noy = orgy - py, foy = orgy + py;
        const float noz = orgz - pz, foz = orgz + pz;
        // near/far quantised planes per axis, swapped once per node by the ray's sign
        const bool sx = ix < 0.0f, sy = iy < 0.0f, sz = iz < 0.0f;
        const unsigned nx0 = sx ? w3.z : w2.x, nx1 = sx ? w3.w : w2.y;
        const unsigned fx0 = sx ? w2.x : w3.z, fx1 = sx ? w2.y : w3.w;
        const unsigned ny0 = sy ? w4.x : w2.z, ny1 = sy ? w4.y : w2.w;
        const unsigned fy0 = sy ? w2.z : w4.x, fy1 = sy ? w2.w : w4.y;
        const unsigned nz0 = sz ? w4.z : w3.x, nz1 = sz ? w4.w : w3.y;
        const unsigned fz0 = sz ? w3.x : w4.z, fz1 = sz ? w3.y : w4.w;

        unsigned hitmask = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const unsigned meta4 = half ? w1.w : w1.z;
            const unsigned nxw = half ? nx1 : nx0, fxw = half ? fx1 : fx0;
            const unsigned nyw = half ? ny1 : ny0, fyw = half ? fy1 : fy0;
            const unsigned nzw = half ? nz1 : nz0, fzw = half ? fz1 : fz0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float tnx = fmaf(byte_f(nxw, j), adjx, nox);
                const float tny = fmaf(byte_f(nyw, j), adjy, noy);
                const float tnz = fmaf(byte_f(nzw, j), adjz, noz);
                const float tfx = fmaf(byte_f(fxw, j), adjx, fox);
                const float tfy = fmaf(byte_f(fyw, j), adjy, foy);
                const float tfz = fmaf(byte_f(fzw, j), adjz, foz);
                const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
                const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, tlimit));
                if (tmin <= tmax) {
                    const unsigned meta = (meta4 >> (8 * j)) & 0xffu;
                    const bool inner = (meta & 0x18u) == 0x18u;
                    const unsigned bi = (inner ? (meta ^ octinv) : meta) & 31u;
                    hitmask |= (meta >> 5) << bi;
                }
            }
        }
        ng = make_uint2(w1.x, (hitmask & 0xff000000u) | (w0.w >> 24));
        unsigned tmask = hitmask & 0x00ffffffu;
        const unsigned tbase = w1.y;

        // ---- triangles of this node
        while (tmask) {
            const int b = __ffs(tmask) - 1;
            tmask &= tmask - 1u;
            const float4 *tp = reinterpret_cast<const float4 *>(tris + (tbase + b));
            const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
            if (STATS) ++n_tris;
            float t;
            if (tri_test(r, p0, p1, p2, t)) {
                const int f = __float_as_int(p0.w);
                if (t < best || (t == best && f < bf)) {
                    best = t;
                    bf = f;
                    tlimit = best * T_SLACK;
                }
            }
        }

        // ---- next node group
        if (!(ng.y & 0xff000000u)) {
            if (sp == 0) break;
            --sp;
            ng = (sp < STACK_SMEM) ? stack[sp * TR_THREADS] : lstack[sp - STACK_SMEM];
        }
    }
    best_t = best;
    best_f = bf;
}

__device__ __forceinline__ void warp_stats(TraceStats *stats, bool valid, bool hit, unsigned nn, unsigned nt)
{
    unsigned long long a = nn, b = nt;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    const unsigned vm = __ballot_sync(0xffffffffu, valid), hm = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats->rays, (unsigned long long)__popc(vm));
        atomicAdd(&stats->hits, (unsigned long long)__popc(hm));
        atomicAdd(&stats->nodes, a);
        atomicAdd(&stats->tris, b);
    }
}

template <bool STATS>
__global__ void __launch_bounds__(TR_THREADS)
k_trace_pixels(const WideNode *__restrict__ nodes, const TriRec *__restrict__ tris, const float *__restrict__ d_scale,
               const uint32_t *__restrict__ pixel, const float *__restrict__ intensity,
               const long long *__restrict__ d_n, long long n_max, int H, int W, const FrameXf *__restrict__ xf,
               long long n_xf, float *__restrict__ t_hit, int32_t *__restrict__ face, float *__restrict__ point,
               double *__restrict__ point64, Accum acc, int has_acc, long long *d_hits, TraceStats *stats)
{
    __shared__ uint2 s_stack[STACK_SMEM * TR_THREADS];
    long long n = *d_n;
    if (n > n_max) n = n_max;
    const long long i = blockIdx.x * (long long)TR_THREADS + threadIdx.x;
    if ((long long)blockIdx.x * TR_THREADS >= n) return;      // whole block idle
    const bool valid = i < n;

    float best = __int_as_float(0x7f800000);
    int bf = -1;
    unsigned nn = 0, nt = 0;
    double dcx = 0.0, dcy = 0.0, dcz = 0.0;
    if (valid) {
        const uint32_t pix = pixel[i];
        const uint32_t hw = (uint32_t)H * (uint32_t)W;
        const uint32_t fr = pix / hw;
        const uint32_t rem = pix - fr * hw;
        const uint32_t y = rem / (uint32_t)W;
        const uint32_t x = rem - y * (uint32_t)W;
        const double *c = xf[(long long)fr < n_xf ? fr : 0].v;
        // compute_rays (:216-221) in float64, operation by operation
        const double xn = __ddiv_rn(__dsub_rn((double)x, c[2]), c[0]);
        const double yn = __ddiv_rn(__dsub_rn((double)y, c[3]), c[1]);
        const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(xn, xn), __dmul_rn(yn, yn)), 1.0));
        dcx = __ddiv_rn(xn, nrm);
        dcy = __ddiv_rn(yn, nrm);
        dcz = __ddiv_rn(1.0, nrm);
        const double *Ri = c + 4;
        const float dx = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[0], dcx), __dmul_rn(Ri[1], dcy)), __dmul_rn(Ri[2], dcz));
        const float dy = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[3], dcx), __dmul_rn(Ri[4], dcy)), __dmul_rn(Ri[5], dcz));
        const float dz = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[6], dcx), __dmul_rn(Ri[7], dcy)), __dmul_rn(Ri[8], dcz));
        const float ox = (float)c[13], oy = (float)c[14], oz = (float)c[15];
        const float pad_abs = 1.9073486e-6f * (__ldg(d_scale) + fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))));
        traverse<STATS>(nodes, tris, pad_abs, ox, oy, oz, dx, dy, dz, s_stack + threadIdx.x, best, bf, nn, nt);
        if (t_hit) t_hit[i] = best;
        if (face) face[i] = bf;
        if (point64) {
            const double t = (double)best;
            const double qn = __longlong_as_double(0x7ff8000000000000ll);
            point64[3 * i + 0] = bf >= 0 ? __dmul_rn(dcx, t) : qn;
            point64[3 * i + 1] = bf >= 0 ? __dmul_rn(dcy, t) : qn;
            point64[3 * i + 2] = bf >= 0 ? __dmul_rn(dcz, t) : qn;
        }
        if (point) {
            if (bf >= 0) {
                // :261-263 with origin 0: p = d (float64) * t (float32), stored as float32
                const double t = (double)best;
                point[3 * i + 0] = (float)__dmul_rn(dcx, t);
                point[3 * i + 1] = (float)__dmul_rn(dcy, t);
                point[3 * i + 2] = (float)__dmul_rn(dcz, t);
            } else {
                const float qnan = __int_as_float(0x7fc00000);
                point[3 * i + 0] = qnan; point[3 * i + 1] = qnan; point[3 * i + 2] = qnan;
            }
        }
    }
    const bool hit = valid && bf >= 0;
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    const int lane = threadIdx.x & 31;
    if (d_hits && lane == 0 && hm) atomicAdd(reinterpret_cast<unsigned long long *>(d_hits), (unsigned long long)__popc(hm));
    if (has_acc && hit) {
        // warp-aggregated: one atomic per distinct face in the warp
        const unsigned peers = __match_any_sync(hm, bf);
        float I = intensity ? intensity[i] : 0.0f;
        I = I > 0.0f ? I : 0.0f;
        const unsigned mx = __reduce_max_sync(peers, __float_as_uint(I));
        if (lane == __ffs(peers) - 1) {
            atomicAdd(&acc.hist[bf], __popc(peers));
            atomicMax(&acc.fmax[bf], mx);
            const int32_t *f3 = acc.F + 3ll * bf;
            atomicMax(&acc.vmax[f3[0]], mx);
            atomicMax(&acc.vmax[f3[1]], mx);
            atomicMax(&acc.vmax[f3[2]], mx);
        }
    }
    if (STATS) warp_stats(stats, valid, hit, nn, nt);
}

template <bool STATS>
__global__ void __launch_bounds__(TR_THREADS)
k_trace_rays6(const WideNode *__restrict__ nodes, const TriRec *__restrict__ tris, const float *__restrict__ d_scale,
              const float *__restrict__ rays6, long long n, float *__restrict__ t_hit, int32_t *__restrict__ face,
              TraceStats *stats)
{
    __shared__ uint2 s_stack[STACK_SMEM * TR_THREADS];
    const long long i = blockIdx.x * (long long)TR_THREADS + threadIdx.x;
    const bool valid = i < n;
    float best = __int_as_float(0x7f800000);
    int bf = -1;
    unsigned nn = 0, nt = 0;
    if (valid) {
        const float *r = rays6 + 6 * i;
        const float ox = r[0], oy = r[1], oz = r[2], dx = r[3], dy = r[4], dz = r[5];
        const float pad_abs = 1.9073486e-6f * (__ldg(d_scale) + fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))));
        traverse<STATS>(nodes, tris, pad_abs, ox, oy, oz, dx, dy, dz, s_stack + threadIdx.x, best, bf, nn, nt);
        if (t_hit) t_hit[i] = best;
        if (face) face[i] = bf;
    }
    if (STATS) warp_stats(stats, valid, valid && bf >= 0, nn, nt);
}

// compute_rays (:196-223) on its own: float64 unit directions in the camera frame
__global__ void k_compute_rays(const int32_t *__restrict__ xs, const int32_t *__restrict__ ys, long long n, FrameXf xf,
                               double *__restrict__ rays3)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *c = xf.v;
    const double xn = __ddiv_rn(__dsub_rn((double)xs[i], c[2]), c[0]);
    const double yn = __ddiv_rn(__dsub_rn((double)ys[i], c[3]), c[1]);
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(xn, xn), __dmul_rn(yn, yn)), 1.0));
    rays3[3 * i + 0] = __ddiv_rn(xn, nrm);
    rays3[3 * i + 1] = __ddiv_rn(yn, nrm);
    rays3[3 * i + 2] = __ddiv_rn(1.0, nrm);
}

}  // namespace

cudaError_t launch_compute_rays(const int32_t *xs, const int32_t *ys, int64_t n, const FrameXf &xf, double *rays3,
                                cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_compute_rays<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xs, ys, n, xf, rays3);
    return cudaGetLastError();
}

cudaError_t launch_trace_pixels(const BvhView &bvh, const uint32_t *pixel, const float *intensity,
                                const long long *d_n, int64_t n_max, int H, int W, const FrameXf *xf,
                                int64_t n_xf, float *t_hit, int32_t *face, float *point, double *point64,
                                const Accum *acc, long long *d_hits, TraceStats *stats, cudaStream_t s)
{
    if (n_max <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_max + TR_THREADS - 1) / TR_THREADS);
    Accum a = acc ? *acc : Accum{nullptr, nullptr, nullptr, nullptr};
    if (stats)
        k_trace_pixels<true><<<grid, TR_THREADS, 0, s>>>(bvh.nodes, bvh.tris, bvh.d_scale, pixel, intensity, d_n, n_max,
                                                         H, W, xf, n_xf, t_hit, face, point, point64, a,
                                                         acc != nullptr, d_hits, stats);
    else
        k_trace_pixels<false><<<grid, TR_THREADS, 0, s>>>(bvh.nodes, bvh.tris, bvh.d_scale, pixel, intensity, d_n,
                                                          n_max, H, W, xf, n_xf, t_hit, face, point, point64, a,
                                                          acc != nullptr, d_hits, stats);
    return cudaGetLastError();
}

cudaError_t launch_trace_rays6(const BvhView &bvh, const float *rays6, int64_t n, float *t_hit, int32_t *face,
                               TraceStats *stats, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + TR_THREADS - 1) / TR_THREADS);
    if (stats)
        k_trace_rays6<true><<<grid, TR_THREADS, 0, s>>>(bvh.nodes, bvh.tris, bvh.d_scale, rays6, n, t_hit, face, stats);
    else
        k_trace_rays6<false><<<grid, TR_THREADS, 0, s>>>(bvh.nodes, bvh.tris, bvh.d_scale, rays6, n, t_hit, face, stats);
    return cudaGetLastError();
}

}  // namespace dp
