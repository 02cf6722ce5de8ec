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

#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------
// Per-lane traversal state.  A warp traces packets of 32 rays in lock step (coherent rays share
// node fetches); packets are handed out by a global counter to persistent warps so that no SM
// idles behind a long-running CTA.
// ------------------------------------------------------------------------------------------
struct RayState {
    RayCtx w;                 // watertight constants
    float ox, oy, oz;         // origin
    float ix, iy, iz;         // clamped reciprocal direction
    float px, py, pz;         // slab padding in t
    float anx, any, anz;      // uncompressed nodes: near plane t = plane * i + an  (an = -o*i - pad)
    float afx, afy, afz;      //                     far plane  t = plane * i + af  (af = -o*i + pad)
    unsigned order;           // three 8-bit slot masks (bytes 0..2 = bit 2, 1, 0 of the slot index): the slots a ray
                              // prefers on that bit, nearest child = the hit slot s that maximises s ^ (7 ^ octant)
    uint2 ng;                 // current node group: (child base, hit bits | imask)
    int sp;
    unsigned next;            // index of the node the next step fetches (valid while the ray is alive)
};

__device__ __forceinline__ void ray_setup(RayState &r, float ox, float oy, float oz, float dx, float dy, float dz,
                                          float scale)
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
    r.w.kx = kx; r.w.ky = ky; r.w.kz = kz;
    r.w.Sx = __fdiv_rn(sel3(dx, dy, dz, kx), dkz);
    r.w.Sy = __fdiv_rn(sel3(dx, dy, dz, ky), dkz);
    r.w.Sz = __fdiv_rn(1.0f, dkz);
    r.w.ox = sel3(ox, oy, oz, kx);
    r.w.oy = sel3(ox, oy, oz, ky);
    r.w.oz = sel3(ox, oy, oz, kz);
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.ix = safe_inv(dx); r.iy = safe_inv(dy); r.iz = safe_inv(dz);
    // every slab plane is moved outwards by pad_abs (in space), i.e. pad_abs*|1/d| in t
    const float pad_abs = 1.9073486e-6f * (scale + fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))));
    r.px = pad_abs * fabsf(r.ix); r.py = pad_abs * fabsf(r.iy); r.pz = pad_abs * fabsf(r.iz);
    // the padding (2^-19 of the scene) exceeds the rounding of o*i and of the fused plane*i + a by a factor of 16
    const float cx = -(ox * r.ix), cy = -(oy * r.iy), cz = -(oz * r.iz);
    r.anx = cx - r.px; r.any = cy - r.py; r.anz = cz - r.pz;
    r.afx = cx + r.px; r.afy = cy + r.py; r.afz = cz + r.pz;
    const unsigned oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
    const unsigned octinv = 7u ^ oct;
    r.order = ((octinv & 4u) ? 0x0fu : 0xf0u) | (((octinv & 2u) ? 0x33u : 0xccu) << 8) | (((octinv & 1u) ? 0x55u : 0xaau) << 16);
    r.ng = make_uint2(0u, 0x80000000u);      // the root as the only inner child of a virtual group
    r.sp = 0;
}

// 8-bit plane index -> float.  Measured on B200 (configs[1], 500k triangles, traversal kernel):
//   mode 1  I2F.U8 (XU pipe, 48 per node step)                                   0.311 ms   <- default
//   mode 0  PRMT into the mantissa of 2^23 + FADD (ALU + FMA pipes, exact)       0.328 ms
//   mode 3  near planes by I2F, far planes by PRMT + FADD                         0.318 ms
// A fourth variant (byte into mantissa bits 8..15 of 2^15 with the bias folded into the FMA addend,
// 0.309 ms) was dropped: it lost a hit on a ray with |1/d| ~ 7e13 in tests/test_gpu_parity.py.
#ifndef DP_PLANE_MODE
#define DP_PLANE_MODE 1
#endif
__device__ __forceinline__ __attribute__((unused)) float qf_magic(unsigned w, int i)
{
    // selector as the immediate, the magic constant in a register: one PRMT, no extra move
    const unsigned magic = 0x4B000000u;
    unsigned f;
    switch (i) {
    case 0: asm("prmt.b32 %0, %1, %2, 0x7440;" : "=r"(f) : "r"(w), "r"(magic)); break;
    case 1: asm("prmt.b32 %0, %1, %2, 0x7441;" : "=r"(f) : "r"(w), "r"(magic)); break;
    case 2: asm("prmt.b32 %0, %1, %2, 0x7442;" : "=r"(f) : "r"(w), "r"(magic)); break;
    default: asm("prmt.b32 %0, %1, %2, 0x7443;" : "=r"(f) : "r"(w), "r"(magic)); break;
    }
    return __uint_as_float(f) - 8388608.0f;
}
__device__ __forceinline__ float qf_i2f(unsigned w, int i) { return (float)((w >> (8 * i)) & 0xffu); }
#if DP_PLANE_MODE == 0
#define qf qf_magic
#define qf_xu qf_magic
#elif DP_PLANE_MODE == 1
#define qf qf_i2f
#define qf_xu qf_i2f
#else
#define qf qf_magic
#define qf_xu qf_i2f
#endif

// Prefetch of the next node (and of queued triangle records) into L1 one phase ahead.  Measured on B200: +2.5 % at
// 5M triangles (BVH 290 MB, beyond L2), -1.5 % at 500k (BVH 29 MB, L2-resident), so the launcher turns it on only
// for hierarchies larger than PREFETCH_MIN_BYTES.
constexpr size_t PREFETCH_MIN_BYTES = 96u << 20;
#ifndef DP_PREFETCH_CHILDREN
#define DP_PREFETCH_CHILDREN 0
#endif
__device__ __forceinline__ void prefetch_l1(const void *p, bool on)
{
    if (on) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// Selection half of a node step: take the nearest pending inner child out of the current node group (the rest of
// the group goes on the stack), remember its node index for the next visit and start fetching it.
__device__ __forceinline__ void node_select(RayState &r, const WideNode *__restrict__ nodes, uint2 *stack, uint2 *lstack,
                                            bool pf)
{
    uint2 ng = r.ng;
    const unsigned hits = ng.y;
    // inner hits sit at bit 24 + slot; the nearest one is the slot s maximising s ^ octinv: filter the candidates
    // bit by bit, most significant first (a filter that would leave nothing is skipped)
    unsigned c = hits >> 24, t;
    t = c & r.order;          c = t ? t : c;
    t = c & (r.order >> 8);   c = t ? t : c;
    t = c & (r.order >> 16);  c = t ? t : c;
    const unsigned slot = (unsigned)__ffs((int)c) - 1u;
    ng.y &= ~(0x01000000u << slot);
    if (ng.y & 0xff000000u) {
        if (r.sp < STACK_SMEM) stack[r.sp * TR_THREADS] = ng; else lstack[r.sp - STACK_SMEM] = ng;
        ++r.sp;
    }
    const unsigned rel = __popc(hits & 0xffu & ((1u << slot) - 1u));
    r.next = ng.x + rel;
    const char *np = reinterpret_cast<const char *>(nodes + r.next);
    prefetch_l1(np, pf);
    prefetch_l1(np + 64, pf);
}

// Visit half: fetch the selected 80-byte node, test its 8 child boxes against [0, tlimit], hand the triangle hits
// back as (tbase, tmask) and make the node's inner hits the current group (or pop one from the stack).  Returns
// false when nothing is left to visit; otherwise the next node has been selected (node_select).
// OCT >= 0: the signs of the ray's direction are compile-time constants (bit 0 = x negative, ...), which removes the
// twelve near/far selects of a visit; the packet loop picks the instantiation when all its rays share the octant
// (DP_OCT_SPECIALISE, off: counted in the SASS -- 228 -> 212 instructions per visit -- but not yet measured).
#ifndef DP_OCT_SPECIALISE
#define DP_OCT_SPECIALISE 0
#endif
template <bool STATS, int OCT = -1>
__device__ __forceinline__ bool node_step(RayState &r, const WideNode *__restrict__ nodes, uint2 *stack, uint2 *lstack,
                                          float tlimit, unsigned &tbase, unsigned &tmask, unsigned &n_nodes, bool pf)
{
    const uint4 *np = reinterpret_cast<const uint4 *>(nodes + r.next);
    const uint4 w0 = __ldg(np), w1 = __ldg(np + 1), w2 = __ldg(np + 2), w3 = __ldg(np + 3), w4 = __ldg(np + 4);
    if (STATS) ++n_nodes;
    const float adjx = __uint_as_float((w0.w & 0xffu) << 23) * r.ix;
    const float adjy = __uint_as_float(((w0.w >> 8) & 0xffu) << 23) * r.iy;
    const float adjz = __uint_as_float(((w0.w >> 16) & 0xffu) << 23) * r.iz;
    const float orgx = (__uint_as_float(w0.x) - r.ox) * r.ix;
    const float orgy = (__uint_as_float(w0.y) - r.oy) * r.iy;
    const float orgz = (__uint_as_float(w0.z) - r.oz) * r.iz;
    const float nox = orgx - r.px, fox = orgx + r.px;
    const float noy = orgy - r.py, foy = orgy + r.py;
    const float noz = orgz - r.pz, foz = orgz + r.pz;
    // near/far quantised planes per axis, swapped once per node by the ray's sign
    const bool sx = OCT < 0 ? r.ix < 0.0f : (OCT & 1) != 0, sy = OCT < 0 ? r.iy < 0.0f : (OCT & 2) != 0,
               sz = OCT < 0 ? r.iz < 0.0f : (OCT & 4) != 0;
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
        // four children at a time: first bit of each child in the hit word and its unary triangle count.  An inner
        // child is encoded as "one triangle at bit 24 + slot", so both kinds go through the same shift
        const unsigned off4 = meta4 & 0x1f1f1f1fu;
        const unsigned cbits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float tnx = fmaf(qf_xu(nxw, j), adjx, nox);
            const float tny = fmaf(qf_xu(nyw, j), adjy, noy);
            const float tnz = fmaf(qf_xu(nzw, j), adjz, noz);
            const float tfx = fmaf(qf(fxw, j), adjx, fox);
            const float tfy = fmaf(qf(fyw, j), adjy, foy);
            const float tfz = fmaf(qf(fzw, j), adjz, foz);
            const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
            const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, tlimit));
            if (tmin <= tmax) hitmask |= ((cbits4 >> (8 * j)) & 0xffu) << ((off4 >> (8 * j)) & 0xffu);
        }
    }
    uint2 ng = make_uint2(w1.x, (hitmask & 0xff000000u) | (w0.w >> 24));
#if DP_PREFETCH_CHILDREN
    if (pf && (hitmask & 0xff000000u)) {
        // the inner children of this node are contiguous: start pulling their lines towards L2 now, long before the
        // farther ones are popped from the stack
        const char *cb = reinterpret_cast<const char *>(nodes + w1.x);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cb));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cb + 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cb + 256));
    }
#endif
    tmask = hitmask & 0x00ffffffu;
    tbase = w1.y;
    if (!(ng.y & 0xff000000u)) {
        if (r.sp == 0) return false;
        --r.sp;
        ng = (r.sp < STACK_SMEM) ? stack[r.sp * TR_THREADS] : lstack[r.sp - STACK_SMEM];
    }
    r.ng = ng;
    node_select(r, nodes, stack, lstack, pf);
    return true;
}

// Selection half for the uncompressed set: as node_select, without the prefetch (the set is L2-resident by construction).
__device__ __forceinline__ void node_select_fat(RayState &r, uint2 *stack, uint2 *lstack)
{
    uint2 ng = r.ng;
    const unsigned hits = ng.y;
    unsigned c = hits >> 24, t;
    t = c & r.order;          c = t ? t : c;
    t = c & (r.order >> 8);   c = t ? t : c;
    t = c & (r.order >> 16);  c = t ? t : c;
    const unsigned slot = (unsigned)__ffs((int)c) - 1u;
    ng.y &= ~(0x01000000u << slot);
    if (ng.y & 0xff000000u) {
        if (r.sp < STACK_SMEM) stack[r.sp * TR_THREADS] = ng; else lstack[r.sp - STACK_SMEM] = ng;
        ++r.sp;
    }
    r.next = ng.x + __popc(hits & 0xffu & ((1u << slot) - 1u));
}

// Visit half for the uncompressed 208-byte node (dp_internal.cuh): thirteen 16-byte loads, per child six FMAs on the
// float planes (near/far chosen by the sign of the ray: compile-time plane blocks when OCT >= 0, per-lane block
// indices otherwise), one compare and ONE predicated OR of the slot's constant hit word (three triangle bits + the
// inner bit); a single AND with the node's valid word then leaves the triangles and inner children that exist.
// No byte->float conversions, no per-node scale, no per-child shifts.
template <bool STATS, int OCT>
__device__ __forceinline__ bool node_step_fat(RayState &r, const uint4 *__restrict__ fat, uint2 *stack, uint2 *lstack,
                                              float tlimit, unsigned &tbase, unsigned &tmask, unsigned &tvalid,
                                              unsigned &n_nodes)
{
    const uint4 *np = fat + (size_t)r.next * FAT_QUADS;
    const uint4 h = __ldg(np);
    if (STATS) ++n_nodes;
    // first 16-byte word of the near block per axis (lo: 1, 3, 5; hi: 7, 9, 11); the far block is the other one
    int bnx, bny, bnz;
    if (OCT >= 0) { bnx = (OCT & 1) ? 7 : 1; bny = (OCT & 2) ? 9 : 3; bnz = (OCT & 4) ? 11 : 5; }
    else { bnx = r.ix < 0.0f ? 7 : 1; bny = r.iy < 0.0f ? 9 : 3; bnz = r.iz < 0.0f ? 11 : 5; }
    const int bfx = 8 - bnx, bfy = 12 - bny, bfz = 16 - bnz;
    unsigned m = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float4 nx = __ldg(reinterpret_cast<const float4 *>(np + bnx + half));
        const float4 ny = __ldg(reinterpret_cast<const float4 *>(np + bny + half));
        const float4 nz = __ldg(reinterpret_cast<const float4 *>(np + bnz + half));
        const float4 fx = __ldg(reinterpret_cast<const float4 *>(np + bfx + half));
        const float4 fy = __ldg(reinterpret_cast<const float4 *>(np + bfy + half));
        const float4 fz = __ldg(reinterpret_cast<const float4 *>(np + bfz + half));
        const float pnx[4] = {nx.x, nx.y, nx.z, nx.w}, pny[4] = {ny.x, ny.y, ny.z, ny.w}, pnz[4] = {nz.x, nz.y, nz.z, nz.w};
        const float pfx[4] = {fx.x, fx.y, fx.z, fx.w}, pfy[4] = {fy.x, fy.y, fy.z, fy.w}, pfz[4] = {fz.x, fz.y, fz.z, fz.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int slot = 4 * half + j;
            const float tnx = fmaf(pnx[j], r.ix, r.anx), tny = fmaf(pny[j], r.iy, r.any), tnz = fmaf(pnz[j], r.iz, r.anz);
            const float tfx = fmaf(pfx[j], r.ix, r.afx), tfy = fmaf(pfy[j], r.iy, r.afy), tfz = fmaf(pfz[j], r.iz, r.afz);
            const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
            const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, tlimit));
            if (tmin <= tmax) m |= (7u << (3 * slot)) | (0x01000000u << slot);
        }
    }
    m &= h.z;
    uint2 ng = make_uint2(h.x, (m & 0xff000000u) | (h.z >> 24));
    tmask = m & 0x00ffffffu;
    tvalid = h.z;
    tbase = h.y;
    if (!(ng.y & 0xff000000u)) {
        if (r.sp == 0) return false;
        --r.sp;
        ng = (r.sp < STACK_SMEM) ? stack[r.sp * TR_THREADS] : lstack[r.sp - STACK_SMEM];
    }
    r.ng = ng;
    node_select_fat(r, stack, lstack);
    return true;
}

int g_tiled = -1;                      // walk dense frames in 8x4 tiles (DP_TILED)

// index in the work order -> output slot.  Dense frames are walked in 8x4-pixel tiles so that the 32
// rays of a warp share most of their path; results are still written at the row-major slot.
__device__ __forceinline__ long long work_to_slot(long long i, bool tiled, int W)
{
    if (!tiled) return i;
    const long long tile = i >> 5;
    const int l = (int)(i & 31);
    const int tpr = W >> 3;
    const long long trow = tile / tpr;
    const int tcol = (int)(tile - trow * tpr);
    return (trow * 4 + (l >> 3)) * W + tcol * 8 + (l & 7);
}

#ifndef DP_TQ_CAP
#define DP_TQ_CAP 256
#endif
constexpr int TQ_CAP = DP_TQ_CAP;               // triangle queue entries per warp (power of two)
#ifndef DP_TQ_FLUSH
#define DP_TQ_FLUSH 32
#endif
constexpr int TQ_FLUSH = DP_TQ_FLUSH;           // pending (ray, triangle) pairs that trigger a test round
constexpr unsigned TQ_TRI_MASK = (1u << 27) - 1u;
constexpr unsigned long long KEY_MISS = (0x7f800000ull << 32) | 0xffffffffull;   // t = +inf, face = -1

// Triangle tests of one warp, compacted: the (ray, triangle) pairs the node steps produce are queued
// in shared memory and tested 32 at a time by all lanes, whichever lane owns the ray; the ray's
// constants come by shuffle from the owner and the result goes to the owner's 64-bit key
// (t bits << 32 | face id) with a shared-memory atomicMin: minimum t, ties to the smaller face id.
template <bool STATS>
__device__ __forceinline__ void tri_batch(const RayState &r, const TriRec *__restrict__ tris, const unsigned *queue,
                                          unsigned long long *best, int qhead, int cnt, int lane, unsigned &n_tris)
{
    const unsigned e = queue[(qhead + lane) & (TQ_CAP - 1)];
    const bool live = lane < cnt;
    const int owner = live ? (int)(e >> 27) : lane;
    RayCtx w;
    w.ox = __shfl_sync(0xffffffffu, r.w.ox, owner);
    w.oy = __shfl_sync(0xffffffffu, r.w.oy, owner);
    w.oz = __shfl_sync(0xffffffffu, r.w.oz, owner);
    w.Sx = __shfl_sync(0xffffffffu, r.w.Sx, owner);
    w.Sy = __shfl_sync(0xffffffffu, r.w.Sy, owner);
    w.Sz = __shfl_sync(0xffffffffu, r.w.Sz, owner);
    const int kpack = __shfl_sync(0xffffffffu, r.w.kx | (r.w.ky << 2) | (r.w.kz << 4), owner);
    w.kx = kpack & 3; w.ky = (kpack >> 2) & 3; w.kz = kpack >> 4;
    if (live) {
        const float4 *tp = reinterpret_cast<const float4 *>(tris + (e & TQ_TRI_MASK));
        const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
        if (STATS) ++n_tris;
        float t;
        if (tri_test(w, p0, p1, p2, t)) {
            t += 0.0f;      // -0 -> +0 so that the bit pattern orders like the value
            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)__float_as_int(p0.w);
            atomicMin(&best[owner], key);
        }
    }
    __syncwarp();
}

// SRC = 0: rays come from (pixel -> dir4[]) of the compaction + ray generation kernels, origins per frame
// SRC = 1: explicit float32 rays [n][6]
#ifndef DP_MIN_BLOCKS
#define DP_MIN_BLOCKS 7      // 72 registers: 7 CTAs (28 warps) per SM; measured ~9 % faster than 80 registers / 6 CTAs
#endif

// ------------------------------------------------------------------------------------------
// Sparse frames (the reference's production case: a thresholded defect blob, a few thousand rays).  With fewer rays
// than the device has warps a launch of k_trace is bound by the LATENCY of its slowest ray: ~100 dependent node
// visits of ~330 instructions each, executed by a lone warp.  k_trace_narrow gives every ray EIGHT lanes, one per
// child slot of the wide node: a visit is one box test per lane (6 conversions, 6 FMAs), a ballot and the selection,
// ~80 instructions; the lanes whose child is a hit leaf test its (<= 3) triangles themselves and the group reduces
// the 64-bit (t, face) key by shuffles.  Same arithmetic, same order, same culling bound as k_trace: results are
// bit-identical.  k_trace reads the ray count on the device and branches here when the frame is sparse (the host does
// not know the count: no extra launch, no read-back).
// ------------------------------------------------------------------------------------------
constexpr int NR_THREADS = 128;                       // = TR_THREADS: 4 warps x 4 rays
constexpr int NR_STACK = 64;                          // stack entries per ray (shared memory)
#ifndef DP_NARROW_MAX_RAYS
#define DP_NARROW_MAX_RAYS 131072                     // up to this many rays the narrow path takes the frame (measured:
                                                      // 16-23 % faster than packets at 35k-110k rays, 30-35 % at 9k-16k)
#endif

// raytab (optional): xn of every column followed by yn of every row, the same two float64 quotients computed once per
// call by the compaction launch when all frames share one K (2 of the 5 float64 divisions of a ray; same bits)
__device__ __forceinline__ void pixel_ray_f64(uint32_t pix, int H, int W, const FrameXf *__restrict__ xf, long long n_xf,
                                              uint32_t &fr, double &dcx, double &dcy, double &dcz,
                                              const double *__restrict__ raytab = nullptr)
{
    const uint32_t hw = (uint32_t)H * (uint32_t)W;
    fr = pix / hw;
    const uint32_t rem = pix - fr * hw;
    const uint32_t y = rem / (uint32_t)W;
    const uint32_t x = rem - y * (uint32_t)W;
    if ((long long)fr >= n_xf) fr = 0;
    const double *c = xf[fr].v;
    const double xn = raytab ? __ldg(raytab + x) : __ddiv_rn(__dsub_rn((double)x, c[2]), c[0]);
    const double yn = raytab ? __ldg(raytab + W + y) : __ddiv_rn(__dsub_rn((double)y, c[3]), c[1]);
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(xn, xn), __dmul_rn(yn, yn)), 1.0));
    dcx = __ddiv_rn(xn, nrm);
    dcy = __ddiv_rn(yn, nrm);
    dcz = __ddiv_rn(1.0, nrm);
}

// The float32 object-frame ray of a compacted pixel (what k_raygen writes to dir4) and its float64 camera-frame unit
// direction (what k_points multiplies by t), computed in place by the traversal when rays are generated in-kernel:
// same operations in the same order, bit-identical results, no 32-byte ray record through HBM.
__device__ __forceinline__ void pixel_ray_object(uint32_t pix, int H, int W, const FrameXf *__restrict__ xf, long long n_xf,
                                                 float &ox, float &oy, float &oz, float &dx, float &dy, float &dz,
                                                 double &dcx, double &dcy, double &dcz, const double *__restrict__ raytab = nullptr)
{
    uint32_t fr;
    pixel_ray_f64(pix, H, W, xf, n_xf, fr, dcx, dcy, dcz, raytab);
    const double *Ri = xf[fr].v + 4;
    dx = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[0], dcx), __dmul_rn(Ri[1], dcy)), __dmul_rn(Ri[2], dcz));
    dy = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[3], dcx), __dmul_rn(Ri[4], dcy)), __dmul_rn(Ri[5], dcz));
    dz = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[6], dcx), __dmul_rn(Ri[7], dcy)), __dmul_rn(Ri[8], dcz));
    const double *ti = xf[fr].v + 13;
    ox = (float)ti[0]; oy = (float)ti[1]; oz = (float)ti[2];
}

// hit point p = d_cam (float64) * t (float32), NaN on a miss  (:261-263 with origin 0; as k_points)
__device__ __forceinline__ void hit_point(bool hit, float tf, double dcx, double dcy, double dcz, double &px, double &py, double &pz)
{
    if (hit) {
        const double t = (double)tf;
        px = __dmul_rn(dcx, t); py = __dmul_rn(dcy, t); pz = __dmul_rn(dcz, t);
    } else {
        px = py = pz = __longlong_as_double(0x7ff8000000000000ll);
    }
}
__device__ __forceinline__ void store_point(long long i, bool hit, float tf, double dcx, double dcy, double dcz,
                                            float *__restrict__ point, double *__restrict__ point64)
{
    double px, py, pz;
    hit_point(hit, tf, dcx, dcy, dcz, px, py, pz);
    if (point64) { point64[3 * i] = px; point64[3 * i + 1] = py; point64[3 * i + 2] = pz; }
    if (point) { point[3 * i] = (float)px; point[3 * i + 1] = (float)py; point[3 * i + 2] = (float)pz; }
}

// Ray-sharded frame over peer memory (peer.cu): the ray's results also go into the result arrays of every OTHER rank,
// mapped into this process (NVLink stores), as the traversal produces them -- the all-gather of the slices is the
// traversal's own epilogue.  Out of line: a launch without peers never executes it and keeps its registers.
__device__ __noinline__ void peer_store(const PeerOut *__restrict__ po, long long i, float tb, int bf, bool with_point, double dcx,
                                        double dcy, double dcz)
{
    double px = 0.0, py = 0.0, pz = 0.0;
    if (with_point) hit_point(bf >= 0, tb, dcx, dcy, dcz, px, py, pz);
    const int n = po->n;
    for (int p = 0; p < n; ++p) {
        if (po->t_hit[p]) po->t_hit[p][i] = tb;
        if (po->face[p]) po->face[p][i] = bf;
        float *q = po->point[p];
        if (with_point && q) { q[3 * i] = (float)px; q[3 * i + 1] = (float)py; q[3 * i + 2] = (float)pz; }
    }
}

// ------------------------------------------------------------------------------------------
// Single-frame ray sharding: with `world` > 1 a launch traces only rank `rank`'s contiguous block of
// the compacted ray list, in units of 32-ray packets -- whole tile rows (4 image rows) when the dense frame is walked in
// 8x4 tiles, so that the block is also a contiguous range of row-major output slots [s_lo, s_hi).
// dp_shard_slots (api.cu) evaluates the same partition on the host.
// ------------------------------------------------------------------------------------------
struct Shard {
    long long p_lo, p_hi;     // packets of the work order
    long long s_lo, s_hi;     // output slots
    bool tiled;
};
__host__ __device__ inline Shard shard_of(long long n, long long total_px, int H, int W, int allow_tiled, int narrow, int rank,
                                          int world)
{
    Shard sh;
    sh.tiled = allow_tiled && n == total_px && (W & 7) == 0 && (H & 3) == 0 && W > 0 && !(narrow && n <= DP_NARROW_MAX_RAYS);
    const long long np = (n + 31) >> 5;
    sh.p_lo = 0; sh.p_hi = np;
    if (world > 1) {
        const long long unit = sh.tiled ? (long long)(W >> 3) : 1;
        const long long units = (np + unit - 1) / unit;
        sh.p_lo = (units * rank / world) * unit;
        sh.p_hi = (units * (rank + 1) / world) * unit;
        if (sh.p_lo > np) sh.p_lo = np;
        if (sh.p_hi > np) sh.p_hi = np;
    }
    // a packet of the tiled order covers 8x4 pixels; W/8 packets = four full image rows
    sh.s_lo = sh.tiled ? (sh.p_lo / (W >> 3)) * 4ll * W : sh.p_lo * 32;
    sh.s_hi = sh.tiled ? (sh.p_hi / (W >> 3)) * 4ll * W : sh.p_hi * 32;
    if (sh.s_lo > n) sh.s_lo = n;
    if (sh.s_hi > n) sh.s_hi = n;
    return sh;
}

template <bool STATS>
__device__ __noinline__ void
trace_narrow(const WideNode *__restrict__ nodes, const TriRec *__restrict__ tris, const float *__restrict__ d_scale,
             const float4 *__restrict__ dir4, const float *__restrict__ intensity, long long n,
             float *__restrict__ t_hit, int32_t *__restrict__ face, Accum acc, int has_acc,
             unsigned long long *work_counter, long long *d_hits, TraceStats *stats, uint2 *s_stack, long long s_lo,
             long long s_hi, const uint32_t *__restrict__ pixel, int H, int W, const FrameXf *__restrict__ xf, long long n_xf,
             float *__restrict__ point, double *__restrict__ point64, const PeerOut *__restrict__ peer_out,
             const double *__restrict__ raytab)
{
    const int lane = threadIdx.x & 31, c = lane & 7, g = lane >> 3;
    const unsigned gmask = 0xffu << (8 * g);
    uint2 *stack = s_stack + (threadIdx.x >> 3) * NR_STACK;
    const float scale = __ldg(d_scale);
    // one work item = four consecutive rays = one warp; a shard's slots start on a packet (32-ray) boundary
    const long long item0 = s_lo >> 2, nitems = ((s_hi + 3) >> 2) - item0;
    n = s_hi;
    unsigned long long nn = 0, nt = 0, nhit = 0, nray = 0;
    for (;;) {
        unsigned long long wv = 0;
        if (lane == 0) wv = atomicAdd(work_counter, 1ull);
        const long long w = (long long)__shfl_sync(0xffffffffu, wv, 0);
        if (w >= nitems) break;
        const long long i = (item0 + w) * 4 + g;
        const bool valid = i < n;
        bool alive = valid;
        RayState r;
        double dcx = 0.0, dcy = 0.0, dcz = 0.0;                       // camera-frame direction (lane c == 0, in-kernel rays)
        {
            float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
            if (valid) {
                if (dir4) {
                    const float4 o4 = __ldg(dir4 + 2 * i), d = __ldg(dir4 + 2 * i + 1);
                    ox = o4.x; oy = o4.y; oz = o4.z; dx = d.x; dy = d.y; dz = d.z;
                } else {
                    // in-kernel ray generation: lane 0 of the ray's eight computes, the others receive
                    if (c == 0) pixel_ray_object(pixel[i], H, W, xf, n_xf, ox, oy, oz, dx, dy, dz, dcx, dcy, dcz, raytab);
                }
            }
            if (!dir4) {
                const int src = lane & 24;
                ox = __shfl_sync(0xffffffffu, ox, src); oy = __shfl_sync(0xffffffffu, oy, src); oz = __shfl_sync(0xffffffffu, oz, src);
                dx = __shfl_sync(0xffffffffu, dx, src); dy = __shfl_sync(0xffffffffu, dy, src); dz = __shfl_sync(0xffffffffu, dz, src);
            }
            ray_setup(r, ox, oy, oz, dx, dy, dz, scale);
        }
        const bool sx = r.ix < 0.0f, sy = r.iy < 0.0f, sz = r.iz < 0.0f;
        unsigned long long best = KEY_MISS;
        uint2 ng = make_uint2(0u, 0x80000000u);                   // the root as the only inner child of a virtual group
        int sp = 0;
        unsigned steps_nodes = 0;
        while (__any_sync(0xffffffffu, alive)) {
            bool hit = false;
            unsigned meta = 0, child_base = 0, tri_base = 0, imask = 0;
            if (alive) {
                // ---- select: nearest pending inner child of the current group (octant order, as node_select)
                const unsigned hits = ng.y;
                unsigned cc = hits >> 24, t;
                t = cc & r.order;          cc = t ? t : cc;
                t = cc & (r.order >> 8);   cc = t ? t : cc;
                t = cc & (r.order >> 16);  cc = t ? t : cc;
                const unsigned slot = (unsigned)__ffs((int)cc) - 1u;
                ng.y &= ~(0x01000000u << slot);
                if (ng.y & 0xff000000u) {
                    stack[sp] = ng;                               // all eight lanes store the same value: no hand-over needed
                    ++sp;
                }
                const unsigned node = ng.x + __popc(hits & 0xffu & ((1u << slot) - 1u));
                // ---- visit: this lane tests child slot c
                const uint4 *np = reinterpret_cast<const uint4 *>(nodes + node);
                const uint4 w0 = __ldg(np), w1 = __ldg(np + 1), w2 = __ldg(np + 2), w3 = __ldg(np + 3), w4 = __ldg(np + 4);
                if (STATS) ++steps_nodes;
                const int sh = 8 * (c & 3);
                const bool hi4 = c >= 4;
                meta = ((hi4 ? w1.w : w1.z) >> sh) & 0xffu;
                child_base = w1.x; tri_base = w1.y; imask = w0.w >> 24;
                const unsigned qlx = ((hi4 ? w2.y : w2.x) >> sh) & 0xffu, qly = ((hi4 ? w2.w : w2.z) >> sh) & 0xffu;
                const unsigned qlz = ((hi4 ? w3.y : w3.x) >> sh) & 0xffu, qhx = ((hi4 ? w3.w : w3.z) >> sh) & 0xffu;
                const unsigned qhy = ((hi4 ? w4.y : w4.x) >> sh) & 0xffu, qhz = ((hi4 ? w4.w : w4.z) >> sh) & 0xffu;
                const float adjx = __uint_as_float((w0.w & 0xffu) << 23) * r.ix;
                const float adjy = __uint_as_float(((w0.w >> 8) & 0xffu) << 23) * r.iy;
                const float adjz = __uint_as_float(((w0.w >> 16) & 0xffu) << 23) * r.iz;
                const float orgx = (__uint_as_float(w0.x) - r.ox) * r.ix;
                const float orgy = (__uint_as_float(w0.y) - r.oy) * r.iy;
                const float orgz = (__uint_as_float(w0.z) - r.oz) * r.iz;
                const float tnx = fmaf((float)(sx ? qhx : qlx), adjx, orgx - r.px), tfx = fmaf((float)(sx ? qlx : qhx), adjx, orgx + r.px);
                const float tny = fmaf((float)(sy ? qhy : qly), adjy, orgy - r.py), tfy = fmaf((float)(sy ? qly : qhy), adjy, orgy + r.py);
                const float tnz = fmaf((float)(sz ? qhz : qlz), adjz, orgz - r.pz), tfz = fmaf((float)(sz ? qlz : qhz), adjz, orgz + r.pz);
                const float tlimit = __uint_as_float((unsigned)(best >> 32)) * T_SLACK;
                const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
                const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, tlimit));
                hit = meta != 0u && tmin <= tmax;
            }
            const unsigned ball = __ballot_sync(0xffffffffu, hit);
            const unsigned hit8 = (ball >> (8 * g)) & 0xffu;
            unsigned long long key = KEY_MISS;
            if (alive) {
                // ---- triangles of this lane's child, when it is a hit leaf
                if (hit && !((imask >> c) & 1u)) {
                    const int cnt = __popc(meta >> 5);
                    const float4 *tp = reinterpret_cast<const float4 *>(tris + (tri_base + (meta & 31u)));
                    for (int k = 0; k < cnt; ++k) {
                        const float4 p0 = __ldg(tp + 3 * k), p1 = __ldg(tp + 3 * k + 1), p2 = __ldg(tp + 3 * k + 2);
                        if (STATS) ++nt;
                        float t;
                        if (tri_test(r.w, p0, p1, p2, t)) {
                            t += 0.0f;
                            const unsigned long long kk = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)__float_as_int(p0.w);
                            key = kk < key ? kk : key;
                        }
                    }
                }
            }
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, d);
                key = o < key ? o : key;
            }
            if (alive) {
                best = key < best ? key : best;
                // ---- the node's inner hits become the current group, or one is popped from the stack
                ng = make_uint2(child_base, ((hit8 & imask) << 24) | imask);
                if (!(ng.y & 0xff000000u)) {
                    if (sp == 0) alive = false;
                    else { --sp; ng = stack[sp]; }
                }
            }
            __syncwarp();
        }
        // ---- result of the ray (lane c == 0 of its group)
        const float tb = __uint_as_float((unsigned)(best >> 32));
        const int bf = (int)(unsigned)best;
        if (valid && c == 0) {
            if (t_hit) t_hit[i] = tb;
            if (face) face[i] = bf;
            if (!dir4 && (point || point64)) store_point(i, bf >= 0, tb, dcx, dcy, dcz, point, point64);
            if (peer_out) peer_store(peer_out, i, tb, bf, !dir4 && point != nullptr, dcx, dcy, dcz);
            if (STATS) { ++nray; nn += steps_nodes; if (stats->ray_nodes) stats->ray_nodes[i] = steps_nodes; }
            if (bf >= 0) {
                ++nhit;
                if (has_acc) {
                    float I = intensity ? intensity[i] : 0.0f;
                    I = I > 0.0f ? I : 0.0f;
                    atomicAdd(&acc.hist[bf], 1);
                    atomicMax(&acc.fmax[bf], __float_as_uint(I));
                }
            }
        }
    }
    // ---- per-warp totals
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        nhit += __shfl_xor_sync(0xffffffffu, nhit, d);
        if (STATS) {
            nn += __shfl_xor_sync(0xffffffffu, nn, d);
            nt += __shfl_xor_sync(0xffffffffu, nt, d);
            nray += __shfl_xor_sync(0xffffffffu, nray, d);
        }
    }
    if (lane == 0) {
        if (d_hits && nhit) atomicAdd(reinterpret_cast<unsigned long long *>(d_hits), nhit);
        if (STATS) {
            atomicAdd(&stats->rays, nray);
            atomicAdd(&stats->hits, nhit);
            atomicAdd(&stats->nodes, nn);
            atomicAdd(&stats->tris, nt);
        }
    }
}

#ifndef DP_LIGHT_FRAC
#define DP_LIGHT_FRAC 0.75f  // packets cheaper than this fraction of the mean are traced last
#endif
#ifndef DP_MIN_BLOCKS_BIG
#define DP_MIN_BLOCKS_BIG 5  // hierarchies beyond L2 (5M triangles: 0.511 -> 0.49 ms): fewer packets share an SM's L1
#endif
// FMT = 0: compressed 80-byte nodes; FMT = 1: the uncompressed 208-byte twin `fat` (L2-resident hierarchies), visited
// by the instantiation of node_step_fat for the packet's octant (its rays disagree on a sign: per-lane plane blocks).
#ifndef DP_FAT_OCT
#define DP_FAT_OCT 1
#endif

template <bool STATS, int SRC, int MINB, int FMT>
__global__ void __launch_bounds__(TR_THREADS, MINB)
k_trace(const WideNode *__restrict__ nodes, const uint4 *__restrict__ fat, const TriRec *__restrict__ tris, const float *__restrict__ d_scale,
        const float4 *__restrict__ dir4, const float *__restrict__ rays6, const float *__restrict__ intensity,
        const long long *__restrict__ d_n, long long n_max, long long total_px, int H, int W,
        const FrameXf *__restrict__ xf, float *__restrict__ t_hit, int32_t *__restrict__ face, Accum acc, int has_acc,
        unsigned long long *work_counter, long long *d_hits, TraceStats *stats, int allow_tiled,
        const OrderState *__restrict__ ord_prev, OrderState *ord_next, int prefetch, int narrow_enabled, int shard_rank,
        int shard_world, const uint32_t *__restrict__ pixel, long long n_xf, float *__restrict__ point,
        double *__restrict__ point64, const PeerOut *__restrict__ peer_out, const double *__restrict__ raytab)
{
    // dir4 == nullptr (SRC 0): rays are generated here from pixel[] and xf[], hit points written here (no k_raygen / k_points)
    // (the learnt packet lists hold packets of the previous launch over the same shard: dp_set_ray_shard drops them)
    const bool pf = prefetch != 0;
    __shared__ uint2 s_stack[STACK_SMEM * TR_THREADS];
    __shared__ unsigned s_queue[(TR_THREADS / 32) * TQ_CAP];
    __shared__ unsigned long long s_best[TR_THREADS];
    __shared__ double s_dcam[SRC == 0 ? 3 * TR_THREADS : 1];     // camera-frame directions of the packet (in-kernel rays)
    const bool want_pts = SRC == 0 && dir4 == nullptr && (point != nullptr || point64 != nullptr);
    uint2 lstack[STACK_LOCAL];
    uint2 *stack = s_stack + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *queue = s_queue + warp * TQ_CAP;
    unsigned long long *best = s_best + warp * 32;
    const unsigned lt = (1u << lane) - 1u;
    long long n = d_n ? *d_n : n_max;
    if (n > n_max) n = n_max;
    const Shard sh = shard_of(n, total_px, H, W, SRC == 0 && allow_tiled, SRC == 0 && narrow_enabled, shard_rank, shard_world);
    if (SRC == 0 && narrow_enabled && n <= DP_NARROW_MAX_RAYS) {
        // sparse frame: eight lanes per ray (work items come from the second counter)
        static_assert(STACK_SMEM * TR_THREADS >= (NR_THREADS / 8) * NR_STACK, "the narrow path borrows the packet stack");
        trace_narrow<STATS>(nodes, tris, d_scale, dir4, intensity, n, t_hit, face, acc, has_acc, work_counter + 1, d_hits, stats,
                            s_stack, sh.s_lo, sh.s_hi, pixel, H, W, xf, n_xf, point, point64, peer_out, raytab);
        return;
    }
    const bool tiled = sh.tiled;
    const float scale = __ldg(d_scale);
    unsigned nn = 0, nt = 0, n_hit_local = 0, n_ray_local = 0;
    // Packets the previous launch over the same rays found expensive (longest ray, in node steps,
    // above 2.5x / 1.5x the mean) are walked first, so that their long dependent chains overlap with
    // the bulk of the work instead of forming the tail of the kernel; the rest follows in natural order.
    const long long np = sh.p_hi - sh.p_lo;                 // packets of this launch (all of them unless sharded)
    const bool use_order = ord_prev != nullptr && ord_prev->n_valid == n;
    const long long c0 = use_order ? (long long)ord_prev->cnt[0] : 0, c1 = use_order ? (long long)ord_prev->cnt[1] : 0,
                    c2 = use_order ? (long long)ord_prev->cnt[2] : 0;
    float thr0 = __int_as_float(0x7f800000), thr1 = thr0, thr2 = -1.0f;
    if (ord_prev != nullptr && ord_prev->n_valid == n && np > 0) {
        const float mean = (float)ord_prev->cost_sum / (float)np;
        thr0 = 2.5f * mean; thr1 = 1.5f * mean; thr2 = DP_LIGHT_FRAC * mean;
    }
    if (ord_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) ord_next->n_valid = n;
    unsigned long long cost_local = 0;

    // lane 0 keeps the next packet's base one fetch ahead, so the atomic's latency hides behind a packet
    unsigned long long nextw = 0;
    if (lane == 0) nextw = atomicAdd(work_counter, 1ull);
    for (;;) {
        const long long w = (long long)__shfl_sync(0xffffffffu, nextw, 0);
        if (w >= np + c0 + c1 + c2) break;
        if (lane == 0) nextw = atomicAdd(work_counter, 1ull);
        long long packet;
        if (w < c0) packet = ord_prev->list0[w];
        else if (w < c0 + c1) packet = ord_prev->list1[w - c0];
        else if (w < c0 + c1 + np) {
            packet = sh.p_lo + (w - c0 - c1);
            if (use_order && ord_prev->flags[packet]) continue;      // traced from a list (before, or at the end)
        } else packet = ord_prev->list2[w - c0 - c1 - np];
        const long long i = packet * 32 + lane;
        bool alive = i < n;
        const bool valid = alive;
        long long slot = 0;
        RayState r;
        {
            float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
            if (valid) {
                slot = work_to_slot(i, tiled, W);
                if (SRC == 0) {
                    if (dir4) {
                        // 32 bytes per ray written by k_raygen: origin (object frame) | direction
                        const float4 o4 = __ldg(dir4 + 2 * slot), d = __ldg(dir4 + 2 * slot + 1);
                        ox = o4.x; oy = o4.y; oz = o4.z;
                        dx = d.x; dy = d.y; dz = d.z;
                    } else {
                        double dcx, dcy, dcz;
                        pixel_ray_object(pixel[slot], H, W, xf, n_xf, ox, oy, oz, dx, dy, dz, dcx, dcy, dcz, raytab);
                        if (want_pts) { s_dcam[threadIdx.x] = dcx; s_dcam[TR_THREADS + threadIdx.x] = dcy; s_dcam[2 * TR_THREADS + threadIdx.x] = dcz; }
                    }
                    if (has_acc && intensity) prefetch_l1(intensity + slot, true);   // read by the packet epilogue
                } else {
                    const float *q = rays6 + 6 * slot;
                    ox = q[0]; oy = q[1]; oz = q[2]; dx = q[3]; dy = q[4]; dz = q[5];
                }
                if (STATS) ++n_ray_local;
            }
            ray_setup(r, ox, oy, oz, dx, dy, dz, scale);
        }
        best[lane] = KEY_MISS;
        if (alive) { if (FMT == 1) node_select_fat(r, stack, lstack); else node_select(r, nodes, stack, lstack, pf); }
        __syncwarp();
        // the packet's octant, from the signs the slab test itself uses; -1 when its rays disagree
        int poct = -1;
        if (DP_OCT_SPECIALISE || (FMT == 1 && DP_FAT_OCT)) {
            const int so = (r.ix < 0.0f ? 1 : 0) | (r.iy < 0.0f ? 2 : 0) | (r.iz < 0.0f ? 4 : 0);
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            const int so0 = __shfl_sync(0xffffffffu, so, vm ? __ffs((int)vm) - 1 : 0);
            poct = __all_sync(0xffffffffu, !valid || so == so0) ? so0 : -1;
        }
        int qhead = 0, qcount = 0;              // warp-uniform
        const unsigned nn_start = nn;
        unsigned steps = 0;

        while (__any_sync(0xffffffffu, alive)) {
            unsigned tmask = 0, tbase = 0, tvalid = 0xffffffffu;
            if (alive) {
                const float tlimit = __uint_as_float((unsigned)(best[lane] >> 32)) * T_SLACK;
                if (FMT == 1) {
#if DP_FAT_OCT
                    switch (poct) {
                    case 0: alive = node_step_fat<STATS, 0>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    case 1: alive = node_step_fat<STATS, 1>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    case 2: alive = node_step_fat<STATS, 2>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    case 3: alive = node_step_fat<STATS, 3>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    case 4: alive = node_step_fat<STATS, 4>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    case 5: alive = node_step_fat<STATS, 5>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    case 6: alive = node_step_fat<STATS, 6>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    case 7: alive = node_step_fat<STATS, 7>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    default: alive = node_step_fat<STATS, -1>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn); break;
                    }
#else
                    alive = node_step_fat<STATS, -1>(r, fat, stack, lstack, tlimit, tbase, tmask, tvalid, nn);
#endif
                } else {
#if DP_OCT_SPECIALISE
                switch (poct) {
                case 0: alive = node_step<STATS, 0>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                case 1: alive = node_step<STATS, 1>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                case 2: alive = node_step<STATS, 2>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                case 3: alive = node_step<STATS, 3>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                case 4: alive = node_step<STATS, 4>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                case 5: alive = node_step<STATS, 5>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                case 6: alive = node_step<STATS, 6>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                case 7: alive = node_step<STATS, 7>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                default: alive = node_step<STATS>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf); break;
                }
#else
                alive = node_step<STATS>(r, nodes, stack, lstack, tlimit, tbase, tmask, nn, pf);
#endif
                }
                ++steps;
            }
            // queue this step's triangles: one warp prefix sum gives every lane the slots of all its triangles, which
            // it then fills in a short private loop; the queue is tested 32 pairs at a time
            if (__any_sync(0xffffffffu, tmask != 0u)) {
                const unsigned k = __popc(tmask);
                unsigned inc = k;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += y;
                }
                const int total = (int)__shfl_sync(0xffffffffu, inc, 31);
                if (qcount + total > TQ_CAP) {
                    // does not fit behind what is pending (at most 31 pairs): test those first; a step never yields
                    // more than 32 x 24 pairs, and more than TQ_CAP only on pathological nodes -> chunked below
                    if (qcount) {
                        tri_batch<STATS>(r, tris, queue, best, qhead, qcount, lane, nt);
                        qhead = (qhead + qcount) & (TQ_CAP - 1);
                        qcount = 0;
                    }
                }
                if (total <= TQ_CAP) {
                    unsigned pos = (unsigned)(qhead + qcount) + inc - k;
                    while (tmask) {
                        const int b = __ffs(tmask) - 1;
                        tmask &= tmask - 1u;
                        const unsigned rec = FMT == 1 ? tbase + __popc(tvalid & ((1u << b) - 1u)) : tbase + b;
                        queue[pos++ & (TQ_CAP - 1)] = ((unsigned)lane << 27) | rec;
                        prefetch_l1(tris + rec, pf);
                    }
                    qcount += total;
                    __syncwarp();
                    while (qcount >= TQ_FLUSH) {
                        tri_batch<STATS>(r, tris, queue, best, qhead, 32, lane, nt);
                        qhead = (qhead + 32) & (TQ_CAP - 1);
                        qcount -= 32;
                    }
                } else {
                    // more pairs than the queue holds: one per lane and round
                    for (;;) {
                        const unsigned contrib = __ballot_sync(0xffffffffu, tmask != 0u);
                        if (!contrib) break;
                        if (tmask) {
                            const int b = __ffs(tmask) - 1;
                            tmask &= tmask - 1u;
                            const unsigned rec = FMT == 1 ? tbase + __popc(tvalid & ((1u << b) - 1u)) : tbase + b;
                            queue[(qhead + qcount + __popc(contrib & lt)) & (TQ_CAP - 1)] = ((unsigned)lane << 27) | rec;
                        }
                        qcount += __popc(contrib);
                        __syncwarp();
                        if (qcount >= 32) {
                            tri_batch<STATS>(r, tris, queue, best, qhead, 32, lane, nt);
                            qhead = (qhead + 32) & (TQ_CAP - 1);
                            qcount -= 32;
                        }
                    }
                }
            }
        }
        if (qcount) tri_batch<STATS>(r, tris, queue, best, qhead, qcount, lane, nt);

        if (ord_next != nullptr) {
            const unsigned c = __reduce_max_sync(0xffffffffu, steps);
            if (lane == 0) {
                cost_local += c;
                const float cf = (float)c;
                const int cls = cf > thr0 ? 0 : (cf > thr1 ? 1 : (cf < thr2 ? 2 : 3));
                ord_next->flags[packet] = cls < 3;
                if (cls < 3) {
                    const unsigned pos = atomicAdd(&ord_next->cnt[cls], 1u);
                    (cls == 0 ? ord_next->list0 : (cls == 1 ? ord_next->list1 : ord_next->list2))[pos] = (uint32_t)packet;
                }
            }
        }
        // ---- results of the packet
        const unsigned long long key = best[lane];
        const float tb = __uint_as_float((unsigned)(key >> 32));
        const int bf = (int)(unsigned)key;
        if (valid) {
            if (t_hit) t_hit[slot] = tb;
            if (face) face[slot] = bf;
            if (STATS && stats->ray_nodes) stats->ray_nodes[slot] = nn - nn_start;
            if (want_pts)
                store_point(slot, bf >= 0, tb, s_dcam[threadIdx.x], s_dcam[TR_THREADS + threadIdx.x], s_dcam[2 * TR_THREADS + threadIdx.x],
                            point, point64);
            if (SRC == 0 && peer_out != nullptr)
                peer_store(peer_out, slot, tb, bf, want_pts && point != nullptr, s_dcam[threadIdx.x], s_dcam[TR_THREADS + threadIdx.x],
                           s_dcam[2 * TR_THREADS + threadIdx.x]);
        }
        const bool hit = valid && bf >= 0;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            ++n_hit_local;
            if (has_acc) {
                // warp-aggregated: one atomic per distinct face in the packet
                const unsigned peers = __match_any_sync(hm, bf);
                float I = intensity ? intensity[slot] : 0.0f;
                I = I > 0.0f ? I : 0.0f;
                const unsigned mx = __reduce_max_sync(peers, __float_as_uint(I));
                if (lane == __ffs(peers) - 1) {
                    // per-vertex maxima are derived from fmax when the accumulators are read (launch_vertex_max)
                    atomicAdd(&acc.hist[bf], __popc(peers));
                    atomicMax(&acc.fmax[bf], mx);
                }
            }
        }
        __syncwarp();
    }
    // ---- per-warp totals
    if (ord_next != nullptr && lane == 0 && cost_local) atomicAdd(&ord_next->cost_sum, cost_local);
    unsigned h = n_hit_local;
#pragma unroll
    for (int d = 16; d; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
    if (d_hits && lane == 0 && h) atomicAdd(reinterpret_cast<unsigned long long *>(d_hits), (unsigned long long)h);
    if (STATS) {
        unsigned long long a = nn, b = nt, c = n_ray_local;
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, d);
            b += __shfl_xor_sync(0xffffffffu, b, d);
            c += __shfl_xor_sync(0xffffffffu, c, d);
        }
        if (lane == 0) {
            atomicAdd(&stats->rays, c);
            atomicAdd(&stats->hits, (unsigned long long)h);
            atomicAdd(&stats->nodes, a);
            atomicAdd(&stats->tris, b);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Subsystem (2): pixel -> ray, float64 exactly as compute_rays (:216-221), then the object-frame
// rotation in float64 and the float32 cast of :251.  One thread per compacted pixel; dir4.w carries
// the frame index.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_raygen(const uint32_t *__restrict__ pixel, const long long *__restrict__ d_n, long long n_max, int H, int W,
         const FrameXf *__restrict__ xf, long long n_xf, float4 *__restrict__ dir4, long long total_px, int allow_tiled,
         int narrow, int shard_rank, int shard_world)
{
    long long n = *d_n;
    if (n > n_max) n = n_max;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (shard_world > 1) {
        const Shard sh = shard_of(n, total_px, H, W, allow_tiled, narrow, shard_rank, shard_world);
        if (i < sh.s_lo || i >= sh.s_hi) return;
    }
    uint32_t fr;
    double dcx, dcy, dcz;
    pixel_ray_f64(pixel[i], H, W, xf, n_xf, fr, dcx, dcy, dcz);
    const double *Ri = xf[fr].v + 4;
    float4 d;
    d.x = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[0], dcx), __dmul_rn(Ri[1], dcy)), __dmul_rn(Ri[2], dcz));
    d.y = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[3], dcx), __dmul_rn(Ri[4], dcy)), __dmul_rn(Ri[5], dcz));
    d.z = (float)__dadd_rn(__dadd_rn(__dmul_rn(Ri[6], dcx), __dmul_rn(Ri[7], dcy)), __dmul_rn(Ri[8], dcz));
    d.w = __uint_as_float(fr);
    // the ray's origin in the object frame (the frame's inverse translation), float32 like the direction
    const double *ti = xf[fr].v + 13;
    dir4[2 * i] = make_float4((float)ti[0], (float)ti[1], (float)ti[2], 0.0f);
    dir4[2 * i + 1] = d;
}

// hit point in the camera frame: p = d_cam (float64) * t (float32)   (:261-263 with origin 0)
__global__ void __launch_bounds__(256)
k_points(const uint32_t *__restrict__ pixel, const float *__restrict__ t_hit, const long long *__restrict__ d_n,
         long long n_max, int H, int W, const FrameXf *__restrict__ xf, long long n_xf, float *__restrict__ point,
         double *__restrict__ point64, long long total_px, int allow_tiled, int narrow, int shard_rank, int shard_world)
{
    long long n = *d_n;
    if (n > n_max) n = n_max;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (shard_world > 1) {
        const Shard sh = shard_of(n, total_px, H, W, allow_tiled, narrow, shard_rank, shard_world);
        if (i < sh.s_lo || i >= sh.s_hi) return;
    }
    const float tf = t_hit[i];
    const bool hit = tf < __int_as_float(0x7f800000);
    double px, py, pz;
    if (hit) {
        uint32_t fr;
        double dcx, dcy, dcz;
        pixel_ray_f64(pixel[i], H, W, xf, n_xf, fr, dcx, dcy, dcz);
        const double t = (double)tf;
        px = __dmul_rn(dcx, t); py = __dmul_rn(dcy, t); pz = __dmul_rn(dcz, t);
    } else {
        px = py = pz = __longlong_as_double(0x7ff8000000000000ll);
    }
    if (point64) { point64[3 * i] = px; point64[3 * i + 1] = py; point64[3 * i + 2] = pz; }
    if (point) { point[3 * i] = (float)px; point[3 * i + 1] = (float)py; point[3 * i + 2] = (float)pz; }
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

// vmax[v] = max over the faces incident to v of fmax[f]  (== max over the rays that hit those faces: max is
// associative, so deriving it once per read is bit-identical to accumulating it per hit)
__global__ void __launch_bounds__(256)
k_vertex_max(const uint32_t *__restrict__ fmax, const int32_t *__restrict__ F, long long nF, uint32_t *__restrict__ vmax)
{
    const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const uint32_t m = fmax[f];
    if (m == 0u) return;
    atomicMax(&vmax[F[3 * f]], m);
    atomicMax(&vmax[F[3 * f + 1]], m);
    atomicMax(&vmax[F[3 * f + 2]], m);
}

int g_order = -1;
int knob_order()
{
    if (g_order < 0) { const char *e = getenv("DP_ORDER"); g_order = e ? atoi(e) : 1; }
    return g_order;
}
int g_narrow = -1;
int knob_narrow()
{
    if (g_narrow < 0) { const char *e = getenv("DP_NARROW"); g_narrow = e ? atoi(e) : 1; }
    return g_narrow;
}
int knob_tiled()
{
    if (g_tiled < 0) { const char *e = getenv("DP_TILED"); g_tiled = e ? atoi(e) : 1; }
    return g_tiled;
}

#ifndef DP_MIN_BLOCKS_FAT
#define DP_MIN_BLOCKS_FAT 6  // 80 registers, 6 CTAs per SM: 0.228 ms against 0.234 (7 CTAs, 72 registers, spills) and 0.264 (8 CTAs)
#endif
// DP_FAT=0: trace the compressed set even when the uncompressed twin exists (read per launch: tests flip it)
int knob_fat()
{
    const char *e = getenv("DP_FAT");
    return e ? atoi(e) : 1;
}

// persistent grid (CTAs resident on the device at once) of the three variants:
// 0 = compressed nodes, hierarchy within L2; 1 = compressed, beyond L2 (fewer CTAs, L1 prefetch); 2 = uncompressed twin
int g_trace_grid[3] = {0, 0, 0};

cudaError_t trace_grid(int variant, int *grid)
{
    if (g_trace_grid[variant] == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if (variant == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<false, 0, DP_MIN_BLOCKS_BIG, 0>, TR_THREADS, 0);
        else if (variant == 2) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<false, 0, DP_MIN_BLOCKS_FAT, 1>, TR_THREADS, 0);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<false, 0, DP_MIN_BLOCKS, 0>, TR_THREADS, 0);
        if (e != cudaSuccess) return e;
        const int target = variant == 1 ? DP_MIN_BLOCKS_BIG : (variant == 2 ? DP_MIN_BLOCKS_FAT : DP_MIN_BLOCKS);
        if (per_sm > target) per_sm = target;       // the variant's point is its residency, not just its registers
        g_trace_grid[variant] = sms * (per_sm > 0 ? per_sm : 1);
    }
    *grid = g_trace_grid[variant];
    return cudaSuccess;
}

int trace_variant(const BvhView &bvh)
{
    if (bvh.fat != nullptr && knob_fat()) return 2;
    return bvh.bytes > PREFETCH_MIN_BYTES ? 1 : 0;
}

}  // namespace

cudaError_t launch_vertex_max(const uint32_t *fmax, const int32_t *F, int64_t nF, uint32_t *vmax, cudaStream_t s)
{
    if (nF <= 0) return cudaSuccess;
    k_vertex_max<<<(unsigned)((nF + 255) / 256), 256, 0, s>>>(fmax, F, nF, vmax);
    return cudaGetLastError();
}

cudaError_t launch_compute_rays(const int32_t *xs, const int32_t *ys, int64_t n, const FrameXf &xf, double *rays3,
                                cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_compute_rays<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xs, ys, n, xf, rays3);
    return cudaGetLastError();
}

cudaError_t launch_raygen(const uint32_t *pixel, const long long *d_n, int64_t n_max, int H, int W, const FrameXf *xf,
                          int64_t n_xf, float4 *dir4, cudaStream_t s, int64_t total_px, RayShard shard)
{
    if (n_max <= 0) return cudaSuccess;
    k_raygen<<<(unsigned)((n_max + 255) / 256), 256, 0, s>>>(pixel, d_n, n_max, H, W, xf, n_xf, dir4, total_px, knob_tiled(),
                                                             knob_narrow(), shard.rank, shard.world);
    return cudaGetLastError();
}

void shard_slots_host(int64_t n, int64_t total_px, int H, int W, RayShard shard, int64_t *lo, int64_t *hi)
{
    const Shard sh = shard_of(n, total_px, H, W, knob_tiled(), knob_narrow(), shard.rank, shard.world);
    *lo = sh.s_lo;
    *hi = sh.s_hi;
}

cudaError_t launch_points(const uint32_t *pixel, const float *t_hit, const long long *d_n, int64_t n_max, int H, int W,
                          const FrameXf *xf, int64_t n_xf, float *point, double *point64, cudaStream_t s, int64_t total_px,
                          RayShard shard)
{
    if (n_max <= 0 || (!point && !point64)) return cudaSuccess;
    k_points<<<(unsigned)((n_max + 255) / 256), 256, 0, s>>>(pixel, t_hit, d_n, n_max, H, W, xf, n_xf, point, point64, total_px,
                                                             knob_tiled(), knob_narrow(), shard.rank, shard.world);
    return cudaGetLastError();
}

cudaError_t launch_trace_pixels(const BvhView &bvh, const float4 *dir4, const float *intensity, const long long *d_n,
                                int64_t n_max, int64_t total_px, int H, int W, const FrameXf *xf, float *t_hit,
                                int32_t *face, const Accum *acc, unsigned long long *work_counter, long long *d_hits,
                                TraceStats *stats, const OrderState *ord_prev, OrderState *ord_next, cudaStream_t s,
                                bool counter_zeroed, RayShard shard, const uint32_t *pixel, int64_t n_xf, float *point,
                                double *point64, const PeerOut *peer_out, const double *raytab)
{
    if (n_max <= 0) return cudaSuccess;
    cudaError_t e;
    int grid = 0;
    const int variant = trace_variant(bvh);
    const int pf = variant == 1;
    if ((e = trace_grid(variant, &grid)) != cudaSuccess) return e;
    const long long want = (n_max + TR_THREADS - 1) / TR_THREADS;
    if (want < grid) grid = (int)want;
    if (!counter_zeroed && (e = cudaMemsetAsync(work_counter, 0, 2 * sizeof(unsigned long long), s)) != cudaSuccess) return e;
    Accum a = acc ? *acc : Accum{nullptr, nullptr, nullptr, nullptr};
    if (knob_order() == 0) { ord_prev = nullptr; ord_next = nullptr; }
    const int narrow = knob_narrow() && d_n != nullptr;
#define DP_LAUNCH_TRACE0(ST, MB, FMT)                                                                                        \
    k_trace<ST, 0, MB, FMT><<<grid, TR_THREADS, 0, s>>>(bvh.nodes, bvh.fat, bvh.tris, bvh.d_scale, dir4, nullptr, intensity, d_n, \
                                                        n_max, total_px, H, W, xf, t_hit, face, a, acc != nullptr, work_counter, \
                                                        d_hits, stats, knob_tiled(), ord_prev, ord_next, pf, narrow, shard.rank, \
                                                        shard.world, pixel, (long long)n_xf, point, point64, peer_out, raytab)
    if (stats) {
        if (variant == 2) DP_LAUNCH_TRACE0(true, DP_MIN_BLOCKS_FAT, 1);
        else if (variant == 1) DP_LAUNCH_TRACE0(true, DP_MIN_BLOCKS_BIG, 0);
        else DP_LAUNCH_TRACE0(true, DP_MIN_BLOCKS, 0);
    } else {
        if (variant == 2) DP_LAUNCH_TRACE0(false, DP_MIN_BLOCKS_FAT, 1);
        else if (variant == 1) DP_LAUNCH_TRACE0(false, DP_MIN_BLOCKS_BIG, 0);
        else DP_LAUNCH_TRACE0(false, DP_MIN_BLOCKS, 0);
    }
#undef DP_LAUNCH_TRACE0
    return cudaGetLastError();
}

cudaError_t launch_trace_rays6(const BvhView &bvh, const float *rays6, int64_t n, float *t_hit, int32_t *face,
                               unsigned long long *work_counter, TraceStats *stats, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    cudaError_t e;
    int grid = 0;
    const int variant = trace_variant(bvh);
    const int pf = variant == 1;
    if ((e = trace_grid(variant, &grid)) != cudaSuccess) return e;
    const long long want = (n + TR_THREADS - 1) / TR_THREADS;
    if (want < grid) grid = (int)want;
    if ((e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s)) != cudaSuccess) return e;
    Accum a{nullptr, nullptr, nullptr, nullptr};
#define DP_LAUNCH_TRACE1(ST, MB, FMT)                                                                                          \
    k_trace<ST, 1, MB, FMT><<<grid, TR_THREADS, 0, s>>>(bvh.nodes, bvh.fat, bvh.tris, bvh.d_scale, nullptr, rays6, nullptr, nullptr, \
                                                        n, 0, 0, 0, nullptr, t_hit, face, a, 0, work_counter, nullptr, stats, 0, \
                                                        nullptr, nullptr, pf, 0, 0, 1, nullptr, 0, nullptr, nullptr, nullptr, nullptr)
    if (stats) {
        if (variant == 2) DP_LAUNCH_TRACE1(true, DP_MIN_BLOCKS_FAT, 1);
        else if (variant == 1) DP_LAUNCH_TRACE1(true, DP_MIN_BLOCKS_BIG, 0);
        else DP_LAUNCH_TRACE1(true, DP_MIN_BLOCKS, 0);
    } else {
        if (variant == 2) DP_LAUNCH_TRACE1(false, DP_MIN_BLOCKS_FAT, 1);
        else if (variant == 1) DP_LAUNCH_TRACE1(false, DP_MIN_BLOCKS_BIG, 0);
        else DP_LAUNCH_TRACE1(false, DP_MIN_BLOCKS, 0);
    }
#undef DP_LAUNCH_TRACE1
    return cudaGetLastError();
}

}  // namespace dp
