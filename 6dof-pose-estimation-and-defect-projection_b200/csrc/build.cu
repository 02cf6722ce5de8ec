// Subsystem (3): BVH construction on the GPU.
//
// Replaces RaycastingScene().add_triangles (/root/reference/src/defect_projection.py:253-254,
// Embree's rtcCommitScene on the host) and the mesh re-posing of :549-550 (float64 transform,
// then the float32 cast of from_legacy, :245).
//
//   scene bounds -> 30-bit Morton codes of the triangle-box centres -> in-house LSD radix sort
//   (4 x 8 bits, stable, key + triangle index) -> Karras 2012 binary hierarchy (duplicate codes
//   disambiguated by the index) -> bottom-up binary boxes -> top-down, surface-area-guided
//   collapse into 8-wide nodes with <= 3 triangles per leaf slot, children placed in the slot
//   whose octant matches their direction from the node centre -> bottom-up fit that quantises
//   the child boxes to 8 bits per plane (conservatively, in double precision).
//
// The same bottom-up fit (ONE launch, k_fit_all) is the refit used after dp_pose_mesh: the topology (meta
// bytes, child and triangle bases, parent links) is shared, only boxes and triangle records are recomputed.
#include "dp_internal.cuh"


#include <math.h>

#include <algorithm>
#include <stdlib.h>

namespace dp {

namespace {

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ void tri_box(const float *__restrict__ V, const int32_t *__restrict__ F, long long f,
                                        float lo[3], float hi[3])
{
    const long long a = F[3 * f], b = F[3 * f + 1], c = F[3 * f + 2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = V[3 * a + k], y = V[3 * b + k], z = V[3 * c + k];
        lo[k] = fminf(x, fminf(y, z));
        hi[k] = fmaxf(x, fmaxf(y, z));
    }
}

// bounds_u[0..2] = min (ordered uint), [3..5] = max.  Grid-stride loop, warp then block reduction, six atomics per
// block (one atomic set per warp made the kernel atomics-bound: 0.61 ms at 5M triangles).
__global__ void __launch_bounds__(256)
k_scene_bounds(const float *__restrict__ V, const int32_t *__restrict__ F, long long n, unsigned *bounds_u)
{
    __shared__ unsigned s_mn[8][3], s_mx[8][3];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float a[3], b[3];
        tri_box(V, F, i, a, b);
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], a[k]); hi[k] = fmaxf(hi[k], b[k]); }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned mn = __reduce_min_sync(0xffffffffu, f2ord(lo[k]));
        const unsigned mx = __reduce_max_sync(0xffffffffu, f2ord(hi[k]));
        if (lane == 0) { s_mn[warp][k] = mn; s_mx[warp][k] = mx; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned mn = s_mn[0][threadIdx.x], mx = s_mx[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) { mn = min(mn, s_mn[w][threadIdx.x]); mx = max(mx, s_mx[w][threadIdx.x]); }
        atomicMin(&bounds_u[threadIdx.x], mn);
        atomicMax(&bounds_u[3 + threadIdx.x], mx);
    }
}

__global__ void k_finish_bounds(const unsigned *bounds_u, float *sbounds, float *d_scale, long long n)
{
    float s = 0.0f;
    for (int k = 0; k < 6; ++k) {
        float v = n > 0 ? ord2f(bounds_u[k]) : 0.0f;
        sbounds[k] = v;
        s = fmaxf(s, fabsf(v));
    }
    *d_scale = s;
}

__device__ __forceinline__ unsigned expand10(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// code = interleave(x,y,z), x most significant; cell = min(1023, (c - lo) * (1024/ext))
// (also clears the binary tree's parent links and completion flags for the passes that follow: two memsets less in
// the build's stream)
__global__ void k_morton(const float *__restrict__ V, const int32_t *__restrict__ F, long long n,
                         const float *__restrict__ sbounds, uint32_t *keys, uint32_t *vals, int32_t *parent, int *flags)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (parent) { parent[2 * i] = -1; parent[2 * i + 1] = -1; }
    if (flags) flags[i] = 0;
    float lo[3], hi[3];
    tri_box(V, F, i, lo, hi);
    unsigned q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float c = __fmul_rn(0.5f, __fadd_rn(lo[k], hi[k]));
        const float ext = __fsub_rn(sbounds[3 + k], sbounds[k]);
        float g = 0.0f;
        if (ext > 0.0f) g = __fmul_rn(__fsub_rn(c, sbounds[k]), __fdiv_rn(1024.0f, ext));
        g = fminf(fmaxf(g, 0.0f), 1023.0f);
        q[k] = (unsigned)g;
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------------------------
// LSD radix sort, RS_BITS = 10 bits per pass, stable: 30-bit Morton codes take three passes (8-bit digits took four),
// the cell keys of the point-cloud grids two.  Tile = 256 threads x 16 keys; every warp owns a contiguous 512-key slice
// of the tile so that warp-local ranks (match_any) are stable.
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef DP_RS_BITS
#define DP_RS_BITS 10
#endif
constexpr int RS_BITS = DP_RS_BITS;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_DPT = RS_BINS / RS_THREADS;     // digits per thread where a thread owns digits
static_assert(RS_BINS % RS_THREADS == 0 && RS_DPT >= 1, "digits are dealt out to the threads of a block");

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint32_t *__restrict__ keys, long long n, int shift, uint32_t *__restrict__ table, int nblocks)
{
    __shared__ unsigned h[RS_BINS];
    for (int k = threadIdx.x; k < RS_BINS; k += RS_THREADS) h[k] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < RS_BINS; k += RS_THREADS) table[(long long)k * nblocks + blockIdx.x] = h[k];
}

// Digit table: table[d * nb + b] = count of digit d in tile b.  One block per digit scans its row in place
// (exclusive, over the tiles) and leaves the row total in totals[d]; the scatter kernel adds the exclusive scan of
// the digit totals itself.
__global__ void __launch_bounds__(256) k_rs_scan_rows(uint32_t *table, int nb, uint32_t *totals)
{
    __shared__ unsigned s_warp[8];
    __shared__ unsigned s_carry;
    uint32_t *row = table + (long long)blockIdx.x * nb;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 256 * 4) {
        const int i0 = base + threadIdx.x * 4;
        unsigned v[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const unsigned x = i0 + q < nb ? row[i0 + q] : 0u; v[q] = sum; sum += x; }
        unsigned inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned woff = 0, wtot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const unsigned c = s_warp[w]; if (w < warp) woff += c; wtot += c; }
        const unsigned off = s_carry + woff + inc - sum;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (i0 + q < nb) row[i0 + q] = off + v[q];
        __syncthreads();
        if (threadIdx.x == 0) s_carry += wtot;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t *__restrict__ keys_out,
             uint32_t *__restrict__ vals_out, long long n, int shift, const uint32_t *__restrict__ table, int nblocks,
             const uint32_t *__restrict__ totals)
{
    __shared__ unsigned cnt[RS_WARPS][RS_BINS];
    __shared__ unsigned s_wsum[RS_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < RS_WARPS * RS_BINS; k += RS_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE + (long long)warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS];
    unsigned short rank[RS_ITEMS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = base + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys[i] : 0u;
        val[r] = ok ? vals[i] : 0u;
        const unsigned digit = ok ? ((key[r] >> shift) & (RS_BINS - 1)) : ((unsigned)RS_BINS | (unsigned)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        unsigned pre = 0;
        if (ok) pre = cnt[warp][digit];
        __syncwarp();
        rank[r] = (unsigned short)(pre + __popc(peers & lt));
        if (ok && lane == (31 - __clz(peers))) cnt[warp][digit] = pre + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        // exclusive scan of the digit totals: thread t owns the RS_DPT consecutive digits t * RS_DPT ...
        unsigned tot[RS_DPT], mine = 0;
#pragma unroll
        for (int q = 0; q < RS_DPT; ++q) { tot[q] = totals[tid * RS_DPT + q]; mine += tot[q]; }
        unsigned inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        unsigned dbase = inc - mine;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
            if (w < warp) dbase += s_wsum[w];
        // ... and turns the per-warp counts of each of them into starting offsets
#pragma unroll
        for (int q = 0; q < RS_DPT; ++q) {
            const int d = tid * RS_DPT + q;
            unsigned run = dbase + table[(long long)d * nblocks + blockIdx.x];
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                const unsigned c = cnt[w][d];
                cnt[w][d] = run;
                run += c;
            }
            dbase += tot[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = base + r * 32 + lane;
        if (i < n) {
            const unsigned digit = (key[r] >> shift) & (RS_BINS - 1);
            const unsigned pos = cnt[warp][digit] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Karras 2012.  Node ids: internal i in [0, n-1), leaf j -> (n-1) + j.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint32_t *__restrict__ keys, long long n, long long i, long long j)
{
    if (j < 0 || j >= n) return -1;
    const uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((unsigned)(i ^ j));
    return __clz(a ^ b);
}

__global__ void k_karras(const uint32_t *__restrict__ keys, long long n, int32_t *left, int32_t *right,
                         int32_t *parent, int32_t *first, int32_t *last)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    long long lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    long long l = 0;
    for (long long t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const long long j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    long long s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const long long gamma = i + s * d + (d < 0 ? -1 : 0);
    const long long lo = i < j ? i : j, hi = i < j ? j : i;
    const int32_t lc = (int32_t)((lo == gamma) ? (n - 1) + gamma : gamma);
    const int32_t rc = (int32_t)((hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1);
    left[i] = lc;
    right[i] = rc;
    parent[lc] = (int32_t)i;
    parent[rc] = (int32_t)i;
    first[i] = (int32_t)lo;
    last[i] = (int32_t)hi;
}

// Surface-area cost model of the collapse (Ylitie, Karras, Laine 2017, section 3.1, restated for this node
// layout): C(n, i) = cheapest way to represent the binary subtree n as at most i children of one wide node.
//   C(n, 1) = min(C_leaf, C_distribute(n, 8) + A_n * c_node),   C_leaf = A_n * P_n * c_prim  (P_n <= LEAF_MAX)
//   C(n, i) = min(C_distribute(n, i), C(n, i - 1)),              i = 2..7
//   C_distribute(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
// ctab[n][i-1] holds C(n, i) for the internal nodes; a single triangle costs A * c_prim for every i.
constexpr float C_NODE = 1.0f;

__device__ __forceinline__ float half_area(const float lo[3], const float hi[3])
{
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ float distribute_cost(const float cl[7], const float cr[7], int j, int *split)
{
    float best = INFINITY;
    int bk = 1;
    const int k0 = j - 7 > 1 ? j - 7 : 1, k1 = j - 1 < 7 ? j - 1 : 7;
    for (int k = k0; k <= k1; ++k) {
        const float c = cl[k - 1] + cr[j - k - 1];
        if (c < best) { best = c; bk = k; }
    }
    if (split) *split = bk;
    return best;
}

// Boxes of the binary nodes (internal 0..n-2, leaves n-1..2n-2): 32 bytes per node = lo.xyz, -, hi.xyz, -, read and
// written as two 16-byte words (one sector per node; the separate 12-byte-stride lo / hi arrays cost six scalar loads
// from two sectors per box, and k_binfit -- the largest kernel of a build -- is made of such loads).
__device__ __forceinline__ void box_load(const float *bbox, long long id, float lo[3], float hi[3])
{
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(bbox + 8 * id));
    const float4 b = __ldcg(reinterpret_cast<const float4 *>(bbox + 8 * id + 4));
    lo[0] = a.x; lo[1] = a.y; lo[2] = a.z;
    hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
}
__device__ __forceinline__ void box_store(float *bbox, long long id, const float lo[3], const float hi[3])
{
    *reinterpret_cast<float4 *>(bbox + 8 * id) = make_float4(lo[0], lo[1], lo[2], 0.0f);
    *reinterpret_cast<float4 *>(bbox + 8 * id + 4) = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// C(id, 1..7) of a node from its children (tables of internal children from ctab, read past L1: another thread may
// have written them; a single triangle costs A * c_prim for every i)
__device__ __forceinline__ void load_table(long long id, long long n, const float *ctab, const float *bbox, float c_prim,
                                           float out[7])
{
    if (id >= n - 1) {
        float lo[3], hi[3];
        box_load(bbox, id, lo, hi);
        const float c = half_area(lo, hi) * c_prim;
#pragma unroll
        for (int i = 0; i < 7; ++i) out[i] = c;
    } else {
        const float4 a = __ldcg(reinterpret_cast<const float4 *>(ctab + 8 * id));
        const float4 b = __ldcg(reinterpret_cast<const float4 *>(ctab + 8 * id + 4));
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z;
    }
}

__device__ __forceinline__ void node_table(const float cl[7], const float cr[7], float A, int P, float c_prim, float C[7])
{
    const float c_leaf = P <= LEAF_MAX ? A * (float)P * c_prim : INFINITY;
    C[0] = fminf(c_leaf, distribute_cost(cl, cr, 8, nullptr) + A * C_NODE);
#pragma unroll
    for (int i = 2; i <= 7; ++i) C[i - 1] = fminf(distribute_cost(cl, cr, i, nullptr), C[i - 2]);
}

// What the collapse's selection walk needs of a binary node, in ONE 32-byte record (one sector) written by k_binfit:
// the walk used to gather it from seven arrays (left, right, first, last, blo, bhi, ctab), seven cache lines per node.
struct __align__(32) BinRec {
    int32_t left, right, count;
    float area;                    // dx*dy + dy*dz + dz*dx of the node's box
    unsigned long long dec;        // node_decisions() (0 without cost tables)
    unsigned long long pad_;
};
static_assert(sizeof(BinRec) == 32, "binary node record must be 32 bytes");

// The cut decisions of a node, precomputed where its tables are in registers, so that the collapse's selection walk is
// one 8-byte load per visited binary node instead of a re-evaluation of the tables of both children (fourteen floats
// from two more cache lines, min-plus loops over local-memory arrays: 60-90 us per LEVEL of the collapse).
//   bits 7(b-2) .. 7(b-2)+6, b = 2..8:  eff(b) in 4 bits = the largest b' <= b with C_distribute(n, b') < C(n, b'-1),
//                                       else 1 ("this node is one child");  split(b) in 3 bits = the k of C_distribute(n, b)
//   bit 63: as ONE child the node is an inner child (its own wide node), not a leaf slot
// Exactly the comparisons the walk used to make, on the same floats: the wide tree is unchanged.
__device__ __forceinline__ unsigned long long node_decisions(const float cl[7], const float cr[7], const float C[7], float A,
                                                             int P, float c_prim)
{
    unsigned long long d = 0;
    int eff = 1;
#pragma unroll
    for (int b = 2; b <= 8; ++b) {
        int k;
        const float dc = distribute_cost(cl, cr, b, &k);
        if (dc < C[b - 2]) eff = b;
        d |= (unsigned long long)(unsigned)(eff | (k << 4)) << (7 * (b - 2));
    }
    const bool inner = P > LEAF_MAX || !(A * (float)P * c_prim == C[0]);
    if (inner) d |= 1ull << 63;
    return d;
}
__device__ __forceinline__ int dec_eff(unsigned long long d, int b) { return b < 2 ? 1 : (int)((d >> (7 * (b - 2))) & 15u); }
__device__ __forceinline__ int dec_split(unsigned long long d, int b) { return (int)((d >> (7 * (b - 2) + 4)) & 7u); }

// Tree rotations at `cur` (Kensler 2008), applied by the thread that completes the node on the way up.  With
// children L = (l0, l1) and R = (r0, r1) the candidates are
//   child <-> grandchild:       L changes places with r_j (R keeps the other one), or R with l_i;
//   grandchild <-> grandchild:  l_i changes places with r_j  (DP_ROTATE_GG, untangles two interleaved children);
// the one that shrinks the summed surface area of the restructured children most is applied.  The node's own box
// and triangle set do not change, so nothing above is affected, and both subtrees are complete, so nobody else is
// reading them.  A restructured child's triangles stop being one interval of the sorted order: its last[] is
// rewritten so that last - first + 1 stays its triangle count (the only use of the interval of a node with more
// than LEAF_MAX triangles); candidates that would leave such a child with LEAF_MAX triangles or fewer are skipped.
struct RotBox { float lo[3], hi[3]; };
__device__ __forceinline__ RotBox rot_union(const RotBox &a, const RotBox &b)
{
    RotBox r;
#pragma unroll
    for (int k = 0; k < 3; ++k) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return r;
}

__device__ __forceinline__ void try_rotate(long long cur, long long n, int32_t *left, int32_t *right, int32_t *parent,
                                           const int32_t *first, int32_t *last, float *bbox, float *ctab,
                                           float c_prim, int gg, BinRec *rec)
{
    auto count = [&](long long id) -> int { return id < n - 1 ? last[id] - first[id] + 1 : 1; };
    auto box = [&](long long id) -> RotBox {
        RotBox r;
        box_load(bbox, id, r.lo, r.hi);
        return r;
    };
    // make node `id` the parent of (c0, c1) with the given box and count, and refresh its cost table
    auto rebuild = [&](long long id, long long c0, long long c1, const RotBox &bx, int cnt) {
        left[id] = (int32_t)c0; right[id] = (int32_t)c1;
        parent[c0] = (int32_t)id; parent[c1] = (int32_t)id;
        box_store(bbox, id, bx.lo, bx.hi);
        last[id] = first[id] + cnt - 1;
        unsigned long long d = 0;
        if (ctab) {
            float cl[7], cr[7], C[7];
            load_table(c0, n, ctab, bbox, c_prim, cl);
            load_table(c1, n, ctab, bbox, c_prim, cr);
            node_table(cl, cr, half_area(bx.lo, bx.hi), cnt, c_prim, C);
            *reinterpret_cast<float4 *>(ctab + 8 * id) = make_float4(C[0], C[1], C[2], C[3]);
            *reinterpret_cast<float4 *>(ctab + 8 * id + 4) = make_float4(C[4], C[5], C[6], 0.0f);
            d = node_decisions(cl, cr, C, half_area(bx.lo, bx.hi), cnt, c_prim);
        }
        BinRec q;
        q.left = (int32_t)c0; q.right = (int32_t)c1; q.count = cnt; q.area = half_area(bx.lo, bx.hi); q.dec = d; q.pad_ = 0;
        rec[id] = q;
    };
    const long long ch[2] = {left[cur], right[cur]};
    const bool inner[2] = {ch[0] < n - 1, ch[1] < n - 1};
    if (!inner[0] && !inner[1]) return;
    RotBox cb[2] = {box(ch[0]), box(ch[1])};
    long long g[2][2] = {{-1, -1}, {-1, -1}};
    RotBox gb[2][2];
    int gc[2][2] = {{0, 0}, {0, 0}};
    float area[2];
    int cc[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        area[s] = half_area(cb[s].lo, cb[s].hi);
        cc[s] = count(ch[s]);
        if (!inner[s]) continue;
        g[s][0] = left[ch[s]]; g[s][1] = right[ch[s]];
#pragma unroll
        for (int w = 0; w < 2; ++w) { gb[s][w] = box(g[s][w]); gc[s][w] = count(g[s][w]); }
    }
    float best_gain = 0.0f;
    int best = -1;                       // 0..3: child 1-s <-> g[s][w] (code 2*s + w);  4..7: g[0][i] <-> g[1][j] (code 4 + 2*i + j)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (!inner[s]) continue;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            if (cc[1 - s] + gc[s][1 - w] <= LEAF_MAX) continue;
            const RotBox nb = rot_union(cb[1 - s], gb[s][1 - w]);
            const float gain = area[s] - half_area(nb.lo, nb.hi);
            if (gain > best_gain && gain > 1e-6f * area[s]) { best_gain = gain; best = 2 * s + w; }
        }
    }
    if (gg && inner[0] && inner[1]) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (gc[1][j] + gc[0][1 - i] <= LEAF_MAX || gc[0][i] + gc[1][1 - j] <= LEAF_MAX) continue;
                const RotBox n0 = rot_union(gb[1][j], gb[0][1 - i]), n1 = rot_union(gb[0][i], gb[1][1 - j]);
                const float gain = area[0] + area[1] - half_area(n0.lo, n0.hi) - half_area(n1.lo, n1.hi);
                if (gain > best_gain && gain > 1e-6f * (area[0] + area[1])) { best_gain = gain; best = 4 + 2 * i + j; }
            }
    }
    if (best < 0) return;
    if (best < 4) {
        const int s = best >> 1, w = best & 1;
        const long long B = ch[s], A = ch[1 - s], moved = g[s][w], stay = g[s][1 - w];
        if (left[cur] == (int32_t)A) left[cur] = (int32_t)moved; else right[cur] = (int32_t)moved;
        parent[moved] = (int32_t)cur;
        rebuild(B, A, stay, rot_union(cb[1 - s], gb[s][1 - w]), cc[1 - s] + gc[s][1 - w]);
    } else {
        const int i = (best >> 1) & 1, j = best & 1;
        rebuild(ch[0], g[1][j], g[0][1 - i], rot_union(gb[1][j], gb[0][1 - i]), gc[1][j] + gc[0][1 - i]);
        rebuild(ch[1], g[0][i], g[1][1 - j], rot_union(gb[0][i], gb[1][1 - j]), gc[0][i] + gc[1][1 - j]);
    }
}

// Bottom-up pass over the binary tree, one thread per triangle walking towards the root; the second thread to arrive
// at a node completes it: box, optional rotation (nodes of at most rot_max triangles), cost table of the collapse.
__global__ void k_binfit(const float *__restrict__ V, const int32_t *__restrict__ F, const uint32_t *__restrict__ sorted_tri,
                         long long n, int32_t *parent, int32_t *left, int32_t *right, const int32_t *first, int32_t *last,
                         float *bbox, int *flags, float *ctab, float c_prim, int rot_min, int rot_max, int rot_gg,
                         BinRec *rec, unsigned *counters, int32_t *wroot, unsigned *cres_words, int n_cres_words)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // the collapse's start state (node 0 = the root, its counters, the cooperative launch's result / barrier block):
    // written here instead of by three small copies in the stream
    if (counters && j < 4) counters[j] = j == 0 ? 1u : 0u;
    if (wroot && j == 0) wroot[0] = 0;
    if (cres_words && j < n_cres_words) cres_words[j] = 0u;
    if (j >= n) return;
    float lo[3], hi[3];
    tri_box(V, F, sorted_tri[j], lo, hi);
    long long id = (n - 1) + j;
    box_store(bbox, id, lo, hi);
    if (n == 1) return;
    long long cur = parent[id];
    for (;;) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1) == 0) return;       // the sibling subtree is not done yet
        const int P = last[cur] - first[cur] + 1;
        if (rot_max > 0 && P <= rot_max && P >= rot_min && P > 2 * LEAF_MAX) try_rotate(cur, n, left, right, parent, first, last, bbox, ctab, c_prim, rot_gg, rec);
        const long long L = left[cur], R = right[cur];
        {
            float llo[3], lhi[3], rlo[3], rhi[3];
            box_load(bbox, L, llo, lhi);
            box_load(bbox, R, rlo, rhi);
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(llo[k], rlo[k]); hi[k] = fmaxf(lhi[k], rhi[k]); }
            box_store(bbox, cur, lo, hi);
        }
        unsigned long long d = 0;
        if (ctab) {
            float cl[7], cr[7], C[7];
            load_table(L, n, ctab, bbox, c_prim, cl);
            load_table(R, n, ctab, bbox, c_prim, cr);
            node_table(cl, cr, half_area(lo, hi), P, c_prim, C);
            *reinterpret_cast<float4 *>(ctab + 8 * cur) = make_float4(C[0], C[1], C[2], C[3]);
            *reinterpret_cast<float4 *>(ctab + 8 * cur + 4) = make_float4(C[4], C[5], C[6], 0.0f);
            d = node_decisions(cl, cr, C, half_area(lo, hi), P, c_prim);
        }
        {
            BinRec q;
            q.left = (int32_t)L; q.right = (int32_t)R; q.count = P; q.area = half_area(lo, hi); q.dec = d; q.pad_ = 0;
            rec[cur] = q;
        }
        if (cur == 0) return;
        cur = parent[cur];
    }
}

// ------------------------------------------------------------------------------------------
// collapse: EIGHT lanes per wide node of the current level.  Lane 0 of the group picks the children (a short serial
// walk over the cost tables); the slot assignment, the child/record allocation and the node words are then
// produced by the group, one candidate child per lane (the levels near the root hold 1, 8, 64 ... nodes, so the
// latency of one node is the latency of the level).
// ------------------------------------------------------------------------------------------
// MODE 0: selection and emission fused (small levels: one launch, lowest latency)
// MODE 1: selection only, ONE THREAD per node, result to `sel` (large levels: all lanes busy with the serial walk)
// MODE 2: emission only, eight lanes per node, selection read from `sel`
struct CollapseArgs {
    long long n;
    const int32_t *left, *right, *first, *last;
    const float *bbox;
    const uint32_t *sorted_tri;
    int32_t *wroot;
    WideNode *nodes;
    int32_t *tri_face;
    unsigned *counters;
    const float *ctab;
    float c_prim;
    int greedy_mode;
    int32_t *wparent;
    int dp_max_count;
    int32_t *sel;
    long long cap_nodes;
    const BinRec *rec;                 // per internal binary node: children, count, area, cut decisions
    unsigned *arrived;                 // [cap_nodes] arrival counters of the bottom-up fit, cleared as the nodes are emitted
};

// one wide node `w` of the level that starts at `begin`; MODE 1: called by ONE thread (gl = 0), else by the eight lanes
// gl = 0..7 of an aligned group
// MODE 0 and 2 must be called by ALL threads of a 256-thread block (`active` = this group has a node): the children
// and records of the block's 32 nodes are allocated with ONE pair of global atomics per block (shared-memory
// aggregation; one pair per node made the two counters the bottleneck of the large levels: 0.65 ms at 5M triangles).
template <int MODE>
__device__ __forceinline__ void collapse_node(const CollapseArgs &A, long long w, long long begin, int gl, bool active = true)
{
    const long long n = A.n;
    const int32_t *__restrict__ left = A.left, *__restrict__ right = A.right, *__restrict__ first = A.first,
                  *__restrict__ last = A.last;
    const float *__restrict__ bbox = A.bbox, *__restrict__ ctab = A.ctab;
    const uint32_t *__restrict__ sorted_tri = A.sorted_tri;
    int32_t *wroot = A.wroot, *tri_face = A.tri_face, *wparent = A.wparent, *sel = A.sel;
    WideNode *nodes = A.nodes;
    unsigned *counters = A.counters;
    const float c_prim = A.c_prim;
    const int greedy_mode = A.greedy_mode, dp_max_count = A.dp_max_count;
    const long long cap_nodes = A.cap_nodes;
    const int lane32 = threadIdx.x & 31;
    const unsigned gmask = 0xffu << (lane32 & 24);
    const int gbase = lane32 & 24;
    const int32_t r = active ? wroot[w] : 0;
    unsigned inner_mask = 0;
    auto count = [&](int32_t id) -> int { return id < n - 1 ? last[id] - first[id] + 1 : 1; };
    int32_t cand[8];
    float carea[8];
    int nc = 0;
    // cost tables decide the cut of subtrees of up to dp_max_count triangles (0 = all); above that the greedy
    // largest-area expansion keeps the upper levels balanced
    const BinRec *__restrict__ rec = A.rec;
    const bool use_dp = ctab != nullptr && (dp_max_count <= 0 || count(r) <= dp_max_count);
    if (!active) {
        nc = 0;                                             // takes part in the block's allocation with nothing to allocate
    } else if (gl != 0 || MODE == 2) {
        // lanes 1..7 wait for lane 0's choice; MODE 2 reads it below
    } else if (use_dp) {
        // cost-optimal cut of the binary subtree, from the decisions k_binfit stored with the tables: a visited node is
        // one 8-byte load (plus its two child ids); `inner` marks the children that become wide nodes themselves
        nc = 0;
        inner_mask = 0;
        if (r >= n - 1 || (w == 0 && !(rec[r].dec >> 63))) {
            cand[0] = r; nc = 1;                                            // a whole mesh of <= LEAF_MAX triangles
        } else {
            int32_t st_id[16];
            int st_b[16], sp = 0;
            st_id[0] = r; st_b[0] = 8; sp = 1;
            bool force = true;                                              // the root of a wide node always distributes
            while (sp > 0) {
                --sp;
                const int32_t id = st_id[sp];
                int b = st_b[sp];
                if (id >= n - 1) { cand[nc++] = id; continue; }
                const BinRec q = rec[id];
                if (!force) {
                    b = dec_eff(q.dec, b);
                    if (b == 1) {
                        if (q.dec >> 63) inner_mask |= 1u << nc;
                        cand[nc++] = id;
                        continue;
                    }
                }
                force = false;
                const int k = dec_split(q.dec, b);
                st_id[sp] = q.right; st_b[sp] = b - k; ++sp;
                st_id[sp] = q.left; st_b[sp] = k; ++sp;
            }
        }
    } else {
        // greedy largest-priority expansion, one record load per new candidate (both children of an expansion in parallel)
        auto prio_of = [&](const BinRec &q) -> float {
            if (greedy_mode == 2) return (float)q.count;
            if (greedy_mode == 3) return q.area * (float)q.count;
            if (greedy_mode == 4) return q.area * sqrtf((float)q.count);
            return q.area;
        };
        int32_t cleft[8], cright[8];
        auto fetch = [&](int32_t id, int slot) {
            cand[slot] = id; carea[slot] = -1.0f; cleft[slot] = cright[slot] = -1;
            if (id < n - 1) {
                const BinRec q = rec[id];
                if (q.count > LEAF_MAX) { carea[slot] = prio_of(q); cleft[slot] = q.left; cright[slot] = q.right; inner_mask |= 1u << slot; }
            }
        };
        inner_mask = 0;
        fetch(r, 0);
        nc = 1;
        if (inner_mask) {
            // the root always expands; then the candidate with the largest priority, until eight
            inner_mask = 0;
            const int32_t l0 = cleft[0], r0 = cright[0];
            fetch(l0, 0);
            fetch(r0, 1);
            nc = 2;
            while (nc < 8) {
                int b = -1;
                float ba = -1.0f;
                for (int j = 0; j < nc; ++j)
                    if (carea[j] > ba) { ba = carea[j]; b = j; }
                if (b < 0) break;
                const int32_t l = cleft[b], rr = cright[b];
                inner_mask &= ~(1u << b);
                fetch(l, b);
                fetch(rr, nc);
                ++nc;
            }
        }
    }
    if (MODE == 1) {
        int32_t *o = sel + 9 * (w - begin);
        for (int c = 0; c < 8; ++c) o[c] = c < nc ? cand[c] : -1;
        o[8] = nc | (int)(inner_mask << 8);
        return;
    }
    if (MODE == 2 && gl == 0 && active) {
        const int32_t *o = sel + 9 * (w - begin);
        for (int c = 0; c < 8; ++c) cand[c] = o[c];
        nc = o[8] & 0xff;
        inner_mask = (unsigned)o[8] >> 8;
    }
    // ---- the group takes over: candidate c lives in lane c
    nc = __shfl_sync(gmask, nc, gbase);
    inner_mask = __shfl_sync(gmask, inner_mask, gbase);
    int32_t my = -1;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int32_t v = __shfl_sync(gmask, cand[c], gbase);
        if (c == gl) my = v;
    }
    const bool have = gl < nc;
    const bool my_inner = have && ((inner_mask >> gl) & 1u);
    float cen[3] = {0.f, 0.f, 0.f}, lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (have) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = bbox[8ll * my + k]; hi[k] = bbox[8ll * my + 4 + k];
            cen[k] = 0.5f * (lo[k] + hi[k]);
        }
    }
    float dlt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float a = lo[k], b = hi[k];
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            a = fminf(a, __shfl_xor_sync(gmask, a, d));
            b = fmaxf(b, __shfl_xor_sync(gmask, b, d));
        }
        dlt[k] = cen[k] - 0.5f * (a + b);
    }
    // ---- slot assignment: slot s (bit k set = towards +axis k) is visited first by rays whose direction is
    //      negative along exactly the axes set in s.  Greedy: the (child, slot) pair with the largest projection
    //      of the child's offset on the slot's diagonal first; ties to the smaller child, then the smaller slot.
    unsigned free_slots = 0xffu;
    int my_slot = -1;
    for (int it = 0; it < nc; ++it) {
        float bc = -INFINITY;
        int bs = 0;
        if (have && my_slot < 0) {
#pragma unroll
            for (int sidx = 0; sidx < 8; ++sidx) {
                if (!((free_slots >> sidx) & 1u)) continue;
                float cost = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) cost += ((sidx >> k) & 1) ? dlt[k] : -dlt[k];
                if (cost > bc) { bc = cost; bs = sidx; }
            }
        }
        // arg max over the group: larger cost wins, ties to the smaller lane
        float wc = bc;
        int wl = (have && my_slot < 0) ? gl : 8, ws = bs;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const float oc = __shfl_xor_sync(gmask, wc, d);
            const int ol = __shfl_xor_sync(gmask, wl, d), os = __shfl_xor_sync(gmask, ws, d);
            const bool take = (ol < 8) && (wl >= 8 || oc > wc || (oc == wc && ol < wl));
            if (take) { wc = oc; wl = ol; ws = os; }
        }
        if (wl == gl) my_slot = ws;
        free_slots &= ~(1u << ws);
    }
    // ---- allocation: inner children and leaf records are numbered in slot order
    const int my_cnt = (have && !my_inner) ? count(my) : 0;
    int k_inner = 0, toff = 0, n_inner = 0, n_leaf_tris = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int sj = __shfl_sync(gmask, my_slot, gbase + j);
        const int ij = __shfl_sync(gmask, (int)my_inner, gbase + j);
        const int cj = __shfl_sync(gmask, my_cnt, gbase + j);
        if (sj < 0) continue;
        n_inner += ij; n_leaf_tris += cj;
        if (sj < my_slot) { k_inner += ij; toff += cj; }
    }
    unsigned cbase = 0, tbase = 0;
    {
        __shared__ unsigned s_alloc[4];                     // demand of the block (inner, records), then its two bases
        if (threadIdx.x == 0) { s_alloc[0] = 0u; s_alloc[1] = 0u; }
        __syncthreads();
        if (gl == 0) {
            if (n_inner) cbase = atomicAdd(&s_alloc[0], (unsigned)n_inner);
            if (n_leaf_tris) tbase = atomicAdd(&s_alloc[1], (unsigned)n_leaf_tris);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_alloc[2] = s_alloc[0] ? atomicAdd(&counters[0], s_alloc[0]) : 0u;
            s_alloc[3] = s_alloc[1] ? atomicAdd(&counters[1], s_alloc[1]) : 0u;
        }
        __syncthreads();
        cbase += s_alloc[2];
        tbase += s_alloc[3];
        __syncthreads();                                    // the next call of this block may zero the counters again
    }
    cbase = __shfl_sync(gmask, cbase, gbase);
    tbase = __shfl_sync(gmask, tbase, gbase);
    unsigned meta = 0, ibit = 0;
    if (have) {
        if (my_inner) {
            ibit = 1u << my_slot;
            meta = 0x20u | (24u + (unsigned)my_slot);
            // beyond the node capacity nothing is written: counters[0] still reports the overflow and the host retries
            // with the larger capacity (build_lbvh)
            if ((long long)cbase + k_inner < cap_nodes) {
                if (wparent) wparent[cbase + k_inner] = (int32_t)w;
                wroot[cbase + k_inner] = my;
            }
        } else {
            const long long f0 = my < n - 1 ? first[my] : my - (n - 1);
            meta = (((1u << my_cnt) - 1u) << 5) | (unsigned)toff;
            for (int k = 0; k < my_cnt; ++k) tri_face[tbase + toff + k] = (int32_t)sorted_tri[f0 + k];
        }
    }
    unsigned meta_lo = (have && my_slot < 4) ? meta << (8 * my_slot) : 0u;
    unsigned meta_hi = (have && my_slot >= 4) ? meta << (8 * (my_slot - 4)) : 0u;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        meta_lo |= __shfl_xor_sync(gmask, meta_lo, d);
        meta_hi |= __shfl_xor_sync(gmask, meta_hi, d);
        ibit |= __shfl_xor_sync(gmask, ibit, d);
    }
    if (gl == 0 && active) {
        nodes[w].w[0] = make_uint4(0u, 0u, 0u, ibit << 24);
        nodes[w].w[1] = make_uint4(cbase, tbase, meta_lo, meta_hi);
        if (A.arrived) A.arrived[w] = 0u;                   // the fit's arrival counter of this node starts at zero
    }
}

// one level per launch (the host reads the node counter back between levels): kept for A/B (DP_COLLAPSE_LAUNCHES=1)
template <int MODE>
__global__ void __launch_bounds__(256) k_collapse(CollapseArgs A, long long begin, long long end)
{
    const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long w = begin + (MODE == 1 ? gt : (gt >> 3));
    if (MODE == 1) {
        if (w < end) collapse_node<1>(A, w, begin, 0);
    } else {
        collapse_node<MODE>(A, w, begin, (int)(gt & 7), w < end);      // every thread of the block (block allocation)
    }
}

// The whole top-down collapse in ONE cooperative launch: the grid walks the levels itself, a grid-wide barrier where the
// per-level version went back to the host (one stream synchronisation + an 8-byte read-back per level: ~25 us each,
// 9-10 levels).  Small levels: eight lanes per node, selection and emission fused; levels of >= 2048 nodes: selection
// with one thread per node, barrier, emission with eight lanes per node (as the per-level launches did).
struct CollapseResult {
    long long n_nodes;
    int n_levels;              // -1: deeper than 126 levels; -2: node capacity exceeded (n_nodes = the demand)
    int pad_;
    long long level_begin[128];
    unsigned bar_count, bar_phase;   // grid barrier (zeroed with the struct before the launch)
    long long published;             // node counter as read by the LAST block to arrive at a barrier
};

// Grid-wide barrier of a cooperative launch (all blocks resident).  The last block to arrive reads the node counter --
// every allocation of the level is complete, none of the next has started -- and publishes it before releasing the
// others; the fence of the released thread invalidates its SM's L1, so that plain loads after the barrier see the
// other blocks' stores (wroot, sel).
__device__ __forceinline__ long long grid_barrier_publish(CollapseResult *res, const unsigned *counter, unsigned &phase)
{
    __shared__ long long s_pub;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        ++phase;
        const unsigned arrived = atomicAdd(&res->bar_count, 1u) + 1u;
        if (arrived == phase * gridDim.x) {
            res->published = (long long)__ldcg(counter);
            __threadfence();
            atomicExch(&res->bar_phase, phase);
        } else {
            while (*reinterpret_cast<volatile unsigned *>(&res->bar_phase) < phase) __nanosleep(32);
        }
        __threadfence();
        s_pub = *reinterpret_cast<volatile long long *>(&res->published);
    }
    __syncthreads();
    return s_pub;
}

__global__ void __launch_bounds__(256) k_collapse_all(CollapseArgs A, CollapseResult *res)
{
    const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    long long begin = 0, end = 1;
    int L = 0;
    unsigned phase = 0;
    if (gt == 0) res->level_begin[0] = 0;
    while (begin < end) {
        if (L + 1 >= 127) { if (gt == 0) { res->n_levels = -1; res->n_nodes = end; } return; }
        const long long lvl = end - begin;
        // MODE 0 / 2: whole blocks walk tiles of 32 nodes (the same trip count for every thread of a block)
        const int gib = threadIdx.x >> 3, gl = threadIdx.x & 7;
        if (lvl < 2048) {
            for (long long base = blockIdx.x * 32ll; base < lvl; base += gridDim.x * 32ll)
                collapse_node<0>(A, begin + base + gib, begin, gl, base + gib < lvl);
        } else {
            for (long long t = gt; t < lvl; t += nthreads) collapse_node<1>(A, begin + t, begin, 0);
            grid_barrier_publish(res, A.counters, phase);
            for (long long base = blockIdx.x * 32ll; base < lvl; base += gridDim.x * 32ll)
                collapse_node<2>(A, begin + base + gib, begin, gl, base + gib < lvl);
        }
        const long long next_end = grid_barrier_publish(res, A.counters, phase);
        ++L;
        if (gt == 0) res->level_begin[L] = end;
        begin = end;
        end = next_end;
        if (end > A.cap_nodes) { if (gt == 0) { res->n_levels = -2; res->n_nodes = end; } return; }
    }
    if (gt == 0) { res->n_levels = L; res->n_nodes = end; }
}

// Fit of one wide node by EIGHT lanes, one per child slot: exact node box into wlo/whi, quantised child boxes into
// dst.  The boxes of the inner children must already be in wlo/whi (read past L1: another thread wrote them); the
// records of the leaf triangles are (re)written here from the vertex array, so the refit needs no separate pass over
// the triangles.  `gmask` names the eight lanes of the group, `s` is this lane's slot.
//
// Quantisation: per axis the smallest power of two 2^e with 255 * 2^e >= extent; child planes are rounded outwards
// in float64 (all products with 2^e and 2^-e are exact) with an explicit containment check, plus 1/128 of a step.
__device__ __forceinline__ void fit_node8(long long w, int s, unsigned gmask, const WideNode *src, WideNode *dst,
                                          TriRec *tris, float *wlo, float *whi, float pad, const float *__restrict__ V,
                                          const int32_t *__restrict__ F, const int32_t *__restrict__ tri_face, uint4 *fat)
{
    const uint4 w1 = src[w].w[1];
    const unsigned imask = src[w].w[0].w >> 24;
    const unsigned meta = ((s < 4 ? w1.z : w1.w) >> (8 * (s & 3))) & 0xffu;
    const bool valid = meta != 0;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (valid) {
        if ((meta & 0x18u) == 0x18u) {
            const long long c = (long long)w1.x + __popc(imask & ((1u << s) - 1u));
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = __ldcg(&wlo[3 * c + k]); hi[k] = __ldcg(&whi[3 * c + k]); }
        } else {
            const int cnt = __popc(meta >> 5);
            const long long t0 = (long long)w1.y + (meta & 31u);
            for (int k = 0; k < cnt; ++k) {
                const int32_t f = tri_face[t0 + k];
                const long long a = F[3ll * f], b = F[3ll * f + 1], c = F[3ll * f + 2];
                TriRec t;
                t.v0 = make_float4(V[3 * a], V[3 * a + 1], V[3 * a + 2], __int_as_float(f));
                t.v1 = make_float4(V[3 * b], V[3 * b + 1], V[3 * b + 2], 0.0f);
                t.v2 = make_float4(V[3 * c], V[3 * c + 1], V[3 * c + 2], 0.0f);
                tris[t0 + k] = t;
                lo[0] = fminf(lo[0], fminf(t.v0.x, fminf(t.v1.x, t.v2.x)));
                lo[1] = fminf(lo[1], fminf(t.v0.y, fminf(t.v1.y, t.v2.y)));
                lo[2] = fminf(lo[2], fminf(t.v0.z, fminf(t.v1.z, t.v2.z)));
                hi[0] = fmaxf(hi[0], fmaxf(t.v0.x, fmaxf(t.v1.x, t.v2.x)));
                hi[1] = fmaxf(hi[1], fmaxf(t.v0.y, fmaxf(t.v1.y, t.v2.y)));
                hi[2] = fmaxf(hi[2], fmaxf(t.v0.z, fmaxf(t.v1.z, t.v2.z)));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] -= pad; hi[k] += pad; }
        }
    }
    if (fat) {
        // uncompressed twin: this slot's six planes (an empty slot keeps +inf / -inf: never hit) and the header
        float *fp = reinterpret_cast<float *>(fat + w * FAT_QUADS);
#pragma unroll
        for (int k = 0; k < 3; ++k) { fp[4 + 8 * k + s] = lo[k]; fp[28 + 8 * k + s] = hi[k]; }
        unsigned vb = 0;
        if (valid) vb = ((meta & 0x18u) == 0x18u) ? (0x01000000u << s) : ((meta >> 5) << (3 * s));
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) vb |= __shfl_xor_sync(gmask, vb, d);
        if (s == 0) fat[w * FAT_QUADS] = make_uint4(w1.x, w1.y, vb, 0u);
    }
    // node box over the eight slots
    float nlo[3], nhi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float a = lo[k], b = hi[k];
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            a = fminf(a, __shfl_xor_sync(gmask, a, d));
            b = fmaxf(b, __shfl_xor_sync(gmask, b, d));
        }
        nlo[k] = a; nhi[k] = b;
    }
    if (!(nlo[0] <= nhi[0])) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { nlo[k] = 0.0f; nhi[k] = 0.0f; }     // empty node
    }
    unsigned eb[3], qa[3], qb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double ext = (double)nhi[k] - (double)nlo[k];
        int e = -126;
        if (ext > 0.0) {
            e = ((__double2hiint(ext) >> 20) & 0x7ff) - 1023 - 7;          // 2^(e+7) <= ext < 2^(e+8)
            if (e < -126) e = -126;
            if (e > 127) e = 127;
            while (e < 127 && 255.0 * __hiloint2double((e + 1023) << 20, 0) < ext) ++e;
        }
        eb[k] = (unsigned)(e + 127);
        const double sc = __hiloint2double((e + 1023) << 20, 0), isc = __hiloint2double((1023 - e) << 20, 0);
        unsigned a = 255u, b = 0u;
        if (valid) {
            const double base = (double)nlo[k];
            double x = floor(((double)lo[k] - base) * isc - 0.0078125);
            x = x < 0.0 ? 0.0 : (x > 255.0 ? 255.0 : x);
            while (x > 0.0 && base + x * sc > (double)lo[k]) x -= 1.0;
            double y = ceil(((double)hi[k] - base) * isc + 0.0078125);
            y = y < 0.0 ? 0.0 : (y > 255.0 ? 255.0 : y);
            while (y < 255.0 && base + y * sc < (double)hi[k]) y += 1.0;
            a = (unsigned)x;
            b = (unsigned)y;
        }
        // bytes of four slots -> one word (lanes 0..3 end up with slots 0..3, lanes 4..7 with slots 4..7)
        a <<= 8 * (s & 3);
        b <<= 8 * (s & 3);
        a |= __shfl_xor_sync(gmask, a, 1); b |= __shfl_xor_sync(gmask, b, 1);
        a |= __shfl_xor_sync(gmask, a, 2); b |= __shfl_xor_sync(gmask, b, 2);
        qa[k] = a; qb[k] = b;
    }
    // w2: qlo_x[0..3] qlo_x[4..7] qlo_y[0..3] qlo_y[4..7] | w3: qlo_z.. qhi_x.. | w4: qhi_y.. qhi_z..
    unsigned *o = reinterpret_cast<unsigned *>(dst + w);
    const int h = s >> 2;
    if ((s & 3) == 0) {
        o[8 + h] = qa[0];  o[10 + h] = qa[1]; o[12 + h] = qa[2];
        o[14 + h] = qb[0]; o[16 + h] = qb[1]; o[18 + h] = qb[2];
    }
    if (s == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { wlo[3 * w + k] = nlo[k]; whi[3 * w + k] = nhi[k]; }
        dst[w].w[0] = make_uint4(__float_as_uint(nlo[0]), __float_as_uint(nlo[1]), __float_as_uint(nlo[2]),
                                 eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24));
        dst[w].w[1] = w1;
    }
}

// The whole bottom-up fit in ONE launch: a group of eight lanes starts at every wide node without inner children
// and walks up; the group that completes the last inner child of a node fits that node (arrival counters are never
// reset: every pass adds exactly the number of inner children, so "last" is a multiple of it).
__global__ void __launch_bounds__(256)
k_fit_all(long long n_nodes, const WideNode *src, WideNode *dst, TriRec *tris, float *wlo, float *whi,
          const float *__restrict__ d_scale, const float *__restrict__ V, const int32_t *__restrict__ F,
          const int32_t *__restrict__ tri_face, const int32_t *__restrict__ wparent, unsigned *arrived, uint4 *fat)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long w = t >> 3;
    const int s = (int)(t & 7);
    const int lane = threadIdx.x & 31;
    const unsigned gmask = 0xffu << (lane & 24);
    if (w >= n_nodes) return;                            // whole groups leave together (n_nodes is per group)
    if (src[w].w[0].w >> 24) return;                     // has inner children: fitted by the last of them
    const float pad = 3.8146973e-6f * (*d_scale);        // 2^-18 * max |coordinate|
    for (;;) {
        fit_node8(w, s, gmask, src, dst, tris, wlo, whi, pad, V, F, tri_face, fat);
        if (w == 0) return;
        const long long p = wparent[w];
        const unsigned need = __popc(src[p].w[0].w >> 24);
        unsigned last = 0;
        __threadfence();
        __syncwarp(gmask);
        if (s == 0) {
            const unsigned old = atomicAdd(&arrived[p], 1u);
            last = ((old + 1u) % need) == 0u;
        }
        last = __shfl_sync(gmask, last, lane & 24);
        if (!last) return;
        __threadfence();
        w = p;
    }
}

struct Mat34 {
    double m[12];
};

// v' = float32(((T0*x + T1*y) + T2*z) + T3) in float64: TriangleMesh.transform (:550) then the
// float32 cast of from_legacy (:245).  Also reduces max |v'| into scale_bits.
template <typename TV>
__global__ void k_pose_vertices(const TV *V, long long nV, Mat34 T, int identity, float *out, double *out64,
                                unsigned *scale_bits)
{
    // `identity`: no arithmetic at all (the vertices as they are, even non-finite ones); V may then be `out` itself
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (i < nV) {
        const double x = V[3 * i], y = V[3 * i + 1], z = V[3 * i + 2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double r = identity ? (k == 0 ? x : (k == 1 ? y : z))
                                      : __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[4 * k], x), __dmul_rn(T.m[4 * k + 1], y)),
                                                            __dmul_rn(T.m[4 * k + 2], z)),
                                                  T.m[4 * k + 3]);
            const float f = (float)r;
            out[3 * i + k] = f;
            if (out64) out64[3 * i + k] = r;
            m = fmaxf(m, fabsf(f));
        }
    }
    const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(scale_bits, mx);
}

// face indices of a device-resident mesh: any index outside [0, nV) raises the flag (dp_set_mesh checks host meshes itself)
__global__ void k_check_faces(const int32_t *__restrict__ F, long long n3, int32_t nV, int *flag)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n3 && (F[i] < 0 || F[i] >= nV)) *flag = 1;
}

__global__ void k_f64_to_f32(const double *__restrict__ a, float *__restrict__ b, long long n)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i];
}

inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

struct Bump {
    char *p;
    size_t off = 0;
    template <typename T>
    T *take(size_t count)
    {
        off = (off + 255) & ~size_t(255);
        T *r = reinterpret_cast<T *>(p + off);
        off += count * sizeof(T);
        return r;
    }
};

cudaError_t ensure_scratch(void **scratch, size_t *have, size_t need)
{
    if (*have >= need) return cudaSuccess;
    if (*scratch) cudaFree(*scratch);
    *scratch = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(scratch, need);
    if (e == cudaSuccess) *have = need;
    return e;
}

}  // namespace

size_t radix_table_entries(int64_t n)
{
    const int64_t nb = (n + RS_TILE - 1) / RS_TILE;
    return (size_t)((size_t)RS_BINS * (nb > 0 ? nb : 1) + RS_BINS);      // counts per (digit, tile) + the digit totals
}

// key_bits: the keys are below 2^key_bits (32 = anything).  An odd number of passes leaves the result in the _tmp
// arrays: *result_in_tmp tells.
cudaError_t radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, int64_t n,
                             uint32_t *table, cudaStream_t s, bool *result_in_tmp, int key_bits)
{
    *result_in_tmp = false;
    if (n <= 1) return cudaSuccess;
    if (key_bits < 1 || key_bits > 32) key_bits = 32;
    const int passes = (key_bits + RS_BITS - 1) / RS_BITS;
    const int nb = (int)((n + RS_TILE - 1) / RS_TILE);
    uint32_t *ki = keys, *vi = vals, *ko = keys_tmp, *vo = vals_tmp;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = RS_BITS * pass;
        k_rs_hist<<<nb, RS_THREADS, 0, s>>>(ki, n, shift, table, nb);
        k_rs_scan_rows<<<RS_BINS, 256, 0, s>>>(table, nb, table + (long long)RS_BINS * nb);
        k_rs_scatter<<<nb, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, table, nb, table + (long long)RS_BINS * nb);
        uint32_t *t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    *result_in_tmp = (passes & 1) != 0;
    return cudaGetLastError();
}

// DP_COLLAPSE=0 selects the greedy largest-area expansion everywhere (kept for A/B measurements); DP_CPRIM overrides
// the triangle/node cost ratio of the surface-area model; DP_HYBRID_COUNT is the largest subtree (in triangles) whose
// cut is taken from the cost tables (0 = every subtree) -- above it the greedy expansion keeps the top of the tree
// balanced and shallow, which is what the lock-step packets pay for.
// DP_COLLAPSE_LAUNCHES=1: one launch + host read-back per level instead of the single cooperative launch (read per
// build: tests flip it)
static int knob_collapse_launches()
{
    const char *e = getenv("DP_COLLAPSE_LAUNCHES");
    return e ? atoi(e) : 0;
}
static int knob_sah_collapse()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DP_COLLAPSE"); v = e ? atoi(e) : 1; }
    return v;
}
static int knob_dp_max_count()
{
    static int v = -1;
    // measured on B200 (1024^2 dense rays): 500k triangles 0.334 -> 0.310 ms, 5M triangles 16.4 -> 14.8 nodes per ray
    if (v < 0) { const char *e = getenv("DP_HYBRID_COUNT"); v = e ? atoi(e) : 512; }
    return v;
}
// Tree rotations in k_binfit: subtrees of DP_ROTATE_MIN .. DP_ROTATE_MAX triangles may rotate (MAX = 0: none).
// Measured on B200 (1024^2 dense rays; profiles/r1d_sweep_rotations*.log): the gain sits at the top of the tree --
// rotating only subtrees of >= 1024 triangles gives 500k triangles 12.79 -> 12.63 nodes per ray (0.297 -> 0.293 ms)
// and 5M triangles 14.80 -> 14.57 (0.439 -> 0.421 ms) at no build cost; rotating everything adds 0.1 / 0.7 ms to the
// build for another 0.1 nodes per ray.  Grandchild <-> grandchild candidates (DP_ROTATE_GG) lower the surface area
// further but not the node visits (12.73 / 14.97), further passes (DP_ROTATE_PASSES) cost a k_binfit each for
// 0.1 nodes per ray: both stay off.
static int knob_rotate_max()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DP_ROTATE_MAX"); v = e ? atoi(e) : 0x7fffffff; }
    return v;
}
static int knob_rotate_min()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DP_ROTATE_MIN"); v = e ? atoi(e) : 1024; }
    return v;
}
static int knob_rotate_gg()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DP_ROTATE_GG"); v = e ? atoi(e) : 0; }
    return v;
}
static int knob_rotate_passes()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DP_ROTATE_PASSES"); v = e ? atoi(e) : 1; }
    return v;
}
static float knob_c_prim()
{
    static float v = -1.0f;
    if (v < 0.0f) { const char *e = getenv("DP_CPRIM"); v = e ? (float)atof(e) : 0.5f; }
    return v;
}

cudaError_t build_lbvh(const float *V, int64_t nV, const int32_t *F, int64_t nF, BvhStorage &out, Topology &topo,
                       void **scratch, size_t *scratch_bytes, uint32_t *morton_out_host, cudaStream_t s)
{
    cudaError_t e;
    const long long n = nF;
    const size_t N = (size_t)(n > 0 ? n : 1);
    size_t need = 4096 + sizeof(CollapseResult) + 256 + 4 * (N * 4 + 256) + (radix_table_entries(n) * 4 + 256) + 4 * (N * 4 + 256) +
                  (2 * N * 4 + 256) + (2 * N * 8 * 4 + 256) + (N * 4 + 256) + (N * 4 + 256) + (8 * N * 4 + 256) + (N * 32 + 256) + (9 * (N / 2 + 64) * 4 + 256) + 1024;
    if ((e = ensure_scratch(scratch, scratch_bytes, need)) != cudaSuccess) return e;
    Bump b{static_cast<char *>(*scratch)};
    unsigned *bounds_u = b.take<unsigned>(8);
    float *sbounds = b.take<float>(8);
    unsigned *counters = b.take<unsigned>(4);
    CollapseResult *cres = b.take<CollapseResult>(1);
    uint32_t *keys = b.take<uint32_t>(N), *vals = b.take<uint32_t>(N);
    uint32_t *keys_t = b.take<uint32_t>(N), *vals_t = b.take<uint32_t>(N);
    uint32_t *table = b.take<uint32_t>(radix_table_entries(n));
    int32_t *left = b.take<int32_t>(N), *right = b.take<int32_t>(N);
    int32_t *first = b.take<int32_t>(N), *last = b.take<int32_t>(N);
    int32_t *parent = b.take<int32_t>(2 * N);
    float *bbox = b.take<float>(2 * N * 8);
    int *flags = b.take<int>(N);
    int32_t *wroot = b.take<int32_t>(N);
    int32_t *selbuf = b.take<int32_t>(9 * (N / 2 + 64));       // selection of one level (two-phase collapse)
    float *ctab = knob_sah_collapse() == 1 ? b.take<float>(8 * N) : nullptr;
    BinRec *rec = b.take<BinRec>(N);
    const float c_prim = knob_c_prim();

    out.n_tris = n;
    topo.n_levels = 0;
    topo.level_begin[0] = 0;

    // scene bounds and scale
    {
        const unsigned init[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
        if ((e = cudaMemcpyAsync(bounds_u, init, sizeof(init), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
        if (n > 0) k_scene_bounds<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, s>>>(V, F, n, bounds_u);
        k_finish_bounds<<<1, 1, 0, s>>>(bounds_u, sbounds, out.d_scale, n);
    }
    if (n == 0) {
        // a root without children: every ray misses
        WideNode root;
        root.w[0] = make_uint4(0, 0, 0, 127u | (127u << 8) | (127u << 16));
        root.w[1] = make_uint4(0, 0, 0, 0);
        root.w[2] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        root.w[3] = make_uint4(0xffffffffu, 0xffffffffu, 0, 0);
        root.w[4] = make_uint4(0, 0, 0, 0);
        if ((e = cudaMemcpyAsync(out.nodes, &root, sizeof(root), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
        if (out.fat && (e = cudaMemsetAsync(out.fat, 0, FAT_NODE_BYTES, s)) != cudaSuccess) return e;   // valid = 0
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
        out.n_nodes = 1;
        topo.n_levels = 1;
        topo.level_begin[1] = 1;
        return cudaSuccess;
    }
    k_morton<<<blocks_for(n, 256), 256, 0, s>>>(V, F, n, sbounds, keys, vals, parent, flags);
    if (morton_out_host) {
        if ((e = cudaMemcpyAsync(morton_out_host, keys, n * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    }
    bool in_tmp = false;
    if ((e = radix_sort_pairs(keys, vals, keys_t, vals_t, n, table, s, &in_tmp, 30)) != cudaSuccess) return e;   // 30-bit Morton codes
    if (in_tmp) { uint32_t *t = keys; keys = keys_t; keys_t = t; t = vals; vals = vals_t; vals_t = t; }

    if (n > 1) k_karras<<<blocks_for(n - 1, 256), 256, 0, s>>>(keys, n, left, right, parent, first, last);
    static_assert(sizeof(CollapseResult) % 4 == 0, "the result block is cleared word by word");
    const int n_cres_words = (int)(sizeof(CollapseResult) / 4);
    k_binfit<<<blocks_for(n > n_cres_words ? n : n_cres_words, 256), 256, 0, s>>>(
        V, F, vals, n, parent, left, right, first, last, bbox, flags, ctab, c_prim, knob_rotate_min(), knob_rotate_max(),
        knob_rotate_gg(), rec, counters, wroot, reinterpret_cast<unsigned *>(cres), n_cres_words);
    // further rotation passes over the rotated tree (each node looks at its new grandchildren once more)
    for (int pass = 1; pass < knob_rotate_passes() && knob_rotate_max() > 0; ++pass) {
        if ((e = cudaMemsetAsync(flags, 0, N * 4, s)) != cudaSuccess) return e;
        k_binfit<<<blocks_for(n, 256), 256, 0, s>>>(V, F, vals, n, parent, left, right, first, last, bbox, flags, ctab,
                                                    c_prim, knob_rotate_min(), knob_rotate_max(), knob_rotate_gg(), rec,
                                                    nullptr, nullptr, nullptr, 0);
    }

    // top-down collapse, one launch per level of the wide tree
    // (counters = {1, 0, 0, 0}: node 0 is the root; wroot[0] = 0; the result block zeroed: all written by k_binfit)
    CollapseArgs ca{n, left, right, first, last, bbox, vals, wroot, out.nodes, topo.tri_face, counters, ctab, c_prim,
                    knob_sah_collapse(), topo.wparent, knob_dp_max_count(), selbuf, out.cap_nodes, rec, topo.arrived};
    long long begin = 0, end = 1;
    int L = 0;
    if (knob_collapse_launches() == 0) {
        // one cooperative launch walks all levels (grid barriers instead of host round trips), one read-back at the end
        static int coop_sms = 0, coop_per_sm = 0, coop_want = 0;
        if (coop_sms == 0) {
            int dev = 0, sms = 0, per_sm = 0;
            if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
            if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_collapse_all, 256, 0)) != cudaSuccess) return e;
            const char *cp = getenv("DP_COOP_PER_SM");             // A/B: blocks per SM of the cooperative launch
            coop_want = cp ? atoi(cp) : 0;
            coop_per_sm = per_sm > 0 ? per_sm : 1;
            coop_sms = sms;
        }
        // Every level ends in one or two grid barriers, whose cost grows with the number of blocks: up to ~2M triangles
        // the levels are short and three blocks per SM are enough (500k: 0.645 -> 0.617 ms, 1M: 0.893 -> 0.863 ms); the
        // long levels of larger meshes want every resident block (5M: 2.70 ms with five, 2.74 with three)
        int per_sm_now = coop_want > 0 ? coop_want : (n <= 2000000 ? 3 : coop_per_sm);
        if (per_sm_now > coop_per_sm) per_sm_now = coop_per_sm;
        const int coop_grid = coop_sms * per_sm_now;
        void *kargs[] = {&ca, &cres};
        if ((e = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_collapse_all), dim3(coop_grid), dim3(256), kargs, 0, s)) !=
            cudaSuccess)
            return e;
        CollapseResult hres;
        if ((e = cudaMemcpyAsync(&hres, cres, sizeof(CollapseResult), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
        if (hres.n_levels == -1) { topo.n_levels = -1; return cudaSuccess; }      // deeper than any ray stack
        if (hres.n_levels == -2) return cudaErrorMemoryAllocation;               // the caller retries with more nodes
        L = hres.n_levels;
        end = hres.n_nodes;
        for (int l = 0; l <= L; ++l) topo.level_begin[l] = hres.level_begin[l];
        topo.level_begin[L] = end;
    } else {
        while (begin < end) {
            if (L + 1 >= 127) { topo.n_levels = -1; return cudaSuccess; }      // deeper than any ray stack: reported by the caller
            const long long lvl = end - begin;
            if (lvl < 2048 || lvl > (long long)(N / 2 + 64)) {
                k_collapse<0><<<blocks_for(lvl * 8, 256), 256, 0, s>>>(ca, begin, end);
            } else {
                k_collapse<1><<<blocks_for(lvl, 128), 128, 0, s>>>(ca, begin, end);
                k_collapse<2><<<blocks_for(lvl * 8, 256), 256, 0, s>>>(ca, begin, end);
            }
            unsigned cnt[2];
            if ((e = cudaMemcpyAsync(cnt, counters, 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
            topo.level_begin[++L] = end;
            begin = end;
            end = cnt[0];
            if (end > out.cap_nodes) return cudaErrorMemoryAllocation;
        }
    }
    topo.n_levels = L;
    out.n_nodes = end;
    k_fit_all<<<blocks_for(out.n_nodes * 8, 256), 256, 0, s>>>(out.n_nodes, out.nodes, out.nodes, out.tris, out.wlo, out.whi,
                                                           out.d_scale, V, F, topo.tri_face, topo.wparent, topo.arrived, out.fat);
    return cudaGetLastError();
}

cudaError_t check_faces(const int32_t *F, int64_t nF, int64_t nV, int *d_flag, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(d_flag, 0, sizeof(int), s);
    if (e != cudaSuccess || nF <= 0) return e;
    k_check_faces<<<blocks_for(3 * nF, 256), 256, 0, s>>>(F, 3 * nF, (int32_t)nV, d_flag);
    return cudaGetLastError();
}

cudaError_t convert_f64_to_f32(const double *src, float *dst, int64_t n, cudaStream_t s)
{
    if (n > 0) k_f64_to_f32<<<blocks_for(n, 256), 256, 0, s>>>(src, dst, n);
    return cudaGetLastError();
}

cudaError_t pose_and_refit(const void *V, int vdtype, int64_t nV, const int32_t *F, int64_t nF, const double *T_host,
                           float *Vposed, double *Vposed64, const BvhStorage &src, BvhStorage &dst,
                           const Topology &topo, cudaStream_t s)
{
    cudaError_t e;
    // T_host == nullptr: the vertices as they are (in-place refit of `src` after dp_update_vertices)
    Mat34 T;
    for (int k = 0; k < 12; ++k) T.m[k] = T_host ? T_host[k] : (k % 5 == 0 ? 1.0 : 0.0);
    const int identity = T_host == nullptr;
    if ((e = cudaMemsetAsync(dst.d_scale, 0, 4, s)) != cudaSuccess) return e;
    if (nV > 0) {
        if (vdtype == 1)
            k_pose_vertices<double><<<blocks_for(nV, 256), 256, 0, s>>>(static_cast<const double *>(V), nV, T, identity, Vposed,
                                                                        Vposed64, reinterpret_cast<unsigned *>(dst.d_scale));
        else
            k_pose_vertices<float><<<blocks_for(nV, 256), 256, 0, s>>>(static_cast<const float *>(V), nV, T, identity, Vposed,
                                                                       Vposed64, reinterpret_cast<unsigned *>(dst.d_scale));
    }
    dst.n_nodes = src.n_nodes;
    dst.n_tris = src.n_tris;
    if (nF == 0) {
        return cudaMemcpyAsync(dst.nodes, src.nodes, sizeof(WideNode), cudaMemcpyDeviceToDevice, s);
    }
    k_fit_all<<<blocks_for(src.n_nodes * 8, 256), 256, 0, s>>>(src.n_nodes, src.nodes, dst.nodes, dst.tris, dst.wlo, dst.whi,
                                                           dst.d_scale, Vposed, F, topo.tri_face, topo.wparent, topo.arrived, dst.fat);
    return cudaGetLastError();
}

}  // namespace dp
