// Point-to-plane ICP inner loop: the producer of the pose the back-projection consumes.
//
// Replaces o3d.pipelines.registration.registration_icp(source, target, max_correspondence_distance, init,
// TransformationEstimationPointToPlane()[, ICPConvergenceCriteria]) as called by
//   /root/reference/src/pose_estimation.py:505-522 (refine_registration), :577-613 (improve_result, <= 50 restarts),
//   :654-660 (predict_z_axis_adjustment, max_iteration = 1).
// Open3D is a third-party dependency (open3d==0.18.0, absent offline): PARITY UNPINNED.  What is restated is its
// published algorithm (cpp/open3d/pipelines/registration/Registration.cpp, TransformationEstimation.cpp, as recalled):
//   pcd = source transformed by init
//   repeat: correspondences = nearest target point of every pcd point within max_correspondence_distance
//           fitness = |corr| / |source|, inlier_rmse = sqrt(sum d^2 / |corr|)
//           r = (s - t) . n_t,  J = [s x n_t, n_t];  solve (sum J J^T) x = -(sum J r)   (6 x 6)
//           update = [Rz(x2) Ry(x1) Rx(x0) | x3..5];  T = update * T;  pcd = update * pcd
//   until |d fitness| < relative_fitness and |d rmse| < relative_rmse, or max_iteration.
//
// One kernel per iteration does the cloud update, the exact nearest-neighbour search (float64 brute force over
// shared-memory tiles of the target, ties to the smaller index -- a k-d tree's tie choice is arbitrary) and the
// block-level reduction of the 29 sums; a second, single-block kernel adds the block partials in a fixed order, so
// results are reproducible run to run.  The 6 x 6 solve and the convergence test run on the host between
// iterations (29 doubles come back per iteration).
#include "dp_internal.cuh"

#include <string.h>

namespace dp {

namespace {

constexpr int ICP_THREADS = 128;
constexpr int ICP_TILE = 1024;        // target points per shared-memory tile
constexpr int ICP_NSUM = 29;          // count, sum d^2, 21 upper-triangle entries of J J^T, 6 entries of J r

struct Xf16 { double m[16]; };

// PointCloud::Transform: (U * [p, 1]).head<3>() / w, evaluated left to right
__device__ __forceinline__ void icp_load_point(double *__restrict__ src, long long i, int apply_update, const Xf16 &U, double &qx,
                                               double &qy, double &qz)
{
    qx = src[3 * i]; qy = src[3 * i + 1]; qz = src[3 * i + 2];
    if (apply_update) {
        double h[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            h[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(U.m[4 * r], qx), __dmul_rn(U.m[4 * r + 1], qy)),
                                       __dmul_rn(U.m[4 * r + 2], qz)), U.m[4 * r + 3]);
        qx = __ddiv_rn(h[0], h[3]); qy = __ddiv_rn(h[1], h[3]); qz = __ddiv_rn(h[2], h[3]);
        src[3 * i] = qx; src[3 * i + 1] = qy; src[3 * i + 2] = qz;
    }
}

// The 29 terms of one correspondence (every operation individually rounded, so that the brute-force and the grid
// search give the same sums bit for bit) and the block's fixed-shape reduction: lanes by xor shuffles, then the
// four warps in order.
__device__ __forceinline__ void icp_emit(bool has, double best, double qx, double qy, double qz, const double *__restrict__ t3,
                                         const double *__restrict__ n3, double (*s_red)[ICP_NSUM], double *__restrict__ partial)
{
    double v[ICP_NSUM];
#pragma unroll
    for (int k = 0; k < ICP_NSUM; ++k) v[k] = 0.0;
    if (has) {
        const double tx = t3[0], ty = t3[1], tz = t3[2];
        const double nx = n3[0], ny = n3[1], nz = n3[2];
        const double r = __dadd_rn(__dadd_rn(__dmul_rn(__dsub_rn(qx, tx), nx), __dmul_rn(__dsub_rn(qy, ty), ny)),
                                   __dmul_rn(__dsub_rn(qz, tz), nz));
        double J[6];
        J[0] = __dsub_rn(__dmul_rn(qy, nz), __dmul_rn(qz, ny));
        J[1] = __dsub_rn(__dmul_rn(qz, nx), __dmul_rn(qx, nz));
        J[2] = __dsub_rn(__dmul_rn(qx, ny), __dmul_rn(qy, nx));
        J[3] = nx; J[4] = ny; J[5] = nz;
        v[0] = 1.0;
        v[1] = best;
        int e = 2;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) v[e++] = __dmul_rn(J[a], J[b]);
#pragma unroll
        for (int a = 0; a < 6; ++a) v[23 + a] = __dmul_rn(J[a], r);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ICP_NSUM; ++k) {
        double x = v[k];
#pragma unroll
        for (int d = 16; d; d >>= 1) x = __dadd_rn(x, __shfl_xor_sync(0xffffffffu, x, d));
        if (lane == 0) s_red[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < ICP_NSUM) {
        double x = s_red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < ICP_THREADS / 32; ++w) x = __dadd_rn(x, s_red[w][threadIdx.x]);
        partial[(long long)blockIdx.x * ICP_NSUM + threadIdx.x] = x;
    }
}

__global__ void __launch_bounds__(ICP_THREADS)
k_icp_step(double *__restrict__ src, long long n, const double *__restrict__ tp, const double *__restrict__ tn, long long m,
           double max_d2, int apply_update, Xf16 U, int32_t *__restrict__ corr, double *__restrict__ partial)
{
    __shared__ double s_t[ICP_TILE * 3];
    __shared__ double s_red[ICP_THREADS / 32][ICP_NSUM];
    const long long i = blockIdx.x * (long long)ICP_THREADS + threadIdx.x;
    const bool ok = i < n;
    double qx = 0.0, qy = 0.0, qz = 0.0;
    if (ok) icp_load_point(src, i, apply_update, U, qx, qy, qz);
    double best = INFINITY;
    long long bi = -1;
    for (long long t0 = 0; t0 < m; t0 += ICP_TILE) {
        const int cnt = (int)((m - t0) < ICP_TILE ? (m - t0) : ICP_TILE);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 3; k += ICP_THREADS) s_t[k] = tp[t0 * 3 + k];
        __syncthreads();
        if (ok) {
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
                const double dx = __dsub_rn(qx, s_t[3 * k]), dy = __dsub_rn(qy, s_t[3 * k + 1]), dz = __dsub_rn(qz, s_t[3 * k + 2]);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (d2 < best) { best = d2; bi = t0 + k; }
            }
        }
    }
    const bool has = ok && bi >= 0 && best <= max_d2;
    if (ok && corr) corr[i] = has ? (int32_t)bi : -1;
    icp_emit(has, best, qx, qy, qz, tp + 3 * (has ? bi : 0), tn + 3 * (has ? bi : 0), s_red, partial);
}


// ------------------------------------------------------------------------------------------ uniform grid over the target
// registration_icp only wants neighbours within max_correspondence_distance, so the exact nearest neighbour can be
// found in the 27 cells around the query of a grid whose cells are at least that wide (built once per call: cell
// keys, the in-house radix sort, cell ranges, target points / normals gathered in cell order).  Same distance
// arithmetic as the brute-force scan, ties to the smaller ORIGINAL index, so both searches return the same index.
__global__ void __launch_bounds__(256) k_icp_bounds(const double *__restrict__ tp, long long m, unsigned long long *bounds)
{
    // bounds[0..2] = ordered min, [3..5] = ordered max
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double v = tp[3 * j + k]; lo[k] = fmin(lo[k], v); hi[k] = fmax(hi[k], v); }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int d = 16; d; d >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
        }
        if ((threadIdx.x & 31) == 0) {
            long long a = __double_as_longlong(lo[k]), b = __double_as_longlong(hi[k]);
            a = a < 0 ? (a ^ 0x7fffffffffffffffll) : a;
            b = b < 0 ? (b ^ 0x7fffffffffffffffll) : b;
            atomicMin(reinterpret_cast<long long *>(&bounds[k]), a);
            atomicMax(reinterpret_cast<long long *>(&bounds[3 + k]), b);
        }
    }
}

__global__ void __launch_bounds__(256)
k_icp_grid_keys(const double *__restrict__ tp, long long m, IcpGrid g, uint32_t *keys, uint32_t *vals)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int cx = grid_coord(tp[3 * j], g.lo[0], g.inv_cell, g.dim[0]);
    const int cy = grid_coord(tp[3 * j + 1], g.lo[1], g.inv_cell, g.dim[1]);
    const int cz = grid_coord(tp[3 * j + 2], g.lo[2], g.inv_cell, g.dim[2]);
    keys[j] = (uint32_t)((cz * g.dim[1] + cy) * g.dim[0] + cx);
    vals[j] = (uint32_t)j;
}

// sorted order: points and normals gathered, cell ranges [start, end)
__global__ void __launch_bounds__(256)
k_icp_grid_gather(const double *__restrict__ tp, const double *__restrict__ tn, const uint32_t *__restrict__ keys,
                  const uint32_t *__restrict__ vals, long long m, double *tps, double *tns, int32_t *cell_start,
                  int32_t *cell_end)
{
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= m) return;
    const uint32_t j = vals[k], key = keys[k];
#pragma unroll
    for (int a = 0; a < 3; ++a) { tps[3 * k + a] = tp[3ll * j + a]; if (tn) tns[3 * k + a] = tn[3ll * j + a]; }
    if (k == 0 || keys[k - 1] != key) cell_start[key] = (int32_t)k;
    if (k == m - 1 || keys[k + 1] != key) cell_end[key] = (int32_t)k + 1;
}

__global__ void __launch_bounds__(ICP_THREADS)
k_icp_step_grid(double *__restrict__ src, long long n, const double *__restrict__ tps, const double *__restrict__ tns,
                const uint32_t *__restrict__ orig, const int32_t *__restrict__ cell_start, const int32_t *__restrict__ cell_end,
                IcpGrid g, double max_d2, int apply_update, Xf16 U, int32_t *__restrict__ corr, double *__restrict__ partial)
{
    __shared__ double s_red[ICP_THREADS / 32][ICP_NSUM];
    const long long i = blockIdx.x * (long long)ICP_THREADS + threadIdx.x;
    const bool ok = i < n;
    double qx = 0.0, qy = 0.0, qz = 0.0;
    if (ok) icp_load_point(src, i, apply_update, U, qx, qy, qz);
    double best = INFINITY;
    long long bk = -1;          // position in cell order
    uint32_t bo = 0xffffffffu;  // original index (tie-break)
    if (ok && qx == qx && qy == qy && qz == qz) {
        const int cx = grid_coord(qx, g.lo[0], g.inv_cell, g.dim[0]);
        const int cy = grid_coord(qy, g.lo[1], g.inv_cell, g.dim[1]);
        const int cz = grid_coord(qz, g.lo[2], g.inv_cell, g.dim[2]);
        for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dim[2] - 1); ++z)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y)
                for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x) {
                    const int cell = (z * g.dim[1] + y) * g.dim[0] + x;
                    const int k0 = cell_start[cell], k1 = cell_end[cell];
                    for (int k = k0; k < k1; ++k) {
                        const double dx = __dsub_rn(qx, tps[3 * k]), dy = __dsub_rn(qy, tps[3 * k + 1]), dz = __dsub_rn(qz, tps[3 * k + 2]);
                        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        const uint32_t o = orig[k];
                        if (d2 < best || (d2 == best && o < bo)) { best = d2; bk = k; bo = o; }
                    }
                }
    }
    const bool has = ok && bk >= 0 && best <= max_d2;
    if (ok && corr) corr[i] = has ? (int32_t)bo : -1;
    icp_emit(has, best, qx, qy, qz, tps + 3 * (has ? bk : 0), tns + 3 * (has ? bk : 0), s_red, partial);
}

// sums[k] = sum over blocks of partial[b][k], blocks taken in index order by a fixed tree
__global__ void __launch_bounds__(256) k_icp_reduce(const double *__restrict__ partial, long long nblocks, double *__restrict__ sums)
{
    __shared__ double s[256];
    for (int k = 0; k < ICP_NSUM; ++k) {
        double x = 0.0;
        for (long long b = threadIdx.x; b < nblocks; b += 256) x += partial[b * ICP_NSUM + k];
        s[threadIdx.x] = x;
        __syncthreads();
        for (int d = 128; d; d >>= 1) {
            if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
            __syncthreads();
        }
        if (threadIdx.x == 0) sums[k] = s[0];
        __syncthreads();
    }
}

}  // namespace

size_t icp_partial_doubles(int64_t n) { return (size_t)((n + ICP_THREADS - 1) / ICP_THREADS) * ICP_NSUM + ICP_NSUM; }

// one ICP iteration on the device: optional in-place update of the working cloud, correspondences, 29 sums
cudaError_t launch_icp_step(double *src, int64_t n, const double *tp, const double *tn, int64_t m, double max_dist,
                            const double *update_host, int32_t *corr, double *partial, double *sums, cudaStream_t s)
{
    if (n <= 0) return cudaMemsetAsync(sums, 0, ICP_NSUM * sizeof(double), s);
    Xf16 U;
    for (int i = 0; i < 16; ++i) U.m[i] = update_host ? update_host[i] : (i % 5 == 0 ? 1.0 : 0.0);
    const long long nb = (n + ICP_THREADS - 1) / ICP_THREADS;
    k_icp_step<<<(unsigned)nb, ICP_THREADS, 0, s>>>(src, n, tp, tn, m, max_dist * max_dist, update_host != nullptr, U, corr, partial);
    k_icp_reduce<<<1, 256, 0, s>>>(partial, nb, sums);
    return cudaGetLastError();
}


// ---- grid: sizes, build, step ----------------------------------------------------------------------------
size_t icp_grid_bytes(int64_t m)
{
    const size_t M = (size_t)(m > 0 ? m : 1);
    return 256 + 4 * (M * 4 + 256) + (radix_table_entries(m) * 4 + 256) + 2 * (M * 24 + 256) +
           2 * ((size_t)ICP_GRID_MAX_DIM * ICP_GRID_MAX_DIM * ICP_GRID_MAX_DIM * 4 + 256);
}

// Builds the grid over the target in `scratch` (icp_grid_bytes(m) bytes).  One 48-byte read-back (the bounding box).
// `tn` may be null (points only); out->tps stays null when the grid is not worth building (fewer than `min_cells`
// cells) or cannot be built (non-finite coordinates).
cudaError_t icp_grid_build(const double *tp, const double *tn, int64_t m, double max_dist, size_t min_cells, void *scratch,
                           IcpGridView *out, cudaStream_t s)
{
    cudaError_t e;
    char *p = static_cast<char *>(scratch);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~size_t(255); char *r = p + off; off += bytes; return r; };
    unsigned long long *bounds = reinterpret_cast<unsigned long long *>(take(64));
    uint32_t *keys = reinterpret_cast<uint32_t *>(take((size_t)m * 4)), *vals = reinterpret_cast<uint32_t *>(take((size_t)m * 4));
    uint32_t *keys_t = reinterpret_cast<uint32_t *>(take((size_t)m * 4)), *vals_t = reinterpret_cast<uint32_t *>(take((size_t)m * 4));
    uint32_t *table = reinterpret_cast<uint32_t *>(take(radix_table_entries(m) * 4));
    double *tps = reinterpret_cast<double *>(take((size_t)m * 24)), *tns = reinterpret_cast<double *>(take((size_t)m * 24));
    const size_t cells_max = (size_t)ICP_GRID_MAX_DIM * ICP_GRID_MAX_DIM * ICP_GRID_MAX_DIM;
    int32_t *cell_start = reinterpret_cast<int32_t *>(take(cells_max * 4)), *cell_end = reinterpret_cast<int32_t *>(take(cells_max * 4));

    const long long init[6] = {0x7ff0000000000000ll, 0x7ff0000000000000ll, 0x7ff0000000000000ll,
                               (long long)0xfff0000000000000ull ^ 0x7fffffffffffffffll,
                               (long long)0xfff0000000000000ull ^ 0x7fffffffffffffffll,
                               (long long)0xfff0000000000000ull ^ 0x7fffffffffffffffll};
    if ((e = cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    long long nb = (m + 256 * 8 - 1) / (256 * 8);
    if (nb > 148 * 8) nb = 148 * 8;
    k_icp_bounds<<<(unsigned)nb, 256, 0, s>>>(tp, m, bounds);
    long long hb[6];
    if ((e = cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    IcpGrid g;
    double ext = 0.0;
    for (int k = 0; k < 3; ++k) {
        long long a = hb[k], b = hb[3 + k];
        a = a < 0 ? (a ^ 0x7fffffffffffffffll) : a;
        b = b < 0 ? (b ^ 0x7fffffffffffffffll) : b;
        double lo, hi;
        memcpy(&lo, &a, 8); memcpy(&hi, &b, 8);
        if (!(lo <= hi)) { out->tps = nullptr; return cudaSuccess; }   // no finite target coordinate
        g.lo[k] = lo;
        out->ext[k] = hi - lo;
        ext = hi - lo > ext ? hi - lo : ext;
    }
    out->tps = nullptr;                                              // "no grid": the caller scans the target instead
    double cell;
    if (max_dist > 0.0) {
        cell = max_dist * (1.0 + 1e-9);                              // >= the search radius, with room for rounding
        if (ext / ICP_GRID_MAX_DIM > cell) cell = ext / ICP_GRID_MAX_DIM;
    } else {
        // no radius (unbounded nearest neighbour, searched ring by ring): a handful of points per occupied cell of
        // a surface-like cloud, which fills about d^2 of the d^3 cells
        double d = ceil(sqrt((double)m / 16.0));
        d = d < 1.0 ? 1.0 : (d > ICP_GRID_MAX_DIM ? (double)ICP_GRID_MAX_DIM : d);
        cell = ext > 0.0 ? ext / d : 1.0;
    }
    if (!(cell > 0.0) || !(cell < 1e300)) return cudaSuccess;        // infinite coordinates
    g.inv_cell = 1.0 / cell;
    for (int k = 0; k < 3; ++k) {
        double d = floor(out->ext[k] * g.inv_cell) + 1.0;
        g.dim[k] = d < 1.0 ? 1 : (d > ICP_GRID_MAX_DIM ? ICP_GRID_MAX_DIM : (int)d);
    }
    const size_t cells = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
    // 27 of `cells` cells per query against one tiled scan of all of them: below ~8 x 27 cells the scan is as fast
    if (cells < min_cells) return cudaSuccess;
    k_icp_grid_keys<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(tp, m, g, keys, vals);
    bool in_tmp = false;
    int key_bits = 1;
    while (key_bits < 32 && ((size_t)1 << key_bits) < cells) ++key_bits;              // cell keys are below `cells`
    if ((e = radix_sort_pairs(keys, vals, keys_t, vals_t, m, table, s, &in_tmp, key_bits)) != cudaSuccess) return e;
    if (in_tmp) { uint32_t *t = keys; keys = keys_t; keys_t = t; t = vals; vals = vals_t; vals_t = t; }
    if ((e = cudaMemsetAsync(cell_start, 0, cells * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(cell_end, 0, cells * 4, s)) != cudaSuccess) return e;
    k_icp_grid_gather<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(tp, tn, keys, vals, m, tps, tns, cell_start, cell_end);
    g.cell = cell;
    out->grid = g;
    out->tps = tps; out->tns = tns; out->orig = vals; out->cell_start = cell_start; out->cell_end = cell_end;
    return cudaGetLastError();
}

cudaError_t launch_icp_step_grid(double *src, int64_t n, const IcpGridView &gv, double max_dist, const double *update_host,
                                 int32_t *corr, double *partial, double *sums, cudaStream_t s)
{
    if (n <= 0) return cudaMemsetAsync(sums, 0, ICP_NSUM * sizeof(double), s);
    Xf16 U;
    for (int i = 0; i < 16; ++i) U.m[i] = update_host ? update_host[i] : (i % 5 == 0 ? 1.0 : 0.0);
    const long long nb = (n + ICP_THREADS - 1) / ICP_THREADS;
    k_icp_step_grid<<<(unsigned)nb, ICP_THREADS, 0, s>>>(src, n, gv.tps, gv.tns, gv.orig, gv.cell_start, gv.cell_end, gv.grid,
                                                         max_dist * max_dist, update_host != nullptr, U, corr, partial);
    k_icp_reduce<<<1, 256, 0, s>>>(partial, nb, sums);
    return cudaGetLastError();
}

}  // namespace dp
