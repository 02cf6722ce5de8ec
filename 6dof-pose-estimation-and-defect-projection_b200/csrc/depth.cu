// Depth-image projection path: the alternative to ray tracing in
// /root/reference/src/defect_projection.py
//   heatmap_to_point3d  :359-395   double Python loop over H x W: intensity = heat/max(heat); keep pixels with
//                                  intensity > thr and depth > 0; back-project with the depth value
//   align_to_surface    :417-460   per defect point: nearest target point (Open3D KDTreeFlann, k = 1), the
//                                  "aligned" point, and that point pushed along its normal by `offset`
//   calc_coordinates    :462-492   the same back-projection for a list of picked pixels
// Selection keeps the reference's row-major order (same ranks-from-ballots + decoupled look-back as
// compact.cu); all arithmetic is float64, operation by operation as the reference's Python evaluates it.
// The nearest-neighbour search is exact brute force in float64 (tiles of the target cloud staged in shared
// memory): d2 = (dx*dx + dy*dy) + dz*dz, ties to the smaller index.
#include "dp_internal.cuh"

namespace dp {

namespace {

constexpr int DS_THREADS = 256;
constexpr int DS_ITEMS = 8;
constexpr int DS_TILE = DS_THREADS * DS_ITEMS;

template <typename T>
__global__ void __launch_bounds__(256) k_max(const T *__restrict__ a, long long n, double *out_max, int *out_nan)
{
    // np.max semantics: NaN if any element is NaN
    double m = -INFINITY;
    int nan = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)a[i];
        if (v != v) nan = 1; else m = fmax(m, v);
    }
    for (int d = 16; d; d >>= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
        nan |= __shfl_xor_sync(0xffffffffu, nan, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nan) atomicOr(out_nan, 1);
        // ordered-int atomic max on the double's bits (all values finite or -inf here)
        long long b = __double_as_longlong(m);
        b = b < 0 ? (b ^ 0x7fffffffffffffffll) : b;
        atomicMax(reinterpret_cast<long long *>(out_max), b);
    }
}

__global__ void k_max_finish(double *mx, const int *nan)
{
    long long b = *reinterpret_cast<long long *>(mx);
    b = b < 0 ? (b ^ 0x7fffffffffffffffll) : b;
    *mx = *nan ? __longlong_as_double(0x7ff8000000000000ll) : __longlong_as_double(b);
}

// one pass: predicate, order-preserving ranks, back-projection of the selected pixels
template <typename T>
__global__ void __launch_bounds__(DS_THREADS)
k_depth_select(const T *__restrict__ heat, int H, int W, const uint16_t *__restrict__ depth, int Hd, int Wd,
               const double *__restrict__ d_max, double thr, double fx, double fy, double cx, double cy,
               double *__restrict__ out4, long long cap, unsigned long long *scratch, long long *d_count)
{
    __shared__ unsigned s_tile, s_warp_tot[DS_THREADS / 32], s_warp_off[DS_THREADS / 32];
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned *>(scratch), 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    unsigned long long *state = scratch + 1;
    const long long n = (long long)H * W;
    const double maxv = *d_max;
    // thread t owns DS_ITEMS consecutive pixels: row-major order = (thread, item)
    const long long e0 = (long long)tile * DS_TILE + (long long)tid * DS_ITEMS;
    unsigned m = 0;
    double inten[DS_ITEMS];
    unsigned short dep[DS_ITEMS];
#pragma unroll
    for (int j = 0; j < DS_ITEMS; ++j) {
        const long long e = e0 + j;
        inten[j] = 0.0; dep[j] = 0;
        if (e < n) {
            const int y = (int)(e / W), x = (int)(e - (long long)y * W);
            if (y < Hd && x < Wd) {
                // a float32 map divides in float32 (:384: np.float32 / np.float32); the rounded quotient, widened, is what
                // the reference compares with the threshold (numpy 1.26.4: in float64) and stores in column 3
                const double v = sizeof(T) == 4 ? (double)__fdiv_rn((float)heat[e], (float)maxv)
                                                : __ddiv_rn((double)heat[e], maxv);
                if (v > thr) {
                    const unsigned short d = depth[(long long)y * Wd + x];
                    if (d > 0) { m |= 1u << j; inten[j] = v; dep[j] = d; }
                }
            }
        }
    }
    const unsigned cnt = __popc(m);
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) s_warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned run = 0;
        for (int w = 0; w < DS_THREADS / 32; ++w) { const unsigned c = s_warp_tot[w]; if (lane == 0) s_warp_off[w] = run; run += c; }
        const unsigned long long total = run;
        const unsigned long long prefix = lookback_exclusive_prefix(state, tile, total, lane);
        if (lane == 0) {
            s_base = (long long)prefix;
            if ((long long)(tile + 1) * DS_TILE >= n) *d_count = (long long)(prefix + total);
        }
    }
    __syncthreads();
    long long off = s_base + s_warp_off[warp] + (inc - cnt);
#pragma unroll
    for (int j = 0; j < DS_ITEMS; ++j) {
        if ((m >> j) & 1u) {
            if (off < cap) {
                const long long e = e0 + j;
                const int y = (int)(e / W), x = (int)(e - (long long)y * W);
                const double d = (double)dep[j];
                // (x - cx) * depth / fx,  (y - cy) * depth / fy,  depth * 0.98   (:387-392)
                out4[4 * off + 0] = __ddiv_rn(__dmul_rn(__dsub_rn((double)x, cx), d), fx);
                out4[4 * off + 1] = __ddiv_rn(__dmul_rn(__dsub_rn((double)y, cy), d), fy);
                out4[4 * off + 2] = __dmul_rn(d, 0.98);
                out4[4 * off + 3] = inten[j];
            }
            ++off;
        }
    }
}

// calc_coordinates (:462-492): picked pixels (x, y) -> [x3d, y3d, depth]; depth 0 is skipped by the caller
__global__ void k_calc_coordinates(const int32_t *__restrict__ xs, const int32_t *__restrict__ ys, long long n,
                                   const uint16_t *__restrict__ depth, int Hd, int Wd, double fx, double fy, double cx,
                                   double cy, double *__restrict__ out3, unsigned char *__restrict__ valid)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = xs[i], y = ys[i];
    double d = 0.0;
    if (x >= 0 && y >= 0 && x < Wd && y < Hd) d = (double)depth[(long long)y * Wd + x];
    valid[i] = d > 0.0;
    out3[3 * i + 0] = __ddiv_rn(__dmul_rn(__dsub_rn((double)x, cx), d), fx);
    out3[3 * i + 1] = __ddiv_rn(__dmul_rn(__dsub_rn((double)y, cy), d), fy);
    out3[3 * i + 2] = d;
}

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;      // target points per shared-memory tile (24 KB of doubles)

// exact nearest neighbour, float64, brute force; q: [n][stride] (x,y,z first), target: [m][3]
__global__ void __launch_bounds__(NN_THREADS)
k_nearest(const double *__restrict__ q, int stride, long long n, const double *__restrict__ tp, long long m,
          const double *__restrict__ normals, double offset, int32_t *__restrict__ idx_out, double *__restrict__ aligned,
          double *__restrict__ offset_pts)
{
    __shared__ double s_t[NN_TILE * 3];
    const long long i = blockIdx.x * (long long)NN_THREADS + threadIdx.x;
    const bool ok = i < n;
    const double qx = ok ? q[i * stride] : 0.0, qy = ok ? q[i * stride + 1] : 0.0, qz = ok ? q[i * stride + 2] : 0.0;
    double best = INFINITY;
    long long bi = -1;
    for (long long t0 = 0; t0 < m; t0 += NN_TILE) {
        const int cnt = (int)((m - t0) < NN_TILE ? (m - t0) : NN_TILE);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 3; k += NN_THREADS) s_t[k] = tp[t0 * 3 + k];
        __syncthreads();
        if (ok) {
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
                const double dx = __dsub_rn(qx, s_t[3 * k]), dy = __dsub_rn(qy, s_t[3 * k + 1]), dz = __dsub_rn(qz, s_t[3 * k + 2]);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (d2 < best) { best = d2; bi = t0 + k; }       // ascending k: ties keep the smaller index
            }
        }
    }
    if (!ok) return;
    if (idx_out) idx_out[i] = (int32_t)bi;
    if (bi >= 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double p = tp[3 * bi + k];
            if (aligned) aligned[3 * i + k] = p;
            // nearest_point + normal * offset   (:452)
            if (offset_pts) offset_pts[3 * i + k] = __dadd_rn(p, __dmul_rn(normals[3 * bi + k], offset));
        }
    }
}

// The same search on the uniform grid of icp.cu (cells sized by the point density): the query's cell, then the shells
// of cells around it, until the best distance found is smaller than the distance to everything not yet visited.
// Same distance arithmetic, ties to the smaller ORIGINAL index: the result equals the scan's bit for bit.
__global__ void __launch_bounds__(NN_THREADS)
k_nearest_grid(const double *__restrict__ q, int stride, long long n, const double *__restrict__ tps,
               const double *__restrict__ tns, const uint32_t *__restrict__ orig, const int32_t *__restrict__ cell_start,
               const int32_t *__restrict__ cell_end, IcpGrid g, double offset, int32_t *__restrict__ idx_out,
               double *__restrict__ aligned, double *__restrict__ offset_pts)
{
    const long long i = blockIdx.x * (long long)NN_THREADS + threadIdx.x;
    if (i >= n) return;
    const double qp[3] = {q[i * stride], q[i * stride + 1], q[i * stride + 2]};
    double best = INFINITY;
    int bk = -1;
    uint32_t bo = 0xffffffffu;
    if (qp[0] == qp[0] && qp[1] == qp[1] && qp[2] == qp[2]) {
        int c[3];
        double slack = 1e-9 * g.cell;                    // cell coordinates are rounded: keep the bound on the safe side
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            c[k] = grid_coord(qp[k], g.lo[k], g.inv_cell, g.dim[k]);
            slack += 1e-12 * (fabs(qp[k]) + fabs(g.lo[k]) + g.cell * g.dim[k]);
        }
        const int rmax = max(max(g.dim[0], g.dim[1]), g.dim[2]);
        for (int r = 0; r < rmax; ++r) {
            const int z0 = max(c[2] - r, 0), z1 = min(c[2] + r, g.dim[2] - 1);
            const int y0 = max(c[1] - r, 0), y1 = min(c[1] + r, g.dim[1] - 1);
            const int x0 = max(c[0] - r, 0), x1 = min(c[0] + r, g.dim[0] - 1);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    // inside the shell only the two end cells of the row are new
                    const bool whole = (z == c[2] - r) || (z == c[2] + r) || (y == c[1] - r) || (y == c[1] + r);
                    const int step = (whole || r == 0) ? 1 : 2 * r;
                    for (int x = c[0] - r; x <= c[0] + r; x += step) {
                        if (x < x0 || x > x1) continue;
                        const int cell = (z * g.dim[1] + y) * g.dim[0] + x;
                        const int k0 = cell_start[cell], k1 = cell_end[cell];
                        for (int k = k0; k < k1; ++k) {
                            const double dx = __dsub_rn(qp[0], tps[3 * k]), dy = __dsub_rn(qp[1], tps[3 * k + 1]),
                                         dz = __dsub_rn(qp[2], tps[3 * k + 2]);
                            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                            const uint32_t o = orig[k];
                            if (d2 < best || (d2 == best && o < bo)) { best = d2; bk = k; bo = o; }
                        }
                    }
                }
            // distance from the query to everything outside the block of cells visited so far (faces at the rim of
            // the grid are open: no target lies beyond them)
            double bound = INFINITY;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (c[k] - r > 0) bound = fmin(bound, qp[k] - (g.lo[k] + (double)(c[k] - r) * g.cell));
                if (c[k] + r + 1 < g.dim[k]) bound = fmin(bound, g.lo[k] + (double)(c[k] + r + 1) * g.cell - qp[k]);
            }
            if (bound == INFINITY) break;                // the whole grid has been visited
            bound -= slack;
            if (bound > 0.0 && best < bound * bound) break;
        }
    }
    if (idx_out) idx_out[i] = bk >= 0 ? (int32_t)bo : -1;
    if (bk >= 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double p = tps[3 * bk + k];
            if (aligned) aligned[3 * i + k] = p;
            if (offset_pts) offset_pts[3 * i + k] = __dadd_rn(p, __dmul_rn(tns[3 * bk + k], offset));
        }
    }
}

}  // namespace

size_t depth_select_scratch_bytes(int64_t n_elems)
{
    return (size_t)((n_elems + DS_TILE - 1) / DS_TILE + 2) * sizeof(unsigned long long);
}

cudaError_t launch_depth_select(const void *heat, int dtype, int H, int W, const uint16_t *depth, int Hd, int Wd,
                                double thr, const double *K, double *out4, int64_t cap, unsigned long long *scratch,
                                long long *d_count, double *d_max, int *d_nan, cudaStream_t s)
{
    cudaError_t e;
    const long long n = (long long)H * W;
    if ((e = cudaMemsetAsync(d_count, 0, sizeof(long long), s)) != cudaSuccess) return e;
    if (n <= 0) return cudaSuccess;
    // max(heatmap): ordered-int encoding of -inf is the start value
    const long long ninf = (long long)0xfff0000000000000ull ^ 0x7fffffffffffffffll;
    if ((e = cudaMemcpyAsync(d_max, &ninf, 8, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d_nan, 0, sizeof(int), s)) != cudaSuccess) return e;
    const unsigned gb = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
    if (dtype == 1) k_max<double><<<gb, 256, 0, s>>>(static_cast<const double *>(heat), n, d_max, d_nan);
    else k_max<float><<<gb, 256, 0, s>>>(static_cast<const float *>(heat), n, d_max, d_nan);
    k_max_finish<<<1, 1, 0, s>>>(d_max, d_nan);
    if ((e = cudaMemsetAsync(scratch, 0, depth_select_scratch_bytes(n), s)) != cudaSuccess) return e;
    const unsigned tiles = (unsigned)((n + DS_TILE - 1) / DS_TILE);
    if (dtype == 1)
        k_depth_select<double><<<tiles, DS_THREADS, 0, s>>>(static_cast<const double *>(heat), H, W, depth, Hd, Wd, d_max, thr,
                                                            K[0], K[4], K[2], K[5], out4, cap, scratch, d_count);
    else
        k_depth_select<float><<<tiles, DS_THREADS, 0, s>>>(static_cast<const float *>(heat), H, W, depth, Hd, Wd, d_max, thr,
                                                           K[0], K[4], K[2], K[5], out4, cap, scratch, d_count);
    return cudaGetLastError();
}

cudaError_t launch_calc_coordinates(const int32_t *xs, const int32_t *ys, int64_t n, const uint16_t *depth, int Hd, int Wd,
                                    const double *K, double *out3, unsigned char *valid, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_calc_coordinates<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xs, ys, n, depth, Hd, Wd, K[0], K[4], K[2], K[5], out3, valid);
    return cudaGetLastError();
}

cudaError_t launch_nearest(const double *q, int stride, int64_t n, const double *target, int64_t m, const double *normals,
                           double offset, int32_t *idx, double *aligned, double *offset_pts, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_nearest<<<(unsigned)((n + NN_THREADS - 1) / NN_THREADS), NN_THREADS, 0, s>>>(q, stride, n, target, m, normals, offset, idx,
                                                                                 aligned, offset_pts);
    return cudaGetLastError();
}

cudaError_t launch_nearest_grid(const double *q, int stride, int64_t n, const IcpGridView &gv, bool has_normals, double offset,
                                int32_t *idx, double *aligned, double *offset_pts, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_nearest_grid<<<(unsigned)((n + NN_THREADS - 1) / NN_THREADS), NN_THREADS, 0, s>>>(
        q, stride, n, gv.tps, has_normals ? gv.tns : nullptr, gv.orig, gv.cell_start, gv.cell_end, gv.grid, offset, idx, aligned,
        has_normals ? offset_pts : nullptr);
    return cudaGetLastError();
}

}  // namespace dp
