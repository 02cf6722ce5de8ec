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

namespace dp {

namespace {

constexpr int ICP_THREADS = 128;
constexpr int ICP_TILE = 1024;        // target points per shared-memory tile
constexpr int ICP_NSUM = 29;          // count, sum d^2, 21 upper-triangle entries of J J^T, 6 entries of J r

struct Xf16 { double m[16]; };

__global__ void __launch_bounds__(ICP_THREADS)
k_icp_step(double *__restrict__ src, long long n, const double *__restrict__ tp, const double *__restrict__ tn, long long m,
           double max_d2, int apply_update, Xf16 U, int32_t *__restrict__ corr, double *__restrict__ partial)
{
    __shared__ double s_t[ICP_TILE * 3];
    __shared__ double s_red[ICP_THREADS / 32][ICP_NSUM];
    const long long i = blockIdx.x * (long long)ICP_THREADS + threadIdx.x;
    const bool ok = i < n;
    double qx = 0.0, qy = 0.0, qz = 0.0;
    if (ok) {
        qx = src[3 * i]; qy = src[3 * i + 1]; qz = src[3 * i + 2];
        if (apply_update) {
            // PointCloud::Transform: (U * [p, 1]).head<3>() / w, evaluated left to right
            double h[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                h[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(U.m[4 * r], qx), __dmul_rn(U.m[4 * r + 1], qy)),
                                           __dmul_rn(U.m[4 * r + 2], qz)), U.m[4 * r + 3]);
            qx = __ddiv_rn(h[0], h[3]); qy = __ddiv_rn(h[1], h[3]); qz = __ddiv_rn(h[2], h[3]);
            src[3 * i] = qx; src[3 * i + 1] = qy; src[3 * i + 2] = qz;
        }
    }
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
    double v[ICP_NSUM];
#pragma unroll
    for (int k = 0; k < ICP_NSUM; ++k) v[k] = 0.0;
    const bool has = ok && bi >= 0 && best <= max_d2;
    if (ok && corr) corr[i] = has ? (int32_t)bi : -1;
    if (has) {
        const double tx = tp[3 * bi], ty = tp[3 * bi + 1], tz = tp[3 * bi + 2];
        const double nx = tn[3 * bi], ny = tn[3 * bi + 1], nz = tn[3 * bi + 2];
        const double r = (qx - tx) * nx + (qy - ty) * ny + (qz - tz) * nz;
        double J[6];
        J[0] = qy * nz - qz * ny;
        J[1] = qz * nx - qx * nz;
        J[2] = qx * ny - qy * nx;
        J[3] = nx; J[4] = ny; J[5] = nz;
        v[0] = 1.0;
        v[1] = best;
        int e = 2;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) v[e++] = J[a] * J[b];
#pragma unroll
        for (int a = 0; a < 6; ++a) v[23 + a] = J[a] * r;
    }
    // fixed-shape tree: lanes by xor shuffles, then the four warps in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ICP_NSUM; ++k) {
        double x = v[k];
#pragma unroll
        for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) s_red[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < ICP_NSUM) {
        double x = s_red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < ICP_THREADS / 32; ++w) x += s_red[w][threadIdx.x];
        partial[(long long)blockIdx.x * ICP_NSUM + threadIdx.x] = x;
    }
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

}  // namespace dp
