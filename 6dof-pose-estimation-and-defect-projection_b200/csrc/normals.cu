// Point-cloud normal estimation: the fallback inside align_to_surface and the target preparation of the ICP row.
//
// Replaces o3d.geometry.PointCloud.estimate_normals(search_param=KDTreeSearchParamHybrid(radius, max_nn)) as called by
//   /root/reference/src/defect_projection.py:181-186 (radius 10, max_nn 30), :431-436 (align_to_surface, radius 0.1),
//   /root/reference/src/pose_estimation.py:301-306 (radius 2, max_nn 5).
// Open3D is a third-party dependency (open3d==0.18.0, absent offline): PARITY UNPINNED.  What is restated is its
// published algorithm (cpp/open3d/geometry/EstimateNormals.cpp, KDTreeFlann.cpp, utility/Eigen.cpp, as recalled):
//   neighbours = the max_nn nearest points (the point itself included) with squared distance < radius^2
//   fewer than 3 neighbours: covariance = identity, else covariance = E[x x^T] - E[x] E[x]^T from nine cumulants
//   normal = eigenvector of the smallest eigenvalue by the closed form of Eberly's "A Robust Eigensolver for 3x3
//            Symmetric Matrices" on covariance / max coefficient (fast_normal_computation = True, the default)
//   zero normal -> the existing normal, or (0, 0, 1); an existing normal keeps its side (flip when the dot is < 0).
// A k-d tree returns equidistant neighbours in an arbitrary order; here neighbours are ordered by (distance, index)
// and the cumulants are summed in that order, every operation individually rounded, so the covariance is
// reproducible and equals the CPU restatement (oracle.estimate_normals) bit for bit.
//
// One thread per point, walked in the cell order of the uniform grid icp.cu builds (cells >= radius, 27 cells per
// query), so that the threads of a warp read the same cells; the sorted neighbour list lives in local memory.
#include "dp_internal.cuh"

namespace dp {

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 cross3(const V3 &a, const V3 &b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot3(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// A = symmetric (a00 a01 a02; . a11 a12; . . a22)
struct Sym3 { double a00, a01, a02, a11, a12, a22; };

__device__ V3 eigenvector0(const Sym3 &A, double ev)
{
    const V3 r0 = {A.a00 - ev, A.a01, A.a02}, r1 = {A.a01, A.a11 - ev, A.a12}, r2 = {A.a02, A.a12, A.a22 - ev};
    const V3 c01 = cross3(r0, r1), c02 = cross3(r0, r2), c12 = cross3(r1, r2);
    const double d0 = dot3(c01, c01), d1 = dot3(c02, c02), d2 = dot3(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) imax = 2;
    if (imax == 0) { const double s = sqrt(d0); return {c01.x / s, c01.y / s, c01.z / s}; }
    if (imax == 1) { const double s = sqrt(d1); return {c02.x / s, c02.y / s, c02.z / s}; }
    const double s = sqrt(d2);
    return {c12.x / s, c12.y / s, c12.z / s};
}

__device__ V3 eigenvector1(const Sym3 &A, const V3 &e0, double ev)
{
    V3 U;
    if (fabs(e0.x) > fabs(e0.y)) {
        const double inv = 1.0 / sqrt(e0.x * e0.x + e0.z * e0.z);
        U = {-e0.z * inv, 0.0, e0.x * inv};
    } else {
        const double inv = 1.0 / sqrt(e0.y * e0.y + e0.z * e0.z);
        U = {0.0, e0.z * inv, -e0.y * inv};
    }
    const V3 V = cross3(e0, U);
    const V3 AU = {A.a00 * U.x + A.a01 * U.y + A.a02 * U.z, A.a01 * U.x + A.a11 * U.y + A.a12 * U.z,
                   A.a02 * U.x + A.a12 * U.y + A.a22 * U.z};
    const V3 AV = {A.a00 * V.x + A.a01 * V.y + A.a02 * V.z, A.a01 * V.x + A.a11 * V.y + A.a12 * V.z,
                   A.a02 * V.x + A.a12 * V.y + A.a22 * V.z};
    double m00 = dot3(U, AU) - ev, m01 = dot3(U, AV), m11 = dot3(V, AV) - ev;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        if (fmax(a00, a01) > 0.0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1.0 / sqrt(1.0 + m00 * m00); m00 *= m01; }
            return {m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z};
        }
        return U;
    }
    if (fmax(a11, a01) > 0.0) {
        if (a11 >= a01) { m01 /= m11; m11 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m11; }
        else { m11 /= m01; m01 = 1.0 / sqrt(1.0 + m11 * m11); m11 *= m01; }
        return {m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z};
    }
    return U;
}

// eigenvector of the smallest eigenvalue of the covariance C (FastEigen3x3)
__device__ V3 smallest_eigenvector(const Sym3 &C)
{
    const double mx = fmax(fmax(fmax(C.a00, C.a01), fmax(C.a02, C.a11)), fmax(C.a12, C.a22));
    if (mx == 0.0) return {0.0, 0.0, 0.0};
    const Sym3 A = {C.a00 / mx, C.a01 / mx, C.a02 / mx, C.a11 / mx, C.a12 / mx, C.a22 / mx};
    const double norm = A.a01 * A.a01 + A.a02 * A.a02 + A.a12 * A.a12;
    if (norm > 0.0) {
        const double q = (A.a00 + A.a11 + A.a22) / 3.0;
        const double b00 = A.a00 - q, b11 = A.a11 - q, b22 = A.a22 - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0);
        const double c00 = b11 * b22 - A.a12 * A.a12;
        const double c01 = A.a01 * b22 - A.a12 * A.a02;
        const double c02 = A.a01 * A.a12 - b11 * A.a02;
        const double det = (b00 * c00 - A.a01 * c01 + A.a02 * c02) / (p * p * p);
        const double half_det = fmin(fmax(det * 0.5, -1.0), 1.0);
        const double angle = acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        const double beta2 = cos(angle) * 2.0;
        const double beta0 = cos(angle + two_thirds_pi) * 2.0;
        const double beta1 = -(beta0 + beta2);
        const double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (half_det >= 0.0) {
            const V3 v2 = eigenvector0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            const V3 v1 = eigenvector1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross3(v1, v2);
        }
        const V3 v0 = eigenvector0(A, e0);
        if (e0 < e1 && e0 < e2) return v0;
        const V3 v1 = eigenvector1(A, v0, e1);
        if (e1 < e0 && e1 < e2) return v1;
        return cross3(v0, v1);
    }
    if (C.a00 < C.a11 && C.a00 < C.a22) return {1.0, 0.0, 0.0};
    if (C.a11 < C.a00 && C.a11 < C.a22) return {0.0, 1.0, 0.0};
    return {0.0, 0.0, 1.0};
}

__global__ void __launch_bounds__(128)
k_estimate_normals(const double *__restrict__ tps, const uint32_t *__restrict__ orig, const int32_t *__restrict__ cell_start,
                   const int32_t *__restrict__ cell_end, IcpGrid g, long long n, double r2, int max_nn,
                   double *__restrict__ normals, int has_normals, int32_t *__restrict__ neighbours)
{
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;      // position in cell order
    if (k >= n) return;
    const double qx = tps[3 * k], qy = tps[3 * k + 1], qz = tps[3 * k + 2];
    const uint32_t me = orig[k];
    double ld2[NORMALS_MAX_NN];
    uint32_t lo[NORMALS_MAX_NN];
    int lk[NORMALS_MAX_NN];
    int cnt = 0;
    if (qx == qx && qy == qy && qz == qz) {
        const int cx = grid_coord(qx, g.lo[0], g.inv_cell, g.dim[0]);
        const int cy = grid_coord(qy, g.lo[1], g.inv_cell, g.dim[1]);
        const int cz = grid_coord(qz, g.lo[2], g.inv_cell, g.dim[2]);
        for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dim[2] - 1); ++z)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y)
                for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x) {
                    const int cell = (z * g.dim[1] + y) * g.dim[0] + x;
                    const int k0 = cell_start[cell], k1 = cell_end[cell];
                    for (int j = k0; j < k1; ++j) {
                        const double dx = __dsub_rn(qx, tps[3 * j]), dy = __dsub_rn(qy, tps[3 * j + 1]), dz = __dsub_rn(qz, tps[3 * j + 2]);
                        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (!(d2 < r2)) continue;
                        const uint32_t o = orig[j];
                        if (cnt == max_nn && !(d2 < ld2[cnt - 1] || (d2 == ld2[cnt - 1] && o < lo[cnt - 1]))) continue;
                        int p = cnt < max_nn ? cnt++ : cnt - 1;
                        while (p > 0 && (d2 < ld2[p - 1] || (d2 == ld2[p - 1] && o < lo[p - 1]))) {
                            ld2[p] = ld2[p - 1]; lo[p] = lo[p - 1]; lk[p] = lk[p - 1];
                            --p;
                        }
                        ld2[p] = d2; lo[p] = o; lk[p] = j;
                    }
                }
    }
    if (neighbours) neighbours[me] = cnt;
    Sym3 C = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
    if (cnt >= 3) {
        double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < cnt; ++i) {
            const int j = lk[i];
            const double x = tps[3 * j], y = tps[3 * j + 1], z = tps[3 * j + 2];
            c[0] = __dadd_rn(c[0], x); c[1] = __dadd_rn(c[1], y); c[2] = __dadd_rn(c[2], z);
            c[3] = __dadd_rn(c[3], __dmul_rn(x, x)); c[4] = __dadd_rn(c[4], __dmul_rn(x, y)); c[5] = __dadd_rn(c[5], __dmul_rn(x, z));
            c[6] = __dadd_rn(c[6], __dmul_rn(y, y)); c[7] = __dadd_rn(c[7], __dmul_rn(y, z)); c[8] = __dadd_rn(c[8], __dmul_rn(z, z));
        }
        const double m = (double)cnt;
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i] = __ddiv_rn(c[i], m);
        C.a00 = __dsub_rn(c[3], __dmul_rn(c[0], c[0]));
        C.a11 = __dsub_rn(c[6], __dmul_rn(c[1], c[1]));
        C.a22 = __dsub_rn(c[8], __dmul_rn(c[2], c[2]));
        C.a01 = __dsub_rn(c[4], __dmul_rn(c[0], c[1]));
        C.a02 = __dsub_rn(c[5], __dmul_rn(c[0], c[2]));
        C.a12 = __dsub_rn(c[7], __dmul_rn(c[1], c[2]));
    }
    V3 nrm = smallest_eigenvector(C);
    double *out = normals + 3ll * me;
    if (nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z == 0.0) {
        if (has_normals) nrm = {out[0], out[1], out[2]};
        else nrm = {0.0, 0.0, 1.0};
    }
    if (has_normals && nrm.x * out[0] + nrm.y * out[1] + nrm.z * out[2] < 0.0) nrm = {-nrm.x, -nrm.y, -nrm.z};
    out[0] = nrm.x; out[1] = nrm.y; out[2] = nrm.z;
}

}  // namespace

cudaError_t launch_estimate_normals(const IcpGridView &gv, int64_t n, double radius, int max_nn, double *normals, int has_normals,
                                    int32_t *neighbours, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_estimate_normals<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(gv.tps, gv.orig, gv.cell_start, gv.cell_end, gv.grid, n,
                                                                  radius * radius, max_nn, normals, has_normals, neighbours);
    return cudaGetLastError();
}

}  // namespace dp
