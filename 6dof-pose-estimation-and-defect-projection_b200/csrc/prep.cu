// The two steps either side of the back-projection.
//
// BEFORE the path -- DataReader.get_heatmap, /root/reference/datareader.py:639-675:
//     heatmap = data - np.min(data); heatmap = heatmap / np.max(heatmap)                     :658-659
//     heatmap_vis = cv2.resize(heatmap, (o, o), interpolation=cv2.INTER_LINEAR)              :664-665
//     heatmap_full = zeros(H, W); heatmap_full[y0:y0+o, x0:x0+o] = heatmap_vis               :669-674
//   cv2 is a third-party dependency; what it computes for CV_32F / CV_64F INTER_LINEAR in the pip wheel
//   (opencv-python 4.13, IPP enabled) was pinned empirically against the real cv2 in the build container
//   (tests/golden/make_golden.py, section E) and is restated here operation by operation:
//     v  = fma(j + 0.5, src/dst, -0.5) in float64;  s = floor(v);  f = v - s
//     s < 0 -> (0, f = 0);  s >= src-1 -> (src-1, f = 0);  f is rounded to the image's type
//     row(y, j) = fma(n[y][s+1] - n[y][s], fx, n[y][s])          (horizontal first, in the image's type)
//     out(i, j) = fma(row(s_y+1, j) - row(s_y, j), fy, row(s_y, j))
//   One kernel produces the padded H x W frame directly from the raw map (normalisation fused, 4 taps per
//   output pixel, no intermediate images).
//
// AFTER the path -- create_intersection_pcd (/root/reference/src/defect_projection.py:268-294), the boolean
// selection of the hits (:259-264), PointCloud.transform (run.py:118, :196-200) and the arrays
// update_dash_data ships to the viewer (src/web_vis.py:203-217):
//     colors = jet((I - min(I)) / (max(I) - min(I)))[:, :3]
//   k_hit_minmax + k_pack_hits: order-preserving selection of the rays that hit (ranks from ballots and
//   decoupled look-back, as compact.cu), colour lookup in matplotlib's 256-entry 'jet' table, and the rigid
//   transform of the hit points, in one pass over the per-ray arrays.
#include "dp_internal.cuh"

namespace dp {

namespace {

// ------------------------------------------------------------------------------------------ ordered atomics
__device__ __forceinline__ long long ord_encode(double v)
{
    const long long b = __double_as_longlong(v);
    return b < 0 ? (b ^ 0x7fffffffffffffffll) : b;
}
__device__ __forceinline__ double ord_decode(long long b)
{
    return __longlong_as_double(b < 0 ? (b ^ 0x7fffffffffffffffll) : b);
}

// mm[0] = ordered max, mm[1] = ordered min, mm[2] = NaN seen (as long long), mm[3] = count
template <typename T>
__global__ void __launch_bounds__(256)
k_minmax(const T *__restrict__ a, const int32_t *__restrict__ select, long long n, long long *mm)
{
    double hi = -INFINITY, lo = INFINITY;
    int nan = 0;
    long long cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (select && select[i] < 0) continue;
        const double v = (double)a[i];
        ++cnt;
        if (v != v) nan = 1; else { hi = fmax(hi, v); lo = fmin(lo, v); }
    }
    for (int d = 16; d; d >>= 1) {
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        nan |= __shfl_xor_sync(0xffffffffu, nan, d);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) {
            atomicMax(&mm[0], ord_encode(hi));
            atomicMin(&mm[1], ord_encode(lo));
            atomicAdd(reinterpret_cast<unsigned long long *>(&mm[3]), (unsigned long long)cnt);
        }
        if (nan) atomicOr(reinterpret_cast<unsigned long long *>(&mm[2]), 1ull);
    }
}

// ------------------------------------------------------------------------------------------ heatmap preparation
template <typename T> struct Arith;
template <> struct Arith<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
};
template <> struct Arith<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};

template <typename T>
__device__ __forceinline__ void lin_coef(int j, int src, double scale, int &s0, int &s1, T &f)
{
    const double v = __fma_rn((double)j + 0.5, scale, -0.5);
    double fl = floor(v);
    double fr = __dsub_rn(v, fl);
    long long s = (long long)fl;
    if (s < 0) { s = 0; fr = 0.0; }
    if (s >= src - 1) { s = src - 1; fr = 0.0; }
    s0 = (int)s;
    s1 = s0 + 1 < src ? s0 + 1 : src - 1;
    f = (T)fr;
}

template <typename T, typename O>
__global__ void __launch_bounds__(256)
k_prepare_heatmap(const T *__restrict__ data, int sh, int sw, int H, int W, int o, int y0, int x0, double scale_y,
                  double scale_x, const long long *__restrict__ mm, O *__restrict__ out)
{
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)H * W) return;
    const int y = (int)(e / W), x = (int)(e - (long long)y * W);
    const int i = y - y0, j = x - x0;
    if (i < 0 || i >= o || j < 0 || j >= o) { out[e] = (O)0; return; }
    // np.min / np.max: NaN if any element is NaN
    T mn, mx;
    if (mm[2]) { mn = (T)NAN; mx = (T)NAN; }
    else { mn = (T)ord_decode(mm[1]); mx = (T)ord_decode(mm[0]); }
    const T range = Arith<T>::sub(mx, mn);          // = np.max(data - min): rounding is monotonic
    int xs0, xs1, ys0, ys1;
    T fx, fy;
    lin_coef<T>(j, sw, scale_x, xs0, xs1, fx);
    lin_coef<T>(i, sh, scale_y, ys0, ys1, fy);
    const T *r0 = data + (long long)ys0 * sw, *r1 = data + (long long)ys1 * sw;
    const T a00 = Arith<T>::div(Arith<T>::sub(__ldg(r0 + xs0), mn), range);
    const T a01 = Arith<T>::div(Arith<T>::sub(__ldg(r0 + xs1), mn), range);
    const T a10 = Arith<T>::div(Arith<T>::sub(__ldg(r1 + xs0), mn), range);
    const T a11 = Arith<T>::div(Arith<T>::sub(__ldg(r1 + xs1), mn), range);
    const T row0 = Arith<T>::fma(Arith<T>::sub(a01, a00), fx, a00);
    const T row1 = Arith<T>::fma(Arith<T>::sub(a11, a10), fx, a10);
    out[e] = (O)Arith<T>::fma(Arith<T>::sub(row1, row0), fy, row0);
}

// ------------------------------------------------------------------------------------------ rigid transform
// Open3D's PointCloud::Transform: p' = (T * [p, 1]).head<3>() / w.  Stated order: ((T0*x + T1*y) + T2*z) + T3.
struct Xf16 { double m[16]; };

__device__ __forceinline__ void xf_point(const Xf16 &T, double x, double y, double z, double &ox, double &oy, double &oz)
{
    double h[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        h[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[4 * r], x), __dmul_rn(T.m[4 * r + 1], y)), __dmul_rn(T.m[4 * r + 2], z)),
                         T.m[4 * r + 3]);
    ox = __ddiv_rn(h[0], h[3]); oy = __ddiv_rn(h[1], h[3]); oz = __ddiv_rn(h[2], h[3]);
}

__global__ void __launch_bounds__(256) k_transform_points(double *__restrict__ p, long long n, Xf16 T)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z;
    xf_point(T, p[3 * i], p[3 * i + 1], p[3 * i + 2], x, y, z);
    p[3 * i] = x; p[3 * i + 1] = y; p[3 * i + 2] = z;
}

// ------------------------------------------------------------------------------------------ hit packing
constexpr int PK_THREADS = 256;
constexpr int PK_ITEMS = 4;
constexpr int PK_TILE = PK_THREADS * PK_ITEMS;

template <typename T>
__global__ void __launch_bounds__(PK_THREADS)
k_pack_hits(const T *__restrict__ inten, const int32_t *__restrict__ face, const uint32_t *__restrict__ pixel,
            const double *__restrict__ point64, long long n, const long long *__restrict__ mm, const double *__restrict__ lut,
            int has_T, Xf16 Tm, double *__restrict__ points, double *__restrict__ colors, int32_t *__restrict__ face_out,
            uint32_t *__restrict__ pixel_out, double *__restrict__ inten_out, long long cap, unsigned long long *scratch,
            long long *d_count)
{
    __shared__ unsigned s_tile, s_warp_tot[PK_THREADS / 32], s_warp_off[PK_THREADS / 32];
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned *>(scratch), 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    unsigned long long *state = scratch + 1;
    const long long e0 = (long long)tile * PK_TILE + (long long)tid * PK_ITEMS;
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < PK_ITEMS; ++j) {
        const long long e = e0 + j;
        if (e < n && (!face || face[e] >= 0)) m |= 1u << j;
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
        for (int w = 0; w < PK_THREADS / 32; ++w) { const unsigned c = s_warp_tot[w]; if (lane == 0) s_warp_off[w] = run; run += c; }
        const unsigned long long total = run;
        const unsigned long long prefix = lookback_exclusive_prefix(state, tile, total, lane);
        if (lane == 0) {
            s_base = (long long)prefix;
            if ((long long)(tile + 1) * PK_TILE >= n) *d_count = (long long)(prefix + total);
        }
    }
    __syncthreads();
    if (!m) return;
    // (I - min) / (max - min) in float64 (:286-289); max == min gives 0/0 = NaN -> the colour map's "bad" entry
    const bool nan = mm[2] != 0;
    const double lo = nan ? NAN : ord_decode(mm[1]), hi = nan ? NAN : ord_decode(mm[0]);
    const double range = __dsub_rn(hi, lo);
    long long off = s_base + s_warp_off[warp] + (inc - cnt);
#pragma unroll
    for (int j = 0; j < PK_ITEMS; ++j) {
        if (!((m >> j) & 1u)) continue;
        if (off < cap) {
            const long long e = e0 + j;
            const double I = (double)inten[e];
            if (colors) {
                const double x = __ddiv_rn(__dsub_rn(I, lo), range);
                double r = 0.0, g = 0.0, b = 0.0;            // NaN -> bad colour (0, 0, 0)
                if (x == x) {
                    // matplotlib Colormap.__call__: xa = x*N; xa < 0 -> under (= lut[0]); xa == N -> N-1;
                    // truncation; > N-1 -> over (= lut[N-1])
                    double xa = __dmul_rn(x, 256.0);
                    int k;
                    if (xa < 0.0) k = 0;
                    else if (xa >= 256.0) k = 255;
                    else k = (int)xa;
                    r = __ldg(lut + 3 * k); g = __ldg(lut + 3 * k + 1); b = __ldg(lut + 3 * k + 2);
                }
                colors[3 * off] = r; colors[3 * off + 1] = g; colors[3 * off + 2] = b;
            }
            if (points) {
                double x = point64[3 * e], y = point64[3 * e + 1], z = point64[3 * e + 2];
                if (has_T) xf_point(Tm, x, y, z, x, y, z);
                points[3 * off] = x; points[3 * off + 1] = y; points[3 * off + 2] = z;
            }
            if (face_out) face_out[off] = face ? face[e] : -1;
            if (pixel_out) pixel_out[off] = pixel ? pixel[e] : (uint32_t)e;
            if (inten_out) inten_out[off] = I;
        }
        ++off;
    }
}


// Lean hit records for the multi-GPU gather: the rays of [first, first + n) that hit, in ray order,
// as rows of 3 words (pixel, t_hit bits, face) or 6 words (+ float32 hit point).  Same ballot ranks + decoupled
// look-back as k_pack_hits; the count also goes to `m_async` (device or mapped pinned-host memory) when given.
__global__ void __launch_bounds__(PK_THREADS)
k_pack_records(const uint32_t *__restrict__ pixel, const float *__restrict__ t_hit, const int32_t *__restrict__ face,
               const float *__restrict__ point, long long n, long long first, uint32_t *__restrict__ rec, long long cap,
               unsigned long long *scratch, long long *d_count, long long *m_async)
{
    __shared__ unsigned s_tile, s_warp_tot[PK_THREADS / 32], s_warp_off[PK_THREADS / 32];
    __shared__ long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned *>(scratch), 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    unsigned long long *state = scratch + 1;
    const long long e0 = (long long)tile * PK_TILE + (long long)tid * PK_ITEMS;
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < PK_ITEMS; ++j) {
        const long long e = e0 + j;
        if (e < n && face[first + e] >= 0) m |= 1u << j;
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
        for (int w = 0; w < PK_THREADS / 32; ++w) { const unsigned c = s_warp_tot[w]; if (lane == 0) s_warp_off[w] = run; run += c; }
        const unsigned long long total = run;
        const unsigned long long prefix = lookback_exclusive_prefix(state, tile, total, lane);
        if (lane == 0) {
            s_base = (long long)prefix;
            if ((long long)(tile + 1) * PK_TILE >= n) {
                *d_count = (long long)(prefix + total);
                if (m_async) *m_async = (long long)(prefix + total);
            }
        }
    }
    __syncthreads();
    if (!m) return;
    const int stride = point ? 6 : 3;
    long long off = s_base + s_warp_off[warp] + (inc - cnt);
#pragma unroll
    for (int j = 0; j < PK_ITEMS; ++j) {
        if (!((m >> j) & 1u)) continue;
        if (off < cap) {
            const long long e = first + e0 + j;
            uint32_t *o = rec + off * stride;
            o[0] = pixel ? pixel[e] : (uint32_t)e;
            o[1] = __float_as_uint(t_hit[e]);
            o[2] = (uint32_t)face[e];
            if (point) { o[3] = __float_as_uint(point[3 * e]); o[4] = __float_as_uint(point[3 * e + 1]); o[5] = __float_as_uint(point[3 * e + 2]); }
        }
        ++off;
    }
}

}  // namespace

// matplotlib's LinearSegmentedColormap('jet', N=256) table (colors._create_lookup_table), float64, RGB.
void jet_lut_host(double *lut /* [256*3] */)
{
    static const double seg_r[][2] = {{0.00, 0.0}, {0.35, 0.0}, {0.66, 1.0}, {0.89, 1.0}, {1.00, 0.5}};
    static const double seg_g[][2] = {{0.000, 0.0}, {0.125, 0.0}, {0.375, 1.0}, {0.640, 1.0}, {0.910, 0.0}, {1.000, 0.0}};
    static const double seg_b[][2] = {{0.00, 0.5}, {0.11, 1.0}, {0.34, 1.0}, {0.65, 0.0}, {1.00, 0.0}};
    const double (*segs[3])[2] = {seg_r, seg_g, seg_b};
    const int cnt[3] = {5, 6, 5};
    const int N = 256;
    for (int c = 0; c < 3; ++c) {
        double xk[8];
        for (int k = 0; k < cnt[c]; ++k) { volatile double t = segs[c][k][0] * (double)(N - 1); xk[k] = t; }
        lut[0 * 3 + c] = segs[c][0][1];
        lut[(N - 1) * 3 + c] = segs[c][cnt[c] - 1][1];
        for (int i = 1; i < N - 1; ++i) {
            const double g = (double)i;
            int k = 0;
            while (k < cnt[c] && xk[k] < g) ++k;       // searchsorted(side='left')
            volatile double num = g - xk[k - 1];
            volatile double den = xk[k] - xk[k - 1];
            volatile double w = num / den;
            volatile double dy = segs[c][k][1] - segs[c][k - 1][1];
            volatile double p = w * dy;
            volatile double v = p + segs[c][k - 1][1];
            double r = v;
            r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
            lut[i * 3 + c] = r;
        }
    }
}

static cudaError_t init_minmax(long long *mm, cudaStream_t s)
{
    const long long init[4] = {(long long)0xfff0000000000000ull ^ 0x7fffffffffffffffll,   // ordered(-inf)
                               (long long)0x7ff0000000000000ull,                            // ordered(+inf)
                               0, 0};
    return cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, s);
}

cudaError_t launch_minmax(const void *a, int dtype, const int32_t *select, int64_t n, long long *mm, cudaStream_t s)
{
    cudaError_t e = init_minmax(mm, s);
    if (e != cudaSuccess || n <= 0) return e;
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == 1) k_minmax<double><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const double *>(a), select, n, mm);
    else k_minmax<float><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const float *>(a), select, n, mm);
    return cudaGetLastError();
}

cudaError_t launch_prepare_heatmap(const void *data, int dtype, int sh, int sw, int H, int W, void *out, int out_dtype,
                                   long long *mm, cudaStream_t s)
{
    cudaError_t e = launch_minmax(data, dtype, nullptr, (int64_t)sh * sw, mm, s);
    if (e != cudaSuccess) return e;
    const long long n = (long long)H * W;
    if (n <= 0) return cudaSuccess;
    const int o = H < W ? H : W;
    const int y0 = (H - o) / 2, x0 = (W - o) / 2;
    const double sy = o > 0 ? (double)sh / (double)o : 0.0, sx = o > 0 ? (double)sw / (double)o : 0.0;
    const unsigned gb = (unsigned)((n + 255) / 256);
#define DP_PREP(T, O)                                                                                           \
    k_prepare_heatmap<T, O><<<gb, 256, 0, s>>>(static_cast<const T *>(data), sh, sw, H, W, o, y0, x0, sy, sx, mm, \
                                               static_cast<O *>(out))
    if (dtype == 1 && out_dtype == 1) DP_PREP(double, double);
    else if (dtype == 1) DP_PREP(double, float);
    else if (out_dtype == 1) DP_PREP(float, double);
    else DP_PREP(float, float);
#undef DP_PREP
    return cudaGetLastError();
}

cudaError_t launch_transform_points(double *p, int64_t n, const double *T_host, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    Xf16 T;
    for (int i = 0; i < 16; ++i) T.m[i] = T_host[i];
    k_transform_points<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, T);
    return cudaGetLastError();
}

size_t pack_scratch_bytes(int64_t n) { return (size_t)((n + PK_TILE - 1) / PK_TILE + 2) * sizeof(unsigned long long); }

cudaError_t launch_pack_hits(const void *inten, int dtype, const int32_t *face, const uint32_t *pixel, const double *point64,
                             int64_t n, const double *T_host, const double *lut, long long *mm, double *points, double *colors,
                             int32_t *face_out, uint32_t *pixel_out, double *inten_out, int64_t cap,
                             unsigned long long *scratch, long long *d_count, cudaStream_t s)
{
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_count, 0, sizeof(long long), s)) != cudaSuccess) return e;
    if ((e = launch_minmax(inten, dtype, face, n, mm, s)) != cudaSuccess) return e;
    if (n <= 0) return cudaSuccess;
    if ((e = cudaMemsetAsync(scratch, 0, pack_scratch_bytes(n), s)) != cudaSuccess) return e;
    Xf16 T;
    for (int i = 0; i < 16; ++i) T.m[i] = T_host ? T_host[i] : (i % 5 == 0 ? 1.0 : 0.0);
    const unsigned tiles = (unsigned)((n + PK_TILE - 1) / PK_TILE);
    if (dtype == 1)
        k_pack_hits<double><<<tiles, PK_THREADS, 0, s>>>(static_cast<const double *>(inten), face, pixel, point64, n, mm, lut,
                                                         T_host != nullptr, T, points, colors, face_out, pixel_out, inten_out,
                                                         cap, scratch, d_count);
    else
        k_pack_hits<float><<<tiles, PK_THREADS, 0, s>>>(static_cast<const float *>(inten), face, pixel, point64, n, mm, lut,
                                                        T_host != nullptr, T, points, colors, face_out, pixel_out, inten_out,
                                                        cap, scratch, d_count);
    return cudaGetLastError();
}

cudaError_t launch_pack_records(const uint32_t *pixel, const float *t_hit, const int32_t *face, const float *point, int64_t n,
                                int64_t first, uint32_t *rec, int64_t cap, unsigned long long *scratch, long long *d_count,
                                long long *m_async, cudaStream_t s)
{
    cudaError_t e;
    if ((e = cudaMemsetAsync(scratch, 0, pack_scratch_bytes(n > 0 ? n : 1), s)) != cudaSuccess) return e;
    // n == 0 still launches one tile: it publishes the zero count
    const unsigned tiles = (unsigned)((n + PK_TILE - 1) / PK_TILE);
    k_pack_records<<<tiles ? tiles : 1, PK_THREADS, 0, s>>>(pixel, t_hit, face, point, n, first, rec, cap, scratch, d_count, m_async);
    return cudaGetLastError();
}

}  // namespace dp
