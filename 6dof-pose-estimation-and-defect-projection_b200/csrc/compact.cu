// Subsystem (1): heatmap threshold filtering + order-preserving stream compaction.
//
// Replaces heatmap_to_points (/root/reference/src/defect_projection.py:165-179):
//     y, x = np.where(heatmap > threshold); intensities = heatmap[y, x]
// Output order is row-major (y ascending, then x), one entry per pixel with value > thr
// (strict; NaN never passes).  A batch of frames is one flat array, so the emitted pixel
// index is frame*H*W + y*W + x.
//
// One pass over the heatmap (HBM streaming, 16-byte loads), ranks from warp ballots,
// tile offsets by decoupled look-back (tiles take their id from an atomic ticket so a
// waiting tile only ever waits for a tile that is already running).
#include "dp_internal.cuh"

#include <string.h>

namespace dp {

namespace {

constexpr int CT_THREADS = 256;
#ifndef DP_CT_CHUNKS
#define DP_CT_CHUNKS 4
#endif
constexpr int CT_CHUNKS = DP_CT_CHUNKS;            // 16-byte (f32) / 32-byte (f64) loads per thread
constexpr int CT_TILE = CT_THREADS * CT_CHUNKS * 4;  // 4096 pixels per tile


template <typename T>
struct Vec4 {
    T v[4];
};

__device__ __forceinline__ void load4(const float *p, Vec4<float> &o)
{
    float4 q = __ldg(reinterpret_cast<const float4 *>(p));
    o.v[0] = q.x; o.v[1] = q.y; o.v[2] = q.z; o.v[3] = q.w;
}
__device__ __forceinline__ void load4(const double *p, Vec4<double> &o)
{
    double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    o.v[0] = a.x; o.v[1] = a.y; o.v[2] = b.x; o.v[3] = b.y;
}

template <typename T>
__device__ __forceinline__ void
compact_tile(const T *__restrict__ heat, long long n, T thr, uint32_t *__restrict__ pixel,
             float *__restrict__ intensity, long long cap, unsigned long long *scratch, long long *d_count, long long *early_n,
             int aligned)
{
    __shared__ unsigned s_tile;
    __shared__ unsigned s_warp_tot[CT_CHUNKS * (CT_THREADS / 32)];
    __shared__ unsigned s_warp_off[CT_CHUNKS * (CT_THREADS / 32)];
    __shared__ unsigned s_block_total;
    __shared__ long long s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned *>(scratch), 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const long long tile_base = (long long)tile * CT_TILE;
    unsigned long long *state = scratch + 1;

    Vec4<T> val[CT_CHUNKS];
    unsigned m[CT_CHUNKS];
    unsigned excl[CT_CHUNKS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int c = 0; c < CT_CHUNKS; ++c) {
        const long long e0 = tile_base + (long long)c * (CT_THREADS * 4) + tid * 4;
        if (aligned && e0 + 3 < n) {
            load4(heat + e0, val[c]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) val[c].v[j] = (e0 + j < n) ? heat[e0 + j] : thr;   // thr > thr is false
        }
        unsigned mm = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) mm |= (val[c].v[j] > thr) ? (1u << j) : 0u;
        m[c] = mm;
        // rank of this thread's first selected pixel among the warp's 128 pixels of chunk c
        unsigned r = 0, tot = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned b = __ballot_sync(0xffffffffu, (mm >> j) & 1u);
            r += __popc(b & lt);
            tot += __popc(b);
        }
        excl[c] = r;
        if (lane == 0) s_warp_tot[c * (CT_THREADS / 32) + warp] = tot;
    }
    __syncthreads();
    if (warp == 0) {
        // the partial counts in pixel order (chunk-major, then warp), 32 at a time: exclusive scan by shuffles
        constexpr int NPART = CT_CHUNKS * (CT_THREADS / 32);
        static_assert(NPART % 32 == 0, "partial counts come in groups of 32");
        unsigned carry = 0;
#pragma unroll
        for (int g = 0; g < NPART; g += 32) {
            const unsigned x = s_warp_tot[g + lane];
            unsigned inc = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += y;
            }
            s_warp_off[g + lane] = carry + inc - x;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 31) s_block_total = carry;
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned long long total = s_block_total;
        const unsigned long long prefix = lookback_exclusive_prefix(state, tile, total, lane);
        if (lane == 0) {
            s_base = (long long)prefix;
            if (tile_base + CT_TILE >= n) {                                        // last tile
                *d_count = (long long)(prefix + total);
                // the ray count as soon as it exists (mapped pinned-host memory): a caller that shards the frame's rays over
                // several GPUs sizes its gather while the traversal still runs
                if (early_n) *reinterpret_cast<volatile long long *>(early_n) = (long long)(prefix + total);
            }
        }
    }
    __syncthreads();
    const long long base = s_base;
#pragma unroll
    for (int c = 0; c < CT_CHUNKS; ++c) {
        if (!m[c]) continue;
        long long off = base + s_warp_off[c * (CT_THREADS / 32) + warp] + excl[c];
        const long long e0 = tile_base + (long long)c * (CT_THREADS * 4) + tid * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((m[c] >> j) & 1u) {
                if (off < cap) {
                    pixel[off] = (uint32_t)(e0 + j);
                    if (intensity) intensity[off] = (float)val[c].v[j];
                }
                ++off;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
k_compact(const T *__restrict__ heat, long long n, T thr, uint32_t *__restrict__ pixel,
          float *__restrict__ intensity, long long cap, unsigned long long *scratch, long long *d_count, long long *early_n,
          int aligned)
{
    compact_tile<T>(heat, n, thr, pixel, intensity, cap, scratch, d_count, early_n, aligned);
}

struct XfPack {
    FrameXf f[8];
};

// The compaction of dp_project: the same tiles, plus the call's resets and uploads (see launch_compact_fused).
// scratch: [0] tile ticket, [1 .. ntiles] look-back states, [ntiles + 1] finished tiles.
template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
k_compact_project(const T *__restrict__ heat, long long n, T thr, uint32_t *__restrict__ pixel,
                  float *__restrict__ intensity, long long cap, unsigned long long *scratch, long long *counts,
                  long long *early_n, int aligned, OrderState *ord_next, FrameXf *xf, int n_xf, const __grid_constant__ XfPack pack,
                  double *raytab, int H, int W)
{
    __shared__ int s_last;
    if (blockIdx.x == 0 && (int)threadIdx.x < 16 * n_xf) xf[threadIdx.x >> 4].v[threadIdx.x & 15] = pack.f[threadIdx.x >> 4].v[threadIdx.x & 15];
    if (raytab) {
        // the two per-pixel quotients of compute_rays (:216-217), once per column / row instead of once per ray
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W + H; i += gridDim.x * blockDim.x) {
            const double *c = pack.f[0].v;
            raytab[i] = i < W ? __ddiv_rn(__dsub_rn((double)i, c[2]), c[0]) : __ddiv_rn(__dsub_rn((double)(i - W), c[3]), c[1]);
        }
    }
    compact_tile<T>(heat, n, thr, pixel, intensity, cap, scratch, counts, early_n, aligned);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned *>(scratch + gridDim.x + 1), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // every tile has published, read and scattered: the scratch and the traversal's counters start the next phase at zero
    for (unsigned i = threadIdx.x; i < gridDim.x + 2; i += blockDim.x) scratch[i] = 0ull;
    if (threadIdx.x < 3) counts[1 + threadIdx.x] = 0;
    if (threadIdx.x == 3 && ord_next) {
        ord_next->n_valid = -1; ord_next->cost_sum = 0; ord_next->cnt[0] = 0; ord_next->cnt[1] = 0; ord_next->cnt[2] = 0;
    }
}

// per-frame counts from the sorted pixel list: count[f] = lb((f+1)*HW) - lb(f*HW)
__global__ void k_frame_counts(const uint32_t *__restrict__ pixel, const long long *d_count, long long cap,
                               long long frame_elems, long long nframes, long long *frame_count)
{
    const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    long long n = *d_count;
    if (n > cap) n = cap;
    long long bounds[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const unsigned long long key = (unsigned long long)(f + k) * (unsigned long long)frame_elems;
        long long lo = 0, hi = n;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if ((unsigned long long)pixel[mid] < key) lo = mid + 1; else hi = mid;
        }
        bounds[k] = lo;
    }
    frame_count[f] = bounds[1] - bounds[0];
}

// ---- prologue / epilogue of dp_project -------------------------------------------------------------------
// Everything a projection has to reset or upload goes through kernel launches, never through the copy engines:
// a cudaMemsetAsync / small cudaMemcpyAsync in the kernel stream queues behind the bulk H2D / D2H transfers of a
// pipelined caller (measured on B200: compaction 0.02 -> 0.1-0.3 ms, ray generation 0.02 -> 0.09-0.25 ms with a
// 12.6 MB copy in flight on another stream).
__global__ void __launch_bounds__(256)
k_project_prologue(unsigned long long *scratch, long long scratch_words, long long *counts, OrderState *ord_next,
                   FrameXf *xf, int n_xf, XfPack pack)
{
    for (long long i = threadIdx.x; i < scratch_words; i += blockDim.x) scratch[i] = 0ull;
    if (threadIdx.x < 4) counts[threadIdx.x] = 0;                 // rays, hits, the two traversal work counters
    if (threadIdx.x == 3 && ord_next) {
        ord_next->n_valid = -1; ord_next->cost_sum = 0; ord_next->cnt[0] = 0; ord_next->cnt[1] = 0; ord_next->cnt[2] = 0;
    }
    if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + 16 * n_xf) {
        const int k = threadIdx.x - 32;
        xf[k >> 4].v[k & 15] = pack.f[k >> 4].v[k & 15];
    }
}

__global__ void k_write_xf(FrameXf *xf, int n_xf, XfPack pack)
{
    const int k = threadIdx.x;
    if (k < 16 * n_xf) xf[k >> 4].v[k & 15] = pack.f[k >> 4].v[k & 15];
}

// counts -> a device address or a mapped pinned-host address (zero-copy store, visible to the host once the
// stream's next event / synchronisation completes)
__global__ void k_publish_counts(const long long *counts, long long *dst)
{
    dst[0] = counts[0];
    dst[1] = counts[1];
    __threadfence_system();
}

}  // namespace

cudaError_t launch_project_prologue(unsigned long long *scratch, int64_t n_elems, long long *counts, OrderState *ord_next,
                                    FrameXf *d_xf, const FrameXf *h_xf, int64_t n_xf, cudaStream_t s)
{
    XfPack pack;
    const int first = (int)(n_xf < 8 ? n_xf : 8);
    for (int i = 0; i < first; ++i) pack.f[i] = h_xf[i];
    const long long words = (long long)(compact_scratch_bytes(n_elems) / sizeof(unsigned long long));
    k_project_prologue<<<1, 256, 0, s>>>(scratch, words, counts, ord_next, d_xf, first, pack);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_xf > 8) {
        if (n_xf <= 128) {
            for (int64_t f0 = 8; f0 < n_xf; f0 += 8) {
                const int c = (int)(n_xf - f0 < 8 ? n_xf - f0 : 8);
                for (int i = 0; i < c; ++i) pack.f[i] = h_xf[f0 + i];
                k_write_xf<<<1, 128, 0, s>>>(d_xf + f0, c, pack);
            }
            e = cudaGetLastError();
        } else {
            e = cudaMemcpyAsync(d_xf + 8, h_xf + 8, (size_t)(n_xf - 8) * sizeof(FrameXf), cudaMemcpyHostToDevice, s);
        }
    }
    return e;
}

cudaError_t launch_compact_fused(const void *heat, int dtype, int64_t n_elems, double thr, uint32_t *pixel, float *intensity,
                                 int64_t cap, unsigned long long *scratch, long long *counts, OrderState *ord_next,
                                 FrameXf *d_xf, const FrameXf *h_xf, int n_xf, long long *early_n, cudaStream_t s, double *raytab,
                                 int H, int W)
{
    XfPack pack;
    memset(&pack, 0, sizeof(pack));
    for (int i = 0; i < n_xf && i < 8; ++i) pack.f[i] = h_xf[i];
    const int64_t ntiles = (n_elems + CT_TILE - 1) / CT_TILE;
    const int aligned = ((uintptr_t)heat % 16) == 0;
    if (dtype == 1)
        k_compact_project<double><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(static_cast<const double *>(heat), n_elems, thr, pixel,
                                                                           intensity, cap, scratch, counts, early_n, aligned,
                                                                           ord_next, d_xf, n_xf, pack, raytab, H, W);
    else
        k_compact_project<float><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(static_cast<const float *>(heat), n_elems, (float)thr,
                                                                          pixel, intensity, cap, scratch, counts, early_n,
                                                                          aligned, ord_next, d_xf, n_xf, pack, raytab, H, W);
    return cudaGetLastError();
}

cudaError_t launch_publish_counts(const long long *counts, long long *dst, cudaStream_t s)
{
    k_publish_counts<<<1, 1, 0, s>>>(counts, dst);
    return cudaGetLastError();
}

size_t compact_scratch_bytes(int64_t n_elems)
{
    const int64_t ntiles = (n_elems + CT_TILE - 1) / CT_TILE;
    return (size_t)(ntiles + 2) * sizeof(unsigned long long);
}

cudaError_t launch_compact(const void *heat, int dtype, int64_t n_elems, int64_t frame_elems, double thr,
                           uint32_t *pixel, float *intensity, int64_t cap, unsigned long long *scratch,
                           long long *d_count, long long *d_frame_count, int64_t nframes, cudaStream_t s, bool scratch_zeroed,
                           long long *early_n)
{
    cudaError_t e;
    if (n_elems <= 0) {
        if ((e = cudaMemsetAsync(d_count, 0, sizeof(long long), s)) != cudaSuccess) return e;
        if (d_frame_count && nframes > 0)
            if ((e = cudaMemsetAsync(d_frame_count, 0, sizeof(long long) * nframes, s)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    const int64_t ntiles = (n_elems + CT_TILE - 1) / CT_TILE;
    if (!scratch_zeroed && (e = cudaMemsetAsync(scratch, 0, compact_scratch_bytes(n_elems), s)) != cudaSuccess) return e;
    const int aligned = ((uintptr_t)heat % 16) == 0;
    if (dtype == 1) {
        k_compact<double><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(static_cast<const double *>(heat), n_elems, thr,
                                                                   pixel, intensity, cap, scratch, d_count, early_n, aligned);
    } else {
        k_compact<float><<<(unsigned)ntiles, CT_THREADS, 0, s>>>(static_cast<const float *>(heat), n_elems,
                                                                  (float)thr, pixel, intensity, cap, scratch,
                                                                  d_count, early_n, aligned);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (d_frame_count && nframes > 0) {
        k_frame_counts<<<(unsigned)((nframes + 127) / 128), 128, 0, s>>>(pixel, d_count, cap, frame_elems, nframes,
                                                                          d_frame_count);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace dp
