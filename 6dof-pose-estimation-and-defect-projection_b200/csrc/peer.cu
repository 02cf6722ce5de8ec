// Multi-GPU result combination over peer-mapped device memory (NVLink 5 / NVSwitch), one process per GPU on one node.
//
// The data path of the projection has no exchange step (rays and frames are independent, mesh + BVH are
// replicated); only RESULTS are combined.  With every rank's exchange window mapped into every other rank's address
// space (CUDA IPC, or plain pointers for contexts of one process) that combination needs no collective library:
//
//   k_peer_combine      one launch per batch and rank: a flag barrier over the peers' control blocks, then the rank
//                       reads every rank's accumulator snapshot (hist | fmax | vmax, one block) straight out of the
//                       peers' memory and folds it into its running totals (integer SUM over the histogram words,
//                       unsigned MAX over the float bits of the maxima: both order-independent, so N GPUs give the
//                       1-GPU result bit for bit), and the gathering rank pulls every rank's compacted hit records into
//                       one array in rank order -- the counts are read from the peers' control blocks inside the kernel,
//                       so the host never learns (or waits for) them.  Replaces two all-reduces, one count exchange
//                       with a host read-back and one padded all-gather.
//   k_peer_frame_done   the end of a ray-sharded frame whose traversal (trace.cu, PeerOut) stored its slice of the
//                       per-ray results into EVERY rank's result arrays as it produced them: signal + wait, after which
//                       every rank holds the whole frame.  Replaces one all-gather per result array.
//
// Synchronisation: rank r signals peer p by a release store (system scope) of the call's epoch into p's
// flag[channel][r] and waits with acquire loads on its own flag[channel][p].  Epochs only grow; every rank issues the
// same sequence of calls per channel (collective semantics).  The exchanged buffers are double-buffered by the caller
// (slot = call parity): a rank that has passed the barrier of call e+1 knows that every peer has finished call e, so
// slot (e & 1) may be overwritten for call e+2.  A wait that sees no progress for PEER_TIMEOUT_NS sets the window's
// error word and gives up (the host reads it back: dp_peer_status) instead of hanging the GPU.
#include "dp_internal.cuh"

#include <stdlib.h>

namespace dp {
namespace {

constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_sys_s64(const long long *p)
{
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One warp: lane p talks to rank p.  `signal`: this rank's arrival is published to the peers first (one warp per
// launch does that); every calling warp then waits until all peers have arrived at `epoch`.  False on a timeout.
__device__ bool peer_signal_wait(const PeerView &pv, int chan, unsigned long long epoch, bool signal)
{
    const int lane = threadIdx.x & 31;
    bool ok = true;
    if (lane < pv.world && lane != pv.rank) {
        PeerCtl *mine = reinterpret_cast<PeerCtl *>(pv.win[pv.rank]);
        if (signal) {
            __threadfence_system();
            st_release_sys(&reinterpret_cast<PeerCtl *>(pv.win[lane])->flag[chan][pv.rank], epoch);
        }
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(&mine->flag[chan][lane]) < epoch) {
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > PEER_TIMEOUT_NS) {
                ok = false;
                atomicExch(&mine->error, 1u + (unsigned)chan);
                break;
            }
        }
    }
    return __all_sync(0xffffffffu, ok);
}

__device__ __forceinline__ uint4 fold4(uint4 a, uint4 b, bool is_max)
{
    if (is_max) return make_uint4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
    return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Accumulator block -> this rank's snapshot slot (in its exchange window); the live block is zeroed for the next batch.
__global__ void __launch_bounds__(256) k_peer_snapshot(uint4 *__restrict__ live, uint4 *__restrict__ stage, long long quads, int reset)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += stride) {
        stage[q] = live[q];
        if (reset) live[q] = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <int WORLD_UNROLL>
__device__ __forceinline__ void fold_range(const PeerView &pv, int slot, uint4 *__restrict__ total, long long quads, long long max_from_q)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += stride) {
        uint4 v[WORLD_UNROLL];
#pragma unroll
        for (int p = 0; p < WORLD_UNROLL; ++p)
            if (p < pv.world) v[p] = ld_relaxed_sys_v4(reinterpret_cast<const uint4 *>(pv.win[p] + pv.stage_off[slot]) + q);
        const bool is_max = q >= max_from_q;
        uint4 a = total[q];
#pragma unroll
        for (int p = 0; p < WORLD_UNROLL; ++p)
            if (p < pv.world) a = fold4(a, v[p], is_max);
        total[q] = a;
    }
}

__global__ void __launch_bounds__(256)
k_peer_combine(PeerView pv, int slot, unsigned long long epoch, uint4 *__restrict__ total, long long quads, long long max_from_q,
               int fold, int gather_root, uint32_t *__restrict__ gathered, long long cap_rows, int row_words, long long *m_out,
               long long *m_async)
{
    __shared__ int s_ok;
    __shared__ long long s_off[PEER_MAX + 1];
    if (threadIdx.x < 32) {
        const bool ok = peer_signal_wait(pv, 0, epoch, blockIdx.x == 0);
        if (threadIdx.x == 0) s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return;
    // ---- every rank: totals (+)= / max= the snapshots of all ranks, read in place over NVLink
    if (fold) {
        if (pv.world <= 2) fold_range<2>(pv, slot, total, quads, max_from_q);
        else if (pv.world <= 4) fold_range<4>(pv, slot, total, quads, max_from_q);
        else if (pv.world <= 8) fold_range<8>(pv, slot, total, quads, max_from_q);
        else fold_range<PEER_MAX>(pv, slot, total, quads, max_from_q);
    }
    // ---- the gathering rank: every rank's hit records, in rank order, unpadded
    if (pv.rank != gather_root || gathered == nullptr) return;
    if (threadIdx.x == 0) {
        long long off = 0;
        for (int p = 0; p < pv.world; ++p) {
            s_off[p] = off;
            long long c = ld_relaxed_sys_s64(&reinterpret_cast<const PeerCtl *>(pv.win[p])->rec_count[slot]);
            if (c < 0) c = 0;
            off += c;
        }
        s_off[pv.world] = off;
        if (blockIdx.x == 0) {
            if (m_out) *m_out = off;
            if (m_async) *m_async = off;
        }
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x, tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (int p = 0; p < pv.world; ++p) {
        long long first = s_off[p], rows = s_off[p + 1] - s_off[p];
        if (first >= cap_rows) break;
        if (first + rows > cap_rows) rows = cap_rows - first;             // the caller sees the total in m_out and knows
        const long long words = rows * row_words;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(pv.win[p] + pv.rec_off[slot]);
        uint32_t *dst = gathered + first * row_words;
        const long long nq = words >> 2;
        // 16-byte loads from the (aligned) source window; the destination offset is a multiple of the row size only
        for (long long q = tid; q < nq; q += stride) {
            const uint4 v = ld_relaxed_sys_v4(reinterpret_cast<const uint4 *>(src) + q);
            dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
        }
        for (long long w = (nq << 2) + tid; w < words; w += stride) dst[w] = ld_relaxed_sys_u32(src + w);
    }
}

__global__ void k_peer_frame_done(PeerView pv, unsigned long long epoch)
{
    peer_signal_wait(pv, 1, epoch, true);
}

}  // namespace

cudaError_t launch_peer_snapshot(void *live, void *stage, size_t bytes, bool reset, cudaStream_t s)
{
    const long long quads = (long long)(bytes >> 4);
    if (quads <= 0) return cudaSuccess;
    long long blocks = (quads + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_peer_snapshot<<<(unsigned)blocks, 256, 0, s>>>(static_cast<uint4 *>(live), static_cast<uint4 *>(stage), quads, reset ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_peer_combine(const PeerView &pv, int slot, unsigned long long epoch, void *total, size_t bytes, size_t max_from,
                                bool fold, int gather_root, uint32_t *gathered, int64_t cap_rows, int row_words, long long *m_out,
                                long long *m_async, cudaStream_t s)
{
    // enough CTAs to keep a few MB of peer loads in flight; they share the SMs with whatever else runs.  Every CTA waits
    // in the barrier, so contexts that emulate several ranks on ONE device (the tests) must keep all their grids
    // co-resident: DP_PEER_BLOCKS bounds the grid there.
    const char *e = getenv("DP_PEER_BLOCKS");             // read per launch: the tests set it per case
    int blocks = e ? atoi(e) : 0;
    if (blocks < 1 || blocks > 148 * 2) blocks = 148 * 2;
    k_peer_combine<<<blocks, 256, 0, s>>>(pv, slot, epoch, static_cast<uint4 *>(total), (long long)(bytes >> 4),
                                           (long long)(max_from >> 4), fold ? 1 : 0, gather_root, gathered, (long long)cap_rows,
                                           row_words, m_out, m_async);
    return cudaGetLastError();
}

cudaError_t launch_peer_frame_done(const PeerView &pv, unsigned long long epoch, cudaStream_t s)
{
    k_peer_frame_done<<<1, 32, 0, s>>>(pv, epoch);
    return cudaGetLastError();
}

}  // namespace dp
