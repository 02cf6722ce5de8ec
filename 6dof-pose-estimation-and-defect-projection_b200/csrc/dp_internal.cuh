// Internal declarations shared by the translation units of libdefectproj.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dp {

// ------------------------------------------------------------------------------------------
// Wide BVH node: 80 bytes = five 16-byte words (loaded as 5 x LDG.128).
//   w0: origin x,y,z (float bits) | ex | ey<<8 | ez<<16 | imask<<24
//   w1: child_base, tri_base, meta[0..3], meta[4..7]
//   w2: qlo_x[0..7], qlo_y[0..7]            (x,y = slots 0..3 / 4..7 of x; z,w = y)
//   w3: qlo_z[0..7], qhi_x[0..7]
//   w4: qhi_y[0..7], qhi_z[0..7]
// meta[slot]: 0 = empty; inner child = 0x20 | (24 + slot); leaf = (unary count) << 5 | offset,
//             unary count = 1, 3, 7 for 1, 2, 3 triangles, offset = first record relative to
//             tri_base (0..23).  e* are biased float exponents: scale = bits(e << 23).
// ------------------------------------------------------------------------------------------
struct __align__(16) WideNode {
    uint4 w[5];
};
static_assert(sizeof(WideNode) == 80, "wide node must be 80 bytes");

// triangle record: 48 bytes = 3 x float4.  v0.w carries the original face id bits.
struct __align__(16) TriRec {
    float4 v0, v1, v2;
};
static_assert(sizeof(TriRec) == 48, "triangle record must be 48 bytes");

constexpr int LEAF_MAX = 3;           // triangles per leaf slot
#ifndef DP_STACK_SMEM
#define DP_STACK_SMEM 10
#endif
constexpr int STACK_SMEM = DP_STACK_SMEM;   // per-ray stack entries kept in shared memory
constexpr int STACK_LOCAL = 64 - STACK_SMEM; // overflow entries in local memory
constexpr int MAX_WIDE_DEPTH = STACK_SMEM + STACK_LOCAL - 2;
constexpr float T_SLACK = 1.0001f;    // culling bound = best * T_SLACK (same as oracle.c)

// ------------------------------------------------------------------------------------------
// Uncompressed twin of the wide node for hierarchies that stay resident in L2: 208 bytes = thirteen 16-byte words.
//   q0      : child_base, tri_base, valid, 0
//             valid bits 3s..3s+2 = the (<= 3) triangle records of leaf slot s (record = tri_base + number of valid
//             bits below), bits 24+s = inner children (child = child_base + number of inner bits below)
//   q1..q6  : lo_x[0..7], lo_y[0..7], lo_z[0..7]   (float32 child planes, empty slot = +inf)
//   q7..q12 : hi_x[0..7], hi_y[0..7], hi_z[0..7]   (empty slot = -inf)
// Same node numbering, same record array, same topology as the compressed set (both are written by fit_node8).  A
// visit needs no byte->float conversion (48 quarter-rate I2F per visit in the compressed format), no per-node scale
// and a constant hit word per slot; the price is 2.6x the bytes per node, paid in L1/L2 hits while the set fits L2.
// ------------------------------------------------------------------------------------------
constexpr int FAT_QUADS = 13;
constexpr size_t FAT_NODE_BYTES = FAT_QUADS * 16;
constexpr size_t FAT_MAX_BYTES = 96u << 20;   // fat nodes + records above this: the compressed set is traced instead

struct BvhView {
    const WideNode *nodes;
    const TriRec *tris;
    const float *d_scale;  // device: max |vertex coordinate| of the mesh the BVH was fitted to
    size_t bytes;          // nodes + triangle records
    const uint4 *fat;      // uncompressed node set (nullptr: not kept for this hierarchy)
};

// per-frame ray generation constants: fx fy cx cy | Rinv (row-major) | tinv
struct FrameXf {
    double v[16];
};

struct Accum {
    int32_t *hist;     // [nF]
    uint32_t *fmax;    // [nF] float bits (non-negative floats order like unsigned ints)
    uint32_t *vmax;    // [nV]
    const int32_t *F;  // [nF*3]
};

// Packet schedule learnt by one traversal launch for the next one over the same rays (device memory).
struct OrderState {
    long long n_valid;               // ray count the lists were built for (-1 = none)
    unsigned long long cost_sum;     // sum over packets of the longest ray (node steps)
    unsigned cnt[3];                 // entries in list0 / list1 / list2
    unsigned pad_;
    uint32_t *list0, *list1;         // packets above 2.5x / 1.5x the mean cost: traced first
    uint32_t *list2;                 // packets below 0.75x the mean cost: traced last (the tail of the launch is made of
                                     // short packets: simulated on the benchmark frame's measured costs, SM-busy 0.90 -> 0.96)
    unsigned char *flags;            // [packets] 1 = the packet is in one of the lists
};

struct TraceStats {
    unsigned long long rays, hits, nodes, tris;
    unsigned *ray_nodes;   // optional [n]: nodes fetched by each ray (debug)
};

// ---- decoupled look-back shared by the order-preserving selections (compact.cu, depth.cu, prep.cu) --------
// state[tile] holds (flag << 62 | value): flag 1 = the tile's own count is published, 2 = its inclusive prefix is.
// Called by ALL 32 lanes of one warp of the tile with the tile's count; returns the exclusive prefix of the tile
// (the same value in every lane) after publishing the tile's inclusive prefix.  Tiles take their id from an atomic
// ticket, so a predecessor a lane waits for is always already running.  32 predecessors are inspected per round.
#ifdef __CUDACC__
constexpr unsigned long long LB_AGG = 1ull << 62, LB_INC = 2ull << 62, LB_VAL = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long lookback_exclusive_prefix(unsigned long long *state, unsigned tile,
                                                                        unsigned long long total, int lane)
{
    // LB_WIDE predecessors per lane and round.  Measured on B200 (1024^2 pixels = 256 tiles, 100-step bench, same box):
    // 4 per lane (2 rounds instead of 8 for the last tile) is 2 us per frame SLOWER than 1 -- the rounds overlap with the
    // tiles' own loads, the extra loads and selects do not -- so 1 stays (scripts/r2/r2_sweep18.sh).
#ifndef DP_LB_WIDE
#define DP_LB_WIDE 1
#endif
    constexpr int LB_WIDE = DP_LB_WIDE;
    unsigned long long prefix = 0;
    if (tile == 0) {
        if (lane == 0) atomicExch(&state[0], LB_INC | total);
        return 0;
    }
    if (lane == 0) atomicExch(&state[tile], LB_AGG | total);
    long long j0 = (long long)tile - 1;                      // lane l looks at tiles j0 - LB_WIDE*l - k, k = 0 .. LB_WIDE-1
    for (;;) {
        unsigned long long sv[LB_WIDE];
#pragma unroll
        for (int k = 0; k < LB_WIDE; ++k) {
            const long long j = j0 - (long long)(LB_WIDE * lane + k);
            sv[k] = LB_INC;                                  // "tiles" before the first one: inclusive prefix 0
            if (j >= 0) sv[k] = *reinterpret_cast<volatile unsigned long long *>(&state[j]);
        }
#pragma unroll
        for (int k = 0; k < LB_WIDE; ++k) {
            const long long j = j0 - (long long)(LB_WIDE * lane + k);
            while ((sv[k] >> 62) == 0) sv[k] = *reinterpret_cast<volatile unsigned long long *>(&state[j]);
        }
        // values up to and including the nearest predecessor that carries an inclusive prefix
        unsigned long long v = 0;
        bool closed = false;
#pragma unroll
        for (int k = 0; k < LB_WIDE; ++k) {
            if (!closed) v += sv[k] & LB_VAL;
            closed = closed || (sv[k] & LB_INC) != 0;
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, closed);
        const int first_inc = __ffs(inc_mask) - 1;           // first lane (nearest tiles) holding an inclusive prefix
        if (first_inc >= 0 && lane > first_inc) v = 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        prefix += v;
        if (first_inc >= 0) break;
        j0 -= 32 * LB_WIDE;
    }
    if (lane == 0) atomicExch(&state[tile], LB_INC | (prefix + total));
    return prefix;
}
#endif

// ---- compact.cu ----------------------------------------------------------------------------
// tile_state: [ntiles+1] u64 scratch (zeroed by the call); d_count: device int64 total;
// d_frame_count: [nframes] device int64 (zeroed by the call) or nullptr.
size_t compact_scratch_bytes(int64_t n_elems);
cudaError_t launch_compact(const void *heat, int dtype, int64_t n_elems, int64_t frame_elems, double thr,
                           uint32_t *pixel, float *intensity, int64_t cap, unsigned long long *scratch,
                           long long *d_count, long long *d_frame_count, int64_t nframes, cudaStream_t s,
                           bool scratch_zeroed = false, long long *early_n = nullptr);
// dp_project's compaction launch doing the call's resets and uploads itself (no prologue launch): block 0 writes the
// per-frame constants (n_xf <= 8), the tile that finishes LAST zeroes counts[1..3] (hits, the traversal's two work
// counters), resets `ord_next` and leaves the look-back scratch clean for the next call (the scratch must be zero on
// entry: zero it once when it is allocated).  n_elems > 0.
cudaError_t launch_compact_fused(const void *heat, int dtype, int64_t n_elems, double thr, uint32_t *pixel, float *intensity,
                                 int64_t cap, unsigned long long *scratch, long long *counts, struct OrderState *ord_next,
                                 struct FrameXf *d_xf, const struct FrameXf *h_xf, int n_xf, long long *early_n, cudaStream_t s,
                                 double *raytab = nullptr, int H = 0, int W = 0);
// raytab [W + H] doubles (optional, all frames share h_xf[0]'s intrinsics): xn(x) = (x - cx) / fx for every column, then
// yn(y) = (y - cy) / fy for every row, written by the first blocks of the compaction for the traversal's ray generation
// dp_project's resets and uploads as kernel launches (no copy-engine work in the kernel stream): zeroes the
// compaction scratch, counts[0..2] (rays, hits, traversal work counter), resets `ord_next`, writes the per-frame
// constants.  launch_publish_counts stores counts[0..1] to a device or mapped pinned-host address.
cudaError_t launch_project_prologue(unsigned long long *scratch, int64_t n_elems, long long *counts, struct OrderState *ord_next,
                                    struct FrameXf *d_xf, const struct FrameXf *h_xf, int64_t n_xf, cudaStream_t s);
cudaError_t launch_publish_counts(const long long *counts, long long *dst, cudaStream_t s);

// ---- trace.cu ------------------------------------------------------------------------------
// single-frame ray sharding: this context traces block `rank` of `world` of every frame's compacted ray list
struct RayShard {
    int rank = 0, world = 1;
};
// output slots [lo, hi) of the shard for a frame of n selected pixels (same partition as the kernels)
void shard_slots_host(int64_t n, int64_t total_px, int H, int W, RayShard shard, int64_t *lo, int64_t *hi);
// rays from pixels: pixel[i] = frame*HW + y*W + x ; xf[frame] ; n read from *d_n (<= n_max)
cudaError_t launch_raygen(const uint32_t *pixel, const long long *d_n, int64_t n_max, int H, int W, const FrameXf *xf,
                          int64_t n_xf, float4 *dir4, cudaStream_t s, int64_t total_px = 0, RayShard shard = RayShard());
cudaError_t launch_points(const uint32_t *pixel, const float *t_hit, const long long *d_n, int64_t n_max, int H, int W,
                          const FrameXf *xf, int64_t n_xf, float *point, double *point64, cudaStream_t s,
                          int64_t total_px = 0, RayShard shard = RayShard());
// dir4[2i] = origin (object frame), dir4[2i+1] = (direction, frame index bits) of compacted ray i, written by launch_raygen
cudaError_t launch_trace_pixels(const BvhView &bvh, const float4 *dir4, const float *intensity, const long long *d_n,
                                int64_t n_max, int64_t total_px, int H, int W, const FrameXf *xf, float *t_hit,
                                int32_t *face, const Accum *acc, unsigned long long *work_counter, long long *d_hits,
                                TraceStats *stats, const OrderState *ord_prev, OrderState *ord_next, cudaStream_t s,
                                bool counter_zeroed = false, RayShard shard = RayShard(), const uint32_t *pixel = nullptr,
                                int64_t n_xf = 0, float *point = nullptr, double *point64 = nullptr,
                                const struct PeerOut *peer_out = nullptr, const double *raytab = nullptr);
// dir4 == nullptr: the traversal generates the rays itself from pixel[] / xf[] and writes the hit points (point, point64)
// per-vertex maxima from the per-face maxima (run when the accumulators are read, not per hit)
cudaError_t launch_vertex_max(const uint32_t *fmax, const int32_t *F, int64_t nF, uint32_t *vmax, cudaStream_t s);
cudaError_t launch_compute_rays(const int32_t *xs, const int32_t *ys, int64_t n, const FrameXf &xf, double *rays3,
                                cudaStream_t s);
cudaError_t launch_trace_rays6(const BvhView &bvh, const float *rays6, int64_t n, float *t_hit, int32_t *face,
                               unsigned long long *work_counter, TraceStats *stats, cudaStream_t s);

// ---- peer.cu: result combination over peer-mapped memory (NVLink / NVSwitch) ---------------------------------
constexpr int PEER_MAX = 16;           // ranks of one exchange (one node)
// Head of every rank's exchange window.  flag[c][r]: the last epoch rank r signalled on channel c (0 = batch combine,
// 1 = ray-sharded frame), written by rank r over NVLink; rec_count[k]: rows in record slot k (written by the owner's
// k_pack_records); error: set by a wait of this rank that timed out.
struct PeerCtl {
    unsigned long long flag[2][PEER_MAX];
    long long rec_count[2];
    unsigned error;
    unsigned pad_[59];
};
static_assert(sizeof(PeerCtl) == 512, "control block is 512 bytes");
// The windows of all ranks as mapped in THIS process (win[rank] = the own one) and the common layout:
// [PeerCtl | snapshot 0 | snapshot 1 | records 0 | records 1 | results 0 | results 1], results k = t_hit f32 [res_cap] |
// face i32 [res_cap] | point f32 [res_cap*3].
struct PeerView {
    char *win[PEER_MAX];
    int rank, world;
    unsigned long long stage_off[2], rec_off[2], res_off[2];
    unsigned long long res_cap;
};
// per-ray result arrays of the OTHER ranks (n entries) for a traversal that stores its slice everywhere
struct PeerOut {
    float *t_hit[PEER_MAX];
    int32_t *face[PEER_MAX];
    float *point[PEER_MAX];
    int n;
};
cudaError_t launch_peer_snapshot(void *live, void *stage, size_t bytes, bool reset, cudaStream_t s);
cudaError_t launch_peer_combine(const PeerView &pv, int slot, unsigned long long epoch, void *total, size_t bytes, size_t max_from,
                                bool fold, int gather_root, uint32_t *gathered, int64_t cap_rows, int row_words, long long *m_out,
                                long long *m_async, cudaStream_t s);
cudaError_t launch_peer_frame_done(const PeerView &pv, unsigned long long epoch, cudaStream_t s);

// ---- depth.cu ------------------------------------------------------------------------------
size_t depth_select_scratch_bytes(int64_t n_elems);
cudaError_t launch_depth_select(const void *heat, int dtype, int H, int W, const uint16_t *depth, int Hd, int Wd,
                                double thr, const double *K, double *out4, int64_t cap, unsigned long long *scratch,
                                long long *d_count, double *d_max, int *d_nan, cudaStream_t s);
cudaError_t launch_calc_coordinates(const int32_t *xs, const int32_t *ys, int64_t n, const uint16_t *depth, int Hd, int Wd,
                                    const double *K, double *out3, unsigned char *valid, cudaStream_t s);
cudaError_t launch_nearest(const double *q, int stride, int64_t n, const double *target, int64_t m, const double *normals,
                           double offset, int32_t *idx, double *aligned, double *offset_pts, cudaStream_t s);
struct IcpGridView;
cudaError_t launch_nearest_grid(const double *q, int stride, int64_t n, const IcpGridView &gv, bool has_normals, double offset,
                                int32_t *idx, double *aligned, double *offset_pts, cudaStream_t s);

// ---- prep.cu -------------------------------------------------------------------------------
// mm: 4 device long longs (ordered max, ordered min, NaN flag, count)
void jet_lut_host(double *lut768);
cudaError_t launch_prepare_heatmap(const void *data, int dtype, int sh, int sw, int H, int W, void *out, int out_dtype,
                                   long long *mm, cudaStream_t s);
cudaError_t launch_transform_points(double *p, int64_t n, const double *T_host, cudaStream_t s);
size_t pack_scratch_bytes(int64_t n);
cudaError_t launch_pack_hits(const void *inten, int dtype, const int32_t *face, const uint32_t *pixel, const double *point64,
                             int64_t n, const double *T_host, const double *lut, long long *mm, double *points, double *colors,
                             int32_t *face_out, uint32_t *pixel_out, double *inten_out, int64_t cap,
                             unsigned long long *scratch, long long *d_count, cudaStream_t s);

cudaError_t launch_pack_records(const uint32_t *pixel, const float *t_hit, const int32_t *face, const float *point, int64_t n,
                                int64_t first, uint32_t *rec, int64_t cap, unsigned long long *scratch, long long *d_count,
                                long long *m_async, cudaStream_t s);

// ---- icp.cu --------------------------------------------------------------------------------
constexpr int ICP_GRID_MIN_POINTS = 256;   // smaller targets are scanned (one shared-memory tile)
constexpr int ICP_GRID_MIN_CELLS = 216;
constexpr int ICP_GRID_MAX_DIM = 96;   // cells per axis of the target grid (2 x 3.5 MB of cell ranges at most)
struct IcpGrid {
    double lo[3];
    double inv_cell, cell;
    int dim[3];
};
struct IcpGridView {
    IcpGrid grid;
    double ext[3];
    const double *tps, *tns;          // target points / normals in cell order
    const uint32_t *orig;             // cell order -> original target index
    const int32_t *cell_start, *cell_end;
};
// cell coordinate of v along one axis, clamped into the grid (NaN -> 0)
__device__ __forceinline__ int grid_coord(double v, double lo, double inv_cell, int dim)
{
    const double c = floor((v - lo) * inv_cell);
    if (!(c >= 0.0)) return 0;                       // below the grid, or NaN (never handed to the conversion)
    return c >= (double)dim ? dim - 1 : (int)c;
}
size_t icp_grid_bytes(int64_t m);
cudaError_t icp_grid_build(const double *tp, const double *tn, int64_t m, double max_dist, size_t min_cells, void *scratch,
                           IcpGridView *out, cudaStream_t s);
cudaError_t launch_icp_step_grid(double *src, int64_t n, const IcpGridView &gv, double max_dist, const double *update_host,
                                 int32_t *corr, double *partial, double *sums, cudaStream_t s);
size_t icp_partial_doubles(int64_t n);
// sums (device, 29 doubles): count, sum d^2, 21 upper-triangle entries of sum J J^T, 6 entries of sum J r
cudaError_t launch_icp_step(double *src, int64_t n, const double *tp, const double *tn, int64_t m, double max_dist,
                            const double *update_host, int32_t *corr, double *partial, double *sums, cudaStream_t s);

// ---- normals.cu ----------------------------------------------------------------------------
constexpr int NORMALS_MAX_NN = 64;
cudaError_t launch_estimate_normals(const IcpGridView &gv, int64_t n, double radius, int max_nn, double *normals, int has_normals,
                                    int32_t *neighbours, cudaStream_t s);

// ---- build.cu ------------------------------------------------------------------------------
struct BuildScratch;   // opaque, owned by the context

struct BvhStorage {
    WideNode *nodes = nullptr;     // [cap_nodes]
    TriRec *tris = nullptr;        // [nF]
    float *wlo = nullptr;          // [cap_nodes*3] exact wide-node boxes (refit)
    float *whi = nullptr;
    uint4 *fat = nullptr;          // [cap_nodes*FAT_QUADS] uncompressed twin (optional)
    int64_t cap_nodes = 0;
    int64_t n_nodes = 0;
    int64_t n_tris = 0;
    float *d_scale = nullptr;      // device scalar: max |vertex coordinate|
};

struct Topology {                  // shared by the object-frame and camera-frame node sets
    int32_t *tri_face = nullptr;   // [nF] record -> original face id
    int32_t *wparent = nullptr;    // [cap_nodes] wide node -> parent wide node (root: unused)
    unsigned *arrived = nullptr;   // [cap_nodes] fit arrivals per node (monotonic, see k_fit_all)
    int64_t level_begin[128];      // wide nodes of level l are [level_begin[l], level_begin[l+1])
    int n_levels = 0;
};

// all launches on `s`; performs one small blocking read-back per tree level
cudaError_t build_lbvh(const float *V, int64_t nV, const int32_t *F, int64_t nF, BvhStorage &out, Topology &topo,
                       void **scratch, size_t *scratch_bytes, uint32_t *morton_out_host, cudaStream_t s);
// v' = float32(T*v) for every vertex, then new triangle records and a bottom-up refit of
// `dst` that copies the topology (meta, bases) of `src`.
// V: float32 (vdtype 0) or float64 (vdtype 1); Vposed64 may be nullptr.
cudaError_t pose_and_refit(const void *V, int vdtype, int64_t nV, const int32_t *F, int64_t nF, const double *T_host,
                           float *Vposed, double *Vposed64, const BvhStorage &src, BvhStorage &dst,
                           const Topology &topo, cudaStream_t s);
cudaError_t convert_f64_to_f32(const double *src, float *dst, int64_t n, cudaStream_t s);
cudaError_t check_faces(const int32_t *F, int64_t nF, int64_t nV, int *d_flag, cudaStream_t s);
cudaError_t radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, int64_t n,
                             uint32_t *table, cudaStream_t s, bool *result_in_tmp, int key_bits = 32);
size_t radix_table_entries(int64_t n);

}  // namespace dp
