/*
 * defectproj.h -- C ABI of libdefectproj.so, the B200 (sm_100a) implementation of the
 * 2D-defect -> 3D-mesh back-projection hot path of
 *     /root/reference/src/defect_projection.py::ray_tracing            (:527-563)
 *
 * The reference has no native boundary on this path: it is Python that calls numpy and
 * open3d.t.geometry.RaycastingScene (Embree).  Each entry point below names the reference
 * lines it replaces; INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add to src/defect_projection.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  No torch / CUDA types in any signature:
 *     `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - every buffer argument is described by one `mem` flag per call:
 *     DP_HOST (pageable or pinned host memory; the call stages it) or DP_DEVICE (memory of
 *     the context's device; the call neither copies nor synchronises more than stated).
 *   - all functions return 0 (DP_OK) or a negative DP_E_* code; the message of the last
 *     failure of a context is returned by dp_last_error().  Nothing throws across the ABI.
 *   - a context is bound to one device and is NOT thread-safe (one context per host thread).
 *   - matrices are row-major doubles.  `pose` is the 4x4 rigid model->camera transform
 *     (what run.py:109-110 followed by src/defect_projection.py:549-550 apply to the mesh).
 *   - there is no CPU fallback: without a CUDA device dp_create() fails with DP_E_CUDA.
 */
#ifndef DEFECTPROJ_H
#define DEFECTPROJ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DP_API
#else
#define DP_API __attribute__((visibility("default")))
#endif

#define DP_ABI_VERSION 1

enum {
    DP_OK = 0,
    DP_E_ARG = -1,    /* bad argument (null pointer, negative size, index out of range ...) */
    DP_E_CUDA = -2,   /* a CUDA runtime call or kernel failed; see dp_last_error()          */
    DP_E_NOMEM = -3,  /* host or device allocation failed / caller capacity too small       */
    DP_E_STATE = -4   /* call order (no mesh, no BVH ...) or BVH too deep for the ray stack */
};

enum { DP_HOST = 0, DP_DEVICE = 1 };
enum { DP_F32 = 0, DP_F64 = 1 };

/* which mesh the rays are cast against */
enum {
    DP_FRAME_OBJECT = 0,  /* static BVH in the model frame; rays go through the inverse pose  */
    DP_FRAME_CAMERA = 1   /* BVH over the mesh posed by dp_pose_mesh(); rays leave (0,0,0)
                             exactly as the reference builds them (:247-251, :545)          */
};

typedef struct dp_ctx dp_ctx;

/* per-ray outputs of dp_project / dp_cast_rays; any pointer may be NULL (not wanted).
 * All arrays are indexed by compacted ray number (row-major pixel order, :165-179). */
typedef struct dp_rays_out {
    uint32_t *pixel;   /* [cap]   frame*H*W + y*W + x        (heatmap_to_points, :175-179)   */
    float *intensity;  /* [cap]   heatmap value as float32                                   */
    float *t_hit;      /* [cap]   +inf on miss                (cast_rays 't_hit', :258-259)  */
    int32_t *face;     /* [cap]   original face id, -1 miss   (cast_rays 'primitive_ids')    */
    float *point;      /* [cap*3] hit point, camera frame     (:261-263); NaN on miss        */
    double *point64;   /* [cap*3] the same in float64: d (float64) * t (float32), as :261-263 */
    int64_t cap;       /* capacity in rays of every non-NULL array above                     */
    int64_t *counts;   /* optional [2], device or PINNED host memory: receives (n_rays, n_hits) by
                          an asynchronous store on the call's stream (for pipelined callers).  In
                          pinned host memory counts[0] is ALSO stored by the compaction kernel as
                          soon as the ray count exists, long before the traversal ends: a caller
                          that set counts[0] = -1 can poll it and size follow-up work (the slice
                          gather of a ray-sharded frame) while the traversal still runs           */
} dp_rays_out;

typedef struct dp_stats {
    int64_t rays;           /* rays traced by the last counted launch                        */
    int64_t hits;
    int64_t nodes_fetched;  /* wide nodes fetched, 80 or 208 bytes each (sum over rays)      */
    int64_t tris_tested;    /* 48-byte triangle records fetched (sum over rays)              */
    int64_t n_wide_nodes;   /* BVH size                                                      */
    int64_t n_tris;
    int32_t wide_depth;
    int32_t reserved;
    float last_build_ms;    /* device time of the last dp_build_bvh (CUDA events)            */
    float last_refit_ms;    /* device time of the last dp_pose_mesh                          */
} dp_stats;

/* ---- life cycle ------------------------------------------------------------------------ */
DP_API int dp_abi_version(void);
DP_API int dp_create(int device, dp_ctx **out);
DP_API void dp_destroy(dp_ctx *ctx);
DP_API const char *dp_last_error(const dp_ctx *ctx);   /* ctx may be NULL: last create error */
DP_API int dp_synchronize(dp_ctx *ctx, void *stream);

/* ---- mesh + BVH  (replaces RaycastingScene() + add_triangles, :245, :253-254) ----------- */
/* V: [nV*3] model-frame vertices, float32 or float64 (`vdtype`).  The object-frame BVH uses
 * the float32 rounding of V (the cast of from_legacy, :245); dp_pose_mesh transforms the
 * vertices at the precision given here, as the reference does with its float64 mesh.
 * F: [nF*3] int32.  The arrays are copied; the caller may free them on return. */
DP_API int dp_set_mesh(dp_ctx *ctx, const void *V, int vdtype, int64_t nV, const int32_t *F, int64_t nF, int mem,
                       void *stream);
/* LBVH over the current object-frame mesh: 30-bit Morton codes, radix sort, Karras hierarchy,
 * surface-area-guided collapse into compressed 8-wide nodes.  Asynchronous apart from one
 * 8-byte read-back per tree level. */
DP_API int dp_build_bvh(dp_ctx *ctx, void *stream);
/* New vertex positions for the mesh of dp_set_mesh (same count, type and faces): the reference hands ray_tracing a
 * freshly transformed copy of the same model on every call (run.py:109-110) and rebuilds its scene from it
 * (:253-254); here the hierarchy keeps its topology and is refitted in place to the new vertices (any valid
 * hierarchy gives the same closest hits; the Morton order is that of the vertices dp_build_bvh saw). */
DP_API int dp_update_vertices(dp_ctx *ctx, const void *V, int vdtype, int64_t nV, int mem, void *stream);
/* Camera-frame copy of the mesh: v' = float32(T * (double)v) (:549-550 then :245), followed by
 * a bottom-up refit of a second set of wide nodes that shares the object-frame topology.
 * T: 16 doubles on the HOST, row-major. */
DP_API int dp_pose_mesh(dp_ctx *ctx, const double *T, void *stream);
/* copy the camera-frame vertices of the last dp_pose_mesh out: [nV*3]; DP_F32 gives the
 * float32 values the BVH holds, DP_F64 the float64 values before that cast (:550). */
DP_API int dp_get_posed_vertices(dp_ctx *ctx, void *V, int vdtype, int mem, void *stream);

/* ---- H1: heatmap threshold + order-preserving compaction (:165-179) ---------------------- */
/* heat: [nframes*H*W] of `dtype`; strict '>' as np.where(heatmap > thr); NaN never passes.
 * pixel/intensity: [cap]; *n (HOST) receives the total count (the call synchronises on it).
 * frame_count: optional [nframes] int64 HOST array receiving the per-frame counts. */
DP_API int dp_compact(dp_ctx *ctx, const void *heat, int dtype, int64_t nframes, int H, int W, double thr,
                      uint32_t *pixel, float *intensity, int64_t cap, int64_t *n, int64_t *frame_count,
                      int mem, void *stream);

/* ---- H2: host helper shared with the tests ---------------------------------------------- */
/* xf[16] = fx fy cx cy | Rinv row-major | tinv for intrinsics K[9] and a rigid pose[16];
 * Rinv = R^T, tinv_k = -((R0k*t0 + R1k*t1) + R2k*t2).  pose == NULL gives the identity. */
DP_API void dp_frame_xform(const double *K, const double *pose, double *xf);

/* compute_rays (:196-223): rays3[i] = normalize(((xs[i]-cx)/fx, (ys[i]-cy)/fy, 1)) in float64,
 * K: 9 HOST doubles; xs, ys: [n] int32; rays3: [n*3] float64. */
DP_API int dp_compute_rays(dp_ctx *ctx, const int32_t *xs, const int32_t *ys, int64_t n, const double *K,
                           double *rays3, int mem, void *stream);

/* ---- H4: closest hit on explicit rays (replaces cast_rays, :256) ------------------------- */
/* rays6: [n*6] float32 (ox,oy,oz,dx,dy,dz) as the reference builds them (:247-251).
 * t_hit: [n] (+inf on miss); face: [n] or NULL. */
DP_API int dp_cast_rays(dp_ctx *ctx, int frame, const float *rays6, int64_t n, float *t_hit, int32_t *face,
                        int mem, void *stream);

/* ---- H1+H2+H4+H6+H7 fused: one call per batch of frames (replaces the body of ray_tracing) */
/* heat: [nframes*H*W]; K: [nK*9] HOST doubles, nK == 1 or nframes; pose: [nframes*16] HOST
 * doubles (ignored, may be NULL, when frame == DP_FRAME_CAMERA: the mesh is already posed).
 * accumulate != 0 adds every hit to the context's per-face / per-vertex accumulators.
 * n_rays / n_hits: HOST, optional; when n_rays != NULL the call synchronises on the stream
 * and, for mem == DP_HOST, copies only the first *n_rays entries of each array back. */
DP_API int dp_project(dp_ctx *ctx, int frame, const void *heat, int dtype, int64_t nframes, int H, int W,
                      double thr, const double *K, int64_t nK, const double *pose, int accumulate,
                      dp_rays_out *out, int64_t *n_rays, int64_t *n_hits, int mem, void *stream);

/* ---- depth-image projection path (src/defect_projection.py:359-492, :613-649) ---------- */
/* heatmap_to_point3d (:359-395): intensity = heat/max(heat); pixels with intensity > thr and depth > 0, in
 * row-major order, back-projected with their depth: points4[i] = ((x-cx)*d/fx, (y-cy)*d/fy, d*0.98, intensity),
 * all float64.  heat: [H*W] of `dtype`; depth: [Hd*Wd] uint16 (pixels outside the depth image are skipped);
 * K: 9 HOST doubles; points4: [cap*4]; *n (HOST) = number of selected pixels (the call synchronises). */
DP_API int dp_depth_backproject(dp_ctx *ctx, const void *heat, int dtype, int H, int W, const uint16_t *depth, int Hd,
                                int Wd, const double *K, double thr, double *points4, int64_t cap, int64_t *n, int mem,
                                void *stream);
/* calc_coordinates (:462-492): picked pixels -> ((x-cx)*d/fx, (y-cy)*d/fy, d); valid[i] = depth > 0 (the
 * reference skips the others).  xs, ys: [n] int32; out3: [n*3] float64; valid: [n] bytes. */
DP_API int dp_calc_coordinates(dp_ctx *ctx, const int32_t *xs, const int32_t *ys, int64_t n, const uint16_t *depth, int Hd,
                               int Wd, const double *K, double *out3, unsigned char *valid, int mem, void *stream);
/* align_to_surface (:417-460): for every query point (first 3 of `stride` doubles) the nearest target point
 * (exact, float64, ties to the smaller index; replaces KDTreeFlann.search_knn_vector_3d(p, 1)):
 * aligned = target[idx], offset_points = target[idx] + normals[idx]*offset.  Any output may be NULL. */
DP_API int dp_align_to_surface(dp_ctx *ctx, const double *query, int stride, int64_t n, const double *target,
                               const double *normals, int64_t m, double offset, double *offset_points,
                               double *aligned_points, int32_t *idx, int mem, void *stream);

/* ---- before the path: heatmap preparation (datareader.py:639-675) ---------------------- */
/* DataReader.get_heatmap: data [src_h*src_w] of `dtype` is min-max normalised ((d - min) / max(d - min), in the
 * array's own type), resized to o x o with o = min(H, W) exactly as cv2.resize(..., INTER_LINEAR) of the
 * opencv-python wheel computes it for CV_32F / CV_64F, and centred in a zero H x W frame (`heatmap_full`).
 * out: [H*W] of `out_dtype` (the reference returns float64; DP_F32 stores the float32 rounding for dp_project). */
DP_API int dp_prepare_heatmap(dp_ctx *ctx, const void *data, int dtype, int src_h, int src_w, int H, int W, void *out,
                              int out_dtype, int mem, void *stream);

/* ---- after the path: viewer payload (src/defect_projection.py:259-294, src/web_vis.py:203-217) */
/* PointCloud.transform (run.py:118, :196-200) in place: p' = (T*[p,1])[:3] / w, float64,
 * ((T0*x + T1*y) + T2*z) + T3 per row.  points3: [n*3]; T: 16 HOST doubles, row-major. */
DP_API int dp_transform_points(dp_ctx *ctx, double *points3, int64_t n, const double *T, int mem, void *stream);
/* Selection of the rays that hit (face[i] >= 0; face == NULL selects all), in ray order, with the colours of
 * create_intersection_pcd (:286-291): jet((I - min) / (max - min))[:, :3] over the SELECTED intensities
 * (matplotlib's 256-entry 'jet' table; max == min gives (0,0,0), the table's NaN colour) and, when T != NULL,
 * the rigid transform above applied to the hit points.  intensity: [n] of `dtype`; pixel, point64: per-ray
 * arrays as dp_rays_out (may be NULL).  Outputs [cap] / [cap*3], any may be NULL.  *m (HOST) = number selected
 * (the call synchronises). */
DP_API int dp_pack_hits(dp_ctx *ctx, const void *intensity, int dtype, const int32_t *face, const uint32_t *pixel,
                        const double *point64, int64_t n, const double *T, double *points, double *colors,
                        int32_t *face_out, uint32_t *pixel_out, double *intensity_out, int64_t cap, int64_t *m, int mem,
                        void *stream);
/* the colour table used above: lut [256*3] float64 RGB (host helper, no device work) */
DP_API void dp_jet_lut(double *lut);

/* o3d.geometry.PointCloud.estimate_normals(search_param=KDTreeSearchParamHybrid(radius, max_nn)) as called by
 * src/defect_projection.py:181-186, :431-436 (inside align_to_surface) and src/pose_estimation.py:301-306:
 * points [n*3] float64; neighbours of a point = its max_nn (<= 64) nearest points, itself included, with squared
 * distance < radius^2, ordered by (distance, index); fewer than 3 -> (0, 0, 1) (or the existing normal); else the
 * eigenvector of the smallest eigenvalue of the neighbours' covariance (Open3D's closed-form FastEigen3x3).
 * normals [n*3] float64: output; with has_normals != 0 also input, and every new normal keeps the side of the old
 * one.  neighbours: optional [n] int32 neighbour counts.  Open3D is absent offline: parity unpinned. */
DP_API int dp_estimate_normals(dp_ctx *ctx, const double *points, int64_t n, double radius, int max_nn, double *normals,
                               int has_normals, int32_t *neighbours, int mem, void *stream);

/* ---- upstream of the path: point-to-plane ICP (src/pose_estimation.py:505-522, :577-613, :654-660) */
/* o3d.pipelines.registration.registration_icp(source, target, max_correspondence_distance, init,
 * TransformationEstimationPointToPlane(), ICPConvergenceCriteria(relative_fitness, relative_rmse, max_iteration)):
 * source [n*3], target [m*3] with normals [m*3], float64; init: 16 HOST doubles (NULL = identity).
 * Per iteration the device finds the exact nearest target point of every source point (ties to the smaller index)
 * within the distance, and reduces the 6 x 6 normal equations in a fixed order; the host solves them and composes
 * update = [Rz Ry Rx | t].  Outputs (HOST): T_out[16] (source -> target), fitness = |corr| / n, inlier_rmse,
 * iterations performed; correspondence: optional [n] int32 (target index or -1) of the final alignment.
 * Open3D defaults: max_iteration 30, relative_fitness = relative_rmse = 1e-6. */
DP_API int dp_icp_point_to_plane(dp_ctx *ctx, const double *source, int64_t n, const double *target,
                                 const double *target_normals, int64_t m, double max_correspondence_distance,
                                 const double *init, int max_iteration, double relative_fitness, double relative_rmse,
                                 double *T_out, double *fitness, double *inlier_rmse, int *iterations,
                                 int32_t *correspondence, int mem, void *stream);

/* ---- H6/H7: accumulators (extensions named by north_star) ------------------ */
DP_API int dp_accum_reset(dp_ctx *ctx, void *stream);
/* hist: [nF] int32, fmax: [nF] float32, vmax: [nV] float32; any may be NULL */
DP_API int dp_accum_get(dp_ctx *ctx, int32_t *hist, float *fmax, float *vmax, int mem, void *stream);
/* The per-vertex maxima are derived from the per-face maxima (vmax[v] = max over the faces incident to v of
 * fmax[f]) when they are read, not per hit: dp_accum_get does it itself; call dp_accum_flush on the stream before
 * reading vmax through dp_accum_device_ptrs. */
DP_API int dp_accum_flush(dp_ctx *ctx, void *stream);
/* device addresses of the accumulators, for in-place NCCL reductions by the host layer */
DP_API int dp_accum_device_ptrs(dp_ctx *ctx, int32_t **hist, float **fmax, float **vmax);
/* The three accumulators live in ONE allocation (hist | fmax | vmax, 256-byte aligned starts, the gaps stay zero):
 * `base` + offsets, `bytes` in total.  A host layer snapshots the whole block with one copy and combines
 * [fmax_off, bytes) of several devices with ONE max-reduction (non-negative floats order like their bits). */
DP_API int dp_accum_layout(dp_ctx *ctx, void **base, int64_t *hist_off, int64_t *fmax_off, int64_t *vmax_off, int64_t *bytes);

/* ---- multi-GPU: single-frame ray sharding and hit records ------------------------------------
 * The reference casts one frame's rays in one call (/root/reference/src/defect_projection.py:247-256); with the mesh
 * and BVH replicated per GPU, that call splits by rays.  After dp_set_ray_shard(rank, world) every dp_project of this
 * context still compacts the whole frame (the ray list and its order are global) but generates, traces and accumulates
 * only block `rank` of `world` of the compacted list: a contiguous range of output slots [lo, hi) -- whole 4-row bands
 * when the dense frame is walked in 8x4 tiles -- which dp_shard_slots returns for any (rank, world) and a frame of n_rays selected pixels.
 * n_rays of dp_project stays the frame's total, n_hits counts the shard.  Slots outside the range are not written.
 * Concatenating the ranks' ranges gives the 1-GPU arrays bit for bit; summing / maxing their accumulators likewise.
 * (0, 1) restores whole frames. */
DP_API int dp_set_ray_shard(dp_ctx *ctx, int rank, int world);
DP_API int dp_shard_slots(int rank, int world, int64_t n_rays, int64_t nframes, int H, int W, int64_t *lo, int64_t *hi);
/* Compacted hit records of the rays [first, first + n) of a projection's per-ray outputs (device memory): the rays with
 * face >= 0, in ray order, as rows of 3 uint32 (pixel, t_hit bits, face) or -- when `point` (float32 [.,3]) is given --
 * 6 uint32 (+ x, y, z bits): the unit a rank contributes to the hit gather, and the viewer's input
 * (/root/reference/src/defect_projection.py:259-264 keeps exactly these rays).  pixel may be NULL (the ray index is
 * stored).  *m receives the count (the call then synchronises); m_async (device or pinned host memory) receives it
 * asynchronously on the stream.  DP_DEVICE only. */
DP_API int dp_pack_records(dp_ctx *ctx, const uint32_t *pixel, const float *t_hit, const int32_t *face, const float *point,
                           int64_t n, int64_t first, uint32_t *records, int64_t cap, int64_t *m, int64_t *m_async, int mem,
                           void *stream);

/* ---- multi-GPU: result exchange over peer-mapped memory (NVLink / NVSwitch) ---------------------
 * One process per GPU on one node.  The data path has no exchange step; what travels are RESULTS: per batch the
 * accumulator block (hist SUM, fmax | vmax MAX) and the compacted hit records, per ray-sharded frame the slices of the
 * per-ray arrays.  Instead of collectives (two all-reduces, a count exchange with a host read-back and a padded
 * all-gather per batch; one all-gather per array and frame) every rank maps every other rank's EXCHANGE WINDOW
 *     [control | accumulator snapshot x2 | hit records x2 | per-ray results x2 (t_hit f32, face i32, point f32 x3)]
 * and the library's own kernels read / write the peers' memory directly:
 *   dp_peer_combine   ONE launch per batch: flag barrier, then every rank folds the snapshots of all ranks into its
 *                     running totals straight out of the peers' memory, and rank `gather_root` pulls all ranks' hit
 *                     records into one array in rank order (the counts are read on the device, never on the host);
 *   dp_project        after dp_peer_results(slot): the traversal stores its slice of t_hit / face / point into the result
 *                     arrays of EVERY rank as it produces them (the gather is the kernel's epilogue), then one tiny
 *                     barrier launch; when it has run, this rank's result slot holds the whole frame.
 * Integer sums and float maxima do not depend on the order: N GPUs give the 1-GPU result bit for bit.
 * All ranks must issue the same sequence of dp_peer_combine calls, and of ray-sharded dp_project calls (collective
 * semantics); slots alternate 0, 1, 0, ... per channel, which is what makes reuse safe without a second barrier.  A wait
 * that sees no progress for 4 s gives up and sets the error word (dp_peer_status) instead of hanging the GPU. */
#define DP_PEER_HANDLE_BYTES 64
enum { DP_PEER_STAGE = 0, DP_PEER_RECORDS = 1, DP_PEER_REC_COUNT = 2, DP_PEER_T_HIT = 3, DP_PEER_FACE = 4, DP_PEER_POINT = 5 };
/* Allocate this context's window (snapshots sized by the current mesh, `record_bytes` per record slot, `result_rays`
 * rays per result slot; either may be 0) and write its CUDA IPC handle (DP_PEER_HANDLE_BYTES) to `handle` (HOST; may be
 * NULL for dp_peer_open_local).  The caller exchanges the handles (e.g. one 64-byte all-gather at start-up). */
DP_API int dp_peer_export(dp_ctx *ctx, int64_t record_bytes, int64_t result_rays, void *handle, int64_t *window_bytes);
/* Map the windows of all ranks: handles [world * DP_PEER_HANDLE_BYTES] in rank order (the own entry is ignored).
 * Every rank must have exported with the same sizes.  Fails with DP_E_CUDA when the devices have no peer access. */
DP_API int dp_peer_open(dp_ctx *ctx, int rank, int world, const void *handles);
/* The same for contexts of ONE process (one thread driving several GPUs, or several contexts on one GPU in the tests):
 * peers [world] contexts in rank order. */
DP_API int dp_peer_open_local(dp_ctx *ctx, int rank, int world, dp_ctx *const *peers);
DP_API int dp_peer_close(dp_ctx *ctx);
/* device address / size of a region of the OWN window: DP_PEER_STAGE (snapshot `slot`), DP_PEER_RECORDS (record slot; pass
 * it as `records` to dp_pack_records), DP_PEER_REC_COUNT (int64 row count of the slot; pass it as `m_async`),
 * DP_PEER_T_HIT / FACE / POINT (result slot: the whole frame after a ray-sharded dp_project) */
DP_API int dp_peer_window(dp_ctx *ctx, int what, int slot, void **ptr, int64_t *bytes);
/* the accumulator block (per-vertex maxima brought up to date first) -> snapshot `slot`; reset != 0 zeroes the live block */
DP_API int dp_peer_snapshot(dp_ctx *ctx, int slot, int reset, void *stream);
/* total: device int32 [accumulator block words] running totals of this rank (+= / max= the snapshots `slot` of all ranks;
 * NULL skips the fold).  gathered (device, on rank gather_root only; gather_root < 0: nobody gathers): [cap_rows *
 * row_words] uint32 rows, the record slots `slot` of all ranks in rank order; m_async (device or pinned host): the total
 * row count (rows beyond cap_rows are dropped).  Asynchronous on `stream`. */
DP_API int dp_peer_combine(dp_ctx *ctx, int slot, void *total, int gather_root, uint32_t *gathered, int64_t cap_rows,
                           int row_words, int64_t *m_async, void *stream);
/* slot 0 / 1: the following ray-sharded dp_project calls (dp_set_ray_shard with world > 1, DP_DEVICE, no t_hit / face /
 * point in dp_rays_out) keep their per-ray results in result slot `slot` of the windows of ALL ranks; -1: off. */
DP_API int dp_peer_results(dp_ctx *ctx, int slot, int with_points);
/* *error: 0, or 1 + channel of a wait of this rank that timed out (synchronises with the device) */
DP_API int dp_peer_status(dp_ctx *ctx, int *error);

/* ---- instrumentation -------------------------------------------------------------------- */
/* enable != 0: the next traversal launches use the counting variant of the kernel */
DP_API int dp_set_stats(dp_ctx *ctx, int enable);
DP_API int dp_get_stats(dp_ctx *ctx, dp_stats *out);
/* device time (ms, CUDA events on the caller's stream) of the stages of the last dp_project made with timing on:
 * [0] H2D (if any) + compaction, [1] ray generation, [2] traversal + accumulation kernel,
 * [3] the whole call on the device (including the hit-point kernel) */
DP_API int dp_last_timings(dp_ctx *ctx, float *ms4);
/* enable != 0: dp_project brackets its stages with CUDA events for dp_last_timings.  Default: DISABLED -- the five
 * event records cost 14 us per call on a B200 (0.336 -> 0.322 ms per 1M-ray frame); dp_last_timings fails with
 * DP_E_STATE for calls made while timing is off. */
DP_API int dp_set_timing(dp_ctx *ctx, int enable);
/* structural dump of the wide BVH for the tests: nodes [n*80 bytes], triangle records
 * [nt*48 bytes] (v0.xyz, face id bits, v1.xyz, 0, v2.xyz, 0).  Pass NULL to query sizes. */
DP_API int dp_debug_dump_bvh(dp_ctx *ctx, int frame, void *nodes, int64_t *n_nodes, void *tris,
                             int64_t *n_tris);
/* nodes fetched by each ray of the last dp_project launched with stats enabled: counts [n] HOST */
DP_API int dp_debug_ray_nodes(dp_ctx *ctx, uint32_t *counts, int64_t n);
/* the in-house radix sort on its own, for the structural tests: sorts (key,value) pairs
 * of HOST arrays in place, stable, ascending by key. */
DP_API int dp_debug_radix_sort(dp_ctx *ctx, uint32_t *keys, uint32_t *vals, int64_t n);
/* Morton codes of the current mesh's triangles in input order: codes [nF] HOST */
DP_API int dp_debug_morton(dp_ctx *ctx, uint32_t *codes);

#ifdef __cplusplus
}
#endif
#endif /* DEFECTPROJ_H */
