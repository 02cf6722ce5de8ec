// C ABI of libdefectproj.so (see include/defectproj.h): context, buffers, call sequencing.
#include "../../include/defectproj.h"
#include "dp_internal.cuh"

#include <math.h>
#include <stdio.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

using namespace dp;

namespace {
thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t need)
    {
        if (need <= cap) return cudaSuccess;
        size_t want = need + need / 4 + 256;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, need);   // retry without head-room
            want = need;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return static_cast<T *>(p); }
};
}  // namespace

struct dp_ctx {
    int device = 0;
    std::string err;

    // mesh (object frame) and its camera-frame copy
    DevBuf V, F, Vposed, V64, Vposed64;
    int vdtype = 0;
    int64_t nV = 0, nF = 0;
    bool has_mesh = false, has_bvh = false, has_cam = false;

    BvhStorage obj, cam;
    DevBuf obj_nodes, obj_tris, obj_wlo, obj_whi, cam_nodes, cam_tris, cam_wlo, cam_whi, scales, tri_face, wparent, arrived, obj_fat, cam_fat;
    Topology topo;
    void *build_scratch = nullptr;
    size_t build_scratch_bytes = 0;

    // accumulators: ONE allocation, hist | fmax | vmax with 256-byte aligned starts (the gaps stay zero), so that a
    // caller can combine fmax and vmax of several contexts with one MAX reduction over [fmax, vmax + nV)
    DevBuf accum;
    struct View {
        void *p = nullptr;
        template <typename T>
        T *as() const { return static_cast<T *>(p); }
    } hist, fmax, vmax;
    size_t accum_bytes = 0;
    RayShard shard;                  // dp_set_ray_shard

    // exchange over peer-mapped memory (peer.cu): the own window, the mapped windows of all ranks, call epochs
    DevBuf peer_win, peer_out;
    PeerView peer{};
    bool peer_exported = false, peer_opened = false;
    bool peer_ipc[PEER_MAX] = {};
    size_t peer_stage_bytes = 0, peer_rec_bytes = 0, peer_win_bytes = 0;
    unsigned long long peer_epoch[2] = {0, 0};
    int peer_res_slot = -1;          // dp_peer_results: result slot the next ray-sharded dp_project stores into (-1: off)
    bool peer_res_points = false;

    // per-call scratch
    DevBuf heat, pixel, inten, t_hit, face, point, point64, rays6, dir4, ray_nodes, order, cost, tmp[8], cscratch, pscratch, raytab, counts, fcounts, xf, stats, jet;
    long long *h_counts = nullptr;   // pinned: [0] rays, [1] hits
    const void *counts_alias_src = nullptr;   // last dp_rays_out.counts address and its device alias
    long long *counts_alias = nullptr;
    bool stats_on = false;
    int order_parity = 0;
    size_t order_np = 0;
    int64_t ray_nodes_n = 0;
    dp_stats last_stats{};
    cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool timings_valid = false;
    bool timing_on = false;          // stage events inside dp_project (dp_set_timing): 5 records = 14 us per call
    float build_ms = 0.f, refit_ms = 0.f;
};

namespace {

int fail(dp_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else
        snprintf(buf, sizeof(buf), "%s", what);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call, what)                                                      \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess)                                             \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? DP_E_NOMEM : DP_E_CUDA, what, e__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

BvhView view_of(const BvhStorage &b)
{
    // the uncompressed twin is traced while it and the records fit L2 together
    const bool fat_ok = b.fat != nullptr && (size_t)b.n_nodes * FAT_NODE_BYTES + (size_t)b.n_tris * sizeof(TriRec) <= FAT_MAX_BYTES;
    return BvhView{b.nodes, b.tris, b.d_scale, (size_t)b.n_nodes * sizeof(WideNode) + (size_t)b.n_tris * sizeof(TriRec),
                   fat_ok ? b.fat : nullptr};
}

}  // namespace

extern "C" {

int dp_abi_version(void) { return DP_ABI_VERSION; }

void dp_frame_xform(const double *K, const double *pose, double *xf)
{
    static const double ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const double *P = pose ? pose : ident;
    xf[0] = K[0]; xf[1] = K[4]; xf[2] = K[2]; xf[3] = K[5];
    // Rinv = R^T
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) xf[4 + 3 * r + c] = P[4 * c + r];
    const double t0 = P[3], t1 = P[7], t2 = P[11];
    for (int k = 0; k < 3; ++k) {
        volatile double a = P[0 * 4 + k] * t0;
        volatile double b = P[1 * 4 + k] * t1;
        volatile double c = P[2 * 4 + k] * t2;
        volatile double s = a + b;
        s = s + c;
        xf[13 + k] = (s == 0.0) ? 0.0 : -s;      // no negative zero: identity pose gives origin +0
    }
}

int dp_create(int device, dp_ctx **out)
{
    if (!out) return fail(nullptr, DP_E_ARG, "dp_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, DP_E_CUDA, "dp_create: no CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= ndev) return fail(nullptr, DP_E_ARG, "dp_create: device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, DP_E_CUDA, "dp_create", e);
    if (prop.major < 10)
        return fail(nullptr, DP_E_CUDA, "dp_create: device is not sm_100 class; the kernels are built for sm_100a only");
    dp_ctx *ctx = new (std::nothrow) dp_ctx;
    if (!ctx) return fail(nullptr, DP_E_NOMEM, "dp_create: out of host memory");
    ctx->device = device;
    DeviceGuard g(device);
    for (int i = 0; i < 10; ++i)
        if ((e = cudaEventCreate(&ctx->ev[i])) != cudaSuccess) { delete ctx; return fail(nullptr, DP_E_CUDA, "dp_create: event", e); }
    if ((e = cudaMallocHost(reinterpret_cast<void **>(&ctx->h_counts), 64)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, DP_E_CUDA, "dp_create: pinned alloc", e);
    }
    if ((e = ctx->counts.ensure(64)) != cudaSuccess || (e = ctx->stats.ensure(sizeof(TraceStats))) != cudaSuccess ||
        (e = ctx->scales.ensure(64)) != cudaSuccess) {
        dp_destroy(ctx);
        return fail(nullptr, DP_E_CUDA, "dp_create: alloc", e);
    }
    cudaMemset(ctx->scales.p, 0, 64);
    cudaMemset(ctx->counts.p, 0xff, 64);      // counts[3] = -1: no learnt packet order yet
    *out = ctx;
    return DP_OK;
}

void dp_destroy(dp_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    dp_peer_close(ctx);
    ctx->peer_win.release();
    ctx->peer_out.release();
    DevBuf *bufs[] = {&ctx->V, &ctx->F, &ctx->Vposed, &ctx->V64, &ctx->Vposed64, &ctx->obj_nodes, &ctx->obj_tris, &ctx->obj_wlo, &ctx->obj_whi,
                      &ctx->cam_nodes, &ctx->cam_tris, &ctx->cam_wlo, &ctx->cam_whi, &ctx->obj_fat, &ctx->cam_fat, &ctx->scales, &ctx->tri_face, &ctx->wparent, &ctx->arrived,
                      &ctx->accum, &ctx->heat, &ctx->pixel, &ctx->inten, &ctx->t_hit,
                      &ctx->face, &ctx->point, &ctx->point64, &ctx->rays6, &ctx->dir4, &ctx->ray_nodes, &ctx->order, &ctx->cost, &ctx->cscratch, &ctx->pscratch, &ctx->raytab, &ctx->counts, &ctx->fcounts, &ctx->xf,
                      &ctx->stats, &ctx->jet};
    for (DevBuf *b : bufs) b->release();
    for (DevBuf &b : ctx->tmp) b.release();
    if (ctx->build_scratch) cudaFree(ctx->build_scratch);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    for (int i = 0; i < 10; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    delete ctx;
}

const char *dp_last_error(const dp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dp_synchronize(dp_ctx *ctx, void *stream)
{
    if (!ctx) return DP_E_ARG;
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)), "dp_synchronize");
    return DP_OK;
}

int dp_set_mesh(dp_ctx *ctx, const void *V, int vdtype, int64_t nV, const int32_t *F, int64_t nF, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (nV < 0 || nF < 0 || (nV > 0 && !V) || (nF > 0 && !F) || (vdtype != DP_F32 && vdtype != DP_F64))
        return fail(ctx, DP_E_ARG, "dp_set_mesh: bad arguments");
    // the traversal's triangle queue packs (owner lane << 27 | record index): 2^27 - 1 triangles at most
    if (nF >= (1LL << 27) || nV > 0x7fffffffLL / 4) return fail(ctx, DP_E_ARG, "dp_set_mesh: mesh too large (at most 2^27 - 1 triangles)");
    if (mem == DP_HOST) {
        for (int64_t i = 0; i < 3 * nF; ++i)
            if (F[i] < 0 || F[i] >= nV) return fail(ctx, DP_E_ARG, "dp_set_mesh: face index out of range");
    }
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t nv = (size_t)(nV > 0 ? nV : 1), nf = (size_t)(nF > 0 ? nF : 1);
    CK(ctx->V.ensure(nv * 12), "dp_set_mesh: V");
    CK(ctx->F.ensure(nf * 12), "dp_set_mesh: F");
    CK(ctx->Vposed.ensure(nv * 12), "dp_set_mesh: Vposed");
    if (vdtype == DP_F64) {
        CK(ctx->V64.ensure(nv * 24), "dp_set_mesh: V64");
        CK(ctx->Vposed64.ensure(nv * 24), "dp_set_mesh: Vposed64");
    }
    {
        const size_t af = (nf * 4 + 255) & ~size_t(255), av = (nv * 4 + 255) & ~size_t(255);
        CK(ctx->accum.ensure(2 * af + av), "dp_set_mesh: accumulators");
        ctx->hist.p = ctx->accum.p;
        ctx->fmax.p = ctx->accum.as<char>() + af;
        ctx->vmax.p = ctx->accum.as<char>() + 2 * af;
        if (ctx->peer_exported && ctx->peer_stage_bytes != 2 * af + av)
            return fail(ctx, DP_E_STATE, "dp_set_mesh: the exchange window was sized for another mesh (dp_peer_close, then export again)");
        ctx->accum_bytes = 2 * af + av;
    }
    const cudaMemcpyKind kind = mem == DP_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (nV) {
        if (vdtype == DP_F64) {
            CK(cudaMemcpyAsync(ctx->V64.p, V, (size_t)nV * 24, kind, s), "dp_set_mesh: copy V");
            CK(convert_f64_to_f32(ctx->V64.as<double>(), ctx->V.as<float>(), nV * 3, s), "dp_set_mesh: convert V");
        } else {
            CK(cudaMemcpyAsync(ctx->V.p, V, (size_t)nV * 12, kind, s), "dp_set_mesh: copy V");
        }
    }
    if (nF) CK(cudaMemcpyAsync(ctx->F.p, F, (size_t)nF * 12, kind, s), "dp_set_mesh: copy F");
    CK(cudaMemsetAsync(ctx->accum.p, 0, ctx->accum_bytes, s), "dp_set_mesh: zero");
    if (mem == DP_HOST) CK(cudaStreamSynchronize(s), "dp_set_mesh: sync");
    else if (nF) {
        // a device-resident mesh cannot be validated on the host: one pass over the indices, one 4-byte read-back
        int *d_flag = reinterpret_cast<int *>(ctx->counts.as<long long>() + 5);
        CK(check_faces(ctx->F.as<int32_t>(), nF, nV, d_flag, s), "dp_set_mesh: index check");
        CK(cudaMemcpyAsync(ctx->h_counts + 5, d_flag, 4, cudaMemcpyDeviceToHost, s), "dp_set_mesh: index check");
        CK(cudaStreamSynchronize(s), "dp_set_mesh: index check");
        if (*reinterpret_cast<int *>(ctx->h_counts + 5)) {
            ctx->has_mesh = false;
            return fail(ctx, DP_E_ARG, "dp_set_mesh: face index out of range");
        }
    }
    ctx->nV = nV;
    ctx->nF = nF;
    ctx->vdtype = vdtype;
    ctx->has_mesh = true;
    ctx->has_bvh = ctx->has_cam = false;
    return DP_OK;
}

int dp_build_bvh(dp_ctx *ctx, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_build_bvh: no mesh (call dp_set_mesh first)");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t nf = (size_t)(ctx->nF > 0 ? ctx->nF : 1);
    // a wide node is rooted at an internal binary node: typically ~nF/7 of them, nF/2 when every one is rooted at a
    // subtree of > LEAF_MAX triangles, never more than nF - 1 (second attempt)
    cudaError_t e = cudaSuccess;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t cap_nodes = attempt == 0 ? nf / 2 + 64 : nf + 64;
        CK(ctx->obj_nodes.ensure(cap_nodes * sizeof(WideNode)), "dp_build_bvh: nodes");
        CK(ctx->obj_tris.ensure(nf * sizeof(TriRec)), "dp_build_bvh: tris");
        CK(ctx->obj_wlo.ensure(cap_nodes * 12), "dp_build_bvh: boxes");
        CK(ctx->obj_whi.ensure(cap_nodes * 12), "dp_build_bvh: boxes");
        CK(ctx->tri_face.ensure(nf * 4), "dp_build_bvh: tri_face");
        CK(ctx->wparent.ensure(cap_nodes * 4), "dp_build_bvh: topology");
        CK(ctx->arrived.ensure(cap_nodes * 4), "dp_build_bvh: topology");
        // uncompressed twin of the node set: only meshes whose records leave room for it in L2 (a wide node per ~8.5
        // triangles: the twin of a 1.3M-triangle mesh is ~32 MB next to 62 MB of records)
        const bool want_fat = nf * sizeof(TriRec) + (nf / 8) * FAT_NODE_BYTES <= FAT_MAX_BYTES;
        ctx->obj.fat = nullptr;
        if (want_fat) {
            CK(ctx->obj_fat.ensure(cap_nodes * FAT_NODE_BYTES), "dp_build_bvh: uncompressed nodes");
            ctx->obj.fat = ctx->obj_fat.as<uint4>();
        }
        ctx->obj.nodes = ctx->obj_nodes.as<WideNode>();
        ctx->obj.tris = ctx->obj_tris.as<TriRec>();
        ctx->obj.wlo = ctx->obj_wlo.as<float>();
        ctx->obj.whi = ctx->obj_whi.as<float>();
        ctx->obj.cap_nodes = (int64_t)cap_nodes;
        ctx->obj.d_scale = ctx->scales.as<float>();
        ctx->topo.tri_face = ctx->tri_face.as<int32_t>();
        ctx->topo.wparent = ctx->wparent.as<int32_t>();
        ctx->topo.arrived = ctx->arrived.as<unsigned>();
        ctx->has_bvh = ctx->has_cam = false;
        CK(cudaEventRecord(ctx->ev[8], s), "dp_build_bvh");
        e = build_lbvh(ctx->V.as<float>(), ctx->nV, ctx->F.as<int32_t>(), ctx->nF, ctx->obj, ctx->topo,
                       &ctx->build_scratch, &ctx->build_scratch_bytes, nullptr, s);
        if (e != cudaErrorMemoryAllocation) break;
        cudaGetLastError();
    }
    CK(e, "dp_build_bvh");
    if (ctx->topo.n_levels < 0) return fail(ctx, DP_E_STATE, "dp_build_bvh: hierarchy deeper than 126 levels");
    CK(cudaEventRecord(ctx->ev[9], s), "dp_build_bvh");
    CK(cudaEventSynchronize(ctx->ev[9]), "dp_build_bvh: kernels");
    cudaEventElapsedTime(&ctx->build_ms, ctx->ev[8], ctx->ev[9]);
    if (ctx->topo.n_levels > MAX_WIDE_DEPTH) return fail(ctx, DP_E_STATE, "dp_build_bvh: BVH too deep for the ray stack");
    ctx->has_bvh = true;
    return DP_OK;
}

int dp_update_vertices(dp_ctx *ctx, const void *V, int vdtype, int64_t nV, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_update_vertices: no mesh (call dp_set_mesh first)");
    if (nV != ctx->nV || vdtype != ctx->vdtype || (nV > 0 && !V))
        return fail(ctx, DP_E_ARG, "dp_update_vertices: vertex count and type must be those of dp_set_mesh");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const cudaMemcpyKind kind = mem == DP_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (nV) {
        if (vdtype == DP_F64) CK(cudaMemcpyAsync(ctx->V64.p, V, (size_t)nV * 24, kind, s), "dp_update_vertices: copy V");
        else CK(cudaMemcpyAsync(ctx->V.p, V, (size_t)nV * 12, kind, s), "dp_update_vertices: copy V");
    }
    ctx->has_cam = false;
    if (ctx->has_bvh) {
        // the object-frame hierarchy keeps its topology and is fitted again, in place, to the new vertices (the
        // "no transform" pose pass rewrites the float32 vertex copy and the scene scale)
        CK(pose_and_refit(vdtype == DP_F64 ? ctx->V64.p : ctx->V.p, vdtype, nV, ctx->F.as<int32_t>(), ctx->nF, nullptr,
                          ctx->V.as<float>(), nullptr, ctx->obj, ctx->obj, ctx->topo, s),
           "dp_update_vertices: refit");
    } else if (nV && vdtype == DP_F64) {
        CK(convert_f64_to_f32(ctx->V64.as<double>(), ctx->V.as<float>(), nV * 3, s), "dp_update_vertices: convert V");
    }
    if (mem == DP_HOST) CK(cudaStreamSynchronize(s), "dp_update_vertices: sync");
    return DP_OK;
}

int dp_pose_mesh(dp_ctx *ctx, const double *T, void *stream)
{
    if (!ctx || !T) return fail(ctx, DP_E_ARG, "dp_pose_mesh: bad arguments");
    if (!ctx->has_bvh) return fail(ctx, DP_E_STATE, "dp_pose_mesh: no BVH (call dp_build_bvh first)");
    if (T[12] != 0.0 || T[13] != 0.0 || T[14] != 0.0 || T[15] != 1.0)
        return fail(ctx, DP_E_ARG, "dp_pose_mesh: last row of the transform must be 0 0 0 1");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t nf = (size_t)(ctx->nF > 0 ? ctx->nF : 1);
    const size_t cap_nodes = (size_t)ctx->obj.cap_nodes;
    CK(ctx->cam_nodes.ensure((size_t)(ctx->obj.n_nodes + 1) * sizeof(WideNode)), "dp_pose_mesh: nodes");
    CK(ctx->cam_tris.ensure(nf * sizeof(TriRec)), "dp_pose_mesh: tris");
    CK(ctx->cam_wlo.ensure((size_t)(ctx->obj.n_nodes + 1) * 12), "dp_pose_mesh: boxes");
    CK(ctx->cam_whi.ensure((size_t)(ctx->obj.n_nodes + 1) * 12), "dp_pose_mesh: boxes");
    (void)cap_nodes;
    ctx->cam.fat = nullptr;
    if (ctx->obj.fat) {
        CK(ctx->cam_fat.ensure((size_t)(ctx->obj.n_nodes + 1) * FAT_NODE_BYTES), "dp_pose_mesh: uncompressed nodes");
        ctx->cam.fat = ctx->cam_fat.as<uint4>();
    }
    ctx->cam.nodes = ctx->cam_nodes.as<WideNode>();
    ctx->cam.tris = ctx->cam_tris.as<TriRec>();
    ctx->cam.wlo = ctx->cam_wlo.as<float>();
    ctx->cam.whi = ctx->cam_whi.as<float>();
    ctx->cam.cap_nodes = ctx->obj.n_nodes + 1;
    ctx->cam.d_scale = ctx->scales.as<float>() + 1;
    CK(cudaEventRecord(ctx->ev[5], s), "dp_pose_mesh");
    CK(pose_and_refit(ctx->vdtype == DP_F64 ? ctx->V64.p : ctx->V.p, ctx->vdtype, ctx->nV, ctx->F.as<int32_t>(), ctx->nF, T,
                      ctx->Vposed.as<float>(), ctx->vdtype == DP_F64 ? ctx->Vposed64.as<double>() : nullptr, ctx->obj,
                      ctx->cam, ctx->topo, s),
       "dp_pose_mesh");
    CK(cudaEventRecord(ctx->ev[6], s), "dp_pose_mesh");
    ctx->has_cam = true;
    ctx->refit_ms = -1.0f;   // resolved lazily by dp_get_stats
    return DP_OK;
}

int dp_get_posed_vertices(dp_ctx *ctx, void *V, int vdtype, int mem, void *stream)
{
    if (!ctx || !V || (vdtype != DP_F32 && vdtype != DP_F64)) return fail(ctx, DP_E_ARG, "dp_get_posed_vertices: bad arguments");
    if (!ctx->has_cam) return fail(ctx, DP_E_STATE, "dp_get_posed_vertices: dp_pose_mesh has not been called");
    if (vdtype == DP_F64 && ctx->vdtype != DP_F64)
        return fail(ctx, DP_E_STATE, "dp_get_posed_vertices: float64 output needs a mesh set with float64 vertices");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (ctx->nV)
        CK(cudaMemcpyAsync(V, vdtype == DP_F64 ? ctx->Vposed64.p : ctx->Vposed.p, (size_t)ctx->nV * (vdtype == DP_F64 ? 24 : 12),
                           mem == DP_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s),
           "dp_get_posed_vertices");
    if (mem == DP_HOST) CK(cudaStreamSynchronize(s), "dp_get_posed_vertices");
    return DP_OK;
}

int dp_compute_rays(dp_ctx *ctx, const int32_t *xs, const int32_t *ys, int64_t n, const double *K, double *rays3, int mem,
                    void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || !K || (n > 0 && (!xs || !ys || !rays3))) return fail(ctx, DP_E_ARG, "dp_compute_rays: bad arguments");
    if (n == 0) return DP_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    FrameXf xf;
    dp_frame_xform(K, nullptr, xf.v);
    const int32_t *dx = xs, *dy = ys;
    double *dr = rays3;
    if (mem == DP_HOST) {
        CK(ctx->pixel.ensure((size_t)n * 8 + 16), "dp_compute_rays: xy");
        CK(ctx->point.ensure((size_t)n * 24 + 16), "dp_compute_rays: rays");
        CK(cudaMemcpyAsync(ctx->pixel.p, xs, (size_t)n * 4, cudaMemcpyHostToDevice, s), "dp_compute_rays: H2D");
        CK(cudaMemcpyAsync(ctx->pixel.as<int32_t>() + n, ys, (size_t)n * 4, cudaMemcpyHostToDevice, s), "dp_compute_rays: H2D");
        dx = ctx->pixel.as<int32_t>();
        dy = dx + n;
        dr = ctx->point.as<double>();
    }
    CK(launch_compute_rays(dx, dy, n, xf, dr, s), "dp_compute_rays: launch");
    if (mem == DP_HOST) {
        CK(cudaMemcpyAsync(rays3, dr, (size_t)n * 24, cudaMemcpyDeviceToHost, s), "dp_compute_rays: D2H");
        CK(cudaStreamSynchronize(s), "dp_compute_rays: kernel");
    }
    return DP_OK;
}

int dp_compact(dp_ctx *ctx, const void *heat, int dtype, int64_t nframes, int H, int W, double thr, uint32_t *pixel,
               float *intensity, int64_t cap, int64_t *n, int64_t *frame_count, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (nframes < 0 || H < 0 || W < 0 || cap < 0 || !n || (dtype != DP_F32 && dtype != DP_F64))
        return fail(ctx, DP_E_ARG, "dp_compact: bad arguments");
    const int64_t frame_elems = (int64_t)H * W, n_elems = nframes * frame_elems;
    if (n_elems > 0 && !heat) return fail(ctx, DP_E_ARG, "dp_compact: heat is NULL");
    if (n_elems > 0xffffffffLL) return fail(ctx, DP_E_ARG, "dp_compact: more than 2^32 pixels in one batch");
    if (cap > 0 && !pixel) return fail(ctx, DP_E_ARG, "dp_compact: pixel is NULL");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t esz = dtype == DP_F64 ? 8 : 4;
    const void *d_heat = heat;
    uint32_t *d_pixel = pixel;
    float *d_int = intensity;
    if (mem == DP_HOST) {
        CK(ctx->heat.ensure((size_t)n_elems * esz + 16), "dp_compact: heat");
        CK(ctx->pixel.ensure((size_t)cap * 4 + 16), "dp_compact: pixel");
        if (intensity) CK(ctx->inten.ensure((size_t)cap * 4 + 16), "dp_compact: intensity");
        if (n_elems) CK(cudaMemcpyAsync(ctx->heat.p, heat, (size_t)n_elems * esz, cudaMemcpyHostToDevice, s), "dp_compact: H2D");
        d_heat = ctx->heat.p;
        d_pixel = ctx->pixel.as<uint32_t>();
        d_int = intensity ? ctx->inten.as<float>() : nullptr;
    }
    CK(ctx->cscratch.ensure(compact_scratch_bytes(n_elems) + 64), "dp_compact: scratch");
    long long *d_fc = nullptr;
    if (frame_count && nframes > 0) {
        CK(ctx->fcounts.ensure((size_t)nframes * 8), "dp_compact: frame counts");
        d_fc = ctx->fcounts.as<long long>();
    }
    CK(launch_compact(d_heat, dtype, n_elems, frame_elems, thr, d_pixel, d_int, cap, ctx->cscratch.as<unsigned long long>(),
                      ctx->counts.as<long long>(), d_fc, nframes, s),
       "dp_compact: launch");
    CK(cudaMemcpyAsync(ctx->h_counts, ctx->counts.p, 8, cudaMemcpyDeviceToHost, s), "dp_compact: count");
    CK(cudaStreamSynchronize(s), "dp_compact: kernel");
    const int64_t total = ctx->h_counts[0];
    *n = total;
    const int64_t ncopy = total < cap ? total : cap;
    if (mem == DP_HOST && ncopy > 0) {
        CK(cudaMemcpyAsync(pixel, d_pixel, (size_t)ncopy * 4, cudaMemcpyDeviceToHost, s), "dp_compact: D2H");
        if (intensity) CK(cudaMemcpyAsync(intensity, d_int, (size_t)ncopy * 4, cudaMemcpyDeviceToHost, s), "dp_compact: D2H");
    }
    if (d_fc) CK(cudaMemcpyAsync(frame_count, d_fc, (size_t)nframes * 8, cudaMemcpyDeviceToHost, s), "dp_compact: D2H");
    CK(cudaStreamSynchronize(s), "dp_compact: D2H");
    if (total > cap) return fail(ctx, DP_E_NOMEM, "dp_compact: capacity too small for the selected pixels");
    return DP_OK;
}

int dp_cast_rays(dp_ctx *ctx, int frame, const float *rays6, int64_t n, float *t_hit, int32_t *face, int mem,
                 void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || (n > 0 && (!rays6 || !t_hit))) return fail(ctx, DP_E_ARG, "dp_cast_rays: bad arguments");
    if (frame == DP_FRAME_OBJECT ? !ctx->has_bvh : !ctx->has_cam)
        return fail(ctx, DP_E_STATE, "dp_cast_rays: no BVH for the requested frame");
    if (frame != DP_FRAME_OBJECT && frame != DP_FRAME_CAMERA) return fail(ctx, DP_E_ARG, "dp_cast_rays: bad frame");
    if (n == 0) return DP_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float *d_rays = rays6;
    float *d_t = t_hit;
    int32_t *d_f = face;
    if (mem == DP_HOST) {
        CK(ctx->rays6.ensure((size_t)n * 24), "dp_cast_rays: rays");
        CK(ctx->t_hit.ensure((size_t)n * 4), "dp_cast_rays: t");
        CK(ctx->face.ensure((size_t)n * 4), "dp_cast_rays: face");
        CK(cudaMemcpyAsync(ctx->rays6.p, rays6, (size_t)n * 24, cudaMemcpyHostToDevice, s), "dp_cast_rays: H2D");
        d_rays = ctx->rays6.as<float>();
        d_t = ctx->t_hit.as<float>();
        d_f = ctx->face.as<int32_t>();
    }
    TraceStats *st = nullptr;
    if (ctx->stats_on) {
        CK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(TraceStats), s), "dp_cast_rays: stats");
        st = ctx->stats.as<TraceStats>();
    }
    const BvhStorage &b = frame == DP_FRAME_OBJECT ? ctx->obj : ctx->cam;
    CK(launch_trace_rays6(view_of(b), d_rays, n, d_t, d_f, ctx->counts.as<unsigned long long>() + 2, st, s),
       "dp_cast_rays: launch");
    if (mem == DP_HOST) {
        CK(cudaMemcpyAsync(t_hit, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "dp_cast_rays: D2H");
        if (face) CK(cudaMemcpyAsync(face, d_f, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "dp_cast_rays: D2H");
        CK(cudaStreamSynchronize(s), "dp_cast_rays: kernel");
    }
    return DP_OK;
}

int dp_project(dp_ctx *ctx, int frame, const void *heat, int dtype, int64_t nframes, int H, int W, double thr,
               const double *K, int64_t nK, const double *pose, int accumulate, dp_rays_out *out, int64_t *n_rays,
               int64_t *n_hits, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (nframes < 0 || H < 0 || W < 0 || !K || (nK != 1 && nK != nframes) || (dtype != DP_F32 && dtype != DP_F64))
        return fail(ctx, DP_E_ARG, "dp_project: bad arguments");
    if (frame != DP_FRAME_OBJECT && frame != DP_FRAME_CAMERA) return fail(ctx, DP_E_ARG, "dp_project: bad frame");
    if (frame == DP_FRAME_OBJECT && !pose && nframes > 0) return fail(ctx, DP_E_ARG, "dp_project: pose is NULL");
    if (frame == DP_FRAME_OBJECT ? !ctx->has_bvh : !ctx->has_cam)
        return fail(ctx, DP_E_STATE, "dp_project: no BVH for the requested frame (dp_build_bvh / dp_pose_mesh)");
    const int64_t frame_elems = (int64_t)H * W, n_elems = nframes * frame_elems;
    if (n_elems > 0 && !heat) return fail(ctx, DP_E_ARG, "dp_project: heat is NULL");
    if (n_elems > 0xffffffffLL) return fail(ctx, DP_E_ARG, "dp_project: more than 2^32 pixels in one batch");
    if (mem == DP_HOST && !n_rays && out && (out->pixel || out->intensity || out->t_hit || out->face || out->point || out->point64))
        return fail(ctx, DP_E_ARG, "dp_project: host outputs need n_rays (the call must synchronise)");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);

    const bool want_pix = out && out->pixel, want_int = out && out->intensity, want_t = out && out->t_hit,
               want_face = out && out->face, want_pt = out && out->point, want_p64 = out && out->point64;
    int64_t cap = n_elems;
    if (out && (want_pix || want_int || want_t || want_face || want_pt || want_p64)) {
        if (out->cap < 0) return fail(ctx, DP_E_ARG, "dp_project: negative capacity");
        if (out->cap < cap) cap = out->cap;
    }
    const bool peer_mode = ctx->peer_res_slot >= 0 && ctx->peer_opened && ctx->shard.world > 1;
    if (peer_mode) {
        if (mem != DP_DEVICE || want_t || want_face || want_pt || want_p64)
            return fail(ctx, DP_E_ARG, "dp_project: with dp_peer_results the per-ray results live in the exchange window "
                                       "(device call, no t_hit / face / point / point64 in dp_rays_out)");
        if (ctx->shard.world != ctx->peer.world || ctx->shard.rank != ctx->peer.rank)
            return fail(ctx, DP_E_STATE, "dp_project: dp_set_ray_shard and dp_peer_open disagree on (rank, world)");
        if ((int64_t)ctx->peer.res_cap < cap) cap = (int64_t)ctx->peer.res_cap;
    }

    // per-frame constants (uploaded by the prologue kernel below)
    std::vector<FrameXf> hxf((size_t)(nframes > 0 ? nframes : 1));
    for (int64_t f = 0; f < nframes; ++f)
        dp_frame_xform(K + 9 * (nK == 1 ? 0 : f), frame == DP_FRAME_OBJECT ? pose + 16 * f : nullptr, hxf[f].v);
    CK(ctx->xf.ensure(hxf.size() * sizeof(FrameXf)), "dp_project: xf");

    const size_t esz = dtype == DP_F64 ? 8 : 4;
    const void *d_heat = heat;
    if (ctx->timing_on) CK(cudaEventRecord(ctx->ev[0], s), "dp_project");
    if (mem == DP_HOST) {
        CK(ctx->heat.ensure((size_t)n_elems * esz + 16), "dp_project: heat");
        if (n_elems) CK(cudaMemcpyAsync(ctx->heat.p, heat, (size_t)n_elems * esz, cudaMemcpyHostToDevice, s), "dp_project: H2D");
        d_heat = ctx->heat.p;
    }
    // ray buffers: the caller's (device) or ours
    uint32_t *d_pixel;
    float *d_int, *d_t = nullptr, *d_pt = nullptr;
    double *d_p64 = nullptr;
    int32_t *d_face = nullptr;
    const bool own = (mem == DP_HOST);
    if (own || !want_pix) { CK(ctx->pixel.ensure((size_t)cap * 4 + 16), "dp_project: pixel"); d_pixel = ctx->pixel.as<uint32_t>(); }
    else d_pixel = out->pixel;
    if (own || !want_int) { CK(ctx->inten.ensure((size_t)cap * 4 + 16), "dp_project: intensity"); d_int = ctx->inten.as<float>(); }
    else d_int = out->intensity;
    if (want_t) { if (own) { CK(ctx->t_hit.ensure((size_t)cap * 4 + 16), "dp_project: t"); d_t = ctx->t_hit.as<float>(); } else d_t = out->t_hit; }
    if (want_face) { if (own) { CK(ctx->face.ensure((size_t)cap * 4 + 16), "dp_project: face"); d_face = ctx->face.as<int32_t>(); } else d_face = out->face; }
    if (want_p64) { if (own) { CK(ctx->point64.ensure((size_t)cap * 24 + 16), "dp_project: point64"); d_p64 = ctx->point64.as<double>(); } else d_p64 = out->point64; }
    if (want_pt) { if (own) { CK(ctx->point.ensure((size_t)cap * 12 + 16), "dp_project: point"); d_pt = ctx->point.as<float>(); } else d_pt = out->point; }
    // ray-sharded frame over peer memory: t_hit / face / point live in the exchange windows; this rank's traversal stores
    // its slice into every rank's arrays (its own window is the local destination) and the call ends with the frame barrier
    const PeerOut *peer_out = nullptr;
    if (peer_mode) {
        char *r = ctx->peer_win.as<char>() + ctx->peer.res_off[ctx->peer_res_slot];
        d_t = reinterpret_cast<float *>(r);
        d_face = reinterpret_cast<int32_t *>(r + ctx->peer.res_cap * 4);
        d_pt = ctx->peer_res_points ? reinterpret_cast<float *>(r + ctx->peer.res_cap * 8) : nullptr;
        peer_out = ctx->peer_out.as<PeerOut>() + ctx->peer_res_slot;
    }

    // <= 8 frames: the compaction launch does the call's resets and uploads itself and leaves its scratch clean (a buffer
    // of its own, zeroed when it is allocated); larger batches go through the prologue launch
    static const bool fused_knob = [] { const char *e = getenv("DP_FUSED_PROLOGUE"); return e ? atoi(e) != 0 : true; }();
    const bool fused_prologue = fused_knob && n_elems > 0 && nframes <= 8;
    if (fused_prologue) {
        const size_t before = ctx->pscratch.cap;            // (a re-allocation may return the same address)
        CK(ctx->pscratch.ensure(compact_scratch_bytes(n_elems) + 64), "dp_project: scratch");
        if (ctx->pscratch.cap != before) {
            CK(cudaMemsetAsync(ctx->pscratch.p, 0, ctx->pscratch.cap, s), "dp_project: scratch");
            CK(cudaMemsetAsync(ctx->counts.p, 0, 32, s), "dp_project: counts");
        }
    } else {
        CK(ctx->cscratch.ensure(compact_scratch_bytes(n_elems) + 64), "dp_project: scratch");
    }
    long long *d_counts = ctx->counts.as<long long>();
    // packet schedule: read the state the previous call wrote, write the other one (double buffer)
    const OrderState *ord_prev = nullptr;
    OrderState *ord_next = nullptr;
    {
        const size_t np = (size_t)(cap / 32 + 2);
        const size_t per = np * 4 * 3 + ((np + 15) & ~size_t(15));
        const void *before = ctx->order.p;
        CK(ctx->order.ensure(2 * per + 256), "dp_project: schedule");
        char *basep = ctx->order.as<char>();
        OrderState *dstate = reinterpret_cast<OrderState *>(basep);
        if (ctx->order.p != before || np != ctx->order_np) {
            OrderState h[2];
            for (int k = 0; k < 2; ++k) {
                char *p = basep + 256 + k * per;
                h[k].n_valid = -1; h[k].cost_sum = 0; h[k].cnt[0] = h[k].cnt[1] = h[k].cnt[2] = 0; h[k].pad_ = 0;
                h[k].list0 = reinterpret_cast<uint32_t *>(p);
                h[k].list1 = reinterpret_cast<uint32_t *>(p + np * 4);
                h[k].list2 = reinterpret_cast<uint32_t *>(p + np * 8);
                h[k].flags = reinterpret_cast<unsigned char *>(p + np * 12);
            }
            CK(cudaMemcpyAsync(dstate, h, sizeof(h), cudaMemcpyHostToDevice, s), "dp_project: schedule");   // (re)allocation only
            ctx->order_parity = 0;
            ctx->order_np = np;
        }
        const int cur = ctx->order_parity, nxt = cur ^ 1;
        ord_prev = dstate + cur;
        ord_next = dstate + nxt;
        ctx->order_parity = nxt;
    }
    // one launch resets the counts, the traversal work counter, the compaction scratch and the schedule state this
    // call will write, and uploads the per-frame constants: no copy-engine work in the kernel stream
    if (!fused_prologue)
        CK(launch_project_prologue(ctx->cscratch.as<unsigned long long>(), n_elems, d_counts, ord_next, ctx->xf.as<FrameXf>(),
                                   hxf.data(), nframes, s),
           "dp_project: prologue");
    // out->counts in pinned host memory: the ray count is stored there by the compaction itself (early), both counts at
    // the end of the call
    long long *early_n = nullptr;
    if (out && out->counts) {
        if (out->counts != ctx->counts_alias_src) {
            cudaPointerAttributes at{};
            ctx->counts_alias = nullptr;
            if (cudaPointerGetAttributes(&at, out->counts) == cudaSuccess && at.devicePointer)
                ctx->counts_alias = static_cast<long long *>(at.devicePointer);
            else
                cudaGetLastError();
            ctx->counts_alias_src = out->counts;
        }
        early_n = ctx->counts_alias;
    }
    // one K for every frame: the compaction launch also tabulates the two per-pixel quotients of the ray generation
    static const bool raytab_knob = [] { const char *e = getenv("DP_RAYTAB"); return e ? atoi(e) != 0 : true; }();
    double *d_raytab = nullptr;
    if (fused_prologue && nK == 1 && raytab_knob) {
        CK(ctx->raytab.ensure((size_t)(W + H) * 8 + 16), "dp_project: ray table");
        d_raytab = ctx->raytab.as<double>();
    }
    if (fused_prologue)
        CK(launch_compact_fused(d_heat, dtype, n_elems, thr, d_pixel, d_int, cap, ctx->pscratch.as<unsigned long long>(), d_counts,
                                ord_next, ctx->xf.as<FrameXf>(), hxf.data(), (int)nframes, early_n, s, d_raytab, H, W),
           "dp_project: compaction");
    else
        CK(launch_compact(d_heat, dtype, n_elems, frame_elems, thr, d_pixel, d_int, cap,
                          ctx->cscratch.as<unsigned long long>(), d_counts, nullptr, nframes, s, true, early_n),
           "dp_project: compaction");
    if (ctx->timing_on) CK(cudaEventRecord(ctx->ev[1], s), "dp_project");

    TraceStats *st = nullptr;
    if (ctx->stats_on) {
        CK(ctx->ray_nodes.ensure((size_t)cap * 4 + 16), "dp_project: stats");
        TraceStats init{};
        init.ray_nodes = ctx->ray_nodes.as<unsigned>();
        CK(cudaMemsetAsync(ctx->ray_nodes.p, 0, (size_t)cap * 4, s), "dp_project: stats");
        CK(cudaMemcpyAsync(ctx->stats.p, &init, sizeof(TraceStats), cudaMemcpyHostToDevice, s), "dp_project: stats");
        st = ctx->stats.as<TraceStats>();
        ctx->ray_nodes_n = cap;
    }
    Accum acc{ctx->hist.as<int32_t>(), ctx->fmax.as<uint32_t>(), ctx->vmax.as<uint32_t>(), ctx->F.as<int32_t>()};
    const BvhStorage &b = frame == DP_FRAME_OBJECT ? ctx->obj : ctx->cam;
    // t_hit is needed by the hit-point kernel even when the caller does not want it
    if ((want_pt || want_p64) && !d_t && !peer_mode) { CK(ctx->t_hit.ensure((size_t)cap * 4 + 16), "dp_project: t"); d_t = ctx->t_hit.as<float>(); }
    // DP_FUSE_RAYS=0: separate ray-generation and hit-point kernels around the traversal (the round-1 sequence, kept for
    // A/B); default: the traversal generates its rays from the compacted pixels and writes the hit points itself --
    // two launches and the 32-byte ray records (write + read) less per frame
    static const int fuse_default = 1;
    const char *fuse_env = getenv("DP_FUSE_RAYS");
    const bool fuse = fuse_env ? atoi(fuse_env) != 0 : fuse_default != 0;
    float4 *d_dir4 = nullptr;
    if (peer_mode && !fuse) return fail(ctx, DP_E_STATE, "dp_project: dp_peer_results needs the fused traversal (DP_FUSE_RAYS=1)");
    if (!fuse) {
        CK(ctx->dir4.ensure((size_t)cap * 32 + 32), "dp_project: rays");
        d_dir4 = ctx->dir4.as<float4>();
        CK(launch_raygen(d_pixel, d_counts, cap, H, W, ctx->xf.as<FrameXf>(), nframes, d_dir4, s, n_elems, ctx->shard),
           "dp_project: ray generation");
    }
    if (ctx->timing_on) CK(cudaEventRecord(ctx->ev[7], s), "dp_project");
    CK(launch_trace_pixels(view_of(b), d_dir4, d_int, d_counts, cap, n_elems, H, W, ctx->xf.as<FrameXf>(),
                           d_t, d_face, accumulate ? &acc : nullptr, reinterpret_cast<unsigned long long *>(d_counts + 2),
                           d_counts + 1, st, ord_prev, ord_next, s, true, ctx->shard, d_pixel, nframes, fuse ? d_pt : nullptr,
                           fuse ? d_p64 : nullptr, peer_out, fuse ? d_raytab : nullptr),
       "dp_project: traversal");
    if (peer_mode) CK(launch_peer_frame_done(ctx->peer, ++ctx->peer_epoch[1], s), "dp_project: frame barrier");
    if (ctx->timing_on) CK(cudaEventRecord(ctx->ev[3], s), "dp_project");
    if (!fuse)
        CK(launch_points(d_pixel, d_t, d_counts, cap, H, W, ctx->xf.as<FrameXf>(), nframes, d_pt, d_p64, s, n_elems, ctx->shard),
           "dp_project: hit points");
    if (out && out->counts) {
        // device memory: plain store; pinned host memory: zero-copy store through its device alias
        if (ctx->counts_alias) CK(launch_publish_counts(d_counts, ctx->counts_alias, s), "dp_project: counts");
        else CK(cudaMemcpyAsync(out->counts, d_counts, 16, cudaMemcpyDefault, s), "dp_project: counts");
    }
    if (ctx->timing_on) CK(cudaEventRecord(ctx->ev[2], s), "dp_project");
    ctx->timings_valid = ctx->timing_on;

    if (n_rays || n_hits) {
        CK(cudaMemcpyAsync(ctx->h_counts, d_counts, 16, cudaMemcpyDeviceToHost, s), "dp_project: counts");
        CK(cudaStreamSynchronize(s), "dp_project: kernels");
        const int64_t total = ctx->h_counts[0];
        if (n_rays) *n_rays = total;
        if (n_hits) *n_hits = ctx->h_counts[1];
        const int64_t nc = total < cap ? total : cap;
        if (own && nc > 0) {
            if (want_pix) CK(cudaMemcpyAsync(out->pixel, d_pixel, (size_t)nc * 4, cudaMemcpyDeviceToHost, s), "dp_project: D2H");
            if (want_int) CK(cudaMemcpyAsync(out->intensity, d_int, (size_t)nc * 4, cudaMemcpyDeviceToHost, s), "dp_project: D2H");
            if (want_t) CK(cudaMemcpyAsync(out->t_hit, d_t, (size_t)nc * 4, cudaMemcpyDeviceToHost, s), "dp_project: D2H");
            if (want_face) CK(cudaMemcpyAsync(out->face, d_face, (size_t)nc * 4, cudaMemcpyDeviceToHost, s), "dp_project: D2H");
            if (want_pt) CK(cudaMemcpyAsync(out->point, d_pt, (size_t)nc * 12, cudaMemcpyDeviceToHost, s), "dp_project: D2H");
            if (want_p64) CK(cudaMemcpyAsync(out->point64, d_p64, (size_t)nc * 24, cudaMemcpyDeviceToHost, s), "dp_project: D2H");
            CK(cudaStreamSynchronize(s), "dp_project: D2H");
        }
        if (total > cap) return fail(ctx, DP_E_NOMEM, "dp_project: output capacity too small for the selected pixels");
    }
    return DP_OK;
}

int dp_depth_backproject(dp_ctx *ctx, const void *heat, int dtype, int H, int W, const uint16_t *depth, int Hd, int Wd,
                         const double *K, double thr, double *points4, int64_t cap, int64_t *n, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (H < 0 || W < 0 || Hd < 0 || Wd < 0 || cap < 0 || !K || !n || (dtype != DP_F32 && dtype != DP_F64))
        return fail(ctx, DP_E_ARG, "dp_depth_backproject: bad arguments");
    const int64_t ne = (int64_t)H * W, nd = (int64_t)Hd * Wd;
    if ((ne > 0 && !heat) || (nd > 0 && !depth) || (cap > 0 && !points4)) return fail(ctx, DP_E_ARG, "dp_depth_backproject: NULL buffer");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t esz = dtype == DP_F64 ? 8 : 4;
    const void *d_heat = heat;
    const uint16_t *d_depth = depth;
    double *d_out = points4;
    if (mem == DP_HOST) {
        CK(ctx->heat.ensure((size_t)ne * esz + 16), "dp_depth_backproject: heat");
        CK(ctx->tmp[0].ensure((size_t)nd * 2 + 16), "dp_depth_backproject: depth");
        CK(ctx->tmp[1].ensure((size_t)cap * 32 + 16), "dp_depth_backproject: out");
        if (ne) CK(cudaMemcpyAsync(ctx->heat.p, heat, (size_t)ne * esz, cudaMemcpyHostToDevice, s), "dp_depth_backproject: H2D");
        if (nd) CK(cudaMemcpyAsync(ctx->tmp[0].p, depth, (size_t)nd * 2, cudaMemcpyHostToDevice, s), "dp_depth_backproject: H2D");
        d_heat = ctx->heat.p;
        d_depth = ctx->tmp[0].as<uint16_t>();
        d_out = ctx->tmp[1].as<double>();
    }
    CK(ctx->cscratch.ensure(depth_select_scratch_bytes(ne) + 64), "dp_depth_backproject: scratch");
    CK(ctx->tmp[2].ensure(64), "dp_depth_backproject: scalars");
    long long *d_counts = ctx->counts.as<long long>();
    CK(launch_depth_select(d_heat, dtype, H, W, d_depth, Hd, Wd, thr, K, d_out, cap, ctx->cscratch.as<unsigned long long>(),
                           d_counts, ctx->tmp[2].as<double>(), reinterpret_cast<int *>(ctx->tmp[2].as<char>() + 16), s),
       "dp_depth_backproject: launch");
    CK(cudaMemcpyAsync(ctx->h_counts, d_counts, 8, cudaMemcpyDeviceToHost, s), "dp_depth_backproject: count");
    CK(cudaStreamSynchronize(s), "dp_depth_backproject: kernels");
    const int64_t total = ctx->h_counts[0];
    *n = total;
    const int64_t nc = total < cap ? total : cap;
    if (mem == DP_HOST && nc > 0) {
        CK(cudaMemcpyAsync(points4, d_out, (size_t)nc * 32, cudaMemcpyDeviceToHost, s), "dp_depth_backproject: D2H");
        CK(cudaStreamSynchronize(s), "dp_depth_backproject: D2H");
    }
    if (total > cap) return fail(ctx, DP_E_NOMEM, "dp_depth_backproject: capacity too small for the selected pixels");
    return DP_OK;
}

int dp_calc_coordinates(dp_ctx *ctx, const int32_t *xs, const int32_t *ys, int64_t n, const uint16_t *depth, int Hd, int Wd,
                        const double *K, double *out3, unsigned char *valid, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || Hd < 0 || Wd < 0 || !K || (n > 0 && (!xs || !ys || !out3 || !valid)) || ((int64_t)Hd * Wd > 0 && !depth))
        return fail(ctx, DP_E_ARG, "dp_calc_coordinates: bad arguments");
    if (n == 0) return DP_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t nd = (int64_t)Hd * Wd;
    const int32_t *dx = xs, *dy = ys;
    const uint16_t *dd = depth;
    double *dout = out3;
    unsigned char *dv = valid;
    if (mem == DP_HOST) {
        CK(ctx->tmp[0].ensure((size_t)nd * 2 + 16), "dp_calc_coordinates: depth");
        CK(ctx->tmp[1].ensure((size_t)n * 24 + 16), "dp_calc_coordinates: out");
        CK(ctx->tmp[3].ensure((size_t)n * 8 + 16), "dp_calc_coordinates: xy");
        CK(ctx->tmp[4].ensure((size_t)n + 16), "dp_calc_coordinates: valid");
        if (nd) CK(cudaMemcpyAsync(ctx->tmp[0].p, depth, (size_t)nd * 2, cudaMemcpyHostToDevice, s), "dp_calc_coordinates: H2D");
        CK(cudaMemcpyAsync(ctx->tmp[3].p, xs, (size_t)n * 4, cudaMemcpyHostToDevice, s), "dp_calc_coordinates: H2D");
        CK(cudaMemcpyAsync(ctx->tmp[3].as<int32_t>() + n, ys, (size_t)n * 4, cudaMemcpyHostToDevice, s), "dp_calc_coordinates: H2D");
        dd = ctx->tmp[0].as<uint16_t>();
        dx = ctx->tmp[3].as<int32_t>();
        dy = dx + n;
        dout = ctx->tmp[1].as<double>();
        dv = ctx->tmp[4].as<unsigned char>();
    }
    CK(launch_calc_coordinates(dx, dy, n, dd, Hd, Wd, K, dout, dv, s), "dp_calc_coordinates: launch");
    if (mem == DP_HOST) {
        CK(cudaMemcpyAsync(out3, dout, (size_t)n * 24, cudaMemcpyDeviceToHost, s), "dp_calc_coordinates: D2H");
        CK(cudaMemcpyAsync(valid, dv, (size_t)n, cudaMemcpyDeviceToHost, s), "dp_calc_coordinates: D2H");
        CK(cudaStreamSynchronize(s), "dp_calc_coordinates: kernels");
    }
    return DP_OK;
}

int dp_align_to_surface(dp_ctx *ctx, const double *query, int stride, int64_t n, const double *target, const double *normals,
                        int64_t m, double offset, double *offset_points, double *aligned_points, int32_t *idx, int mem,
                        void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || m < 0 || stride < 3 || (n > 0 && !query) || (m > 0 && !target) || (offset_points && !normals))
        return fail(ctx, DP_E_ARG, "dp_align_to_surface: bad arguments");
    if (n > 0 && m == 0) return fail(ctx, DP_E_ARG, "dp_align_to_surface: empty target cloud");
    if (n == 0) return DP_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const double *dq = query, *dt = target, *dn = normals;
    double *doff = offset_points, *dal = aligned_points;
    int32_t *di = idx;
    if (mem == DP_HOST) {
        CK(ctx->tmp[0].ensure((size_t)n * stride * 8 + 16), "dp_align_to_surface: query");
        CK(ctx->tmp[1].ensure((size_t)m * 24 + 16), "dp_align_to_surface: target");
        CK(ctx->tmp[2].ensure((size_t)m * 24 + 64), "dp_align_to_surface: normals");
        CK(ctx->tmp[3].ensure((size_t)n * 24 + 16), "dp_align_to_surface: out");
        CK(ctx->tmp[4].ensure((size_t)n * 24 + 16), "dp_align_to_surface: out");
        CK(ctx->tmp[5].ensure((size_t)n * 4 + 16), "dp_align_to_surface: idx");
        CK(cudaMemcpyAsync(ctx->tmp[0].p, query, (size_t)n * stride * 8, cudaMemcpyHostToDevice, s), "dp_align_to_surface: H2D");
        CK(cudaMemcpyAsync(ctx->tmp[1].p, target, (size_t)m * 24, cudaMemcpyHostToDevice, s), "dp_align_to_surface: H2D");
        if (normals) CK(cudaMemcpyAsync(ctx->tmp[2].p, normals, (size_t)m * 24, cudaMemcpyHostToDevice, s), "dp_align_to_surface: H2D");
        dq = ctx->tmp[0].as<double>();
        dt = ctx->tmp[1].as<double>();
        dn = normals ? ctx->tmp[2].as<double>() : nullptr;
        doff = offset_points ? ctx->tmp[3].as<double>() : nullptr;
        dal = aligned_points ? ctx->tmp[4].as<double>() : nullptr;
        di = idx ? ctx->tmp[5].as<int32_t>() : nullptr;
    }
    // Large searches go through the uniform grid over the target (built per call; ring search, same result bit for
    // bit); DP_NN_GRID=0 keeps the tiled scan of the whole target for A/B measurements.
    const char *knob = getenv("DP_NN_GRID");
    IcpGridView gv;
    gv.tps = nullptr;
    if ((knob ? atoi(knob) : 1) && m >= 4096 && m < (int64_t)1 << 31 && (double)n * (double)m >= 134217728.0) {
        CK(ctx->tmp[6].ensure(icp_grid_bytes(m)), "dp_align_to_surface: grid");
        CK(icp_grid_build(dt, dn, m, 0.0, 1, ctx->tmp[6].p, &gv, s), "dp_align_to_surface: grid build");
    }
    if (gv.tps) CK(launch_nearest_grid(dq, stride, n, gv, dn != nullptr, offset, di, dal, doff, s), "dp_align_to_surface: launch");
    else CK(launch_nearest(dq, stride, n, dt, m, dn, offset, di, dal, doff, s), "dp_align_to_surface: launch");
    if (mem == DP_HOST) {
        if (offset_points) CK(cudaMemcpyAsync(offset_points, doff, (size_t)n * 24, cudaMemcpyDeviceToHost, s), "dp_align_to_surface: D2H");
        if (aligned_points) CK(cudaMemcpyAsync(aligned_points, dal, (size_t)n * 24, cudaMemcpyDeviceToHost, s), "dp_align_to_surface: D2H");
        if (idx) CK(cudaMemcpyAsync(idx, di, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "dp_align_to_surface: D2H");
        CK(cudaStreamSynchronize(s), "dp_align_to_surface: kernels");
    }
    return DP_OK;
}

int dp_estimate_normals(dp_ctx *ctx, const double *points, int64_t n, double radius, int max_nn, double *normals,
                        int has_normals, int32_t *neighbours, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || n >= (int64_t)1 << 31 || !(radius > 0.0) || max_nn < 1 || max_nn > NORMALS_MAX_NN ||
        (n > 0 && (!points || !normals)))
        return fail(ctx, DP_E_ARG, "dp_estimate_normals: bad arguments (max_nn is limited to 64)");
    if (n == 0) return DP_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const double *dp_ = points;
    double *dn = normals;
    int32_t *dc = neighbours;
    if (mem == DP_HOST) {
        CK(ctx->tmp[0].ensure((size_t)n * 24 + 16), "dp_estimate_normals: points");
        CK(ctx->tmp[1].ensure((size_t)n * 24 + 16), "dp_estimate_normals: normals");
        CK(cudaMemcpyAsync(ctx->tmp[0].p, points, (size_t)n * 24, cudaMemcpyHostToDevice, s), "dp_estimate_normals: H2D");
        if (has_normals)
            CK(cudaMemcpyAsync(ctx->tmp[1].p, normals, (size_t)n * 24, cudaMemcpyHostToDevice, s), "dp_estimate_normals: H2D");
        dp_ = ctx->tmp[0].as<double>();
        dn = ctx->tmp[1].as<double>();
        if (neighbours) {
            CK(ctx->tmp[5].ensure((size_t)n * 4 + 16), "dp_estimate_normals: counts");
            dc = ctx->tmp[5].as<int32_t>();
        }
    }
    // the grid of the ICP row over the cloud itself: cells at least `radius` wide, so the neighbours are in 27 cells
    IcpGridView gv;
    gv.tps = nullptr;
    CK(ctx->tmp[6].ensure(icp_grid_bytes(n)), "dp_estimate_normals: grid");
    CK(icp_grid_build(dp_, nullptr, n, radius, 1, ctx->tmp[6].p, &gv, s), "dp_estimate_normals: grid build");
    if (!gv.tps) return fail(ctx, DP_E_ARG, "dp_estimate_normals: the cloud has non-finite coordinates");
    CK(launch_estimate_normals(gv, n, radius, max_nn, dn, has_normals, dc, s), "dp_estimate_normals: launch");
    if (mem == DP_HOST) {
        CK(cudaMemcpyAsync(normals, dn, (size_t)n * 24, cudaMemcpyDeviceToHost, s), "dp_estimate_normals: D2H");
        if (neighbours) CK(cudaMemcpyAsync(neighbours, dc, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "dp_estimate_normals: D2H");
        CK(cudaStreamSynchronize(s), "dp_estimate_normals: kernels");
    }
    return DP_OK;
}

int dp_prepare_heatmap(dp_ctx *ctx, const void *data, int dtype, int src_h, int src_w, int H, int W, void *out, int out_dtype,
                       int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (src_h <= 0 || src_w <= 0 || H < 0 || W < 0 || !data || (dtype != DP_F32 && dtype != DP_F64) ||
        (out_dtype != DP_F32 && out_dtype != DP_F64) || ((int64_t)H * W > 0 && !out))
        return fail(ctx, DP_E_ARG, "dp_prepare_heatmap: bad arguments");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t isz = dtype == DP_F64 ? 8 : 4, osz = out_dtype == DP_F64 ? 8 : 4;
    const int64_t ns = (int64_t)src_h * src_w, no = (int64_t)H * W;
    const void *d_in = data;
    void *d_out = out;
    if (mem == DP_HOST) {
        CK(ctx->tmp[0].ensure((size_t)ns * isz + 16), "dp_prepare_heatmap: in");
        CK(ctx->tmp[1].ensure((size_t)no * osz + 16), "dp_prepare_heatmap: out");
        CK(cudaMemcpyAsync(ctx->tmp[0].p, data, (size_t)ns * isz, cudaMemcpyHostToDevice, s), "dp_prepare_heatmap: H2D");
        d_in = ctx->tmp[0].p;
        d_out = ctx->tmp[1].p;
    }
    CK(ctx->tmp[2].ensure(64), "dp_prepare_heatmap: scalars");
    CK(launch_prepare_heatmap(d_in, dtype, src_h, src_w, H, W, d_out, out_dtype, ctx->tmp[2].as<long long>(), s),
       "dp_prepare_heatmap: launch");
    if (mem == DP_HOST) {
        if (no) CK(cudaMemcpyAsync(out, d_out, (size_t)no * osz, cudaMemcpyDeviceToHost, s), "dp_prepare_heatmap: D2H");
        CK(cudaStreamSynchronize(s), "dp_prepare_heatmap: kernels");
    }
    return DP_OK;
}

int dp_transform_points(dp_ctx *ctx, double *points3, int64_t n, const double *T, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || !T || (n > 0 && !points3)) return fail(ctx, DP_E_ARG, "dp_transform_points: bad arguments");
    if (n == 0) return DP_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double *d = points3;
    if (mem == DP_HOST) {
        CK(ctx->tmp[0].ensure((size_t)n * 24 + 16), "dp_transform_points: buffer");
        CK(cudaMemcpyAsync(ctx->tmp[0].p, points3, (size_t)n * 24, cudaMemcpyHostToDevice, s), "dp_transform_points: H2D");
        d = ctx->tmp[0].as<double>();
    }
    CK(launch_transform_points(d, n, T, s), "dp_transform_points: launch");
    if (mem == DP_HOST) {
        CK(cudaMemcpyAsync(points3, d, (size_t)n * 24, cudaMemcpyDeviceToHost, s), "dp_transform_points: D2H");
        CK(cudaStreamSynchronize(s), "dp_transform_points: kernels");
    }
    return DP_OK;
}

void dp_jet_lut(double *lut)
{
    if (lut) jet_lut_host(lut);
}

int dp_pack_hits(dp_ctx *ctx, const void *intensity, int dtype, const int32_t *face, const uint32_t *pixel, const double *point64,
                 int64_t n, const double *T, double *points, double *colors, int32_t *face_out, uint32_t *pixel_out,
                 double *intensity_out, int64_t cap, int64_t *m, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || cap < 0 || !m || (dtype != DP_F32 && dtype != DP_F64) || (n > 0 && !intensity) || (points && !point64))
        return fail(ctx, DP_E_ARG, "dp_pack_hits: bad arguments");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!ctx->jet.p) {
        double lut[768];
        jet_lut_host(lut);
        CK(ctx->jet.ensure(sizeof(lut)), "dp_pack_hits: colour table");
        CK(cudaMemcpy(ctx->jet.p, lut, sizeof(lut), cudaMemcpyHostToDevice), "dp_pack_hits: colour table");
    }
    const size_t isz = dtype == DP_F64 ? 8 : 4;
    const void *d_in = intensity;
    const int32_t *d_face = face;
    const uint32_t *d_pix = pixel;
    const double *d_p64 = point64;
    double *d_pts = points, *d_col = colors, *d_io = intensity_out;
    int32_t *d_fo = face_out;
    uint32_t *d_po = pixel_out;
    if (mem == DP_HOST) {
        // inputs: tmp[0] intensity | tmp[1] face + pixel | tmp[3] point64 ; outputs: tmp[4] points | tmp[5] colours |
        // tmp[6] face_out + pixel_out | tmp[7] intensity_out
        CK(ctx->tmp[0].ensure((size_t)n * isz + 16), "dp_pack_hits: in");
        CK(ctx->tmp[1].ensure((size_t)n * 8 + 16), "dp_pack_hits: in");
        if (n) CK(cudaMemcpyAsync(ctx->tmp[0].p, intensity, (size_t)n * isz, cudaMemcpyHostToDevice, s), "dp_pack_hits: H2D");
        d_in = ctx->tmp[0].p;
        if (face) {
            if (n) CK(cudaMemcpyAsync(ctx->tmp[1].p, face, (size_t)n * 4, cudaMemcpyHostToDevice, s), "dp_pack_hits: H2D");
            d_face = ctx->tmp[1].as<int32_t>();
        }
        if (pixel) {
            if (n) CK(cudaMemcpyAsync(ctx->tmp[1].as<int32_t>() + n, pixel, (size_t)n * 4, cudaMemcpyHostToDevice, s), "dp_pack_hits: H2D");
            d_pix = ctx->tmp[1].as<uint32_t>() + n;
        }
        if (point64) {
            CK(ctx->tmp[3].ensure((size_t)n * 24 + 16), "dp_pack_hits: in");
            if (n) CK(cudaMemcpyAsync(ctx->tmp[3].p, point64, (size_t)n * 24, cudaMemcpyHostToDevice, s), "dp_pack_hits: H2D");
            d_p64 = ctx->tmp[3].as<double>();
        }
        if (points) { CK(ctx->tmp[4].ensure((size_t)cap * 24 + 16), "dp_pack_hits: out"); d_pts = ctx->tmp[4].as<double>(); }
        if (colors) { CK(ctx->tmp[5].ensure((size_t)cap * 24 + 16), "dp_pack_hits: out"); d_col = ctx->tmp[5].as<double>(); }
        if (face_out || pixel_out) {
            CK(ctx->tmp[6].ensure((size_t)cap * 8 + 16), "dp_pack_hits: out");
            if (face_out) d_fo = ctx->tmp[6].as<int32_t>();
            if (pixel_out) d_po = ctx->tmp[6].as<uint32_t>() + cap;
        }
        if (intensity_out) { CK(ctx->tmp[7].ensure((size_t)cap * 8 + 16), "dp_pack_hits: out"); d_io = ctx->tmp[7].as<double>(); }
    }
    CK(ctx->cscratch.ensure(pack_scratch_bytes(n) + 64), "dp_pack_hits: scratch");
    CK(ctx->tmp[2].ensure(64), "dp_pack_hits: scalars");
    long long *d_counts = ctx->counts.as<long long>();
    CK(launch_pack_hits(d_in, dtype, d_face, d_pix, d_p64, n, T, ctx->jet.as<double>(), ctx->tmp[2].as<long long>(), d_pts, d_col,
                        d_fo, d_po, d_io, cap, ctx->cscratch.as<unsigned long long>(), d_counts, s),
       "dp_pack_hits: launch");
    CK(cudaMemcpyAsync(ctx->h_counts, d_counts, 8, cudaMemcpyDeviceToHost, s), "dp_pack_hits: count");
    CK(cudaStreamSynchronize(s), "dp_pack_hits: kernels");
    const int64_t total = ctx->h_counts[0];
    *m = total;
    const int64_t nc = total < cap ? total : cap;
    if (mem == DP_HOST && nc > 0) {
        if (points) CK(cudaMemcpyAsync(points, d_pts, (size_t)nc * 24, cudaMemcpyDeviceToHost, s), "dp_pack_hits: D2H");
        if (colors) CK(cudaMemcpyAsync(colors, d_col, (size_t)nc * 24, cudaMemcpyDeviceToHost, s), "dp_pack_hits: D2H");
        if (face_out) CK(cudaMemcpyAsync(face_out, d_fo, (size_t)nc * 4, cudaMemcpyDeviceToHost, s), "dp_pack_hits: D2H");
        if (pixel_out) CK(cudaMemcpyAsync(pixel_out, d_po, (size_t)nc * 4, cudaMemcpyDeviceToHost, s), "dp_pack_hits: D2H");
        if (intensity_out) CK(cudaMemcpyAsync(intensity_out, d_io, (size_t)nc * 8, cudaMemcpyDeviceToHost, s), "dp_pack_hits: D2H");
        CK(cudaStreamSynchronize(s), "dp_pack_hits: D2H");
    }
    if (total > cap) return fail(ctx, DP_E_NOMEM, "dp_pack_hits: capacity too small for the selected rays");
    return DP_OK;
}

int dp_set_ray_shard(dp_ctx *ctx, int rank, int world)
{
    if (!ctx) return DP_E_ARG;
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, DP_E_ARG, "dp_set_ray_shard: need 0 <= rank < world");
    if (rank != ctx->shard.rank || world != ctx->shard.world) ctx->order_np = 0;   // the learnt packet lists are per shard
    ctx->shard.rank = rank;
    ctx->shard.world = world;
    return DP_OK;
}

int dp_shard_slots(int rank, int world, int64_t n_rays, int64_t nframes, int H, int W, int64_t *lo, int64_t *hi)
{
    if (!lo || !hi || n_rays < 0 || nframes < 0 || H < 0 || W < 0 || world < 1 || rank < 0 || rank >= world) return DP_E_ARG;
    RayShard sh;
    sh.rank = rank;
    sh.world = world;
    shard_slots_host(n_rays, nframes * (int64_t)H * W, H, W, sh, lo, hi);
    return DP_OK;
}

int dp_pack_records(dp_ctx *ctx, const uint32_t *pixel, const float *t_hit, const int32_t *face, const float *point,
                    int64_t n, int64_t first, uint32_t *records, int64_t cap, int64_t *m, int64_t *m_async, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || first < 0 || cap < 0 || (n > 0 && (!face || !t_hit)) || (cap > 0 && !records) || (!m && !m_async))
        return fail(ctx, DP_E_ARG, "dp_pack_records: bad arguments");
    if (mem != DP_DEVICE) return fail(ctx, DP_E_ARG, "dp_pack_records: device buffers only (the per-ray outputs of dp_project)");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(ctx->tmp[7].ensure(pack_scratch_bytes(n) + 64), "dp_pack_records: scratch");
    long long *d_count = ctx->counts.as<long long>() + 4;
    long long *d_async = nullptr;
    if (m_async) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, m_async) == cudaSuccess && at.devicePointer) d_async = static_cast<long long *>(at.devicePointer);
        else { cudaGetLastError(); return fail(ctx, DP_E_ARG, "dp_pack_records: m_async must be device or pinned host memory"); }
    }
    CK(launch_pack_records(pixel, t_hit, face, point, n, first, records, cap, ctx->tmp[7].as<unsigned long long>(), d_count,
                           d_async, s),
       "dp_pack_records: launch");
    if (m) {
        CK(cudaMemcpyAsync(ctx->h_counts + 4, d_count, 8, cudaMemcpyDeviceToHost, s), "dp_pack_records: count");
        CK(cudaStreamSynchronize(s), "dp_pack_records: kernel");
        *m = ctx->h_counts[4];
        if (*m > cap) return fail(ctx, DP_E_NOMEM, "dp_pack_records: capacity too small for the hits");
    }
    return DP_OK;
}

// ---- exchange over peer-mapped memory -------------------------------------------------------------------------
int dp_peer_export(dp_ctx *ctx, int64_t record_bytes, int64_t result_rays, void *handle, int64_t *window_bytes)
{
    if (!ctx) return DP_E_ARG;
    if (record_bytes < 0 || result_rays < 0) return fail(ctx, DP_E_ARG, "dp_peer_export: negative size");
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_peer_export: no mesh (the window holds accumulator snapshots)");
    if (ctx->peer_opened) return fail(ctx, DP_E_STATE, "dp_peer_export: windows are open (dp_peer_close first)");
    DeviceGuard g(ctx->device);
    const size_t a256 = 255;
    const size_t stage = (ctx->accum_bytes + a256) & ~a256, rec = ((size_t)record_bytes + a256) & ~a256,
                 res = ((size_t)result_rays * 20 + a256) & ~a256;
    size_t off = sizeof(PeerCtl);
    PeerView &pv = ctx->peer;
    memset(&pv, 0, sizeof(pv));
    for (int k = 0; k < 2; ++k) { pv.stage_off[k] = off; off += stage; }
    for (int k = 0; k < 2; ++k) { pv.rec_off[k] = off; off += rec; }
    for (int k = 0; k < 2; ++k) { pv.res_off[k] = off; off += res; }
    pv.res_cap = (unsigned long long)result_rays;
    // a fresh allocation of its own (cudaMalloc, never sub-allocated): that is what an IPC handle names
    ctx->peer_win.release();
    CK(ctx->peer_win.ensure(off), "dp_peer_export: window");
    CK(cudaMemset(ctx->peer_win.p, 0, off), "dp_peer_export: zero");
    ctx->peer_stage_bytes = ctx->accum_bytes;
    ctx->peer_rec_bytes = (size_t)record_bytes;
    ctx->peer_win_bytes = off;
    ctx->peer_epoch[0] = ctx->peer_epoch[1] = 0;
    ctx->peer_exported = true;
    ctx->peer_res_slot = -1;
    if (handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == DP_PEER_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ctx->peer_win.p), "dp_peer_export: cudaIpcGetMemHandle");
        memcpy(handle, &h, sizeof(h));
    }
    if (window_bytes) *window_bytes = (int64_t)off;
    return DP_OK;
}

static int peer_finish_open(dp_ctx *ctx, int rank, int world)
{
    ctx->peer.rank = rank;
    ctx->peer.world = world;
    ctx->peer.win[rank] = ctx->peer_win.as<char>();
    // result arrays of the other ranks, per result slot, for the traversal's epilogue (device table)
    PeerOut po[2];
    memset(po, 0, sizeof(po));
    for (int k = 0; k < 2; ++k) {
        int n = 0;
        for (int p = 0; p < world; ++p) {
            if (p == rank) continue;
            char *r = ctx->peer.win[p] + ctx->peer.res_off[k];
            po[k].t_hit[n] = reinterpret_cast<float *>(r);
            po[k].face[n] = reinterpret_cast<int32_t *>(r + ctx->peer.res_cap * 4);
            po[k].point[n] = reinterpret_cast<float *>(r + ctx->peer.res_cap * 8);
            ++n;
        }
        po[k].n = n;
    }
    CK(ctx->peer_out.ensure(sizeof(po)), "dp_peer_open: table");
    CK(cudaMemcpy(ctx->peer_out.p, po, sizeof(po), cudaMemcpyHostToDevice), "dp_peer_open: table");
    ctx->peer_opened = true;
    return DP_OK;
}

int dp_peer_open(dp_ctx *ctx, int rank, int world, const void *handles)
{
    if (!ctx) return DP_E_ARG;
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !handles)
        return fail(ctx, DP_E_ARG, "dp_peer_open: need 0 <= rank < world <= 16 and the handles of all ranks");
    if (!ctx->peer_exported) return fail(ctx, DP_E_STATE, "dp_peer_open: dp_peer_export first");
    if (ctx->peer_opened) return fail(ctx, DP_E_STATE, "dp_peer_open: already open");
    DeviceGuard g(ctx->device);
    for (int p = 0; p < world; ++p) {
        ctx->peer_ipc[p] = false;
        if (p == rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)p * DP_PEER_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < p; ++q)
                if (ctx->peer_ipc[q]) { cudaIpcCloseMemHandle(ctx->peer.win[q]); ctx->peer_ipc[q] = false; }
            return fail(ctx, DP_E_CUDA, "dp_peer_open: cudaIpcOpenMemHandle (peer access between the devices of one node is needed)", e);
        }
        ctx->peer.win[p] = static_cast<char *>(ptr);
        ctx->peer_ipc[p] = true;
    }
    return peer_finish_open(ctx, rank, world);
}

int dp_peer_open_local(dp_ctx *ctx, int rank, int world, dp_ctx *const *peers)
{
    if (!ctx) return DP_E_ARG;
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !peers)
        return fail(ctx, DP_E_ARG, "dp_peer_open_local: need 0 <= rank < world <= 16 and the contexts of all ranks");
    if (!ctx->peer_exported) return fail(ctx, DP_E_STATE, "dp_peer_open_local: dp_peer_export first");
    if (ctx->peer_opened) return fail(ctx, DP_E_STATE, "dp_peer_open_local: already open");
    DeviceGuard g(ctx->device);
    for (int p = 0; p < world; ++p) {
        ctx->peer_ipc[p] = false;
        if (p == rank) continue;
        const dp_ctx *o = peers[p];
        if (!o || !o->peer_exported || o->peer_win_bytes != ctx->peer_win_bytes)
            return fail(ctx, DP_E_STATE, "dp_peer_open_local: every context needs an exported window of the same layout");
        if (o->device != ctx->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->device, o->device), "dp_peer_open_local");
            if (!can) return fail(ctx, DP_E_CUDA, "dp_peer_open_local: no peer access between the devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(ctx, DP_E_CUDA, "dp_peer_open_local: cudaDeviceEnablePeerAccess", e);
        }
        ctx->peer.win[p] = o->peer_win.as<char>();
    }
    return peer_finish_open(ctx, rank, world);
}

int dp_peer_close(dp_ctx *ctx)
{
    if (!ctx) return DP_E_ARG;
    DeviceGuard g(ctx->device);
    if (ctx->peer_opened) cudaDeviceSynchronize();
    for (int p = 0; p < PEER_MAX; ++p) {
        if (ctx->peer_ipc[p]) cudaIpcCloseMemHandle(ctx->peer.win[p]);
        ctx->peer_ipc[p] = false;
        ctx->peer.win[p] = nullptr;
    }
    ctx->peer_opened = false;
    ctx->peer_exported = false;     // the window stays allocated until the next export / dp_destroy (peers may still map it)
    ctx->peer_res_slot = -1;
    return DP_OK;
}

int dp_peer_window(dp_ctx *ctx, int what, int slot, void **ptr, int64_t *bytes)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->peer_exported) return fail(ctx, DP_E_STATE, "dp_peer_window: no exchange window");
    if (slot < 0 || slot > 1) return fail(ctx, DP_E_ARG, "dp_peer_window: slot is 0 or 1");
    char *w = ctx->peer_win.as<char>();
    const PeerView &pv = ctx->peer;
    void *p = nullptr;
    int64_t b = 0;
    switch (what) {
    case DP_PEER_STAGE: p = w + pv.stage_off[slot]; b = (int64_t)ctx->peer_stage_bytes; break;
    case DP_PEER_RECORDS: p = w + pv.rec_off[slot]; b = (int64_t)ctx->peer_rec_bytes; break;
    case DP_PEER_REC_COUNT: p = &reinterpret_cast<PeerCtl *>(w)->rec_count[slot]; b = 8; break;
    case DP_PEER_T_HIT: p = w + pv.res_off[slot]; b = (int64_t)pv.res_cap * 4; break;
    case DP_PEER_FACE: p = w + pv.res_off[slot] + pv.res_cap * 4; b = (int64_t)pv.res_cap * 4; break;
    case DP_PEER_POINT: p = w + pv.res_off[slot] + pv.res_cap * 8; b = (int64_t)pv.res_cap * 12; break;
    default: return fail(ctx, DP_E_ARG, "dp_peer_window: unknown region");
    }
    if (ptr) *ptr = p;
    if (bytes) *bytes = b;
    return DP_OK;
}

int dp_peer_snapshot(dp_ctx *ctx, int slot, int reset, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->peer_exported) return fail(ctx, DP_E_STATE, "dp_peer_snapshot: no exchange window");
    if (slot < 0 || slot > 1) return fail(ctx, DP_E_ARG, "dp_peer_snapshot: slot is 0 or 1");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(launch_vertex_max(ctx->fmax.as<uint32_t>(), ctx->F.as<int32_t>(), ctx->nF, ctx->vmax.as<uint32_t>(), s), "dp_peer_snapshot: vertex maxima");
    CK(launch_peer_snapshot(ctx->accum.p, ctx->peer_win.as<char>() + ctx->peer.stage_off[slot], ctx->accum_bytes, reset != 0, s),
       "dp_peer_snapshot: launch");
    return DP_OK;
}

int dp_peer_combine(dp_ctx *ctx, int slot, void *total, int gather_root, uint32_t *gathered, int64_t cap_rows, int row_words,
                    int64_t *m_async, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->peer_opened) return fail(ctx, DP_E_STATE, "dp_peer_combine: windows are not open (dp_peer_open)");
    if (slot < 0 || slot > 1 || cap_rows < 0 || (gathered && row_words < 1) || gather_root >= ctx->peer.world)
        return fail(ctx, DP_E_ARG, "dp_peer_combine: bad arguments");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    long long *d_async = nullptr;
    if (m_async) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, m_async) == cudaSuccess && at.devicePointer) d_async = static_cast<long long *>(at.devicePointer);
        else { cudaGetLastError(); return fail(ctx, DP_E_ARG, "dp_peer_combine: m_async must be device or pinned host memory"); }
    }
    const unsigned long long epoch = ++ctx->peer_epoch[0];
    CK(launch_peer_combine(ctx->peer, slot, epoch, total, ctx->accum_bytes, (size_t)(ctx->fmax.as<char>() - ctx->accum.as<char>()),
                           total != nullptr, gather_root, gathered, cap_rows, row_words, nullptr, d_async, s),
       "dp_peer_combine: launch");
    return DP_OK;
}

int dp_peer_results(dp_ctx *ctx, int slot, int with_points)
{
    if (!ctx) return DP_E_ARG;
    if (slot < -1 || slot > 1) return fail(ctx, DP_E_ARG, "dp_peer_results: slot is 0, 1 or -1 (off)");
    if (slot >= 0 && !ctx->peer_opened) return fail(ctx, DP_E_STATE, "dp_peer_results: windows are not open (dp_peer_open)");
    if (slot >= 0 && ctx->peer.res_cap == 0) return fail(ctx, DP_E_STATE, "dp_peer_results: the window was exported without result arrays");
    ctx->peer_res_slot = slot;
    ctx->peer_res_points = with_points != 0;
    return DP_OK;
}

int dp_peer_status(dp_ctx *ctx, int *error)
{
    if (!ctx || !error) return DP_E_ARG;
    *error = 0;
    if (!ctx->peer_exported) return DP_OK;
    DeviceGuard g(ctx->device);
    unsigned e = 0;
    CK(cudaMemcpy(&e, &ctx->peer_win.as<PeerCtl>()->error, 4, cudaMemcpyDeviceToHost), "dp_peer_status");
    *error = (int)e;
    return DP_OK;
}

int dp_accum_layout(dp_ctx *ctx, void **base, int64_t *hist_off, int64_t *fmax_off, int64_t *vmax_off, int64_t *bytes)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_accum_layout: no mesh");
    if (base) *base = ctx->accum.p;
    if (hist_off) *hist_off = 0;
    if (fmax_off) *fmax_off = (int64_t)(ctx->fmax.as<char>() - ctx->accum.as<char>());
    if (vmax_off) *vmax_off = (int64_t)(ctx->vmax.as<char>() - ctx->accum.as<char>());
    if (bytes) *bytes = (int64_t)ctx->accum_bytes;
    return DP_OK;
}

int dp_accum_reset(dp_ctx *ctx, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_accum_reset: no mesh");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(cudaMemsetAsync(ctx->accum.p, 0, ctx->accum_bytes, s), "dp_accum_reset");
    return DP_OK;
}

int dp_accum_flush(dp_ctx *ctx, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_accum_flush: no mesh");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(launch_vertex_max(ctx->fmax.as<uint32_t>(), ctx->F.as<int32_t>(), ctx->nF, ctx->vmax.as<uint32_t>(), s), "dp_accum_flush");
    return DP_OK;
}

int dp_accum_get(dp_ctx *ctx, int32_t *hist, float *fmax, float *vmax, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_accum_get: no mesh");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (vmax) {
        const int rc = dp_accum_flush(ctx, stream);
        if (rc != DP_OK) return rc;
    }
    const cudaMemcpyKind kind = mem == DP_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (hist && ctx->nF) CK(cudaMemcpyAsync(hist, ctx->hist.p, (size_t)ctx->nF * 4, kind, s), "dp_accum_get");
    if (fmax && ctx->nF) CK(cudaMemcpyAsync(fmax, ctx->fmax.p, (size_t)ctx->nF * 4, kind, s), "dp_accum_get");
    if (vmax && ctx->nV) CK(cudaMemcpyAsync(vmax, ctx->vmax.p, (size_t)ctx->nV * 4, kind, s), "dp_accum_get");
    if (mem == DP_HOST) CK(cudaStreamSynchronize(s), "dp_accum_get");
    return DP_OK;
}

int dp_accum_device_ptrs(dp_ctx *ctx, int32_t **hist, float **fmax, float **vmax)
{
    if (!ctx) return DP_E_ARG;
    if (!ctx->has_mesh) return fail(ctx, DP_E_STATE, "dp_accum_device_ptrs: no mesh");
    if (hist) *hist = ctx->hist.as<int32_t>();
    if (fmax) *fmax = ctx->fmax.as<float>();
    if (vmax) *vmax = ctx->vmax.as<float>();
    return DP_OK;
}

int dp_set_timing(dp_ctx *ctx, int enable)
{
    if (!ctx) return DP_E_ARG;
    ctx->timing_on = enable != 0;
    if (!ctx->timing_on) ctx->timings_valid = false;
    return DP_OK;
}

int dp_set_stats(dp_ctx *ctx, int enable)
{
    if (!ctx) return DP_E_ARG;
    ctx->stats_on = enable != 0;
    return DP_OK;
}

int dp_get_stats(dp_ctx *ctx, dp_stats *out)
{
    if (!ctx || !out) return fail(ctx, DP_E_ARG, "dp_get_stats: bad arguments");
    DeviceGuard g(ctx->device);
    CK(cudaDeviceSynchronize(), "dp_get_stats");
    TraceStats st{};
    CK(cudaMemcpy(&st, ctx->stats.p, sizeof(st), cudaMemcpyDeviceToHost), "dp_get_stats");
    memset(out, 0, sizeof(*out));
    out->rays = (int64_t)st.rays;
    out->hits = (int64_t)st.hits;
    out->nodes_fetched = (int64_t)st.nodes;
    out->tris_tested = (int64_t)st.tris;
    out->n_wide_nodes = ctx->has_bvh ? ctx->obj.n_nodes : 0;
    out->n_tris = ctx->has_bvh ? ctx->obj.n_tris : 0;
    out->wide_depth = ctx->has_bvh ? ctx->topo.n_levels : 0;
    out->last_build_ms = ctx->build_ms;
    if (ctx->has_cam && ctx->refit_ms < 0.0f) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]) == cudaSuccess) ctx->refit_ms = ms;
    }
    out->last_refit_ms = ctx->refit_ms < 0.0f ? 0.0f : ctx->refit_ms;
    return DP_OK;
}

int dp_last_timings(dp_ctx *ctx, float *ms4)
{
    if (!ctx || !ms4) return fail(ctx, DP_E_ARG, "dp_last_timings: bad arguments");
    if (!ctx->timings_valid) return fail(ctx, DP_E_STATE, "dp_last_timings: no dp_project call with timing enabled yet (dp_set_timing)");
    DeviceGuard g(ctx->device);
    CK(cudaEventSynchronize(ctx->ev[2]), "dp_last_timings");
    CK(cudaEventElapsedTime(&ms4[0], ctx->ev[0], ctx->ev[1]), "dp_last_timings");
    CK(cudaEventElapsedTime(&ms4[1], ctx->ev[1], ctx->ev[7]), "dp_last_timings");
    CK(cudaEventElapsedTime(&ms4[2], ctx->ev[7], ctx->ev[3]), "dp_last_timings");
    CK(cudaEventElapsedTime(&ms4[3], ctx->ev[0], ctx->ev[2]), "dp_last_timings");
    return DP_OK;
}

int dp_debug_dump_bvh(dp_ctx *ctx, int frame, void *nodes, int64_t *n_nodes, void *tris, int64_t *n_tris)
{
    if (!ctx) return DP_E_ARG;
    if (frame == DP_FRAME_OBJECT ? !ctx->has_bvh : !ctx->has_cam) return fail(ctx, DP_E_STATE, "dp_debug_dump_bvh: no BVH");
    DeviceGuard g(ctx->device);
    const BvhStorage &b = frame == DP_FRAME_OBJECT ? ctx->obj : ctx->cam;
    CK(cudaDeviceSynchronize(), "dp_debug_dump_bvh");
    if (n_nodes) *n_nodes = b.n_nodes;
    if (n_tris) *n_tris = b.n_tris;
    if (nodes) CK(cudaMemcpy(nodes, b.nodes, (size_t)b.n_nodes * sizeof(WideNode), cudaMemcpyDeviceToHost), "dp_debug_dump_bvh");
    if (tris && b.n_tris) CK(cudaMemcpy(tris, b.tris, (size_t)b.n_tris * sizeof(TriRec), cudaMemcpyDeviceToHost), "dp_debug_dump_bvh");
    return DP_OK;
}

int dp_debug_ray_nodes(dp_ctx *ctx, uint32_t *counts, int64_t n)
{
    if (!ctx || !counts || n < 0) return fail(ctx, DP_E_ARG, "dp_debug_ray_nodes: bad arguments");
    if (n > ctx->ray_nodes_n) return fail(ctx, DP_E_STATE, "dp_debug_ray_nodes: no counted dp_project launch of that size");
    DeviceGuard g(ctx->device);
    CK(cudaDeviceSynchronize(), "dp_debug_ray_nodes");
    if (n) CK(cudaMemcpy(counts, ctx->ray_nodes.p, (size_t)n * 4, cudaMemcpyDeviceToHost), "dp_debug_ray_nodes");
    return DP_OK;
}

int dp_debug_radix_sort(dp_ctx *ctx, uint32_t *keys, uint32_t *vals, int64_t n)
{
    if (!ctx || n < 0 || (n > 0 && (!keys || !vals))) return fail(ctx, DP_E_ARG, "dp_debug_radix_sort: bad arguments");
    if (n == 0) return DP_OK;
    DeviceGuard g(ctx->device);
    DevBuf k, v, kt, vt, tb;
    cudaError_t e = cudaSuccess;
    if ((e = k.ensure((size_t)n * 4)) == cudaSuccess && (e = v.ensure((size_t)n * 4)) == cudaSuccess &&
        (e = kt.ensure((size_t)n * 4)) == cudaSuccess && (e = vt.ensure((size_t)n * 4)) == cudaSuccess &&
        (e = tb.ensure(radix_table_entries(n) * 4)) == cudaSuccess) {
        bool in_tmp = false;
        if ((e = cudaMemcpy(k.p, keys, (size_t)n * 4, cudaMemcpyHostToDevice)) == cudaSuccess &&
            (e = cudaMemcpy(v.p, vals, (size_t)n * 4, cudaMemcpyHostToDevice)) == cudaSuccess &&
            (e = radix_sort_pairs(k.as<uint32_t>(), v.as<uint32_t>(), kt.as<uint32_t>(), vt.as<uint32_t>(), n,
                                  tb.as<uint32_t>(), nullptr, &in_tmp)) == cudaSuccess &&
            (e = cudaDeviceSynchronize()) == cudaSuccess) {
            e = cudaMemcpy(keys, in_tmp ? kt.p : k.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(vals, in_tmp ? vt.p : v.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
        }
    }
    k.release(); v.release(); kt.release(); vt.release(); tb.release();
    CK(e, "dp_debug_radix_sort");
    return DP_OK;
}

int dp_debug_morton(dp_ctx *ctx, uint32_t *codes)
{
    if (!ctx || !codes) return fail(ctx, DP_E_ARG, "dp_debug_morton: bad arguments");
    if (!ctx->has_bvh) return fail(ctx, DP_E_STATE, "dp_debug_morton: no BVH");
    DeviceGuard g(ctx->device);
    cudaError_t e = build_lbvh(ctx->V.as<float>(), ctx->nV, ctx->F.as<int32_t>(), ctx->nF, ctx->obj, ctx->topo,
                               &ctx->build_scratch, &ctx->build_scratch_bytes, codes, nullptr);
    CK(e, "dp_debug_morton");
    CK(cudaDeviceSynchronize(), "dp_debug_morton");
    ctx->has_cam = false;
    return DP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ point-to-plane ICP
namespace {
// x = A^-1 b for the symmetric 6 x 6 normal equations (Gaussian elimination, partial pivoting); returns det(A)
double solve6(double A[6][6], double b[6], double x[6])
{
    double det = 1.0;
    int perm_sign = 1;
    for (int c = 0; c < 6; ++c) {
        int p = c;
        for (int r = c + 1; r < 6; ++r)
            if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
        if (A[p][c] == 0.0) return 0.0;
        if (p != c) {
            for (int k = 0; k < 6; ++k) { const double t = A[c][k]; A[c][k] = A[p][k]; A[p][k] = t; }
            const double t = b[c]; b[c] = b[p]; b[p] = t;
            perm_sign = -perm_sign;
        }
        det *= A[c][c];
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int k = c; k < 6; ++k) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double acc = b[r];
        for (int k = r + 1; k < 6; ++k) acc -= A[r][k] * x[k];
        x[r] = acc / A[r][r];
    }
    return det * perm_sign;
}

void mat4_mul(const double *A, const double *B, double *C)
{
    double T[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += A[4 * r + k] * B[4 * k + c];
            T[4 * r + c] = acc;
        }
    memcpy(C, T, sizeof(T));
}

// Open3D TransformVector6dToMatrix4d: R = Rz(x2) * Ry(x1) * Rx(x0), t = x3..5
void se3_from_vector6(const double x[6], double *M)
{
    const double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]), cg = cos(x[2]), sg = sin(x[2]);
    const double Rx[16] = {1, 0, 0, 0, 0, ca, -sa, 0, 0, sa, ca, 0, 0, 0, 0, 1};
    const double Ry[16] = {cb, 0, sb, 0, 0, 1, 0, 0, -sb, 0, cb, 0, 0, 0, 0, 1};
    const double Rz[16] = {cg, -sg, 0, 0, sg, cg, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double T[16];
    mat4_mul(Ry, Rx, T);
    mat4_mul(Rz, T, M);
    M[3] = x[3]; M[7] = x[4]; M[11] = x[5];
}
}  // namespace

extern "C" int dp_icp_point_to_plane(dp_ctx *ctx, const double *source, int64_t n, const double *target,
                                     const double *target_normals, int64_t m, double max_correspondence_distance,
                                     const double *init, int max_iteration, double relative_fitness, double relative_rmse,
                                     double *T_out, double *fitness, double *inlier_rmse, int *iterations,
                                     int32_t *correspondence, int mem, void *stream)
{
    if (!ctx) return DP_E_ARG;
    if (n < 0 || m < 0 || max_iteration < 0 || !T_out || !(max_correspondence_distance > 0.0) || (n > 0 && !source) ||
        (m > 0 && (!target || !target_normals)))
        return fail(ctx, DP_E_ARG, "dp_icp_point_to_plane: bad arguments");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static const double ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double T[16];
    memcpy(T, init ? init : ident, sizeof(T));
    // working copy of the source cloud (Open3D transforms a copy in place, iteration after iteration)
    CK(ctx->tmp[0].ensure((size_t)n * 24 + 16), "dp_icp_point_to_plane: cloud");
    double *d_src = ctx->tmp[0].as<double>();
    const double *d_tp = target, *d_tn = target_normals;
    int32_t *d_corr = correspondence;
    if (n) CK(cudaMemcpyAsync(d_src, source, (size_t)n * 24, mem == DP_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s),
              "dp_icp_point_to_plane: source");
    if (mem == DP_HOST) {
        CK(ctx->tmp[1].ensure((size_t)m * 24 + 16), "dp_icp_point_to_plane: target");
        CK(ctx->tmp[2].ensure((size_t)m * 24 + 16), "dp_icp_point_to_plane: normals");
        if (m) {
            CK(cudaMemcpyAsync(ctx->tmp[1].p, target, (size_t)m * 24, cudaMemcpyHostToDevice, s), "dp_icp_point_to_plane: H2D");
            CK(cudaMemcpyAsync(ctx->tmp[2].p, target_normals, (size_t)m * 24, cudaMemcpyHostToDevice, s), "dp_icp_point_to_plane: H2D");
        }
        d_tp = ctx->tmp[1].as<double>();
        d_tn = ctx->tmp[2].as<double>();
        if (correspondence) {
            CK(ctx->tmp[5].ensure((size_t)n * 4 + 16), "dp_icp_point_to_plane: corr");
            d_corr = ctx->tmp[5].as<int32_t>();
        }
    }
    CK(ctx->tmp[3].ensure(icp_partial_doubles(n) * 8 + 16), "dp_icp_point_to_plane: partials");
    CK(ctx->tmp[4].ensure(64 * 8), "dp_icp_point_to_plane: sums");
    double *d_partial = ctx->tmp[3].as<double>(), *d_sums = ctx->tmp[4].as<double>();
    double sums[29];
    bool identity_init = true;
    for (int i = 0; i < 16; ++i) identity_init = identity_init && T[i] == ident[i];
    // Targets of some size get a uniform grid (cells >= the search radius, built once per call); DP_ICP_GRID=0 keeps the
    // tiled scan of the whole target (same correspondences and sums bit for bit, kept for A/B measurements).
    IcpGridView gv;
    gv.tps = nullptr;
    const char *knob = getenv("DP_ICP_GRID");                        // read per call: tests flip it within one process
    const int use_grid = knob ? atoi(knob) : 1;
    if (use_grid && n > 0 && m >= ICP_GRID_MIN_POINTS && m < (int64_t)1 << 31) {
        CK(ctx->tmp[6].ensure(icp_grid_bytes(m)), "dp_icp_point_to_plane: grid");
        CK(icp_grid_build(d_tp, d_tn, m, max_correspondence_distance, ICP_GRID_MIN_CELLS, ctx->tmp[6].p, &gv, s), "dp_icp_point_to_plane: grid build");
    }
    auto evaluate = [&](const double *update) -> int {
        cudaError_t e = gv.tps ? launch_icp_step_grid(d_src, n, gv, max_correspondence_distance, update, d_corr, d_partial, d_sums, s)
                               : launch_icp_step(d_src, n, d_tp, d_tn, m, max_correspondence_distance, update, d_corr, d_partial,
                                                 d_sums, s);
        if (e != cudaSuccess) return fail(ctx, DP_E_CUDA, "dp_icp_point_to_plane: launch", e);
        if ((e = cudaMemcpyAsync(sums, d_sums, sizeof(sums), cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
            (e = cudaStreamSynchronize(s)) != cudaSuccess)
            return fail(ctx, DP_E_CUDA, "dp_icp_point_to_plane: kernels", e);
        return DP_OK;
    };
    int rc = evaluate(identity_init ? nullptr : T);
    if (rc != DP_OK) return rc;
    double fit = n > 0 ? sums[0] / (double)n : 0.0, rmse = sums[0] > 0.0 ? sqrt(sums[1] / sums[0]) : 0.0;
    int it = 0;
    while (it < max_iteration) {
        double update[16];
        memcpy(update, ident, sizeof(update));
        if (sums[0] > 0.0) {
            double A[6][6], b[6], x[6] = {0, 0, 0, 0, 0, 0};
            int e = 2;
            for (int a = 0; a < 6; ++a)
                for (int c = a; c < 6; ++c) { A[a][c] = sums[e]; A[c][a] = sums[e]; ++e; }
            for (int a = 0; a < 6; ++a) b[a] = -sums[23 + a];
            const double det = solve6(A, b, x);
            bool okx = fabs(det) >= 1e-6 && det == det && fabs(det) <= 1.7e308;     // Open3D's determinant check
            for (int a = 0; a < 6; ++a) okx = okx && x[a] == x[a];
            if (okx) se3_from_vector6(x, update);
        }
        mat4_mul(update, T, T);
        const double pf = fit, pr = rmse;
        rc = evaluate(update);
        if (rc != DP_OK) return rc;
        ++it;
        fit = n > 0 ? sums[0] / (double)n : 0.0;
        rmse = sums[0] > 0.0 ? sqrt(sums[1] / sums[0]) : 0.0;
        if (fabs(pf - fit) < relative_fitness && fabs(pr - rmse) < relative_rmse) break;
    }
    memcpy(T_out, T, sizeof(T));
    if (fitness) *fitness = fit;
    if (inlier_rmse) *inlier_rmse = rmse;
    if (iterations) *iterations = it;
    if (mem == DP_HOST && correspondence && n) {
        CK(cudaMemcpyAsync(correspondence, d_corr, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "dp_icp_point_to_plane: D2H");
        CK(cudaStreamSynchronize(s), "dp_icp_point_to_plane: D2H");
    }
    return DP_OK;
}
