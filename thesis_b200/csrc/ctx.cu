// ctx.cu -- host side of the C ABI (include/rbpf_b200.h): handle, device
// memory, the write-path LUT, launch sequencing.  No torch types, no CPU
// fallback: every compute entry point enqueues CUDA kernels.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/rbpf_b200.h"
#include "common.cuh"

struct rbpf_ctx {
    rbpf_config cfg;
    RbCtx d;                       // device pointers + dims, passed to kernels by value
    cudaStream_t stream;
    std::string err;
    std::vector<void *> allocs;
    double *d_px, *d_py, *d_dist;  // scan
    float4 *d_beamf;               // the same beams in float32 for the weight stage: (px, py, inside the weight stage's range gate, 0)
    double *d_rot;                 // rotation table
    uint32_t *d_lutx, *d_luty, *d_clut, *d_rlut;
    double *d_prev;                // 2 * RB_MAXB: previous scan endpoints (x then y)
    double *h_prev;                // pinned staging
    double *d_z;                   // N*K*3 host-supplied normals
    double *d_u01;
    double *d_tile;                // 800*800 export buffer
    int *d_slice;                  // 29*29 debug slice
    unsigned long long *d_refstats;
    double *h_scan;                // pinned staging: 2 slots x (px, py, dist | prev x, prev y | float4 beams), used alternately
    cudaEvent_t stage_ev[2];       // "the copies out of slot i have completed"
    int stage_slot;
    int have_scan;
    // optional per-stage CUDA-event timing of rbpf_step
    int *d_mg_slots;               // 2*N staging ints (migration)
    uint32_t *d_mg_mark, *d_mg_list;
    int *d_mg_count;
    int mg_n, mg_tiles;
    uint32_t *d_pull_mark = nullptr;   // world x pool_tiles claim table of the pull migration (allocated on first use)
    unsigned char *d_pull_rank = nullptr;
    uint32_t *d_pull_local = nullptr;  // local sub-tile of every claim
    bool pulled_pending = false;       // rbpf_migrate_pull_async: the payload copies run on another stream ...
    cudaEvent_t ev_alloc = nullptr, ev_copied = nullptr;   // ... between these two events
    void *phys[9];                 // pool, pt x2, pose x2, cov x2, exists x2 as allocated (index 1 + 2*k + parity)
    int parity;                    // which of the double buffers is current (flips with every commit)
    bool refs_pending = false;     // rbpf_resample_apply_local_deferred: the reference-count pass has not been launched yet
    cudaEvent_t refs_gate = nullptr; // ... and has to wait for this event (every peer has finished pulling)
    uint32_t *refs_pt = nullptr;   // ... on the page tables of the particles that were resampled
    struct PeerMap { bool attached = false, ipc = false; void *base[9] = {}; };
    std::vector<PeerMap> peers;    // by rank
    std::vector<cudaEvent_t> tev;  // (RB_NSTAGES + 1) events per recorded step
    int t_max_steps, t_steps;
    // deferred error reporting of rbpf_step: the flags are copied to pinned memory after every step
    // and looked at two steps later (the host stays one step ahead of the device)
    RbFlags *h_flags = nullptr;    // 2 pinned slots
    cudaEvent_t flag_ev[2] = {nullptr, nullptr};
    int flag_pending[2] = {0, 0};
    int last_adj = 0;              // mode of the last match (rbpf_get_match_slice re-runs it)
    void *ckpt_host = nullptr;     // pinned bounce buffer of the checkpoint calls, allocated on first use
    // device-side snapshot (rbpf_snapshot / rbpf_restore): shadow copies of every mutable buffer
    struct Snap { void *live, *shadow; size_t bytes; };
    std::vector<Snap> snap;
    int snap_parity = 0, snap_use_dup = 0, snap_valid = 0;
    unsigned long long snap_step_no = 0;
};

#define RB_STAGE_DOUBLES (7 * RB_MAXB)   // px, py, dist, prev x, prev y, RB_MAXB float4 (= 2 RB_MAXB doubles)
#define RB_NSTAGES 8               // set_scan, match, weight, raycast_prepare, raycast_cast, weight_fallback, resample_plan, resample_apply

#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
            return RBPF_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

template <typename T>
static cudaError_t dalloc(rbpf_ctx *h, T **p, size_t n)
{
    void *v = nullptr;
    cudaError_t e = cudaMalloc(&v, n * sizeof(T) > 0 ? n * sizeof(T) : 16);
    if (e == cudaSuccess) { h->allocs.push_back(v); *p = (T *)v; }
    return e;
}

static double rot_step_host() { return acos(1.0 - (RB_CS * RB_CS) / (2.0 * RB_MATCH_MAX_R * RB_MATCH_MAX_R)); }
static int rot_count_host() { return (int)floor((M_PI / 6.0) / rot_step_host()); }

extern "C" double rbpf_rot_step(void) { return rot_step_host(); }
extern "C" int32_t rbpf_rot_count(void) { return rot_count_host(); }

// Write-path LUT: lattice cell k -> storage coordinate, replaying
// int((k*0.05 - c)/0.05 + 400.0) of GridMap.set_*_pos (gridmap.py:92-95) for the
// tile that contains k*0.05 (hybridmap.py:44-45,193-208).  The float64 result is
// one cell low for some k; negative indices wrap like ndarray[-1].
// Entries are packed for the ray-cast inner loop:
//   RB_LUT_OFF = this axis' part of the byte offset inside the sub-tile (blocked
//   layout, RB_OFF_X / RB_OFF_Y), RB_LUT_SUB = sub-tile index along the axis,
//   RB_LUT_TILE = reference-tile index along the axis, bits 30 / 31 = this lattice
//   cell shares its storage cell with k+1 / k-1 (the aliasing the ray-cast has to
//   apply in order).
static void build_lut(int h, std::vector<uint32_t> &lut, bool y_axis)
{
    const int n = 800 * (2 * h + 1);
    lut.resize(n);
    for (int q = 0; q < n; q++) {
        const int k = q - 800 * h - 400;
        volatile double X = (double)k * RB_CS;
        int t = (int)floor((double)(k + 400) / 800.0);
        for (int tt = t - 1; tt <= t + 1; tt++) {                 // containment on the 40 m lattice
            double c = 40.0 * tt;
            if (X >= c - 20.0 && X < c + 20.0) { t = tt; break; }
        }
        if (t < -h) t = -h;
        if (t > h) t = h;
        volatile double rel = X - 40.0 * t;
        volatile double qd = rel / RB_CS;
        int idx = (int)(qd + 400.0);
        if (idx < 0) idx += RB_DIM;
        if (idx >= RB_DIM) idx = RB_DIM - 1;
        const int u = 800 * (t + h) + idx;
        const int o = u % RB_SUB;
        lut[q] = (uint32_t)(y_axis ? RB_OFF_Y(o) : RB_OFF_X(o)) | ((uint32_t)(u / RB_SUB) << 15) | ((uint32_t)(u / RB_DIM) << 24);
    }
    for (int q = 0; q + 1 < n; q++)
        if ((lut[q] & 0x3FFFFFFFu) == (lut[q + 1] & 0x3FFFFFFFu)) {
            lut[q] |= 1u << RB_LUT_NEXT_BIT;
            lut[q + 1] |= 1u << RB_LUT_PREV_BIT;
        }
}

extern "C" const char *rbpf_last_error(rbpf_handle h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int rbpf_destroy(rbpf_handle h)
{
    if (!h) return RBPF_ERR_ARG;
    cudaSetDevice(h->cfg.device);
    if (h->pulled_pending) cudaEventSynchronize(h->ev_copied);   // payload copies of the last pull still read and write our buffers
    cudaStreamSynchronize(h->stream);
    for (auto &pm : h->peers)
        if (pm.attached && pm.ipc)
            for (void *b : pm.base) cudaIpcCloseMemHandle(b);
    if (h->ev_alloc) cudaEventDestroy(h->ev_alloc);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    for (void *p : h->allocs) cudaFree(p);
    for (auto &s2 : h->snap) cudaFree(s2.shadow);
    for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) cudaEventDestroy(h->stage_ev[i]);
    if (h->h_scan) cudaFreeHost(h->h_scan);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->ckpt_host) cudaFreeHost(h->ckpt_host);
    for (int i = 0; i < 2; i++) if (h->flag_ev[i]) cudaEventDestroy(h->flag_ev[i]);
    delete h;
    return RBPF_OK;
}

extern "C" int rbpf_create(const rbpf_config *cfg, rbpf_handle *out)
{
    if (!cfg || !out) return RBPF_ERR_ARG;
    *out = nullptr;
    rbpf_ctx *h = new rbpf_ctx();
    h->cfg = *cfg;
    h->h_scan = nullptr;
    h->have_scan = 0;
    h->t_max_steps = 0;
    h->t_steps = 0;
    auto fail = [&](int code, const std::string &msg) {
        static std::string last;
        last = msg;
        fprintf(stderr, "rbpf_create: %s\n", msg.c_str());
        for (void *p : h->allocs) cudaFree(p);
        if (h->h_scan) cudaFreeHost(h->h_scan);
        if (h->h_flags) cudaFreeHost(h->h_flags);
        for (int i = 0; i < 2; i++) if (h->flag_ev[i]) cudaEventDestroy(h->flag_ev[i]);
        delete h;
        return code;
    };
    if (cfg->n_particles < 1 || cfg->n_beams < 1 || cfg->n_beams > RB_MAXB || cfg->n_samples < 1 ||
        cfg->n_samples > RB_MAXK || cfg->world_tiles_x < 1 || cfg->world_tiles_y < 1 ||
        (cfg->world_tiles_x & 1) == 0 || (cfg->world_tiles_y & 1) == 0 ||
        cfg->world_tiles_x * cfg->world_tiles_y > 64 || cfg->pool_subtiles < 1 || cfg->world < 1 || cfg->rank < 0 ||
        cfg->rank >= cfg->world)
        return fail(RBPF_ERR_ARG, "bad configuration");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || cfg->device < 0 || cfg->device >= ndev)
        return fail(RBPF_ERR_CUDA, "no usable CUDA device (there is no CPU fallback)");
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(RBPF_ERR_CUDA, "cudaSetDevice failed");
    h->stream = (cudaStream_t)(uintptr_t)cfg->stream;

    RbCtx &d = h->d;
    memset(&d, 0, sizeof(d));
    d.N = cfg->n_particles; d.B = cfg->n_beams; d.K = cfg->n_samples;
    d.rank = cfg->rank; d.world = cfg->world; d.n_global = cfg->n_particles * cfg->world;
    d.tiles_x = cfg->world_tiles_x; d.tiles_y = cfg->world_tiles_y;
    d.txh = (d.tiles_x - 1) / 2; d.tyh = (d.tiles_y - 1) / 2;
    d.subs_x = d.tiles_x * RB_SUBS_PER_TILE; d.subs_y = d.tiles_y * RB_SUBS_PER_TILE;
    d.nsub = d.subs_x * d.subs_y;
    d.ux_max = d.tiles_x * RB_DIM; d.uy_max = d.tiles_y * RB_DIM;
    d.pool_tiles = cfg->pool_subtiles;
    d.seed = cfg->seed;
    d.refine = (cfg->flags & RBPF_FLAG_NDT_REFINE) ? 1 : 0;
    d.step_no = 0;
    d.nk = rot_count_host();
    d.rot_step = rot_step_host();

    const size_t N = (size_t)d.N;
    cudaError_t e = cudaSuccess;
#define A(ptr, n) if (e == cudaSuccess) e = dalloc(h, &(ptr), (size_t)(n))
    A(d.pool, ((size_t)d.pool_tiles + 1) * RB_SUB_BYTES);              // + the sink sub-tile of the cast kernel (never referenced by a page table)
    A(d.refcnt, d.pool_tiles);
    A(d.free_list, d.pool_tiles);
    A(d.free_count, 4);
    A(d.pose, N * 3); A(d.pose2, N * 3);
    A(d.cov, N * 9); A(d.cov2, N * 9);
    A(d.weight, N);
    A(d.pt, N * d.nsub); A(d.pt2, N * d.nsub);
    A(d.exists, N); A(d.exists2, N);
    A(h->d_px, RB_MAXB); A(h->d_py, RB_MAXB); A(h->d_dist, RB_MAXB); A(h->d_beamf, RB_MAXB);
    A(h->d_rot, 2 * (2 * d.nk + 1));
    A(h->d_lutx, 800 * d.tiles_x);
    A(h->d_luty, 800 * d.tiles_y);
    A(h->d_clut, 800 * (d.tiles_x + d.tiles_y));
    A(h->d_rlut, 800 * (d.tiles_x + d.tiles_y));
    A(d.cast_work, 4);
    A(d.cast_done, N);
    A(d.pulled, N);
    A(d.m_pose, N * 3); A(d.m_cov, N * 9); A(d.m_score, N); A(d.m_valid, N); A(d.m_best, N * 4); A(d.m_refine, N * 2);
    A(d.w_all, d.n_global); A(d.plan_scal, 4); A(d.ancestors, d.n_global); A(d.mult, N); A(d.dup_of, N);
    A(d.stats, 1); A(d.flags, 1);
    A(h->d_z, N * d.K * 3);
    A(h->d_u01, 2);
    A(h->d_tile, RB_DIM * RB_DIM);
    A(h->d_prev, 2 * RB_MAXB);
    A(h->d_slice, RB_SLICE_W * RB_SLICE_W);
    A(h->d_refstats, 8);
    A(h->d_mg_slots, 2 * N);
    A(h->d_mg_mark, d.pool_tiles);
    A(h->d_mg_list, d.pool_tiles);
    A(h->d_mg_count, 4);
#undef A
    if (e != cudaSuccess) return fail(RBPF_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    {
        void *phys[9] = {d.pool, d.pt, d.pt2, d.pose, d.pose2, d.cov, d.cov2, d.exists, d.exists2};
        memcpy(h->phys, phys, sizeof(phys));
        h->parity = 0;
    }
    if (cudaMallocHost((void **)&h->h_scan, 2 * RB_STAGE_DOUBLES * sizeof(double)) != cudaSuccess)
        return fail(RBPF_ERR_CUDA, "cudaMallocHost failed");
    if (cudaMallocHost((void **)&h->h_flags, 2 * sizeof(RbFlags)) != cudaSuccess)
        return fail(RBPF_ERR_CUDA, "cudaMallocHost failed");
    memset(h->h_flags, 0, 2 * sizeof(RbFlags));
    for (int i = 0; i < 2; i++)
        if (cudaEventCreateWithFlags(&h->flag_ev[i], cudaEventDisableTiming) != cudaSuccess)
            return fail(RBPF_ERR_CUDA, "cudaEventCreate failed");
    h->h_prev = nullptr;
    h->stage_slot = 0;
    for (int i = 0; i < 2; i++)
        if (cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming) != cudaSuccess)
            return fail(RBPF_ERR_CUDA, "cudaEventCreate failed");
    d.prev_x = h->d_prev; d.prev_y = h->d_prev + RB_MAXB; d.n_prev = 0;
    d.px = h->d_px; d.py = h->d_py; d.dist = h->d_dist; d.beamf = h->d_beamf;
    d.rot_cs = h->d_rot;
    d.lutx = h->d_lutx;
    d.luty = h->d_luty;
    d.clut = h->d_clut;
    d.rlut = h->d_rlut;

    std::vector<double> rot(2 * (2 * d.nk + 1));
    for (int k = -d.nk; k <= d.nk; k++) {
        rot[2 * (k + d.nk)] = cos(k * d.rot_step);
        rot[2 * (k + d.nk) + 1] = sin(k * d.rot_step);
    }
    std::vector<uint32_t> lutx, luty;
    build_lut(d.txh, lutx, false);
    build_lut(d.tyh, luty, true);
    // cast LUT: the same storage coordinates re-packed so that the entries of the two axes add up to one word
    std::vector<uint32_t> clut(lutx.size() + luty.size());
    for (size_t q = 0; q < lutx.size(); q++)
        clut[q] = RB_LUT_OFF(lutx[q]) | (RB_LUT_SUB(lutx[q]) << 15) | (((lutx[q] >> RB_LUT_NEXT_BIT) & 1u) << 30) | (((lutx[q] >> RB_LUT_PREV_BIT) & 1u) << 31);
    for (size_t q = 0; q < luty.size(); q++)
        clut[lutx.size() + q] = RB_LUT_OFF(luty[q]) | ((RB_LUT_SUB(luty[q]) * (uint32_t)d.subs_x) << 15) |
                                (((luty[q] >> RB_LUT_NEXT_BIT) & 1u) << 28) | (((luty[q] >> RB_LUT_PREV_BIT) & 1u) << 29);
    // read LUT: page-table slot and byte offset of a storage coordinate, per axis, packed so that x + y is slot | offset << 12
    std::vector<uint32_t> rlut((size_t)d.ux_max + (size_t)d.uy_max);
    for (int ux = 0; ux < d.ux_max; ux++) rlut[ux] = (uint32_t)(ux / RB_SUB) | ((uint32_t)RB_OFF_X(ux % RB_SUB) << 12);
    for (int uy = 0; uy < d.uy_max; uy++)
        rlut[(size_t)d.ux_max + uy] = (uint32_t)((uy / RB_SUB) * d.subs_x) | ((uint32_t)RB_OFF_Y(uy % RB_SUB) << 12);
    if (cudaMemcpy(h->d_rlut, rlut.data(), rlut.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(RBPF_ERR_CUDA, "table upload failed");
    if (cudaMemcpy(h->d_rot, rot.data(), rot.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_clut, clut.data(), clut.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_lutx, lutx.data(), lutx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_luty, luty.data(), luty.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(RBPF_ERR_CUDA, "table upload failed");
    h->mg_n = h->mg_tiles = 0;
    cudaMemsetAsync(h->d_mg_mark, 0xFF, sizeof(uint32_t) * (size_t)d.pool_tiles, h->stream);
    cudaMemsetAsync(d.m_refine, 0, sizeof(int) * 2 * (size_t)N, h->stream);
    cudaMemsetAsync(d.mult, 0, sizeof(int) * (size_t)N, h->stream);      // zero between resamples (resample_refs_kernel resets it)
    rb_launch_init(d, h->stream);
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess)
        return fail(RBPF_ERR_CUDA, std::string("init: ") + cudaGetErrorString(e));
    *out = h;
    return RBPF_OK;
}

// Sticky device-side errors -> status.  They stay set until rbpf_clear_errors.
static int flags_status(rbpf_ctx *h, const RbFlags &f)
{
    if (f.pool_exhausted) {
        h->err = "tile pool exhausted (raise pool_subtiles): scans since then were not integrated and resampling is frozen";
        return RBPF_ERR_POOL;
    }
    if (f.resample_error_sticky) { h->err = "Incorrect number of resampled weights."; return RBPF_ERR_RESAMPLE; }   // main.py:67
    if (f.world_overflow) { h->err = "matcher window / ray-cast internal bound hit"; return RBPF_ERR_WORLD; }
    return RBPF_OK;
}

static int check_flags(rbpf_ctx *h)
{
    RbFlags f;
    CK(cudaMemcpyAsync(&f, h->d.flags, sizeof(f), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return flags_status(h, f);
}

static int flush_refs(rbpf_ctx *h);
static int flush_pulled(rbpf_ctx *h);

extern "C" int rbpf_clear_errors(rbpf_handle h)
{
    if (!h) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    RbFlags f;
    CK(cudaMemcpy(&f, h->d.flags, sizeof(f), cudaMemcpyDeviceToHost));
    f.pool_exhausted = f.world_overflow = f.resample_error = f.resample_error_sticky = 0;
    CK(cudaMemcpy(h->d.flags, &f, sizeof(f), cudaMemcpyHostToDevice));
    h->flag_pending[0] = h->flag_pending[1] = 0;
    h->err.clear();
    return RBPF_OK;
}

extern "C" int rbpf_synchronize(rbpf_handle h)
{
    if (!h) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return check_flags(h);
}

extern "C" int rbpf_set_scan(rbpf_handle h, const double *ranges, const double *angles, int32_t n_beams)
{
    if (!h || !ranges || !angles || n_beams < 1 || n_beams > RB_MAXB) { if (h) h->err = "set_scan: bad arguments"; return RBPF_ERR_ARG; }
    h->d.B = n_beams;                                            // loaders differ in beam count; buffers hold RB_MAXB
    CK(cudaSetDevice(h->cfg.device));
    // two pinned slots used alternately: only wait for the copies issued two calls ago
    h->stage_slot ^= 1;
    CK(cudaEventSynchronize(h->stage_ev[h->stage_slot]));
    double *px = h->h_scan + (size_t)h->stage_slot * RB_STAGE_DOUBLES, *py = px + RB_MAXB, *dist = py + RB_MAXB;
    for (int j = 0; j < n_beams; j++) {                          // Scan.__init__ lidar.py:76-80 (host libm, like the reference)
        px[j] = ranges[j] * cos(angles[j]);
        py[j] = ranges[j] * sin(angles[j]);
        dist[j] = sqrt(px[j] * px[j] + py[j] * py[j]);          // hybridmap.py:105,217 ; robot.py:129
    }
    CK(cudaMemcpyAsync(h->d_px, px, n_beams * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_py, py, n_beams * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_dist, dist, n_beams * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    float4 *bf = reinterpret_cast<float4 *>(px + 5 * RB_MAXB);   // weight stage: float32 beams + the range gate of robot.py:130
    for (int j = 0; j < n_beams; j++)
        bf[j] = make_float4((float)px[j], (float)py[j], (dist[j] < RB_W_MAX_R && dist[j] > RB_W_MIN_R) ? 1.0f : 0.0f, 0.0f);
    CK(cudaMemcpyAsync(h->d_beamf, bf, n_beams * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->stage_ev[h->stage_slot], h->stream));
    h->have_scan = 1;
    return RBPF_OK;
}

extern "C" int rbpf_motion(rbpf_handle h, int32_t family, const double *u, double dt, const double *par)
{
    if (!h || !u || family < 0 || family > 2) { if (h) h->err = "motion: bad arguments"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    rb_launch_motion(h->d, family, u, dt, par, h->stream);
    CK(cudaGetLastError());
    return RBPF_OK;
}

extern "C" int rbpf_scan_match(rbpf_handle h)
{
    if (!h || !h->have_scan) { if (h) h->err = "scan_match: no scan set"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    if (h->pulled_pending) {
        // the sub-tiles of migrated particles are still arriving on the copy stream: match the local particles now,
        // the migrated ones behind the copies
        rb_launch_match(h->d, 0, h->stream, 1, false);
        { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
        rb_launch_match(h->d, 0, h->stream, 2, true);
    } else {
        rb_launch_match(h->d, 0, h->stream);
    }
    h->last_adj = 0;
    CK(cudaGetLastError());
    return RBPF_OK;
}

extern "C" int rbpf_scan_match_adj(rbpf_handle h, const double *last_scan_xy, int32_t n_points)
{
    if (!h || !h->have_scan || !last_scan_xy || n_points < 0 || n_points > RB_MAXB) {
        if (h) h->err = "scan_match_adj: bad arguments";
        return RBPF_ERR_ARG;
    }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
    // shares the slot of the preceding rbpf_set_scan (its event is re-recorded after this copy)
    CK(cudaEventSynchronize(h->stage_ev[h->stage_slot]));
    double *hp = h->h_scan + (size_t)h->stage_slot * RB_STAGE_DOUBLES + 3 * RB_MAXB;
    for (int q = 0; q < n_points; q++) { hp[q] = last_scan_xy[2 * q]; hp[RB_MAXB + q] = last_scan_xy[2 * q + 1]; }
    CK(cudaMemcpyAsync(h->d_prev, hp, 2 * RB_MAXB * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->stage_ev[h->stage_slot], h->stream));
    h->d.n_prev = n_points;
    rb_launch_match(h->d, 1, h->stream);
    h->last_adj = 1;
    CK(cudaGetLastError());
    return RBPF_OK;
}

extern "C" int rbpf_weight(rbpf_handle h, const double *z)
{
    if (h) h->d.use_dup = 0;       // samples make duplicates diverge
    if (!h || !h->have_scan) { if (h) h->err = "weight: no scan set"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
    const double *zd = nullptr;
    if (z) {
        CK(cudaMemcpyAsync(h->d_z, z, (size_t)h->d.N * h->d.K * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        zd = h->d_z;
    }
    rb_launch_weight(h->d, zd, nullptr, 0, h->stream);
    CK(cudaGetLastError());
    return RBPF_OK;
}

extern "C" int rbpf_weight_guesses(rbpf_handle h, const double *guesses)
{
    if (h) h->d.use_dup = 0;
    if (!h || !h->have_scan || !guesses) { if (h) h->err = "weight_guesses: no scan set or no samples"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
    CK(cudaMemcpyAsync(h->d_z, guesses, (size_t)h->d.N * h->d.K * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rb_launch_weight(h->d, nullptr, h->d_z, 0, h->stream);
    CK(cudaGetLastError());
    return RBPF_OK;
}

// The reference-count pass of a sharded resample may be deferred (rbpf_resample_apply_local_deferred): it is launched
// here, behind its gate, by the first call that needs the pool's bookkeeping -- reference counts, free list, `mult`,
// or the old page tables as the target of the next gather.  Motion, matching and weighting only read tiles.
// Sub-tiles of migrated particles may still be arriving on the copy stream (rbpf_migrate_pull_async): everything that
// touches tiles of arbitrary particles waits for them here; rbpf_scan_match matches the local particles first.
static int flush_pulled(rbpf_ctx *h)
{
    if (!h->pulled_pending) return RBPF_OK;
    h->pulled_pending = false;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamWaitEvent(h->stream, h->ev_copied, 0));
    return RBPF_OK;
}

static int flush_refs(rbpf_ctx *h)
{
    { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
    if (!h->refs_pending) return RBPF_OK;
    h->refs_pending = false;
    CK(cudaSetDevice(h->cfg.device));
    if (h->refs_gate) CK(cudaStreamWaitEvent(h->stream, h->refs_gate, 0));
    RbCtx d = h->d;
    d.pt = h->refs_pt;
    rb_launch_resample_refs(d, h->stream);
    CK(cudaGetLastError());
    return RBPF_OK;
}

extern "C" int rbpf_integrate(rbpf_handle h, int32_t fallback_weights)
{
    if (h) h->d.use_dup = 0;
    if (!h || !h->have_scan) { if (h) h->err = "integrate: no scan set"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    rb_launch_raycast_prepare(h->d, h->stream);
    rb_launch_raycast_cast(h->d, h->stream);
    if (fallback_weights) rb_launch_weight(h->d, nullptr, nullptr, 1, h->stream);
    CK(cudaGetLastError());
    return RBPF_OK;
}

static void swap_buffers(rbpf_ctx *h)
{
    RbCtx &d = h->d;
    d.use_dup = 1;                 // set by the resample that just ran; cleared by anything but motion
    std::swap(d.pose, d.pose2);
    std::swap(d.cov, d.cov2);
    std::swap(d.pt, d.pt2);
    std::swap(d.exists, d.exists2);
    h->parity ^= 1;
}

static int resample_common(rbpf_ctx *h, const double *weights_all_dev, const double *u01, int32_t *ancestors_out,
                           int32_t *did_resample)
{
    const double *ud = nullptr;
    if (u01) {
        CK(cudaMemcpyAsync(h->d_u01, u01, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        ud = h->d_u01;
    }
    rb_launch_resample(h->d, weights_all_dev, ud, h->stream);
    CK(cudaGetLastError());
    if (ancestors_out || did_resample) {
        RbFlags f;
        CK(cudaMemcpyAsync(&f, h->d.flags, sizeof(f), cudaMemcpyDeviceToHost, h->stream));
        if (ancestors_out)
            CK(cudaMemcpyAsync(ancestors_out, h->d.ancestors, sizeof(int) * (size_t)h->d.n_global,
                               cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (did_resample) *did_resample = f.did_resample;
        if (f.resample_error) {
            h->err = "Incorrect number of resampled weights.";            // main.py:67
            return RBPF_ERR_RESAMPLE;
        }
    }
    return RBPF_OK;
}

extern "C" int rbpf_resample(rbpf_handle h, const double *u01, int32_t *ancestors_out, int32_t *did_resample)
{
    if (!h) return RBPF_ERR_ARG;
    if (h->d.world != 1) { h->err = "resample: sharded handle, use rbpf_resample_global"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    int rc = resample_common(h, h->d.weight, u01, ancestors_out, did_resample);
    rb_launch_resample_apply(h->d, h->stream);                  // identity gather when nothing triggered / on error
    CK(cudaGetLastError());
    swap_buffers(h);
    h->d.step_no++;
    return rc;
}

extern "C" int rbpf_step(rbpf_handle h, const double *ranges, const double *angles, int32_t n_beams)
{
    if (!h) return RBPF_ERR_ARG;
    if (h->d.world != 1) { h->err = "step: sharded handle, drive the stages from thesis_b200.dist"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    // errors of the step before the previous one (pool exhausted, world overflow, resample assertion):
    // their flags were copied to pinned memory behind that step; waiting for them keeps the host at
    // most two steps ahead of the device
    const int fs = (int)(h->d.step_no & 1ull);
    if (h->flag_pending[fs]) {
        CK(cudaEventSynchronize(h->flag_ev[fs]));
        h->flag_pending[fs] = 0;
        const int frc = flags_status(h, h->h_flags[fs]);
        if (frc) return frc;
    }
    const bool timed = h->t_max_steps > 0 && h->t_steps < h->t_max_steps;
    cudaEvent_t *ev = timed ? &h->tev[(size_t)h->t_steps * (RB_NSTAGES + 1)] : nullptr;
#define MARK(i) if (timed) cudaEventRecord(ev[i], h->stream)
    MARK(0);
    int rc = rbpf_set_scan(h, ranges, angles, n_beams);
    if (rc) return rc;
    MARK(1);
    rb_launch_match(h->d, 0, h->stream);
    h->last_adj = 0;
    MARK(2);
    h->d.use_dup = 0;
    rb_launch_weight(h->d, nullptr, nullptr, 0, h->stream);
    MARK(3);
    rb_launch_raycast_prepare(h->d, h->stream);
    MARK(4);
    rb_launch_raycast_cast(h->d, h->stream);
    MARK(5);
    rb_launch_weight(h->d, nullptr, nullptr, 1, h->stream);
    MARK(6);
    rb_launch_resample(h->d, h->d.weight, nullptr, h->stream);
    MARK(7);
    rb_launch_resample_apply(h->d, h->stream);
    MARK(8);
#undef MARK
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h->h_flags[fs], h->d.flags, sizeof(RbFlags), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->flag_ev[fs], h->stream));
    h->flag_pending[fs] = 1;
    if (timed) h->t_steps++;
    swap_buffers(h);
    h->d.step_no++;
    return RBPF_OK;
}

extern "C" int rbpf_timing_enable(rbpf_handle h, int32_t max_steps)
{
    if (!h || max_steps < 0) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
    h->tev.clear();
    h->tev.resize((size_t)max_steps * (RB_NSTAGES + 1));
    for (auto &e : h->tev) CK(cudaEventCreate(&e));
    h->t_max_steps = max_steps;
    h->t_steps = 0;
    return RBPF_OK;
}

extern "C" int rbpf_timing_read(rbpf_handle h, double *ms_out, int32_t *steps_out)
{
    if (!h || !ms_out || !steps_out) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < RB_NSTAGES; s++) ms_out[s] = 0.0;
    for (int i = 0; i < h->t_steps; i++)
        for (int s = 0; s < RB_NSTAGES; s++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->tev[(size_t)i * (RB_NSTAGES + 1) + s], h->tev[(size_t)i * (RB_NSTAGES + 1) + s + 1]));
            ms_out[s] += ms;
        }
    *steps_out = h->t_steps;
    h->t_steps = 0;
    return RBPF_OK;
}

// ---- state access -----------------------------------------------------------

static int copy_out(rbpf_ctx *h, void *dst, const void *src, size_t bytes)
{
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBPF_OK;
}

static int copy_in(rbpf_ctx *h, void *dst, const void *src, size_t bytes)
{
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBPF_OK;
}

extern "C" int rbpf_get_poses(rbpf_handle h, double *o) { return h && o ? copy_out(h, o, h->d.pose, sizeof(double) * 3 * h->d.N) : RBPF_ERR_ARG; }
extern "C" int rbpf_get_covs(rbpf_handle h, double *o) { return h && o ? copy_out(h, o, h->d.cov, sizeof(double) * 9 * h->d.N) : RBPF_ERR_ARG; }
extern "C" int rbpf_get_weights(rbpf_handle h, double *o) { return h && o ? copy_out(h, o, h->d.weight, sizeof(double) * h->d.N) : RBPF_ERR_ARG; }
extern "C" int rbpf_set_poses(rbpf_handle h, const double *i) { if (h) h->d.use_dup = 0; return h && i ? copy_in(h, h->d.pose, i, sizeof(double) * 3 * h->d.N) : RBPF_ERR_ARG; }
extern "C" int rbpf_set_covs(rbpf_handle h, const double *i) { if (h) h->d.use_dup = 0; return h && i ? copy_in(h, h->d.cov, i, sizeof(double) * 9 * h->d.N) : RBPF_ERR_ARG; }
extern "C" int rbpf_set_weights(rbpf_handle h, const double *i) { return h && i ? copy_in(h, h->d.weight, i, sizeof(double) * h->d.N) : RBPF_ERR_ARG; }

extern "C" int rbpf_get_match(rbpf_handle h, double *pose, double *cov, double *score, int32_t *valid, int32_t *best)
{
    if (!h) return RBPF_ERR_ARG;
    const size_t N = h->d.N;
    int rc = RBPF_OK;
    if (pose && !rc) rc = copy_out(h, pose, h->d.m_pose, sizeof(double) * 3 * N);
    if (cov && !rc) rc = copy_out(h, cov, h->d.m_cov, sizeof(double) * 9 * N);
    if (score && !rc) rc = copy_out(h, score, h->d.m_score, sizeof(double) * N);
    if (valid && !rc) rc = copy_out(h, valid, h->d.m_valid, sizeof(int) * N);
    if (best && !rc) rc = copy_out(h, best, h->d.m_best, sizeof(int) * 4 * N);
    return rc;
}

extern "C" int rbpf_get_resample_cumsum(rbpf_handle h, double *out)
{
    if (!h || !out) return RBPF_ERR_ARG;
    return copy_out(h, out, h->d.w_all, sizeof(double) * (size_t)h->d.n_global);
}

extern "C" int rbpf_get_match_refine(rbpf_handle h, int32_t *out)
{
    if (!h || !out) return RBPF_ERR_ARG;
    return copy_out(h, out, h->d.m_refine, sizeof(int) * 2 * (size_t)h->d.N);
}

extern "C" int rbpf_set_refine(rbpf_handle h, int32_t on)
{
    if (!h) return RBPF_ERR_ARG;
    h->d.refine = on ? 1 : 0;
    return RBPF_OK;
}

extern "C" int rbpf_set_match(rbpf_handle h, const double *pose, const double *cov, const int32_t *valid)
{
    if (!h || !pose || !cov || !valid) return RBPF_ERR_ARG;
    const size_t N = h->d.N;
    int rc = copy_in(h, h->d.m_pose, pose, sizeof(double) * 3 * N);
    if (!rc) rc = copy_in(h, h->d.m_cov, cov, sizeof(double) * 9 * N);
    if (!rc) rc = copy_in(h, h->d.m_valid, valid, sizeof(int) * N);
    return rc;
}

extern "C" int rbpf_get_match_slice(rbpf_handle h, int32_t particle, int32_t *out)
{
    if (!h || !out || particle < 0 || particle >= h->d.N || !h->have_scan) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
    CK(cudaMemsetAsync(h->d_slice, 0, sizeof(int) * RB_SLICE_W * RB_SLICE_W, h->stream));
    rb_launch_match_slice(h->d, particle, h->d_slice, h->last_adj, h->stream);
    CK(cudaGetLastError());
    return copy_out(h, out, h->d_slice, sizeof(int) * RB_SLICE_W * RB_SLICE_W);
}

extern "C" int rbpf_export_tile(rbpf_handle h, int32_t particle, int32_t cx, int32_t cy, double *out, int32_t *exists)
{
    if (!h || !out || particle < 0 || particle >= h->d.N || cx % 40 || cy % 40) return RBPF_ERR_ARG;
    { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
    const int tx = cx / 40, ty = cy / 40;
    memset(out, 0, sizeof(double) * RB_DIM * RB_DIM);
    if (exists) *exists = 0;
    if (tx < -h->d.txh || tx > h->d.txh || ty < -h->d.tyh || ty > h->d.tyh) return RBPF_OK;
    unsigned long long mask = 0;
    int rc = copy_out(h, &mask, h->d.exists + particle, sizeof(mask));
    if (rc) return rc;
    if (!((mask >> ((ty + h->d.tyh) * h->d.tiles_x + (tx + h->d.txh))) & 1ull)) return RBPF_OK;
    if (exists) *exists = 1;
    rb_launch_export_tile(h->d, particle, tx, ty, h->d_tile, h->stream);
    CK(cudaGetLastError());
    return copy_out(h, out, h->d_tile, sizeof(double) * RB_DIM * RB_DIM);
}

extern "C" int rbpf_list_tiles(rbpf_handle h, int32_t particle, int32_t *out_xy, int32_t max_tiles, int32_t *n)
{
    if (!h || !n || particle < 0 || particle >= h->d.N) return RBPF_ERR_ARG;
    unsigned long long mask = 0;
    int rc = copy_out(h, &mask, h->d.exists + particle, sizeof(mask));
    if (rc) return rc;
    int cnt = 0;
    for (int ty = -h->d.tyh; ty <= h->d.tyh; ty++)
        for (int tx = -h->d.txh; tx <= h->d.txh; tx++)
            if ((mask >> ((ty + h->d.tyh) * h->d.tiles_x + (tx + h->d.txh))) & 1ull) {
                if (out_xy && cnt < max_tiles) { out_xy[2 * cnt] = 40 * tx; out_xy[2 * cnt + 1] = 40 * ty; }
                cnt++;
            }
    *n = cnt;
    return RBPF_OK;
}

extern "C" int rbpf_stats(rbpf_handle h, rbpf_stats_t *out)
{
    if (!h || !out) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    RbStats s;
    int fc = 0;
    unsigned long long rs[3];
    rb_launch_refstats(h->d, h->d_refstats, h->stream);
    CK(cudaMemcpyAsync(&s, h->d.stats, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&fc, h->d.free_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(rs, h->d_refstats, sizeof(rs), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    out->pool_subtiles = h->d.pool_tiles;
    out->pool_in_use = h->d.pool_tiles - (uint32_t)(fc < 0 ? 0 : fc);
    out->cow_copies = s.cow_copies;
    out->fresh_allocs = s.fresh_allocs;
    out->cells_dropped = s.cells_dropped;
    out->resamples = s.resamples;
    out->match_failed = s.match_failed;
    out->shared_refs = rs[0];
    out->total_refs = rs[1];
    out->refcount_sum = rs[2];
    out->match_evals = s.match_evals;
    out->match_visits = s.match_visits;
    out->match_points = s.match_points;
    out->match_runs = s.match_runs;
    out->ndt_evals = s.ndt_evals;
    out->ndt_accepted = s.ndt_accepted;
    out->match_failed_zero = s.match_failed_zero;
    return RBPF_OK;
}

extern "C" int rbpf_match_phase_clocks(rbpf_handle h, uint64_t *out16)
{
    if (!h || !out16) return RBPF_ERR_ARG;
    RbStats s;
    int rc = copy_out(h, &s, h->d.stats, sizeof(s));
    if (rc) return rc;
    for (int i = 0; i < 16; i++) out16[i] = s.match_clk[i];
    return RBPF_OK;
}

// ---- multi-GPU resampling -------------------------------------------------------

extern "C" int rbpf_weights_device_ptr(rbpf_handle h, uint64_t *dev_ptr)
{
    if (!h || !dev_ptr) return RBPF_ERR_ARG;
    *dev_ptr = (uint64_t)(uintptr_t)h->d.weight;
    return RBPF_OK;
}

// Global systematic resample on the all-gathered weights; every rank computes
// the identical ancestor vector (main.py:46-67).  Does not move any particle.
extern "C" int rbpf_resample_global(rbpf_handle h, uint64_t weights_all_dev, int32_t n_global, const double *u01,
                                    int32_t *did_resample, int32_t *ancestors_out)
{
    if (!h || !weights_all_dev || n_global != h->d.n_global) { if (h) h->err = "resample_global: bad arguments"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    return resample_common(h, (const double *)(uintptr_t)weights_all_dev, u01, ancestors_out, did_resample);
}

static int upload_ints(rbpf_ctx *h, int *dst, const int32_t *src, int n)
{
    if (n > 0) CK(cudaMemcpyAsync(dst, src, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    return RBPF_OK;
}

// Sender, step 1: how many distinct sub-tiles go with these local particles.
extern "C" int rbpf_migrate_count(rbpf_handle h, const int32_t *src_slots, int32_t n, int32_t *n_subtiles, int64_t *bytes)
{
    if (!h || n < 0 || n > h->d.N || (n > 0 && !src_slots) || !n_subtiles) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    int rc = upload_ints(h, h->d_mg_slots, src_slots, n);
    if (rc) return rc;
    rb_launch_migrate_claim(h->d, h->d_mg_slots, n, h->d_mg_mark, h->d_mg_list, h->d_mg_count, h->stream);
    CK(cudaGetLastError());
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, h->d_mg_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *n_subtiles = cnt;
    h->mg_n = n;
    h->mg_tiles = cnt;
    if (bytes) *bytes = (int64_t)rb_migrate_bytes(n, cnt, h->d.nsub);
    return RBPF_OK;
}

// Sender, step 2: pack what rbpf_migrate_count just claimed into dev_buf.
extern "C" int rbpf_migrate_pack(rbpf_handle h, uint64_t dev_buf)
{
    if (!h || (!dev_buf && h->mg_n > 0)) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    rb_launch_migrate_pack(h->d, h->d_mg_slots, h->mg_n, h->mg_tiles, h->d_mg_mark, h->d_mg_list, h->d_mg_count,
                           (unsigned char *)(uintptr_t)dev_buf, h->stream);
    CK(cudaGetLastError());
    h->mg_n = h->mg_tiles = 0;
    return RBPF_OK;
}

// Receiver: adopt n_particles records / n_subtiles payloads from dev_buf; local
// destination slot dst_slots[i] becomes a copy of received record rec_index[i].
extern "C" int rbpf_migrate_unpack(rbpf_handle h, uint64_t dev_buf, int32_t n_particles, int32_t n_subtiles,
                                   const int32_t *dst_slots, const int32_t *rec_index, int32_t m)
{
    if (!h || n_particles < 0 || m < 0 || m > h->d.N || (m > 0 && (!dst_slots || !rec_index || !dev_buf))) return RBPF_ERR_ARG;
    if ((uint32_t)n_subtiles > h->d.pool_tiles) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    int rc = upload_ints(h, h->d_mg_slots, dst_slots, m);
    if (!rc) rc = upload_ints(h, h->d_mg_slots + h->d.N, rec_index, m);
    if (rc) return rc;
    rb_launch_migrate_unpack(h->d, (const unsigned char *)(uintptr_t)dev_buf, n_particles, n_subtiles, h->d_mg_slots,
                             h->d_mg_slots + h->d.N, m, h->d_mg_list, h->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));                       // the staging arrays are reused by the next peer
    return RBPF_OK;
}

// Local half of the resample (gather of local ancestors, reference counts); call
// after every rbpf_migrate_pack and before any rbpf_migrate_unpack.
extern "C" int rbpf_resample_apply_local(rbpf_handle h)
{
    if (!h) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    rb_launch_resample_apply(h->d, h->stream);
    CK(cudaGetLastError());
    return RBPF_OK;
}

// Same, with the reference-count pass deferred (include/rbpf_b200.h): the gather runs now, the counts -- and with them
// every free and every in-place write of a tile a peer may still be pulling -- wait for `gate_event`.
extern "C" int rbpf_resample_apply_local_deferred(rbpf_handle h, uint64_t gate_event)
{
    if (!h) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    rb_launch_resample_gather(h->d, h->stream);
    CK(cudaGetLastError());
    h->refs_pending = true;
    h->refs_gate = (cudaEvent_t)(uintptr_t)gate_event;
    h->refs_pt = h->d.pt;
    return RBPF_OK;
}

// Make the resampled buffers current (after the last rbpf_migrate_unpack).
extern "C" int rbpf_resample_commit(rbpf_handle h)
{
    if (!h) return RBPF_ERR_ARG;
    swap_buffers(h);
    h->d.step_no++;
    return RBPF_OK;
}

// ---- pull migration over peer memory ----------------------------------------------

extern "C" int rbpf_peer_export(rbpf_handle h, rbpf_peer_view *out)
{
    if (!h || !out) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < 9; i++) {
        cudaIpcMemHandle_t ih;
        CK(cudaIpcGetMemHandle(&ih, h->phys[i]));
        static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(out->ipc[i], &ih, 64);
        out->ptr[i] = (uint64_t)(uintptr_t)h->phys[i];
    }
    out->parity = h->parity;
    out->pid = (int64_t)getpid();
    out->device = h->cfg.device;
    out->n_particles = h->d.N;
    out->pool_subtiles = h->d.pool_tiles;
    out->nsub = h->d.nsub;
    return RBPF_OK;
}

extern "C" int rbpf_peer_attach(rbpf_handle h, int32_t peer_rank, const rbpf_peer_view *view)
{
    if (!h || !view || peer_rank < 0 || peer_rank >= h->d.world || peer_rank == h->d.rank) return RBPF_ERR_ARG;
    if (view->n_particles != h->d.N || view->pool_subtiles != h->d.pool_tiles || view->nsub != h->d.nsub ||
        view->parity != h->parity) {
        h->err = "peer_attach: the peer's particle count, pool size, world extent or buffer parity differs";
        return RBPF_ERR_ARG;
    }
    CK(cudaSetDevice(h->cfg.device));
    if ((int)h->peers.size() < h->d.world) h->peers.resize(h->d.world);
    rbpf_ctx::PeerMap &pm = h->peers[peer_rank];
    if (pm.attached) { h->err = "peer_attach: already attached"; return RBPF_ERR_ARG; }
    if (view->pid == (int64_t)getpid()) {                       // same process (tests): the pointers are valid as they are
        if (view->device != h->cfg.device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, h->cfg.device, view->device));
            if (!can) { h->err = "peer_attach: no peer access between the two devices"; return RBPF_ERR_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(view->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
        }
        for (int i = 0; i < 9; i++) pm.base[i] = (void *)(uintptr_t)view->ptr[i];
        pm.ipc = false;
    } else {
        for (int i = 0; i < 9; i++) {
            cudaIpcMemHandle_t ih;
            memcpy(&ih, view->ipc[i], 64);
            cudaError_t e = cudaIpcOpenMemHandle(&pm.base[i], ih, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (int k = 0; k < i; k++) cudaIpcCloseMemHandle(pm.base[k]);
                h->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
                cudaGetLastError();
                return RBPF_ERR_CUDA;
            }
        }
        pm.ipc = true;
    }
    pm.attached = true;
    return RBPF_OK;
}

// Receiver: every local slot whose ancestor (rbpf_resample_global) lives on another rank
// becomes a copy of that particle, read through the peer mappings (the job runs in
// lockstep: the peers' current buffers have this handle's parity).  The plan is the
// ancestor vector on the device: nothing is copied to or from the host, nothing waits.
extern "C" int rbpf_migrate_pull_async(rbpf_handle h, uint64_t copy_stream)
{
    if (!h) return RBPF_ERR_ARG;
    if (h->d.world > RB_MAX_WORLD) { h->err = "migrate_pull: world larger than RB_MAX_WORLD"; return RBPF_ERR_ARG; }
    RbPeers peers;
    memset(&peers, 0, sizeof(peers));
    const int q = h->parity;
    for (int r = 0; r < h->d.world; r++) {
        if (r == h->d.rank) continue;
        if (r >= (int)h->peers.size() || !h->peers[r].attached) { h->err = "migrate_pull: a peer is not attached"; return RBPF_ERR_ARG; }
        const rbpf_ctx::PeerMap &pm = h->peers[r];
        peers.p[r].pool = (const int8_t *)pm.base[0];
        peers.p[r].pt = (const uint32_t *)pm.base[1 + q];
        peers.p[r].pose = (const double *)pm.base[3 + q];
        peers.p[r].cov = (const double *)pm.base[5 + q];
        peers.p[r].exists = (const unsigned long long *)pm.base[7 + q];
    }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    if (!h->d_pull_mark) {                                      // first use: one claim table per source rank
        const size_t n = (size_t)h->d.world * h->d.pool_tiles;
        CK(cudaMalloc((void **)&h->d_pull_mark, n * sizeof(uint32_t)));
        h->allocs.push_back(h->d_pull_mark);
        CK(cudaMalloc((void **)&h->d_pull_rank, h->d.pool_tiles));
        h->allocs.push_back(h->d_pull_rank);
        CK(cudaMalloc((void **)&h->d_pull_local, (size_t)h->d.pool_tiles * sizeof(uint32_t)));
        h->allocs.push_back(h->d_pull_local);
        CK(cudaEventCreateWithFlags(&h->ev_alloc, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
        CK(cudaMemsetAsync(h->d_pull_mark, 0xFF, n * sizeof(uint32_t), h->stream));
    }
    cudaStream_t cs = copy_stream ? (cudaStream_t)(uintptr_t)copy_stream : h->stream;
    rb_launch_migrate_pull(h->d, peers, h->d_pull_mark, h->d_mg_list, h->d_pull_rank, h->d_pull_local, h->d_mg_count, h->stream, cs,
                           h->ev_alloc);
    CK(cudaGetLastError());
    if (cs != h->stream) {                                      // the payloads arrive behind this event
        CK(cudaEventRecord(h->ev_copied, cs));
        h->pulled_pending = true;
    }
    return RBPF_OK;
}

extern "C" int rbpf_migrate_pull(rbpf_handle h) { return rbpf_migrate_pull_async(h, 0); }


extern "C" int64_t rbpf_migrate_bytes(rbpf_handle h, int32_t n_particles, int32_t n_subtiles)
{
    return h ? (int64_t)rb_migrate_bytes(n_particles, n_subtiles, h->d.nsub) : 0;
}

// HybridMap.get_occupied_points (hybridmap.py:303-313) on the device: threshold and
// compact the particle's sub-tiles.  out_xy may be NULL (count only); at most
// max_points pairs are written, *n receives the total number of occupied cells.
extern "C" int rbpf_occupied_points(rbpf_handle h, int32_t particle, double *out_xy, int64_t max_points, int64_t *n)
{
    if (!h || !n || particle < 0 || particle >= h->d.N || max_points < 0) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_pulled(h); if (rc_) return rc_; }
    double *dev = nullptr;
    if (out_xy && max_points > 0) CK(cudaMalloc((void **)&dev, sizeof(double) * 2 * (size_t)max_points));
    rb_launch_occupied_points(h->d, particle, dev, (unsigned long long)max_points, h->d_refstats + 3, h->stream);
    unsigned long long cnt = 0;
    cudaError_t e = cudaMemcpyAsync(&cnt, h->d_refstats + 3, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess && dev) {
        const unsigned long long m = cnt < (unsigned long long)max_points ? cnt : (unsigned long long)max_points;
        e = cudaMemcpy(out_xy, dev, sizeof(double) * 2 * (size_t)m, cudaMemcpyDeviceToHost);
    }
    if (dev) cudaFree(dev);
    if (e != cudaSuccess) { h->err = std::string("occupied_points: ") + cudaGetErrorString(e); return RBPF_ERR_CUDA; }
    *n = (int64_t)cnt;
    return RBPF_OK;
}

// ---- device-side snapshot -----------------------------------------------------------
// The whole mutable state copied device-to-device into shadow buffers (allocated on first use: it
// doubles the handle's memory), and back.  bench.py uses it to time the device-resident loop and the
// host-buffer loop on the SAME scans; it is also the cheap way to branch a filter.
extern "C" int rbpf_snapshot(rbpf_handle h)
{
    if (!h) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    RbCtx &d = h->d;
    const size_t N = d.N;
    if (h->snap.empty()) {
        // double-buffered arrays by their physical allocation (h->phys), so that a restore is independent of the parity
        const std::pair<void *, size_t> bufs[] = {
            {h->phys[0], (size_t)d.pool_tiles * RB_SUB_BYTES}, {d.refcnt, sizeof(uint32_t) * d.pool_tiles},
            {d.free_list, sizeof(uint32_t) * d.pool_tiles}, {d.free_count, sizeof(int) * 4},
            {h->phys[1], sizeof(uint32_t) * N * d.nsub}, {h->phys[2], sizeof(uint32_t) * N * d.nsub},
            {h->phys[3], sizeof(double) * 3 * N}, {h->phys[4], sizeof(double) * 3 * N},
            {h->phys[5], sizeof(double) * 9 * N}, {h->phys[6], sizeof(double) * 9 * N},
            {h->phys[7], sizeof(unsigned long long) * N}, {h->phys[8], sizeof(unsigned long long) * N},
            {d.weight, sizeof(double) * N}, {d.dup_of, sizeof(int) * N}, {d.ancestors, sizeof(int) * (size_t)d.n_global},
            {d.stats, sizeof(RbStats)}, {d.flags, sizeof(RbFlags)},
        };
        for (const auto &b : bufs) {
            void *sh = nullptr;
            cudaError_t e = cudaMalloc(&sh, b.second ? b.second : 16);
            if (e != cudaSuccess) {
                for (auto &s2 : h->snap) cudaFree(s2.shadow);
                h->snap.clear();
                h->err = std::string("snapshot: cudaMalloc: ") + cudaGetErrorString(e);
                cudaGetLastError();
                return RBPF_ERR_CUDA;
            }
            h->snap.push_back({b.first, sh, b.second});
        }
    }
    for (const auto &s2 : h->snap) CK(cudaMemcpyAsync(s2.shadow, s2.live, s2.bytes, cudaMemcpyDeviceToDevice, h->stream));
    h->snap_parity = h->parity;
    h->snap_use_dup = d.use_dup;
    h->snap_step_no = d.step_no;
    h->snap_valid = 1;
    return RBPF_OK;
}

extern "C" int rbpf_restore(rbpf_handle h)
{
    if (!h) return RBPF_ERR_ARG;
    if (!h->snap_valid) { h->err = "restore: no snapshot"; return RBPF_ERR_ARG; }
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    for (const auto &s2 : h->snap) CK(cudaMemcpyAsync(s2.live, s2.shadow, s2.bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (h->parity != h->snap_parity) swap_buffers(h);
    h->d.use_dup = h->snap_use_dup;
    h->d.step_no = h->snap_step_no;
    h->flag_pending[0] = h->flag_pending[1] = 0;
    return RBPF_OK;
}

// ---- checkpoint / resume of the whole particle set --------------------------------
// The reference pickles particle 0 only every 50 frames (main.py:183-210); here the
// complete set is written: state, page tables, reference counts and every sub-tile
// in use (sparse, by pool index), through a pinned bounce buffer.
struct CkptHeader {
    char magic[8];
    int32_t N, nsub, tiles_x, tiles_y, rank, world;
    uint32_t pool_tiles, in_use;
    uint64_t step_no;
};

static const size_t CKPT_CH = 32u << 20;

static int ckpt_io(rbpf_ctx *h, FILE *f, void *dev, size_t bytes, bool write)
{
    const size_t CH = CKPT_CH;
    if (!h->ckpt_host) CK(cudaMallocHost(&h->ckpt_host, CH));
    void *host = h->ckpt_host;
    int rc = RBPF_OK;
    for (size_t off = 0; off < bytes && rc == RBPF_OK; off += CH) {
        const size_t n = bytes - off < CH ? bytes - off : CH;
        if (write) {
            if (cudaMemcpy(host, (char *)dev + off, n, cudaMemcpyDeviceToHost) != cudaSuccess || fwrite(host, 1, n, f) != n) rc = RBPF_ERR_CUDA;
        } else {
            if (fread(host, 1, n, f) != n || cudaMemcpy((char *)dev + off, host, n, cudaMemcpyHostToDevice) != cudaSuccess) rc = RBPF_ERR_CUDA;
        }
    }
    if (rc) h->err = write ? "checkpoint: write failed" : "checkpoint: read failed (truncated file?)";
    return rc;
}

extern "C" int rbpf_checkpoint_write(rbpf_handle h, const char *path)
{
    if (!h || !path) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    CK(cudaStreamSynchronize(h->stream));
    const RbCtx &d = h->d;
    std::vector<uint32_t> rc(d.pool_tiles);
    CK(cudaMemcpy(rc.data(), d.refcnt, sizeof(uint32_t) * d.pool_tiles, cudaMemcpyDeviceToHost));
    CkptHeader hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, "RBPFCK01", 8);
    hd.N = d.N; hd.nsub = d.nsub; hd.tiles_x = d.tiles_x; hd.tiles_y = d.tiles_y; hd.rank = d.rank; hd.world = d.world;
    hd.pool_tiles = d.pool_tiles; hd.step_no = d.step_no;
    for (uint32_t t = 0; t < d.pool_tiles; t++) hd.in_use += rc[t] > 0;
    FILE *f = fopen(path, "wb");
    if (!f) { h->err = std::string("checkpoint: cannot open ") + path; return RBPF_ERR_ARG; }
    int r = fwrite(&hd, sizeof(hd), 1, f) == 1 ? RBPF_OK : RBPF_ERR_CUDA;
    const size_t N = d.N;
    if (!r) r = ckpt_io(h, f, d.pose, sizeof(double) * 3 * N, true);
    if (!r) r = ckpt_io(h, f, d.cov, sizeof(double) * 9 * N, true);
    if (!r) r = ckpt_io(h, f, d.weight, sizeof(double) * N, true);
    if (!r) r = ckpt_io(h, f, d.exists, sizeof(unsigned long long) * N, true);
    if (!r) r = ckpt_io(h, f, d.pt, sizeof(uint32_t) * N * d.nsub, true);
    if (!r) r = fwrite(rc.data(), sizeof(uint32_t), d.pool_tiles, f) == d.pool_tiles ? RBPF_OK : RBPF_ERR_CUDA;
    for (uint32_t t = 0; t < d.pool_tiles && !r; t++)
        if (rc[t] > 0) r = ckpt_io(h, f, d.pool + (size_t)t * RB_SUB_BYTES, RB_SUB_BYTES, true);
    fclose(f);
    return r;
}

extern "C" int rbpf_checkpoint_read(rbpf_handle h, const char *path)
{
    if (!h || !path) return RBPF_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    { const int rc_ = flush_refs(h); if (rc_) return rc_; }
    CK(cudaStreamSynchronize(h->stream));
    RbCtx &d = h->d;
    FILE *f = fopen(path, "rb");
    if (!f) { h->err = std::string("checkpoint: cannot open ") + path; return RBPF_ERR_ARG; }
    CkptHeader hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, "RBPFCK01", 8) != 0 || hd.N != d.N || hd.nsub != d.nsub ||
        hd.tiles_x != d.tiles_x || hd.tiles_y != d.tiles_y || hd.pool_tiles != d.pool_tiles || hd.rank != d.rank ||
        hd.world != d.world) {
        fclose(f);
        h->err = "checkpoint: header does not match this handle's configuration (particles, world extent, pool, rank / world)";
        return RBPF_ERR_ARG;
    }
    const size_t N = d.N;
    {   // nothing of the handle is touched unless the file holds everything the header promises
        const size_t want = sizeof(hd) + sizeof(double) * 13 * N + sizeof(unsigned long long) * N + sizeof(uint32_t) * N * d.nsub +
                            sizeof(uint32_t) * d.pool_tiles + (size_t)hd.in_use * RB_SUB_BYTES;
        if (hd.in_use > d.pool_tiles || fseek(f, 0, SEEK_END) != 0 || (size_t)ftell(f) != want ||
            fseek(f, (long)sizeof(hd), SEEK_SET) != 0) {
            fclose(f);
            h->err = "checkpoint: file size does not match its header (truncated?)";
            return RBPF_ERR_ARG;
        }
    }
    std::vector<uint32_t> rc(d.pool_tiles);
    int r = ckpt_io(h, f, d.pose, sizeof(double) * 3 * N, false);
    if (!r) r = ckpt_io(h, f, d.cov, sizeof(double) * 9 * N, false);
    if (!r) r = ckpt_io(h, f, d.weight, sizeof(double) * N, false);
    if (!r) r = ckpt_io(h, f, d.exists, sizeof(unsigned long long) * N, false);
    if (!r) r = ckpt_io(h, f, d.pt, sizeof(uint32_t) * N * d.nsub, false);
    if (!r) r = fread(rc.data(), sizeof(uint32_t), d.pool_tiles, f) == d.pool_tiles ? RBPF_OK : RBPF_ERR_CUDA;
    std::vector<uint32_t> freel;
    for (uint32_t t = 0; t < d.pool_tiles && !r; t++) {
        if (rc[t] > 0) r = ckpt_io(h, f, d.pool + (size_t)t * RB_SUB_BYTES, RB_SUB_BYTES, false);
        else freel.push_back(t);
    }
    fclose(f);
    if (r) return r;
    // free list: pop order ascending like a fresh handle
    std::vector<uint32_t> fl(d.pool_tiles, 0u);
    for (size_t i = 0; i < freel.size(); i++) fl[i] = freel[freel.size() - 1 - i];
    const int fc = (int)freel.size();
    CK(cudaMemcpy(d.refcnt, rc.data(), sizeof(uint32_t) * d.pool_tiles, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.free_list, fl.data(), sizeof(uint32_t) * d.pool_tiles, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.free_count, &fc, sizeof(int), cudaMemcpyHostToDevice));
    d.step_no = hd.step_no;
    d.use_dup = 0;
    h->flag_pending[0] = h->flag_pending[1] = 0;
    return RBPF_OK;
}
