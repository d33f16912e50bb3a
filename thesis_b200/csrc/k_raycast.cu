// k_raycast.cu -- stage 4: log-odds ray-cast map integration into the
// copy-on-write tile pool.
//
// Reference: HybridMap.update hybridmap.py:95-145, get_affected_points :274-301,
// GridMap.set_occupied/empty/nearby_pos gridmap.py:86-117.
//
// Two launches, both one warp per particle:
//   raycast_prepare  marks (a superset of) the sub-tiles the sweep will write,
//                    then makes each of them private: unallocated -> fresh zero
//                    sub-tile, shared (refcount > 1) -> copy unless this particle
//                    turns out to be the last holder.  No cell is modified here,
//                    so sharers can copy while the future owner waits.
//   raycast_cast     replays the reference's update order exactly: beams in
//                    order, lanes across 32 consecutive cells of one ray
//                    (closed-form Bresenham for the first cell of a lane, then
//                    +32 cells incrementally), saturating int8 read-modify-write.
//                    The reference's clamps make the result order-dependent
//                    (SURVEY 3.4-4), hence no atomics and no beam parallelism
//                    inside a particle.  Per-beam constants are computed
//                    lane-parallel (one beam per lane) and broadcast; chunks whose
//                    cells are all plain "empty" updates take a fast path; a cell
//                    whose clamped value does not change is not stored back.
#include <stdlib.h>

#include "common.cuh"

#ifndef RC_WARPS
#define RC_WARPS 4
#endif
#ifndef RC_INFLIGHT
#define RC_INFLIGHT 4
#endif
// RC_INFLIGHT: chunks of 32 ray cells whose loads are issued together (fast path)
#define RC_INFLIGHT_G 2   // same for the general path, which only sees the tail of a ray

// One beam's ray on the 0.05 m lattice, everything the per-cell closed form needs.
struct Ray {
    int ex, ey;      // end cell
    int len;         // number of cells (0 for the reference's empty-list quirk, hybridmap.py:278-281)
    int occ;         // end cell is an obstacle (range <= 15 m)
};

struct RayStep {     // derived per-beam constants (uniform across the warp)
    int steep;       // major axis is y
    int smaj, smin;  // signed unit steps along the major / minor axis
    unsigned d2, D, D2;   // 2*minor extent, major extent, 2*major extent
    float inv_D2;
};

__device__ __forceinline__ Ray ray_of_beam(const RbCtx &c, int j, double x, double y, double cs_, double sn_, int sx,
                                           int sy)
{
    Ray r;
    double gx, gy;
    rb_xform(cs_, sn_, x, y, c.px[j], c.py[j], gx, gy);
    r.ex = rb_trunc(gx / RB_CS);                                   // hybridmap.py:106
    r.ey = rb_trunc(gy / RB_CS);
    r.occ = 1;
    double d = c.dist[j];
    if (d > RB_CLIP_R) {                                           // hybridmap.py:107-113
        double scale = 15.0 / d;
        int nex = rb_trunc((double)sx + scale * (double)(r.ex - sx));
        int ney = rb_trunc((double)sy + scale * (double)(r.ey - sy));
        r.ex = nex;
        r.ey = ney;
        r.occ = 0;
    }
    int dx = r.ex - sx, dy = r.ey - sy;
    int adx = abs(dx), ady = abs(dy);
    if (adx == 0) r.len = dy >= 0 ? dy + 1 : 0;                    // hybridmap.py:278-279
    else if (ady == 0) r.len = dx >= 0 ? dx + 1 : 0;               // hybridmap.py:280-281
    else r.len = max(adx, ady) + 1;
    return r;
}

__device__ __forceinline__ RayStep ray_step(int sx, int sy, const Ray &r)
{
    RayStep s;
    const int dx = r.ex - sx, dy = r.ey - sy;
    const int adx = abs(dx), ady = abs(dy);
    // degenerate rays only ever step in the positive direction (range(y0, y1+1))
    s.steep = ady > adx || adx == 0;
    const int amaj = s.steep ? ady : adx, amin = s.steep ? adx : ady;
    const int dmaj = s.steep ? dy : dx, dmin = s.steep ? dx : dy;
    s.smaj = (adx == 0 || ady == 0) ? 1 : (dmaj > 0 ? 1 : -1);
    s.smin = dmin > 0 ? 1 : (dmin < 0 ? -1 : 0);
    s.D = (unsigned)max(amaj, 1);
    s.D2 = 2u * s.D;
    s.d2 = 2u * (unsigned)amin;
    s.inv_D2 = 1.0f / (float)s.D2;
    return s;
}

// n-th cell of the reference's integer Bresenham in closed form:
// minor(n) = floor((2 n d + D) / (2 D)), D = major extent, d = minor extent.
// The quotient comes from a float reciprocal (operands < 2^19) fixed up exactly.
__device__ __forceinline__ void ray_cell(int sx, int sy, const RayStep &s, int n, int &kx, int &ky)
{
    const unsigned num = (unsigned)n * s.d2 + s.D;
    unsigned m = __float2uint_rz(__uint2float_rz(num) * s.inv_D2);
    int rem = (int)(num - m * s.D2);
    if (rem < 0) { m--; rem += (int)s.D2; }
    if (rem >= (int)s.D2) m++;
    const int maj = s.smaj * n, mn = s.smin * (int)m;
    kx = sx + (s.steep ? mn : maj);
    ky = sy + (s.steep ? maj : mn);
}

__device__ __forceinline__ bool particle_frame(const RbCtx &c, int p, double &x, double &y, double &cs_, double &sn_,
                                               int &sx, int &sy)
{
    const double *pose = c.pose + 3 * (size_t)p;
    x = pose[0];
    y = pose[1];
    double th = pose[2];
    rb_sincos(th, &sn_, &cs_);
    int tx, ty, ix, iy;
    rb_read_axis(x, tx, ix);
    rb_read_axis(y, ty, iy);
    if (!rb_tile_exists(c, c.exists[p], tx, ty)) return false;    // hybridmap.py:98-100
    sx = rb_trunc(x / RB_CS);                                      // hybridmap.py:102
    sy = rb_trunc(y / RB_CS);
    return true;
}

// ------------------------------------------------------------- prepare --

#ifndef RC_COPY_UNROLL
#define RC_COPY_UNROLL 5                // 16-byte loads in flight per lane of a copy-on-write sub-tile copy
#endif
constexpr int kCopyUnroll = RC_COPY_UNROLL;
__global__ void __launch_bounds__(RC_WARPS * 32) raycast_prepare_kernel(RbCtx c)
{
    __shared__ uint32_t mask_s[RC_WARPS][52];                      // 64 tiles * 25 sub-tiles = 1600 bits
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * RC_WARPS + warp;
    if (blockIdx.x == 0 && threadIdx.x == 0) *c.cast_work = 0;      // work counter of raycast_cast2_kernel, which runs next
    if (p >= c.N) return;
    uint32_t *mask = mask_s[warp];
    const int nwords = (c.nsub + 31) >> 5;
    for (int w = lane; w < nwords; w += 32) mask[w] = 0;
    __syncwarp();

    double x, y, cs_, sn_;
    int sx, sy;
    if (!particle_frame(c, p, x, y, cs_, sn_, sx, sy)) return;

    // superset of touched sub-tiles: for every 32-cell segment of every ray mark
    // the sub-tiles of its two end cells and the two mixed corners (a segment
    // spans at most 2x2 sub-tiles because 33 < 160 and the LUT is monotone).
    for (int j = lane; j < c.B; j += 32) {
        const Ray r = ray_of_beam(c, j, x, y, cs_, sn_, sx, sy);
        const RayStep st = ray_step(sx, sy, r);
        // sub-tile coordinates of the cells at n = 0, 32, 64, ... and of the last cell: consecutive
        // segments share an end, and most of them stay inside the sub-tile that is already marked
        int pax = 0, pay = 0, last = -1;
        for (int n0 = 0; n0 < r.len; n0 += 32) {
            const int n1 = min(n0 + 32, r.len - 1);
            int ax, ay, bx, by;
            if (n0 == 0) {
                ray_cell(sx, sy, st, 0, ax, ay);
                const uint32_t qx = rb_write_lut(c.lutx, ax, c.txh), qy = rb_write_lut(c.luty, ay, c.tyh);
                // when one end is outside the world, the cells of the segment that are inside still lie
                // between the inside end and the world border: use the border sub-tile of that axis
                pax = qx == RB_NONE ? (ax < 0 ? 0 : c.subs_x - 1) : (int)RB_LUT_SUB(qx);
                pay = qy == RB_NONE ? (ay < 0 ? 0 : c.subs_y - 1) : (int)RB_LUT_SUB(qy);
            }
            ray_cell(sx, sy, st, n1, bx, by);
            const uint32_t pbx = rb_write_lut(c.lutx, bx, c.txh), pby = rb_write_lut(c.luty, by, c.tyh);
            const int sbx = pbx == RB_NONE ? (bx < 0 ? 0 : c.subs_x - 1) : (int)RB_LUT_SUB(pbx);
            const int sby = pby == RB_NONE ? (by < 0 ? 0 : c.subs_y - 1) : (int)RB_LUT_SUB(pby);
            const int s0 = pay * c.subs_x + pax, s1 = sby * c.subs_x + sbx;
            if (s0 != s1 || s0 != last) {
                const int s2 = pay * c.subs_x + sbx, s3 = sby * c.subs_x + pax;
                atomicOr(&mask[s0 >> 5], 1u << (s0 & 31));
                if (s1 != s0) {
                    atomicOr(&mask[s1 >> 5], 1u << (s1 & 31));
                    if (s2 != s0 && s2 != s1) atomicOr(&mask[s2 >> 5], 1u << (s2 & 31));
                    if (s3 != s0 && s3 != s1) atomicOr(&mask[s3 >> 5], 1u << (s3 & 31));
                }
                last = s1;
            }
            pax = sbx; pay = sby;
        }
    }
    __syncwarp();

    uint32_t *pt = c.pt + (size_t)p * c.nsub;
    for (int w = 0; w < nwords; w++) {
        uint32_t bits = mask[w];
        while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            int sub = 32 * w + b;
            uint32_t told = pt[sub];
            // lane 0 decides: 0 = keep in place, 1 = fresh zero tile, 2 = copy
            int action = 0;
            uint32_t tnew = RB_NONE;
            if (lane == 0) {
                if (told == RB_NONE) action = 1;
                else if (atomicAdd(&c.refcnt[told], 0u) > 1u) {
                    uint32_t old = atomicSub(&c.refcnt[told], 1u);
                    if (old == 1u) atomicAdd(&c.refcnt[told], 1u);    // last holder: keep it (an add, not a store: a sharer that
                                                                      // ran out of pool may be putting its own reference back)
                    else action = 2;
                }
                if (action) {
                    int idx = atomicSub(c.free_count, 1) - 1;
                    if (idx < 0) {
                        atomicAdd(c.free_count, 1);                        // keep the stack pointer valid for the pushes of resample_refs
                        atomicExch(&c.flags->pool_exhausted, 1);
                        if (action == 2) atomicAdd(&c.refcnt[told], 1u);   // undo: stay a sharer
                        action = 0;
                    } else {
                        tnew = c.free_list[idx];
                        c.refcnt[tnew] = 1u;
                        atomicAdd(action == 1 ? &c.stats->fresh_allocs : &c.stats->cow_copies, 1ull);
                    }
                }
            }
            action = __shfl_sync(0xffffffffu, action, 0);
            tnew = __shfl_sync(0xffffffffu, tnew, 0);
            if (action == 0) continue;
            uint4 *dst = reinterpret_cast<uint4 *>(c.pool + (size_t)tnew * RB_SUB_BYTES);
            if (action == 1) {
                const uint4 z = make_uint4(0, 0, 0, 0);
                for (int q = lane; q < RB_SUB_BYTES / 16; q += 32) dst[q] = z;
            } else {
                const uint4 *src = reinterpret_cast<const uint4 *>(c.pool + (size_t)told * RB_SUB_BYTES);
#pragma unroll kCopyUnroll
                for (int q = lane; q < RB_SUB_BYTES / 16; q += 32) dst[q] = src[q];
            }
            if (lane == 0) pt[sub] = tnew;
        }
    }
}

// ---------------------------------------------------------------- cast --

__device__ __forceinline__ int apply_ops(int t, int ops)
{
    // bit0 = empty (-0.3, floor -3.0), bit1 = occupied (+0.8, cap 3.0), bit2 = nearby (+0.2, cap 3.0)
    if (ops & 1) t = max(t - RB_T_EMP, -RB_T_MAX);
    if (ops & 2) t = min(t + RB_T_OCC, RB_T_MAX);
    if (ops & 4) t = min(t + RB_T_NEAR, RB_T_MAX);
    return t;
}

// Per-beam constants computed lane-parallel (one beam per lane) and broadcast.
struct BeamPack {
    int w0;          // len (bits 0-11) | occ << 12 | steep << 13 | (smaj + 1) << 14 | (smin + 1) << 16
    int w1;          // D (bits 0-12) | step_q << 13        (D = major extent >= 1)
    int w2;          // d2 (bits 0-13) | step_e << 14       (d2 = 2 * minor extent)
    int end_tile;    // reference tile of the end cell (ty * tiles_x + tx) or -1
    unsigned magic;  // ceil(2^32 / D2): floor(num / D2) == umulhi(num, magic) for num * D2 < 2^32
};

// Same lookup from a table staged in shared memory (plain load).
__device__ __forceinline__ uint32_t rb_write_lut_s(const uint32_t *lut, int k, int half_tiles)
{
    const unsigned q = (unsigned)(k + 800 * half_tiles + 400);
    if (q >= (unsigned)(800 * (2 * half_tiles + 1))) return RB_NONE;
    return lut[q];
}

// LUT_SMEM: the two per-axis write LUTs (800 entries per reference tile and
// axis) are staged in shared memory, which takes two of the three global loads
// per cell off the L1 tag pipeline; for worlds too large for that the kernel
// reads them through the read-only path instead.
template <bool LUT_SMEM>
// no minimum-blocks bound: 56 registers, 9 CTAs per SM; forcing 10 or more spills and measured slower (3.28 -> 3.5 ms
// at 16,384 particles), and an explicit bound of 1 lets ptxas take more registers (4.45 ms)
__global__ void __launch_bounds__(RC_WARPS * 32) raycast_cast_kernel(RbCtx c)
{
    extern __shared__ uint32_t lut_s[];
    const unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * RC_WARPS + warp;
    if (LUT_SMEM) {
        const int nx_ = 800 * c.tiles_x, ny_ = 800 * c.tiles_y;
        for (int i = threadIdx.x; i < nx_; i += RC_WARPS * 32) lut_s[i] = c.lutx[i];
        for (int i = threadIdx.x; i < ny_; i += RC_WARPS * 32) lut_s[nx_ + i] = c.luty[i];
        __syncthreads();
    }
    if (p >= c.N) return;
    if (c.flags->pool_exhausted) return;                           // prepare could not privatise: skip the scan
    double x, y, cs_, sn_;
    int sx, sy;
    if (!particle_frame(c, p, x, y, cs_, sn_, sx, sy)) return;
    const uint32_t *pt = c.pt + (size_t)p * c.nsub;
    const unsigned long long ex_mask = c.exists[p];
    unsigned long long ex_new = 0ull;
    unsigned dropped = 0;
    int cached_sub = -1;
    int8_t *cached_base = nullptr;
    const uint32_t *lutx = LUT_SMEM ? lut_s : c.lutx, *luty = LUT_SMEM ? lut_s + 800 * c.tiles_x : c.luty;
#define RC_LUT(l, k, h) (LUT_SMEM ? rb_write_lut_s(l, k, h) : rb_write_lut(l, k, h))
    const int txh = c.txh, tyh = c.tyh, subs_x = c.subs_x, tiles_x = c.tiles_x;
    // every ray starts in the robot's cell: its sub-tile is the page-table hit of each beam's first cells
    int sub0 = -1;
    int8_t *base0 = nullptr;
    {
        const uint32_t p0x = RC_LUT(lutx, sx, txh), p0y = RC_LUT(luty, sy, tyh);
        if (p0x != RB_NONE && p0y != RB_NONE) {
            sub0 = (int)RB_LUT_SUB(p0y) * subs_x + (int)RB_LUT_SUB(p0x);
            const uint32_t t0 = pt[sub0];
            if (t0 == RB_NONE) sub0 = -1; else base0 = c.pool + (size_t)t0 * RB_SUB_BYTES;
        }
    }

    for (int j0 = 0; j0 < c.B; j0 += 32) {
        BeamPack mine;
        mine.w0 = 0; mine.w1 = 1; mine.w2 = 0; mine.end_tile = -1; mine.magic = 0u;
        if (j0 + lane < c.B) {
            const Ray r = ray_of_beam(c, j0 + lane, x, y, cs_, sn_, sx, sy);
            const RayStep st = ray_step(sx, sy, r);
            const uint32_t pex = RC_LUT(lutx, r.ex, txh), pey = RC_LUT(luty, r.ey, tyh);
            mine.end_tile = (pex == RB_NONE || pey == RB_NONE) ? -1 : (int)(RB_LUT_TILE(pey) * tiles_x + RB_LUT_TILE(pex));
            int len = r.len;
            if (len > 4095 || st.D > 4095u) len = 0;               // cannot happen: rays are clipped at 15 m = 300 cells
            // lanes advance 32 cells per chunk: minor += step_q (+1 on remainder overflow), e += step_e
            const unsigned step_q = (32u * st.d2) / st.D2, step_e = 32u * st.d2 - step_q * st.D2;
            // start and end cell inside the world => every cell of the ray is (bounding box)
            const int inside = pex != RB_NONE && pey != RB_NONE && RC_LUT(lutx, sx, txh) != RB_NONE && RC_LUT(luty, sy, tyh) != RB_NONE;
            mine.w0 = len | (r.occ << 12) | (st.steep << 13) | ((st.smaj + 1) << 14) | ((st.smin + 1) << 16) | (inside << 18);
            mine.magic = (unsigned)((0x100000000ull + st.D2 - 1) / st.D2);
            mine.w1 = (int)(st.D | (step_q << 13));
            mine.w2 = (int)(st.d2 | (step_e << 14));
        }
        const int nb = min(32, c.B - j0);
        for (int b = 0; b < nb; b++) {
            const int w0 = __shfl_sync(FULL, mine.w0, b);
            const int len = w0 & 0xfff;
            if (len == 0) continue;                                 // hybridmap.py:278-281 empty list
            const int w1 = __shfl_sync(FULL, mine.w1, b), w2 = __shfl_sync(FULL, mine.w2, b);
            const int end_tile = __shfl_sync(FULL, mine.end_tile, b);
            const unsigned magic = __shfl_sync(FULL, mine.magic, b);
            const int occ = (w0 >> 12) & 1, steep = (w0 >> 13) & 1;
            const int smaj = ((w0 >> 14) & 3) - 1, smin = ((w0 >> 16) & 3) - 1;
            const int D = w1 & 0x1fff, step_q = w1 >> 13, d2 = w2 & 0x3fff, step_e = w2 >> 14;
            const int D2 = 2 * D;
            const int n_occ = occ ? len - 1 : -1, n_near = occ ? len - 2 : -1;
            cached_sub = sub0;
            cached_base = base0;
            const uint32_t *lmaj = steep ? luty : lutx, *lmin = steep ? lutx : luty;
            const int hmaj = steep ? tyh : txh, hmin = steep ? txh : tyh;
            // LUT bit that says "this lattice cell shares its storage cell with the
            // next / previous cell of the ray along that axis" 
#define SH_MAJ_NEXT (smaj > 0 ? RB_LUT_NEXT_BIT : RB_LUT_PREV_BIT)
#define SH_MAJ_PREV (smaj > 0 ? RB_LUT_PREV_BIT : RB_LUT_NEXT_BIT)
#define SH_MIN_NEXT (smin > 0 ? RB_LUT_NEXT_BIT : RB_LUT_PREV_BIT)
#define SH_MIN_PREV (smin > 0 ? RB_LUT_PREV_BIT : RB_LUT_NEXT_BIT)
            // state of cell n = lane (closed form), then +32 cells per chunk incrementally:
            // minor(n) = floor((n*d2 + D) / D2), e = remainder
            int n = lane, e, kmaj, kmin;
            {
                // num <= 31 * 2D + D and D2 <= 8190, so num * D2 < 2^32 and the magic quotient is exact
                const unsigned num = (unsigned)lane * (unsigned)d2 + (unsigned)D;
                const unsigned m = __umulhi(num, magic);
                e = (int)(num - m * (unsigned)D2);
                kmaj = (steep ? sy : sx) + smaj * lane;
                kmin = (steep ? sx : sy) + smin * (int)m;
            }
            // Up to RC_INFLIGHT chunks (32 consecutive cells each) of the ray are in
            // flight together: cells of one ray are distinct storage cells except for
            // the aliasing pairs, which the earlier cell's lane applies in order.
            int n0 = 0;
            // Fast path: chunks whose 32 cells all have n <= len - 4 are plain "empty"
            // updates (-0.3, floor -3.0) whose successor is one too; no lane is idle, no
            // end/nearby handling, no world-border checks.
            if ((w0 >> 18) & 1) {
                const uint32_t *__restrict__ fmaj = lmaj + (800 * hmaj + 400), *__restrict__ fmin = lmin + (800 * hmin + 400);
                while (n0 + 32 <= len - 3) {
                    const int nch = min(RC_INFLIGHT, (len - 3 - n0) >> 5);
                    int8_t *addr[RC_INFLIGHT];
                    int dec[RC_INFLIGHT], t[RC_INFLIGHT];
#pragma unroll
                    for (int u = 0; u < RC_INFLIGHT; u++) {
                        addr[u] = nullptr; dec[u] = 0;
                        if (u >= nch) continue;                          // warp-uniform
                        const uint32_t pmaj = __ldg(&fmaj[kmaj]), pmin = __ldg(&fmin[kmin]);
                        const uint32_t px_ = steep ? pmin : pmaj, py_ = steep ? pmaj : pmin;
                        const int sub = (int)RB_LUT_SUB(py_) * subs_x + (int)RB_LUT_SUB(px_);
                        if (sub != cached_sub) {
                            const uint32_t tt = pt[sub];
                            cached_sub = sub;
                            cached_base = tt == RB_NONE ? nullptr : c.pool + (size_t)tt * RB_SUB_BYTES;
                            const int tile = (int)(RB_LUT_TILE(py_) * tiles_x + RB_LUT_TILE(px_));
                            if (!((ex_mask >> tile) & 1ull)) ex_new |= 1ull << tile;
                            if (!cached_base) atomicExch(&c.flags->world_overflow, 2);
                        }
                        int d_ = RB_T_EMP;
                        bool skip = false;
                        if (pmaj >> RB_LUT_NEXT_BIT) {                      // rare: an aliasing pair along the major axis
                            const bool bump_prev = e < d2, bump_next = e + d2 >= D2;
                            skip = n >= 1 && ((pmaj >> SH_MAJ_PREV) & 1u) && (!bump_prev || ((pmin >> SH_MIN_PREV) & 1u));
                            if (((pmaj >> SH_MAJ_NEXT) & 1u) && (!bump_next || ((pmin >> SH_MIN_NEXT) & 1u))) d_ = 2 * RB_T_EMP;
                        }
                        if (cached_base && !skip) {
                            addr[u] = cached_base + RB_LUT_OFF(py_) + RB_LUT_OFF(px_);
                            dec[u] = d_;
                        }
                        n += 32;
                        kmaj += 32 * smaj;
                        e += step_e;
                        int dq = step_q;
                        if (e >= D2) { e -= D2; dq++; }
                        kmin += smin * dq;
                    }
#pragma unroll
                    for (int u = 0; u < RC_INFLIGHT; u++) t[u] = addr[u] ? (int)*addr[u] : 0;
#pragma unroll
                    for (int u = 0; u < RC_INFLIGHT; u++)
                        if (addr[u]) {
                            const int v = max(t[u] - dec[u], -RB_T_MAX);
                            if (v != t[u]) *addr[u] = (int8_t)v;
                        }
                    __syncwarp();
                    n0 += 32 * nch;
                }
            }
            // general path: the last <= 34 cells of the ray (end / nearby handling, idle lanes)
            // and rays that leave the world
            for (; n0 < len; n0 += 32 * RC_INFLIGHT_G) {
                int ops[RC_INFLIGHT_G], t[RC_INFLIGHT_G];
                int8_t *addr[RC_INFLIGHT_G];
#pragma unroll
                for (int u = 0; u < RC_INFLIGHT_G; u++) {
                    ops[u] = 0; addr[u] = nullptr;
                    if (n0 + 32 * u >= len) continue;               // warp-uniform
                    if (n < len) {
                        const uint32_t pmaj = RC_LUT(lmaj, kmaj, hmaj), pmin = RC_LUT(lmin, kmin, hmin);
                        if (pmaj == RB_NONE || pmin == RB_NONE) {
                            dropped++;
                        } else {
                            const uint32_t px_ = steep ? pmin : pmaj, py_ = steep ? pmaj : pmin;
                            const int sub = (int)RB_LUT_SUB(py_) * subs_x + (int)RB_LUT_SUB(px_);
                            const int tile = (int)(RB_LUT_TILE(py_) * tiles_x + RB_LUT_TILE(px_));
                            if (sub != cached_sub) {
                                const uint32_t tt = pt[sub];
                                cached_sub = sub;
                                cached_base = tt == RB_NONE ? nullptr : c.pool + (size_t)tt * RB_SUB_BYTES;
                                if (!((ex_mask >> tile) & 1ull)) ex_new |= 1ull << tile;   // HybridMapEntry allocation :125-131
                                if (!cached_base) atomicExch(&c.flags->world_overflow, 2);  // cannot happen after prepare
                            }
                            // does this cell share its storage cell with the previous / next ray cell?
                            bool alias_prev = false, alias_next = false;
                            if (pmaj >> RB_LUT_NEXT_BIT) {                  // rare: an aliasing pair along the major axis
                                const bool bump_prev = e < d2, bump_next = e + d2 >= D2;
                                alias_prev = n >= 1 && ((pmaj >> SH_MAJ_PREV) & 1u) && (!bump_prev || ((pmin >> SH_MIN_PREV) & 1u));
                                alias_next = n + 1 < len && ((pmaj >> SH_MAJ_NEXT) & 1u) && (!bump_next || ((pmin >> SH_MIN_NEXT) & 1u));
                            }
                            if (cached_base && !alias_prev) {
                                int o = n == n_occ ? 2 : 1;                                 // hybridmap.py:137-138 / :144
                                if (n == n_near && tile == end_tile) o |= 4;                // hybridmap.py:139-142
                                if (alias_next) {                                           // next cell's ops, applied after ours
                                    int o2 = n + 1 == n_occ ? 2 : 1;
                                    if (n + 1 == n_near && tile == end_tile) o2 |= 4;
                                    o |= o2 << 3;
                                }
                                ops[u] = o;
                                addr[u] = cached_base + RB_LUT_OFF(py_) + RB_LUT_OFF(px_);   // blocked layout: the two axis parts add up
                            }
                        }
                    }
                    // advance this lane by 32 cells
                    n += 32;
                    kmaj += 32 * smaj;
                    e += step_e;
                    int dq = step_q;
                    if (e >= D2) { e -= D2; dq++; }
                    kmin += smin * dq;
                }
#pragma unroll
                for (int u = 0; u < RC_INFLIGHT_G; u++) t[u] = ops[u] ? (int)*addr[u] : 0;
#pragma unroll
                for (int u = 0; u < RC_INFLIGHT_G; u++)
                    if (ops[u]) {
                        int v = apply_ops(t[u], ops[u] & 7);
                        if (ops[u] >> 3) v = apply_ops(v, ops[u] >> 3);
                        if (v != t[u]) *addr[u] = (int8_t)v;        // saturated cells (most of a built map) are not rewritten
                    }
                __syncwarp();                                       // the next beam must see these stores
            }
        }
    }
    // publish newly created reference tiles and dropped-cell count
    for (int o = 16; o > 0; o >>= 1) {
        ex_new |= __shfl_xor_sync(FULL, ex_new, o);
        dropped += __shfl_xor_sync(FULL, dropped, o);
    }
    if (lane == 0) {
        if (ex_new) c.exists[p] = ex_mask | ex_new;
        if (dropped) atomicAdd(&c.stats->cells_dropped, (unsigned long long)dropped);
    }
}

// ------------------------------------------------------- cast, version 2 --
// Same update order as raycast_cast_kernel, a third of the instructions.  What changed:
//   * persistent CTAs (grid = resident CTAs), warps fetch particles from a work counter, so the
//     two cast LUTs (800 entries per reference tile and axis) are staged in shared memory once
//     per CTA instead of once per four particles;
//   * the cast LUT entries of the two axes ADD UP to one word: bits 0-14 byte offset inside the
//     sub-tile, bits 15-25 page-table slot, bits 28-31 the aliasing flags (x: 30/31, y: 28/29);
//   * the page-table entries a sweep can reach (5 x 5 sub-tiles around the robot's: rays are
//     clipped at 300 cells, a sub-tile has 160) sit in a per-warp table in shared memory, indexed
//     by slot - slot of the window corner: no per-lane cache, no branch on a sub-tile change;
//   * new reference tiles (HybridMapEntry allocation, hybridmap.py:125-131) are found per beam
//     from the tiles of the two end cells; a ray that changes tile along both axes takes the
//     per-cell path;
//   * the minor coordinate of cell n is the closed form floor((n d2 + D) / D2) by a magic
//     multiply (exact: n <= 300 cells after the 15 m clip), no error term is carried;
//   * a ray is ceil(len / 32) chunks: "empty" chunks over [0, len - 32) and ONE tail chunk
//     [len - 32, len) with the end / nearby cells in lanes 31 / 30; every cell's update is
//     v = min(max(t - a, -30) + b, 30) with (a, b) per lane, which also covers two ray cells
//     that share a storage cell (the earlier lane applies both, the later one nothing);
//   * per-beam constants go through shared memory (three 16-byte broadcasts per beam).
// Rays that leave the world take cast_general_ray (the old per-cell path).
#ifndef RC2_WARPS
#define RC2_WARPS 16                 // 2 CTAs of 16 warps per SM: the staged LUTs take 2 x 32 KB, the rest of the 256 KB stays L1
#endif                               // (8 warps x 4 CTAs: 6.67 ms, 16 x 2: 6.32, 32 x 1: 6.26 at 65,536 particles)
#ifndef RC2_MINBLOCKS
#define RC2_MINBLOCKS 2
#endif

#define RC2_OFFMASK 0x00007fffu
#define RC2_PART_BEAMS 96            // beams per work item in split mode (multiple of 32)
#ifndef RC2_SPLIT_BELOW
#define RC2_SPLIT_BELOW 4             // whole sweeps when every resident warp gets at least this many particles (measured: parts win 6 % at 8,192 particles, 1 % at 16,384, lose 2 % at 32,768)
#endif
#define RC2_TBL_ENTRIES 168          // slot table of a warp: 4 * subs_x + 5 <= 165 sub-tile base pointers (subs_x <= 40)
#define RC2_WARP_WORDS (384 + 16)    // per warp: 32 records of 48 bytes, the particle's frame
static_assert(RC2_WARPS * RC2_TBL_ENTRIES <= 8192, "slot-table entry index has 13 bits");

// frame words (doubles): 0 x, 1 y, 2 cos, 3 sin; then ints at double index 4: sx, sy
__device__ __forceinline__ Ray ray_of_beam_p(const double *px, const double *py, const double *dist, int j, double x, double y,
                                             double cs_, double sn_, int sx, int sy)
{
    Ray r;
    double gx, gy;
    rb_xform(cs_, sn_, x, y, px[j], py[j], gx, gy);
    r.ex = rb_trunc(gx / RB_CS);                                   // hybridmap.py:106
    r.ey = rb_trunc(gy / RB_CS);
    r.occ = 1;
    const double d = dist[j];
    if (d > RB_CLIP_R) {                                           // hybridmap.py:107-113
        const double scale = 15.0 / d;
        const int nex = rb_trunc((double)sx + scale * (double)(r.ex - sx));
        const int ney = rb_trunc((double)sy + scale * (double)(r.ey - sy));
        r.ex = nex;
        r.ey = ney;
        r.occ = 0;
    }
    const int dx = r.ex - sx, dy = r.ey - sy;
    const int adx = abs(dx), ady = abs(dy);
    if (adx == 0) r.len = dy >= 0 ? dy + 1 : 0;                    // hybridmap.py:278-279
    else if (ady == 0) r.len = dx >= 0 ? dx + 1 : 0;               // hybridmap.py:280-281
    else r.len = max(adx, ady) + 1;
    return r;
}

// The old per-cell path for one beam (all lanes on the same beam) -- rays with a cell outside the world and rays
// that change reference tile along both axes (rare: the world is sized for the log, tiles are 800 cells).
// Scalars in, the newly touched reference tiles out: nothing of the caller has its address taken.
__device__ __noinline__ unsigned long long cast_general_ray(int8_t *pool, const uint32_t *pt, const uint32_t *lutx, const uint32_t *luty,
                                                            RbFlags *flags, RbStats *stats, int txh, int tyh, int subs_x, int tiles_x,
                                                            int lane, const double *px, const double *py, const double *dist, int j,
                                                            const double *frame, unsigned long long ex_mask)
{
    unsigned long long ex_new = 0ull;
    unsigned dropped = 0;
    const int sx = reinterpret_cast<const int *>(frame + 4)[0], sy = reinterpret_cast<const int *>(frame + 4)[1];
    const Ray r = ray_of_beam_p(px, py, dist, j, frame[0], frame[1], frame[2], frame[3], sx, sy);
    const int len = r.len, occ = r.occ;
    const uint32_t pex = rb_write_lut(lutx, r.ex, txh), pey = rb_write_lut(luty, r.ey, tyh);
    const int end_tile = (pex == RB_NONE || pey == RB_NONE) ? -1 : (int)(RB_LUT_TILE(pey) * tiles_x + RB_LUT_TILE(pex));
    const RayStep st = ray_step(sx, sy, r);
    const int steep = st.steep, smaj = st.smaj, smin = st.smin;
    const int D2 = (int)st.D2, d2 = (int)st.d2;
    const int step_q = (int)((32u * st.d2) / st.D2), step_e = (int)(32u * st.d2 - (unsigned)step_q * st.D2);
    const int n_occ = occ ? len - 1 : -1, n_near = occ ? len - 2 : -1;
    const uint32_t *lmaj = steep ? luty : lutx, *lmin = steep ? lutx : luty;
    const int hmaj = steep ? tyh : txh, hmin = steep ? txh : tyh;
    int n = lane, e, kmaj, kmin;
    {
        const unsigned num = (unsigned)lane * st.d2 + st.D;
        const unsigned m = num / st.D2;
        e = (int)(num - m * st.D2);
        kmaj = (steep ? sy : sx) + smaj * lane;
        kmin = (steep ? sx : sy) + smin * (int)m;
    }
    int cached_sub = -1;
    int8_t *cached_base = nullptr;
    for (int n0 = 0; n0 < len; n0 += 32) {
        int ops = 0;
        int8_t *addr = nullptr;
        if (n < len) {
            const uint32_t pmaj = rb_write_lut(lmaj, kmaj, hmaj), pmin = rb_write_lut(lmin, kmin, hmin);
            if (pmaj == RB_NONE || pmin == RB_NONE) {
                dropped++;
            } else {
                const uint32_t px_ = steep ? pmin : pmaj, py_ = steep ? pmaj : pmin;
                const int sub = (int)RB_LUT_SUB(py_) * subs_x + (int)RB_LUT_SUB(px_);
                const int tile = (int)(RB_LUT_TILE(py_) * tiles_x + RB_LUT_TILE(px_));
                if (sub != cached_sub) {
                    const uint32_t tt = pt[sub];
                    cached_sub = sub;
                    cached_base = tt == RB_NONE ? nullptr : pool + (size_t)tt * RB_SUB_BYTES;
                    if (!((ex_mask >> tile) & 1ull)) ex_new |= 1ull << tile;   // HybridMapEntry allocation :125-131
                    if (!cached_base) atomicExch(&flags->world_overflow, 2);    // cannot happen after prepare
                }
                bool alias_prev = false, alias_next = false;
                if (pmaj >> RB_LUT_NEXT_BIT) {
                    const bool bump_prev = e < d2, bump_next = e + d2 >= D2;
                    alias_prev = n >= 1 && ((pmaj >> SH_MAJ_PREV) & 1u) && (!bump_prev || ((pmin >> SH_MIN_PREV) & 1u));
                    alias_next = n + 1 < len && ((pmaj >> SH_MAJ_NEXT) & 1u) && (!bump_next || ((pmin >> SH_MIN_NEXT) & 1u));
                }
                if (cached_base && !alias_prev) {
                    int o = n == n_occ ? 2 : 1;                                 // hybridmap.py:137-138 / :144
                    if (n == n_near && tile == end_tile) o |= 4;                // hybridmap.py:139-142
                    if (alias_next) {
                        int o2 = n + 1 == n_occ ? 2 : 1;
                        if (n + 1 == n_near && tile == end_tile) o2 |= 4;
                        o |= o2 << 3;
                    }
                    ops = o;
                    addr = cached_base + RB_LUT_OFF(py_) + RB_LUT_OFF(px_);
                }
            }
        }
        n += 32;
        kmaj += 32 * smaj;
        e += step_e;
        int dq = step_q;
        if (e >= D2) { e -= D2; dq++; }
        kmin += smin * dq;
        if (ops) {
            const int t = (int)*addr;
            int v = apply_ops(t, ops & 7);
            if (ops >> 3) v = apply_ops(v, ops >> 3);
            if (v != t) *addr = (int8_t)v;
        }
        __syncwarp();
    }
    if (dropped) atomicAdd(&stats->cells_dropped, (unsigned long long)dropped);
    return ex_new;
}

// Per-beam constants of beams j0 .. j0 + 31 (one beam per lane) into the warp's record table; returns the
// reference tiles this lane's beam creates.  Out of line: the float64 frame and the per-beam geometry do not
// take registers away from the cell loop.  A record is 12 words (three 16-byte broadcasts per beam):
//   0    : len (bits 0-11) | occ << 12 | steep << 13 | fast << 14 | near_ok << 15 | (step along the major axis < 0) << 16
//   1-3  : d2 = 2 * minor extent, D = major extent (>= 1), magic = ceil(2^32 / 2 D)
//   4-5  : byte offsets (from the start of the dynamic shared memory) of the LUT entries of cell 0's major / minor coordinate
//   6-7  : 4 * step along the major axis, 4 * step along the minor axis (bytes per LUT entry)
//   8    : the two aliasing flags of the major axis inside the LUT sum
//   9    : chunks of 32 cells, tail included
//   10-11: updates (a | b << 8) of the end cell (+0.8 if it is an obstacle, hybridmap.py:137-138, else -0.3, :144) and of
//          the cell before it (-0.3 when the ray passed, then +0.2 if the end's tile holds it, :139-142)
__device__ __noinline__ unsigned long long cast_setup_beams(const double *px, const double *py, const double *dist, const uint32_t *lutx,
                                                            const uint32_t *luty, int txh, int tyh, int tiles_x, int B, int j0, int lane,
                                                            const double *frame, int4 *recs, int nx, int ny, int lut_off,
                                                            unsigned long long ex_mask)
{
    const int ox = 800 * txh + 400, oy = 800 * tyh + 400;
    unsigned long long ex_new = 0ull;
    int4 ra = make_int4(0, 0, 1, 0), rb = make_int4(0, 0, 4, 0), rc = make_int4(0, 1, RB_T_EMP, RB_T_EMP);
    if (j0 + lane < B) {
        const int sx = reinterpret_cast<const int *>(frame + 4)[0], sy = reinterpret_cast<const int *>(frame + 4)[1];
        const Ray r = ray_of_beam_p(px, py, dist, j0 + lane, frame[0], frame[1], frame[2], frame[3], sx, sy);
        const RayStep st = ray_step(sx, sy, r);
        const uint32_t psx = rb_write_lut(lutx, sx, txh), psy = rb_write_lut(luty, sy, tyh);
        const uint32_t pex = rb_write_lut(lutx, r.ex, txh), pey = rb_write_lut(luty, r.ey, tyh);
        int len = r.len;
        if (len > 4095 || st.D > 4095u) len = 0;                       // cannot happen: rays are clipped at 15 m = 300 cells
        // start and end cell inside the world => every cell of the ray is (bounding box);
        // the magic quotient is exact for D <= 700 and the slot table covers 300 cells (the 15 m clip keeps D at 300)
        int fast = pex != RB_NONE && pey != RB_NONE && psx != RB_NONE && psy != RB_NONE && st.D <= 300u && len <= 301;
        // the end cell (hence every cell: rays are monotone along both axes) lies in the 5 x 5 sub-tile window of the slot table
        if (fast && (abs((int)RB_LUT_SUB(pex) - (int)RB_LUT_SUB(psx)) > 2 || abs((int)RB_LUT_SUB(pey) - (int)RB_LUT_SUB(psy)) > 2)) fast = 0;
        int near_ok = 0;
        if (fast && len > 0) {
            const int tsx = (int)RB_LUT_TILE(psx), tsy = (int)RB_LUT_TILE(psy), tex = (int)RB_LUT_TILE(pex), tey = (int)RB_LUT_TILE(pey);
            const int end_tile = tey * tiles_x + tex;
            if (tsx != tex && tsy != tey) {
                fast = 0;                                               // which corner tile the ray crosses is the per-cell path's business
            } else {
                // one axis changes tile at most once (300 < 800 cells): the ray's cells lie in the start tile (exists,
                // hybridmap.py:98-100) and the end tile
                if (!((ex_mask >> end_tile) & 1ull)) ex_new = 1ull << end_tile;
                if (r.occ && len >= 2) {                                // hybridmap.py:139-142: the cell before the end, if the end's tile holds it
                    int qx, qy;
                    ray_cell(sx, sy, st, len - 2, qx, qy);
                    const uint32_t pqx = rb_write_lut(lutx, qx, txh), pqy = rb_write_lut(luty, qy, tyh);
                    near_ok = (int)(RB_LUT_TILE(pqy) * tiles_x + RB_LUT_TILE(pqx)) == end_tile;
                }
            }
        }
        ra.x = len | (r.occ << 12) | (st.steep << 13) | (fast << 14) | (near_ok << 15) | ((st.smaj < 0) << 16);
        ra.y = (int)st.d2;
        ra.z = (int)st.D;
        ra.w = (int)(unsigned)((0x100000000ull + st.D2 - 1) / st.D2);
        rb.x = lut_off + 4 * (st.steep ? nx + sy + oy : sx + ox);
        rb.y = lut_off + 4 * (st.steep ? sx + ox : nx + sy + oy);
        rb.z = 4 * st.smaj;
        rb.w = 4 * st.smin;
        rc.x = (int)(3u << (st.steep ? 28 : 30));
        rc.y = max((len + 31) >> 5, 1);
        rc.z = r.occ ? (RB_T_OCC << 8) : RB_T_EMP;
        rc.w = near_ok ? (RB_T_EMP | (RB_T_NEAR << 8)) : RB_T_EMP;
    }
    recs[3 * lane] = ra;
    recs[3 * lane + 1] = rb;
    recs[3 * lane + 2] = rc;
    return ex_new;
}

struct CastBeam {    // warp-uniform constants of the beam in flight
    const char *smem;          // start of the dynamic shared memory = start of the slot tables
    int pminb;                 // byte offset of the LUT entry of the minor coordinate of cell 0
    int smin4;                 // 4 * step along the minor axis (bytes per LUT entry)
    unsigned kadd;             // (first entry of this warp's slot table - slot of the window corner) << 15
    int w0, d2, Dm;            // record word 0, 2 * minor extent, major extent
    unsigned magic, amask;     // amask: the two aliasing flags of the major axis inside x + y
    int ab_end, ab_near;       // updates (a | b << 8) of the end cell and of the cell before it
    int8_t *sink;              // slot-table value of an unallocated sub-tile (the spare sub-tile behind the pool)
    int *overflow;
};

// Cell n of the ray (pmaj = byte offset of the LUT entry of its major coordinate, num = n d2 + D): address of its
// storage cell and its update (a | b << 8).  ab_in = update if the cell had its storage cell to itself.
template <bool TAIL>
__device__ __forceinline__ int8_t *cast_cell(const CastBeam &k, int lane, int n, int pmaj, unsigned num, int ab_in, int &ab)
{
    const unsigned m = __umulhi(num, k.magic);
    // offset (bits 0-14) | entry of the slot table (15-27) | aliasing flags (28-31)
    const unsigned s = *reinterpret_cast<const uint32_t *>(k.smem + pmaj) +
                       *reinterpret_cast<const uint32_t *>(k.smem + (k.pminb + k.smin4 * (int)m)) + k.kadd;
    int8_t *base = *reinterpret_cast<int8_t *const *>(k.smem + ((s >> 12) & 0xfff8u));
    if (TAIL && base == k.sink && ab_in) atomicExch(k.overflow, 2);  // an unallocated sub-tile under the ray's end: cannot happen after prepare
    ab = ab_in;
    if (s & k.amask) {                                               // shares its storage cell with a neighbour along the major axis
        const int steep = (k.w0 >> 13) & 1, len = k.w0 & 0xfff, D2 = 2 * k.Dm;
        const int sh_maj = steep ? 28 : 30, sh_min = steep ? 30 : 28;
        const unsigned A = (s >> sh_maj) & 3u, Bm = (s >> sh_min) & 3u;       // bit 0: with k + 1, bit 1: with k - 1
        const int e = (int)(num - m * (unsigned)D2);
        const bool bump_prev = e < k.d2, bump_next = e + k.d2 >= D2;          // the minor coordinate changes from n - 1 / to n + 1
        const unsigned f = (k.w0 >> 16) & 1 ? 2u : 1u, bk = 3u - f;           // flag of the direction of travel along the major axis
        const unsigned fm = k.smin4 > 0 ? 1u : 2u, bm = 3u - fm;
        const bool with_prev = n >= 1 && (A & bk) && (!bump_prev || (Bm & bm));
        const bool with_next = n + 1 < len && (A & f) && (!bump_next || (Bm & fm));
        const int ab_next = !TAIL ? RB_T_EMP : lane == 30 ? k.ab_end : lane == 29 ? k.ab_near : RB_T_EMP;
        if (ab_in) ab = with_prev ? 0 : with_next ? ab_in + ab_next : ab_in;  // the earlier cell's lane applies both, in order
    }
    return base + (s & RC2_OFFMASK);
}

// NCH chunks of 32 consecutive cells, the loads of all chunks in flight together.  TAIL: the last chunk is the
// ray's tail [len - 32, len) and the one before it the last of the "empty" chunks (-0.3, floor -3.0), whose
// lanes past cell len - 33 load a cell of the ray and store nothing; all other chunks are 32 empty cells.
template <int NCH, bool TAIL>
__device__ __forceinline__ void cast_group(const CastBeam &k, int lane, int &pmaj, unsigned &num, int &n, int nf, int smaj128,
                                           unsigned d2x32, int t_pmaj, unsigned t_num, int t_n, int t_ab)
{
    int8_t *addr[NCH];
    int ab[NCH], t[NCH];
#pragma unroll
    for (int u = 0; u < NCH; u++) {
        if (TAIL && u == NCH - 1) {
            addr[u] = cast_cell<true>(k, lane, t_n, t_pmaj, t_num, t_ab, ab[u]);
        } else {
            // the last empty chunk sits before the tail, or closes a full group when only the tail is left after it
            const bool maybe_partial = TAIL ? u == NCH - 2 : u == NCH - 1;
            addr[u] = cast_cell<false>(k, lane, n, pmaj, num, maybe_partial ? (n < nf ? RB_T_EMP : 0) : RB_T_EMP, ab[u]);
            n += 32;
            pmaj += smaj128;
            num += d2x32;
        }
    }
#pragma unroll
    for (int u = 0; u < NCH; u++) t[u] = (int)*addr[u];
#pragma unroll
    for (int u = 0; u < NCH; u++) {
        int v;
        if (TAIL && u == NCH - 1) v = min(max(t[u] - (ab[u] & 0xff), -RB_T_MAX) + (ab[u] >> 8), RB_T_MAX);
        else v = max(t[u] - ab[u], -RB_T_MAX);                        // ab 0: v == t (cells are never below the floor)
        if (v != t[u]) *addr[u] = (int8_t)v;                          // saturated cells (most of a built map) are not rewritten
    }
}

__global__ void __launch_bounds__(RC2_WARPS * 32, RC2_MINBLOCKS) raycast_cast2_kernel(RbCtx c)
{
    extern __shared__ __align__(16) uint32_t smem2[];
    const unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nx = 800 * c.tiles_x, ny = 800 * c.tiles_y;
    // slot tables first (their entry index has to fit the 13 bits between offset and flags), then the LUTs, then records + frames
    int8_t **tbl = reinterpret_cast<int8_t **>(smem2) + warp * RC2_TBL_ENTRIES;
    uint32_t *lut_s = smem2 + 2 * RC2_WARPS * RC2_TBL_ENTRIES;       // x entries, then y entries
    int4 *recs = reinterpret_cast<int4 *>(lut_s + nx + ny + warp * RC2_WARP_WORDS);
    double *frame = reinterpret_cast<double *>(lut_s + nx + ny + warp * RC2_WARP_WORDS + 384);
    for (int i = threadIdx.x; i < nx + ny; i += RC2_WARPS * 32) lut_s[i] = c.clut[i];
    __syncthreads();
    if (c.flags->pool_exhausted) return;                             // prepare could not privatise: skip the scan
    const int subs_x = c.subs_x;
    int8_t *const sink = c.pool + (size_t)c.pool_tiles * RB_SUB_BYTES;
    const int lane_m32 = lane - 32;

    // Work items: whole sweeps (cast_ipp == 1), or -- when a GPU holds so few particles that whole sweeps would leave the
    // last round of the persistent grid half empty (8,192 particles on 4,736 resident warps: two rounds for 1.7 rounds of
    // work) -- parts of RC2_PART_BEAMS beams.  Item i is part i / N of particle i % N: every first part is handed out
    // before any second part, a part waits for its predecessor's release (which is running or done: items are handed out
    // in order and every CTA of the grid is resident), so the beams of a particle are still applied in order.
    const int ipp = c.cast_ipp, n_items = c.N * ipp;
    for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(c.cast_work, 1);
        it = __shfl_sync(FULL, it, 0);
        if (it >= n_items) break;
        const int part = it / c.N, p = it - part * c.N;
        const int j_begin = ipp > 1 ? part * RC2_PART_BEAMS : 0, j_end = ipp > 1 ? min(c.B, j_begin + RC2_PART_BEAMS) : c.B;
        if (part > 0) {
            if (lane == 0) {
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(c.cast_done + p) : "memory");
                    if (seen < part) __nanosleep(200);
                } while (seen < part);
            }
            __syncwarp();
        }
#define RC2_RELEASE_PART()                                                                                       \
        if (ipp > 1) {                                                                                           \
            __threadfence();                                                                                     \
            __syncwarp();                                                                                        \
            if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(c.cast_done + p), "r"(part + 1) : "memory"); \
        }
        const uint32_t *pt = c.pt + (size_t)p * c.nsub;
        int sub_lo = 0;                                              // slot of the window corner (may lie outside the world)
        {
            double x, y, cs_, sn_;
            int sx, sy;
            if (!particle_frame(c, p, x, y, cs_, sn_, sx, sy)) { RC2_RELEASE_PART() continue; }
            __syncwarp();
            if (lane == 0) {
                frame[0] = x; frame[1] = y; frame[2] = cs_; frame[3] = sn_;
                reinterpret_cast<int *>(frame + 4)[0] = sx;
                reinterpret_cast<int *>(frame + 4)[1] = sy;
            }
            // the 5 x 5 sub-tiles around the robot's; unallocated -> the sink sub-tile behind the pool
            const int ox = 800 * c.txh + 400, oy = 800 * c.tyh + 400;
            if ((unsigned)(sx + ox) < (unsigned)nx && (unsigned)(sy + oy) < (unsigned)ny) {
                const unsigned s0 = lut_s[sx + ox] + lut_s[nx + sy + oy];
                const int sub0 = (int)((s0 >> 15) & 0x7ffu);
                const int subx0 = sub0 % subs_x, suby0 = sub0 / subs_x;
                sub_lo = (suby0 - 2) * subs_x + (subx0 - 2);
                if (lane < 25) {
                    const int dy = lane / 5, dx = lane - 5 * dy;
                    const int qx = subx0 - 2 + dx, qy = suby0 - 2 + dy;
                    uint32_t tt = RB_NONE;
                    if ((unsigned)qx < (unsigned)subs_x && (unsigned)qy < (unsigned)c.subs_y) tt = pt[qy * subs_x + qx];
                    tbl[dy * subs_x + dx] = tt == RB_NONE ? sink : c.pool + (size_t)tt * RB_SUB_BYTES;
                }
            }
        }
        const unsigned long long ex_mask = c.exists[p];
        unsigned long long ex_new = 0ull;
        CastBeam k;
        k.smem = reinterpret_cast<const char *>(smem2);
        k.kadd = (unsigned)(warp * RC2_TBL_ENTRIES - sub_lo) << 15;
        k.sink = sink;
        k.overflow = &c.flags->world_overflow;

        for (int j0 = j_begin; j0 < j_end; j0 += 32) {
            __syncwarp();
            ex_new |= cast_setup_beams(c.px, c.py, c.dist, c.lutx, c.luty, c.txh, c.tyh, c.tiles_x, j_end, j0, lane, frame, recs, nx, ny,
                                       8 * RC2_WARPS * RC2_TBL_ENTRIES, ex_mask);
            __syncwarp();
            const int nb = min(32, j_end - j0);
            for (int b = 0; b < nb; b++) {
                const int4 ra = recs[3 * b];
                const int w0 = ra.x, len = w0 & 0xfff;
                if (len == 0) continue;                              // hybridmap.py:278-281 empty list
                if (!((w0 >> 14) & 1)) {
                    ex_new |= cast_general_ray(c.pool, pt, c.lutx, c.luty, c.flags, c.stats, c.txh, c.tyh, subs_x, c.tiles_x, lane,
                                               c.px, c.py, c.dist, j0 + b, frame, ex_mask);
                    continue;
                }
                const int4 rb = recs[3 * b + 1], rc = recs[3 * b + 2];
                k.w0 = w0; k.d2 = ra.y; k.Dm = ra.z; k.magic = (unsigned)ra.w;
                k.pminb = rb.y; k.smin4 = rb.w;
                k.amask = (unsigned)rc.x; k.ab_end = rc.z; k.ab_near = rc.w;
                const int smaj4 = rb.z;
                // tail chunk: cells [len - 32, len), lane 31 = end cell, lane 30 = the cell before it; lanes before the ray's
                // start (len < 32) sit on cell 0 and do nothing
                const int nt_ = len + lane_m32, nt = max(nt_, 0);
                int t_ab = lane == 31 ? k.ab_end : lane == 30 ? k.ab_near : RB_T_EMP;
                if (nt_ < 0) t_ab = 0;
                const int t_pmaj = rb.x + smaj4 * nt;
                const unsigned t_num = (unsigned)nt * (unsigned)k.d2 + (unsigned)k.Dm;
                // empty chunks over [0, len - 32)
                const int nf = len - 32;
                int pmaj = rb.x + smaj4 * lane, n = lane;
                unsigned num = (unsigned)lane * (unsigned)k.d2 + (unsigned)k.Dm;
                const unsigned d2x32 = 32u * (unsigned)k.d2;
                const int smaj128 = 32 * smaj4;
                int left = rc.y;                                     // chunks of the ray, tail included
#define RC2_GROUP(K, T) cast_group<K, T>(k, lane, pmaj, num, n, nf, smaj128, d2x32, t_pmaj, t_num, nt, t_ab)
#ifndef RC2_NCH
#define RC2_NCH 4                    // chunks of 32 cells whose loads are in flight together
#endif
                while (left > RC2_NCH) { RC2_GROUP(RC2_NCH, false); left -= RC2_NCH; }
#if RC2_NCH >= 6
                if (left == 6) RC2_GROUP(6, true);
                else
#endif
#if RC2_NCH >= 5
                if (left == 5) RC2_GROUP(5, true);
                else
#endif
                if (left == 4) RC2_GROUP(4, true);
                else if (left == 3) RC2_GROUP(3, true);
                else if (left == 2) RC2_GROUP(2, true);
                else RC2_GROUP(1, true);
                __syncwarp();                                        // the next beam must see these stores
            }
        }
        // publish newly created reference tiles
        for (int o = 16; o > 0; o >>= 1) ex_new |= __shfl_xor_sync(FULL, ex_new, o);
        if (lane == 0 && ex_new) c.exists[p] = ex_mask | ex_new;
        RC2_RELEASE_PART()
    }
#undef RC2_RELEASE_PART
}

// ------------------------------------------------------- atomics variant --
// EXPERIMENT, not the product path (RBPF_CAST_ATOMICS=1): the variant BASELINE.json's north_star names.
// The beams of a particle are split over the four warps of a CTA and every cell update is a
// compare-and-swap on the 32-bit word that holds the byte (there are no byte atomics), in whatever
// order the warps get there.  It measures the best case of an atomic formulation: it does NOT replay
// the reference's beam order, so a cell that both saturates and is touched by several beams of one
// sweep can end up different -- an exact version would add a membership test against
// the <= 2 B order-sensitive cells to every update and a sequential replay of those.  Numbers in
// profiles/README.md; the ordered kernel above stays the default because it is exact and not slower.
__global__ void __launch_bounds__(RC_WARPS * 32) raycast_cast_atomic_kernel(RbCtx c)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x;
    if (p >= c.N || c.flags->pool_exhausted) return;
    double x, y, cs_, sn_;
    int sx, sy;
    if (!particle_frame(c, p, x, y, cs_, sn_, sx, sy)) return;
    const uint32_t *pt = c.pt + (size_t)p * c.nsub;
    const unsigned long long ex_mask = c.exists[p];
    unsigned long long ex_new = 0ull;
    unsigned dropped = 0;
    for (int j = warp; j < c.B; j += RC_WARPS) {
        const Ray r = ray_of_beam(c, j, x, y, cs_, sn_, sx, sy);
        if (r.len == 0) continue;
        const RayStep st = ray_step(sx, sy, r);
        const uint32_t pex = rb_write_lut(c.lutx, r.ex, c.txh), pey = rb_write_lut(c.luty, r.ey, c.tyh);
        const int end_tile = (pex == RB_NONE || pey == RB_NONE) ? -1 : (int)(RB_LUT_TILE(pey) * c.tiles_x + RB_LUT_TILE(pex));
        for (int n = lane; n < r.len; n += 32) {
            int kx, ky;
            ray_cell(sx, sy, st, n, kx, ky);
            const uint32_t px_ = rb_write_lut(c.lutx, kx, c.txh), py_ = rb_write_lut(c.luty, ky, c.tyh);
            if (px_ == RB_NONE || py_ == RB_NONE) { dropped++; continue; }
            const int sub = (int)RB_LUT_SUB(py_) * c.subs_x + (int)RB_LUT_SUB(px_);
            const int tile = (int)(RB_LUT_TILE(py_) * c.tiles_x + RB_LUT_TILE(px_));
            if (!((ex_mask >> tile) & 1ull)) ex_new |= 1ull << tile;
            const uint32_t tt = pt[sub];
            if (tt == RB_NONE) continue;
            int ops = (r.occ && n == r.len - 1) ? 2 : 1;
            if (r.occ && n == r.len - 2 && tile == end_tile) ops |= 4;
            int8_t *a = c.pool + (size_t)tt * RB_SUB_BYTES + RB_LUT_OFF(py_) + RB_LUT_OFF(px_);
            uint32_t *wp = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(a) & ~(uintptr_t)3);
            const int sh8 = (int)(reinterpret_cast<uintptr_t>(a) & 3) * 8;
            uint32_t old = *wp;
            for (;;) {
                const int t = (int)(int8_t)(old >> sh8);
                const int v = apply_ops(t, ops);
                if (v == t) break;                                           // saturated: nothing to write
                const uint32_t want = (old & ~(0xffu << sh8)) | ((uint32_t)(uint8_t)(int8_t)v << sh8);
                const uint32_t seen = atomicCAS(wp, old, want);
                if (seen == old) break;
                old = seen;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        ex_new |= __shfl_xor_sync(0xffffffffu, ex_new, o);
        dropped += __shfl_xor_sync(0xffffffffu, dropped, o);
    }
    if (lane == 0) {
        if (ex_new) atomicOr(&c.exists[p], ex_new);
        if (dropped) atomicAdd(&c.stats->cells_dropped, (unsigned long long)dropped);
    }
}

void rb_launch_raycast_prepare(const RbCtx &c, cudaStream_t s)
{
    raycast_prepare_kernel<<<(c.N + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, s>>>(c);
}

void rb_launch_raycast_cast(const RbCtx &c, cudaStream_t s)
{
    static const int use_atomics = getenv("RBPF_CAST_ATOMICS") && atoi(getenv("RBPF_CAST_ATOMICS")) > 0;
    if (use_atomics) {                                            // experiment only, see raycast_cast_atomic_kernel
        raycast_cast_atomic_kernel<<<c.N, RC_WARPS * 32, 0, s>>>(c);
        return;
    }
    static const int use_v1 = getenv("RBPF_CAST_V1") && atoi(getenv("RBPF_CAST_V1")) > 0;
    const size_t lut_bytes2 = sizeof(uint32_t) * 800 * (size_t)(c.tiles_x + c.tiles_y);
    if (!use_v1 && c.subs_x <= 40 && lut_bytes2 <= 64 * 1024) {      // larger worlds: the staged LUTs / the slot table do not fit
        const size_t smem = sizeof(uint32_t) * (800 * (size_t)(c.tiles_x + c.tiles_y) + RC2_WARPS * (RC2_WARP_WORDS + 2 * RC2_TBL_ENTRIES));
        static size_t smem_set = 0;
        static int resident = 0;
        if (smem != smem_set) {
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaFuncSetAttribute(raycast_cast2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raycast_cast2_kernel, RC2_WARPS * 32, smem);
            resident = sms * (per_sm > 0 ? per_sm : 1);
            smem_set = smem;
        }
        // split sweeps into parts when whole sweeps would fill fewer than RC2_SPLIT_BELOW rounds of the resident warps
        static const int force_ipp = getenv("RBPF_CAST_PARTS") ? atoi(getenv("RBPF_CAST_PARTS")) : 0;      // 1: never split (A/B)
        RbCtx d = c;
        d.cast_ipp = 1;
        if (force_ipp != 1 && c.N < RC2_SPLIT_BELOW * resident * RC2_WARPS) d.cast_ipp = (c.B + RC2_PART_BEAMS - 1) / RC2_PART_BEAMS;
        if (d.cast_ipp > 1) cudaMemsetAsync(c.cast_done, 0, sizeof(int) * (size_t)c.N, s);
        const long long items = (long long)c.N * d.cast_ipp;
        const int want = (int)((items + RC2_WARPS - 1) / RC2_WARPS);
        raycast_cast2_kernel<<<want < resident ? want : resident, RC2_WARPS * 32, smem, s>>>(d);
        return;
    }
    const int blocks = (c.N + RC_WARPS - 1) / RC_WARPS;
#ifdef RC_SMEM_LUT   // measured slower on B200 (5.55 vs 4.02 ms at 16,384 particles): staging 32 KB per CTA costs more than the L1 hits it saves
    const size_t lut_bytes = sizeof(uint32_t) * 800 * (size_t)(c.tiles_x + c.tiles_y);
    if (lut_bytes <= 64 * 1024) {
        cudaFuncSetAttribute(raycast_cast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        raycast_cast_kernel<true><<<blocks, RC_WARPS * 32, lut_bytes, s>>>(c);
        return;
    }
#endif
    raycast_cast_kernel<false><<<blocks, RC_WARPS * 32, 0, s>>>(c);
}
