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
//                    order, lanes across the cells of one ray (closed-form
//                    Bresenham), saturating int8 read-modify-write.  The
//                    reference's clamps make the result order-dependent
//                    (SURVEY 3.4-4), hence no atomics and no beam parallelism
//                    inside a particle.
#include "common.cuh"

#define RC_WARPS 4

struct Ray {
    int ex, ey;      // end cell (lattice)
    int len;         // number of cells (0 for the reference's empty-list quirk)
    int occ;         // end cell is an obstacle (range <= 15 m)
};

// End cell and length of beam j for a particle at (x, y) with start cell (sx, sy).
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

// n-th cell of the reference's integer Bresenham, closed form:
// minor(n) = floor((2 n d + D) / (2 D)) with D = major extent, d = minor extent.
__device__ __forceinline__ void ray_cell(int sx, int sy, const Ray &r, int n, int &kx, int &ky)
{
    int dx = r.ex - sx, dy = r.ey - sy;
    int adx = abs(dx), ady = abs(dy);
    if (adx == 0) { kx = sx; ky = sy + n; return; }
    if (ady == 0) { kx = sx + n; ky = sy; return; }
    int xs = dx > 0 ? 1 : -1, ys = dy > 0 ? 1 : -1;
    if (ady > adx) {
        unsigned m = (2u * (unsigned)n * (unsigned)adx + (unsigned)ady) / (2u * (unsigned)ady);
        kx = sx + xs * (int)m;
        ky = sy + ys * n;
    } else {
        unsigned m = (2u * (unsigned)n * (unsigned)ady + (unsigned)adx) / (2u * (unsigned)adx);
        kx = sx + xs * n;
        ky = sy + ys * (int)m;
    }
}

__device__ __forceinline__ bool particle_frame(const RbCtx &c, int p, double &x, double &y, double &cs_, double &sn_,
                                               int &sx, int &sy)
{
    const double *pose = c.pose + 3 * (size_t)p;
    x = pose[0];
    y = pose[1];
    double th = pose[2];
    cs_ = cos(th);
    sn_ = sin(th);
    int tx, ty, ix, iy;
    rb_read_axis(x, tx, ix);
    rb_read_axis(y, ty, iy);
    if (!rb_tile_exists(c, c.exists[p], tx, ty)) return false;    // hybridmap.py:98-100
    sx = rb_trunc(x / RB_CS);                                      // hybridmap.py:102
    sy = rb_trunc(y / RB_CS);
    return true;
}

// ------------------------------------------------------------- prepare --

__global__ void __launch_bounds__(RC_WARPS * 32) raycast_prepare_kernel(RbCtx c)
{
    __shared__ uint32_t mask_s[RC_WARPS][52];                      // 64 tiles * 25 sub-tiles = 1600 bits
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * RC_WARPS + warp;
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
        Ray r = ray_of_beam(c, j, x, y, cs_, sn_, sx, sy);
        for (int n0 = 0; n0 < r.len; n0 += 32) {
            int n1 = min(n0 + 31, r.len - 1), ax, ay, bx, by;
            ray_cell(sx, sy, r, n0, ax, ay);
            ray_cell(sx, sy, r, n1, bx, by);
            int uax = rb_write_axis(c, ax, c.txh), uay = rb_write_axis(c, ay, c.tyh);
            int ubx = rb_write_axis(c, bx, c.txh), uby = rb_write_axis(c, by, c.tyh);
            int sax = uax < 0 ? -1 : uax / RB_SUB, say = uay < 0 ? -1 : uay / RB_SUB;
            int sbx = ubx < 0 ? -1 : ubx / RB_SUB, sby = uby < 0 ? -1 : uby / RB_SUB;
            // when one end is outside the world, cells of the segment inside the
            // world still lie in the row/column of the inside end or up to the
            // world border; mark the border sub-tile of that axis as well.
            if (sax < 0) sax = ax < 0 ? 0 : c.subs_x - 1;
            if (sbx < 0) sbx = bx < 0 ? 0 : c.subs_x - 1;
            if (say < 0) say = ay < 0 ? 0 : c.subs_y - 1;
            if (sby < 0) sby = by < 0 ? 0 : c.subs_y - 1;
            int s0 = say * c.subs_x + sax, s1 = sby * c.subs_x + sbx, s2 = say * c.subs_x + sbx,
                s3 = sby * c.subs_x + sax;
            atomicOr(&mask[s0 >> 5], 1u << (s0 & 31));
            atomicOr(&mask[s1 >> 5], 1u << (s1 & 31));
            atomicOr(&mask[s2 >> 5], 1u << (s2 & 31));
            atomicOr(&mask[s3 >> 5], 1u << (s3 & 31));
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
                    if (old == 1u) atomicExch(&c.refcnt[told], 1u);   // last holder: keep it
                    else action = 2;
                }
                if (action) {
                    int idx = atomicSub(c.free_count, 1) - 1;
                    if (idx < 0) {
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
#pragma unroll 5
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

__global__ void __launch_bounds__(RC_WARPS * 32) raycast_cast_kernel(RbCtx c)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * RC_WARPS + warp;
    if (p >= c.N) return;
    if (c.flags->pool_exhausted) return;                           // prepare could not privatise: skip the scan
    double x, y, cs_, sn_;
    int sx, sy;
    if (!particle_frame(c, p, x, y, cs_, sn_, sx, sy)) return;
    const uint32_t *pt = c.pt + (size_t)p * c.nsub;
    unsigned long long ex_mask = c.exists[p], ex_new = 0ull;
    unsigned long long dropped = 0;
    int cached_sub = -1;
    int8_t *cached_base = nullptr;

    for (int j0 = 0; j0 < c.B; j0 += 32) {
        Ray mine;
        mine.ex = mine.ey = mine.len = mine.occ = 0;
        if (j0 + lane < c.B) mine = ray_of_beam(c, j0 + lane, x, y, cs_, sn_, sx, sy);
        const int nb = min(32, c.B - j0);
        for (int b = 0; b < nb; b++) {
            Ray r;
            r.ex = __shfl_sync(0xffffffffu, mine.ex, b);
            r.ey = __shfl_sync(0xffffffffu, mine.ey, b);
            r.len = __shfl_sync(0xffffffffu, mine.len, b);
            r.occ = __shfl_sync(0xffffffffu, mine.occ, b);
            // reference tile of the end cell, for the "nearby" rule (hybridmap.py:141)
            int uex = rb_write_axis(c, r.ex, c.txh), uey = rb_write_axis(c, r.ey, c.tyh);
            for (int n0 = 0; n0 < r.len; n0 += 32) {
                const int n = n0 + lane;
                const bool active = n < r.len;
                int ops = 0;
                uint32_t id = 0xFFFFFFFFu;
                int8_t *addr = nullptr;
                if (active) {
                    int kx, ky;
                    ray_cell(sx, sy, r, n, kx, ky);
                    int ux = rb_write_axis(c, kx, c.txh), uy = rb_write_axis(c, ky, c.tyh);
                    if (ux < 0 || uy < 0) {
                        dropped++;
                    } else {
                        if (r.occ && n == r.len - 1) ops = 2;                       // hybridmap.py:137-138
                        else ops = 1;                                               // hybridmap.py:144
                        if (r.occ && n == r.len - 2 && uex >= 0 && uey >= 0 && ux / RB_DIM == uex / RB_DIM &&
                            uy / RB_DIM == uey / RB_DIM)
                            ops |= 4;                                               // hybridmap.py:139-142
                        int sub = (uy / RB_SUB) * c.subs_x + ux / RB_SUB;
                        if (sub != cached_sub) {
                            uint32_t t = pt[sub];
                            cached_sub = sub;
                            cached_base = t == RB_NONE ? nullptr : c.pool + (size_t)t * RB_SUB_BYTES;
                        }
                        if (cached_base) {
                            addr = cached_base + (uy % RB_SUB) * RB_SUB + (ux % RB_SUB);
                            id = (uint32_t)uy * (uint32_t)c.ux_max + (uint32_t)ux;
                        } else {
                            ops = 0;                                                // cannot happen after prepare
                            atomicExch(&c.flags->world_overflow, 2);
                        }
                        int tbit = (uy / RB_DIM) * c.tiles_x + ux / RB_DIM;          // HybridMapEntry allocation :125-131
                        if (!((ex_mask >> tbit) & 1ull)) ex_new |= 1ull << tbit;
                    }
                }
                // two consecutive lattice cells can alias to one storage cell
                // (SURVEY 3.4-2): the earlier lane applies both op sets in order.
                uint32_t id_next = __shfl_down_sync(0xffffffffu, id, 1);
                int ops_next = __shfl_down_sync(0xffffffffu, ops, 1);
                uint32_t id_prev = __shfl_up_sync(0xffffffffu, id, 1);
                bool dup_of_prev = lane > 0 && id != 0xFFFFFFFFu && id == id_prev;
                bool absorbs_next = lane < 31 && id != 0xFFFFFFFFu && id == id_next;
                if (ops && !dup_of_prev) {
                    int t = *addr;
                    t = apply_ops(t, ops);
                    if (absorbs_next) t = apply_ops(t, ops_next);
                    *addr = (int8_t)t;
                }
                __syncwarp();
            }
        }
    }
    // publish newly created reference tiles and dropped-cell count
    for (int o = 16; o > 0; o >>= 1) {
        ex_new |= __shfl_xor_sync(0xffffffffu, ex_new, o);
        dropped += __shfl_xor_sync(0xffffffffu, dropped, o);
    }
    if (lane == 0) {
        if (ex_new) c.exists[p] = ex_mask | ex_new;
        if (dropped) atomicAdd(&c.stats->cells_dropped, dropped);
    }
}

void rb_launch_raycast_prepare(const RbCtx &c, cudaStream_t s)
{
    raycast_prepare_kernel<<<(c.N + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, s>>>(c);
}

void rb_launch_raycast_cast(const RbCtx &c, cudaStream_t s)
{
    raycast_cast_kernel<<<(c.N + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, s>>>(c);
}
