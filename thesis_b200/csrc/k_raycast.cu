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
        const Ray r = ray_of_beam(c, j, x, y, cs_, sn_, sx, sy);
        const RayStep st = ray_step(sx, sy, r);
        for (int n0 = 0; n0 < r.len; n0 += 32) {
            int n1 = min(n0 + 31, r.len - 1), ax, ay, bx, by;
            ray_cell(sx, sy, st, n0, ax, ay);
            ray_cell(sx, sy, st, n1, bx, by);
            const uint32_t pax = rb_write_lut(c.lutx, ax, c.txh), pay = rb_write_lut(c.luty, ay, c.tyh);
            const uint32_t pbx = rb_write_lut(c.lutx, bx, c.txh), pby = rb_write_lut(c.luty, by, c.tyh);
            // when one end is outside the world, the cells of the segment that are
            // inside still lie between the inside end and the world border: use the
            // border sub-tile of that axis
            int sax = pax == RB_NONE ? (ax < 0 ? 0 : c.subs_x - 1) : (int)((pax >> 8) & 0xfff);
            int sbx = pbx == RB_NONE ? (bx < 0 ? 0 : c.subs_x - 1) : (int)((pbx >> 8) & 0xfff);
            int say = pay == RB_NONE ? (ay < 0 ? 0 : c.subs_y - 1) : (int)((pay >> 8) & 0xfff);
            int sby = pby == RB_NONE ? (by < 0 ? 0 : c.subs_y - 1) : (int)((pby >> 8) & 0xfff);
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
    const unsigned long long ex_mask = c.exists[p];
    unsigned long long ex_new = 0ull;
    unsigned dropped = 0;
    int cached_sub = -1;
    int8_t *cached_base = nullptr;
    const uint32_t *__restrict__ lutx = c.lutx, *__restrict__ luty = c.luty;
    const int txh = c.txh, tyh = c.tyh, subs_x = c.subs_x, tiles_x = c.tiles_x;

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
            if (r.len == 0) continue;
            const RayStep st = ray_step(sx, sy, r);
            // reference tile of the end cell, for the "nearby" rule (hybridmap.py:141)
            const uint32_t pex = rb_write_lut(lutx, r.ex, txh), pey = rb_write_lut(luty, r.ey, tyh);
            const int end_tile = (pex == RB_NONE || pey == RB_NONE) ? -1 : (int)((pey >> 20) * tiles_x + (pex >> 20));
            const int n_last = r.len - 1, n_near = r.occ ? r.len - 2 : -1, n_occ = r.occ ? n_last : -1;
            for (int n0 = 0; n0 < r.len; n0 += 32) {
                const int n = n0 + lane;
                int ops = 0;
                uint32_t id = 0xFFFFFFFFu;
                int8_t *addr = nullptr;
                if (n <= n_last) {
                    int kx, ky;
                    ray_cell(sx, sy, st, n, kx, ky);
                    const uint32_t px_ = rb_write_lut(lutx, kx, txh), py_ = rb_write_lut(luty, ky, tyh);
                    if (px_ == RB_NONE || py_ == RB_NONE) {
                        dropped++;
                    } else {
                        const int tile = (int)((py_ >> 20) * tiles_x + (px_ >> 20));
                        ops = n == n_occ ? 2 : 1;                                   // hybridmap.py:137-138 / :144
                        if (n == n_near && tile == end_tile) ops |= 4;              // hybridmap.py:139-142
                        const int sub = (int)((py_ >> 8) & 0xfff) * subs_x + (int)((px_ >> 8) & 0xfff);
                        if (sub != cached_sub) {
                            const uint32_t t = pt[sub];
                            cached_sub = sub;
                            cached_base = t == RB_NONE ? nullptr : c.pool + (size_t)t * RB_SUB_BYTES;
                        }
                        if (cached_base) {
                            const uint32_t off = (py_ & 0xff) * RB_SUB + (px_ & 0xff);
                            addr = cached_base + off;
                            id = ((uint32_t)sub << 15) | off;                       // nsub <= 1600, off < 25600
                        } else {
                            ops = 0;                                                // cannot happen after prepare
                            atomicExch(&c.flags->world_overflow, 2);
                        }
                        if (!((ex_mask >> tile) & 1ull)) ex_new |= 1ull << tile;    // HybridMapEntry allocation :125-131
                    }
                }
                // two consecutive lattice cells can alias to one storage cell
                // (SURVEY 3.4-2): the earlier lane applies both op sets in order.
                const uint32_t id_next = __shfl_down_sync(0xffffffffu, id, 1);
                const uint32_t id_prev = __shfl_up_sync(0xffffffffu, id, 1);
                const int ops_next = __shfl_down_sync(0xffffffffu, ops, 1);
                const bool dup_of_prev = lane > 0 && id != 0xFFFFFFFFu && id == id_prev;
                const bool absorbs_next = lane < 31 && id != 0xFFFFFFFFu && id == id_next;
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
        if (dropped) atomicAdd(&c.stats->cells_dropped, (unsigned long long)dropped);
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
