// k_misc.cu -- initialisation, tile export for parity dumps, pool statistics.
#include "common.cuh"

// particles = [Robot(eng) ...] main.py:87 : pose 0, cov 0, weight 1.0
// (robot.py:20-28), one blank reference tile at (0,0) (hybridmap.py:66-70).
__global__ void init_kernel(RbCtx c)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < (size_t)c.N * c.nsub; i += stride) { c.pt[i] = RB_NONE; c.pt2[i] = RB_NONE; }
    for (size_t i = tid; i < (size_t)c.N * 3; i += stride) { c.pose[i] = 0.0; c.pose2[i] = 0.0; }
    for (size_t i = tid; i < (size_t)c.N * 9; i += stride) { c.cov[i] = 0.0; c.cov2[i] = 0.0; }
    const unsigned long long origin = 1ull << (c.tyh * c.tiles_x + c.txh);
    for (size_t i = tid; i < (size_t)c.N; i += stride) {
        c.weight[i] = 1.0;
        c.exists[i] = origin;
        c.exists2[i] = origin;
        c.m_valid[i] = 0;
    }
    for (size_t i = tid; i < c.pool_tiles; i += stride) {
        c.refcnt[i] = 0u;
        c.free_list[i] = (uint32_t)(c.pool_tiles - 1 - i);       // pop order 0, 1, 2, ...
    }
    if (tid == 0) {
        *c.free_count = (int)c.pool_tiles;
        RbStats z = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        *c.stats = z;
        RbFlags f = {};
        *c.flags = f;
    }
}

void rb_launch_init(const RbCtx &c, cudaStream_t s) { init_kernel<<<1024, 256, 0, s>>>(c); }

// One reference tile of one particle as 800x800 float64 [ix][iy] (gridmap.py:32).
__global__ void export_tile_kernel(RbCtx c, int p, int tx, int ty, double *__restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= RB_DIM * RB_DIM) return;
    const int ix = idx / RB_DIM, iy = idx - ix * RB_DIM;
    const int t = rb_cell_tenths(c, p, 800 * (tx + c.txh) + ix, 800 * (ty + c.tyh) + iy);
    out[idx] = (double)t / 10.0;
}

void rb_launch_export_tile(const RbCtx &c, int particle, int tx, int ty, double *out_dev, cudaStream_t s)
{
    export_tile_kernel<<<(RB_DIM * RB_DIM + 255) / 256, 256, 0, s>>>(c, particle, tx, ty, out_dev);
}

// out[0] = page-table entries whose sub-tile is shared, out[1] = allocated entries,
// out[2] = sum of all reference counts (== out[1] when the pool accounting is intact).
__global__ void refstats_kernel(RbCtx c, unsigned long long *out)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long shared = 0, total = 0;
    for (size_t i = tid; i < (size_t)c.N * c.nsub; i += stride) {
        uint32_t t = c.pt[i];
        if (t == RB_NONE) continue;
        total++;
        if (c.refcnt[t] > 1u) shared++;
    }
    for (int o = 16; o > 0; o >>= 1) {
        shared += __shfl_xor_sync(0xffffffffu, shared, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    unsigned long long rsum = 0;
    for (size_t i = tid; i < c.pool_tiles; i += stride) rsum += c.refcnt[i];
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], shared); atomicAdd(&out[1], total); atomicAdd(&out[2], rsum); }
}

void rb_launch_refstats(const RbCtx &c, unsigned long long *out2_dev, cudaStream_t s)
{
    cudaMemsetAsync(out2_dev, 0, 3 * sizeof(unsigned long long), s);
    refstats_kernel<<<512, 256, 0, s>>>(c, out2_dev);
}

// Occupied cells (log-odds > 1.0) of one particle as the reference's plotting feed
// HybridMap.get_occupied_points (hybridmap.py:303-313): coordinates in cell units,
// ((i - 400) * 0.05 + centre) / 0.05.  One CTA per sub-tile; unordered append.
__global__ void __launch_bounds__(256) occupied_points_kernel(RbCtx c, int p, double *__restrict__ out_xy,
                                                            unsigned long long cap, unsigned long long *count)
{
    const int sub = blockIdx.x;
    const uint32_t t = c.pt[(size_t)p * c.nsub + sub];
    if (t == RB_NONE) return;
    const int sx = sub % c.subs_x, sy = sub / c.subs_x;
    const int8_t *tile = c.pool + (size_t)t * RB_SUB_BYTES;
    for (int idx = threadIdx.x; idx < RB_SUB_BYTES; idx += blockDim.x) {
        const int x = idx % RB_SUB, y = idx / RB_SUB;
        if (tile[RB_OFF_Y(y) + RB_OFF_X(x)] <= RB_T_OCC_THRESH) continue;
        const unsigned long long slot = atomicAdd(count, 1ull);
        if (!out_xy || slot >= cap) continue;
        const int ux = sx * RB_SUB + x, uy = sy * RB_SUB + y;
        const int tx = ux / RB_DIM - c.txh, ty = uy / RB_DIM - c.tyh;
        const int ix = ux % RB_DIM, iy = uy % RB_DIM;
        out_xy[2 * slot] = (((double)ix - 400.0) * RB_CS + 40.0 * tx) / RB_CS;
        out_xy[2 * slot + 1] = (((double)iy - 400.0) * RB_CS + 40.0 * ty) / RB_CS;
    }
}

void rb_launch_occupied_points(const RbCtx &c, int particle, double *out_dev, unsigned long long cap,
                               unsigned long long *count_dev, cudaStream_t s)
{
    cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), s);
    occupied_points_kernel<<<c.nsub, 256, 0, s>>>(c, particle, out_dev, cap, count_dev);
}
