// common.cuh -- shared constants, device context and device helpers of the
// B200-native RBPF update.  Compiled with -fmad=false: every a*b+c below is two
// IEEE roundings unless written as fma(), because cell indices must replay the
// reference's float64 expressions exactly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rb_math.h"                       // sin / cos / exp shared bit for bit with the CPU oracle

// ---- constants of the reference (file:line into the reference checkout) ----
#define RB_CS 0.05              // hybridmap.py:67  cell size [m]
#define RB_TILE_LEN 40.0        // hybridmap.py:68  reference tile side [m]
#define RB_DIM 800              // gridmap.py:31    cells per reference tile side
#define RB_T_OCC 8              // gridmap.py:20    +0.80 in tenths
#define RB_T_NEAR 2             // gridmap.py:21    +0.20
#define RB_T_EMP 3              // gridmap.py:23    -0.30 (magnitude)
#define RB_T_MAX 30             // gridmap.py:22,24 caps +-3.0
#define RB_T_OCC_THRESH 10      // gridmap.py:17    occupied iff L > 1.0
#define RB_MATCH_MAX_R 11.0     // hybridmap.py:20
#define RB_MATCH_MIN_R 1e-3     // hybridmap.py:218
#define RB_CLIP_R 15.0          // hybridmap.py:107-108
#define RB_W_MAX_R 25.0         // robot.py:130
#define RB_W_MIN_R 0.01         // robot.py:130
#define RB_RESAMPLE_TRIGGER 200.0 // main.py:50
#define RB_NT_MAX 14            // 0.7 m window clamp (robot.py:64-65) in cells

// ---- device layout ----
#define RB_SUB 160                          // cells per sub-tile side (divides 800)
#define RB_SUB_BYTES (RB_SUB * RB_SUB)      // int8 tenths
// Cells of a sub-tile are stored in blocks of 8 (x) by 4 (y) = one 32-byte sector,
// 20 blocks per block row: a ray of 32 consecutive cells touches ~5 (along x) to
// ~9 (along y) sectors instead of 1 to 32 with plain rows.
#define RB_BLK_X 8
#define RB_BLK_Y 4
#define RB_BLKS_PER_ROW (RB_SUB / RB_BLK_X)  // 20
// the two per-axis parts of a cell's byte offset add up to the offset
#define RB_OFF_X(x) ((((x) >> 3) << 5) + ((x) & 7))
#define RB_OFF_Y(y) (((y) >> 2) * (RB_BLKS_PER_ROW * 32) + (((y) & 3) << 3))
// packed write-LUT entry: bits 0-14 offset part, 15-23 sub-tile index along the axis,
// 24-29 reference-tile index along the axis, bit 30 / 31 = shares its storage cell with k+1 / k-1
#define RB_LUT_OFF(p) ((p) & 0x7fffu)
#define RB_LUT_SUB(p) (((p) >> 15) & 0x1ffu)
#define RB_LUT_TILE(p) (((p) >> 24) & 0x3fu)
#define RB_LUT_NEXT_BIT 30
#define RB_LUT_PREV_BIT 31
#define RB_SUBS_PER_TILE (RB_DIM / RB_SUB)  // 5
#define RB_NONE 0xFFFFFFFFu                 // unallocated page-table entry (reads as log-odds 0)
#define RB_MAXB 384                         // max beams per sweep
#define RB_MAXK 32                          // max proposal samples (one lane each)

// matcher window (see k_match.cu)
#define RB_WIN_R 237
#define RB_BM_ROWS (2 * RB_WIN_R + 1)       // 475
#define RB_BM_STRIDE 17                     // 16 data words + 1 (funnel-shift hi), odd => conflict-free rows
#define RB_RAW_ROWS (RB_BM_ROWS + 2)
#define RB_RAW_STRIDE 18
#define RB_SLICE_W (2 * RB_NT_MAX + 1)      // 29

struct RbStats {                            // device-side counters
    unsigned long long cow_copies, fresh_allocs, cells_dropped, resamples, match_failed, match_evals, match_visits, match_points, match_runs,
        ndt_evals, ndt_accepted;
    // match_kernel phase split (SM clocks summed over CTAs resp. warps, see rbpf_match_phase_clocks)
    unsigned long long match_clk[16];
    unsigned long long match_failed_zero;   // failed matches whose optimum is the zero correction (matchScanCustom.m:55)
};

struct RbFlags {                            // device-side status words
    int pool_exhausted;                     // set by raycast prepare
    int resample_error;                     // main.py:66-67 assertion would have fired
    int did_resample;                       // last resample triggered
    int world_overflow;                     // matcher window left the world / internal bound hit
    int remote_needed;                      // multi-GPU: local slots whose ancestor is remote
    int resample_error_sticky;              // resample_error of any plan since the last rbpf_clear_errors
    int pad[2];
};

struct RbPeer {                             // another rank's current particle state, mapped into this process
    const int8_t *pool;
    const uint32_t *pt;
    const double *pose, *cov;
    const unsigned long long *exists;
};

#define RB_MAX_WORLD 16
struct RbPeers { RbPeer p[RB_MAX_WORLD]; };  // by rank; this rank's own entry is unused

struct RbCtx {
    int N, B, K;                            // local particles, beams, samples
    int rank, world, n_global;              // sharding
    int tiles_x, tiles_y, txh, tyh;         // world extent in reference tiles, half extents
    int subs_x, subs_y, nsub;               // sub-tile grid of the world
    int ux_max, uy_max;                     // storage extent in cells
    uint32_t pool_tiles;
    // tile pool
    int8_t *pool;
    uint32_t *refcnt;
    uint32_t *free_list;
    int *free_count;
    // particle state (current buffers)
    double *pose, *cov, *weight;
    uint32_t *pt;                           // N * nsub page table
    unsigned long long *exists;             // N  bitmask of existing reference tiles
    // alternate buffers (resample target)
    double *pose2, *cov2;
    uint32_t *pt2;
    unsigned long long *exists2;
    // scan
    const double *px, *py, *dist;
    const float4 *beamf;                    // (px, py, 1 if RB_W_MIN_R < dist < RB_W_MAX_R else 0, 0) in float32 (weight stage)
    // matcher
    const double *rot_cs;                   // (2*nk+1) * 2 : cos, sin of k*step
    int nk;
    double rot_step;
    double *m_pose, *m_cov, *m_score;
    int *m_valid, *m_best, *m_refine;       // m_refine: {NDT evaluations, refined pose accepted} per particle
    // scan-to-previous-scan matching (hybridmap.py:147-191): global endpoints of the previous scan
    const double *prev_x, *prev_y;
    int n_prev;
    // write-path LUT (lattice cell k -> storage coordinate)
    const uint32_t *lutx, *luty;            // per axis, packed (RB_LUT_*)
    const uint32_t *clut;                   // cast LUT (k_raycast.cu, version 2): x entries then y entries, packed so that x + y is
                                            // offset (0-14) | page-table slot (15-25) | aliasing flags (x: 30/31, y: 28/29)
    const uint32_t *rlut;                   // read LUT: storage coordinate -> page-table slot (bits 0-11) | byte offset in the sub-tile (12-31),
                                            // ux_max x entries then uy_max y entries; the entries of the two axes add up (k_weight.cu)
    int *cast_work;                         // work counter of the persistent cast kernel (zeroed by raycast_prepare)
    unsigned char *pulled;                  // N: 1 = the particle arrived with the last sharded resample (its sub-tiles may still be in flight)
    int *cast_done;                         // N: parts of a particle's sweep already cast (split mode of the cast kernel, zeroed per launch)
    int cast_ipp;                           // work items per particle (1: whole sweeps; set by the launcher)
    // resample
    double *w_all;                          // n_global adjusted weights / cumsum scratch
    double *plan_scal;                      // slice, start of the last plan (main.py:57,59)
    int *ancestors;                         // n_global
    int *mult;                              // N  local descendants of each old local particle
    // Duplicates made by the last resample are bit-identical (pose, covariance, shared
    // page table) until the next weight stage: dup_of[j] = first local slot with the same
    // ancestor.  The matcher runs once per representative and its result is copied.
    int *dup_of;                            // N
    int refine;                             // host-set: run the NDT stage (matchScanCustom.m:32-50) after the grid search
    int use_dup;                            // host-set: dup_of is valid for this launch
    RbStats *stats;
    RbFlags *flags;
    unsigned long long seed;
    unsigned long long step_no;
};

// ---------------------------------------------------------------- helpers --

// Python int(): truncation toward zero of a float64.
__device__ __forceinline__ int rb_trunc(double v) { return __double2int_rz(v); }

// Scan.from_global_reference lidar.py:111-128 with np.matmul's accumulation
// order (round(c*px), fma(-s,py,.), + x), see oracle/rbpf_oracle.c xform().
__device__ __forceinline__ void rb_xform(double c, double s, double x, double y, double px, double py,
                                         double &gx, double &gy)
{
    gx = fma(-s, py, c * px) + x;
    gy = fma(c, py, s * px) + y;
}

// Read path, one axis: HybridMapEntry.is_in_map hybridmap.py:44-45 on the 40 m
// lattice followed by GridMap.get_cell gridmap.py:120-128.  Returns the tile
// lattice index in t and the tile-local index in idx.
//
// The reference's index expression int(rel/40*800 + 400) needs an IEEE divide.
// rel*20 + 400 differs from it by < 1e-12, so its truncation is the same unless
// the value sits within 1e-6 of an integer; only then is the exact expression
// replayed.  The tile candidate uses a multiply and is fixed up by exact compares.
__device__ __forceinline__ void rb_read_axis(double g, int &t, int &idx)
{
    t = __double2int_rd(g * 0.025 + 0.5);
    double c = RB_TILE_LEN * (double)t;
    if (g < c - 20.0) { t--; c -= RB_TILE_LEN; }
    else if (g >= c + 20.0) { t++; c += RB_TILE_LEN; }
    const double rel = g - c;
    double v = rel * 20.0 + 400.0;
    if (fabs(v - rint(v)) < 1e-6) v = rel / RB_TILE_LEN * 800.0 + 400.0;
    idx = rb_trunc(v);
}

// GridMap.index_to_distance gridmap.py:333-334 plus the tile centre.
__device__ __forceinline__ double rb_cell_corner(int idx, int t)
{
    return ((double)idx - 400.0) * RB_TILE_LEN / 800.0 + RB_TILE_LEN * (double)t;
}

// int8 log-odds (tenths) at storage cell (ux, uy) of particle p; 0 when the
// sub-tile is unallocated or the cell is outside the world (reference: None -> 0).
__device__ __forceinline__ int rb_cell_tenths(const RbCtx &c, int p, int ux, int uy)
{
    if ((unsigned)ux >= (unsigned)c.ux_max || (unsigned)uy >= (unsigned)c.uy_max) return 0;
    int sub = (uy / RB_SUB) * c.subs_x + ux / RB_SUB;
    uint32_t t = c.pt[(size_t)p * c.nsub + sub];
    if (t == RB_NONE) return 0;
    return c.pool[(size_t)t * RB_SUB_BYTES + RB_OFF_Y(uy % RB_SUB) + RB_OFF_X(ux % RB_SUB)];
}

// Read path split for memory-level parallelism: page-table slot and byte offset
// of the cell under (gx, gy); sub = -1 outside the world.
__device__ __forceinline__ void rb_locate(const RbCtx &c, double gx, double gy, int &sub, int &off)
{
    int tx, ty, ix, iy;
    rb_read_axis(gx, tx, ix);
    rb_read_axis(gy, ty, iy);
    if (tx < -c.txh || tx > c.txh || ty < -c.tyh || ty > c.tyh) { sub = -1; off = 0; return; }
    const int ux = 800 * (tx + c.txh) + ix, uy = 800 * (ty + c.tyh) + iy;
    sub = (uy / RB_SUB) * c.subs_x + ux / RB_SUB;
    off = RB_OFF_Y(uy % RB_SUB) + RB_OFF_X(ux % RB_SUB);
}

// The same location with a third of the float64 work, for the weight stage's K x B lookups.  With
// v = g * 20 (cells), tile containment g in [40 t - 20, 40 t + 20) and the index int((g - 40 t)/40*800 + 400)
// are floor(v) split by integer arithmetic -- unless v lies within 1e-6 of an integer, where the last
// ulp of the reference's own expression decides: those coordinates take rb_read_axis.
__device__ __forceinline__ void rb_locate_fast(const RbCtx &c, double gx, double gy, int &sub, int &off)
{
    const double vx = gx * 20.0, vy = gy * 20.0;
    const int kx = __double2int_rd(vx), ky = __double2int_rd(vy);
    const double dx = vx - (double)kx, dy = vy - (double)ky;
    if (!(dx > 1e-6 && dx < 1.0 - 1e-6 && dy > 1e-6 && dy < 1.0 - 1e-6 && fabs(vx) < 1e6 && fabs(vy) < 1e6)) {
        rb_locate(c, gx, gy, sub, off);
        return;
    }
    // storage coordinate: 800 (t + half) + idx = k + 400 + 800 half
    const int ux = kx + 400 + 800 * c.txh, uy = ky + 400 + 800 * c.tyh;
    if ((unsigned)ux >= (unsigned)c.ux_max || (unsigned)uy >= (unsigned)c.uy_max) { sub = -1; off = 0; return; }
    sub = (uy / RB_SUB) * c.subs_x + ux / RB_SUB;
    off = RB_OFF_Y(uy % RB_SUB) + RB_OFF_X(ux % RB_SUB);
}

// HybridMap.get_odds_at hybridmap.py:85-93 in tenths (None -> 0).
__device__ __forceinline__ int rb_odds_tenths(const RbCtx &c, int p, double gx, double gy)
{
    int tx, ty, ix, iy;
    rb_read_axis(gx, tx, ix);
    rb_read_axis(gy, ty, iy);
    if (tx < -c.txh || tx > c.txh || ty < -c.tyh || ty > c.tyh) return 0;
    return rb_cell_tenths(c, p, 800 * (tx + c.txh) + ix, 800 * (ty + c.tyh) + iy);
}

__device__ __forceinline__ bool rb_tile_exists(const RbCtx &c, unsigned long long mask, int tx, int ty)
{
    if (tx < -c.txh || tx > c.txh || ty < -c.tyh || ty > c.tyh) return false;
    return (mask >> ((ty + c.tyh) * c.tiles_x + (tx + c.txh))) & 1ull;
}

// Write path, one axis: lattice cell k -> packed storage location through the
// LUT built on the host from int((k*0.05 - c)/0.05 + 400.0) (gridmap.py:92-95),
// packed as RB_LUT_*.  RB_NONE outside the world.
__device__ __forceinline__ uint32_t rb_write_lut(const uint32_t *__restrict__ lut, int k, int half_tiles)
{
    const unsigned q = (unsigned)(k + 800 * half_tiles + 400);
    if (q >= (unsigned)(800 * (2 * half_tiles + 1))) return RB_NONE;
    return __ldg(&lut[q]);
}

// Philox4x32-10 (counter-based RNG) for device-side draws.
__device__ __forceinline__ void rb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, unsigned long long key,
                                          uint32_t out[4])
{
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double rb_u01(uint32_t hi, uint32_t lo)   // (0,1), 53 bits
{
    unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// ---- launchers (one per kernel file) ----
void rb_launch_motion(const RbCtx &c, int family, const double *u, double dt, const double *par, cudaStream_t s);
void rb_launch_match(const RbCtx &c, int adj, cudaStream_t s, int sel = 0, bool copy_dups = true);   // sel 1: all but pulled particles, 2: only those
void rb_launch_match_slice(const RbCtx &c, int particle, int *slice_dev, int adj, cudaStream_t s);
void rb_launch_weight(const RbCtx &c, const double *z_dev, const double *guesses_dev, int fallback_phase, cudaStream_t s);
void rb_launch_raycast_prepare(const RbCtx &c, cudaStream_t s);
void rb_launch_raycast_cast(const RbCtx &c, cudaStream_t s);
void rb_launch_resample(const RbCtx &c, const double *weights_all, const double *u01_dev, cudaStream_t s);
void rb_launch_resample_apply(const RbCtx &c, cudaStream_t s);
void rb_launch_resample_gather(const RbCtx &c, cudaStream_t s);
void rb_launch_resample_refs(const RbCtx &c, cudaStream_t s);
void rb_launch_export_tile(const RbCtx &c, int particle, int tx, int ty, double *out_dev, cudaStream_t s);
void rb_launch_init(const RbCtx &c, cudaStream_t s);
void rb_launch_refstats(const RbCtx &c, unsigned long long *out2_dev, cudaStream_t s);
size_t rb_match_smem_bytes();
void rb_launch_occupied_points(const RbCtx &c, int particle, double *out_dev, unsigned long long cap,
                               unsigned long long *count_dev, cudaStream_t s);
void rb_launch_migrate_claim(const RbCtx &c, const int *slots_dev, int n, uint32_t *mark, uint32_t *list, int *count,
                             cudaStream_t s);
void rb_launch_migrate_pack(const RbCtx &c, const int *slots_dev, int n, int n_tiles, uint32_t *mark, uint32_t *list,
                            int *count, unsigned char *buf, cudaStream_t s);
void rb_launch_migrate_unpack(const RbCtx &c, const unsigned char *buf, int n, int n_tiles, const int *dst_slots_dev,
                              const int *rec_idx_dev, int m, uint32_t *map, cudaStream_t s);
size_t rb_migrate_bytes(int n, int n_tiles, int nsub);
void rb_launch_migrate_pull(const RbCtx &c, const RbPeers &peers, uint32_t *mark, uint32_t *list, unsigned char *list_rank,
                            uint32_t *list_local, int *count, cudaStream_t s, cudaStream_t copy_stream, cudaEvent_t ev_alloc);
