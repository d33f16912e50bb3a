// k_match.cu -- stage 2: per-particle correlative scan-to-map matching.
//
// Reference contract: HybridMap.get_scan_match hybridmap.py:210-261 (front-end:
// which beams become "curr" points, snapped to cell corners, relative to the
// guess) and matchScanCustom.m:1-58 (search window, isValidPose gate, NaN
// covariance on failure).  The search arithmetic itself lives in MathWorks'
// matchScansGrid, which is not in the reference tree: the algorithm below is the
// restatement defined in oracle/rbpf_oracle.c orc_match() -- PARITY UNPINNED
// against MATLAB, bit-exact against the oracle.
//
// One CTA per particle.
//   1. curr points (<= B) -> shared memory.
//   2. the particle's occupancy around the guess cell is gathered from the tile
//      pool with 32-byte vector loads (one aligned 32-cell word per thread),
//      thresholded (tenths > 10) into a bitmap, 3x3-dilated (proximity kernel):
//      475 rows x 512 bits in shared memory.
//   3. every warp takes rotations k = warp, warp + W, ...; lanes first rotate and
//      rasterise the points of that rotation, then each lane owns one row shift
//      j of the translation window: for every point it funnel-shifts the 29 bits
//      [x - nx, x + nx] of bitmap row (y + j) out of two shared-memory words
//      (row stride 17 words -> the 29 lanes hit 29 different banks) and adds
//      them into bit-sliced counters with carry-save adders (LOP3).  One pass
//      over the points therefore scores 29 x 29 translations.
//   4. bit-sliced max + tie-break key per lane, warp-shuffle argmax, CTA argmax.
//   5. covariance from the score slice at the best rotation and the score line
//      at the best translation (integer moments, weights 2^(score - best)).
//   5b. optional NDT refinement of the grid optimum (matchScanCustom.m:32-50) on the
//      same bitmap, thread = beam: ndt_partial / ndt_refine below, DESIGN.md 3.1-8.
//
// Branch and bound over rotations (exact): rotations are grouped by MT_GROUP
// consecutive lattice steps.  A point at range <= 11 m moves at most one cell per
// rotation step, so scoring the group's middle rotation against the bitmap
// dilated by MT_GRAD cells bounds the score of every rotation of the group at
// every translation from above (mt_bound_pass: four row shifts per lane, four points
// per instruction).  The rotations around the guess are scored first
// (seeds), phase A then scores every group once (231/8 = 29 passes instead of
// 231), phase B visits groups in decreasing bound and scores a member rotation
// only while its bound can still beat the best key so far.  Every pass gives up
// as soon as its best partial count plus the points still to come cannot reach
// the best complete score (admissible).  The result is identical to the
// exhaustive search of the oracle; ~31 full-pass equivalents instead of 231.
#include "common.cuh"

#ifdef WT_LIBM   // experiment only
#define rb_sincos(a, s, c) sincos(a, s, c)
#endif

#ifndef MT_GROUP
#define MT_GROUP 8                      // rotations per group
#endif
#define MT_GRAD (MT_GROUP / 2 + 1)      // dilation radius: MT_GROUP/2 cells of motion + 1 cell of rounding
#define MT_MAXGROUPS 64
#define MT_ROT_LINE_HALF 16             // rotation-variance support (oracle ROT_LINE_HALF)
#ifndef MT_WARPS
#define MT_WARPS 12
#endif
#ifndef MT_SEEDS
#define MT_SEEDS 12                     // rotations around the guess scored first (by as many warps)
#endif
#ifndef MT_ABORT_EVERY
#define MT_ABORT_EVERY 64              // points between two abort checks of a scoring pass (power of two, multiple of 8)
#endif
#define MT_THREADS (MT_WARPS * 32)
#define MT_PLANES 9                     // bit-sliced counters up to 511 >= RB_MAXB
// staging buffer: the 477 map rows the 3x3 dilation reads, starting on a multiple of 4 in
// storage coordinates (a 32-byte sector of the pool holds 8 x 4 cells) -> up to 3 rows more
#define MT_RAW_ROWS 484
#define MT_NRB ((MT_RAW_ROWS + 31) / 32) // 16 bands of 32 rows
#define MT_NBLK (MT_NRB * RB_RAW_STRIDE) // 288 blocks of 32 rows x 32 cells
#ifndef MT_GATHER_U
#define MT_GATHER_U 4                   // blocks (= 32-byte sectors per lane) in flight per warp and round of the gather
#endif
#define MT_SEG_ROWS 20                  // rows per thread of the fused dilation pass (24 segments x 16 words)

// Phase clocks (include/rbpf_b200.h rbpf_match_phase_clocks): thread 0 of a CTA adds the SM
// clocks between two barriers to slot i; MT_WCLK adds a warp's own busy time.
#define MT_CLK(i)                                                                         \
    if (tid == 0 && !slice_out) {                                                         \
        const long long t_ = clock64();                                                   \
        atomicAdd(&c.stats->match_clk[i], (unsigned long long)(t_ - clk_prev));           \
        clk_prev = t_;                                                                    \
    }
#define MT_WCLK(i, t0)                                                                    \
    if (lane == 0 && !slice_out) atomicAdd(&c.stats->match_clk[i], (unsigned long long)(clock64() - (t0)));

struct MatchShared {
    double gx, gy, gth, cs0, sn0, fx, fy, rx, ry;
    int M, nx, ny, g0xu, g0yu, x0, y0, t0x, t0y, ok, overflow;
    int slot_warp;                      // warp whose slot holds the counters of the best rotation
    int a0, dy0;                        // first staged row (storage coordinates, multiple of 4); y0 - 1 - a0
    int nblk;                           // blocks to gather
    unsigned short row_lo[RB_BM_ROWS], row_hi[RB_BM_ROWS];   // columns of a bitmap row within 11.5 m of the guess (:239)
    uint32_t need[MT_NRB];              // per band: which 32-cell words a lookup can reach
    unsigned short blist[MT_NBLK];      // compacted list of those blocks
    int ord_cnt[4 * MT_WARPS];          // far-first ordering of the points: per range class and warp
    uint32_t ptw[25];                   // page-table entries of the <= 5 x 5 sub-tiles under the staging window
    int sxb0, syb0;                     // sub-tile coordinates of ptw[0] (may be negative at the world border)
    unsigned long long best_key;
    long long mom[9];                   // W0 Wx Wy Wxx Wyy Wxy T0 T1 T2
    long long mom_part[MT_WARPS][9];    // per-warp partial moments (64-bit shared-memory atomics are CAS loops)
    int group_ub[MT_MAXGROUPS];         // phase A: upper bound of every rotation group
    int group_order[MT_MAXGROUPS];      // groups by decreasing bound
    int next_group;                     // phase A work queue
    int next_item;                      // phase B work queue
    int evals;                          // scoring passes started (statistics)
    int visits;                         // points visited by those passes (an aborted pass visits fewer than M)
};

__host__ __device__ inline size_t mt_bm_words() { return ((size_t)RB_BM_ROWS * RB_BM_STRIDE + 3) & ~(size_t)3; }  // keeps raw/pts 16-B aligned
__host__ __device__ inline size_t mt_raw_words() { return (size_t)MT_RAW_ROWS * RB_RAW_STRIDE; }

size_t rb_match_smem_bytes()
{
    size_t words = 2 * mt_bm_words() + mt_raw_words();
    words = (words + 1) & ~(size_t)1;
    return words * 4 + 2 * RB_MAXB * sizeof(double) + RB_MAXB * sizeof(float2) + sizeof(MatchShared);
}

__device__ __forceinline__ unsigned long long mt_key(int score, int i, int j, int k)
{
    // higher score; then smaller |k|, |i|, |j|; then negative before positive (oracle match_key)
    unsigned long long key = (unsigned long long)(unsigned)score << 32;
    key |= (unsigned long long)(255 - abs(k)) << 24;
    key |= (unsigned long long)(31 - abs(i)) << 19;
    key |= (unsigned long long)(31 - abs(j)) << 14;
    key |= (unsigned long long)(k < 0) << 13;
    key |= (unsigned long long)(i < 0) << 12;
    key |= (unsigned long long)(j < 0) << 11;
    return key;
}

__device__ __forceinline__ void mt_key_decode(unsigned long long key, int &score, int &i, int &j, int &k)
{
    score = (int)(key >> 32);
    int ak = 255 - (int)((key >> 24) & 0xff), ai = 31 - (int)((key >> 19) & 0x1f), aj = 31 - (int)((key >> 14) & 0x1f);
    k = ((key >> 13) & 1) ? -ak : ak;
    i = ((key >> 12) & 1) ? -ai : ai;
    j = ((key >> 11) & 1) ? -aj : aj;
}

// carry-save adder: (h, l) = a + b + c per bit
#define CSA(h, l, a, b, c)                          \
    {                                               \
        uint32_t u_ = (a) ^ (b);                    \
        h = ((a) & (b)) | (u_ & (c));               \
        l = u_ ^ (c);                               \
    }

// 32 thresholded cells -> 32 bits (bit b = cell x0 + b)
__device__ __forceinline__ uint32_t mt_pack4(uint32_t bytes)
{
    uint32_t m = __vcmpgts4(bytes, 0x0A0A0A0Au) & 0x80808080u;              // tenths > 10  (gridmap.py:153)
    return (m * 0x00204081u) >> 28;
}

// Packed point: (byte offset of the bitmap word << 5) | bit shift, for row shift 0.  The
// scoring loop turns it into a shared-memory address with one multiply-add-high on the
// FMA pipe (the ALU pipe is the busy one: LOP3 carry-save adders and funnel shifts).
#define MT_PT_WORD_SHIFT 7
__device__ __forceinline__ uint32_t mt_lds(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t mt_lds4(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+4];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t mt_pt_addr(uint32_t pk, uint32_t lane_base)
{
    uint32_t a;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(pk), "r"(1u << 27), "r"(lane_base));   // (pk >> 5) + lane_base
    return a;
}

__device__ __forceinline__ int mt_decode(const uint32_t pl[MT_PLANES], int bit)
{
    int s = 0;
#pragma unroll
    for (int p = 0; p < MT_PLANES; p++) s |= (int)((pl[p] >> bit) & 1u) << p;
    return s;
}

// Best key of one lane's row: bit-sliced max over the translation bits, then the
// column closest to the window centre (negative side first), oracle match_key order.
__device__ __forceinline__ unsigned long long mt_lane_key(const uint32_t pl[MT_PLANES], uint32_t colmask, int nx, int j, int k)
{
    uint32_t cand = colmask;
    int sc = 0;
#pragma unroll
    for (int pbit = MT_PLANES - 1; pbit >= 0; pbit--) {
        uint32_t t = cand & pl[pbit];
        if (t) { cand = t; sc |= 1 << pbit; }
    }
    uint32_t lowm = cand & ((2u << nx) - 1u);                                // bits 0..nx   (i <= 0)
    uint32_t highm = cand >> nx;                                             // bit d = i = +d
    int dn = lowm ? nx - (31 - __clz(lowm)) : 99;
    int dp = highm ? __ffs(highm) - 1 : 99;
    int i = dn <= dp ? -dn : dp;
    return mt_key(sc, i, j, k);
}

// Rasterise the curr points for rotation k into packed bitmap addresses.
//
// The lattice point is floor(v) of a float64 expression (same operations as the oracle).  A
// float32 evaluation of v (points pre-scaled to cells, fused multiply-adds) is within 5e-5
// of it (|c| <= 220 cells, |v| < 256: the two points' conversions 7.6e-6 and 3.8e-6 (|sin| <= 0.5), the
// conversions of cos / sin 1.3e-5 and 6.6e-6, two roundings of 7.6e-6), so its floor is the same
// unless v32 lies within MT_RAS_EPS of an integer -- only those points (0.05 %) take the float64
// expression.  (2e-3 until round 2: one warp iteration in eight ran both paths.)
#ifndef MT_RAS_EPS
#define MT_RAS_EPS 2.5e-4f
#endif
struct MtRot {                                                             // one rotation of the search, per lane
    float ckf, skf, fxf, fyf;
    int k, xoff, yoff;
};

__device__ __forceinline__ MtRot mt_rot(const RbCtx &c, const MatchShared *sh, int k, int shift_i, int shift_j)
{
    MtRot r;
    r.k = k;
    r.ckf = (float)c.rot_cs[2 * (k + c.nk)];
    r.skf = (float)c.rot_cs[2 * (k + c.nk) + 1];
    r.fxf = (float)(sh->fx * 20.0 + 0.5);
    r.fyf = (float)(sh->fy * 20.0 + 0.5);
    r.xoff = sh->g0xu - sh->x0 + shift_i;
    r.yoff = RB_WIN_R + shift_j;
    return r;
}

// the oracle's float64 expression, for the few points whose float32 value lies next to a lattice boundary
__device__ __noinline__ int2 mt_raster_exact(const double *__restrict__ cs, double cx, double cy, double fx, double fy)
{
    const double ck = cs[0], sk = cs[1];
    const double rxq = (ck * cx - sk * cy) + fx;
    const double ryq = (sk * cx + ck * cy) + fy;
    return make_int2(__double2int_rd(rxq * 20.0 + 0.5), __double2int_rd(ryq * 20.0 + 0.5));   // nearest lattice point (see oracle)
}

__device__ __forceinline__ uint32_t mt_raster_point(const RbCtx &c, MatchShared *sh, const double *ccx, const double *ccy,
                                                    const float2 *__restrict__ ccf, const MtRot &r, int q, int span_i, int span_j, int &ovf)
{
    const float2 cf = ccf[q];
    const float vx = fmaf(r.ckf, cf.x, fmaf(-r.skf, cf.y, r.fxf)), vy = fmaf(r.skf, cf.x, fmaf(r.ckf, cf.y, r.fyf));
    const float flx = floorf(vx), fly = floorf(vy);
    int ox = (int)flx, oy = (int)fly;
    const float dx = vx - flx, dy = vy - fly;
    // out of line: if-converted, the float64 path would be issued (predicated off) for every point of every pass
    if (!(dx > MT_RAS_EPS && dx < 1.0f - MT_RAS_EPS && dy > MT_RAS_EPS && dy < 1.0f - MT_RAS_EPS)) {
        const int2 o = mt_raster_exact(c.rot_cs + 2 * (r.k + c.nk), ccx[q], ccy[q], sh->fx, sh->fy);
        ox = o.x; oy = o.y;
    }
    int bx = ox + r.xoff, by = oy + r.yoff;
    // cannot happen for |c| < 11 m; branch-free (the flag travels in a register: the if-converted store to shared memory
    // was seven instructions issued, predicated off, for every point)
    const bool bad = (unsigned)bx >= (unsigned)(32 * (RB_BM_STRIDE - 1) - span_i) || (unsigned)by >= (unsigned)(RB_BM_ROWS - span_j);
    ovf |= (int)bad;
    bx = bad ? 0 : bx;
    by = bad ? 0 : by;
    return ((uint32_t)(by * RB_BM_STRIDE + (bx >> 5)) << MT_PT_WORD_SHIFT) | (uint32_t)(bx & 31);
}

// One scoring pass: the points of rotation k are rasterised MT_CHUNK at a time, just before
// they are scored, so a pass that is given up has rasterised only what it looked at.  Hit
// masks go into bit-sliced counters (carry-save adders over 16 points, then a ripple into
// the upper planes).
//
// ABORT: after every chunk the partial counts are compared with the best complete score
// found so far (best_key, shared): when no translation can reach it even if all remaining
// points hit, the pass is abandoned (returns false).  The bound is admissible, so the search
// result is unchanged.
#ifndef MT_CHUNK
#define MT_CHUNK 64                     // multiple of 32
#endif
template <bool ABORT>
__device__ __forceinline__ bool mt_pass(const RbCtx &c, MatchShared *sh, const double *ccx, const double *ccy,
                                        const float2 *__restrict__ ccf, int k, int nx, int ny, uint32_t bm_lane, uint32_t *pts,
                                        int M, int lane, uint32_t pl[MT_PLANES], uint32_t rowmask = 0u,
                                        const volatile unsigned long long *best_key = nullptr, int *visited = nullptr)
{
    const MtRot r = mt_rot(c, sh, k, -nx, -ny);
    int ovf = 0;
    uint32_t ones = 0, twos = 0, fours = 0, eights = 0;
#pragma unroll
    for (int p = 0; p < MT_PLANES; p++) pl[p] = 0;
#define MT_LOAD8(Q)                                                                     \
    uint32_t h[8];                                                                      \
    {                                                                                   \
        const uint4 pa = *reinterpret_cast<const uint4 *>(pts + (Q));                   \
        const uint4 pb = *reinterpret_cast<const uint4 *>(pts + (Q) + 4);               \
        const uint32_t pk[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};        \
        _Pragma("unroll") for (int e = 0; e < 8; e++) {                                 \
            const uint32_t a = mt_pt_addr(pk[e], bm_lane);                              \
            h[e] = __funnelshift_r(mt_lds(a), mt_lds4(a), pk[e]);                       \
        }                                                                               \
    }
#define MT_TREE8(E8)                                                                    \
    {                                                                                   \
        uint32_t t0, t1, f0, f1;                                                        \
        CSA(t0, ones, ones, h[0], h[1]);                                                \
        CSA(t1, ones, ones, h[2], h[3]);                                                \
        CSA(f0, twos, twos, t0, t1);                                                    \
        CSA(t0, ones, ones, h[4], h[5]);                                                \
        CSA(t1, ones, ones, h[6], h[7]);                                                \
        CSA(f1, twos, twos, t0, t1);                                                    \
        CSA(E8, fours, fours, f0, f1);                                                  \
    }
    for (int q0 = 0; q0 < M; q0 += MT_CHUNK) {
        if (ABORT && q0) {
            // hits a translation must have by now to still reach the best complete score
            const int need = (int)(*best_key >> 32) - (M - q0);
            if (need > 0) {
                // bit-sliced "count >= need" over this lane's row, most significant plane first
                uint32_t gt = 0u, eq = rowmask;
#pragma unroll
                for (int pbit = MT_PLANES - 1; pbit >= 0; pbit--) {
                    const uint32_t plane = pbit >= 4 ? pl[pbit] : (pbit == 3 ? eights : (pbit == 2 ? fours : (pbit == 1 ? twos : ones)));
                    const uint32_t tb = (need >> pbit) & 1 ? 0xffffffffu : 0u;
                    gt |= eq & plane & ~tb;
                    eq &= ~(plane ^ tb);
                }
                if (!__any_sync(0xffffffffu, (gt | eq) != 0u)) {
                    if (visited) *visited += q0;
                    if (ovf) sh->overflow = 1;
                    return false;
                }
            }
        }
        const int n = min(MT_CHUNK, M - q0);
        __syncwarp();                                                       // the previous chunk has been read
#pragma unroll
        for (int e = 0; e < MT_CHUNK; e += 32)
            if (e + lane < n) pts[e + lane] = mt_raster_point(c, sh, ccx, ccy, ccf, r, q0 + e + lane, 2 * nx, 2 * ny, ovf);
        __syncwarp();
        int q = 0;
        for (; q + 16 <= n; q += 16) {
            uint32_t ea, eb, s16;
            { MT_LOAD8(q) MT_TREE8(ea) }
            { MT_LOAD8(q + 8) MT_TREE8(eb) }
            CSA(s16, eights, eights, ea, eb);
            // ripple the sixteens into planes 4..8
#pragma unroll
            for (int p = 4; p < MT_PLANES; p++) { uint32_t t = pl[p] & s16; pl[p] ^= s16; s16 = t; }
        }
        if (q + 8 <= n) {                                                   // only the last chunk gets here
            uint32_t e8;
            { MT_LOAD8(q) MT_TREE8(e8) }
            uint32_t t = eights & e8; eights ^= e8; e8 = t;
#pragma unroll
            for (int p = 4; p < MT_PLANES; p++) { uint32_t t2 = pl[p] & e8; pl[p] ^= e8; e8 = t2; }
            q += 8;
        }
        for (; q < n; q++) {                                                // tail: plain ripple add
            const uint32_t pk = pts[q];
            const uint32_t a = mt_pt_addr(pk, bm_lane);
            uint32_t carry = __funnelshift_r(mt_lds(a), mt_lds4(a), pk), t;
            t = ones & carry; ones ^= carry; carry = t;
            t = twos & carry; twos ^= carry; carry = t;
            t = fours & carry; fours ^= carry; carry = t;
            t = eights & carry; eights ^= carry; carry = t;
#pragma unroll
            for (int p = 4; p < MT_PLANES; p++) { t = pl[p] & carry; pl[p] ^= carry; carry = t; }
        }
    }
#undef MT_LOAD8
#undef MT_TREE8
    pl[0] = ones; pl[1] = twos; pl[2] = fours; pl[3] = eights;
    if (visited) *visited += M;
    if (ovf) sh->overflow = 1;
    return true;
}

// Group-bound pass (phase A).  A bound only has to be admissible, so a lane need not own ONE row shift: bmg row r is the OR of
// the dilated rows r .. r + MT_BSUB - 1 (built by the dilation walker), and a lookup at row y + MT_BSUB * b bounds the hits of the
// MT_BSUB row shifts MT_BSUB * b .. MT_BSUB * b + MT_BSUB - 1 at once.  With MT_BSUB = 4 eight lanes cover the 29 row shifts and the
// four quarters of the warp score four different points per instruction: a quarter of the instructions of mt_pass per point.  The
// quarters' bit-sliced counters are added (two shuffle rounds of nine full adders) for every abort check and at the end.
#define MT_BSUB 4
#ifndef MT_BCHUNK
#define MT_BCHUNK 64                    // points between two abort checks of a bound pass (multiple of 32)
#endif
#define MT_BQ (MT_BCHUNK / 4)           // points per quarter and chunk
#define MT_NULLPT ((uint32_t)(RB_BM_STRIDE - 1) << MT_PT_WORD_SHIFT)   // the zeroed pad word of a row, shift 0: never a hit

__device__ __forceinline__ void mt_quarter_sum(uint32_t T[MT_PLANES])
{
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        uint32_t carry = 0u;
#pragma unroll
        for (int p = 0; p < MT_PLANES; p++) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, T[p], o);
            const uint32_t s = T[p] ^ other ^ carry;
            carry = (T[p] & other) | ((T[p] ^ other) & carry);
            T[p] = s;
        }
    }
}

// Returns false when the pass was given up (no translation of the group can reach the best complete score); else ub = the
// largest bound over the translation window.
__device__ __forceinline__ bool mt_bound_pass(const RbCtx &c, MatchShared *sh, const double *ccx, const double *ccy,
                                              const float2 *__restrict__ ccf, int k, int nx, int ny, uint32_t bmg_addr, uint32_t *pts,
                                              int M, int lane, uint32_t colmask, const volatile unsigned long long *best_key,
                                              int *visited, int &ub)
{
    const MtRot r = mt_rot(c, sh, k, -nx, -ny);
    const int qd = lane >> 3, b = lane & 7;
    const bool rows_ok = MT_BSUB * b <= 2 * ny;
    const uint32_t bm_lane = bmg_addr + (uint32_t)((rows_ok ? MT_BSUB * b : 0) * RB_BM_STRIDE * 4);
    const uint32_t rowmask = rows_ok ? colmask : 0u;
    const uint32_t *ptq = pts + MT_BQ * qd;
    int ovf = 0;
    uint32_t ones = 0, twos = 0, fours = 0, eights = 0, pl[MT_PLANES];
#pragma unroll
    for (int p = 0; p < MT_PLANES; p++) pl[p] = 0;
#define MT_LOAD8(Q)                                                                     \
    uint32_t h[8];                                                                      \
    {                                                                                   \
        const uint4 pa = *reinterpret_cast<const uint4 *>(ptq + (Q));                   \
        const uint4 pb = *reinterpret_cast<const uint4 *>(ptq + (Q) + 4);               \
        const uint32_t pk[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};        \
        _Pragma("unroll") for (int e = 0; e < 8; e++) {                                 \
            const uint32_t a = mt_pt_addr(pk[e], bm_lane);                              \
            h[e] = __funnelshift_r(mt_lds(a), mt_lds4(a), pk[e]);                       \
        }                                                                               \
    }
#define MT_TREE8(E8)                                                                    \
    {                                                                                   \
        uint32_t t0, t1, f0, f1;                                                        \
        CSA(t0, ones, ones, h[0], h[1]);                                                \
        CSA(t1, ones, ones, h[2], h[3]);                                                \
        CSA(f0, twos, twos, t0, t1);                                                    \
        CSA(t0, ones, ones, h[4], h[5]);                                                \
        CSA(t1, ones, ones, h[6], h[7]);                                                \
        CSA(f1, twos, twos, t0, t1);                                                    \
        CSA(E8, fours, fours, f0, f1);                                                  \
    }
    for (int q0 = 0; q0 < M; q0 += MT_BCHUNK) {
        if (q0) {
            const int need = (int)(*best_key >> 32) - (M - q0);
            if (need > 0) {
                uint32_t T[MT_PLANES] = {ones, twos, fours, eights, pl[4], pl[5], pl[6], pl[7], pl[8]};
                mt_quarter_sum(T);
                uint32_t gt = 0u, eq = rowmask;
#pragma unroll
                for (int pbit = MT_PLANES - 1; pbit >= 0; pbit--) {
                    const uint32_t tb = (need >> pbit) & 1 ? 0xffffffffu : 0u;
                    gt |= eq & T[pbit] & ~tb;
                    eq &= ~(T[pbit] ^ tb);
                }
                if (!__any_sync(0xffffffffu, (gt | eq) != 0u)) {
                    *visited += q0;
                    if (ovf) sh->overflow = 1;
                    return false;
                }
            }
        }
        const int n = min(MT_BCHUNK, M - q0), npad = (n + 31) & ~31;
        __syncwarp();                                                       // the previous chunk has been read
#pragma unroll
        for (int e = 0; e < MT_BCHUNK; e += 32) {
            const int idx = e + lane;                                       // point idx of the chunk goes to quarter idx & 3
            if (idx < npad)
                pts[(idx & 3) * MT_BQ + (idx >> 2)] = idx < n ? mt_raster_point(c, sh, ccx, ccy, ccf, r, q0 + idx, 2 * nx, 2 * ny, ovf) : MT_NULLPT;
        }
        __syncwarp();
        const int nq = npad >> 2;                                           // multiple of 8
        int q = 0;
        for (; q + 16 <= nq; q += 16) {
            uint32_t ea, eb, s16;
            { MT_LOAD8(q) MT_TREE8(ea) }
            { MT_LOAD8(q + 8) MT_TREE8(eb) }
            CSA(s16, eights, eights, ea, eb);
#pragma unroll
            for (int p = 4; p < MT_PLANES; p++) { uint32_t t = pl[p] & s16; pl[p] ^= s16; s16 = t; }
        }
        if (q < nq) {
            uint32_t e8;
            { MT_LOAD8(q) MT_TREE8(e8) }
            uint32_t t = eights & e8; eights ^= e8; e8 = t;
#pragma unroll
            for (int p = 4; p < MT_PLANES; p++) { uint32_t t2 = pl[p] & e8; pl[p] ^= e8; e8 = t2; }
        }
    }
#undef MT_LOAD8
#undef MT_TREE8
    *visited += M;
    if (ovf) sh->overflow = 1;
    uint32_t T[MT_PLANES] = {ones, twos, fours, eights, pl[4], pl[5], pl[6], pl[7], pl[8]};
    mt_quarter_sum(T);
    uint32_t cand = rowmask;
    int sc = 0;
#pragma unroll
    for (int pbit = MT_PLANES - 1; pbit >= 0; pbit--) {
        const uint32_t t = cand & T[pbit];
        if (t) { cand = t; sc |= 1 << pbit; }
    }
    if (!rowmask) sc = 0;
    for (int o = 4; o > 0; o >>= 1) sc = max(sc, __shfl_xor_sync(0xffffffffu, sc, o));
    ub = sc;
    return true;
}

// curr point of beam j relative to the guess position (hybridmap.py:216-228,236,240; adj: :165-172)
// Returns bit 1 when the beam is a curr point of the matcher (|c| < 11 m, :240) and bit 0 when it is a
// curr point of hybridmap.py:216-228 at all -- those span the 72 x 72-cell windows of the reference set
// (:230-234), also the ones the 11 m filter drops.  (ax, ay) = its cell corner in map coordinates.
__device__ __forceinline__ int mt_curr_point(const RbCtx &c, const MatchShared *sh, unsigned long long exists, double d, double bpx,
                                             double bpy, int adj, double &qx, double &qy, double &ax, double &ay)
{
    double gx, gy;
    rb_xform(sh->cs0, sh->sn0, sh->gx, sh->gy, bpx, bpy, gx, gy);
    ax = ay = 0.0;
    if (adj) {                                                             // hybridmap.py:165-172
        qx = gx - sh->gx; qy = gy - sh->gy;
        return sqrt(qx * qx + qy * qy) < RB_MATCH_MAX_R ? 2 : 0;
    }
    if (!(d < RB_MATCH_MAX_R && d > RB_MATCH_MIN_R)) return 0;
    int tx, ty, ix, iy;
    rb_read_axis(gx, tx, ix);
    rb_read_axis(gy, ty, iy);
    if (!rb_tile_exists(c, exists, tx, ty)) return 0;
    ax = rb_cell_corner(ix, tx); ay = rb_cell_corner(iy, ty);
    qx = ax - sh->gx; qy = ay - sh->gy;
    return sqrt(qx * qx + qy * qy) < RB_MATCH_MAX_R ? 3 : 1;
}

// ---- NDT refinement stage (matchScanCustom.m:32-50) --------------------------------
// Restated in oracle/rbpf_oracle.c (ndt_eval / ndt_refine), PARITY UNPINNED against
// MathWorks matchScans, bit-exact against the oracle: exp / sin / cos are the same
// polynomials in plain IEEE arithmetic, thread = beam, per-beam terms are summed by an
// xor-butterfly inside each warp and then over the warps in order.
#define NDT_TERMS 16
#define NDT_MAX_ITERS 500
#define NDT_LIMIT 235.0
#define NDT_STEP_T 5e-3
#define NDT_STEP_R 5e-5
#define NDT_T1 0.33333333333333331
#define NDT_T2 0.66666666666666663

__device__ __forceinline__ double ndt_exp_neg(double e)
{
    if (!(e < 700.0)) return 0.0;
    const double x = -e;
    const double kf = floor(fma(x, 1.4426950408889634, 0.5));
    const double r = fma(-kf, 1.90821492927058770002e-10, fma(-kf, 0.693147180369123816490, x));
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p * __longlong_as_double((long long)((int)kf + 1023) << 52);
}

__device__ __forceinline__ void ndt_sincos(double a, double &sn, double &cs)
{
    const double z = a * a;
    double s = -1.0 / 121645100408832000.0;
    s = s * z + 1.0 / 355687428096000.0;
    s = s * z - 1.0 / 1307674368000.0;
    s = s * z + 1.0 / 6227020800.0;
    s = s * z - 1.0 / 39916800.0;
    s = s * z + 1.0 / 362880.0;
    s = s * z - 1.0 / 5040.0;
    s = s * z + 1.0 / 120.0;
    s = s * z - 1.0 / 6.0;
    s = s * z + 1.0;
    sn = s * a;
    double c = 1.0 / 2432902008176640000.0;
    c = c * z - 1.0 / 6402373705728000.0;
    c = c * z + 1.0 / 20922789888000.0;
    c = c * z - 1.0 / 87178291200.0;
    c = c * z + 1.0 / 479001600.0;
    c = c * z - 1.0 / 3628800.0;
    c = c * z + 1.0 / 40320.0;
    c = c * z - 1.0 / 720.0;
    c = c * z + 1.0 / 24.0;
    c = c * z - 0.5;
    c = c * z + 1.0;
    cs = c;
}

// One beam's terms (score, gradient, Hessian, curvature model) at p, folded over the
// warp; lane 2e of every warp writes the warp's sum of term e to red[warp][e].
__device__ __forceinline__ void ndt_partial(const uint32_t *__restrict__ bm, double *red, int xoff0, double fx, double fy,
                                            bool has, double cx, double cy, double p0, double p1, double sn, double cs)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a[NDT_TERMS];
#pragma unroll
    for (int e = 0; e < NDT_TERMS; e++) a[e] = 0.0;
    if (has) {
        const double X = cs * cx - sn * cy, Y = sn * cx + cs * cy;
        const double X20 = X * 20.0, Y20 = Y * 20.0;
        const double u = (X + fx) * 20.0 + p0, v = (Y + fy) * 20.0 + p1;
        if (fabs(u) < NDT_LIMIT && fabs(v) < NDT_LIMIT) {
            const int iu = __double2int_rd(u + 0.5), iv = __double2int_rd(v + 0.5);
            const int bx0 = iu - 1 + xoff0, by0 = iv - 1 + RB_WIN_R;
            uint32_t rows[3];
#pragma unroll
            for (int dr = 0; dr < 3; dr++) {
                const int w = (by0 + dr) * RB_BM_STRIDE + (bx0 >> 5);
                rows[dr] = __funnelshift_r(bm[w], bm[w + 1], bx0 & 31) & 7u;
            }
            double f = 0.0, fu = 0.0, fv = 0.0, fuu = 0.0, fuv = 0.0, fvv = 0.0, cuu = 0.0, cuv = 0.0, cvv = 0.0;
#pragma unroll
            for (int by = 0; by < 2; by++)
#pragma unroll
                for (int bx = 0; bx < 2; bx++) {
                    const int pat = (int)((rows[by] >> bx) & 3u) | ((int)((rows[by + 1] >> bx) & 3u) << 2);
                    const int n = __popc(pat);
                    if (n < 3) continue;
                    double mx = 0.5, my = 0.5, Bd = 4.0, Bxy = 0.0;
                    if (n == 3) {
                        const int miss = pat == 14 ? 0 : (pat == 13 ? 1 : (pat == 11 ? 2 : 3));
                        mx = (miss & 1) ? NDT_T1 : NDT_T2;
                        my = (miss >> 1) ? NDT_T1 : NDT_T2;
                        Bd = 6.0;
                        Bxy = ((miss & 1) == (miss >> 1)) ? 3.0 : -3.0;
                    }
                    const double ru = u - (double)(iu - 1 + bx), rv = v - (double)(iv - 1 + by);
                    const double tu = ru - 0.5, tv = rv - 0.5;
                    const double au = fmin(fabs(tu), 1.0), av = fmin(fabs(tv), 1.0);
                    const double su = tu < 0.0 ? -1.0 : 1.0, sv = tv < 0.0 ? -1.0 : 1.0;
                    const double wx = fma(-(au * au), fma(-2.0, au, 3.0), 1.0), wy = fma(-(av * av), fma(-2.0, av, 3.0), 1.0);
                    const double wx1 = su * (6.0 * au * (au - 1.0)), wy1 = sv * (6.0 * av * (av - 1.0));
                    const double wx2 = fma(12.0, au, -6.0), wy2 = fma(12.0, av, -6.0);
                    const double dx = ru - mx, dy = rv - my;
                    const double a1 = fma(Bd, dx, Bxy * dy), a2 = fma(Bxy, dx, Bd * dy);
                    const double G = ndt_exp_neg(0.5 * fma(dx, a1, dy * a2));
                    const double W = wx * wy, WG = W * G, Wu = wx1 * wy, Wv = wx * wy1;
                    const double Gu = -(a1 * G), Gv = -(a2 * G);
                    f += WG;
                    fu = fma(-a1, WG, fma(Wu, G, fu));
                    fv = fma(-a2, WG, fma(Wv, G, fv));
                    fuu = fma(fma(a1, a1, -Bd), WG, fma(2.0 * Wu, Gu, fma(wx2 * wy, G, fuu)));
                    fuv = fma(fma(a1, a2, -Bxy), WG, fma(Wv, Gu, fma(Wu, Gv, fma(wx1 * wy1, G, fuv))));
                    fvv = fma(fma(a2, a2, -Bd), WG, fma(2.0 * Wv, Gv, fma(wx * wy2, G, fvv)));
                    cuu = fma(fmax(-wx2, 0.0) * wy, G, fma(WG, Bd, cuu));
                    cuv = fma(WG, Bxy, cuv);
                    cvv = fma(wx * fmax(-wy2, 0.0), G, fma(WG, Bd, cvv));
                }
            const double J3x = -Y20, J3y = X20;
            const double h13 = fuu * J3x + fuv * J3y, h23 = fuv * J3x + fvv * J3y;
            const double c13 = cuu * J3x + cuv * J3y, c23 = cuv * J3x + cvv * J3y;
            a[0] = f; a[1] = fu; a[2] = fv; a[3] = fu * J3x + fv * J3y;
            a[4] = fuu; a[5] = fuv; a[6] = h13; a[7] = fvv; a[8] = h23;
            a[9] = (J3x * h13 + J3y * h23) - (fu * X20 + fv * Y20);
            a[10] = cuu; a[11] = cuv; a[12] = c13; a[13] = cvv; a[14] = c23;
            a[15] = J3x * c13 + J3y * c23;
        }
    }
    // warp reduction: a butterfly whose every step halves the terms a lane carries
    // (same pair sums as the plain xor-butterfly); lane l ends with term l >> 1.
#define NDT_FOLD(HALF, O)                                                        \
    {                                                                            \
        const bool up = lane & (O);                                              \
        _Pragma("unroll") for (int i = 0; i < (HALF); i++) {                     \
            const double send = up ? a[i] : a[i + (HALF)];                       \
            const double keep = up ? a[i + (HALF)] : a[i];                       \
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, (O));               \
        }                                                                        \
    }
    NDT_FOLD(8, 16)
    NDT_FOLD(4, 8)
    NDT_FOLD(2, 4)
    NDT_FOLD(1, 2)
#undef NDT_FOLD
    a[0] = a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
    if (!(lane & 1)) red[warp * NDT_TERMS + (lane >> 1)] = a[0];
}

__device__ __forceinline__ bool ndt_solve(double A00, double A10, double A20, double A11, double A21, double A22,
                                          const double *g, double *d)
{
    if (!(A00 > 1e-12)) return false;
    const double i00 = 1.0 / sqrt(A00), l10 = A10 * i00, l20 = A20 * i00;
    const double t1 = A11 - l10 * l10;
    if (!(t1 > 1e-12)) return false;
    const double i11 = 1.0 / sqrt(t1), l21 = (A21 - l20 * l10) * i11;
    const double t2 = (A22 - l20 * l20) - l21 * l21;
    if (!(t2 > 1e-12)) return false;
    const double i22 = 1.0 / sqrt(t2);
    const double y0 = g[0] * i00, y1 = (g[1] - l10 * y0) * i11, y2 = ((g[2] - l20 * y0) - l21 * y1) * i22;
    d[2] = y2 * i22;
    d[1] = (y1 - l21 * d[2]) * i11;
    d[0] = ((y0 - l10 * d[1]) - l20 * d[2]) * i00;
    return true;
}

// The ascent loop of the oracle's ndt_refine() as a coroutine: every thread takes part in
// the score evaluations, warp 0 alone (all lanes alike) owns the optimiser state, decides,
// and posts the next correction to evaluate -- or the end -- in ctl[].
//   red : MT_WARPS x NDT_TERMS doubles, ctl : 8 doubles {go, p0, p1, p2, S, evals, sin p2, cos p2}.
__device__ __noinline__ void ndt_refine(const uint32_t *__restrict__ bm, double *red, double *ctl, int xoff0, double fx,
                                        double fy, bool has, double cx, double cy, double q0, double q1, double q2)
{
    const int lane = threadIdx.x & 31;
    const bool boss = threadIdx.x < 32;
    double t[NDT_TERMS], p[3] = {q0, q1, q2}, pn[3] = {q0, q1, q2}, d[3], lam = 1e-3;
    int ne = 0, it = 0;
    bool newton_ok = true, newton = false, first = true;
    double sn, cs;
    ndt_sincos(q2, sn, cs);
    for (;;) {
        ndt_partial(bm, red, xoff0, fx, fy, has, cx, cy, pn[0], pn[1], sn, cs);
        __syncthreads();
        if (boss) {
            // lanes 0-15 sum warps 0-5 of their term, lanes 16-31 warps 6-11, in order; then the halves
            double s, tn[NDT_TERMS];
            {
                const double *col = red + (lane >> 4) * (MT_WARPS / 2) * NDT_TERMS + (lane & 15);
                s = col[0];
#pragma unroll
                for (int w = 1; w < MT_WARPS / 2; w++) s = s + col[w * NDT_TERMS];
                const double other = __shfl_xor_sync(0xffffffffu, s, 16);
                s = lane < 16 ? s + other : other + s;
            }
#pragma unroll
            for (int e = 0; e < NDT_TERMS; e++) tn[e] = __shfl_sync(0xffffffffu, s, e);
            ne++;
            bool done = false;
            // ---- the evaluation that just came back (oracle: second half of the loop body)
            if (first) {
                first = false;
#pragma unroll
                for (int e = 0; e < NDT_TERMS; e++) t[e] = tn[e];
            } else {
                if (tn[0] > t[0]) {
                    const double gain = tn[0] - t[0];
#pragma unroll
                    for (int e = 0; e < NDT_TERMS; e++) t[e] = tn[e];
                    for (int q = 0; q < 3; q++) p[q] = pn[q];
                    if (!newton) lam = lam * 0.1 < 1e-3 ? 1e-3 : lam * 0.1;
                    newton_ok = true;
                    if (gain < 1e-6) done = true;
                } else if (newton) {
                    newton_ok = false;
                } else {
                    lam *= 10.0;
                    if (lam > 1e9) done = true;
                }
                it++;
            }
            // ---- iterations that need no evaluation, up to the next proposal (first half)
            bool go = false;
            while (!done && it < NDT_MAX_ITERS) {
                newton = newton_ok && ndt_solve(-t[4], -t[5], -t[6], -t[7], -t[8], -t[9], t + 1, d) &&
                         fabs(d[0]) < 1.0 && fabs(d[1]) < 1.0 && fabs(d[2]) < 0.01;
                bool ok = newton;
                if (!newton)
                    ok = ndt_solve(t[10] + lam * t[10], t[11], t[12], t[13] + lam * t[13], t[14], t[15] + lam * t[15], t + 1, d);
                if (!ok && !newton && !(t[10] > 0.0)) break;
                if (ok) {
                    if (fabs(d[0]) < NDT_STEP_T && fabs(d[1]) < NDT_STEP_T && fabs(d[2]) < NDT_STEP_R) break;
                    for (int q = 0; q < 3; q++) pn[q] = p[q] + d[q];
                    ok = fabs(pn[0]) < 64.0 && fabs(pn[1]) < 64.0 && fabs(pn[2]) < 1.0;
                }
                if (ok) { go = true; break; }
                if (newton) {
                    newton_ok = false;
                } else {
                    lam *= 10.0;
                    if (lam > 1e9) break;
                }
                it++;
            }
            if (go) ndt_sincos(pn[2], sn, cs);
            if (lane == 0) {
                ctl[6] = sn;
                ctl[7] = cs;
                ctl[0] = go ? 1.0 : 0.0;
                ctl[1] = go ? pn[0] : p[0];
                ctl[2] = go ? pn[1] : p[1];
                ctl[3] = go ? pn[2] : p[2];
                ctl[4] = t[0];
                ctl[5] = (double)ne;
            }
        }
        __syncthreads();
        if (ctl[0] == 0.0) break;
        if (!boss) { pn[0] = ctl[1]; pn[1] = ctl[2]; sn = ctl[6]; cs = ctl[7]; }
    }
}

// adj = 0: scan-to-map (HybridMap.get_scan_match hybridmap.py:210-261)
// adj = 1: scan-to-previous-scan (HybridMap.get_scan_adj hybridmap.py:147-191): curr points are
//          the unsnapped endpoints, occupancy is the previous scan rasterised on the lattice.
__global__ void __launch_bounds__(MT_THREADS, 2) match_kernel(RbCtx c, int p_offset, int *__restrict__ slice_out, int adj, int sel)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *bm = smem;
    uint32_t *bmg = bm + mt_bm_words();                                     // bm dilated by MT_GRAD, rows r .. r + MT_BSUB - 1 ORed (group bounds)
    uint32_t *raw = bmg + mt_bm_words();                                    // staging, then per-warp point lists
    size_t w_end = (2 * mt_bm_words() + mt_raw_words() + 1) & ~(size_t)1;
    double *ccx = reinterpret_cast<double *>(smem + w_end);
    double *ccy = ccx + RB_MAXB;
    float2 *ccf = reinterpret_cast<float2 *>(ccy + RB_MAXB);                // the same points in cells, float32
    MatchShared *sh = reinterpret_cast<MatchShared *>(ccf + RB_MAXB);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x + p_offset;
    if (c.use_dup && !slice_out && c.dup_of[p] != p) return;                // a bit-identical duplicate: result copied afterwards
    if (sel && (c.pulled[p] != 0) != (sel == 2)) return;                    // two launches around the arrival of migrated sub-tiles
    const double *pose = c.pose + 3 * (size_t)p, *cov = c.cov + 9 * (size_t)p;
    long long clk_prev = clock64();
    if (tid < MT_NRB) sh->need[tid] = 0u;
    // this thread's beam (one per thread) and the tile mask are on their way while thread 0 sets up the frame
    const bool has_beam = tid < c.B;
    const double b_d = has_beam ? c.dist[tid] : 0.0, b_px = has_beam ? c.px[tid] : 0.0, b_py = has_beam ? c.py[tid] : 0.0;
    const unsigned long long exists = c.exists[p];

    // ---- 0. per-particle frame -------------------------------------------
    if (tid == 0) {
        sh->gx = pose[0]; sh->gy = pose[1]; sh->gth = pose[2];
        { double sn_, cs_; rb_sincos(pose[2], &sn_, &cs_); sh->cs0 = cs_; sh->sn0 = sn_; }
        // search window, robot.py:62-65
        double p0 = sqrt(cov[0]) * 30.0, p1 = sqrt(cov[4]) * 30.0;
        sh->rx = fmax(fmin(4 * p0, 0.7), 0.1);
        sh->ry = fmax(fmin(4 * p1, 0.7), 0.1);
        sh->nx = min(__double2int_rd(sh->rx / RB_CS + 1e-9), RB_NT_MAX);
        sh->ny = min(__double2int_rd(sh->ry / RB_CS + 1e-9), RB_NT_MAX);
        int tx, ty, ix, iy;
        rb_read_axis(pose[0], tx, ix);
        rb_read_axis(pose[1], ty, iy);
        sh->ok = !(tx < -c.txh || tx > c.txh || ty < -c.tyh || ty > c.tyh);
        sh->t0x = tx; sh->t0y = ty;
        sh->g0xu = 800 * (tx + c.txh) + ix;
        sh->g0yu = 800 * (ty + c.tyh) + iy;
        sh->fx = pose[0] - rb_cell_corner(ix, tx);
        sh->fy = pose[1] - rb_cell_corner(iy, ty);
        sh->x0 = (sh->g0xu - RB_WIN_R) & ~31;
        sh->y0 = sh->g0yu - RB_WIN_R;
        sh->a0 = (sh->y0 - 1) & ~3;
        sh->dy0 = sh->y0 - 1 - sh->a0;
        sh->sxb0 = (sh->x0 - 32 + 160 * 1024) / RB_SUB - 1024;             // floor division, also for negative coordinates
        sh->syb0 = (sh->a0 + 160 * 1024) / RB_SUB - 1024;
        sh->nblk = 0;
        sh->slot_warp = -1;
        sh->M = 0;
        sh->overflow = 0;
        sh->best_key = 0ull;
        sh->next_item = 0;
        sh->next_group = 0;
        sh->evals = 0;
        sh->visits = 0;
#pragma unroll
        for (int q = 0; q < 9; q++) sh->mom[q] = 0;
    }
    __syncthreads();

    // ---- 1. curr points, hybridmap.py:216-228,236,240 ----------------------
    if (tid < 25 && !adj) {                                                 // page-table entries under the window
        const int sxb = sh->sxb0 + tid % 5, syb = sh->syb0 + tid / 5;
        sh->ptw[tid] = (sxb >= 0 && sxb < c.subs_x && syb >= 0 && syb < c.subs_y) ? c.pt[(size_t)p * c.nsub + syb * c.subs_x + sxb] : RB_NONE;
    }
    static_assert(MT_THREADS >= RB_MAXB, "one thread per beam");
    double cp_x = 0.0, cp_y = 0.0;                                          // this thread's beam: cell corner in map coordinates
    int cp_kind = 0;
    {
        double qx = 0.0, qy = 0.0;
        cp_kind = has_beam ? mt_curr_point(c, sh, exists, b_d, b_px, b_py, adj, qx, qy, cp_x, cp_y) : 0;
        const bool is_pt = cp_kind & 2;
        const unsigned bal = __ballot_sync(0xffffffffu, is_pt);
        int base = 0;
        if (bal && lane == 0) base = atomicAdd(&sh->M, __popc(bal));        // one slot range per warp
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_pt) {
            const int slot = base + __popc(bal & ((1u << lane) - 1u));
            ccx[slot] = qx;
            ccy[slot] = qy;
        }
    }
    __syncthreads();
#ifndef MT_NO_SHUFFLE
    // Spread the points: consecutive list entries are consecutive beams (one wall
    // segment), so the first 64 points of a pass would tell little about the rest.
    // A stride permutation makes every prefix a sample of the whole sweep and the
    // abortable passes give up earlier.  Scores are sums: the order changes nothing.
    {
        double *tx_ = reinterpret_cast<double *>(bmg), *ty_ = tx_ + RB_MAXB;     // bmg is free until the dilation
        const int Mp = sh->M;
        const int s_ = Mp % 37 ? 37 : (Mp % 41 ? 41 : 43);                      // a prime that does not divide M
        for (int q = tid; q < Mp; q += MT_THREADS) { tx_[q] = ccx[q]; ty_[q] = ccy[q]; }
        __syncthreads();
        // ... and far points first (stable within four range classes): a wrong rotation moves a
        // point by range x angle, so the far points are the ones that miss, and the passes that
        // cannot win are given up after fewer points.
        double mx_ = 0.0, my_ = 0.0;
        int key = -1;
        if (tid < Mp) {
            const int src = (int)(((unsigned)tid * (unsigned)s_) % (unsigned)Mp);
            mx_ = tx_[src]; my_ = ty_[src];
            const double r2 = mx_ * mx_ + my_ * my_;
            key = r2 >= 64.0 ? 0 : (r2 >= 25.0 ? 1 : (r2 >= 6.25 ? 2 : 3));           // 8 m, 5 m, 2.5 m
        }
        int intra = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const unsigned mb = __ballot_sync(0xffffffffu, key == b);
            if (key == b) intra = __popc(mb & ((1u << lane) - 1u));
            if (lane == 0) sh->ord_cnt[b * MT_WARPS + warp] = __popc(mb);
        }
        __syncthreads();
        if (warp == 0) {                                                        // exclusive prefix, class-major then warp
            static_assert(4 * MT_WARPS <= 64, "two entries per lane");
            const int e0 = 2 * lane, e1 = 2 * lane + 1;
            const int n0 = e0 < 4 * MT_WARPS ? sh->ord_cnt[e0] : 0, n1 = e1 < 4 * MT_WARPS ? sh->ord_cnt[e1] : 0;
            int inc = n0 + n1;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            const int ex = inc - n0 - n1;
            if (e0 < 4 * MT_WARPS) sh->ord_cnt[e0] = ex;
            if (e1 < 4 * MT_WARPS) sh->ord_cnt[e1] = ex + n0;
        }
        __syncthreads();
        if (key >= 0) {
            const int slot = sh->ord_cnt[key * MT_WARPS + warp] + intra;
            ccx[slot] = mx_;
            ccy[slot] = my_;
        }
        __syncthreads();
    }
#endif
    if (tid < sh->M) ccf[tid] = make_float2((float)(ccx[tid] * 20.0), (float)(ccy[tid] * 20.0));
    MT_CLK(0)
    const int a0 = sh->a0, dy0 = sh->dy0;

    // ---- 2. occupancy bitmap around the guess cell ---------------------------
    if (adj) {
        // previous scan rasterised relative to the guess cell, like curr points at rotation 0
        for (int idx = tid; idx < MT_RAW_ROWS * RB_RAW_STRIDE; idx += MT_THREADS) raw[idx] = 0u;
        if (tid < MT_NRB) sh->need[tid] = 0xffffffffu;
        __syncthreads();
        const int xb = sh->g0xu - sh->x0 + 32, yb = RB_WIN_R + 1 + dy0;
        for (int q = tid; q < c.n_prev; q += MT_THREADS) {
            const double qx = c.prev_x[q] - sh->gx, qy = c.prev_y[q] - sh->gy;
            if (!(sqrt(qx * qx + qy * qy) < RB_MATCH_MAX_R)) continue;     // hybridmap.py:171
            const int a = __double2int_rd((qx + sh->fx) * 20.0 + 0.5) + xb, b = __double2int_rd((qy + sh->fy) * 20.0 + 0.5) + yb;
            if (a < 0 || a >= 32 * RB_RAW_STRIDE || b < 0 || b >= MT_RAW_ROWS) continue;
            atomicOr(&raw[b * RB_RAW_STRIDE + (a >> 5)], 1u << (a & 31));
        }
    } else {
        // ---- 2a. the reference's `ref` set, hybridmap.py:230-239 -------------------------------
        // Occupied cells reach the matcher only when they lie in the 72 x 72-cell window
        // [j - 36, j + 36) of some curr point in some tile (GridMap._get_rel_cell /
        // get_nearby_occ_points, gridmap.py:130-155: float expression and dec quirk replayed per point
        // and tile) and within 11.5 m of the guess.  Union of equal squares = dilation of their centres:
        // every point marks its 72-bit row span at its (per tile) centre row, a running OR over 72 rows
        // (prefix / suffix ORs over blocks of 72 rows) does the rest; tile extents clip both ways.  The
        // mask is built in `raw`; the gather then fetches only 32 x 32-cell blocks whose mask is not empty.
        const int x0 = sh->x0, y0 = sh->y0;
        uint32_t *H = bm, *Gp = bmg;                                        // both free until the dilation pass
        for (int idx = tid; idx < MT_RAW_ROWS * RB_RAW_STRIDE; idx += MT_THREADS) raw[idx] = 0u;
        // rows: |ref| < 11.5 m (:239) as a column interval [lo, hi) per bitmap row, float64 predicate as in the oracle:
        // sqrt(dx^2 + dy^2) < 11.5 holds exactly when dx^2 + dy^2 < 132.25 (sqrt is correctly rounded and monotone, 132.25 and
        // 11.5 are exact, and the double below 132.25 has a root 0.7 ulp below 11.5), so the per-column test is one fused
        // multiply-add on a table of the 512 column offsets -- the float64 divisions of index_to_distance happen once per column
        // and row instead of once per probe.  The table sits in Gp, which is free until the first prefix pass.
        double *colx = reinterpret_cast<double *>(Gp);
        for (int cb = tid; cb < 512; cb += MT_THREADS) {
            const int ux = x0 + cb;
            const int tX = (ux + 800 * 1024) / 800 - 1024, ix = ux - 800 * tX;
            colx[cb] = rb_cell_corner(ix, tX - c.txh) - sh->gx;
        }
        __syncthreads();
        for (int r = tid; r < RB_BM_ROWS; r += MT_THREADS) {
            const int uy = y0 + r;
            const int tY = (uy + 800 * 1024) / 800 - 1024, iy = uy - 800 * tY;
            const double dyr = rb_cell_corner(iy, tY - c.tyh) - sh->gy;
            int lo = 0, hi = 0;
            const double lim = RB_MATCH_MAX_R + 0.5, lim2 = lim * lim;
            const double dy2 = dyr * dyr;
            if (dy2 < lim2) {
                auto inside = [&](int cbit) { const double dxr = colx[cbit]; return dxr * dxr + dy2 < lim2; };
                // column c is dx = (c - cg) * 0.05 - fx away from the guess (0 <= fx < 0.05): a float32 estimate of the interval
                // ends, then the float64 predicate decides (the interval is contiguous: the column offsets are monotone)
                const int half = (int)(sqrtf((float)(lim2 - dy2)) * 20.0f);
                const int cg = sh->g0xu - x0;
                lo = min(max(cg - half, 0), 512);
                hi = min(max(cg + half + 2, lo), 512);
                while (lo > 0 && inside(lo - 1)) lo--;
                while (lo < hi && !inside(lo)) lo++;
                while (hi < 512 && inside(hi)) hi++;
                while (hi > lo && !inside(hi - 1)) hi--;
            }
            sh->row_lo[r] = (unsigned short)lo;
            sh->row_hi[r] = (unsigned short)hi;
        }
        // tiles under the bitmap window (storage tile indices), at most 2 x 2
        const int TXa = max(x0 / 800, 0), TXb = min((x0 + 511) / 800, c.tiles_x - 1);
        const int TYa = max(y0 < 0 ? -1 : y0 / 800, 0), TYb = min((y0 + RB_BM_ROWS - 1) / 800, c.tiles_y - 1);
        for (int TY = TYa; TY <= TYb; TY++) {
            for (int idx = tid; idx < RB_BM_ROWS * RB_BM_STRIDE; idx += MT_THREADS) H[idx] = 0u;
            __syncthreads();
            if (cp_kind & 1) {
                for (int TX = (x0 < 0 ? 0 : TXa); TX <= TXb; TX++) {
                    if (!((exists >> (TY * c.tiles_x + TX)) & 1ull)) continue;   // the reference walks its own tile list
                    const double xr = cp_x - RB_TILE_LEN * (double)(TX - c.txh), yr = cp_y - RB_TILE_LEN * (double)(TY - c.tyh);
                    int decx = 0, decy = 0;
                    if (yr < -20.0) decy = 1; else if (xr < -20.0) decx = 1;         // gridmap.py:131-136
                    const int jx = rb_trunc(xr / RB_TILE_LEN * 800.0 + 400.0) - decx, jy = rb_trunc(yr / RB_TILE_LEN * 800.0 + 400.0) - decy;
                    if (jy - 36 >= 800 || jy + 36 <= 0) continue;                    // window misses the tile (rows)
                    const int xa = max(jx - 36, 0), xb = min(jx + 36, 800);
                    if (xa >= xb) continue;
                    const int row = 800 * TY + jy - y0;                              // centre row in the bitmap (always inside: |c| < 11.1 m)
                    int ca = 800 * TX + xa - x0, cb = 800 * TX + xb - x0;            // columns [ca, cb)
                    ca = max(ca, 0); cb = min(cb, 512);
                    if (row < 0 || row >= RB_BM_ROWS || ca >= cb) continue;
                    uint32_t *hr = H + row * RB_BM_STRIDE;
                    for (int wq = ca >> 5; wq <= (cb - 1) >> 5; wq++) {
                        const int b0 = max(ca - 32 * wq, 0), b1 = min(cb - 32 * wq, 32);
                        const uint32_t bits = (b1 >= 32 ? 0xffffffffu : ((1u << b1) - 1u)) & ~((1u << b0) - 1u);
                        atomicOr(&hr[wq], bits);
                    }
                }
            }
            __syncthreads();
            // prefix ORs of every block of 72 rows into Gp, suffix ORs in place (one thread per word column and block)
            if (tid < 16 * 7) {
                const int wq = tid & 15, blk = tid >> 4;
                const int ra = 72 * blk, rb_ = min(ra + 72, RB_BM_ROWS);
                uint32_t acc = 0u;
                for (int r = ra; r < rb_; r += 8) {                         // eight loads in flight, then the OR chain
                    uint32_t v[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = r + e < rb_ ? H[(r + e) * RB_BM_STRIDE + wq] : 0u;
#pragma unroll
                    for (int e = 0; e < 8; e++) { acc |= v[e]; if (r + e < rb_) Gp[(r + e) * RB_BM_STRIDE + wq] = acc; }
                }
                acc = 0u;
                for (int r = rb_ - 1; r >= ra; r -= 8) {
                    uint32_t v[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = r - e >= ra ? H[(r - e) * RB_BM_STRIDE + wq] : 0u;
#pragma unroll
                    for (int e = 0; e < 8; e++) { acc |= v[e]; if (r - e >= ra) H[(r - e) * RB_BM_STRIDE + wq] = acc; }
                }
            }
            __syncthreads();
            // rows of this tile row: window [r - 35, r + 36] of centre rows, clipped to 11.5 m; into raw (+ block list)
            const int rlo = max(800 * TY - y0, 0), rhi = min(800 * TY + 800 - y0, RB_BM_ROWS);
            // one thread per half row (8 words): the row's clip interval and window rows are set up once, one shared-memory
            // atomic per half row says which 32-cell words of the band hold anything
            for (int it = tid; it < (rhi - rlo) * 2; it += MT_THREADS) {
                const int r = rlo + (it >> 1), w0 = (it & 1) * 8;
                const int a_ = r - 35, b_ = min(r + 36, RB_BM_ROWS - 1);
                const bool use_h = a_ > 0, use_g = a_ <= 0 || a_ / 72 != b_ / 72;
                const uint32_t *ha = H + max(a_, 0) * RB_BM_STRIDE + w0, *gb = Gp + b_ * RB_BM_STRIDE + w0;
                const int lo = (int)sh->row_lo[r] - 32 * w0, hi = (int)sh->row_hi[r] - 32 * w0;
                const int rho = r + dy0 + 1;
                uint32_t *out = raw + rho * RB_RAW_STRIDE + w0 + 1;
                uint32_t nb = 0u;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    uint32_t v = (use_h ? ha[e] : 0u) | (use_g ? gb[e] : 0u);
                    const int l = lo - 32 * e, h_ = hi - 32 * e;
                    uint32_t keep = 0u;
                    if (h_ > 0 && l < 32) keep = (h_ >= 32 ? 0xffffffffu : ((1u << h_) - 1u)) & ~((1u << max(l, 0)) - 1u);
                    v &= keep;
                    if (v) { out[e] = v; nb |= 1u << (w0 + e + 1); }
                }
                if (nb) atomicOr(&sh->need[rho >> 5], nb);
            }
            __syncthreads();
        }
        static_assert(MT_NBLK <= MT_THREADS, "one thread per block of the staging window");
        {
            bool want = false;
            if (tid < MT_NBLK) {
                const int rb = tid / RB_RAW_STRIDE, wc = tid - rb * RB_RAW_STRIDE;
                want = (sh->need[rb] >> wc) & 1u;
            }
            const unsigned wb = __ballot_sync(0xffffffffu, want);              // one atomic per warp
            int base = 0;
            if (lane == 0 && wb) base = atomicAdd(&sh->nblk, __popc(wb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (want) sh->blist[base + __popc(wb & ((1u << lane) - 1u))] = (unsigned short)tid;
        }
        MT_CLK(3)
        __syncthreads();
        const int nblk = sh->nblk;
        unsigned char *rawb = reinterpret_cast<unsigned char *>(raw);
        // One warp per block, one lane per 32-byte sector (8 cells of 4 rows): whole sectors
        // in flight, MT_GATHER_U blocks per round, page-table entries from shared memory (one
        // global-memory latency per round).  Thresholded bytes go straight to their place.
        const int sx = lane & 3, sy = lane >> 2;
        const int sxb0 = sh->sxb0, syb0 = sh->syb0;
        for (int i0 = warp; i0 < nblk; i0 += MT_GATHER_U * MT_WARPS) {
            uint4 v[MT_GATHER_U][2];
            int row0[MT_GATHER_U], wcs[MT_GATHER_U];
#pragma unroll
            for (int u = 0; u < MT_GATHER_U; u++) {
                const int i = i0 + u * MT_WARPS;
                v[u][0] = v[u][1] = make_uint4(0u, 0u, 0u, 0u);
                row0[u] = -1;
                if (i >= nblk) continue;
                const int b = sh->blist[i], rb = b / RB_RAW_STRIDE, wc = b - rb * RB_RAW_STRIDE;
                const int r0 = 32 * rb + 4 * sy;
                if (r0 >= MT_RAW_ROWS) continue;
                row0[u] = r0; wcs[u] = wc;
                const int uy = a0 + r0, ux = x0 - 32 + 32 * wc + 8 * sx;
                if (uy >= 0 && uy < c.uy_max && ux >= 0 && ux < c.ux_max) {
                    const int syb = uy / RB_SUB, sxb = ux / RB_SUB;
                    const uint32_t t = sh->ptw[(syb - syb0) * 5 + (sxb - sxb0)];
                    if (t != RB_NONE) {
                        const uint4 *src = reinterpret_cast<const uint4 *>(c.pool + (size_t)t * RB_SUB_BYTES +
                                                                           RB_OFF_Y(uy - syb * RB_SUB) + RB_OFF_X(ux - sxb * RB_SUB));
                        v[u][0] = __ldg(src);
                        v[u][1] = __ldg(src + 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < MT_GATHER_U; u++) {
                if (row0[u] < 0) continue;
                unsigned char *dst = rawb + ((size_t)row0[u] * RB_RAW_STRIDE + wcs[u]) * 4 + sx;
                dst[0] &= (unsigned char)(mt_pack4(v[u][0].x) | (mt_pack4(v[u][0].y) << 4));
                dst[RB_RAW_STRIDE * 4] &= (unsigned char)(mt_pack4(v[u][0].z) | (mt_pack4(v[u][0].w) << 4));
                dst[2 * RB_RAW_STRIDE * 4] &= (unsigned char)(mt_pack4(v[u][1].x) | (mt_pack4(v[u][1].y) << 4));
                dst[3 * RB_RAW_STRIDE * 4] &= (unsigned char)(mt_pack4(v[u][1].z) | (mt_pack4(v[u][1].w) << 4));
            }
        }
    }
    __syncthreads();
    MT_CLK(1)

    // ---- 2b. both dilations in one pass down the rows --------------------------------
    //   bm  = raw dilated by 1 (a lookup counts when the cell or one of its 8 neighbours is
    //         occupied: the proximity kernel of our matcher)
    //   bmg = raw dilated by 1 + MT_GRAD (upper bound for every rotation of a group), rows r .. r + MT_BSUB - 1
    //         ORed into row r (a bound lookup covers MT_BSUB row shifts, mt_bound_pass)
    // A thread owns one 32-cell word column over MT_SEG_ROWS rows and walks down the staged
    // rows once: three loads per row, the vertical windows are running ORs in registers
    // (13 + 3 = 16 rows: doubling 2, 4, 8, 16).  Rows of raw beyond the buffer
    // count as empty, like rows of bm beyond the window.
    {
        static_assert(MT_GRAD == 5 && MT_BSUB == 4, "the running OR below is a 16-row window: 13 rows of dilation + MT_BSUB - 1");
        static_assert((MT_THREADS / 16) * MT_SEG_ROWS >= RB_BM_ROWS, "segments must cover the window");
        const int w = tid & 15, r0 = (tid >> 4) * MT_SEG_ROWS, r1 = min(r0 + MT_SEG_ROWS, RB_BM_ROWS);
        bool live = false;
        if (r0 < RB_BM_ROWS) {
            const int rb0 = max(r0 + dy0 + 1 - 6, 0) >> 5, rb1 = min(r1 + dy0 + 6 + MT_BSUB - 1, MT_RAW_ROWS - 1) >> 5;
            for (int rb = rb0; rb <= rb1; rb++) live |= (sh->need[rb] >> (w + 1)) & 1u;
        }
        if (r0 < RB_BM_ROWS && !live) {
            for (int r = r0; r < r1; r++) { bm[r * RB_BM_STRIDE + w] = 0u; bmg[r * RB_BM_STRIDE + w] = 0u; }
        } else if (r0 < RB_BM_ROWS) {
            // step s reads raw row rho = rs + s; bm row rho - dy0 - 2 and bmg row rho - dy0 - 10 complete there
            const int rs = r0 + dy0 - 5;
            uint32_t h1p = 0, h1pp = 0;                                     // 3-wide rows rho-1, rho-2
            uint32_t h6p = 0, A[2] = {0, 0}, Bq[4] = {0, 0, 0, 0}, Cq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int s_ = 0; s_ < MT_SEG_ROWS + 15; s_++) {
                const int rho = rs + s_;
                uint32_t l = 0, cw = 0, rw = 0;
                if (rho >= 0 && rho < MT_RAW_ROWS) {
                    const uint32_t *row = raw + rho * RB_RAW_STRIDE + w;
                    l = row[0]; cw = row[1]; rw = row[2];
                }
                const uint32_t h1 = cw | (cw << 1) | (cw >> 1) | (l >> 31) | (rw << 31);
                uint32_t h6 = h1;
#pragma unroll
                for (int d = 2; d <= MT_GRAD + 1; d++) h6 |= __funnelshift_l(l, cw, d) | __funnelshift_r(cw, rw, d);
                const int rbm = rho - dy0 - 2;
                if (rbm >= r0 && rbm < r1) bm[rbm * RB_BM_STRIDE + w] = h1 | h1p | h1pp;
                h1pp = h1p; h1p = h1;
                const uint32_t a_ = h6 | h6p;                               // rows rho-1 .. rho
                const uint32_t b_ = a_ | A[s_ & 1];                         // | a(rho-2): rows rho-3 .. rho
                const uint32_t c_ = b_ | Bq[s_ & 3];                        // | b(rho-4): rows rho-7 .. rho
                const uint32_t o_ = c_ | Cq[s_ & 7];                        // | c(rho-8): rows rho-15 .. rho
                h6p = h6; A[s_ & 1] = a_; Bq[s_ & 3] = b_; Cq[s_ & 7] = c_;
                const int rg = rho - dy0 - 10;                              // dilated rows rg .. rg + 3 = raw rows rg - 6 .. rg + 9
                if (rg >= r0 && rg < r1) bmg[rg * RB_BM_STRIDE + w] = o_;
            }
        }
        if (tid < RB_BM_ROWS) { bm[tid * RB_BM_STRIDE + 16] = 0u; bmg[tid * RB_BM_STRIDE + 16] = 0u; }
        if (tid + MT_THREADS < RB_BM_ROWS) { bm[(tid + MT_THREADS) * RB_BM_STRIDE + 16] = 0u; bmg[(tid + MT_THREADS) * RB_BM_STRIDE + 16] = 0u; }
    }
    __syncthreads();
    MT_CLK(2)

    const int M = sh->M, nx = sh->nx, ny = sh->ny;
    const int nrows = 2 * ny + 1, ncols = 2 * nx + 1;
    const uint32_t colmask = ncols >= 32 ? 0xffffffffu : ((1u << ncols) - 1u);
    uint32_t *pts = raw + warp * RB_MAXB;                                   // raw is dead now: per-warp point lists
    const uint32_t lane_off = (uint32_t)((lane < nrows ? lane : 0) * RB_BM_STRIDE * 4);
    const uint32_t bm_lane = (uint32_t)__cvta_generic_to_shared(bm) + lane_off;
    const uint32_t bmg_addr = (uint32_t)__cvta_generic_to_shared(bmg);
    const int nrot = 2 * c.nk + 1, ngroups = (nrot + MT_GROUP - 1) / MT_GROUP;

    // ---- 3a0. seed the best key with MT_SEEDS rotations around the guess ---------
    // (odometry is usually close: a strong bound lets the group passes below give up
    // early).  The other warps start on the group bounds right away.
    const int seed_lo = -(MT_SEEDS / 2), seed_hi = seed_lo + MT_SEEDS - 1;
    int visited = 0;                                                        // per warp (all lanes count alike)
    // A pass that raises the best key keeps its bit-sliced counters in this warp's slot (raw is
    // dead, the point lists take its first MT_WARPS * RB_MAXB words): the warp that holds the
    // final best key has the score slice of the best rotation without another pass.
    uint32_t *slot = raw + MT_WARPS * RB_MAXB + warp * (MT_PLANES * 32);
    unsigned long long saved_key = 0ull;
    long long wclk = clock64();
    if (sh->ok && warp < MT_SEEDS && seed_lo + warp >= -c.nk && seed_lo + warp <= c.nk) {
        const int k = seed_lo + warp;
        uint32_t pl[MT_PLANES];
        mt_pass<false>(c, sh, ccx, ccy, ccf, k, nx, ny, bm_lane, pts, M, lane, pl, 0u, nullptr, &visited);
        unsigned long long key = 0ull;
        if (lane < nrows) key = mt_lane_key(pl, colmask, nx, lane - ny, k);
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            if (other > key) key = other;
        }
        unsigned long long old = 0ull;
        if (lane == 0) {
            old = atomicMax(const_cast<unsigned long long *>(&sh->best_key), key);
            atomicAdd(&sh->evals, 1);
        }
        old = __shfl_sync(0xffffffffu, old, 0);
        if (key > old) {                                                    // best so far: keep its counters for the covariance
            saved_key = key;
#pragma unroll
            for (int pb = 0; pb < MT_PLANES; pb++) slot[pb * 32 + lane] = pl[pb];
        }
    }

    MT_WCLK(10, wclk)
    if (lane == 0 && !slice_out && visited) atomicAdd(&c.stats->match_clk[13], (unsigned long long)visited);
    int visited_seed = visited;
    wclk = clock64();
    // ---- 3a. upper bound of every rotation group (shared queue) -------------------
    if (sh->ok) {
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(&sh->next_group, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= ngroups) break;
            const int kmid = min(g * MT_GROUP + MT_GROUP / 2, nrot - 1) - c.nk;
            if (g * MT_GROUP - c.nk >= seed_lo && min(g * MT_GROUP + MT_GROUP - 1, nrot - 1) - c.nk <= seed_hi) {
                if (lane == 0) { sh->group_ub[g] = -1; atomicAdd(&sh->evals, -1); }   // every member is a seed: nothing to bound
                continue;
            }
            int sc = 0;
            const bool done = mt_bound_pass(c, sh, ccx, ccy, ccf, kmid, nx, ny, bmg_addr, pts, M, lane, colmask, &sh->best_key, &visited, sc);
            if (!done) {                                                    // even the bound cannot reach the seeded best
                if (lane == 0) sh->group_ub[g] = -1;
                continue;
            }
            if (lane == 0) sh->group_ub[g] = sc;
        }
    }
    MT_WCLK(11, wclk)
    if (lane == 0 && !slice_out) atomicAdd(&c.stats->match_clk[14], (unsigned long long)(visited - visited_seed));
    visited_seed = visited;
    __syncthreads();
    MT_CLK(4)
    if (warp == 0 && sh->ok) {                                              // rank the groups by decreasing bound
        for (int me = lane; me < ngroups; me += 32) {
            const int mine = sh->group_ub[me];
            int rank = 0;
            for (int g = 0; g < ngroups; g++) {
                const int o = sh->group_ub[g];
                rank += (o > mine) || (o == mine && g < me);
            }
            sh->group_order[rank] = me;
        }
    }
    __syncthreads();
    MT_CLK(5)
    wclk = clock64();

    // ---- 3b. score member rotations while their group's bound can still win -----
    if (sh->ok) {
        volatile unsigned long long *vbest = &sh->best_key;
        for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(&sh->next_item, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= ngroups * MT_GROUP) break;
            const int g = sh->group_order[item / MT_GROUP];
            const int ridx = g * MT_GROUP + item % MT_GROUP;
            if (ridx >= nrot) continue;
            const int k = ridx - c.nk, ub = sh->group_ub[g];
            if (k >= seed_lo && k <= seed_hi) continue;                     // scored as a seed
            const unsigned long long cur = *vbest;
            if (ub < (int)(cur >> 32)) break;                               // groups are sorted: nothing left can win
            if (mt_key(ub, 0, 0, k) < cur) continue;                        // this rotation cannot beat the best key
            uint32_t pl[MT_PLANES];
            const bool done = mt_pass<true>(c, sh, ccx, ccy, ccf, k, nx, ny, bm_lane, pts, M, lane, pl, lane < nrows ? colmask : 0u, vbest,
                                            &visited);
            if (lane == 0) atomicAdd(&sh->evals, 1);
            if (!done) continue;                                            // cannot reach the best score any more
            unsigned long long key = 0ull;
            if (lane < nrows) key = mt_lane_key(pl, colmask, nx, lane - ny, k);
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                if (other > key) key = other;
            }
            unsigned long long old = 0ull;
            if (lane == 0) old = atomicMax(const_cast<unsigned long long *>(&sh->best_key), key);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (key > old) {
                saved_key = key;
#pragma unroll
                for (int pb = 0; pb < MT_PLANES; pb++) slot[pb * 32 + lane] = pl[pb];
            }
        }
    }
    MT_WCLK(12, wclk)
    if (lane == 0 && !slice_out) atomicAdd(&c.stats->match_clk[15], (unsigned long long)(visited - visited_seed));
    __syncthreads();
    MT_CLK(6)
    if (lane == 0 && visited) atomicAdd(&sh->visits, visited);
    int bs, bi, bj, bk;
    mt_key_decode(sh->best_key, bs, bi, bj, bk);
    if (!sh->ok) { bs = 0; bi = bj = bk = 0; }

    const double step = c.rot_step;
    const bool valid = sh->ok && fabs((double)bi * RB_CS) < sh->rx && fabs((double)bj * RB_CS) < sh->ry &&
                       fabs((double)bk * step) < 3.14159265358979323846 / 6.0 && (bi != 0 || bj != 0 || bk != 0);

    // ---- 5. covariance --------------------------------------------------------
    if (valid || slice_out) {
        if (sh->ok && saved_key == sh->best_key && lane == 0) sh->slot_warp = warp;    // this warp holds the slice of the best rotation
        if (sh->ok) {                                                       // rotation line at the best translation (all warps)
            long long T0 = 0, T1 = 0, T2 = 0;
            const int klo = max(bk - MT_ROT_LINE_HALF, -c.nk), khi = min(bk + MT_ROT_LINE_HALF, c.nk);
            // A rotation whose score ends more than 40 below the best has weight zero: with misses counted as the points go by
            // (far points first, they miss first) such a rotation is given up as soon as bs - (M - misses) > 40 is certain.
            const int miss_max = 40 - (bs - M);                             // misses a rotation can afford
            int ovf = 0;
            for (int k = klo + warp; k <= khi; k += MT_WARPS) {
                const MtRot r = mt_rot(c, sh, k, bi, bj);
                int s = 0, seen = 0;
                for (int q0 = 0; q0 < M; q0 += 32) {
                    const int q = q0 + lane;
                    bool hit = false;
                    if (q < M) {
                        const uint32_t pk = mt_raster_point(c, sh, ccx, ccy, ccf, r, q, 0, 0, ovf);
                        hit = (bm[pk >> MT_PT_WORD_SHIFT] >> (pk & 31)) & 1u;
                    }
                    s += __popc(__ballot_sync(0xffffffffu, hit));
                    seen = min(q0 + 32, M);
                    if (seen - s > miss_max) break;
                }
                if (seen < M) continue;
                const int d = bs - s;
                if (d <= 40) {
                    const long long w = 1ll << (40 - d);
                    T0 += w; T1 += w * k; T2 += w * k * k;
                }
            }
            if (ovf) sh->overflow = 1;
            if (lane == 0) { sh->mom_part[warp][6] = T0; sh->mom_part[warp][7] = T1; sh->mom_part[warp][8] = T2; }
        } else if (lane == 0) {
            sh->mom_part[warp][6] = sh->mom_part[warp][7] = sh->mom_part[warp][8] = 0;
        }
        __syncthreads();
        if (sh->ok && sh->slot_warp >= 0) {                                 // translation slice at the best rotation: every thread decodes a few cells
            const uint32_t *best_slot = raw + MT_WARPS * RB_MAXB + sh->slot_warp * (MT_PLANES * 32);
            long long mo[6] = {0, 0, 0, 0, 0, 0};
            for (int cell = tid; cell < nrows * ncols; cell += MT_THREADS) {
                const int jr = cell / ncols, b = cell - jr * ncols;
                int sc = 0;
#pragma unroll
                for (int pb = 0; pb < MT_PLANES; pb++) sc |= (int)((best_slot[pb * 32 + jr] >> b) & 1u) << pb;
                const int i = b - nx, j = jr - ny, d = bs - sc;
                if (slice_out) slice_out[(j + RB_NT_MAX) * RB_SLICE_W + (i + RB_NT_MAX)] = sc;
                if (d > 40) continue;
                const long long w = 1ll << (40 - d);
                mo[0] += w; mo[1] += w * i; mo[2] += w * j; mo[3] += w * (i * i); mo[4] += w * (j * j); mo[5] += w * (i * j);
            }
#pragma unroll
            for (int e = 0; e < 6; e++) {
                for (int o = 16; o > 0; o >>= 1) mo[e] += __shfl_xor_sync(0xffffffffu, mo[e], o);
                if (lane == 0) sh->mom_part[warp][e] = mo[e];
            }
        } else if (lane == 0) {
#pragma unroll
            for (int e = 0; e < 6; e++) sh->mom_part[warp][e] = 0;
        }
        __syncthreads();
        if (tid < 9) {
            long long t = 0;
#pragma unroll
            for (int wq = 0; wq < MT_WARPS; wq++) t += sh->mom_part[wq][tid];
            sh->mom[tid] = t;
        }
    }
    __syncthreads();
    MT_CLK(7)

    // ---- 5b. NDT refinement, matchScanCustom.m:32-50 ---------------------------------
    static_assert(MT_THREADS >= RB_MAXB && MT_WARPS % 2 == 0, "NDT stage: one thread per beam, two halves of warps");
    double nd_p[3] = {(double)bi, (double)bj, (double)bk * step}, nd_S = 0.0;
    int nd_evals = 0;
    bool nd_accept = false;
    if (c.refine && valid) {
        double qx = 0.0, qy = 0.0;
        double ax_, ay_;
        const bool has = has_beam && (mt_curr_point(c, sh, exists, b_d, b_px, b_py, adj, qx, qy, ax_, ay_) & 2);
        double *red = reinterpret_cast<double *>(bmg), *ctl = red + MT_WARPS * NDT_TERMS;   // bmg is dead after phase A
        ndt_refine(bm, red, ctl, sh->g0xu - sh->x0, sh->fx, sh->fy, has, qx, qy, nd_p[0], nd_p[1], nd_p[2]);
        nd_p[0] = ctl[1]; nd_p[1] = ctl[2]; nd_p[2] = ctl[3]; nd_S = ctl[4]; nd_evals = (int)ctl[5];
        const bool ok = fabs(nd_p[0] * RB_CS) < sh->rx && fabs(nd_p[1] * RB_CS) < sh->ry &&
                        fabs(nd_p[2]) < 3.14159265358979323846 / 6.0 && (nd_p[0] != 0.0 || nd_p[1] != 0.0 || nd_p[2] != 0.0);
        nd_accept = ok && nd_S * 2.0 > (double)bs;                          // :38-39
    }

    // ---- 6. result --------------------------------------------------------------
    MT_CLK(8)
    if (tid == 0) {
        double *op = c.m_pose + 3 * (size_t)p, *oc = c.m_cov + 9 * (size_t)p;
        op[0] = sh->gx + (double)bi * RB_CS;                                // hybridmap.py:253-255
        op[1] = sh->gy + (double)bj * RB_CS;
        op[2] = sh->gth + (double)bk * step;
        int *ob = c.m_best + 4 * (size_t)p;
        ob[0] = bi; ob[1] = bj; ob[2] = bk; ob[3] = M;
        c.m_refine[2 * (size_t)p] = nd_evals;
        c.m_refine[2 * (size_t)p + 1] = nd_accept ? 1 : 0;
        if (nd_evals && !slice_out) {
            atomicAdd(&c.stats->ndt_evals, (unsigned long long)nd_evals);
            atomicAdd(&c.stats->ndt_accepted, nd_accept ? 1ull : 0ull);
        }
        c.m_valid[p] = valid ? 1 : 0;
        if (sh->overflow) atomicExch(&c.flags->world_overflow, 1);
        if (!slice_out) {
            atomicAdd(&c.stats->match_evals, (unsigned long long)(sh->evals + ngroups));
            atomicAdd(&c.stats->match_visits, (unsigned long long)sh->visits);
            atomicAdd(&c.stats->match_points, (unsigned long long)M);
            atomicAdd(&c.stats->match_runs, 1ull);
        }
        if (!valid) {                                                       // matchScanCustom.m:26-28
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            for (int q = 0; q < 9; q++) oc[q] = nan;
            c.m_score[p] = 0.0;
            if (!slice_out) {
                atomicAdd(&c.stats->match_failed, 1ull);
                if (sh->ok && bi == 0 && bj == 0 && bk == 0) atomicAdd(&c.stats->match_failed_zero, 1ull);
            }
        } else {
            const double W0 = (double)sh->mom[0], T0 = (double)sh->mom[6];
            const double mx = (double)sh->mom[1] / W0, my = (double)sh->mom[2] / W0, mt = (double)sh->mom[7] / T0;
            const double q = RB_CS * RB_CS, qt = step * step;
            for (int e = 0; e < 9; e++) oc[e] = 0.0;
            oc[0] = ((double)sh->mom[3] / W0 - mx * mx) * q + q / 12.0;
            oc[4] = ((double)sh->mom[4] / W0 - my * my) * q + q / 12.0;
            oc[1] = oc[3] = ((double)sh->mom[5] / W0 - mx * my) * q;
            oc[8] = ((double)sh->mom[8] / T0 - mt * mt) * qt + qt / 12.0;
            c.m_score[p] = (double)bs;
            if (nd_accept) {                                                // :40-41, covariance stays the grid stage's
                op[0] = sh->gx + nd_p[0] * RB_CS;
                op[1] = sh->gy + nd_p[1] * RB_CS;
                op[2] = sh->gth + nd_p[2];
                c.m_score[p] = nd_S;
            }
        }
    }
}

// The opt-in to > 48 KB of dynamic shared memory is per device: set it on every
// launch (a host-side call of about a microsecond) rather than caching a flag that
// would be wrong for a second device in the same process.
static void match_set_attr()
{
    cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rb_match_smem_bytes());
}

// Duplicates of the last resample take their representative's result (the guess is
// identical, so guess + correction is too).
__global__ void __launch_bounds__(256) match_copy_dups_kernel(RbCtx c)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= c.N) return;
    const int r = c.dup_of[p];
    if (r == p) return;
    for (int q = 0; q < 3; q++) c.m_pose[3 * (size_t)p + q] = c.m_pose[3 * (size_t)r + q];
    for (int q = 0; q < 9; q++) c.m_cov[9 * (size_t)p + q] = c.m_cov[9 * (size_t)r + q];
    for (int q = 0; q < 4; q++) c.m_best[4 * (size_t)p + q] = c.m_best[4 * (size_t)r + q];
    for (int q = 0; q < 2; q++) c.m_refine[2 * (size_t)p + q] = c.m_refine[2 * (size_t)r + q];
    c.m_score[p] = c.m_score[r];
    c.m_valid[p] = c.m_valid[r];
    if (!c.m_valid[r]) {
        atomicAdd(&c.stats->match_failed, 1ull);
        const int *ob = c.m_best + 4 * (size_t)r;
        if (ob[0] == 0 && ob[1] == 0 && ob[2] == 0) atomicAdd(&c.stats->match_failed_zero, 1ull);
    }
}

void rb_launch_match(const RbCtx &c, int adj, cudaStream_t s, int sel, bool copy_dups)
{
    match_set_attr();
    match_kernel<<<c.N, MT_THREADS, rb_match_smem_bytes(), s>>>(c, 0, nullptr, adj, sel);
    if (c.use_dup && copy_dups) match_copy_dups_kernel<<<(c.N + 255) / 256, 256, 0, s>>>(c);
}

// Debug/test entry: re-run the last match (same mode) for one particle and dump the
// score slice at its best rotation (29x29 int32).  Overwrites that particle's match
// outputs with identical values.
void rb_launch_match_slice(const RbCtx &c, int particle, int *slice_dev, int adj, cudaStream_t s)
{
    match_set_attr();
    match_kernel<<<1, MT_THREADS, rb_match_smem_bytes(), s>>>(c, particle, slice_dev, adj, 0);
}
