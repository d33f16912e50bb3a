// k_weight.cu -- stages 1b + 3: proposal sampling around the matcher result,
// observation-likelihood weighting and moment normalisation, fused; plus the
// NaN-covariance fallback.
//
// Reference: Robot.map_update robot.py:73-114, Robot._generate_sample_weight
// robot.py:118-139, HybridMap.get_odds_at hybridmap.py:85-93.
//
// One warp per particle, one lane per proposal sample (K <= 32).  Each lane
// walks the B beams for its own sample; the K samples of a beam land in the same
// or neighbouring cells, so a warp-wide lookup touches a handful of sectors.
// Log-odds are summed as integer tenths (exact), the K-term moments are reduced
// in the reference's sequential order so that results are run-to-run identical.
#include "common.cuh"

#ifdef WT_LIBM   // experiment only: CUDA libm instead of the shared IEEE routines (breaks bit parity with the oracle)
#define rb_sincos(a, s, c) sincos(a, s, c)
#define rb_exp(x) exp(x)
#endif

#ifndef WT_WARPS
#define WT_WARPS 4
#endif
#ifndef WT_INFLIGHT
#define WT_INFLIGHT 4                   // beam lookups in flight per lane
#endif
#ifndef WT_F32_EPS
#define WT_F32_EPS 2.5e-4f              // float32 cell lookups: distance from a lattice boundary below which float64 decides (error bound 1.7e-4)
#endif

// The float64 cell lookup for the lanes the float32 evaluation cannot decide (out of line: rare, and its
// registers stay out of the beam loop): floor(20 g) from two fused multiply-adds, and the reference's own
// expressions (rb_xform + rb_read_axis) within 1e-6 of a lattice boundary.  Returns (page-table slot or -1, offset).
__device__ __noinline__ int2 wt_locate_f64(double bpx, double bpy, double cs_, double sn_, double g0, double g1, int txh, int tyh,
                                           int subs_x, int ux_max, int uy_max)
{
    const double c20 = cs_ * 20.0, s20 = sn_ * 20.0, x20 = g0 * 20.0, y20 = g1 * 20.0;
    const double vx = fma(c20, bpx, fma(-s20, bpy, x20)), vy = fma(s20, bpx, fma(c20, bpy, y20));
    const int kx = __double2int_rd(vx), ky = __double2int_rd(vy);
    const double dx = vx - (double)kx, dy = vy - (double)ky;
    int ux, uy;
    if (dx > 1e-6 && dx < 1.0 - 1e-6 && dy > 1e-6 && dy < 1.0 - 1e-6 && fabs(vx) < 1e6 && fabs(vy) < 1e6) {
        ux = kx + 400 + 800 * txh; uy = ky + 400 + 800 * tyh;
    } else {
        double gx, gy;
        rb_xform(cs_, sn_, g0, g1, bpx, bpy, gx, gy);
        int tx, ty, ix, iy;
        rb_read_axis(gx, tx, ix);
        rb_read_axis(gy, ty, iy);
        if (tx < -txh || tx > txh || ty < -tyh || ty > tyh) return make_int2(-1, 0);
        ux = 800 * (tx + txh) + ix; uy = 800 * (ty + tyh) + iy;
    }
    if ((unsigned)ux >= (unsigned)ux_max || (unsigned)uy >= (unsigned)uy_max) return make_int2(-1, 0);
    return make_int2((uy / RB_SUB) * subs_x + ux / RB_SUB, RB_OFF_Y(uy % RB_SUB) + RB_OFF_X(ux % RB_SUB));
}

// fallback_phase == 0: particles with a valid match (robot.py:80-114)
// fallback_phase == 1: particles whose match failed, after the map update
//                      (robot.py:73-78: weight += 1 + sum of log-odds at the odometry pose)
// 8 resident CTAs of 4 warps: 64 registers per thread (66 without the bound, 7 CTAs); measured 0.96 -> 0.89 ms at
// 16,384 particles.  Higher bounds spill and lose.
#ifndef WT_MINBLOCKS
#define WT_MINBLOCKS 8
#endif
// z: N*K*3 standard normals (our transform mean + chol(cov) z), or guesses: N*K*3 proposal samples drawn by
// the caller exactly as the reference draws them (np.random.multivariate_normal, robot.py:81), or neither
// (Philox normals on the device).
__global__ void __launch_bounds__(WT_WARPS * 32, WT_MINBLOCKS) weight_kernel(RbCtx c, const double *__restrict__ z,
                                                                             const double *__restrict__ guesses, int fallback_phase)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * WT_WARPS + warp;
    if (p >= c.N) return;
    const int valid = c.m_valid[p];
    const unsigned FULL = 0xffffffffu;

    if (fallback_phase) {
        if (valid) return;
        const double *pose = c.pose + 3 * (size_t)p;
        double x = pose[0], y = pose[1], cs_, sn_;
        rb_sincos(pose[2], &sn_, &cs_);
        int S = 0;
        for (int j = lane; j < c.B; j += 32) {
            double d = c.dist[j];
            if (d < RB_W_MAX_R && d > RB_W_MIN_R) {
                double gx, gy;
                rb_xform(cs_, sn_, x, y, c.px[j], c.py[j], gx, gy);
                S += rb_odds_tenths(c, p, gx, gy);
            }
        }
        for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(FULL, S, o);
        if (lane == 0) c.weight[p] = (double)(10 + S) / 10.0 + c.weight[p];      // robot.py:76-77
        return;
    }
    if (!valid) return;

    const int K = c.K;
    const double *mp = c.m_pose + 3 * (size_t)p, *mc = c.m_cov + 9 * (size_t)p;
    const double m0 = mp[0], m1 = mp[1], m2 = mp[2];
    // lower Cholesky factor of the matcher covariance (our sampling transform,
    // stands in for np.random.multivariate_normal robot.py:81)
    double l00 = sqrt(mc[0]);
    double l10 = mc[3] / l00, l20 = mc[6] / l00;
    double l11 = sqrt(mc[4] - l10 * l10);
    double l21 = (mc[7] - l20 * l10) / l11;
    double l22 = sqrt((mc[8] - l20 * l20) - l21 * l21);
    const double TWO_PI = 6.283185307179586476925286766559;
    double nrm = TWO_PI * sqrt(TWO_PI) * ((l00 * l11) * l22);

    double g0 = m0, g1 = m1, g2 = m2, w = 0.0;
    if (lane < K) {
        double z0 = 0.0, z1 = 0.0, z2 = 0.0;
        if (guesses) {
        } else if (z) {
            const double *zz = z + ((size_t)p * K + lane) * 3;
            z0 = zz[0]; z1 = zz[1]; z2 = zz[2];
        } else {
            uint32_t r[4], q[4];
            unsigned long long gp = (unsigned long long)c.rank * (unsigned long long)c.N + (unsigned long long)p;
            rb_philox((uint32_t)gp, (uint32_t)(gp >> 32) ^ ((uint32_t)lane << 8), (uint32_t)c.step_no, 0x57u, c.seed, r);
            rb_philox((uint32_t)gp, (uint32_t)(gp >> 32) ^ ((uint32_t)lane << 8), (uint32_t)c.step_no, 0x58u, c.seed, q);
            double u1 = rb_u01(r[0], r[1]), u2 = rb_u01(r[2], r[3]);
            double u3 = rb_u01(q[0], q[1]), u4 = rb_u01(q[2], q[3]);
            double ra = sqrt(-2.0 * log(u1)), rb2 = sqrt(-2.0 * log(u3));
            z0 = ra * cos(TWO_PI * u2);
            z1 = ra * sin(TWO_PI * u2);
            z2 = rb2 * cos(TWO_PI * u4);
        }
        if (guesses) {
            const double *gg = guesses + ((size_t)p * K + lane) * 3;
            g0 = gg[0]; g1 = gg[1]; g2 = gg[2];
        } else {
            g0 = m0 + l00 * z0;
            g1 = m1 + (l10 * z0 + l11 * z1);
            g2 = m2 + ((l20 * z0 + l21 * z1) + l22 * z2);
        }
        double d0 = g0 - m0, d1 = g1 - m1, d2 = g2 - m2;
        double y0 = d0 / l00;
        double y1 = (d1 - l10 * y0) / l11;
        double y2 = ((d2 - l20 * y0) - l21 * y1) / l22;
        double maha = (y0 * y0 + y1 * y1) + y2 * y2;
        double pr = rb_exp(-0.5 * maha) / nrm * 10.0;                            // robot.py:87
        // observation weight of this sample, robot.py:118-139
        double cs_, sn_;
        rb_sincos(g2, &sn_, &cs_);
        int S = 0;
        const uint32_t *pt = c.pt + (size_t)p * c.nsub;
        // WT_INFLIGHT beams in flight: locate (ALU) -> page-table entries -> cells.
        // The cell of a beam end is floor(V) per axis with V = 20 g = x20 + c20 px - s20 py (rb_locate_fast: within 1e-11
        // of 20 x the reference's own float64 chain, same floor unless V lies within 1e-6 of an integer).  V is
        // evaluated in float32 RELATIVE to floor(x20): |c20 px| <= 500 cells (ranges below 25 m), so the float32 value
        // is within 1.7e-4 of V - floor(x20) (conversions of c20, s20, px, py: 3e-5 each; the two fused roundings 1.5e-5
        // and 3e-5) and its floor is right unless it lies within WT_F32_EPS of an integer; those lanes (0.1 %) take the
        // float64 expression, and there the reference's own expressions within 1e-6 of a lattice boundary
        // (rb_xform + rb_locate).  (Absolute float32 coordinates need 4e-3 and diverge in 38 % of the iterations.)
        const double x20 = g0 * 20.0, y20 = g1 * 20.0;
        const int offx = 400 + 800 * c.txh, offy = 400 + 800 * c.tyh;
        const double X0d = floor(x20), Y0d = floor(y20);
        const bool f32_ok = fabs(x20) < 1e6 && fabs(y20) < 1e6;
        const int X0 = f32_ok ? (int)X0d + offx : 0, Y0 = f32_ok ? (int)Y0d + offy : 0;
        const float xr = (float)(x20 - X0d), yr = (float)(y20 - Y0d), cf = (float)(cs_ * 20.0), sf = (float)(sn_ * 20.0);
        for (int j = 0; j < c.B; j += WT_INFLIGHT) {
            int sub[WT_INFLIGHT], off[WT_INFLIGHT];
#pragma unroll
            for (int u = 0; u < WT_INFLIGHT; u++) {
                const int jj = j + u;
                sub[u] = -1; off[u] = 0;
                if (jj < c.B) {
                    const float4 bf = __ldg(&c.beamf[jj]);
                    if (bf.z != 0.0f) {                                             // robot.py:130 range gate
                        const float vxf = fmaf(cf, bf.x, fmaf(-sf, bf.y, xr)), vyf = fmaf(sf, bf.x, fmaf(cf, bf.y, yr));
                        const float flx = floorf(vxf), fly = floorf(vyf);
                        const float dxf = vxf - flx, dyf = vyf - fly;
                        if (f32_ok && dxf > WT_F32_EPS && dxf < 1.0f - WT_F32_EPS && dyf > WT_F32_EPS && dyf < 1.0f - WT_F32_EPS) {
                            const int ux = (int)flx + X0, uy = (int)fly + Y0;
                            if ((unsigned)ux < (unsigned)c.ux_max && (unsigned)uy < (unsigned)c.uy_max) {
#ifdef WT_NO_RLUT
                                sub[u] = (uy / RB_SUB) * c.subs_x + ux / RB_SUB;
                                off[u] = RB_OFF_Y(uy % RB_SUB) + RB_OFF_X(ux % RB_SUB);
#else
                                // page-table slot | offset << 12 from two table entries (L1 hits) instead of two divisions by 160,
                                // two remainders and the block layout arithmetic: the kernel is issue-bound, its LSU pipe is idle
                                const uint32_t so = __ldg(&c.rlut[ux]) + __ldg(&c.rlut[c.ux_max + uy]);
                                sub[u] = (int)(so & 0xfffu);
                                off[u] = (int)(so >> 12);
#endif
                            }
                        } else {
                            const int2 so = wt_locate_f64(c.px[jj], c.py[jj], cs_, sn_, g0, g1, c.txh, c.tyh, c.subs_x, c.ux_max, c.uy_max);
                            sub[u] = so.x; off[u] = so.y;
                        }
                    }
                }
            }
            uint32_t t[WT_INFLIGHT];
#pragma unroll
            for (int u = 0; u < WT_INFLIGHT; u++) t[u] = sub[u] >= 0 ? pt[sub[u]] : RB_NONE;
#pragma unroll
            for (int u = 0; u < WT_INFLIGHT; u++)
                if (t[u] != RB_NONE) S += c.pool[(size_t)t[u] * RB_SUB_BYTES + off[u]];
        }
        w = ((double)(10 + S) / 10.0) * pr;
    }
    // min over the K samples (robot.py:89), order-independent
    double mn = lane < K ? w : __longlong_as_double(0x7ff0000000000000ll);
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL, mn, o));
    double wk_mine = (w - mn) + 1e-2;                                          // robot.py:90
    // sequential moments, robot.py:92-108 (every lane runs the same chain)
    double mean0 = 0.0, mean1 = 0.0, mean2 = 0.0, norm = 0.0;
    for (int k = 0; k < K; k++) {
        double wk = __shfl_sync(FULL, wk_mine, k);
        double a0 = __shfl_sync(FULL, g0, k), a1 = __shfl_sync(FULL, g1, k), a2 = __shfl_sync(FULL, g2, k);
        mean0 = mean0 + a0 * wk;
        mean1 = mean1 + a1 * wk;
        mean2 = mean2 + a2 * wk;
        norm = norm + wk;
    }
    mean0 = mean0 / norm; mean1 = mean1 / norm; mean2 = mean2 / norm;
    double s00 = 0, s01 = 0, s02 = 0, s10 = 0, s11 = 0, s12 = 0, s20 = 0, s21 = 0, s22 = 0;
    for (int k = 0; k < K; k++) {
        double wk = __shfl_sync(FULL, wk_mine, k);
        double d0 = __shfl_sync(FULL, g0, k) + (-mean0), d1 = __shfl_sync(FULL, g1, k) + (-mean1),
               d2 = __shfl_sync(FULL, g2, k) + (-mean2);
        s00 = s00 + (d0 * d0) * wk; s01 = s01 + (d0 * d1) * wk; s02 = s02 + (d0 * d2) * wk;
        s10 = s10 + (d1 * d0) * wk; s11 = s11 + (d1 * d1) * wk; s12 = s12 + (d1 * d2) * wk;
        s20 = s20 + (d2 * d0) * wk; s21 = s21 + (d2 * d1) * wk; s22 = s22 + (d2 * d2) * wk;
    }
    if (lane == 0) {
        double *pose = c.pose + 3 * (size_t)p, *cov = c.cov + 9 * (size_t)p;
        pose[0] = mean0; pose[1] = mean1; pose[2] = mean2;                      // robot.py:109-113
        cov[0] = s00 / norm; cov[1] = s01 / norm; cov[2] = s02 / norm;
        cov[3] = s10 / norm; cov[4] = s11 / norm; cov[5] = s12 / norm;
        cov[6] = s20 / norm; cov[7] = s21 / norm; cov[8] = s22 / norm;
        double total = norm + mn * (double)K;                                   // robot.py:108
        c.weight[p] = total + c.weight[p];                                      // robot.py:114
    }
}

void rb_launch_weight(const RbCtx &c, const double *z_dev, const double *guesses_dev, int fallback_phase, cudaStream_t s)
{
    int blocks = (c.N + WT_WARPS - 1) / WT_WARPS;
    weight_kernel<<<blocks, WT_WARPS * 32, 0, s>>>(c, z_dev, guesses_dev, fallback_phase);
}
