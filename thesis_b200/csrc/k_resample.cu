// k_resample.cu -- stage 5: systematic resampling and the copy-on-write
// duplication that replaces the reference's deep copies.
//
// Reference: main.resample main.py:46-79 (weights pinned to float64, SURVEY
// 3.4-7), Robot.copy robot.py:141-149, HybridMap.copy hybridmap.py:315-320.
//
// resample_plan   one CTA over the (all-gathered) weights of ALL ranks:
//                   min/max and trigger (main.py:50), -inf -> 0 (:53), additive
//                   shift of the non-zero entries (:54-55), the running sum with
//                   the reference's left-to-right float64 roundings (:57,:61-62;
//                   float64 addition is not associative and the ancestors must be
//                   bit-exact: an exact integer prefix sum while the sum stays in
//                   one binade, a single-thread chain otherwise), then ancestors
//                   in parallel.
//                 Every rank runs it on identical input and gets identical output.
// resample_gather builds the new particle slots of this rank from the ancestor
//                 vector: pose, covariance, weight <- 1.0 (main.py:77-78),
//                 page table and tile-existence mask.  Slots whose ancestor
//                 lives on another rank are left for the migration step.
// resample_refs   fixes sub-tile reference counts: an old particle with m local
//                 descendants contributes m - 1 (or releases its tiles if m = 0).
#include "common.cuh"

#ifndef RS_THREADS
#define RS_THREADS 1024                 // one CTA: the running sum is a chain of windows, every window one block scan
#endif
#ifndef RS_EPT
#define RS_EPT 4                        // consecutive elements per thread
#endif
#define RS_CHUNK (RS_THREADS * RS_EPT)  // 4,096 elements per window
#define RS_WARPS (RS_THREADS / 32)
#define RS_SEQ 64                       // elements of a sequential stretch (start of the sum, non-finite or denormal carries)
static_assert(RS_WARPS <= 32, "the block scan folds one warp total per lane");

// The running sum c_i = fl(c_{i-1} + v_i) (main.py:57 == :62), float64, round to nearest even, is inherently a
// chain of N dependent adds -- except that while the sum stays inside one binade [2^k, 2^(k+1)) every add rounds to a
// multiple of u = 2^(k-52).  With S = c/u (an integer in [2^52, 2^53)) and v = (m + f) u, m integer, 0 <= f < 1:
//     S' = S + m + [f > 1/2]          (f == 1/2, a tie, depends on the parity of S + m)
// so a stretch without ties that does not leave the binade is an exact INTEGER prefix sum.  A window of RS_CHUNK
// elements is scanned in parallel in units of the carry's u; the first element that is a tie, is not a positive
// normal number, or takes the sum to 2^53 u or beyond ends the window: everything before it is final, the element
// itself is added with one genuine float64 add by the thread that owns it, and the next window starts behind it
// in the (possibly new) binade.  The sum of N positive weights crosses about log2(N) binades, almost all of them
// within the first few elements, which a single thread chains (RS_SEQ elements); 65,536 weights take about 16 + 12
// windows instead of 65,536 dependent adds -- with every rounding of the reference reproduced.
__global__ void __launch_bounds__(RS_THREADS) resample_plan_kernel(RbCtx c, const double *__restrict__ w_in,
                                                                   const double *__restrict__ u01_in)
{
    __shared__ double red_mx[RS_WARPS], red_mn[RS_WARPS], red_mn2[RS_WARPS];
    __shared__ long long scan_tot[RS_WARPS];
    __shared__ int s_stop[2];
    __shared__ double s_mn2, s_carry;
    __shared__ int s_do;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NG = c.n_global;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double *w = c.w_all;                                                     // the running sum

    // max / min of the raw weights (main.py:50) and min after -inf -> 0 (:53-54), one pass
    double mx = -INF, mn = INF, mn2 = INF;
    for (int i = tid; i < NG; i += RS_THREADS) {
        const double v = w_in[i];
        mx = fmax(mx, v); mn = fmin(mn, v); mn2 = fmin(mn2, v == -INF ? 0.0 : v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mn2 = fmin(mn2, __shfl_xor_sync(0xffffffffu, mn2, o));
    }
    if (lane == 0) { red_mx[warp] = mx; red_mn[warp] = mn; red_mn2[warp] = mn2; }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < RS_WARPS; k++) { mx = fmax(mx, red_mx[k]); mn = fmin(mn, red_mn[k]); mn2 = fmin(mn2, red_mn2[k]); }
        s_do = (mx - mn > RB_RESAMPLE_TRIGGER) ? 1 : 0;
        if (c.flags->pool_exhausted) s_do = 0;                               // maps missed a scan: freeze the set until the caller reacts
        c.flags->did_resample = s_do;
        c.flags->resample_error = 0;
        c.flags->remote_needed = 0;
        if (s_do) c.stats->resamples += 1ull;
        s_mn2 = mn2;
        s_carry = 0.0;
        s_stop[0] = s_stop[1] = 0x7fffffff;
    }
    __syncthreads();
    if (!s_do) return;                                                       // particles unchanged (identity ancestors)
    const double shift = s_mn2 < 0.0 ? fabs(s_mn2) : 0.0;
    const bool do_shift = s_mn2 < 0.0;
    const unsigned long long MANT = (1ull << 52) - 1ull;
    auto load = [&](int i) {
        double v = 0.0;
        if (i < NG) {
            v = w_in[i];
            if (v == -INF) v = 0.0;                                          // main.py:53
            if (do_shift && v != 0.0) v += shift;                            // main.py:55
        }
        return v;
    };
    int pos = 0, pre_pos = -1, it = 0;
    double carry = 0.0;                                                      // uniform over the block
    double v_next[RS_EPT];
    while (pos < NG) {
        const unsigned long long cb = (unsigned long long)__double_as_longlong(carry);
        const int kexp = (int)((cb >> 52) & 0x7ffull);
        const bool fast_ok = !(cb >> 63) && kexp >= 54 && kexp < 0x7ff;      // carry > 0, normal, u normal
        if (pos == 0 || !fast_ok) {
            // sequential stretch by one thread: the start of the sum (carry 0, a binade crossing every few elements) and
            // carries the integer scan cannot express (non-finite, denormal, negative: pathological input)
            const int n = min(RS_SEQ, NG - pos);
            if (tid == 0) {
                double cur = carry;
                for (int e0 = 0; e0 < n; e0 += 8) {
                    double x[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) x[e] = load(pos + e0 + e);
#pragma unroll
                    for (int e = 0; e < 8; e++) { cur += x[e]; if (e0 + e < n) { w[pos + e0 + e] = cur; s_carry = cur; } }
                }
            }
            __syncthreads();
            carry = s_carry;
            pos += n;
            __syncthreads();
            continue;
        }
        const int i0 = pos + RS_EPT * tid;
        double v[RS_EPT];
        if (pre_pos == pos) {
#pragma unroll
            for (int e = 0; e < RS_EPT; e++) v[e] = v_next[e];
        } else {
#pragma unroll
            for (int e = 0; e < RS_EPT; e++) v[e] = load(i0 + e);
        }
        pre_pos = pos + RS_CHUNK;                                            // the usual next window is on its way
#pragma unroll
        for (int e = 0; e < RS_EPT; e++) v_next[e] = load(i0 + RS_CHUNK + e);
        long long a[RS_EPT];
        int first_bad = RS_EPT;                                              // first element of this thread that needs a real add
#pragma unroll
        for (int e = RS_EPT - 1; e >= 0; e--) {
            a[e] = 0;
            if (i0 + e < NG && v[e] != 0.0) {
                const unsigned long long vb = (unsigned long long)__double_as_longlong(v[e]);
                const int ve = (int)((vb >> 52) & 0x7ffull);
                if ((vb >> 63) || ve == 0x7ff || ve == 0) first_bad = e;     // negative, inf / nan, denormal
                else {
                    const unsigned long long M = (vb & MANT) | (1ull << 52);
                    const int sft = kexp - ve;
                    if (sft < 1) first_bad = e;                              // v >= 2^k: the sum leaves the binade
                    else if (sft <= 54) {
                        const unsigned long long r = M & ((1ull << sft) - 1ull), half = 1ull << (sft - 1);
                        a[e] = (long long)(M >> sft);
                        if (r > half) a[e] += 1;
                        else if (r == half) first_bad = e;                   // tie
                    }                                                        // sft > 54: v < u/4, rounds away
                }
            }
        }
        // inclusive integer scan: inside the thread, across the warp, across the 32 warps (every warp folds the totals itself)
#pragma unroll
        for (int e = 1; e < RS_EPT; e++) a[e] += a[e - 1];
        long long pw = a[RS_EPT - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long q = __shfl_up_sync(0xffffffffu, pw, o);
            if (lane >= o) pw += q;
        }
        if (lane == 31) scan_tot[warp] = pw;
        if (tid == 0) s_stop[(it + 1) & 1] = 0x7fffffff;                     // the other buffer: its readers passed the last barrier
        __syncthreads();
        long long tw = lane < RS_WARPS ? scan_tot[lane] : 0, ti = tw;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long q = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += q;
        }
        const long long total = __shfl_sync(0xffffffffu, ti, RS_WARPS - 1);
        const long long before = __shfl_sync(0xffffffffu, ti - tw, warp) + (pw - a[RS_EPT - 1]);   // exclusive prefix of this thread
        const unsigned long long S_in = (cb & MANT) | (1ull << 52);
        // first element that ends the window: needs a real add, or takes the sum out of the binade (the prefix is non-decreasing)
        int stop_e = first_bad;
#pragma unroll
        for (int e = RS_EPT - 1; e >= 0; e--)
            if (e < stop_e && S_in + (unsigned long long)(before + a[e]) >= (1ull << 53)) stop_e = e;
        int my_stop = (stop_e < RS_EPT && i0 + stop_e < NG) ? i0 + stop_e : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) my_stop = min(my_stop, __shfl_xor_sync(0xffffffffu, my_stop, o));
        if (lane == 0 && my_stop != 0x7fffffff) atomicMin(&s_stop[it & 1], my_stop);
        __syncthreads();
        const int stop = s_stop[it & 1];
        const unsigned long long hi = (unsigned long long)kexp << 52;
#pragma unroll
        for (int e = 0; e < RS_EPT; e++)
            if (i0 + e < NG && i0 + e < stop)
                w[i0 + e] = __longlong_as_double((long long)(hi | ((S_in + (unsigned long long)(before + a[e])) & MANT)));
        if (stop == 0x7fffffff) {
            carry = __longlong_as_double((long long)(hi | ((S_in + (unsigned long long)total) & MANT)));
            pos += RS_CHUNK;
        } else {
            if (stop >= i0 && stop < i0 + RS_EPT) {                          // the owner adds its element for real
                const int e = stop - i0;
                const long long pre = before + (e ? a[e - 1] : 0);           // a[] below `stop` is exact
                const double cprev = __longlong_as_double((long long)(hi | ((S_in + (unsigned long long)pre) & MANT)));
                const double cur = cprev + v[e];
                w[stop] = cur;
                s_carry = cur;
            }
            __syncthreads();
            carry = s_carry;
            pos = stop + 1;
        }
        it++;
    }
    if (tid == 0) {
        double slice = carry / (double)NG;                                   // main.py:57
        double u;
        if (u01_in) u = *u01_in;
        else {
            uint32_t r[4];
            rb_philox(0x5eedu, 0u, (uint32_t)c.step_no, 0x52u, c.seed, r);
            u = rb_u01(r[0], r[1]);
        }
        c.plan_scal[0] = slice;
        c.plan_scal[1] = u * slice;                                          // main.py:59
    }
}

// Ancestors from the running sum, in parallel over all SMs (every rank runs it on identical input):
// emitted-after-i = max(0, floor((c_i - start)/slice) + 1)  (main.py:63-64; the running sum is
// non-decreasing because all adjusted weights are >= 0).  Identity when nothing triggered.
__global__ void __launch_bounds__(256) resample_ancestors_kernel(RbCtx c)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, NG = c.n_global;
    if (i >= NG) return;
    if (!c.flags->did_resample) { c.ancestors[i] = i; return; }
    const double *w = c.w_all;
    const double slice = c.plan_scal[0], start = c.plan_scal[1];
    bool bad = false;
    const double f = floor((w[i] - start) / slice);
    const double fp = i ? floor((w[i - 1] - start) / slice) : -1.0;
    if (!(f > -4e18 && f < 4e18)) bad = true;                                // math.floor would raise
    else {
        long long e1 = (long long)f + 1, e0 = i ? (long long)fp + 1 : 0;
        if (e0 < 0) e0 = 0;
        if (e1 < 0) e1 = 0;
        if (i == NG - 1 && e1 != NG) bad = true;                             // main.py:66-67
        if (e1 > NG) { e1 = NG; }
        for (long long s = e0; s < e1; s++) c.ancestors[s] = i;
    }
    if (bad) { c.flags->resample_error = 1; c.flags->resample_error_sticky = 1; }
}

// Local slot j of this rank is global slot rank*N + j.  One warp per new slot: copies the ancestor's
// state and page table, counts the ancestor's local descendants (mult, zero between resamples) and
// finds dup_of[j] = first local slot with the same ancestor (the ancestor vector is non-decreasing, so
// that is a lower bound found by bisection).
__global__ void __launch_bounds__(256) resample_gather_kernel(RbCtx c)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= c.N) return;
    const int j = warp;
    const int err = c.flags->resample_error;                                  // on error: particles unchanged
    const int base = c.rank * c.N;
    const int ag = err ? base + j : c.ancestors[base + j];
    const int a = ag - base;
    if (lane == 1) {
        int rep = j;
        if (!err && c.flags->did_resample) {
            int lo = base, hi = base + j;                 // first index in [base, base + j] with ancestors[idx] >= ag
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (c.ancestors[mid] < ag) lo = mid + 1; else hi = mid;
            }
            rep = lo - base;
        }
        c.dup_of[j] = rep;
    }
    if (a < 0 || a >= c.N) {                                                  // remote ancestor: migration fills it
        if (lane == 0) atomicAdd(&c.flags->remote_needed, 1);
        return;
    }
    const int did = c.flags->did_resample && !err;
    if (lane < 3) c.pose2[3 * (size_t)j + lane] = c.pose[3 * (size_t)a + lane];
    if (lane < 9) c.cov2[9 * (size_t)j + lane] = c.cov[9 * (size_t)a + lane];
    if (lane == 0) {
        atomicAdd(&c.mult[a], 1);
        c.exists2[j] = c.exists[a];
        if (did) c.weight[j] = 1.0;                                           // main.py:77-78 (a == j when !did)
    }
    const uint32_t *src = c.pt + (size_t)a * c.nsub;
    uint32_t *dst = c.pt2 + (size_t)j * c.nsub;
    for (int e = lane; e < c.nsub; e += 32) dst[e] = src[e];
}

// Reference counts: an old particle with m local descendants contributes m - 1, or releases its
// sub-tiles when m = 0.  Leaves mult zero for the next resample.
__global__ void __launch_bounds__(256) resample_refs_kernel(RbCtx c)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= c.N) return;
    const int m = c.mult[warp];
    __syncwarp();
    if (lane == 0) c.mult[warp] = 0;
    if (m == 1) return;
    const uint32_t *src = c.pt + (size_t)warp * c.nsub;
    for (int e = lane; e < c.nsub; e += 32) {
        uint32_t t = src[e];
        if (t == RB_NONE) continue;
        if (m == 0) {
            if (atomicSub(&c.refcnt[t], 1u) == 1u) {                          // last reference: back to the free list
                const int idx = atomicAdd(c.free_count, 1);
                if (idx >= 0 && (uint32_t)idx < c.pool_tiles) c.free_list[idx] = t;
                else atomicExch(&c.flags->world_overflow, 3);                 // cannot happen: the counter is kept in [0, pool_tiles]
            }
        } else {
            atomicAdd(&c.refcnt[t], (unsigned)(m - 1));
        }
    }
}

void rb_launch_resample(const RbCtx &c, const double *weights_all, const double *u01_dev, cudaStream_t s)
{
    resample_plan_kernel<<<1, RS_THREADS, 0, s>>>(c, weights_all, u01_dev);
    resample_ancestors_kernel<<<(c.n_global + 255) / 256, 256, 0, s>>>(c);
}

// Applies the planned ancestors to this rank's particles (local part).
void rb_launch_resample_apply(const RbCtx &c, cudaStream_t s)
{
    rb_launch_resample_gather(c, s);
    rb_launch_resample_refs(c, s);
}

void rb_launch_resample_gather(const RbCtx &c, cudaStream_t s)
{
    resample_gather_kernel<<<(c.N * 32 + 255) / 256, 256, 0, s>>>(c);
}

// c.pt must be the page tables the gather read (the old particles'), c.mult the counts it left.
void rb_launch_resample_refs(const RbCtx &c, cudaStream_t s)
{
    resample_refs_kernel<<<(c.N * 32 + 255) / 256, 256, 0, s>>>(c);
}
