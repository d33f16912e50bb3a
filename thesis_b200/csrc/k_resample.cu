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

#define RS_THREADS 256                  // 8 warps: the kernel is a chain of dependent chunk scans, not throughput work
#define RS_EPT 4                        // consecutive elements per thread
#define RS_CHUNK (RS_THREADS * RS_EPT)  // 1,024
static_assert(RS_THREADS == 256, "the block scan assumes 8 warps");

__global__ void __launch_bounds__(RS_THREADS) resample_plan_kernel(RbCtx c, const double *__restrict__ w_in,
                                                                   const double *__restrict__ u01_in)
{
    __shared__ double red_mx[RS_THREADS / 32], red_mn[RS_THREADS / 32];
    __shared__ double chunk[RS_CHUNK];
    __shared__ long long scan_tot[RS_THREADS / 32];
    __shared__ double s_mn2, s_carry;
    __shared__ int s_do;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NG = c.n_global;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double *w = c.w_all;                                                     // scratch: adjusted weights, then cumsum

    // max / min of the raw weights (main.py:50)
    double mx = -INF, mn = INF;
    for (int i = tid; i < NG; i += RS_THREADS) { double v = w_in[i]; mx = fmax(mx, v); mn = fmin(mn, v); }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) { red_mx[warp] = mx; red_mn[warp] = mn; }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < RS_THREADS / 32; k++) { mx = fmax(mx, red_mx[k]); mn = fmin(mn, red_mn[k]); }
        s_do = (mx - mn > RB_RESAMPLE_TRIGGER) ? 1 : 0;
        if (c.flags->pool_exhausted) s_do = 0;                               // maps missed a scan: freeze the set until the caller reacts
        c.flags->did_resample = s_do;
        c.flags->resample_error = 0;
        c.flags->remote_needed = 0;
        if (s_do) c.stats->resamples += 1ull;
    }
    __syncthreads();
    if (!s_do) return;                                                       // particles unchanged (identity ancestors)
    // -inf -> 0, then min of the result (main.py:53-54)
    mn = INF;
    for (int i = tid; i < NG; i += RS_THREADS) {
        double v = w_in[i];
        if (v == -INF) v = 0.0;
        w[i] = v;
        mn = fmin(mn, v);
    }
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    __syncthreads();
    if (lane == 0) red_mn[warp] = mn;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < RS_THREADS / 32; k++) mn = fmin(mn, red_mn[k]);
        s_mn2 = mn;
        s_carry = 0.0;
    }
    __syncthreads();
    const double shift = s_mn2 < 0.0 ? fabs(s_mn2) : 0.0;
    const bool do_shift = s_mn2 < 0.0;
    // running sum c_i = fl(c_{i-1} + v_i) (main.py:57 == :62), float64, round to nearest even:
    // inherently a chain of N dependent adds -- except that while the sum stays inside one
    // binade [2^k, 2^(k+1)) every add rounds to a multiple of u = 2^(k-52).  With S = c/u
    // (an integer in [2^52, 2^53)) and v = (m + f) u, m integer, 0 <= f < 1:
    //     S' = S + m + [f > 1/2]          (f == 1/2, a tie, depends on the parity of S + m)
    // so a chunk without ties that does not leave the binade is an exact INTEGER prefix sum,
    // done in parallel (RS_EPT consecutive elements per thread, two barriers).  Chunks with a tie,
    // a binade crossing, a denormal or the very first chunk take the sequential chain.
    const int n_chunks = (NG + RS_CHUNK - 1) / RS_CHUNK;
    const unsigned long long MANT = (1ull << 52) - 1ull;
    double carry = 0.0;                                                      // uniform over the block
    auto load = [&](int i) {
        double v = 0.0;
        if (i < NG) {
            v = w[i];
            if (do_shift && v != 0.0) v += shift;                            // main.py:55
        }
        return v;
    };
    double v_next[RS_EPT];
#pragma unroll
    for (int e = 0; e < RS_EPT; e++) v_next[e] = load(RS_EPT * tid + e);
    for (int k = 0; k < n_chunks; k++) {
        const int i0 = k * RS_CHUNK + RS_EPT * tid;
        double v[RS_EPT];
#pragma unroll
        for (int e = 0; e < RS_EPT; e++) { v[e] = v_next[e]; v_next[e] = load(i0 + RS_CHUNK + e); }
        const unsigned long long cb = (unsigned long long)__double_as_longlong(carry);
        const int kexp = (int)((cb >> 52) & 0x7ffull);
        const bool fast_ok = !(cb >> 63) && kexp >= 54 && kexp < 0x7ff;      // carry > 0, normal, u normal
        long long a[RS_EPT];
        bool slow = !fast_ok;
#pragma unroll
        for (int e = 0; e < RS_EPT; e++) {
            a[e] = 0;
            if (i0 + e < NG && fast_ok && v[e] != 0.0) {
                const unsigned long long vb = (unsigned long long)__double_as_longlong(v[e]);
                const int ve = (int)((vb >> 52) & 0x7ffull);
                if ((vb >> 63) || ve == 0x7ff || ve == 0) slow = true;       // negative, inf / nan, denormal
                else {
                    const unsigned long long M = (vb & MANT) | (1ull << 52);
                    const int sft = kexp - ve;
                    if (sft < 1) slow = true;                                // v >= 2^k: the sum leaves the binade
                    else if (sft <= 54) {
                        const unsigned long long r = M & ((1ull << sft) - 1ull), half = 1ull << (sft - 1);
                        a[e] = (long long)(M >> sft);
                        if (r > half) a[e] += 1;
                        else if (r == half) slow = true;                     // tie
                    }                                                        // sft > 54: v < u/4, rounds away
                }
            }
        }
        bool done = false;
        if (!__syncthreads_or(slow)) {
            // inclusive integer scan: inside the thread, across the warp, across the 8 warps
#pragma unroll
            for (int e = 1; e < RS_EPT; e++) a[e] += a[e - 1];
            long long p = a[RS_EPT - 1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long q = __shfl_up_sync(0xffffffffu, p, o);
                if (lane >= o) p += q;
            }
            if (lane == 31) scan_tot[warp] = p;
            __syncthreads();
            long long before = p - a[RS_EPT - 1], total = 0;                 // exclusive prefix of this thread inside the warp
#pragma unroll
            for (int q = 0; q < RS_THREADS / 32; q++) {
                const long long t = scan_tot[q];
                if (q < warp) before += t;
                total += t;
            }
            const unsigned long long S_in = (cb & MANT) | (1ull << 52);
            if (S_in + (unsigned long long)total < (1ull << 53)) {            // stays in the binade
                const unsigned long long hi = (unsigned long long)kexp << 52;
#pragma unroll
                for (int e = 0; e < RS_EPT; e++)
                    if (i0 + e < NG) w[i0 + e] = __longlong_as_double((long long)(hi | ((S_in + (unsigned long long)(before + a[e])) & MANT)));
                carry = __longlong_as_double((long long)(hi | ((S_in + (unsigned long long)total) & MANT)));
                done = true;
            }
            __syncthreads();                                                 // scan_tot is reused by the next chunk
        }
        if (!done) {                                                         // sequential chain for this chunk
            // (a[] may hold partial prefix sums here; the chain works on the values themselves)
#pragma unroll
            for (int e = 0; e < RS_EPT; e++) chunk[RS_EPT * tid + e] = v[e];
            __syncthreads();
            if (tid == 0) {
                const int n = min(RS_CHUNK, NG - k * RS_CHUNK);
                double cur = carry;
                int e0 = 0;
                for (; e0 + 8 <= n; e0 += 8) {
                    double x[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) x[e] = chunk[e0 + e];
#pragma unroll
                    for (int e = 0; e < 8; e++) { cur += x[e]; x[e] = cur; }
#pragma unroll
                    for (int e = 0; e < 8; e++) chunk[e0 + e] = x[e];
                }
                for (; e0 < n; e0++) { cur += chunk[e0]; chunk[e0] = cur; }
                s_carry = cur;
            }
            __syncthreads();
#pragma unroll
            for (int e = 0; e < RS_EPT; e++)
                if (i0 + e < NG) w[i0 + e] = chunk[RS_EPT * tid + e];
            carry = s_carry;
            __syncthreads();
        }
    }
    if (tid == 0) s_carry = carry;
    __syncthreads();
    if (tid == 0) {
        double slice = s_carry / (double)NG;                                 // main.py:57
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
    int blocks = (c.N * 32 + 255) / 256;
    resample_gather_kernel<<<blocks, 256, 0, s>>>(c);
    resample_refs_kernel<<<blocks, 256, 0, s>>>(c);
}
