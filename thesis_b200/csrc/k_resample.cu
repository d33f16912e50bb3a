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

#define RS_THREADS 1024
static_assert(RS_THREADS == 1024, "the block scan assumes 32 warps");

__global__ void __launch_bounds__(RS_THREADS) resample_plan_kernel(RbCtx c, const double *__restrict__ w_in,
                                                                   const double *__restrict__ u01_in)
{
    __shared__ double red_mx[32], red_mn[32];
    __shared__ double chunk[RS_THREADS];
    __shared__ long long scan_tot[32];
    __shared__ double s_mn2, s_carry, s_slice, s_start;
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
    if (!s_do) {                                                             // particles unchanged
        for (int i = tid; i < NG; i += RS_THREADS) c.ancestors[i] = i;
        return;
    }
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
    // done in parallel (one element per thread, two barriers).  Chunks with a tie, a
    // binade crossing, a denormal or the very first chunk take the sequential chain.
    const int n_chunks = (NG + RS_THREADS - 1) / RS_THREADS;
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
    double v_next = load(tid);
    for (int k = 0; k < n_chunks; k++) {
        const int i = k * RS_THREADS + tid;
        const bool active = i < NG;
        const double v = v_next;
        v_next = load(i + RS_THREADS);
        const unsigned long long cb = (unsigned long long)__double_as_longlong(carry);
        const int kexp = (int)((cb >> 52) & 0x7ffull);
        const bool fast_ok = !(cb >> 63) && kexp >= 54 && kexp < 0x7ff;      // carry > 0, normal, u normal
        long long a = 0;
        bool slow = !fast_ok;
        if (active && fast_ok && v != 0.0) {
            const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
            const int ve = (int)((vb >> 52) & 0x7ffull);
            if ((vb >> 63) || ve == 0x7ff || ve == 0) slow = true;           // negative, inf / nan, denormal
            else {
                const unsigned long long M = (vb & MANT) | (1ull << 52);
                const int sft = kexp - ve;
                if (sft < 1) slow = true;                                    // v >= 2^k: the sum leaves the binade
                else if (sft <= 54) {
                    const unsigned long long r = M & ((1ull << sft) - 1ull), half = 1ull << (sft - 1);
                    a = (long long)(M >> sft);
                    if (r > half) a += 1;
                    else if (r == half) slow = true;                         // tie
                }                                                            // sft > 54: v < u/4, rounds away
            }
        }
        bool done = false;
        if (!__syncthreads_or(slow)) {
            // inclusive integer scan over the block
            long long p = a;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long q = __shfl_up_sync(0xffffffffu, p, o);
                if (lane >= o) p += q;
            }
            if (lane == 31) scan_tot[warp] = p;
            __syncthreads();
            long long t = scan_tot[lane];                                    // RS_THREADS / 32 == 32 warps
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long q = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += q;
            }
            const long long total = __shfl_sync(0xffffffffu, t, 31);
            const long long before = warp ? __shfl_sync(0xffffffffu, t, warp - 1) : 0ll;
            const unsigned long long S_in = (cb & MANT) | (1ull << 52);
            if (S_in + (unsigned long long)total < (1ull << 53)) {            // stays in the binade
                const unsigned long long hi = (unsigned long long)kexp << 52;
                if (active) w[i] = __longlong_as_double((long long)(hi | ((S_in + (unsigned long long)(before + p)) & MANT)));
                carry = __longlong_as_double((long long)(hi | ((S_in + (unsigned long long)total) & MANT)));
                done = true;
            }
            __syncthreads();                                                 // scan_tot is reused by the next chunk
        }
        if (!done) {                                                         // sequential chain for this chunk
            chunk[tid] = v;
            __syncthreads();
            if (tid == 0) {
                const int n = min(RS_THREADS, NG - k * RS_THREADS);
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
            if (active) w[i] = chunk[tid];
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
        s_slice = slice;
        s_start = u * slice;                                                 // main.py:59
    }
    __syncthreads();
    const double slice = s_slice, start = s_start;
    // emitted-after-i = max(0, floor((c_i - start)/slice) + 1)  (main.py:63-64;
    // the running sum is non-decreasing because all adjusted weights are >= 0)
    bool bad = false;
    for (int i = tid; i < NG; i += RS_THREADS) {
        double f = floor((w[i] - start) / slice);
        double fp = i ? floor((w[i - 1] - start) / slice) : -1.0;
        if (!(f > -4e18 && f < 4e18)) { bad = true; continue; }             // math.floor would raise
        long long e1 = (long long)f + 1, e0 = i ? (long long)fp + 1 : 0;
        if (e0 < 0) e0 = 0;
        if (e1 < 0) e1 = 0;
        if (i == NG - 1 && e1 != NG) bad = true;                             // main.py:66-67
        if (e1 > NG) { e1 = NG; }
        for (long long s = e0; s < e1; s++) c.ancestors[s] = i;
    }
    if (bad) { c.flags->resample_error = 1; c.flags->resample_error_sticky = 1; }
}

// Local slot j of this rank is global slot rank*N + j.
__global__ void __launch_bounds__(256) resample_gather_kernel(RbCtx c)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= c.N) return;
    const int j = warp;
    const int err = c.flags->resample_error;                                  // on error: particles unchanged
    const int a = err ? j : c.ancestors[c.rank * c.N + j] - c.rank * c.N;
    if (a < 0 || a >= c.N) {                                                  // remote ancestor: migration fills it
        if (lane == 0) atomicAdd(&c.flags->remote_needed, 1);
        return;
    }
    const int did = c.flags->did_resample && !err;
    if (lane < 3) c.pose2[3 * (size_t)j + lane] = c.pose[3 * (size_t)a + lane];
    if (lane < 9) c.cov2[9 * (size_t)j + lane] = c.cov[9 * (size_t)a + lane];
    if (lane == 0) {
        c.exists2[j] = c.exists[a];
        if (did) c.weight[j] = 1.0;                                           // main.py:77-78 (a == j when !did)
    }
    const uint32_t *src = c.pt + (size_t)a * c.nsub;
    uint32_t *dst = c.pt2 + (size_t)j * c.nsub;
    for (int e = lane; e < c.nsub; e += 32) dst[e] = src[e];
}

// Number of local descendants of every old local particle.
__global__ void __launch_bounds__(256) resample_mult_kernel(RbCtx c)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.N) return;
    const int a = c.flags->resample_error ? j : c.ancestors[c.rank * c.N + j] - c.rank * c.N;
    if (a >= 0 && a < c.N) atomicAdd(&c.mult[a], 1);
}

__global__ void __launch_bounds__(256) resample_refs_kernel(RbCtx c)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= c.N) return;
    const int m = c.mult[warp];
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
}

// dup_of[j] = first local slot whose ancestor equals slot j's (ancestors are non-decreasing,
// so that is a lower bound found by bisection).
__global__ void __launch_bounds__(256) resample_dups_kernel(RbCtx c)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.N) return;
    int rep = j;
    if (!c.flags->resample_error && c.flags->did_resample) {
        const int base = c.rank * c.N;
        const int a = c.ancestors[base + j];
        int lo = base, hi = base + j;                 // first index in [base, base + j] with ancestors[idx] >= a
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c.ancestors[mid] < a) lo = mid + 1; else hi = mid;
        }
        rep = lo - base;
    }
    c.dup_of[j] = rep;
}

// Applies the planned ancestors to this rank's particles (local part).
void rb_launch_resample_apply(const RbCtx &c, cudaStream_t s)
{
    cudaMemsetAsync(c.mult, 0, sizeof(int) * (size_t)c.N, s);
    resample_mult_kernel<<<(c.N + 255) / 256, 256, 0, s>>>(c);
    int blocks = (c.N * 32 + 255) / 256;
    resample_gather_kernel<<<blocks, 256, 0, s>>>(c);
    resample_refs_kernel<<<blocks, 256, 0, s>>>(c);
    resample_dups_kernel<<<(c.N + 255) / 256, 256, 0, s>>>(c);
}
