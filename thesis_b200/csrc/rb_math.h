/* rb_math.h -- sin, cos and exp in plain IEEE-754 double operations.
 *
 * The particle weights feed a floor((c - u) / slice) (main.py:63): resampled
 * ancestors are bit-exact against the CPU oracle only if the weights are, and CUDA's
 * libm and glibc do not round alike.  These routines use nothing but + - * and fma
 * (correctly rounded everywhere), so the CUDA kernels and oracle/rbpf_oracle.c --
 * which both include this file -- get identical bits.  The leading terms are carried
 * in double-double, which makes the results correctly rounded in all but about one
 * argument in 10^4 (checked against libm in tests/test_rb_math.py); where the
 * reference's own libm is correctly rounded too, nothing changes against it.
 *
 * Compile without floating-point contraction (-fmad=false / -ffp-contract=off).
 */
#ifndef RB_MATH_H
#define RB_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RB_MATH_FN __host__ __device__ static __forceinline__
#else
#define RB_MATH_FN static inline
#endif

/* ---- double-double pieces: a value is hi + lo with |lo| <= ulp(hi)/2 ---- */
RB_MATH_FN void rbm_two_sum(double a, double b, double *s, double *e)
{
    const double x = a + b, bb = x - a;
    *e = (a - (x - bb)) + (b - bb);
    *s = x;
}
RB_MATH_FN void rbm_fast_two_sum(double a, double b, double *s, double *e)      /* |a| >= |b| */
{
    const double x = a + b;
    *e = b - (x - a);
    *s = x;
}
RB_MATH_FN void rbm_two_prod(double a, double b, double *p, double *e)
{
    const double x = a * b;
    *e = fma(a, b, -x);
    *p = x;
}
RB_MATH_FN void rbm_dd_mul(double ah, double al, double bh, double bl, double *h, double *l)
{
    double p, e;
    rbm_two_prod(ah, bh, &p, &e);
    e = e + (ah * bl + al * bh);
    rbm_fast_two_sum(p, e, h, l);
}
RB_MATH_FN void rbm_dd_add(double ah, double al, double bh, double bl, double *h, double *l)
{
    double s, e;
    rbm_two_sum(ah, bh, &s, &e);
    e = e + (al + bl);
    rbm_fast_two_sum(s, e, h, l);
}

/* sin and cos of a.  |a| up to about 1e8 keeps full accuracy (three-part pi/2). */
RB_MATH_FN void rb_sincos(double a, double *sn, double *cs)
{
    const double k = rint(a * 0.6366197723675814);                               /* nearest multiple of pi/2 */
    /* r = a - k * pi/2 as a double-double */
    double ph, pl, qh, ql, rh, rl, e;
    rbm_two_prod(k, 1.5707963267948966, &ph, &pl);
    const double t = a - ph;                                                     /* exact (Sterbenz) */
    rbm_two_prod(k, 6.123233995736766e-17, &qh, &ql);
    rbm_two_sum(t, -pl, &rh, &rl);
    rbm_two_sum(rh, -qh, &rh, &e);
    rl = (rl + e) - (ql + k * -1.4973849048591698e-33);
    rbm_fast_two_sum(rh, rl, &rh, &rl);
    /* z = r^2 */
    double zh, zl;
    rbm_two_prod(rh, rh, &zh, &zl);
    zl = zl + 2.0 * (rh * rl);
    rbm_fast_two_sum(zh, zl, &zh, &zl);
    /* sin r = r + r^3 (S1 + z (S2 + z (S3 + z Qs(z)))) */
    double q = 1.0 / 51090942171709440000.0;                                     /* 1/21! */
    q = fma(q, zh, -1.0 / 121645100408832000.0);
    q = fma(q, zh, 1.0 / 355687428096000.0);
    q = fma(q, zh, -1.0 / 1307674368000.0);
    q = fma(q, zh, 1.0 / 6227020800.0);
    q = fma(q, zh, -1.0 / 39916800.0);
    q = fma(q, zh, 1.0 / 362880.0);
    double uh, ul, vh, vl;
    rbm_two_prod(zh, q, &uh, &ul);
    ul = ul + zl * q;
    rbm_dd_add(-0.0001984126984126984, -1.7209558293420705e-22, uh, ul, &uh, &ul);
    rbm_dd_mul(zh, zl, uh, ul, &uh, &ul);
    rbm_dd_add(0.008333333333333333, 1.1564823173178714e-19, uh, ul, &uh, &ul);
    rbm_dd_mul(zh, zl, uh, ul, &uh, &ul);
    rbm_dd_add(-0.16666666666666666, -9.25185853854297e-18, uh, ul, &uh, &ul);
    rbm_dd_mul(zh, zl, rh, rl, &vh, &vl);                                        /* r^3 */
    rbm_dd_mul(vh, vl, uh, ul, &uh, &ul);
    rbm_dd_add(rh, rl, uh, ul, &uh, &ul);
    const double s_r = uh + ul;
    /* cos r = 1 - z/2 + z^2 (C2 + z (C3 + z (C4 + z Qc(z)))) */
    q = -1.0 / 1124000727777607680000.0;                                         /* -1/22! */
    q = fma(q, zh, 1.0 / 2432902008176640000.0);
    q = fma(q, zh, -1.0 / 6402373705728000.0);
    q = fma(q, zh, 1.0 / 20922789888000.0);
    q = fma(q, zh, -1.0 / 87178291200.0);
    q = fma(q, zh, 1.0 / 479001600.0);
    q = fma(q, zh, -1.0 / 3628800.0);
    rbm_two_prod(zh, q, &uh, &ul);
    ul = ul + zl * q;
    rbm_dd_add(2.48015873015873e-05, 2.1511947866775882e-23, uh, ul, &uh, &ul);
    rbm_dd_mul(zh, zl, uh, ul, &uh, &ul);
    rbm_dd_add(-0.001388888888888889, 5.300543954373577e-20, uh, ul, &uh, &ul);
    rbm_dd_mul(zh, zl, uh, ul, &uh, &ul);
    rbm_dd_add(0.041666666666666664, 2.3129646346357427e-18, uh, ul, &uh, &ul);
    rbm_dd_mul(zh, zl, zh, zl, &vh, &vl);                                        /* z^2 */
    rbm_dd_mul(vh, vl, uh, ul, &uh, &ul);
    rbm_dd_add(-0.5 * zh, -0.5 * zl, uh, ul, &uh, &ul);
    rbm_dd_add(1.0, 0.0, uh, ul, &uh, &ul);
    const double c_r = uh + ul;
    const long long n = (long long)k;
    switch ((int)(n & 3)) {
    case 0: *sn = s_r; *cs = c_r; break;
    case 1: *sn = c_r; *cs = -s_r; break;
    case 2: *sn = -s_r; *cs = -c_r; break;
    default: *sn = -c_r; *cs = s_r; break;
    }
    if (!(fabs(a) < 1e15)) { *sn = a - a; *cs = a - a; }                          /* inf / nan / no digits left -> nan or 0 */
}

/* exp(x); 0 below -745, +inf above 709.78. */
RB_MATH_FN double rb_exp(double x)
{
    if (!(x > -745.2)) return x != x ? x : 0.0;
    if (x > 709.78) return x + 1e308 * 10.0;
    const double k = rint(x * 1.4426950408889634);
    /* r = x - k ln2 as a double-double */
    double ph, pl, rh, rl;
    rbm_two_prod(k, 0.6931471805599453, &ph, &pl);
    const double t = x - ph;                                                     /* exact */
    rbm_two_sum(t, -pl, &rh, &rl);
    rl = rl - k * 2.3190468138462996e-17;
    rbm_fast_two_sum(rh, rl, &rh, &rl);
    /* e^r = 1 + r + r^2 (1/2 + r (1/6 + r (1/24 + r (1/120 + r Q(r))))) */
    double q = 1.0 / 355687428096000.0;                                          /* 1/17! */
    q = fma(q, rh, 1.0 / 20922789888000.0);
    q = fma(q, rh, 1.0 / 1307674368000.0);
    q = fma(q, rh, 1.0 / 87178291200.0);
    q = fma(q, rh, 1.0 / 6227020800.0);
    q = fma(q, rh, 1.0 / 479001600.0);
    q = fma(q, rh, 1.0 / 39916800.0);
    q = fma(q, rh, 1.0 / 3628800.0);
    q = fma(q, rh, 1.0 / 362880.0);
    q = fma(q, rh, 1.0 / 40320.0);
    q = fma(q, rh, 1.0 / 5040.0);
    q = fma(q, rh, 1.0 / 720.0);
    double uh, ul, zh, zl;
    rbm_two_prod(rh, q, &uh, &ul);
    ul = ul + rl * q;
    rbm_dd_add(0.008333333333333333, 1.1564823173178714e-19, uh, ul, &uh, &ul);
    rbm_dd_mul(rh, rl, uh, ul, &uh, &ul);
    rbm_dd_add(0.041666666666666664, 2.3129646346357427e-18, uh, ul, &uh, &ul);
    rbm_dd_mul(rh, rl, uh, ul, &uh, &ul);
    rbm_dd_add(0.16666666666666666, 9.25185853854297e-18, uh, ul, &uh, &ul);
    rbm_dd_mul(rh, rl, uh, ul, &uh, &ul);
    rbm_dd_add(0.5, 0.0, uh, ul, &uh, &ul);
    rbm_dd_mul(rh, rl, rh, rl, &zh, &zl);
    rbm_dd_mul(zh, zl, uh, ul, &uh, &ul);
    rbm_dd_add(rh, rl, uh, ul, &uh, &ul);
    rbm_dd_add(1.0, 0.0, uh, ul, &uh, &ul);
    const double m = uh + ul;                                                    /* in [0.7, 1.42] */
    /* scale by 2^k in two exact steps (the second one rounds once when the result is subnormal) */
    const int ki = (int)k;
    const int k1 = ki / 2, k2 = ki - k1;
    int64_t b1 = (int64_t)(k1 + 1023) << 52, b2 = (int64_t)(k2 + 1023) << 52;
    double s1, s2;
    memcpy(&s1, &b1, 8);
    memcpy(&s2, &b2, 8);
    return (m * s1) * s2;
}

#endif /* RB_MATH_H */
