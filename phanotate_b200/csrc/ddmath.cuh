// Double-double arithmetic (a value = hi + lo, two IEEE doubles, |lo| <= ulp(hi)/2; about 104 bits) for the
// certified integer weights of fast.cuh.  Every routine is a few fused multiply-adds: the B200's FP64 pipe
// does the closed forms of ORF and overlap weights at a small fraction of the cost of the 224-bit fixed
// point.  Error bounds (relative, for normalised inputs): add <= 2^-103, mul <= 2^-102, recip/div <= 2^-100
// (Joldes, Muller, Popescu, "Tight and rigorous error bounds for basic building blocks of double-word
// arithmetic", ACM TOMS 2017); the callers budget 2^-100 per operation.
#pragma once
#include "fxpow.cuh"

struct DD {
    double hi, lo;
};
PB_HD double pb_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
PB_HD DD dd_two_sum(double a, double b) {
    DD r;
    r.hi = a + b;
    const double bb = r.hi - a;
    r.lo = (a - (r.hi - bb)) + (b - bb);
    return r;
}
PB_HD DD dd_quick(double a, double b) {      // |a| >= |b|
    DD r;
    r.hi = a + b;
    r.lo = b - (r.hi - a);
    return r;
}
PB_HD DD dd_add(const DD& a, const DD& b) {
    DD s = dd_two_sum(a.hi, b.hi);
    const DD t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_quick(s.hi, s.lo);
    s.lo += t.lo;
    return dd_quick(s.hi, s.lo);
}
PB_HD DD dd_mul(const DD& a, const DD& b) {
    const double p = a.hi * b.hi;
    double e = pb_fma(a.hi, b.hi, -p);
    e = pb_fma(a.hi, b.lo, e);
    e = pb_fma(a.lo, b.hi, e);
    return dd_quick(p, e);
}
PB_HD DD dd_mul_d(const DD& a, double b) {
    const double p = a.hi * b;
    double e = pb_fma(a.hi, b, -p);
    e = pb_fma(a.lo, b, e);
    return dd_quick(p, e);
}
PB_HD DD dd_from_d(double v) {
    DD r;
    r.hi = v;
    r.lo = 0.0;
    return r;
}
PB_HD DD dd_recip(const DD& b) {             // 1/b with two correction steps
    const double q1 = 1.0 / b.hi;
    DD t = dd_mul_d(b, q1);
    t.hi = -t.hi;
    t.lo = -t.lo;
    DD r = dd_add(dd_from_d(1.0), t);
    const double q2 = r.hi / b.hi;
    t = dd_mul_d(b, q2);
    t.hi = -t.hi;
    t.lo = -t.lo;
    r = dd_add(r, t);
    const double q3 = r.hi / b.hi;
    DD q = dd_quick(q1, q2);
    return dd_add(q, dd_from_d(q3));
}
// non-negative Dec with a coefficient below 2^96 and exponent in (-PB_NP10DD, 0] -> DD
PB_HD bool dd_from_dec(const Dec& d, DD& out) {
    out = dd_from_d(0.0);
    if (dec_is_zero(d)) return true;
    if (d.neg || d.c.w[3] != 0 || d.e > 0 || -d.e >= PB_NP10DD) return false;
    const double a = (double)d.c.w[2] * 18446744073709551616.0, b = (double)d.c.w[1] * 4294967296.0, c = (double)d.c.w[0];
    DD s = dd_two_sum(a, b);
    const DD t = dd_two_sum(s.hi, c);
    const DD cc = dd_quick(t.hi, s.lo + t.lo);          // exact: the coefficient has at most 96 bits
    DD p;
    p.hi = TBL(p10neg_dd)[-d.e][0];
    p.lo = TBL(p10neg_dd)[-d.e][1];
    out = dd_mul(cc, p);
    return true;
}
PB_HD DD dd_neg(const DD& a) {
    DD r;
    r.hi = -a.hi;
    r.lo = -a.lo;
    return r;
}
PB_HD DD dd_div(const DD& a, const DD& b) { return dd_mul(a, dd_recip(b)); }
PB_HD DD dd_table(const double* row) {
    DD r;
    r.hi = row[0];
    r.lo = row[1];
    return r;
}
// -ln(1 - p) for 0 < p <= 1/8:  2 atanh(z), z = p / (2 - p);  series in z^2 <= 2^-7.8, 14 terms -> < 2^-109
PB_HDNI DD dd_neglog1m(const DD& p) {
    const DD z = dd_div(p, dd_add(dd_from_d(2.0), dd_neg(p)));
    const DD t = dd_mul(z, z);
    DD s = dd_table(TBL(inv_odd_dd)[13]);
#pragma unroll 1
    for (int k = 12; k >= 0; k--) s = dd_add(dd_mul(s, t), dd_table(TBL(inv_odd_dd)[k]));
    DD r = dd_mul(z, s);
    r.hi *= 2.0;
    r.lo *= 2.0;
    return r;
}
// exp(T) = P * 2^K for 0 <= T < 700, P in [0.70, 1.42]:  T = K ln2 + r, exp(r) = exp(r/512)^512, Taylor to degree 9
// (|r/512| <= 6.8e-4: remainder < 2^-113).  Relative error < 2^-92 (nine squarings amplify 2^-102 by 2^9 and add theirs).
PB_HDNI DD dd_exp_split(const DD& T, int* K) {
    const double kd = floor(T.hi * 1.4426950408889634 + 0.5);
    DD r = dd_add(T, dd_neg(dd_mul_d(dd_table(TBL(ln2_dd)), kd)));
    r.hi *= 0.001953125;
    r.lo *= 0.001953125;
    DD p = dd_table(TBL(invfact_dd)[9]);
#pragma unroll 1
    for (int n = 8; n >= 0; n--) p = dd_add(dd_mul(p, r), dd_table(TBL(invfact_dd)[n]));
#pragma unroll 1
    for (int q = 0; q < 9; q++) p = dd_mul(p, p);
    *K = (int)kd;
    return p;
}
