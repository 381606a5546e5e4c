// float(Decimal): correctly rounded (half-even) conversion of a Dec to an IEEE double, as CPython
// does for '%E' % weight (phanotate.py:75-76).  Normal range only.
#pragma once
#include <math.h>
#include "dec.cuh"

template <int N>
PB_HD double wide_to_double_rne(const Wide<N>& v, bool sticky, int exp2) {   // v * 2^exp2, v != 0
    int bl = w_bitlen(v);
    u64 m;
    if (bl <= 53) {
        m = ((u64)v.w[1] << 32) | v.w[0];
        return ldexp((double)m, exp2);       // exact (sticky only set when bl > 53 by construction)
    }
    int sh = bl - 54;                        // keep 54 bits: 53 + round bit
    Wide<N> t = w_shr(v, sh);
    u64 top = ((u64)t.w[1] << 32) | t.w[0];
    bool rest = sticky;
    if (!rest && sh > 0) {                   // any bit below the kept ones?
        Wide<N> back = w_shl(t, sh);
        rest = w_cmp(back, v) != 0;
    }
    u64 mant = top >> 1;
    bool rb = top & 1;
    if (rb && (rest || (mant & 1))) mant++;
    return ldexp((double)mant, exp2 + sh + 1);
}

PB_HDNI double dec_to_double(const Dec& d, bool* ok) {
    *ok = true;
    if (dec_is_zero(d)) return d.neg ? -0.0 : 0.0;
    double r;
    if (d.e >= 0) {
        if (d.e > 100) {
            // astronomically large weights (an ORF of many kb in AT-rich sequence): exact up to the largest double, inf beyond
            // (float(Decimal('1E+400')) is inf in CPython too)
            if (d.e > 330) r = HUGE_VAL;
            else {
                Wide<48> n = w_resize<48>(d.c);
                w_mul_pow10(n, d.e);
                r = wide_to_double_rne(n, false, 0);
            }
            return d.neg ? -r : r;
        }
        Wide<16> n = w_resize<16>(d.c);
        w_mul_pow10(n, d.e);
        r = wide_to_double_rne(n, false, 0);
    } else {
        int k = -d.e;
        if (k >= PB_NPOW10) {
            *ok = false;
            return 0.0;
        }
        Wide<8> den = w_pow10<8>(k);
        int s = 56 + w_bitlen(den) - w_bitlen(d.c);
        if (s < 0) s = 0;
        Wide<16> n = w_shl(w_resize<16>(d.c), s);
        Wide<16> q;
        Wide<8> rem;
        w_divmod<16, 8>(n, den, q, rem);
        r = wide_to_double_rne(q, !w_is_zero(rem), -s);
    }
    return d.neg ? -r : r;
}
