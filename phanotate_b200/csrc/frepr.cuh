// Decimal(str(x)) for a Python float x  (reference orfs.py:126: Decimal(str(self.weight_rbs))).
// str(float) is the shortest digit string that round-trips (David Gay's dtoa mode 0 as used by
// CPython's float_repr); Decimal() of that string keeps its digits verbatim, including the
// ".0" CPython appends to integral values in fixed notation.  Free-format digit generation after
// Steele-White / Burger-Dybvig with exact 256-bit integers; ties on the last digit go to even
// as in dtoa.c.
#pragma once
#include <string.h>
#include "dec.cuh"

PB_HDNI Dec dec_from_double_repr(double v, bool* ok) {
    Dec out;
    w_zero(out.c);
    out.e = 0;
    out.neg = 0;
    *ok = true;
    if (v == 0.0) {                       // '0.0'
        out.e = -1;
        return out;
    }
    u64 bits;
#ifdef __CUDA_ARCH__
    bits = (u64)__double_as_longlong(v);
#else
    memcpy(&bits, &v, 8);
#endif
    out.neg = (int)(bits >> 63);
    int be = (int)((bits >> 52) & 0x7FF);
    u64 f = bits & 0xFFFFFFFFFFFFFull;
    if (be == 0x7FF) {
        *ok = false;
        return out;
    }
    int e;
    if (be == 0) e = -1074;
    else {
        f |= 1ull << 52;
        e = be - 1075;
    }
    if (e > 60 || e < -150) {             // outside the range this path produces (ratios of two probabilities)
        *ok = false;
        return out;
    }
    const bool even = (f & 1ull) == 0;
    const bool closer = (f == (1ull << 52)) && be > 1;
    typedef Wide<8> W;
    W r = w_from_u64<8>(f), s = w_from_u64<8>(1), mp = w_from_u64<8>(1), mm = w_from_u64<8>(1);
    if (e >= 0) {
        r = w_shl(r, e + (closer ? 2 : 1));
        s = w_from_u64<8>(closer ? 4 : 2);
        mp = w_shl(mp, e + (closer ? 1 : 0));
        mm = w_shl(mm, e);
    } else {
        r = w_shl(r, closer ? 2 : 1);
        s = w_shl(s, -e + (closer ? 2 : 1));
        mp = w_from_u64<8>(closer ? 2 : 1);
    }
    int fl = 64;
    while (!((f >> (fl - 1)) & 1ull)) fl--;
    int k = (int)ceil((double)(e + fl - 1) * 0.30102999566398120 - 1e-10);
    if (k >= 0) w_mul_pow10(s, k);
    else {
        w_mul_pow10(r, -k);
        w_mul_pow10(mp, -k);
        w_mul_pow10(mm, -k);
    }
    {
        W t = r;
        w_add(t, mp);
        int c = w_cmp(t, s);
        if (even ? (c >= 0) : (c > 0)) k += 1;
        else {
            w_mul_small(r, 10);
            w_mul_small(mp, 10);
            w_mul_small(mm, 10);
        }
    }
    W coef;
    w_zero(coef);
    int n = 0;
    for (;;) {
        // d = floor(r / s), r = r mod s  (d in 0..9)
        u32 d = 0;
        W s8 = w_shl(s, 3), s4 = w_shl(s, 2), s2 = w_shl(s, 1);
        if (w_cmp(r, s8) >= 0) { w_sub(r, s8); d += 8; }
        if (w_cmp(r, s4) >= 0) { w_sub(r, s4); d += 4; }
        if (w_cmp(r, s2) >= 0) { w_sub(r, s2); d += 2; }
        if (w_cmp(r, s) >= 0) { w_sub(r, s); d += 1; }
        int c1 = w_cmp(r, mm);
        bool tc1 = even ? (c1 <= 0) : (c1 < 0);
        W t = r;
        w_add(t, mp);
        int c2 = w_cmp(t, s);
        bool tc2 = even ? (c2 >= 0) : (c2 > 0);
        n++;
        if (!tc1 && !tc2) {
            w_mul_small(coef, 10);
            w_add_small(coef, d);
            w_mul_small(r, 10);
            w_mul_small(mp, 10);
            w_mul_small(mm, 10);
            if (n > 17) {
                *ok = false;
                return out;
            }
            continue;
        }
        if (tc1 && tc2) {
            W r2 = w_shl(r, 1);
            int c = w_cmp(r2, s);
            if (c > 0 || (c == 0 && (d & 1u))) d += 1;
        } else if (tc2) {
            d += 1;
        }
        w_mul_small(coef, 10);
        w_add_small(coef, d);     // d == 10 carries correctly (cannot exceed the digit count by construction)
        break;
    }
    // strip trailing zeros the carry may have produced
    for (;;) {
        W t = coef;
        u32 rem = w_div_p10(t, 1);
        if (rem != 0 || w_is_zero(coef)) break;
        coef = t;
        n--;
    }
    int nd = w_ndigits(coef);
    if (nd > n) {             // carry created an extra leading digit (99.. -> 100..): already stripped above
        k += nd - n;
        n = nd;
    }
    int decpt = k;            // value = 0.DIGITS * 10^decpt
    if (decpt > -4 && decpt <= 16) {
        if (decpt >= n) {     // integral: digits, zeros, ".0"
            w_mul_pow10(coef, decpt - n + 1);
            out.e = -1;
        } else {
            out.e = decpt - n;
        }
    } else {
        out.e = decpt - n;
    }
    out.c = w_resize<4>(coef);
    return out;
}
