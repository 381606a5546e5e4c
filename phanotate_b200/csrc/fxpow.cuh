// x ** y for a non-integer y, bit-compatible with Python's Decimal.__pow__ on this path.
//
// libmpdec (_mpd_qpow_real) computes exp(y * ln x) with correctly rounded ln and exp at
// prec + 4 + 19 = 51 digits and rounds the result to 28 digits.  We evaluate exp(y * ln x) in
// binary fixed point (Q32.192, absolute error ~1e-56) and round ONCE to 28 digits half-even.  The
// two agree unless exp(y ln x) lies within ~1e-50 (relative) of a 28-digit rounding boundary, which
// cannot be hit by chance (p ~ 1e-22 per call; exp of a non-zero rational is never exactly on a
// boundary).  tests/test_dec_host.py checks 10^5+ random (x, y) pairs from the path's value ranges
// against Python's decimal.
//
// exp(T):  T = K ln2 + r, r in [0, ln2);  r = j1/2^8 + j2/2^16 + j3/2^24 + u;
//          exp(r) = E1[j1] E2[j2] E3[j3] sum_{n<=8} u^n/n!   (u < 2^-24)
// ln(X):   L0 = log(double(X)); Z = X exp(-L0) = 1+eps, |eps| < 2^-45;
//          ln X = L0 + eps - eps^2/2 + eps^3/3 - eps^4/4
#pragma once
#include <math.h>
#include <string.h>
#include "dec.cuh"

#define FX_N 7
#define FX_FR 192
typedef Wide<FX_N> Fx;

struct SFx {     // signed fixed point
    Fx m;
    i32 neg;
};

PB_HD Fx fx_table(const u32* row) {
    Fx r;
#pragma unroll
    for (int i = 0; i < FX_N; i++) r.w[i] = row[i];
    return r;
}
PB_HD Fx fx_one() {
    Fx r;
    w_zero(r);
    r.w[6] = 1;
    return r;
}
// a*b >> 192 from the partial products a_i*b_j with i+j >= CUT and j <= JMAX only.
// Dropped columns sum to < 2^(32*CUT - 348) (absolute), so CUT = 4 costs < 2^-220; larger CUTs are
// used where the result is multiplied by further small factors (Horner steps of exp).
// JMAX = 5 is for operands known to be < 1 (limb 6 zero).
template <int CUT, int JMAX>
PB_HD Fx fx_mul_cols(const Fx& a, const Fx& b) {
    Fx r;
    u64 acc = 0;
    u32 hi = 0;
#pragma unroll
    for (int k = CUT; k <= 12; k++) {
#pragma unroll
        for (int i = 0; i < FX_N; i++) {
            const int j = k - i;
            if (j >= 0 && j <= JMAX) {
                u64 p = (u64)a.w[i] * b.w[j];
                acc += p;
                hi += (acc < p) ? 1u : 0u;
            }
        }
        if (k >= 6) r.w[k - 6] = (u32)acc;
        acc = (acc >> 32) | ((u64)hi << 32);
        hi = 0;
    }
    return r;
}
PB_HDNI Fx fx_mul(const Fx& a, const Fx& b) { return fx_mul_cols<4, 6>(a, b); }
// Horner step of exp: p*u with u < 2^-24; `rest` = number of further multiplications by u the result
// still goes through, which relaxes the absolute accuracy needed here by 2^(24*rest)
PB_HDNI Fx fx_mul_u(const Fx& p, const Fx& u, int rest) {
    // three variants only: instruction-cache footprint matters more than the last few skipped columns
    // admissible cut per `rest` (0..7): 4,5,6,7,7,8,9,10
    if (rest <= 2) return fx_mul_cols<4, 5>(p, u);
    if (rest <= 5) return fx_mul_cols<7, 5>(p, u);
    return fx_mul_cols<9, 5>(p, u);
}
PB_HD double fx_to_double(const Fx& a) {
    int bl = w_bitlen(a);
    if (bl == 0) return 0.0;
    u64 top;
    if (bl >= 64) {
        Fx s = w_shr(a, bl - 64);
        top = ((u64)s.w[1] << 32) | s.w[0];
        return ldexp((double)top, bl - 64 - FX_FR);
    }
    top = ((u64)a.w[1] << 32) | a.w[0];
    return ldexp((double)top, -FX_FR);
}
// exact for 0 <= v < 2^32 whose lowest set bit is >= 2^-192
PB_HD Fx fx_from_double(double v) {
    Fx r;
    w_zero(r);
    if (!(v > 0.0)) return r;
    int ex;
    double m = frexp(v, &ex);               // v = m * 2^ex, m in [0.5,1)
    u64 mi = (u64)ldexp(m, 53);             // 53-bit integer mantissa
    int sh = ex - 53 + FX_FR;               // value = mi * 2^(ex-53); fixed = mi << sh
    r.w[0] = (u32)mi;
    r.w[1] = (u32)(mi >> 32);
    if (sh >= 0) return w_shl(r, sh);
    return w_shr(r, -sh);
}

// Dec (non-negative, value < 2^32) -> Fx
PB_HDNI Fx fx_from_dec(const Dec& x, bool* ok) {
    Fx r;
    w_zero(r);
    *ok = true;
    if (dec_is_zero(x)) return r;
    if (x.e >= 0) {
        Wide<8> c = w_resize<8>(x.c);
        if (w_ndigits(c) + x.e > 9) {
            *ok = false;
            return r;
        }
        w_mul_pow10(c, x.e);
        r.w[6] = c.w[0];
        return r;
    }
    int k = -x.e;
    if (k >= PB_NINV10) {
        *ok = false;
        return r;
    }
    Fx inv = fx_table(TBL(inv10_w7)[k]);
    int S = (int)TBL(inv10_shift)[k];
    Wide<11> p = w_mul(x.c, inv);
    Wide<11> s = w_shr(p, S);
    for (int i = FX_N; i < 11; i++)
        if (s.w[i]) *ok = false;
    return w_resize<FX_N>(s);
}

// exp(T) for signed T as P * 2^K with P in [1, 2) (Q32.192); |T| <= 700.
PB_HDNI Fx fx_exp_core(const SFx& T, int* Kout, bool* ok) {
    *ok = true;
    *Kout = 0;
    const Fx ln2 = fx_table(TBL(fx_ln2));
    Fx r;
    int K;
    if (w_is_zero(T.m)) return fx_one();
    double td = fx_to_double(T.m);
    if (td > 700.0) {
        *ok = false;
        return fx_one();
    }
    int k = (int)floor(td * 1.4426950408889634);
    if (!T.neg) {
        // T = k ln2 + r
        Fx kl = ln2;
        u32 carry = w_mul_small(kl, (u32)k);
        (void)carry;
        r = T.m;
        if (w_cmp(r, kl) < 0) {
            k -= 1;
            w_sub(kl, ln2);
        }
        w_sub(r, kl);
        if (w_cmp(r, ln2) >= 0) {
            w_sub(r, ln2);
            k += 1;
        }
        K = k;
    } else {
        // -t = -(k+1) ln2 + r,  r = (k+1) ln2 - t in (0, ln2]
        int kk = k + 1;
        Fx kl = ln2;
        w_mul_small(kl, (u32)kk);
        if (w_cmp(kl, T.m) < 0) {
            kk += 1;
            w_add(kl, ln2);
        }
        r = kl;
        w_sub(r, T.m);
        if (w_cmp(r, ln2) >= 0) {
            w_sub(r, ln2);
            kk -= 1;
        }
        K = -kk;
    }
    u32 top = r.w[5];
    int j1 = (int)(top >> 24), j2 = (int)((top >> 16) & 255u), j3 = (int)((top >> 8) & 255u);
    if (r.w[6] != 0) {   // r >= 1 cannot happen (ln2 < 1)
        *ok = false;
        return fx_one();
    }
    Fx u = r;
    u.w[5] = top & 255u;
    // Horner for sum u^n / n!, n = 0..8
    Fx p = fx_table(TBL(fx_invfact)[8]);
#pragma unroll 1
    for (int n = 7; n >= 0; n--) {
        p = fx_mul_u(p, u, n);
        Fx cn = fx_table(TBL(fx_invfact)[n]);
        w_add(p, cn);
    }
    if (j1) p = fx_mul(p, fx_table(TBL(fx_e1)[j1]));
    if (j2) p = fx_mul(p, fx_table(TBL(fx_e2)[j2]));
    if (j3) p = fx_mul(p, fx_table(TBL(fx_e3)[j3]));
    *Kout = K;
    return p;
}
// exp(T) for signed T; result must be < 2^32.
PB_HDNI Fx fx_exp(const SFx& T, bool* ok) {
    int K;
    Fx p = fx_exp_core(T, &K, ok);
    if (!*ok) return fx_one();
    if (K > 0) {
        if (K >= 30) {
            *ok = false;
            return fx_one();
        }
        p = w_shl(p, K);
    } else if (K < 0) {
        if (-K >= 32 * FX_N) w_zero(p);
        else p = w_shr(p, -K);
    }
    return p;
}

// ln(X), X > 0
PB_HDNI SFx fx_ln(const Fx& X, bool* ok) {
    SFx L;
    w_zero(L.m);
    L.neg = 0;
    *ok = true;
    Fx one = fx_one();
    if (w_cmp(X, one) == 0) return L;
    double xd = fx_to_double(X);
    double l0 = log(xd);
    SFx L0;
    L0.neg = l0 < 0.0;
    L0.m = fx_from_double(fabs(l0));
    SFx mL0 = L0;
    mL0.neg ^= 1;
    Fx E0 = fx_exp(mL0, ok);
    Fx Z = fx_mul(X, E0);
    // eps = Z - 1
    Fx eps;
    int eneg;
    if (w_cmp(Z, one) >= 0) {
        eps = Z;
        w_sub(eps, one);
        eneg = 0;
    } else {
        eps = one;
        w_sub(eps, Z);
        eneg = 1;
    }
    if (w_bitlen(eps) > FX_FR - 40) *ok = false;   // the double seed must leave |eps| < 2^-40
    Fx e2 = fx_mul(eps, eps);
    Fx e3 = fx_mul(e2, eps);
    Fx e4 = fx_mul(e3, eps);
    Fx t2 = w_shr(e2, 1);
    Fx t4 = w_shr(e4, 2);
    Fx t3 = e3;
    {   // divide by 3
        u64 rem = 0;
        for (int i = FX_N - 1; i >= 0; i--) {
            u64 x = (rem << 32) | t3.w[i];
            t3.w[i] = (u32)(x / 3u);
            rem = x % 3u;
        }
    }
    // log1p(eps) = eps - e2/2 + e3/3 - e4/4 with eps signed: odd powers carry eneg, even are positive
    // pos = (eneg ? 0 : eps) + (eneg ? 0 : t3);  negs = t2 + t4 + (eneg ? eps + t3 : 0)
    Fx pos, ngv;
    w_zero(pos);
    w_zero(ngv);
    w_add(ngv, t2);
    w_add(ngv, t4);
    if (eneg) {
        w_add(ngv, eps);
        w_add(ngv, t3);
    } else {
        w_add(pos, eps);
        w_add(pos, t3);
    }
    // L = L0 + pos - ngv
    if (L0.neg) w_add(ngv, L0.m);
    else w_add(pos, L0.m);
    if (w_cmp(pos, ngv) >= 0) {
        L.m = pos;
        w_sub(L.m, ngv);
        L.neg = 0;
    } else {
        L.m = ngv;
        w_sub(L.m, pos);
        L.neg = 1;
    }
    return L;
}

// Fx (positive) -> Dec rounded half-even to prec digits
PB_HDNI Dec fx_to_dec(const Fx& V, int prec, bool* ok) {
    Dec r;
    w_zero(r.c);
    r.e = 0;
    r.neg = 0;
    *ok = true;
    int bl = w_bitlen(V);
    if (bl == 0) {
        *ok = false;
        return r;
    }
    int l2 = bl - 1 - FX_FR;                        // floor(log2 V)
    int adj = (int)floor((double)l2 * 0.30102999566398120);
    int scale = prec - 1 - adj;
    Wide<5> I;
    Wide<6> F;
    for (int iter = 0; iter < 4; iter++) {
        if (scale < 0 || scale > 38) {
            *ok = false;
            return r;
        }
        Wide<4> p10 = w_pow10<4>(scale);
        Wide<11> P = w_mul(V, p10);
#pragma unroll
        for (int i = 0; i < 6; i++) F.w[i] = P.w[i];
#pragma unroll
        for (int i = 0; i < 5; i++) I.w[i] = P.w[i + 6];
        Wide<5> hi = w_pow10<5>(prec), lo = w_pow10<5>(prec - 1);
        if (w_cmp(I, hi) >= 0) {
            scale--;
            continue;
        }
        if (w_cmp(I, lo) < 0) {
            scale++;
            continue;
        }
        break;
    }
    bool up;
    if (F.w[5] & 0x80000000u) {
        bool rest = (F.w[5] & 0x7FFFFFFFu) | F.w[4] | F.w[3] | F.w[2] | F.w[1] | F.w[0];
        up = rest || (I.w[0] & 1u);
    } else {
        up = false;
    }
    if (up) w_add_small(I, 1u);
    Wide<5> hi = w_pow10<5>(prec);
    if (w_cmp(I, hi) == 0) {
        I = w_pow10<5>(prec - 1);
        scale--;
    }
    r.c = w_resize<4>(I);
    r.e = -scale;
    return r;
}

// x ** y, x > 0 (Dec), y given in fixed point with sign (non-integer on this path)
PB_HDNI Dec dec_pow_fx(const Dec& x, const Fx& y, int yneg, int prec, bool* ok) {
    if (dec_is_one_abs(x)) {          // _qcheck_pow_one, non-integer exponent: 1.000...0 with prec digits
        Dec r;
        r.c = w_pow10<4>(prec - 1);
        r.e = -(prec - 1);
        r.neg = 0;
        *ok = true;
        return r;
    }
    bool ok1, ok2, ok3, ok4;
    Fx X = fx_from_dec(x, &ok1);
    SFx L = fx_ln(X, &ok2);
    SFx T;
    T.m = fx_mul(L.m, y);
    T.neg = L.neg ^ yneg;
    Fx V = fx_exp(T, &ok3);
    Dec r = fx_to_dec(V, prec, &ok4);
    *ok = ok1 && ok2 && ok3 && ok4;
    return r;
}
// ln of a = round28(V) given V = exp(T) in full precision:  ln a = T + log1p((a - V)/V).
// |(a-V)/V| <= 5e-29, so log1p is its argument to 2^-187 and 1/V is only needed to ~100 bits
// (one Newton step from a double seed).  Saves a full ln (one exp) per use.
PB_HDNI SFx fx_ln_of_rounded(const Dec& a, const Fx& V, const SFx& T, bool* ok) {
    bool o1;
    Fx Af = fx_from_dec(a, &o1);
    *ok = o1;
    Fx d;
    int dneg;
    if (w_cmp(Af, V) >= 0) {
        d = Af;
        w_sub(d, V);
        dneg = 0;
    } else {
        d = V;
        w_sub(d, Af);
        dneg = 1;
    }
    // r ~ 1/V : r0 from double, r1 = r0 * (2 - V*r0)
    Fx r0 = fx_from_double(1.0 / fx_to_double(V));
    Fx t = fx_mul(V, r0);
    Fx two;
    w_zero(two);
    two.w[6] = 2;
    w_sub(two, t);
    Fx r1 = fx_mul(r0, two);
    Fx delta = fx_mul(d, r1);
    // L = T + delta (signed)
    SFx L;
    if (T.neg == dneg) {
        L.m = T.m;
        w_add(L.m, delta);
        L.neg = T.neg;
    } else if (w_cmp(T.m, delta) >= 0) {
        L.m = T.m;
        w_sub(L.m, delta);
        L.neg = T.neg;
    } else {
        L.m = delta;
        w_sub(L.m, T.m);
        L.neg = dneg;
    }
    return L;
}

// Same with ln(x) supplied (ORF scoring reuses ln(1-pstop) for three exponents)
PB_HDNI Dec dec_pow_ln(const SFx& lnx, const Fx& y, int prec, bool* ok, SFx* Tout = 0, Fx* Vout = 0) {
    bool ok3, ok4;
    SFx T;
    T.m = fx_mul(lnx.m, y);
    T.neg = lnx.neg;
    Fx V = fx_exp(T, &ok3);
    Dec r = fx_to_dec(V, prec, &ok4);
    *ok = ok3 && ok4;
    if (Tout) *Tout = T;
    if (Vout) *Vout = V;
    return r;
}
