// Decimal floating point with General-Decimal-Arithmetic semantics, as far as the PHANOTATE hot
// path uses it: the reference does all scoring in Python's decimal.Decimal (prec 28,
// ROUND_HALF_EVEN; functions.py:7,26-46,140-141,153,174-178,262-298; orfs.py:4,126,168-173) and
// the results are compared digit for digit, so every operation here reproduces libmpdec 2.5.1:
//   add/sub/mul/div : exact result rounded half-even to `prec` digits, GDA exponent rules
//                     (ideal exponent of exact quotients, zero operands, sticky tiny operand);
//   integer power   : libmpdec's _mpd_qpow_int -- left-to-right square-and-multiply at working
//                     precision prec + ndigits(n) + 2, one final rounding (this is NOT the
//                     correctly rounded power, and the difference is visible);
//   1**y            : _qcheck_pow_one.
// Only finite numbers; exponents stay far from Emax/Emin on this path.
#pragma once
#include "wide.cuh"

#define PB_PREC 28

struct Dec {
    Wide<4> c;   // coefficient, < 10^38
    i32 e;       // exponent
    i32 neg;     // sign (1 = negative)
};

PB_HD Dec dec_from_u64(u64 v, int neg = 0) {
    Dec r;
    r.c = w_from_u64<4>(v);
    r.e = 0;
    r.neg = neg;
    return r;
}
PB_HD Dec dec_neg(Dec a) {
    a.neg ^= 1;
    return a;
}
PB_HD bool dec_is_zero(const Dec& a) { return w_is_zero(a.c); }

// round an exact (coefficient, exponent) to prec digits, half-even ("mpd_qfinalize")
template <int N>
PB_HD Dec dec_round(Wide<N> c, i32 e, i32 neg, int prec, bool sticky = false) {
    int nd = w_ndigits(c);
    if (nd > prec) {
        int k = nd - prec;
        w_round_drop(c, k, sticky, (bool*)0);
        e += k;
        Wide<N> lim = w_pow10<N>(prec);
        if (w_cmp(c, lim) == 0) {   // 999..9 rounded up to 10^prec
            c = w_pow10<N>(prec - 1);
            e += 1;
        }
    }
    Dec r;
    r.c = w_resize<4>(c);
    r.e = e;
    r.neg = neg;
    return r;
}

PB_HDNI Dec dec_round8(const Wide<8>& c, i32 e, i32 neg, int prec, bool sticky) { return dec_round<8>(c, e, neg, prec, sticky); }

PB_HDNI Dec dec_mul(const Dec& a, const Dec& b, int prec = PB_PREC) {
    Wide<8> p = w_mul(a.c, b.c);
    return dec_round8(p, a.e + b.e, a.neg ^ b.neg, prec, false);
}

PB_HDNI Dec dec_add(const Dec& a_in, const Dec& b_in, int prec = PB_PREC) {
    bool az = dec_is_zero(a_in), bz = dec_is_zero(b_in);
    if (az && bz) {
        Dec r;
        w_zero(r.c);
        r.e = a_in.e < b_in.e ? a_in.e : b_in.e;
        r.neg = a_in.neg & b_in.neg;
        return r;
    }
    if (az || bz) {
        const Dec& x = az ? b_in : a_in;
        const Dec& z = az ? a_in : b_in;
        Wide<8> c = w_resize<8>(x.c);
        i32 e = x.e;
        if (z.e < x.e) {                       // pad with zeros towards the smaller exponent, up to prec digits
            int room = prec - w_ndigits(c);
            int sh = x.e - z.e;
            if (sh > room) sh = room;
            if (sh > 0) {
                w_mul_pow10(c, sh);
                e -= sh;
            }
        }
        return dec_round8(c, e, x.neg, prec, false);
    }
    Dec big = a_in, small = b_in;
    if (big.e < small.e) {
        big = b_in;
        small = a_in;
    }
    int shift = big.e - small.e;
    if (shift > 0) {
        int Db = w_ndigits(big.c), Ds = w_ndigits(small.c);
        int ex = big.e - 1 + ((Db > prec) ? 0 : Db - prec - 1);
        if (small.e + Ds - 1 < ex) {           // small only matters as a sticky digit (libmpdec _mpd_qaddsub)
            small.c = w_from_u64<4>(1);
            small.e = ex;
            shift = big.e - ex;
        }
    }
    Wide<8> B = w_resize<8>(big.c);
    w_mul_pow10(B, shift);
    Wide<8> S = w_resize<8>(small.c);
    i32 neg;
    if (big.neg == small.neg) {
        w_add(B, S);
        neg = big.neg;
    } else {
        int cmp = w_cmp(B, S);
        if (cmp == 0) {
            Dec r;
            w_zero(r.c);
            r.e = small.e;
            r.neg = 0;
            return r;
        }
        if (cmp > 0) {
            w_sub(B, S);
            neg = big.neg;
        } else {
            w_sub(S, B);
            B = S;
            neg = small.neg;
        }
    }
    return dec_round8(B, small.e, neg, prec, false);
}
PB_HD Dec dec_sub(const Dec& a, const Dec& b, int prec = PB_PREC) { return dec_add(a, dec_neg(b), prec); }

PB_HDNI Dec dec_div(const Dec& a, const Dec& b, int prec = PB_PREC) {
    Dec r;
    if (dec_is_zero(a)) {
        w_zero(r.c);
        r.e = a.e - b.e;
        r.neg = a.neg ^ b.neg;
        return r;
    }
    int Da = w_ndigits(a.c), Db = w_ndigits(b.c);
    int shift = Db - Da + prec + 1;            // > 0 because Da <= 38 < prec+1+Db is not guaranteed in general,
    Wide<8> A = w_resize<8>(a.c);              // but on this path operands carry <= prec digits
    Wide<4> Bv = b.c;
    if (shift > 0) w_mul_pow10(A, shift);
    Wide<8> Bw = w_resize<8>(Bv);
    if (shift < 0) w_mul_pow10(Bw, -shift);
    Wide<8> Q;
    Wide<8> R;
    w_divmod<8, 8>(A, Bw, Q, R);
    i32 e = a.e - b.e - shift;
    bool inexact = !w_is_zero(R);
    if (!inexact && shift > 0) {               // exact: move towards the ideal exponent a.e - b.e
        e += w_strip_zeros(Q, shift);
    }
    return dec_round8(Q, e, a.neg ^ b.neg, prec, inexact);
}

// a*b rounded half-even to W digits by ONE reciprocal multiplication (a has da digits, b has db):
//   P = a*b;  k = digits(P) - W;  P/10^k = P * floor(2^256/10^k) / 2^256 + [0, 2^-30)
// so quotient and rounding direction are known unless the top 32 fraction bits are within 2^-24 of
// 0, 1/2 or 1 -- then, or for k outside 20..47, *ok = false and the caller uses the exact dec_mul.
PB_HDNI Dec dec_mul_fast(const Dec& a, int da, const Dec& b, int db, int W, bool* ok) {
    Dec r;
    r.neg = a.neg ^ b.neg;
    r.e = 0;
    w_zero(r.c);
    *ok = false;
    Wide<8> P = w_mul(a.c, b.c);
    int nd = da + db - 1;
    if (nd >= PB_NPOW10 - 1) return r;
    {
        Wide<8> lim = w_pow10<8>(nd);
        if (w_cmp(P, lim) >= 0) nd += 1;
    }
    const int k = nd - W;
    if (k < 20 || k > 47 || P.w[7] != 0) return r;
    Wide<6> R;
#pragma unroll
    for (int i = 0; i < 6; i++) R.w[i] = TBL(inv10_256)[k][i];
    Wide<7> P7 = w_resize<7>(P);
    Wide<13> Q = w_mul(P7, R);
    if (Q.w[12] != 0) return r;
    Wide<4> I;
    I.w[0] = Q.w[8];
    I.w[1] = Q.w[9];
    I.w[2] = Q.w[10];
    I.w[3] = Q.w[11];
    const u32 fr = Q.w[7];
    if (fr >= 0x7FFFFF00u && !(fr > 0x80000000u && fr < 0xFFFFFF00u)) return r;
    if (fr > 0x80000000u) w_add_small(I, 1u);
    Wide<4> hi = w_pow10<4>(W), lo = w_pow10<4>(W - 1);
    i32 e = a.e + b.e + k;
    if (w_cmp(I, hi) == 0) {
        I = lo;
        e += 1;
    }
    if (w_cmp(I, lo) < 0 || w_cmp(I, hi) >= 0) return r;
    r.c = I;
    r.e = e;
    *ok = true;
    return r;
}

// value exactly one?
PB_HD bool dec_is_one_abs(const Dec& a) {
    if (a.e > 0 || a.e < -37) return false;
    Wide<4> p = w_pow10<4>(-a.e);
    return w_cmp(a.c, p) == 0;
}

// x ** n for a non-negative integer n (libmpdec mpd_qpow integer branch)
PB_HDNI Dec dec_powi(const Dec& x, u32 n, int prec = PB_PREC) {
    if (n == 0) return dec_from_u64(1);
    if (dec_is_one_abs(x)) {                   // _qcheck_pow_one: 1.000**3 = 1.000000000, at most prec digits
        i64 sh = (i64)n * (i64)(-x.e);
        if (sh > prec - 1) sh = prec - 1;
        Dec r;
        r.c = w_pow10<4>((int)sh);
        r.e = -(i32)sh;
        r.neg = (x.neg && (n & 1)) ? 1 : 0;
        return r;
    }
    int nd = 1;
    for (u32 t = n; t >= 10; t /= 10) nd++;
    int wprec = prec + nd + 2;
    Dec r = x;
    int top = 31;
    while (!((n >> top) & 1)) top--;
    const int dx = w_ndigits(x.c);
    int dr = dx;                                   // digits of r: wprec once a product has been rounded
    // square-and-multiply as ONE stream of products (r*r or r*x): lanes of a warp stay on the same
    // instruction whatever the bits of their exponents are
    int b = top - 1;
    bool pending = false;
    for (;;) {
        const bool mulx = pending;
        if (!mulx) {
            if (b < 0) break;
            pending = ((n >> b) & 1) != 0;
            b--;
        } else pending = false;
        const Dec& other = mulx ? x : r;
        const int dother = mulx ? dx : dr;
        bool ok;
        Dec t = dec_mul_fast(r, dr, other, dother, wprec, &ok);
        if (!ok) {
            t = dec_mul(r, other, wprec);
            dr = w_ndigits(t.c);
        } else dr = wprec;
        r = t;
    }
    i32 neg = (x.neg && (n & 1)) ? 1 : 0;
    return dec_round(r.c, r.e, neg, prec);
}

// Integer the shortest-path solver sees for an edge weight: the integer part of weight*1000
// (edges.py:17-23 prints str(weight*1000); fastpathz keeps whole numbers, CHANGELOG.md:13,57).
// Returns false if it does not fit N limbs minus a sign bit.
template <int N>
PB_HD bool dec_to_milli_int(const Dec& a, Wide<N>& mag) {
    Wide<8> c = w_resize<8>(a.c);
    int e = a.e + 3;
    if (e >= 0) {
        if (w_is_zero(c)) {
            w_zero(mag);
            return true;
        }
        if (w_ndigits(c) + e > 9 * N) return false;   // 10^(9N) < 2^(32N-1)
        if (w_ndigits(c) + e > 76) return false;
        w_mul_pow10(c, e);
    } else {
        int k = -e;
        if (k >= 39) {
            w_zero(c);
        } else {
            while (k > 9) {
                w_div_p10(c, 9);
                k -= 9;
            }
            w_div_p10(c, k);
        }
    }
    for (int i = N; i < 8; i++)
        if (c.w[i]) return false;
    mag = w_resize<N>(c);
    if (N <= 8 && (mag.w[N - 1] >> 31)) return false;
    return true;
}
