// Certified integer weights: what the shortest-path solve needs from an ORF or overlap edge is not its
// 28-digit Decimal weight but the INTEGER trunc(weight * 1000) (edges.py:22, phanotate.py:55-59,
// CHANGELOG.md:13,57).  That integer is decided here from a closed form with a rigorous error bound;
// only where the bound straddles an integer boundary (or the weight is astronomically large) does the
// item go through the literal chain (hold.cuh: every one of the reference's 28-digit multiplications
// replayed).  The literal chain also runs -- after the solve -- for the ORFs that were actually
// called, because their Decimal weight and '%E' score are output, and on demand for every ORF / edge
// when the caller asks for the ORF table or the edge dump (pb200_get_orfs, pb200_build_edges).
//
// ORF (functions.py:286-301, orfs.py:122-127).  With x = 1 - pstop (the literal 28-digit Decimal),
// n_k = number of codons of the ORF in GC-frame class k and e_k = pos_max[im]*pos_min[il]:
//     W_true = startw * Decimal(str(weight_rbs)) * x ** -(sum n_k e_k)
// The reference's value differs from W_true only by roundings, each bounded:
//     factor  F_k = (x ** pos_max) ** pos_min   two Decimal powers, <= 1 ulp each, pos_min <= 1  -> 2.001e-27
//     product hold *= F_k                       half an ulp per step                              -> 0.5e-27
//     1/hold, * startw, * weight_rbs            half an ulp each                                  -> 1.5e-27
// so |W_ref / W(x) - 1| <= (2.51 n + 1.51) * 1e-27 for n codon steps (tests/test_certified.py measures the
// actual ratio on the golden ORF tables: <= 4 % of the bound).  x itself is not needed as a Decimal either:
//     pstop_true = Pt Pa (Pa + 2 Pg) = nt na (na + 2 ng) / len^3             (orfs.py:162-173, exact rational)
//     pstop_ref  = pstop_true (1 +- 3.51e-27)    three quotients, five products, two sums, half an ulp each
//     x_ref      = x_true (1 +- 4.01e-27)        for x >= 1/2, incl. the rounding of 1 - pstop
// and W = C x^-E with E <= n turns that into another 4.01 n e-27: total (6.52 n + 1.51) e-27.  We use
// (7n + 8) * 2^-89 (2^-89 = 1.6e-27), which also swallows the ~2^-170 relative error of the Q32.192 evaluation.
#pragma once
#include "graph.cuh"

// number of set bits of M in [g0, g1) at positions = rg (mod 3), for five masks at once
PB_HD void count_frame_bits5(u64* const* M, i64 g0, i64 g1, int rg, u32* n) {
#pragma unroll
    for (int k = 0; k < 5; k++) n[k] = 0;
    if (g1 <= g0) return;
    const i64 w0 = g0 >> 6, w1 = (g1 - 1) >> 6;
    for (i64 w = w0; w <= w1; w++) {
        u64 m = frame_pat(w, rg);
        if (w == w0) m &= ~0ull << (g0 & 63);
        if (w == w1) m &= ~0ull >> (63 - ((g1 - 1) & 63));
#pragma unroll
        for (int k = 0; k < 5; k++) n[k] += (u32)pb_popc64(M[k][w] & m);
    }
}

// Per-contig / global fixed-point constants of the closed form.  item = contig*28 + r (RBS bin);
// items 0..8 of contig 0 also convert the start-codon weights (index 8 = no start codon -> 1000).
PB_HDN void st_fast_tables(const Batch& B, i64 item) {
    const i64 c = item / 28;
    if (c >= B.nc) return;
    const int r = (int)(item % 28);
    CStat* cs = B.cs + c;
    bool ok;
    Fx v = fx_from_dec(cs->wrbs[r], &ok);
    if (!ok || v.w[6] >= (1u << 20)) {
        cs->fast_ok = 0;               // (benign race: every writer stores 0)
        w_zero(v);
    }
    B.wr_fx[item] = v;
    if (c == 0 && r < 9) {
        Fx s;
        w_zero(s);
        s.w[6] = 1000;
        if (r < 8) {
            Dec d = B.P.startw[r];
            d.e += 3;
            bool o2;
            s = fx_from_dec(d, &o2);
            if (!o2 || d.neg || s.w[6] >= (1u << 11)) w_zero(s);     // zero = "not usable": the ORF goes literal
        }
        B.sw_fx[r] = s;
    }
}

// Certified integer weight of ORF oi, or a slot in the literal list.  item = ORF
PB_HDN void st_orf_fast(const Batch& B, i64 oi) {
    if (oi >= B.no) return;
    const int c = contig_of_orf(B, oi);
    const CStat* cs = B.cs + c;
    const int start = B.o_start[oi], stop = B.o_stop[oi];
    const bool rev = B.o_frame[oi] < 0;
    const int n = orf_steps(start, stop, rev);
    bool fast = cs->fast_ok && n > 0 && n < 100000;
    Wide<10> big;
    w_zero(big);
    // x_true = 1 - pstop_true, pstop_true = Pt Pa (Pa + 2 Pg) = nt na (na + 2 ng) / len^3 exactly (orfs.py:162-173)
    u32 na, nt, ng, len;
    orf_base_counts(B, oi, na, nt, ng, len);
    {
        U4 cn;
        cn.x = na;
        cn.y = nt;
        cn.z = ng;
        cn.w = len;
        B.o_cnt[oi] = cn;
    }
    Fx X;
    w_zero(X);
    if (len < 1 || len >= (1u << 17)) fast = false;
    if (fast) {
        const u64 num = (u64)nt * na * ((u64)na + 2ull * ng), den = (u64)len * len * len;
        if (num == 0 || num >= den) fast = false;
        else {
            Wide<9> N, Q;
            Wide<2> D, Rm;
            w_zero(N);
            N.w[6] = (u32)num;
            N.w[7] = (u32)(num >> 32);
            D.w[0] = (u32)den;
            D.w[1] = (u32)(den >> 32);
            w_divmod<9, 2>(N, D, Q, Rm);                       // floor(num * 2^192 / den) < 2^192
            X = fx_one();
            const Fx q = w_resize<FX_N>(Q);
            w_sub(X, q);
            if (!w_is_zero(Rm)) {                               // round the quotient up so that X <= x_true < X + 2^-192
                Fx ulp;
                w_zero(ulp);
                ulp.w[0] = 1;
                w_sub(X, ulp);
            }
            if (!(X.w[6] == 0 && (X.w[5] >> 31))) fast = false; // certified for 1/2 <= x < 1 only
        }
    }
    if (fast) {
        // codon counts per factor class over [start, stop) (forward) / (stop, start] (reverse), functions.py:289-298
        const i64 cb = B.coff[c];
        const i64 g0 = rev ? cb + stop : cb + start - 1, g1 = rev ? cb + start : cb + stop - 1;
        u32 nk[6];
        count_frame_bits5(rev ? B.cR : B.cF, g0, g1, (int)((cb + start - 1) % 3), nk);
        nk[5] = (u32)n - (nk[0] + nk[1] + nk[2] + nk[3] + nk[4]);
        if (nk[5] > (u32)n) fast = false;
        Fx E;
        w_zero(E);
#pragma unroll 1
        for (int k = 0; k < 6; k++) {
            Fx t = cs->fe[k];
            w_mul_small(t, nk[k]);
            w_add(E, t);
        }
        bool o1 = true, o2, o3;
        const SFx lnx = fx_ln(X, &o2);
        SFx T;
        T.m = fx_mul(lnx.m, E);
        T.neg = 0;                                             // exp(-E ln x), ln x < 0
        if (!o1 || !o2 || !lnx.neg || fx_to_double(T.m) > 60.0) fast = false;
        if (fast) {
            int K;
            const Fx P = fx_exp_core(T, &K, &o3);
            const int sw = B.o_sw[oi];
            const Fx SW = B.sw_fx[sw < 0 ? 8 : sw];
            if (!o3 || K < 0 || K > 90 || w_is_zero(SW)) fast = false;
            else {
                const Fx Mv = fx_mul(fx_mul(P, SW), B.wr_fx[(i64)c * 28 + B.o_rbs[oi]]);
                big = w_shl(w_resize<10>(Mv), K);              // |weight| * 1000 in Q128.192
            }
        }
    }
    if (fast) {
        Wide<4> I;
        I.w[0] = big.w[6];
        I.w[1] = big.w[7];
        I.w[2] = big.w[8];
        I.w[3] = big.w[9];
        const u64 fr = ((u64)big.w[5] << 32) | big.w[4];       // top 64 bits of the fraction
        const int bl = w_bitlen(I) + 1;                        // |weight|*1000 < 2^bl
        const u64 en = 7ull * (u64)n + 8ull;
        u64 errU;                                              // error bound in units of 2^-64: en * 2^(bl-89) * 2^64
        if (bl > 68) fast = false;
        else {
            errU = (bl >= 25) ? (en << (bl - 25)) : ((en >> (25 - bl)) + 1);
            errU += 2;
            if (errU >> 62) fast = false;
            else if (fr < errU || fr > ~errU) fast = false;
        }
        if (fast) {
            B.o_wint[oi] = wint_from_mag(w_resize<WN>(I), !w_is_zero(I));
            B.o_lit[oi] = 0;
        }
    }
    if (!fast) {
        const u32 pos = PB_ATOMIC_ADD_RET(&B.lit_cnt[0], 1u);
        B.lit_ids[pos] = (i32)oi;
        B.o_lit[oi] = 2;                                       // queued
    }
}

// ORFs on the path whose Decimal weight is still owed.  item = call
PB_HDN void st_lit_calls(const Batch& B, i64 k) {
    if (k >= B.ncalls) return;
    const i32 oi = B.call_orf[k];
    if (B.o_lit[oi] == 0) {
        const u32 pos = PB_ATOMIC_ADD_RET(&B.lit_cnt[1], 1u);
        B.lit_ids[pos] = oi;
        B.o_lit[oi] = 2;
    }
}
// every ORF without a literal weight (pb200_get_orfs / pb200_build_edges after a certified run).  item = ORF
PB_HDN void st_lit_rest(const Batch& B, i64 oi) {
    if (oi >= B.no) return;
    if (B.o_lit[oi] == 0) {
        const u32 pos = PB_ATOMIC_ADD_RET(&B.lit_cnt[1], 1u);
        B.lit_ids[pos] = (i32)oi;
        B.o_lit[oi] = 2;
    }
}

// ------------------------------------------------------------------------------------------------
// Overlap edges (functions.py:26-34,140-141,386): weight = 1/o**len (+20 if 'diff'), o = 1 - ave([o1,o2]).
// Closed form in double-double arithmetic (two IEEE doubles, ~2^-104 per operation):
//     W_true = (1 - (o1+o2)/2) ** -len  (+ 20)
// Roundings of the reference, each of relative size <= 0.5e-27 of its result:
//     o1, o2 themselves      -> pstop_ref = pstop_true (1 +- 3.51e-27) (above), pbar <= 1/2
//     o1+o2, /2, 1-pbar      -> |o_ref - o_true| <= 1.001e-27 + 1.76e-27, i.e. <= 5.52e-27 relative for
//                               o >= 0.5, amplified by len <= 502 in the power          -> 2.77e-24
//     _mpd_qpow_int at 33 digits (<= 16 products) + rounding to 28, 1/x, + 20          -> 1.52e-27
// total <= 2.78e-24 < 2^-78.2; the double-double evaluation adds < 2^-87.  The integer
// trunc(W*1000) is accepted when W*1000 (1 +- 2^-78) does not straddle an integer; else the literal
// chain (st_ov_pbar/pow/weight) computes it.
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

// pstop of the ORF behind node n as the exact rational nt na (na + 2 ng) / len^3 (counts left by st_orf_fast),
// or the contig's Decimal pstop for a node without an ORF (functions.py:373-385)
PB_HD bool dd_node_pstop(const Batch& B, int c, i32 n, DD& out) {
    const i32 oi = B.n_oidx[n];
    if (oi < 0) return dd_from_dec(B.cs[c].pstop, out);
    const U4 cn = B.o_cnt[oi];
    if (cn.w < 1 || cn.w >= (1u << 17)) return false;
    const double num = (double)cn.y * (double)cn.x * ((double)cn.x + 2.0 * (double)cn.z);   // < 2^53: exact
    const double den = (double)cn.w * (double)cn.w * (double)cn.w;                             // < 2^51: exact
    const double q1 = num / den;
    const double r = pb_fma(-q1, den, num);                                                    // exact remainder
    out = dd_quick(q1, r / den);
    return true;
}
// Certified integer weight of overlap edge k, or a slot in the literal list.  item = overlap edge
PB_HDN void st_ov_fast(const Batch& B, i64 k) {
    if (k >= B.nov) return;
    const i32 x = B.ov_src[k], e = B.ov_dst[k];
    const int c = contig_of_node(B, x);
    const int len = B.n_pos[x] - B.n_pos[e] + 3;
    DD o1, o2;
    bool fast = dd_node_pstop(B, c, e, o1) && dd_node_pstop(B, c, x, o2) && len >= 1 && len <= 502;
    i64 I = 0;
    if (fast) {
        DD pbar = dd_add(o1, o2);
        pbar.hi *= 0.5;
        pbar.lo *= 0.5;
        pbar.hi = -pbar.hi;
        pbar.lo = -pbar.lo;
        const DD o = dd_add(dd_from_d(1.0), pbar);
        if (!(o.hi >= 0.5 && o.hi <= 1.0)) fast = false;
        else {
            int top = 31;
            while (!((u32)len >> top)) top--;
            DD r = o;
            for (int b = top - 1; b >= 0; b--) {
                r = dd_mul(r, r);
                if ((len >> b) & 1) r = dd_mul(r, o);
            }
            DD w = dd_mul_d(dd_recip(r), 1000.0);
            if (B.ov_diff[k]) w = dd_add(w, dd_from_d(20000.0));
            if (!(w.hi >= 1.0 && w.hi < 4.0e18)) fast = false;
            else {
                const double ih = floor(w.hi), il = floor(w.lo);
                double f = (w.hi - ih) + (w.lo - il);
                I = (i64)ih + (i64)il;
                if (f >= 1.0) {
                    I += 1;
                    f -= 1.0;
                }
                const double thr = w.hi * 3.3087224502121107e-24 + 8.8817841970012523e-16;   // 2^-78, 2^-50
                if (!(f >= thr && f <= 1.0 - thr)) fast = false;
            }
        }
    }
    if (fast) B.ov_w64[k] = I;
    else {
        const u32 pos = PB_ATOMIC_ADD_RET(&B.lit_cnt[2], 1u);
        B.ovlit_ids[pos] = (i32)k;
    }
}
