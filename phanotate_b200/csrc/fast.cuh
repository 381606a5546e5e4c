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
// (7n + 8) * 2^-89 (2^-89 = 1.6e-27) plus 2^-90 for the double-double evaluation of the closed form
// (ddmath.cuh: < 2^-92).  The fast path is taken for 0 < pstop <= 1/8 and |weight|*1000 < 2^68.
#pragma once
#include "graph.cuh"

// Per-contig / global double-double constants of the closed form.  item = contig*28 + r (RBS bin);
// items 0..8 of contig 0 also convert the start-codon weights (index 8 = no start codon -> 1000).
PB_HDN void st_fast_tables(const Batch& B, i64 item) {
    const i64 c = item / 28;
    if (c >= B.nc) return;
    const int r = (int)(item % 28);
    CStat* cs = B.cs + c;
    DD v;
    if (!dd_from_dec(cs->wrbs[r], v) || !(v.hi < 1.0e9)) {
        cs->fast_ok = 0;               // (benign race: every writer stores 0)
        v = dd_from_d(0.0);
    }
    B.wr_dd[item] = v;
    if (c == 0 && r < 9) {
        DD s = dd_from_d(1000.0);
        if (r < 8) {
            DD w;
            if (dd_from_dec(B.P.startw[r], w) && w.hi <= 1.0e6) s = dd_mul_d(w, 1000.0);
            else s = dd_from_d(0.0);   // zero = "not usable": the ORF goes literal
        }
        B.sw_dd[r] = s;
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
    DD V = dd_from_d(0.0);              // |weight| * 1000
    if (len < 1 || len >= (1u << 17)) fast = false;
    if (fast) {
        const double num = (double)nt * (double)na * ((double)na + 2.0 * (double)ng);   // < 2^53: exact
        const double den = (double)len * (double)len * (double)len;                      // < 2^51: exact
        if (!(num > 0.0 && num * 8.0 <= den)) fast = false;                             // certified for 0 < pstop <= 1/8
        else {
            const double q1 = num / den;
            const DD p = dd_quick(q1, pb_fma(-q1, den, num) / den);
            const DD Lx = dd_neglog1m(p);                                               // -ln(1 - pstop) > 0
            // codon counts per factor class over [start, stop) (forward) / (stop, start] (reverse), functions.py:289-298
            const i64 cb = B.coff[c];
            const i64 g0 = rev ? cb + stop : cb + start - 1, g1 = rev ? cb + start : cb + stop - 1;
            u32 nk[6];
            count_frame_bits5(rev ? B.cR : B.cF, g0, g1, (int)((cb + start - 1) % 3), nk);
            nk[5] = (u32)n - (nk[0] + nk[1] + nk[2] + nk[3] + nk[4]);
            if (nk[5] > (u32)n) fast = false;
            DD E = dd_from_d(0.0);
#pragma unroll 1
            for (int k = 0; k < 6; k++) E = dd_add(E, dd_mul_d(cs->fe[k], (double)nk[k]));
            const DD T = dd_mul(E, Lx);                                                 // weight = C * exp(T)
            const int sw = B.o_sw[oi];
            const DD SW = B.sw_dd[sw < 0 ? 8 : sw];
            if (!(T.hi >= 0.0 && T.hi < 60.0) || !(SW.hi > 0.0)) fast = false;
            if (fast) {
                int K;
                const DD P = dd_exp_split(T, &K);
                const DD Mv = dd_mul(dd_mul(P, SW), B.wr_dd[(i64)c * 28 + B.o_rbs[oi]]);
                V.hi = ldexp(Mv.hi, K);
                V.lo = ldexp(Mv.lo, K);
            }
        }
    }
    if (fast && !(V.hi >= 0.0 && V.hi < 2.9e20)) fast = false;                          // < 2^68
    if (fast) {
        // integer part and fraction of hi + lo
        const double ih = floor(V.hi), il = floor(V.lo);
        double f = (V.hi - ih) + (V.lo - il);                                           // in [0, 2)
        i64 adj = (i64)il;
        if (f >= 1.0) {
            f -= 1.0;
            adj += 1;
        }
        const double top = floor(ih * 2.3283064365386963e-10);                          // ih / 2^32, exact
        const u64 low = (u64)(ih - top * 4294967296.0), t64 = (u64)top;                 // ih = t64 * 2^32 + low, t64 < 2^36
        u64 lo64 = (t64 << 32) + low, hi64 = t64 >> 32;
        const u64 old = lo64;
        lo64 += (u64)adj;
        if (adj >= 0) hi64 += (lo64 < old) ? 1u : 0u;
        else hi64 -= (lo64 > old) ? 1u : 0u;
        // relative error bound (7n + 8) 2^-89 of the reference's roundings + 2^-90 for the evaluation, absolute 2^-36 for f
        const double thr = V.hi * (((double)(7 * n + 8) + 0.5) * 1.6155871338926322e-27) + 1.4551915228366852e-11;
        if (!(thr < 0.25 && f >= thr && f <= 1.0 - thr) || (hi64 >> 8)) fast = false;
        if (fast) {
            Wide<WN> I;
            w_zero(I);
            I.w[0] = (u32)lo64;
            I.w[1] = (u32)(lo64 >> 32);
            I.w[2] = (u32)hi64;
            I.w[3] = (u32)(hi64 >> 32);
            B.o_wint[oi] = wint_from_mag(I, !w_is_zero(I));
            B.o_v[oi] = V;
            B.o_lit[oi] = 0;
        }
    }
    if (!fast) {
        const u32 pos = PB_ATOMIC_ADD_RET(&B.lit_cnt[0], 1u);
        B.lit_ids[pos] = (i32)oi;
        B.o_lit[oi] = 2;                                       // queued
    }
}

// float(Orf.weight) (what '%E' prints, phanotate.py:75-76) from the closed form: W = V/1000 = hi + lo is within
// rel = (7n + 8.5) 2^-89 + 2^-100 of the reference's Decimal, so hi is ITS nearest double as soon as
// |lo| + rel |hi| stays clear of half an ulp of hi (a quarter below a power of two).  Fails ~2^-26 of the time.
PB_HD bool certified_score(const DD& V, int n, double* score) {
    if (!(V.hi > 0.0)) return false;
    const DD W = dd_mul(V, dd_table(TBL(p10neg_dd)[3]));
    const double hi = W.hi, alo = fabs(W.lo);
    int ex;
    const double m = frexp(hi, &ex);                              // hi = m 2^ex, m in [0.5, 1): ulp = 2^(ex-53)
    if (ex < -900 || ex > 900) return false;
    const double half = ldexp(1.0, ex - 54), room = (m == 0.5) ? 0.5 * half : half;
    const double rel = ((double)(7 * n + 8) + 0.5) * 1.6155871338926322e-27 + 7.8886090522101181e-31;
    if (!(alo + hi * rel * 1.0000001 < room * 0.9999999)) return false;
    *score = -hi;
    return true;
}
// Calls whose score cannot be certified (or every call when PB200_CALL_WEIGHTS asks for the Decimal weights)
// go through the literal chain after the solve.  item = call
PB_HDN void st_lit_calls(const Batch& B, i64 k) {
    if (k >= B.ncalls) return;
    const i32 oi = B.call_orf[k];
    if (oi < 0) return;                       // a tRNA call: weight -20, nothing owed
    if (B.o_lit[oi] == 0 && !(B.flags & PB200_CALL_WEIGHTS)) {
        const bool rev = B.o_frame[oi] < 0;
        double sc;
        if (certified_score(B.o_v[oi], orf_steps(B.o_start[oi], B.o_stop[oi], rev), &sc)) {
            B.call_score[k] = sc;
            return;
        }
    }
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
// Gap edges (functions.py:36-46) for 0 <= len <= 300: trunc(1000 / g**(len/3)) (+ 20000 for 'diff') from
// exp((len/3) * -ln g) in double-double arithmetic, g = 1 - pstop of the contig (its literal Decimal).  Roundings of
// the reference: the power (<= 1 ulp), 1/x, + 20 (half an ulp each): <= 2.01e-27 relative; accepted when the value
// (1 +- 2^-86) does not straddle an integer, else (and for g outside [7/8, 1)) the Decimal arithmetic runs
// right here.  The Decimal tables gap_same / gap_diff themselves are only built for the edge dump.  item = contig*GAPN + len+2
PB_HDN void st_gap_fast(const Batch& B, i64 item) {
    const i64 c = item / GAPN;
    if (c >= B.nc) return;
    CStat* cs = B.cs + c;
    if (cs->L < 1) return;
    const int len = (int)(item % GAPN) - 2;
    DD g;
    bool fast = dd_from_dec(cs->g, g) && g.hi < 1.0 && g.hi >= 0.875;
    i64 vs = 0, vd = 0;
    if (fast) {
        const DD p = dd_add(dd_from_d(1.0), dd_neg(g));
        const DD T = dd_mul_d(dd_neglog1m(p), fabs((double)len / 3.0));    // Decimal(len/3) is exactly this double
        int K;
        DD P = dd_exp_split(T, &K);
        P.hi = ldexp(P.hi, K);
        P.lo = ldexp(P.lo, K);
        if (len < 0) P = dd_recip(P);                                      // 1 / g**(len/3) = g**(|len|/3)
        const DD V = dd_mul_d(P, 1000.0);
#pragma unroll 1
        for (int d = 0; d < 2 && fast; d++) {
            const DD W = d ? dd_add(V, dd_from_d(20000.0)) : V;
            if (!(W.hi >= 1.0 && W.hi < 1.0e15)) fast = false;
            else {
                const double ih = floor(W.hi), il = floor(W.lo);
                double f = (W.hi - ih) + (W.lo - il);
                i64 I = (i64)ih + (i64)il;
                if (f >= 1.0) {
                    I += 1;
                    f -= 1.0;
                }
                const double thr = W.hi * 1.2924697071141057e-26 + 8.8817841970012523e-16;   // 2^-86, 2^-50
                if (!(f >= thr && f <= 1.0 - thr)) fast = false;
                if (d) vd = I;
                else vs = I;
            }
        }
    }
    if (!fast) {                                   // the reference's Decimal arithmetic for this one entry
        Dec pw;
        bool ok = true;
        if (len % 3 == 0 && len >= 0) pw = dec_powi(cs->g, (u32)(len / 3));
        else {
            const double y = (double)len / 3.0;
            pw = dec_pow_fx(cs->g, fx_from_double(fabs(y)), y < 0, PB_PREC, &ok);
        }
        if (!ok) PB_ATOMIC_OR(&cs->err, (u32)ERR_RANGE);
        const Dec same = dec_div(dec_one(), pw), diff = dec_add(same, dec_twenty());
        Wide<2> m;
        const bool f1 = dec_to_milli_int<2>(same, m);
        vs = (i64)(((u64)m.w[1] << 32) | m.w[0]);
        const bool f2 = dec_to_milli_int<2>(diff, m);
        vd = (i64)(((u64)m.w[1] << 32) | m.w[0]);
        if (!f1 || !f2) PB_ATOMIC_OR(&cs->err, (u32)ERR_OVERFLOW);
    }
    B.gapi_same[item] = vs;
    B.gapi_diff[item] = vd;
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
