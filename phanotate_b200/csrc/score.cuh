// Scoring stages: contig statistics, gap-score tables, per-ORF pstop / GC-frame product / weight.
// (functions.py:153-181, 253-301; orfs.py:122-127,162-173; functions.py:26-46)
#pragma once
#include "pipeline.cuh"

PB_HD Dec dec_one() { return dec_from_u64(1); }

// Pt*Pa*Pa + Pt*Pg*Pa + Pt*Pa*Pg, left to right, every operation rounded (functions.py:178, orfs.py:173)
PB_HDNI Dec pstop_formula(const Dec& Pa, const Dec& Pt, const Dec& Pg) {
    Dec t1 = dec_mul(dec_mul(Pt, Pa), Pa);
    Dec t2 = dec_mul(dec_mul(Pt, Pg), Pa);
    Dec t3 = dec_mul(dec_mul(Pt, Pa), Pg);
    return dec_add(dec_add(t1, t2), t3);
}

// Stage 5: per-contig statistics.  item = contig
PB_HDN void st_contig_stats(const Batch& B, i64 c) {
    if (c >= B.nc) return;
    CStat* cs = B.cs + c;
    const int L = (int)(B.coff[c + 1] - B.coff[c]);
    cs->L = L;
    if (L < 1) {
        cs->err |= ERR_RANGE;
        return;
    }
    // base frequencies over both strands: fa = ft = #a+#t, fg = fc = #g+#c (functions.py:159-166,174-178)
    Dec twoL = dec_from_u64(2ull * (u64)L);
    Dec Pa = dec_div(dec_from_u64(cs->nAT), twoL);
    Dec Pg = dec_div(dec_from_u64(cs->nGC), twoL);
    cs->pstop = pstop_formula(Pa, Pa, Pg);
    cs->g = dec_sub(dec_one(), cs->pstop);
    cs->g100 = dec_powi(cs->g, 100);
    {   // score_gap(len > 300) = g**100 + len: the rounding position only depends on the digit count of len
        Wide<2> m;
        bool o3 = dec_to_milli_int<2>(dec_add(cs->g100, dec_from_u64(301)), m);
        cs->gap_hi3 = (i64)(((u64)m.w[1] << 32) | m.w[0]) - 301000;
        bool o4 = dec_to_milli_int<2>(dec_add(cs->g100, dec_from_u64(1000)), m);
        cs->gap_hi4 = (i64)(((u64)m.w[1] << 32) | m.w[0]) - 1000000;
        if (!o3 || !o4) cs->err |= ERR_OVERFLOW;
    }
    // (the RBS likelihood ratios per score bin are converted by st_rbs_weights, 28 threads per contig;
    //  ln g by st_contig_lng, only where the literal gap tables are built)
    u32 nz = 0;
    for (int r = 1; r < 28; r++) nz += cs->hist_bg[r];
    cs->hist_bg[0] = 2u * (u32)L - nz;
    // GC-frame exponents (functions.py:262-263,281-284): counts start at 1, divided by their max
    u32 mx = 1, mn = 1;
    for (int k = 1; k < 4; k++) {
        if (cs->cmax[k] + 1 > mx) mx = cs->cmax[k] + 1;
        if (cs->cmin[k] + 1 > mn) mn = cs->cmin[k] + 1;
    }
    for (int k = 0; k < 4; k++) {
        u32 a = (k ? cs->cmax[k] : 0) + 1, b = (k ? cs->cmin[k] : 0) + 1;
        cs->pos_max[k] = dec_div(dec_from_u64(a), dec_from_u64(mx));
        cs->pos_min[k] = dec_div(dec_from_u64(b), dec_from_u64(mn));
        cs->max_one[k] = (a == mx);
        cs->min_one[k] = (b == mn);
        bool o1, o2;
        cs->fmax[k] = fx_from_dec(cs->pos_max[k], &o1);
        cs->fmin[k] = fx_from_dec(cs->pos_min[k], &o2);
        if (!o1 || !o2) cs->err |= ERR_RANGE;
    }
    cs->fast_ok = (cs->err & ERR_RANGE) ? 0 : 1;
    for (int k = 0; k < 6; k++) {
        const int im = k / 2 + 1, il = (k % 2) + 1 + (((k % 2) + 1 >= im) ? 1 : 0);
        DD a, b;
        if (!dd_from_dec(cs->pos_max[im], a) || !dd_from_dec(cs->pos_min[il], b)) cs->fast_ok = 0;
        cs->fe[k] = dd_mul(a, b);
    }
}

// ln g in fixed point for the real-exponent gap scores.  item = contig
PB_HDN void st_contig_lng(const Batch& B, i64 c) {
    if (c >= B.nc) return;
    CStat* cs = B.cs + c;
    if (cs->L < 1) return;
    bool ok = true, ok2 = true;
    if (dec_is_one_abs(cs->g)) {
        w_zero(cs->ln_g.m);
        cs->ln_g.neg = 0;
    } else {
        Fx X = fx_from_dec(cs->g, &ok);
        cs->ln_g = fx_ln(X, &ok2);
    }
    if (!ok || !ok2) PB_ATOMIC_OR(&cs->err, (u32)ERR_RANGE);
}
// RBS likelihood ratio per score bin, IEEE doubles, then Decimal(str(float)) (functions.py:155-156,180-181,254-257;
// orfs.py:126).  item = contig*28 + bin
PB_HDN void st_rbs_weights(const Batch& B, i64 item) {
    const i64 c = item / 28;
    if (c >= B.nc) return;
    const int r = (int)(item % 28);
    CStat* cs = B.cs + c;
    if (cs->L < 1) return;
    const double ybg = 28.0 + 2.0 * (double)cs->L;
    const double ytr = 28.0 + (double)(u32)(B.corf[c + 1] - B.corf[c]);
    const double bg = (1.0 + (double)cs->hist_bg[r]) / ybg;
    const double tr = (1.0 + (double)cs->hist_tr[r]) / ytr;
    bool okr;
    cs->wrbs[r] = dec_from_double_repr(tr / bg, &okr);
    if (!okr) PB_ATOMIC_OR(&cs->err, (u32)ERR_RANGE);
}

// score_gap for length <= 300 (functions.py:36-46): 1/g**Decimal(length/3) (+ 1/0.05 if 'diff'), as three
// small kernels: integer exponents (length % 3 == 0), real exponents, and the reciprocal / +20 / integer part.
PB_HD Dec dec_twenty() {   // 1/Decimal('0.05') == Decimal('2E+1')
    Dec d = dec_from_u64(2);
    d.e = 1;
    return d;
}
// item = contig*101 + m, length = 3m
PB_HDN void st_gap_pow_int(const Batch& B, i64 item) {
    const i64 c = item / 101;
    if (c >= B.nc) return;
    const int m = (int)(item % 101);
    const CStat* cs = B.cs + c;
    if (cs->L < 1) return;
    B.gap_same[c * GAPN + (3 * m + 2)] = dec_powi(cs->g, (u32)m);
}
// item = contig*202 + j, length = 3*(j/2) - 2 + (j&1)  (the lengths in -2..299 that are not multiples of 3)
PB_HDN void st_gap_pow_real(const Batch& B, i64 item) {
    const i64 c = item / 202;
    if (c >= B.nc) return;
    const int j = (int)(item % 202);
    const int len = 3 * (j >> 1) - 2 + (j & 1);
    CStat* cs = B.cs + c;
    if (cs->L < 1) return;
    const double y = (double)len / 3.0;                   // Python float; Decimal(float) is exact
    const Fx Y = fx_from_double(fabs(y));
    Dec pw;
    bool ok = true;
    if (dec_is_one_abs(cs->g)) pw = dec_pow_fx(cs->g, Y, y < 0, PB_PREC, &ok);
    else {
        SFx T;
        T.m = fx_mul(cs->ln_g.m, Y);
        T.neg = cs->ln_g.neg ^ (y < 0 ? 1 : 0);
        bool o3, o4;
        Fx V = fx_exp(T, &o3);
        pw = fx_to_dec(V, PB_PREC, &o4);
        ok = o3 && o4;
    }
    if (!ok) PB_ATOMIC_OR(&cs->err, (u32)ERR_RANGE);
    B.gap_same[c * GAPN + (len + 2)] = pw;
}
// item = contig*GAPN + (len+2): gap_same holds g**(len/3) on entry
PB_HDN void st_gap_lut(const Batch& B, i64 item) {
    i64 c = item / GAPN;
    if (c >= B.nc) return;
    CStat* cs = B.cs + c;
    if (cs->L < 1) return;
    Dec same = dec_div(dec_one(), B.gap_same[item]);
    Dec diff = dec_add(same, dec_twenty());
    B.gap_same[item] = same;
    B.gap_diff[item] = diff;
    Wide<2> m;
    bool f1 = dec_to_milli_int<2>(same, m);
    B.gapi_same[item] = (i64)(((u64)m.w[1] << 32) | m.w[0]);
    bool f2 = dec_to_milli_int<2>(diff, m);
    B.gapi_diff[item] = (i64)(((u64)m.w[1] << 32) | m.w[0]);
    if (!f1 || !f2) PB_ATOMIC_OR(&cs->err, (u32)ERR_OVERFLOW);
}
// score_gap for any length as a Dec (used for terminals, bridges and the edge dump)
PB_HDN Dec gap_score(const Batch& B, int c, int len, bool diff) {
    if (len > 300) return dec_add(B.cs[c].g100, dec_from_u64((u64)len));   // functions.py:40-41 (no inversion, no +20)
    i64 k = (i64)c * GAPN + (len + 2);
    return diff ? B.gap_diff[k] : B.gap_same[k];
}

PB_HD WInt wint_from_mag(const Wide<WN>& mag, bool neg) {
    if (!neg) return mag;
    WInt r;
    u64 carry = 1;
#pragma unroll
    for (int i = 0; i < WN; i++) {
        carry += (u64)(~mag.w[i]);
        r.w[i] = (u32)carry;
        carry >>= 32;
    }
    return r;
}
PB_HD WInt wint_from_i64(i64 v) {
    WInt r;
    u32 ext = v < 0 ? 0xFFFFFFFFu : 0u;
    r.w[0] = (u32)(u64)v;
    r.w[1] = (u32)((u64)v >> 32);
#pragma unroll
    for (int i = 2; i < WN; i++) r.w[i] = ext;
    return r;
}
PB_HD bool wint_less(const WInt& a, const WInt& b) {   // signed compare
    u32 sa = a.w[WN - 1] >> 31, sb = b.w[WN - 1] >> 31;
    if (sa != sb) return sa > sb;
    return w_cmp(a, b) < 0;
}
PB_HD bool wint_is_inf(const WInt& a) { return a.w[WN - 1] == 0x7FFFFFFFu; }
PB_HD WInt wint_inf() {
    WInt r;
#pragma unroll
    for (int i = 0; i < WN; i++) r.w[i] = 0xFFFFFFFFu;
    r.w[WN - 1] = 0x7FFFFFFFu;
    return r;
}
// |w| < 2^110 ?  (sums over a contig's <= 2^14-edge paths then stay inside 128 bits)
PB_HD bool wint_is_narrow(const WInt& w) {
    const u32 ext = (w.w[WN - 1] >> 31) ? 0xFFFFFFFFu : 0u;
    bool ok = true;
#pragma unroll
    for (int i = 4; i < WN; i++) ok = ok && w.w[i] == ext;
    const u32 top = w.w[3] ^ ext;               // bits 96..127 relative to the sign
    return ok && (top >> 14) == 0;
}
// Weights beyond 2^240 (hold.cuh: an ORF of ~9 kb in AT-rich sequence already weighs 1e108; the reference's Decimal and
// the GMP-backed fastpathz have no limit): o_wint holds a marker, the value is formed from the Decimal weight when the solve
// needs it, as a 2048-bit integer (|w| < ~1e560; a double overflows at 1.8e308).
#define WH 64
typedef Wide<WH> HInt;
PB_HD WInt wint_huge_marker() {
    WInt r;
#pragma unroll
    for (int i = 0; i < WN; i++) r.w[i] = 0x48554745u;
    r.w[WN - 1] = 0x7FFFFFFEu;
    return r;
}
PB_HD bool wint_is_huge_marker(const WInt& a) { return a.w[WN - 1] == 0x7FFFFFFEu; }
// trunc(d * 1000) as a two's complement 1280-bit integer; false if even that is too narrow
PB_HDNI bool dec_to_hint(const Dec& d, HInt& out) {
    HInt mag = w_resize<WH>(d.c);
    const int e = d.e + 3;
    if (e < 0 || w_ndigits(d.c) + e > 9 * WH - 12) return false;
    w_mul_pow10(mag, e);
    if (d.neg && !w_is_zero(mag)) {
        u64 carry = 1;
        for (int i = 0; i < WH; i++) {
            carry += (u64)(~mag.w[i]);
            mag.w[i] = (u32)carry;
            carry >>= 32;
        }
    }
    out = mag;
    return true;
}
PB_HDNI bool dec_to_wint(const Dec& d, WInt& out) {
    Wide<WN> mag;
    bool ok = dec_to_milli_int<WN>(d, mag);
    if (ok && (mag.w[WN - 1] >> 16)) ok = false;      // |w| < 2^240: sums of 2^14 weights stay below the INF marker
    out = wint_from_mag(mag, d.neg && !w_is_zero(mag));
    return ok;
}

