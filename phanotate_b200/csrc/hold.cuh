// ORF scoring, split for the GPU (functions.py:286-301, orfs.py:122-127,162-173):
//
//   st_orf_factors  per ORF, uniform work: pstop, the six factors ((1-pstop)**pos_max[i])**pos_min[j]
//                   (i != j), each also in "reciprocal-scaled" form for the fast multiply; length bin.
//   st_len_scatter  counting sort of the ORFs by codon count (longest first) so that the lanes of
//                   a warp run the same number of product steps.
//   hold_run        per ORF in sorted order: hold *= factor[class(codon)] for every codon, then
//                   Orf.score().  One step is one 28-digit Decimal multiplication with half-even
//                   rounding; the fast path does it with ONE 3x4-limb product:
//                       a*b/10^s = a * floor(b*2^124/10^s) / 2^124 + [0, 2^-30)
//                   so integer part and rounding direction are known unless the top 32 fraction bits
//                   fall within 2^-27 of 0, 1/2 or 1 -- then (p ~ 1e-8) the exact generic dec_mul runs.
#pragma once
#include "score.cuh"

struct HoldFac {      // one factor b (28 digits) prepared for the fast multiply; 48 bytes = 3 x 16
    u32 c27[4];       // floor(b * 2^124 / 10^27)
    u32 c28[4];       // floor(b * 2^124 / 10^28)
    u32 btop[2];      // b >> 30
    i32 e;            // exponent of b
    u32 ok;           // 1 if b has exactly 28 digits (fast path usable)
};
#define HOLD_BINS 2048

PB_HD int fac_index(int imax, int imin) { return (imax - 1) * 2 + (imin - 1) - (imin > imax ? 1 : 0); }
PB_HD int orf_steps(int start, int stop, bool rev) { return rev ? (start - stop + 2) / 3 : (stop - start + 2) / 3; }

PB_HDNI void holdfac_prepare(const Dec& b, HoldFac& f) {
    const u32 R27[6] = {0x524f8e02u, 0x7aa9a3eeu, 0xbaf51326u, 0x8f03f243u, 0xf3a68dbcu, 0x00000004u};   // 2^252/10^27
    const u32 R28[6] = {0x3b6e5b00u, 0x0c4429feu, 0xc5e54eb7u, 0x41806506u, 0x7ec3daf9u, 0x00000000u};   // 2^252/10^28
    // a factor with fewer than 28 digits (e.g. 1 - pstop = 0.95968 exactly) is scaled to 28 digits: the
    // rounded product with a 28-digit multiplicand only depends on its value
    Wide<4> c4 = b.c;
    f.e = b.e;
    int nd = w_ndigits(c4);
    if (nd >= 1 && nd < 28) {
        w_mul_pow10(c4, 28 - nd);
        f.e -= 28 - nd;
    }
    Wide<4> lo = w_pow10<4>(27), hi = w_pow10<4>(28);
    f.ok = (!b.neg && w_cmp(c4, lo) >= 0 && w_cmp(c4, hi) < 0) ? 1u : 0u;
    Wide<3> bc = w_resize<3>(c4);
    Wide<6> r;
#pragma unroll
    for (int i = 0; i < 6; i++) r.w[i] = R27[i];
    Wide<9> p = w_mul(bc, r);
#pragma unroll
    for (int i = 0; i < 4; i++) f.c27[i] = p.w[i + 4];          // >> 128
#pragma unroll
    for (int i = 0; i < 6; i++) r.w[i] = R28[i];
    p = w_mul(bc, r);
#pragma unroll
    for (int i = 0; i < 4; i++) f.c28[i] = p.w[i + 4];
    Wide<4> t = w_shr(c4, 30);
    f.btop[0] = t.w[0];
    f.btop[1] = t.w[1];
}

// one multiplication hold * b.  a = coefficient of hold (28 digits, three limbs), eh its exponent.
// Returns false when the rounding cannot be decided from 32 fraction bits (caller runs the exact path).
PB_HD bool hold_step_fast(u32& a0, u32& a1, u32& a2, i32& eh, const U4& c27, const U4& c28, const U4& misc) {
    // which power of ten is dropped: a*b >= 10^55 <=> 28 digits are dropped
    const u64 atop = ((u64)a2 << 34) | ((u64)a1 << 2) | (a0 >> 30);
    const u64 btop = ((u64)misc.y << 32) | misc.x;
#ifdef __CUDA_ARCH__
    const u64 hi = __umul64hi(atop, btop);
#else
    const u64 hi = (u64)(((unsigned __int128)atop * btop) >> 64);
#endif
    const bool big = hi >= 0x06867a5a867f103bull;        // floor(10^55 / 2^124)
    const u32 c0 = big ? c28.x : c27.x, c1 = big ? c28.y : c27.y, c2 = big ? c28.z : c27.z, c3 = big ? c28.w : c27.w;
    // p = a * c, seven limbs, column by column
    u32 p[7];
    {
        u64 acc;
        u32 hi3;
        u64 t;
#define MAC(x, y)                  \
    t = (u64)(x) * (y);            \
    acc += t;                      \
    hi3 += (acc < t) ? 1u : 0u;
#define NEXT(k)                              \
    p[k] = (u32)acc;                         \
    acc = (acc >> 32) | ((u64)hi3 << 32);    \
    hi3 = 0;
        acc = 0;
        hi3 = 0;
        MAC(a0, c0) NEXT(0)
        MAC(a0, c1) MAC(a1, c0) NEXT(1)
        MAC(a0, c2) MAC(a1, c1) MAC(a2, c0) NEXT(2)
        MAC(a0, c3) MAC(a1, c2) MAC(a2, c1) NEXT(3)
        MAC(a1, c3) MAC(a2, c2) NEXT(4)
        MAC(a2, c3) NEXT(5)
        p[6] = (u32)acc;
#undef MAC
#undef NEXT
    }
    // integer part = p >> 124, fraction top 32 bits = bits 92..123
    u32 i0 = (p[3] >> 28) | (p[4] << 4), i1 = (p[4] >> 28) | (p[5] << 4), i2 = (p[5] >> 28) | (p[6] << 4);
    const u32 i3 = p[6] >> 28;
    const u32 fr = (p[2] >> 28) | (p[3] << 4);
    // 10^27 <= I < 10^28 must hold (otherwise the scale guess was off by one ulp: exact path)
    const bool ge27 = (i2 > 0x033b2e3cu) || (i2 == 0x033b2e3cu && (i1 > 0x9fd0803cu || (i1 == 0x9fd0803cu && i0 >= 0xe8000000u)));
    const bool lt28 = (i2 < 0x204fce5eu) || (i2 == 0x204fce5eu && (i1 < 0x3e250261u || (i1 == 0x3e250261u && i0 < 0x10000000u)));
    if (i3 != 0 || !ge27 || !lt28) return false;
    if (fr < 0x7FFFFF00u) {
        // round down
    } else if (fr > 0x80000000u && fr < 0xFFFFFF00u) {
        i0 += 1;                                   // round up; carry
        if (i0 == 0) {
            i1 += 1;
            if (i1 == 0) i2 += 1;
        }
        if (i2 == 0x204fce5eu && i1 == 0x3e250261u && i0 == 0x10000000u) {   // 10^28 -> 10^27, exponent + 1
            i0 = 0xe8000000u;
            i1 = 0x9fd0803cu;
            i2 = 0x033b2e3cu;
            eh += 1;
        }
    } else {
        return false;
    }
    a0 = i0;
    a1 = i1;
    a2 = i2;
    eh += (i32)misc.z + (big ? 28 : 27);
    return true;
}

// a * b for 28-digit positive operands through hold_step_fast (b prepared in fb), else the generic multiply
PB_HD Dec dec_mul28(const Dec& a, const Dec& b, const HoldFac& fb) {
    if (fb.ok && a.c.w[3] == 0 && !a.neg && !b.neg) {
        Wide<4> lo = w_pow10<4>(27);
        if (w_cmp(a.c, lo) >= 0) {           // < 10^28 is implied by prec 28
            u32 a0 = a.c.w[0], a1 = a.c.w[1], a2 = a.c.w[2];
            i32 eh = a.e;
            U4 c27, c28, misc;
            c27.x = fb.c27[0]; c27.y = fb.c27[1]; c27.z = fb.c27[2]; c27.w = fb.c27[3];
            c28.x = fb.c28[0]; c28.y = fb.c28[1]; c28.z = fb.c28[2]; c28.w = fb.c28[3];
            misc.x = fb.btop[0]; misc.y = fb.btop[1]; misc.z = (u32)fb.e; misc.w = fb.ok;
            if (hold_step_fast(a0, a1, a2, eh, c27, c28, misc)) {
                Dec r;
                r.c.w[0] = a0; r.c.w[1] = a1; r.c.w[2] = a2; r.c.w[3] = 0;
                r.e = eh;
                r.neg = 0;
                return r;
            }
        }
    }
    return dec_mul(a, b);
}
// Decimal(a) / Decimal(b) for small non-negative integers (orfs.py:169-172: count / Decimal(len(seq))):
// the General Decimal Arithmetic division with a one-limb divisor, magic = floor((2^64-1)/b)
PB_HDNI Dec dec_div_u32(u32 a, u32 b, u64 magic, int prec) {
    Dec r;
    w_zero(r.c);
    r.e = 0;
    r.neg = 0;
    if (a == 0) return r;
    int Da = 1, Db = 1;
    for (u32 t = a; t >= 10; t /= 10) Da++;
    for (u32 t = b; t >= 10; t /= 10) Db++;
    const int shift = Db - Da + prec + 1;
    Wide<5> A = w_from_u64<5>(a);
    w_mul_pow10(A, shift);
    u32 rem = 0;
#pragma unroll
    for (int i = 4; i >= 0; i--) {
        u64 x = ((u64)rem << 32) | A.w[i];
        A.w[i] = div_u64_magic(x, b, magic, rem);
    }
    i32 e = -shift;
    const bool inexact = rem != 0;
    if (!inexact) {                                // exact: towards the ideal exponent 0
        e += w_strip_zeros(A, shift);
    }
    return dec_round<5>(A, e, 0, prec, inexact);
}
PB_HD int contig_of_orf(const Batch& B, i64 oi) { return B.o_contig[oi]; }
// The literal chain (S2..S5, hold, finish) works on SLOTS: slot s stands for ORF lit_ids[s] (or ORF s
// when lit_all); its scratch arrays (o_lnx, o_A, o_lnA, o_fac, o_hf, o_bin, o_hold) are indexed by slot.
PB_HD i64 orf_of_slot(const Batch& B, i64 s) { return B.lit_all ? s : (i64)B.lit_ids[s]; }
// Stage 7a (split into small kernels: each keeps its instruction working set inside the I-cache).
// letters of the ORF's own strand-oriented sequence (orfs.py:162-168): popcounts over the base masks
PB_HD void orf_base_counts(const Batch& B, i64 oi, u32& na, u32& nt, u32& ng, u32& len) {
    const int c = contig_of_orf(B, oi);
    const int L = B.cs[c].L;
    const int start = B.o_start[oi], stop = B.o_stop[oi];
    const bool rev = B.o_frame[oi] < 0;
    int x0 = rev ? stop - 1 : start - 1, x1 = rev ? start + 2 : stop + 2;   // extent of orf.seq (functions.py:206,219,234,246)
    if (x0 < 0) x0 = 0;
    if (x1 > L) x1 = L;
    const i64 cb = B.coff[c];
    u32 n4[4];
    count_bases4(B, cb + x0, cb + x1, n4);
    if (!rev) na = n4[0], nt = n4[3], ng = n4[2];
    else na = n4[3], nt = n4[0], ng = n4[1];
    len = (u32)(x1 - x0);
}
// Orf.p_stop (orfs.py:162-173) in Decimal arithmetic, every operation rounded to 28 digits
PB_HDNI Dec orf_pstop_dec(const Batch& B, i64 oi) {
    u32 na, nt, ng, len;
    orf_base_counts(B, oi, na, nt, ng, len);
    const u64 magic = ~0ull / len;
    const Dec Pa = dec_div_u32(na, len, magic, PB_PREC), Pt = dec_div_u32(nt, len, magic, PB_PREC),
              Pg = dec_div_u32(ng, len, magic, PB_PREC);
    // Pt*Pa*Pa + Pt*Pg*Pa + Pt*Pa*Pg, left to right (orfs.py:173); Pa and Pg are the repeated multipliers
    HoldFac fa, fg;
    holdfac_prepare(Pa, fa);
    holdfac_prepare(Pg, fg);
    const Dec m1 = dec_mul28(Pt, Pa, fa), m2 = dec_mul28(Pt, Pg, fg);
    const Dec t1 = dec_mul28(m1, Pa, fa), t2 = dec_mul28(m2, Pa, fa), t3 = dec_mul28(m1, Pg, fg);
    return dec_add(dec_add(t1, t2), t3);
}
// S1: base composition -> pstop, x = 1 - pstop.  item = slot
PB_HDN void st_orf_pstop(const Batch& B, i64 sl) {
    if (sl >= B.nlit) return;
    const i64 oi = orf_of_slot(B, sl);
    const Dec pstop = orf_pstop_dec(B, oi);
    B.o_pstop[oi] = pstop;
    B.o_x[oi] = dec_sub(dec_one(), pstop);
}
// S2: ln(1 - pstop).  item = ORF
PB_HDN void st_orf_lnx(const Batch& B, i64 sl) {
    if (sl >= B.nlit) return;
    const i64 oi = orf_of_slot(B, sl);
    const Dec x = B.o_x[oi];
    SFx lnx;
    w_zero(lnx.m);
    lnx.neg = 0;
    if (!dec_is_one_abs(x)) {
        bool o1, o2;
        Fx X = fx_from_dec(x, &o1);
        lnx = fx_ln(X, &o2);
        if (!(o1 && o2)) PB_ATOMIC_OR(&B.cs[contig_of_orf(B, oi)].err, (u32)ERR_RANGE);
    }
    B.o_lnx[sl] = lnx;
}
// S3: A_im = x ** pos_max[im] and ln(A_im).  item = (im-1)*no + ORF: a warp works on one exponent index, so
// the "exponent is exactly 1 -> plain copy" case (one of the three per contig) does not split warps
PB_HDN void st_orf_powA(const Batch& B, i64 item_in) {
    if (item_in >= (i64)B.nlit * 3) return;
    const i64 sl = item_in % B.nlit;
    const i64 oi = orf_of_slot(B, sl);
    const int im = (int)(item_in / B.nlit) + 1;
    const i64 item = sl * 3 + (im - 1);
    const int c = contig_of_orf(B, oi);
    const CStat* cs = B.cs + c;
    const Dec x = B.o_x[oi];
    const bool xone = dec_is_one_abs(x);
    Dec A;
    SFx lnA = B.o_lnx[sl];
    bool ok = true;
    if (cs->max_one[im]) {
        A = x;                                                        // x ** Decimal(1) == x
    } else if (xone) {
        A = dec_pow_fx(x, cs->fmax[im], 0, PB_PREC, &ok);             // 1.000...0 (_qcheck_pow_one)
    } else {
        SFx T;
        Fx V;
        bool o2;
        A = dec_pow_ln(lnA, cs->fmax[im], PB_PREC, &ok, &T, &V);
        lnA = fx_ln_of_rounded(A, V, T, &o2);
        ok = ok && o2;
    }
    if (!ok) PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_RANGE);
    B.o_A[item] = A;
    B.o_lnA[item] = lnA;
}
// S4: F_k = A_im ** pos_min[il].  item = k*no + ORF (same reason)
PB_HDN void st_orf_powF(const Batch& B, i64 item_in) {
    if (item_in >= (i64)B.nlit * 6) return;
    const i64 sl = item_in % B.nlit;
    const i64 oi = orf_of_slot(B, sl);
    const int k = (int)(item_in / B.nlit);
    const i64 item = sl * 6 + k;
    const int im = k / 2 + 1;
    const int il = (k % 2) + 1 + (((k % 2) + 1 >= im) ? 1 : 0);      // inverse of fac_index
    const int c = contig_of_orf(B, oi);
    const CStat* cs = B.cs + c;
    const Dec A = B.o_A[sl * 3 + (im - 1)];
    Dec f;
    bool ok = true;
    if (cs->min_one[il]) f = A;
    else if (dec_is_one_abs(A)) f = dec_pow_fx(A, cs->fmin[il], 0, PB_PREC, &ok);
    else f = dec_pow_ln(B.o_lnA[sl * 3 + (im - 1)], cs->fmin[il], PB_PREC, &ok);
    if (!ok) PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_RANGE);
    B.o_fac[item] = f;
}
// S5: prepared factors, length bin, histogram for the counting sort.  item = ORF
PB_HDN void st_orf_prepare(const Batch& B, i64 item) {
    const i64 sl = item / 6;
    if (sl >= B.nlit) return;
    const i64 oi = orf_of_slot(B, sl);
    const int k = (int)(item % 6);
    HoldFac hf;
    holdfac_prepare(B.o_fac[item], hf);
    U4* dst = (U4*)(B.o_hf + item);            // three 16-byte stores
    U4 v0, v1, v2;
    v0.x = hf.c27[0]; v0.y = hf.c27[1]; v0.z = hf.c27[2]; v0.w = hf.c27[3];
    v1.x = hf.c28[0]; v1.y = hf.c28[1]; v1.z = hf.c28[2]; v1.w = hf.c28[3];
    v2.x = hf.btop[0]; v2.y = hf.btop[1]; v2.z = (u32)hf.e; v2.w = hf.ok;
    dst[0] = v0;
    dst[1] = v1;
    dst[2] = v2;
    if (k == 0) {
        const bool rev = B.o_frame[oi] < 0;
        int n = orf_steps(B.o_start[oi], B.o_stop[oi], rev);
        int bin = n < HOLD_BINS - 1 ? n : HOLD_BINS - 1;
        B.o_bin[sl] = (unsigned short)bin;
        PB_ATOMIC_ADD(&B.len_hist[HOLD_BINS - 1 - bin], 1u);          // reversed: longest ORFs first
    }
}
PB_HDN void st_len_scatter(const Batch& B, i64 sl) {
    if (sl >= B.nlit) return;
    int bin = B.o_bin[sl];
    u32 pos = PB_ATOMIC_ADD_RET(&B.len_cursor[HOLD_BINS - 1 - bin], 1u);
    B.o_order[B.len_hist[HOLD_BINS - 1 - bin] + pos] = (i32)sl;
}

// Stage 7b+c: the per-codon product of the ORF in slot sl.  S holds its six HoldFac as
// 18 U4 words with stride BD (shared memory on the GPU: S[(k*3+v)*BD + t]).
// GC-frame factor class (0..5) of the codon that starts at batch position g: which of the strand's five class masks has
// the bit; none = class 5 (the scan leaves no per-base class byte: only these few literal ORFs ever ask).  The five mask
// words of the current 64 positions stay in registers: a walk along an ORF reloads them every 21 codons.
struct ClassWin {
    i64 w;
    u64 m[5];
};
PB_HD int factor_class_at(ClassWin& cw, u64* const* M, i64 g) {
    const i64 w = g >> 6;
    if (w != cw.w) {
        cw.w = w;
#pragma unroll
        for (int q = 0; q < 5; q++) cw.m[q] = M[q][w];
    }
    const int b = (int)(g & 63);
    int k = 5;
#pragma unroll
    for (int q = 4; q >= 0; q--)
        if ((cw.m[q] >> b) & 1ull) k = q;
    return k;
}
PB_HDN void hold_run(const Batch& B, i32 sl, const U4* S, int BD, int t) {
    const i64 oi = orf_of_slot(B, sl);
    const int c = contig_of_orf(B, oi);
    u64* const* M = B.o_frame[oi] < 0 ? B.cR : B.cF;
    const i64 gb = B.coff[c] - 1;                  // position b of the contig = batch position gb + b
    ClassWin cw;
    cw.w = -1;
    const int start = B.o_start[oi], stop = B.o_stop[oi];
    const bool rev = B.o_frame[oi] < 0;
    const int step = rev ? -3 : 3;
    const int n = orf_steps(start, stop, rev);
    // the fast path needs all six factors to carry exactly 28 digits
    bool fastok = true;
#pragma unroll
    for (int k = 0; k < 6; k++) fastok = fastok && (S[(k * 3 + 2) * BD + t].w != 0);
    Dec hold = dec_one();
    int it = 0, b = start;
    // generic steps until the product carries 28 digits (normally just the first: 1 * f == f exactly)
    const Wide<4> lo27 = w_pow10<4>(27);
    while (it < n && !(fastok && hold.c.w[3] == 0 && w_cmp(hold.c, lo27) >= 0)) {
        const int k = factor_class_at(cw, M, gb + b);
        hold = dec_mul(hold, B.o_fac[(i64)sl * 6 + k]);             // functions.py:293,298
        it++;
        b += step;
    }
    if (it < n) {
        u32 a0 = hold.c.w[0], a1 = hold.c.w[1], a2 = hold.c.w[2];
        i32 eh = hold.e;
        // software pipeline: class of step it+1 and operands of step it are fetched one step ahead
        int k = factor_class_at(cw, M, gb + b);
        U4 c27 = S[(k * 3 + 0) * BD + t], c28 = S[(k * 3 + 1) * BD + t], misc = S[(k * 3 + 2) * BD + t];
        b += step;
        int knext = (it + 1 < n) ? factor_class_at(cw, M, gb + b) : 0;
        for (; it < n; it++) {
            const int kcur = k;
            const U4 d27 = c27, d28 = c28, dmisc = misc;
            k = knext;
            c27 = S[(k * 3 + 0) * BD + t];
            c28 = S[(k * 3 + 1) * BD + t];
            misc = S[(k * 3 + 2) * BD + t];
            b += step;
            if (it + 2 < n) knext = factor_class_at(cw, M, gb + b);
            if (hold_step_fast(a0, a1, a2, eh, d27, d28, dmisc)) continue;
            // undecidable from 32 fraction bits: exact multiplication, then back to the fast path
            Dec h;
            h.c.w[0] = a0;
            h.c.w[1] = a1;
            h.c.w[2] = a2;
            h.c.w[3] = 0;
            h.e = eh;
            h.neg = 0;
            h = dec_mul(h, B.o_fac[(i64)sl * 6 + kcur]);
            a0 = h.c.w[0];
            a1 = h.c.w[1];
            a2 = h.c.w[2];
            eh = h.e;
        }
        hold.c.w[0] = a0;
        hold.c.w[1] = a1;
        hold.c.w[2] = a2;
        hold.c.w[3] = 0;
        hold.e = eh;
        hold.neg = 0;
    }
    B.o_hold[sl] = hold;
}
// Stage 7c: Orf.score (orfs.py:122-127): weight = -(1/hold * start-codon weight * Decimal(str(weight_rbs))).  item = ORF
PB_HDN void st_orf_finish(const Batch& B, i64 sl) {
    if (sl >= B.nlit) return;
    const i64 oi = orf_of_slot(B, sl);
    const int c = contig_of_orf(B, oi);
    CStat* cs = B.cs + c;
    Dec sc = dec_div(dec_one(), B.o_hold[sl]);
    int sw = B.o_sw[oi];
    if (sw >= 0) sc = dec_mul(sc, B.P.startw[sw]);
    sc = dec_mul(sc, cs->wrbs[B.o_rbs[oi]]);
    sc.neg ^= 1;
    B.o_weight[oi] = sc;
    WInt wi;
    if (!dec_to_wint(sc, wi)) {
        // beyond 256 bits: the contig is solved with 2048-bit distances, this weight formed from the Decimal on demand
        HInt h;
        if (dec_to_hint(sc, h)) {
            wi = wint_huge_marker();
            cs->huge = 1;
            PB_ATOMIC_ADD(B.lit_cnt + 5, 1u);
        } else {
            PB_ATOMIC_OR(&cs->err, (u32)ERR_OVERFLOW);
        }
    }
    B.o_wint[oi] = wi;
    if (!wint_is_narrow(wi)) cs->wide = 1;         // (benign race: every writer stores 1)
    B.o_lit[oi] = 1;
}
