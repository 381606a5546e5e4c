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
// so |W_ref / W_true - 1| <= (2.51 n + 1.51) * 1e-27 for n codon steps (tests/test_filter_bound.py
// measures the actual ratio on the golden ORF tables: <= 4 % of the bound).  We use (3n + 8) * 2^-89
// (2^-89 = 1.6e-27), which also swallows the ~2^-170 relative error of the Q32.192 evaluation.
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
    const Dec x = B.o_x[oi];
    Wide<10> big;
    w_zero(big);
    if (fast && (dec_is_one_abs(x) || dec_is_zero(x) || x.neg)) fast = false;
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
        bool o1, o2, o3;
        const Fx X = fx_from_dec(x, &o1);
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
        const u64 en = 3ull * (u64)n + 8ull;
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
