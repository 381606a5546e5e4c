// Graph stages (functions.py:307-454) and the exact shortest path (phanotate.py:53-65 + the
// fastpathz contract, CHANGELOG.md:11-13,54,57).
//
// Nodes are sorted by (contig, position).  entry = forward start / reverse stop-key, exit =
// forward stop-key / reverse start; every connector goes exit -> entry, every ORF edge entry ->
// exit.  Gap edges (exit l -> entry r, 0 < r-l < 500) are never materialised for the
// solve: their weight is a table lookup on r-l.  Overlap edges (exit r -> entry l, backwards) need
// a 33-digit integer power each and are stored in CSR by source node.  Bridges over >500-bp
// uncovered runs are rare and stored per contig.
#pragma once
#include "hold.cuh"
#include "dec2double.cuh"

PB_HD bool kind_is_entry(int k) { return k == K_FSTART || k == K_RSTOP; }
// first entry of contig c's region of call_tmp: its ORFs and its tRNAs can all be on the path
PB_HD i32 call_base(const Batch& B, int c) { return B.corf[c] + (B.nt > 0 ? B.ctrna[c] : 0); }
PB_HD int contig_of_node(const Batch& B, i32 ni) { return B.n_contig[ni]; }

// Stage 8: other_end[] and the pstop used for overlap averaging, per node, with the reference's
// bare-position dict semantics at the six positions where two roles share a key (orfs.py:17-32,
// functions.py:362-385).  item = node
PB_HDN void st_node_attrs(const Batch& B, i64 ni64) {
    if (ni64 >= B.nn) return;
    const i32 ni = (i32)ni64;
    const int c = contig_of_node(B, ni);
    const int kind = B.n_kind[ni] & 3;
    const int p = B.n_pos[ni];
    i32 oth, oidx;
    if (kind == K_FSTART || kind == K_RSTART) {
        oth = B.o_stop[B.n_orf[ni]];
        oidx = -1;
    } else {
        oth = B.n_pos[B.n_mate[ni]];
        oidx = B.n_orf[ni];
    }
    i32 twin = -1;
    if (ni > B.cnode[c] && B.n_pos[ni - 1] == p) twin = ni - 1;
    else if (ni + 1 < B.cnode[c + 1] && B.n_pos[ni + 1] == p) twin = ni + 1;
    if (twin >= 0) {
        const int tk = B.n_kind[twin] & 3;
        i32 F = -1, R = -1, C = -1, V = -1;
        if (kind == K_FSTART) F = ni; else if (kind == K_RSTOP) R = ni; else if (kind == K_FSTOP) C = ni; else V = ni;
        if (tk == K_FSTART) F = twin; else if (tk == K_RSTOP) R = twin; else if (tk == K_FSTOP) C = twin; else V = twin;
        if (F >= 0 && R >= 0) {
            // left end: forward ORF starting at p and reverse family keyed at p share other_end[p]
            i32 trigF = B.n_trig[B.n_mate[F]], trigR = B.n_trig[R];
            if (trigF < trigR) {         // family R is inserted later and overwrites
                oth = B.n_pos[B.n_mate[R]];
                oidx = B.n_orf[R];
            } else {                     // the forward ORF is inserted later: other_end[p] = its stop
                oth = B.o_stop[B.n_orf[F]];
                oidx = B.n_orf[F];       // p is a stop key but other_end[p] is not in its family -> get_orf(p, other_end[p])
            }
        } else if (C >= 0 && V >= 0) {
            // right end: forward family keyed at p is inserted first, the reverse ORF starting at p last
            oth = B.o_stop[B.n_orf[V]];
            if (B.n_pos[B.n_mate[C]] == oth) oidx = B.n_orf[C];
            else oidx = B.n_orf[V];
        } else {
            PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_INTERNAL);
        }
    }
    B.n_oth[ni] = oth;
    B.n_oidx[ni] = oidx;
    B.n_pk[ni] = ((u32)p << 4) | (u32)(B.n_kind[ni] & 15);
}

// overlap predicate for (left entry e at l, right exit x at r), functions.py:400-438
PB_HD int overlap_kind(const Batch& B, i32 e, i32 x) {     // 0 none, 1 'same', 2 'diff'
    const int ke = B.n_kind[e] & 3, kx = B.n_kind[x] & 3;
    const int fe = B.n_kind[e] >> 2, fx = B.n_kind[x] >> 2;
    const int l = B.n_pos[e], r = B.n_pos[x];
    const int lo = B.n_oth[e], ro = B.n_oth[x];
    if (ke == K_FSTART && kx == K_FSTOP) return (fe != fx && r < lo && ro < l) ? 1 : 0;
    if (ke == K_RSTOP && kx == K_RSTART) return (fe != fx && r < lo && ro < l) ? 1 : 0;
    if (ke == K_RSTOP && kx == K_FSTOP) return (ro + 3 < l && r < lo) ? 2 : 0;
    if (ke == K_FSTART && kx == K_RSTART) return (ro < l && r < lo) ? 2 : 0;
    return 0;
}
// the pstop a node contributes to ave([o1,o2]) (functions.py:373-385); an ORF whose Decimal pstop has not been
// materialised yet (certified run) gets it computed on the spot
PB_HD Dec node_o(const Batch& B, int c, i32 n) {
    const i32 oi = B.n_oidx[n];
    if (oi < 0) return B.cs[c].pstop;
    if (B.o_lit[oi] == 1) return B.o_pstop[oi];
    return orf_pstop_dec(B, oi);
}
// score_overlap(r-l+3, dir, ave([o1,o2]))  (functions.py:26-34,140-141,386)
PB_HDN Dec overlap_score(const Batch& B, int c, i32 e, i32 x, bool diff) {
    Dec t = dec_add(dec_from_u64(0), node_o(B, c, e));
    t = dec_add(t, node_o(B, c, x));
    Dec pbar = dec_div(t, dec_from_u64(2));
    Dec o = dec_sub(dec_one(), pbar);
    Dec sc = dec_powi(o, (u32)(B.n_pos[x] - B.n_pos[e] + 3));
    sc = dec_div(dec_one(), sc);
    if (diff) sc = dec_add(sc, dec_twenty());
    return sc;
}
// Stage 9/10: overlap edges out of exit node ni.  fill=false counts, fill=true writes.
// The counting pass evaluates the predicate over the entries within 500 bp upstream and leaves WHICH of them matched as a
// bit mask (bit k = node ni-1-k; the window is ~18 nodes, 64 covers all but pathologically dense stretches), so the
// filling pass only expands the mask: no second round of other_end loads and predicate tests.
PB_HDN void overlaps_of(const Batch& B, i32 ni, bool fill) {
    const u32 wx = B.n_pk[ni];                      // position << 4 | kind | frame << 2
    const int kx = (int)(wx & 3), fx = (int)((wx >> 2) & 3), r = (int)(wx >> 4);
    u32 cnt = 0;
    if (!kind_is_entry(kx)) {
        u32 k = fill ? B.ov_cnt[ni] : 0;
        if (fill) {
            const u32 have = B.ov_cnt[ni + 1] - k;
            u64 m = B.ov_mask[ni];
            if (have == 0) return;
            if (m) {                                    // the counting pass's matches, nearest entry first (as it found them)
                while (m) {
                    const int b = pb_ctz64(m);
                    m &= m - 1;
                    const i32 j = ni - 1 - b;
                    const int ke = (int)(B.n_pk[j] & 3);
                    B.ov_dst[k] = j;
                    B.ov_src[k] = ni;
                    // 'diff' = the two ORFs lie on different strands (functions.py:419-438): stop-key/stop-key or start/start
                    B.ov_diff[k] = (u8)((ke == K_RSTOP && kx == K_FSTOP) || (ke == K_FSTART && kx == K_RSTART));
                    k++;
                }
                return;
            }
        }
        const int c = contig_of_node(B, ni);
        const i32 first = B.cnode[c];
        const int ro = B.n_oth[ni];
        u64 mask = 0;
        bool fits = true;
        for (i32 j = ni - 1; j >= first; j--) {
            const u32 we = B.n_pk[j];
            const int l = (int)(we >> 4), ke = (int)(we & 3), fe = (int)((we >> 2) & 3);
            if (r - l >= 500) break;
            if (l >= r || !kind_is_entry(ke)) continue;
            // the overlap predicate of overlap_kind() on the packed node words (functions.py:400-438)
            const int lo = B.n_oth[j];
            int ok;
            if (ke == K_FSTART && kx == K_FSTOP) ok = (fe != fx && r < lo && ro < l) ? 1 : 0;
            else if (ke == K_RSTOP && kx == K_RSTART) ok = (fe != fx && r < lo && ro < l) ? 1 : 0;
            else if (ke == K_RSTOP && kx == K_FSTOP) ok = (ro + 3 < l && r < lo) ? 2 : 0;
            else ok = (ro < l && r < lo) ? 2 : 0;          // K_FSTART entry, K_RSTART exit
            if (!ok) continue;
            if (fill) {
                B.ov_dst[k] = j;
                B.ov_src[k] = ni;
                B.ov_diff[k] = (u8)(ok == 2);
                k++;
            } else {
                const i32 d = ni - 1 - j;
                if (d < 64) mask |= 1ull << d;
                else fits = false;
            }
            cnt++;
        }
        if (!fill) B.ov_mask[ni] = fits ? mask : 0ull;     // (0 with a non-zero count: the filling pass tests again)
    }
    if (!fill) B.ov_cnt[ni] = cnt;
}
PB_HDN void st_ov_count(const Batch& B, i64 ni) {
    if (ni < B.nn) overlaps_of(B, (i32)ni, false);
}
PB_HDN void st_ov_fill(const Batch& B, i64 ni) {
    if (ni < B.nn) overlaps_of(B, (i32)ni, true);
}
// Stage 10b-d: weight of one overlap edge in three small kernels (instruction-cache footprint):
//   pbar:   o = 1 - ave([o1,o2])                     (functions.py:140-141,386; 26-27)
//   pow:    o ** Decimal(r-l+3), a 33-digit square-and-multiply (functions.py:30)
//   weight: 1/score (+ 1/0.05 if 'diff'), integer weight (functions.py:31-34)
// item = overlap edge
// The chain works on SLOTS: slot sl stands for edge ovlit_ids[sl] (or edge sl when ov_all); ov_w[] is per slot.
PB_HD i64 edge_of_slot(const Batch& B, i64 sl) { return B.ov_all ? sl : (i64)B.ovlit_ids[sl]; }
PB_HDN void st_ov_pbar(const Batch& B, i64 sl) {
    if (sl >= B.novlit) return;
    const i64 k = edge_of_slot(B, sl);
    const i32 x = B.ov_src[k], e = B.ov_dst[k];
    const int c = contig_of_node(B, x);
    Dec t = dec_add(dec_from_u64(0), node_o(B, c, e));
    t = dec_add(t, node_o(B, c, x));
    Dec pbar = dec_div(t, dec_from_u64(2));
    B.ov_w[sl] = dec_sub(dec_one(), pbar);
}
PB_HDN void st_ov_pow(const Batch& B, i64 sl) {
    if (sl >= B.novlit) return;
    const i64 k = edge_of_slot(B, sl);
    const i32 x = B.ov_src[k], e = B.ov_dst[k];
    B.ov_w[sl] = dec_powi(B.ov_w[sl], (u32)(B.n_pos[x] - B.n_pos[e] + 3));
}
PB_HDN void st_ov_weight(const Batch& B, i64 sl) {
    if (sl >= B.novlit) return;
    const i64 k = edge_of_slot(B, sl);
    const int c = contig_of_node(B, B.ov_src[k]);
    Dec sc = dec_div(dec_one(), B.ov_w[sl]);
    if (B.ov_diff[k]) sc = dec_add(sc, dec_twenty());
    B.ov_w[sl] = sc;
    WInt wi;
    if (!dec_to_wint(sc, wi)) PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_OVERFLOW);
    B.ov_wint[k] = wi;
    if (!wint_is_narrow(wi)) B.cs[c].wide = 1;
    bool small = (wi.w[1] >> 30) == 0;               // 0 <= weight < 2^62 (overlap weights are positive)
#pragma unroll
    for (int i = 2; i < WN; i++) small = small && wi.w[i] == 0;
    B.ov_w64[k] = small ? (i64)(((u64)wi.w[1] << 32) | wi.w[0]) : OV_W64_WIDE;
}

// Stage 11: bridges over uncovered runs longer than 500 bp (functions.py:320-354).
// Coverage = union over families of [min(start,stop), min(max(start,stop), L-1)) of the longest ORF.
// Interval starts are entry nodes (a reverse stop-key node, or the farthest start of a forward
// family), already sorted by position, so "last covered base before node i" is an exclusive prefix
// maximum over the node list: one warp per contig computes it (reach_contig), then every interval
// start that begins more than 500 bp after it enumerates its bridging pairs (item = node).
PB_HD bool bridge_interval(const Batch& B, i32 i, int L, int& mi, int& me) {
    const int kind = B.n_kind[i] & 3;
    if (kind == K_RSTOP || (kind == K_FSTART && B.n_mate[B.n_mate[i]] == i)) {
        mi = B.n_pos[i];
        int ma = B.n_pos[B.n_mate[i]];
        me = ma < L - 1 ? ma : L - 1;           // covered: mi .. me-1
        return me > mi;
    }
    return false;
}
PB_HDNI WInt bridge_wint_long(const Batch& B, int c, int len) {
    WInt wi;
    if (!dec_to_wint(gap_score(B, c, len, false), wi)) PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_OVERFLOW);
    if (!wint_is_narrow(wi)) B.cs[c].wide = 1;
    return wi;
}
// The bridging pairs of interval start i (position `base`, last covered base `last`, base - last > 500): counted, or
// written from slot k on.  left: exits with last-500 < l <= last+1 ; right: entries with base-1 <= r < base+500
PB_HDN u32 bridge_pairs(const Batch& B, i32 i, int c, int base, int last, bool fill, u32 k) {
    const i32 nb = B.cnode[c], ne = B.cnode[c + 1];
    u32 cnt = 0;
    i32 r0 = i;
    while (r0 > nb && B.n_pos[r0 - 1] >= base - 1) r0--;          // same-position twins sort before i
    for (i32 r = r0; r < ne && B.n_pos[r] < base + 500; r++) {
        if (B.n_pos[r] < base - 1 || !kind_is_entry(B.n_kind[r] & 3)) continue;
        for (i32 l = i - 1; l >= nb && B.n_pos[l] > last - 500; l--) {
            if (B.n_pos[l] > last + 1 || kind_is_entry(B.n_kind[l] & 3)) continue;
            int len = B.n_pos[r] - B.n_pos[l] - 3;
            if (B.n_pos[r] - B.n_pos[l] < 500) PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_PARALLEL);
            if (fill) {
                B.n_brs[l] |= 1;
                B.br_src[k] = l;
                B.br_dst[k] = r;
                // score_gap(len > 300) = g**100 + len (functions.py:40-41): its integer is len*1000 + a per-contig constant
                // for 3- and 4-digit lengths; longer gaps take the Decimal route
                B.br_wint[k] = (len <= 9999) ? wint_from_i64((i64)len * 1000 + (len <= 999 ? B.cs[c].gap_hi3 : B.cs[c].gap_hi4))
                                             : bridge_wint_long(B, c, len);
                k++;
            }
            cnt++;
        }
    }
    return cnt;
}
// exclusive prefix maximum of the interval ends over nodes [nb, ne) starting from `run`; returns the total.
// write = false: only the total.  write = true: n_reach, and with it the number of bridging pairs of every node (br_cnt:
// non-zero only for the rare interval start that begins more than 500 bp after the last covered base)
PB_HDN int reach_range(const Batch& B, int c, i32 nb, i32 ne, int run, bool write, int lane, int NL) {
    const int L = (int)(B.coff[c + 1] - B.coff[c]);
    for (i32 base = nb; base < ne; base += NL) {
        const i32 i = base + lane;
        int v = 0, mi = 0, me;
        const bool iv = i < ne && bridge_interval(B, i, L, mi, me);
        if (iv) v = me - 1;
        int incl = v;
#ifdef __CUDA_ARCH__
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o && t > incl) incl = t;
        }
        int prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        if (lane == 0) prev = 0;
        int excl = prev > run ? prev : run;
        int tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
#else
        int excl = run, tot = incl;
#endif
        if (write && i < ne) {
            B.n_reach[i] = excl;
            B.br_cnt[i] = (iv && mi > excl && mi - excl > 500) ? bridge_pairs(B, i, c, mi, excl, false, 0) : 0u;
        }
        if (tot > run) run = tot;
    }
    return run;
}
PB_HDN void reach_contig(const Batch& B, int c, int lane, int NL) {
    if (B.ch_cnt[c + 1] > B.ch_cnt[c]) return;    // a long contig: reach_chunk_* (chunk.cuh)
    reach_range(B, c, B.cnode[c], B.cnode[c + 1], 0, true, lane, NL);
}
// the pairs written (after the exclusive scan of br_cnt).  item = node
PB_HDN void st_br_fill(const Batch& B, i64 i) {
    if (i >= B.nn) return;
    const u32 k = B.br_cnt[i];
    if (B.br_cnt[i + 1] == k) return;             // (all but a few thousand nodes of a batch)
    const int c = contig_of_node(B, (i32)i);
    bridge_pairs(B, (i32)i, c, B.n_pos[i], B.n_reach[i], true, k);
}

// ------------------------------------------------------------------------------------------------
// Stage 12: exact single-source shortest path source -> target.  One warp per contig: nodes are
// visited in position order, each visit pushes the node's out-edges (lanes over edges); an
// improvement through a backward (overlap) edge rewinds the sweep to the improved node, so the
// result is the Bellman-Ford fixpoint with strict '<' relaxation.
#ifdef __CUDA_ARCH__
#define PB_SYNCWARP() __syncwarp()
#else
#define PB_SYNCWARP()
#endif

// The distances are exact integers.  Two widths of the same algorithm: 128 bit (all weights of the contig below
// 2^110, i.e. practically always: one 16-byte load/store and two 64-bit add/compare steps per relaxation) and 256
// bit (CStat.wide: some ORF weight is astronomically large).
struct
#ifdef __CUDACC__
    __align__(16)
#endif
    I128 {
    u64 lo;
    i64 hi;
};
// An exact tie: remember the edge.  The sweep keeps the first edge IT saw reach a distance; the reference's solver
// (Bellman-Ford over the edges in Graph.iteredges() order, strict '<', phanotate.py:56-64) keeps the first edge IT sees,
// and st_tie_fix settles the difference afterwards from these records.
// Recording is a handful of inline instructions in the cold branch of relax (one atomic counter, three stores): the
// events are threaded into per-contig lists afterwards (st_tie_link).  v = -3 - contig stands for the contig's target.
PB_HD TieEv* tie_slot(const Batch& B, int c, i32 v, i32 from) {
    const u32 k = PB_ATOMIC_ADD_RET(B.tie_n, 1u);
    if (k >= (u32)B.tie_cap) {
        PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_TIES);
        return nullptr;
    }
    TieEv* e = B.tie_ev + k;
    e->v = v;
    e->from = from;
    return e;
}
struct D256 {
    typedef WInt T;
    static PB_HD T inf() { return wint_inf(); }
    static PB_HD bool is_inf(const T& a) { return wint_is_inf(a); }
    static PB_HD bool less(const T& a, const T& b) { return wint_less(a, b); }
    static PB_HD bool eq(const T& a, const T& b) { return w_cmp(a, b) == 0; }
    static PB_HD T add(T a, const T& b) {
        w_add(a, b);
        return a;
    }
    static PB_HD T from_i64(i64 v) { return wint_from_i64(v); }
    static PB_HD T load_w(const WInt* p) { return *p; }
    static PB_HD T orf_w(const Batch& B, i32 orf) { return B.o_wint[orf]; }
    static PB_HD T* dist(const Batch& B) { return B.dist; }
    static PB_HD WInt to_wint(const T& a) { return a; }
    static PB_HD void tie(const Batch& B, int c, i32 v, i32 from, const T& cand) {
        TieEv* e = tie_slot(B, c, v, from);
        if (e) {
            e->pad = 0;
            e->cand = cand;
        }
    }
};
struct D128 {
    typedef I128 T;
    static PB_HD T inf() {
        T r;
        r.lo = ~0ull;
        r.hi = 0x7FFFFFFFFFFFFFFFll;
        return r;
    }
    static PB_HD bool is_inf(const T& a) { return a.hi == 0x7FFFFFFFFFFFFFFFll; }
    static PB_HD bool less(const T& a, const T& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
    static PB_HD bool eq(const T& a, const T& b) { return a.hi == b.hi && a.lo == b.lo; }
    static PB_HD T add(T a, const T& b) {
        const u64 lo = a.lo + b.lo;
        a.hi = (i64)((u64)a.hi + (u64)b.hi + (lo < a.lo ? 1ull : 0ull));
        a.lo = lo;
        return a;
    }
    static PB_HD T from_i64(i64 v) {
        T r;
        r.lo = (u64)v;
        r.hi = v < 0 ? -1ll : 0ll;
        return r;
    }
    static PB_HD T load_w(const WInt* p) {      // low 128 bits of the two's complement value
        const U4 q = *(const U4*)p;
        T r;
        r.lo = ((u64)q.y << 32) | q.x;
        r.hi = (i64)(((u64)q.w << 32) | q.z);
        return r;
    }
    static PB_HD T orf_w(const Batch& B, i32 orf) { return load_w(B.o_wint + orf); }
    static PB_HD T* dist(const Batch& B) { return B.dist128; }
    static PB_HD WInt to_wint(const T& a) {
        WInt r;
        r.w[0] = (u32)a.lo;
        r.w[1] = (u32)(a.lo >> 32);
        r.w[2] = (u32)(u64)a.hi;
        r.w[3] = (u32)((u64)a.hi >> 32);
        const u32 ext = a.hi < 0 ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int i = 4; i < WN; i++) r.w[i] = ext;
        if (is_inf(a)) r = wint_inf();
        return r;
    }
    static PB_HD void tie(const Batch& B, int c, i32 v, i32 from, const T& cand) {
        TieEv* e = tie_slot(B, c, v, from);
        if (e) {
            e->pad = 1;                           // the low 128 bits; st_tie_fix extends the sign
            u64* w = (u64*)&e->cand;
            w[0] = cand.lo;
            w[1] = (u64)cand.hi;
        }
    }
};
// 2048-bit distances: contigs with an ORF weight beyond 2^240 (CStat.huge; score.cuh).  Same algorithm, values in local
// memory; rare by construction (an ORF of many kb in AT-rich sequence), so nothing here is tuned.
struct DHuge {
    typedef HInt T;
    static PB_HD T inf() {
        T r;
        for (int i = 0; i < WH; i++) r.w[i] = 0xFFFFFFFFu;
        r.w[WH - 1] = 0x7FFFFFFFu;
        return r;
    }
    static PB_HD bool is_inf(const T& a) { return a.w[WH - 1] == 0x7FFFFFFFu; }
    static PB_HD bool less(const T& a, const T& b) {
        const u32 sa = a.w[WH - 1] >> 31, sb = b.w[WH - 1] >> 31;
        if (sa != sb) return sa > sb;
        return w_cmp(a, b) < 0;
    }
    static PB_HD bool eq(const T& a, const T& b) { return w_cmp(a, b) == 0; }
    static PB_HD T add(T a, const T& b) {
        w_add(a, b);
        return a;
    }
    static PB_HD T from_i64(i64 v) {
        T r;
        const u32 ext = v < 0 ? 0xFFFFFFFFu : 0u;
        r.w[0] = (u32)(u64)v;
        r.w[1] = (u32)((u64)v >> 32);
        for (int i = 2; i < WH; i++) r.w[i] = ext;
        return r;
    }
    static PB_HD T load_w(const WInt* p) {         // sign-extended
        T r;
        const u32 ext = (p->w[WN - 1] >> 31) ? 0xFFFFFFFFu : 0u;
        for (int i = 0; i < WN; i++) r.w[i] = p->w[i];
        for (int i = WN; i < WH; i++) r.w[i] = ext;
        return r;
    }
    static PB_HDN T orf_w(const Batch& B, i32 orf) {
        if (!wint_is_huge_marker(B.o_wint[orf])) return load_w(B.o_wint + orf);
        T r;
        dec_to_hint(B.o_weight[orf], r);          // (checked when the marker was set)
        return r;
    }
    static PB_HD T* dist(const Batch& B) { return B.dist_huge; }
    static PB_HD WInt to_wint(const T& a) {       // the low 256 bits (the tie records and the target compare on those)
        WInt r;
        for (int i = 0; i < WN; i++) r.w[i] = a.w[i];
        if (is_inf(a)) r = wint_inf();
        return r;
    }
    static PB_HD void tie(const Batch& B, int c, i32 v, i32 from, const T& cand) {
        TieEv* e = tie_slot(B, c, v, from);
        if (e) {
            e->pad = 0;
            for (int i = 0; i < WN; i++) e->cand.w[i] = cand.w[i];
        }
    }
};
// integer weight of score_gap(len,'same'|'diff') as the solver sees it; len < 500 (connect), <= 2000 (terminals)
PB_HD i64 gap_w64(const Batch& B, int c, int len, bool diff, bool* ok) {
    *ok = true;
    if (len <= 300) {
        const i64 k = (i64)c * GAPN + (len + 2);
        return diff ? B.gapi_diff[k] : B.gapi_same[k];
    }
    if (len <= 999) return (i64)len * 1000 + B.cs[c].gap_hi3;
    if (len <= 9999) return (i64)len * 1000 + B.cs[c].gap_hi4;
    *ok = false;
    return 0;
}

// A sweep over part of a contig (chunk.cuh): nodes [s, e) only, distances and dirty flags in the chunk's private arrays
// (indexed by node id), no parents, no tie records, no target -- the distances are all that is kept.  base = position
// the stand-in source sits at (0: the contig's real source, functions.py:444-447).
struct SolveRange {
    i32 s, e;
    struct I128* dist;
    u8* dirty;
    i32 base;
};
// relaxation of node v whose current distance `cur` the caller already holds
template <class D, bool CH = false>
PB_HD bool relax(const Batch& B, u32& ties, typename D::T* dist, i32 v, const typename D::T& cur, const typename D::T& cand,
                 i32 from, u8* dirty = nullptr) {
    if (CH) {
        if (D::less(cand, cur)) {
            dist[v] = cand;
            dirty[v] = 1;
            return true;
        }
        return false;
    }
    if (D::less(cand, cur)) {
        dist[v] = cand;
        B.parent[v] = from;
        B.dirty[v] = 1;
        return true;
    }
    if (!D::is_inf(cur) && D::eq(cand, cur) && B.parent[v] != from) {
        ties++;
        D::tie(B, B.n_contig[v], v, from, cand);
    }
    return false;
}

// Loads that do not depend on each other are issued together (node word, distance of the candidate target,
// edge ranges), so a visit is two or three memory round trips deep instead of five.
// CH = true: the sweep of one chunk of a long contig over the node range R (see SolveRange).
template <class D, bool CH = false>
PB_HDN void solve_contig_t(const Batch& B, int c, int lane, int NL, const SolveRange* R = nullptr) {
    typedef typename D::T T;
    const i32 nb = CH ? R->s : B.cnode[c], ne = CH ? R->e : B.cnode[c + 1];
    CStat* cs = B.cs + c;
    const int L = cs->L;
    T* dist = CH ? (T*)R->dist : D::dist(B);
    u8* dirty = CH ? R->dirty : B.dirty;
    const int base = CH ? R->base : 0;
    const u32* pk = B.n_pk;                       // position << 4 | kind | frame << 2
    u32 ties = 0;
    bool okw = true;
    // tRNA nodes of the contig (trna.cuh): behind the regular nodes of the batch, edges in the explicit list te_*
    const bool TR = !CH && B.nt > 0 && B.ctrna[c + 1] > B.ctrna[c];
    const i32 tnb = TR ? B.nn + 2 * B.ctrna[c] : 0, tne = TR ? B.nn + 2 * B.ctrna[c + 1] : 0;
    const u32 teb = TR ? B.te_cnt[2 * B.ctrna[c]] : 0, tee = TR ? B.te_cnt[2 * B.ctrna[c + 1]] : 0;
    for (int part = 0; part < (TR ? 2 : 1); part++)
        for (i32 i = (part ? tnb : nb) + lane; i < (part ? tne : ne); i += NL) {
            const u32 w = pk[i];
            T d0 = D::inf();
            i32 p0 = -1;
            u8 f0 = 0;
            // source -> entry nodes within 2000 bp of the left end (functions.py:444-447)
            if ((int)(w >> 4) - base <= 2000 && kind_is_entry((int)(w & 3))) {
                bool o;
                d0 = D::from_i64(gap_w64(B, c, (int)(w >> 4) - base, false, &o));
                okw = okw && o;
                p0 = -2;
                f0 = 1;
            }
            dist[i] = d0;
            if (!CH) B.parent[i] = p0;
            dirty[i] = f0;
        }
    T tdist = D::inf();
    i32 tpar = -1;
    PB_SYNCWARP();
    const u32 brb = B.br_cnt[B.cnode[c]], bre = B.br_cnt[B.cnode[c + 1]];
    i32 i = nb;
    int budget = 64 * (ne - nb) + 1024 + (TR ? 4096 * (tne - tnb) : 0);
    for (;;) {
    while (i < ne) {
        // next dirty node at or after i
#ifdef __CUDA_ARCH__
        {
            bool found = false;
            while (i < ne) {
                i32 j = i + lane;
                unsigned m = __ballot_sync(0xFFFFFFFFu, j < ne && dirty[j]);
                if (m) {
                    i += __ffs(m) - 1;
                    found = true;
                    break;
                }
                i += 32;
            }
            if (!found) break;
        }
#else
        while (i < ne && !dirty[i]) i++;
        if (i >= ne) break;
#endif
        if (--budget < 0) {
            if (lane == 0) PB_ATOMIC_OR(&cs->err, (u32)ERR_INTERNAL);
            break;
        }
        const i32 u = i;
        const u32 wu = pk[u];
        const T Du = dist[u];
        const int kind = (int)(wu & 3), pu = (int)(wu >> 4);
        PB_SYNCWARP();
        if (lane == 0) dirty[u] = 0;
        i32 rewind = 0x7FFFFFFF;
        if (kind == K_FSTART) {
            if (lane == 0) {
                const i32 v = B.n_mate[u], orf = B.n_orf[u];
                if (!CH || v < ne) {
                    const T cur = dist[v];
                    relax<D, CH>(B, ties, dist, v, cur, D::add(Du, D::orf_w(B, orf)), u, dirty);
                }
            }
        } else if (kind == K_RSTOP) {
            const int farpos = (int)(pk[B.n_mate[u]] >> 4);
            for (i32 j = u + 1 + lane; j < ne; j += NL) {
                const u32 wj = pk[j];
                const i32 mj = B.n_mate[j];
                if ((int)(wj >> 4) > farpos) break;
                if ((int)(wj & 3) == K_RSTART && mj == u) {
                    const i32 orf = B.n_orf[j];
                    const T cur = dist[j];
                    relax<D, CH>(B, ties, dist, j, cur, D::add(Du, D::orf_w(B, orf)), u, dirty);
                }
            }
        } else {
            const u32 ovb = B.ov_cnt[u], ove = B.ov_cnt[u + 1];
            // gap edges to entries within 500 bp downstream (functions.py:360-438)
            for (i32 j = u + 1 + lane; j < ne; j += NL) {
                const u32 wj = pk[j];
                const T cur = dist[j];
                const int kj = (int)(wj & 3), d = (int)(wj >> 4) - pu;
                if (d >= 500) break;
                if (d <= 0 || !kind_is_entry(kj)) continue;
                bool diff = (kind == K_FSTOP) ? (kj == K_RSTOP) : (kj == K_FSTART);
                if (kind == K_RSTART && kj == K_FSTART && d <= 2) continue;      // functions.py:431
                bool o;
                relax<D, CH>(B, ties, dist, j, cur, D::add(Du, D::from_i64(gap_w64(B, c, d - 3, diff, &o))), u, dirty);
            }
            // overlap edges (backwards)
            if (ove > ovb) {
                for (u32 k = ovb + lane; k < ove; k += NL) {
                    const i64 w64 = B.ov_w64[k];
                    const i32 v = B.ov_dst[k];
                    if (CH && v < nb) continue;
                    const T cur = dist[v];
                    const T cand = D::add(Du, w64 != OV_W64_WIDE ? D::from_i64(w64) : D::load_w(B.ov_wint + k));
                    if (relax<D, CH>(B, ties, dist, v, cur, cand, u, dirty) && v < rewind) rewind = v;
                }
#ifdef __CUDA_ARCH__
                rewind = (i32)__reduce_min_sync(0xFFFFFFFFu, (unsigned)rewind);
#endif
            }
            // bridges (rare: n_brs flags the exit nodes that have any)
            if (bre > brb && (B.n_brs[u] & 1)) {
                for (u32 k = brb + lane; k < bre; k += NL) {
                    if (B.br_src[k] != u) continue;
                    const i32 v = B.br_dst[k];
                    if (CH && v >= ne) continue;
                    const T cur = dist[v];
                    relax<D, CH>(B, ties, dist, v, cur, D::add(Du, D::load_w(B.br_wint + k)), u, dirty);
                }
            }
            // gap edges into tRNA entry nodes (functions.py:388-399)
            if (TR && (B.n_brs[u] & 2)) {
                for (u32 k = teb + lane; k < tee; k += NL) {
                    if (B.te_src[k] != u) continue;
                    const i32 v = B.te_dst[k];
                    const T cur = dist[v];
                    relax<D, CH>(B, ties, dist, v, cur, D::add(Du, D::from_i64(B.te_w[k])), u, dirty);
                }
            }
            // exit -> target within 2000 bp of the right end (functions.py:448-451)
            if (!CH && L - pu <= 2000) {
                bool o;
                const T cand = D::add(Du, D::from_i64(gap_w64(B, c, L - pu, false, &o)));
                okw = okw && o;
                if (D::less(cand, tdist)) {
                    tdist = cand;
                    tpar = u;
                } else if (lane == 0 && !D::is_inf(tdist) && D::eq(cand, tdist) && tpar != u) {
                    ties++;
                    D::tie(B, c, -3 - c, u, cand);
                }
            }
        }
        PB_SYNCWARP();
        i = (rewind < u) ? rewind : u + 1;
    }
    if (!TR) break;
    // the contig's tRNA nodes: every improved one pushes its explicit edges; an improvement of a regular node takes
    // the sweep back to it, one of another tRNA node repeats this pass
    i32 back = 0x7FFFFFFF;
    bool again = false;
    for (i32 t = tnb; t < tne; t++) {
        if (!dirty[t]) continue;
        if (--budget < 0) {
            if (lane == 0) PB_ATOMIC_OR(&cs->err, (u32)ERR_INTERNAL);
            break;
        }
        const T Du = dist[t];
        const int kind = (int)(pk[t] & 3), pu = (int)(pk[t] >> 4);
        PB_SYNCWARP();
        if (lane == 0) dirty[t] = 0;
        for (u32 k = teb + lane; k < tee; k += NL) {
            if (B.te_src[k] != t) continue;
            const i32 v = B.te_dst[k];
            const T cur = dist[v];
            if (relax<D, CH>(B, ties, dist, v, cur, D::add(Du, D::from_i64(B.te_w[k])), t, dirty)) {
                if (v < ne) back = v < back ? v : back;
                else again = true;
            }
        }
        if (!kind_is_entry(kind) && L - pu <= 2000) {
            bool o;
            const T cand = D::add(Du, D::from_i64(gap_w64(B, c, L - pu, false, &o)));
            okw = okw && o;
            if (D::less(cand, tdist)) {
                tdist = cand;
                tpar = t;
            } else if (lane == 0 && !D::is_inf(tdist) && D::eq(cand, tdist) && tpar != t) {
                ties++;
                D::tie(B, c, -3 - c, t, cand);
            }
        }
        PB_SYNCWARP();
    }
#ifdef __CUDA_ARCH__
    back = (i32)__reduce_min_sync(0xFFFFFFFFu, (unsigned)back);
    again = __any_sync(0xFFFFFFFFu, again);
#endif
    if (budget < 0) break;
    if (back < 0x7FFFFFFF) i = back;
    else if (again) i = ne;
    else break;
    }
    if (CH) return;
    if (ties) PB_ATOMIC_ADD(&cs->n_ties, ties);
    if (!okw && lane == 0) PB_ATOMIC_OR(&cs->err, (u32)ERR_OVERFLOW);
    if (lane == 0) {
        B.tdist[c] = D::to_wint(tdist);
        B.tparent[c] = tpar;
    }
}
#ifdef __CUDACC__
// The 128-bit sweep as the kernel runs it.  Same visits, same relaxations, same order as solve_contig_t (which stays
// the plain statement: the host build and the 256-bit contigs run it); what changes is where the operands come from.
// Every iteration the warp loads the records of the 32 nodes at the sweep position in ONE round of independent loads
// (dirty flag, packed word, distance, mate, ORF, overlap range: one node per lane).  The ballot over the dirty flags
// picks the node u to visit, its record arrives by shuffles, and the lanes behind it already hold the packed word and
// the current distance of u's gap / reverse-ORF targets: a visit is one or two memory round trips deep instead of four.
// Targets beyond the 32-node window (dense stretches, long reverse families) take the loops of the plain statement.
template <int NL>
__device__ __forceinline__ I128 shfl128(unsigned mask, const I128& a, int src) {
    I128 r;
    r.lo = (u64)__shfl_sync(mask, (unsigned long long)a.lo, src, NL);
    r.hi = (i64)__shfl_sync(mask, (long long)a.hi, src, NL);
    return r;
}
// NL = lanes per contig (32: a warp; 16: two contigs share a warp, each half with its own control flow -- twice the
// contigs in flight per SM at the same register cost, which is what a latency-bound sweep wants); lane = 0..NL-1,
// mask = the lanes of this contig inside the warp.
template <int NL, bool CH = false>
__device__ void solve_contig_win(const Batch& B, int c, int lane, unsigned mask, const SolveRange* R = nullptr) {
    const int lane0 = __ffs((int)mask) - 1;       // position of lane 0 of this group inside the warp
    typedef D128 D;
    typedef I128 T;
    const i32 nb = CH ? R->s : B.cnode[c], ne = CH ? R->e : B.cnode[c + 1];
    CStat* cs = B.cs + c;
    const int L = cs->L;
    T* dist = CH ? R->dist : B.dist128;
    u8* dirty = CH ? R->dirty : B.dirty;
    const int base = CH ? R->base : 0;
    const u32* pk = B.n_pk;
    u32 ties = 0;
    bool okw = true;
    for (i32 i = nb + lane; i < ne; i += NL) {
        const u32 w = pk[i];
        T d0 = D::inf();
        i32 p0 = -1;
        u8 f0 = 0;
        if ((int)(w >> 4) - base <= 2000 && kind_is_entry((int)(w & 3))) {       // source -> entry (functions.py:444-447)
            bool o;
            d0 = D::from_i64(gap_w64(B, c, (int)(w >> 4) - base, false, &o));
            okw = okw && o;
            p0 = -2;
            f0 = 1;
        }
        dist[i] = d0;
        if (!CH) B.parent[i] = p0;
        dirty[i] = f0;
    }
    T tdist = D::inf();
    i32 tpar = -1;
    __syncwarp(mask);
    const u32 brb = B.br_cnt[B.cnode[c]], bre = B.br_cnt[B.cnode[c + 1]];
    i32 i = nb;
    int budget = 64 * (ne - nb) + 1024;
    while (i < ne) {
        // ---- the window: node i + lane
        const i32 j0 = i + lane;
        const bool in = j0 < ne;
        u8 dj = 0;
        u32 wj = 0, cj = 0;
        i32 mj = -1, oj = -1;
        T Tj = D::inf();
        if (in) {
            dj = dirty[j0];
            wj = pk[j0];
            Tj = dist[j0];
            mj = B.n_mate[j0];
            oj = B.n_orf[j0];
            cj = B.ov_cnt[j0];
        }
        const unsigned m = (__ballot_sync(mask, dj != 0) >> lane0) & (NL == 32 ? 0xFFFFFFFFu : ((1u << (NL & 31)) - 1u));
        if (!m) {
            i += NL;
            continue;
        }
        if (--budget < 0) {
            if (lane == 0) atomicOr(&cs->err, (u32)ERR_INTERNAL);
            break;
        }
        const int f = __ffs((int)m) - 1;
        if (NL == 32) {
            // Forward-start nodes have ONE out-edge (their ORF edge, to the family's stop): all the dirty ones in front of
            // the first dirty node of another kind are visited in this one iteration, a lane each -- a visit costs a
            // round of dependent loads whatever it does, and these are 38 % of all visits (measured: solve 6.90 -> 6.75 ms;
            // taking the dirty starts BEHIND another dirty node too revisits them after that node's push).  Starts of one family share
            // their stop: those lanes take turns in node order (what the one-by-one sweep does), so every equal offer
            // still meets the distance it ties with.
            const unsigned mnf = __ballot_sync(0xFFFFFFFFu, dj != 0 && (int)(wj & 3) != K_FSTART);
            const unsigned upto = mnf ? ((1u << (__ffs((int)mnf) - 1)) - 1u) : 0xFFFFFFFFu;
            const unsigned mfs = m & upto;                           // dirty forward starts before the first other dirty node
            if (mfs & (mfs - 1)) {                                   // two or more: the batch (one alone takes the path below)
                const bool me = (mfs >> lane) & 1u;
                const bool inr = me && (!CH || mj < ne);
                int turn = 0;
                T wme = D::inf();
                if (me) {
                    dirty[j0] = 0;
                    const unsigned grp = __match_any_sync(mfs, mj);  // the starts of my family in the batch
                    turn = __popc(grp & ((1u << lane) - 1u));
                    if (inr) wme = D::add(Tj, D::load_w(B.o_wint + oj));
                }
                const int turns = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(me ? turn : 0));
                for (int t = 0; t <= turns; t++) {
                    if (inr && turn == t) {
                        const T cur = dist[mj];
                        relax<D, CH>(B, ties, dist, mj, cur, wme, j0, dirty);
                    }
                    __syncwarp();
                }
                continue;                                            // (same window again: the stops may lie inside it)
            }
        }
        const i32 u = i + f;
        const u32 wu = __shfl_sync(mask, wj, f, NL);
        const T Du = shfl128<NL>(mask, Tj, f);
        const i32 mate_u = __shfl_sync(mask, mj, f, NL), orf_u = __shfl_sync(mask, oj, f, NL);
        const int kind = (int)(wu & 3), pu = (int)(wu >> 4);
        if (lane == 0) dirty[u] = 0;
        i32 rewind = 0x7FFFFFFF;
        const bool behind = in && lane > f;                 // a node after u inside the window
        if (kind == K_FSTART) {
            if (lane == 0 && (!CH || mate_u < ne)) {
                const T cur = dist[mate_u];
                relax<D, CH>(B, ties, dist, mate_u, cur, D::add(Du, D::load_w(B.o_wint + orf_u)), u, dirty);
            }
        } else if (kind == K_RSTOP) {
            // the starts of this reverse family: nodes up to its farthest start (n_mate of the stop-key node)
            if (behind && j0 <= mate_u && (int)(wj & 3) == K_RSTART && mj == u)
                relax<D, CH>(B, ties, dist, j0, Tj, D::add(Du, D::load_w(B.o_wint + oj)), u, dirty);
            const i32 last = (CH && mate_u >= ne) ? ne - 1 : mate_u;
            for (i32 j = i + NL + lane; j <= last; j += NL) {
                const u32 w2 = pk[j];
                const i32 m2 = B.n_mate[j];
                if ((int)(w2 & 3) == K_RSTART && m2 == u) {
                    const i32 orf = B.n_orf[j];
                    const T cur = dist[j];
                    relax<D, CH>(B, ties, dist, j, cur, D::add(Du, D::load_w(B.o_wint + orf)), u, dirty);
                }
            }
        } else {
            const u32 ovb = __shfl_sync(mask, cj, f, NL);
            u32 ove = __shfl_sync(mask, cj, (f + 1) & (NL - 1), NL);
            if (f == NL - 1 || u + 1 >= ne) ove = B.ov_cnt[u + 1];
            // gap edges to entries within 500 bp downstream (functions.py:360-438)
            {
                const int kj = (int)(wj & 3), d = (int)(wj >> 4) - pu;
                if (behind && d > 0 && d < 500 && kind_is_entry(kj) && !(kind == K_RSTART && kj == K_FSTART && d <= 2)) {
                    const bool diff = (kind == K_FSTOP) ? (kj == K_RSTOP) : (kj == K_FSTART);
                    bool o;
                    relax<D, CH>(B, ties, dist, j0, Tj, D::add(Du, D::from_i64(gap_w64(B, c, d - 3, diff, &o))), u, dirty);
                }
            }
            const u32 wlast = __shfl_sync(mask, wj, NL - 1, NL);
            if (i + NL < ne && (int)(wlast >> 4) - pu < 500) {
                for (i32 j = i + NL + lane; j < ne; j += NL) {
                    const u32 w2 = pk[j];
                    const T cur = dist[j];
                    const int kj = (int)(w2 & 3), d = (int)(w2 >> 4) - pu;
                    if (d >= 500) break;
                    if (d <= 0 || !kind_is_entry(kj)) continue;
                    const bool diff = (kind == K_FSTOP) ? (kj == K_RSTOP) : (kj == K_FSTART);
                    if (kind == K_RSTART && kj == K_FSTART && d <= 2) continue;      // functions.py:431
                    bool o;
                    relax<D, CH>(B, ties, dist, j, cur, D::add(Du, D::from_i64(gap_w64(B, c, d - 3, diff, &o))), u, dirty);
                }
            }
            // overlap edges (backwards)
            if (ove > ovb) {
                for (u32 k = ovb + lane; k < ove; k += NL) {
                    const i64 w64 = B.ov_w64[k];
                    const i32 v = B.ov_dst[k];
                    if (CH && v < nb) continue;
                    const T cur = dist[v];
                    const T cand = D::add(Du, w64 != OV_W64_WIDE ? D::from_i64(w64) : D::load_w(B.ov_wint + k));
                    if (relax<D, CH>(B, ties, dist, v, cur, cand, u, dirty) && v < rewind) rewind = v;
                }
                rewind = (i32)__reduce_min_sync(mask, (unsigned)rewind);
            }
            // bridges (rare: n_brs flags the exit nodes that have any)
            if (bre > brb && (B.n_brs[u] & 1)) {
                for (u32 k = brb + lane; k < bre; k += NL) {
                    if (B.br_src[k] != u) continue;
                    const i32 v = B.br_dst[k];
                    if (CH && v >= ne) continue;
                    const T cur = dist[v];
                    relax<D, CH>(B, ties, dist, v, cur, D::add(Du, D::load_w(B.br_wint + k)), u, dirty);
                }
            }
            // exit -> target within 2000 bp of the right end (functions.py:448-451)
            if (!CH && L - pu <= 2000) {
                bool o;
                const T cand = D::add(Du, D::from_i64(gap_w64(B, c, L - pu, false, &o)));
                okw = okw && o;
                if (D::less(cand, tdist)) {
                    tdist = cand;
                    tpar = u;
                } else if (lane == 0 && !D::is_inf(tdist) && D::eq(cand, tdist) && tpar != u) {
                    ties++;
                    D::tie(B, c, -3 - c, u, cand);
                }
            }
        }
        __syncwarp(mask);
        i = (rewind < u) ? rewind : u + 1;
    }
    if (CH) return;
    if (ties) atomicAdd(&cs->n_ties, ties);
    if (!okw && lane == 0) atomicOr(&cs->err, (u32)ERR_OVERFLOW);
    if (lane == 0) {
        B.tdist[c] = D::to_wint(tdist);
        B.tparent[c] = tpar;
    }
}
#endif
PB_HD bool contig_is_wide(const Batch& B, int c) { return B.cs[c].wide || (B.flags & PB200_SOLVE_WIDE); }
PB_HD bool contig_is_huge(const Batch& B, int c) { return B.cs[c].huge != 0; }
// a long contig whose solve runs as chunks (chunk.cuh) instead of one sweep
PB_HD bool contig_chunked(const Batch& B, int c) { return B.ch_cnt && B.ch_cnt[c + 1] > B.ch_cnt[c] && !contig_is_wide(B, c); }
// ... and taking part in the current attempt (the second one is only for the contigs whose first failed its check)
PB_HD bool chunk_active(const Batch& B, int c) { return contig_chunked(B, c) && (B.ch_round == 0 || B.cs[c].chunk_retry); }
PB_HDN void solve_contig(const Batch& B, int c, int lane, int NL) {
    if (contig_is_huge(B, c)) solve_contig_t<DHuge>(B, c, lane, NL);
    else if (contig_is_wide(B, c)) solve_contig_t<D256>(B, c, lane, NL);
    else if (!contig_chunked(B, c)) solve_contig_t<D128>(B, c, lane, NL);
}

// Stage 12b: exact ties.  Distances do not depend on the relaxation order, parents do: where two edges into a node
// offer the same final distance, the reference's solver keeps the one it processes first -- Bellman-Ford passes over
// the edges in Graph.iteredges() order, i.e. grouped by source node in node insertion order sigma (graphs.py:121-126,
// functions.py:311-318), strict '<'.  A node x whose parent is y reached its final distance while y's group was
// processed, so x's own group offers it in the same pass if sigma(x) > sigma(y) and one pass later otherwise; hence
//     pass(x -> v) = 1 + number of edges y -> z on the parent chain source .. x -> v's source side with sigma(z) < sigma(y)
// and the reference's parent of v is the tight in-edge with the smallest (pass, sigma(x)).  The tight in-edges of v are
// the sweep's parent plus the recorded ties whose value is the final distance.  Only contigs with ties do any work.
// (Checked against the oracle's edge-order Bellman-Ford: tests/test_certified.py, tests/tie_audit.py.)
#define TIE_MAXN 256          /* ties of one contig settled out of thread-local arrays */
#define TIE_BIGN 16384        /* ... out of the batch-wide scratch; beyond this (an exact repeat over megabases) ERR_TIES */
// insertion index of node a inside its family's run of the node order: forward family = nearest start, stop, then the
// other starts by descending position; reverse family = stop-key node, then the starts by ascending position
PB_HDN int tie_sigma_idx(const Batch& B, i32 a, i32 fam) {
    if ((B.n_kind[fam] & 3) == K_FSTOP) {
        if (a == fam) return 1;
        int r = 1;
        for (i32 j = a + 1; j < fam; j++)
            if ((B.n_kind[j] & 3) == K_FSTART && B.n_mate[j] == fam) r++;
        return r == 1 ? 0 : r;
    }
    if (a == fam) return 0;
    int r = 0;
    for (i32 j = fam + 1; j <= a; j++)
        if ((B.n_kind[j] & 3) == K_RSTART && B.n_mate[j] == fam) r++;
    return r;
}
// sigma(a) < sigma(b); -2 = the source node, inserted after every CDS node (functions.py:440-443)
PB_HDN bool tie_sigma_less(const Batch& B, i32 a, i32 b) {
    if (a == b || a == -2) return false;
    if (b == -2) return true;
    // tRNA nodes (indices from nn) are inserted by add_trnas after every CDS node, entry then exit per hit (functions.py:496-508)
    if (a >= B.nn || b >= B.nn) return (a >= B.nn && b >= B.nn) ? a < b : b >= B.nn;
    const int ka = B.n_kind[a] & 3, kb = B.n_kind[b] & 3;
    const i32 fa = (ka == K_FSTOP || ka == K_RSTOP) ? a : B.n_mate[a];
    const i32 fb = (kb == K_FSTOP || kb == K_RSTOP) ? b : B.n_mate[b];
    const i32 ta = B.n_trig[fa], tb = B.n_trig[fb];
    if (ta != tb) return ta < tb;
    if (fa != fb) return fa < fb;                 // (two families emitted at one scan position: custom stop codon sets)
    return tie_sigma_idx(B, a, fa) < tie_sigma_idx(B, b, fb);
}
// pass in which the edge x -> (some node) offers x's final distance; -1: an unsettled tie node lies on the chain,
// -2: the chain is broken
// (flag != null: flag[y] says "y is an unsettled tie node" instead of the scan over the arrays)
PB_HDN int tie_chain_pass(const Batch& B, i32 x, const i32* tv, const u8* done, int n, i32 limit, const u8* flag) {
    if (x == -2) return 1;
    int cnt = 1;
    i32 y = x;
    for (i32 steps = 0; steps <= limit; steps++) {
        if (flag) {
            if (flag[y]) return -1;
        } else
            for (int a = 0; a < n; a++)
                if (!done[a] && tv[a] == y) return -1;
        const i32 py = B.parent[y];
        if (py == -2) return cnt + 1;             // sigma(y) < sigma(source) always
        if (py < 0) return -2;
        if (tie_sigma_less(B, y, py)) cnt++;
        y = py;
    }
    return -2;
}
// The candidates of one tie node usually share an ancestor a few edges back, and everything above it is common to
// their pass counts: walk each parent chain for at most TIE_K nodes, remembering the descents so far.
//   node[0] = x, node[i+1] = parent(node[i]) (the source, -2, ends the list); cum[i] = descents on the edges between
//   node[i] and x.  Returns the length, -1 if an unsettled tie node lies on the walked part, -2 if the chain is broken.
#define TIE_K 48
#define TIE_K_SHORT 6
PB_HDN int tie_walk(const Batch& B, i32 x, i32* node, int* cum, const i32* tv, const u8* done, int n, int kmax, const u8* flag) {
    int len = 0, c = 0;
    i32 y = x;
    while (len < kmax) {
        node[len] = y;
        cum[len] = c;
        len++;
        if (y == -2) break;
        if (flag) {
            if (flag[y]) return -1;
        } else
            for (int a = 0; a < n; a++)
                if (!done[a] && tv[a] == y) return -1;
        const i32 py = B.parent[y];
        if (py < 0 && py != -2) return -2;
        if (tie_sigma_less(B, y, py)) c++;
        y = py;
    }
    return len;
}
// thread the recorded events into per-contig lists.  item = event slot
PB_HDN void st_tie_link(const Batch& B, i64 k) {
    u32 n = *B.tie_n;
    if (n > (u32)B.tie_cap) n = (u32)B.tie_cap;
    if (k >= (i64)n) return;
    TieEv* e = B.tie_ev + k;
    const int c = (e->v <= -3) ? (-3 - e->v) : B.n_contig[e->v];
    if (e->v <= -3) e->v = -3;
    // (pad bit 1: recorded by the chunked solve's check, bit 4 = in its second attempt -- void once the contig fell back to
    // the one-warp sweep, or when it belongs to an attempt that was repeated)
    if ((e->pad & 2) && (B.cs[c].chunk_viol || ((e->pad >> 4) & 1) != (B.cs[c].chunk_retry ? 1 : 0))) return;
    if (e->pad & 1) {
        const u32 ext = (e->cand.w[3] >> 31) ? 0xFFFFFFFFu : 0u;
        for (int i = 4; i < WN; i++) e->cand.w[i] = ext;
    }
    e->next = PB_ATOMIC_EXCH(&B.cs[c].tie_head, (i32)k + 1);
}
PB_HDN void st_tie_fix(const Batch& B, i64 c64) {
    if (c64 >= B.nc) return;
    const int c = (int)c64;
    CStat* cs = B.cs + c;
    if (!cs->tie_head) return;
    const bool wide = contig_is_wide(B, c);
    const i32 nodes = B.cnode[c + 1] - B.cnode[c] + (B.nt > 0 ? 2 * (B.ctrna[c + 1] - B.ctrna[c]) : 0);
    // the events at FINAL distances are the second (third ...) tight in-edges: counted first, then laid out oldest first --
    // the sweep records in position order, so a tie node's chain mostly runs over nodes settled before it
    int n = 0;
    for (i32 k = cs->tie_head; k;) {
        TieEv* e = B.tie_ev + (k - 1);
        const WInt fin = (e->v == -3) ? B.tdist[c]
                         : contig_is_huge(B, c) ? DHuge::to_wint(B.dist_huge[e->v])
                         : (wide ? B.dist[e->v] : D128::to_wint(B.dist128[e->v]));
        e->pad = (w_cmp(fin, e->cand) == 0) ? 8 : 0;               // (pad is free from here on: 8 = tight at the final distance)
        n += e->pad ? 1 : 0;
        k = e->next;
    }
    if (n == 0) return;
    i32 tvl[TIE_MAXN], tfl[TIE_MAXN];
    u8 donel[TIE_MAXN];
    i32* tv = tvl;
    i32* tf = tfl;
    u8* done = donel;
    u8* flag = nullptr;
    if (n <= TIE_MAXN) {
        // small (every ordinary contig): thread-local arrays, filled oldest first
        int at = n;
        for (i32 k = cs->tie_head; k;) {
            const TieEv* e = B.tie_ev + (k - 1);
            if (e->pad == 8) {
                --at;
                tvl[at] = e->v;
                tfl[at] = e->from;
                donel[at] = 0;
            }
            k = e->next;
        }
    } else {
        // large (exact tandem repeats over tens of kb: a tie per repeat unit): a region of the batch-wide scratch, and a
        // per-node flag "unsettled tie node" (the solve's dirty flags: all clear once a sweep has ended) instead of scans
        const u32 base = n <= TIE_BIGN ? PB_ATOMIC_ADD_RET(B.tie_n + 1, (u32)n) : 0u;
        if (n > TIE_BIGN || base + (u32)n > (u32)B.tie_cap) {
            cs->err |= ERR_TIES;
            return;
        }
        tv = B.tie_tv + base;
        tf = B.tie_tf + base;
        done = B.tie_done + base;
        flag = B.dirty;
        int at = n;
        for (i32 k = cs->tie_head; k;) {
            const TieEv* e = B.tie_ev + (k - 1);
            if (e->pad == 8) {
                --at;
                tv[at] = e->v;
                tf[at] = e->from;
                done[at] = 0;
                if (e->v >= 0) B.dirty[e->v] = 1;
            }
            k = e->next;
        }
    }
    i32 ynode[TIE_K], znode[TIE_K];
    int ycum[TIE_K], zcum[TIE_K];
    for (int round = 0; round <= n; round++) {
        bool open = false;
        for (int a = 0; a < n; a++) {
            if (done[a]) continue;
            const i32 v = tv[a];
            i32 best = (v == -3) ? B.tparent[c] : B.parent[v];
            // the chains of a node's candidates usually meet within a few edges: a short walk first, the long one only
            // if no common node turns up (each step is a handful of dependent loads for a lone thread)
            bool wait = false, broken = false;
            for (int b = a; b < n && !wait && !broken; b++) {
                if (done[b] || tv[b] != v) continue;
                const i32 x = tf[b];
                int dy = -1, dz = -1;             // descents of both chains below their first common node
                for (int kmax = TIE_K_SHORT; dy < 0 && kmax <= TIE_K && !wait && !broken; kmax = (kmax == TIE_K ? TIE_K + 1 : TIE_K)) {
                    const int ylen = tie_walk(B, best, ynode, ycum, tv, done, n, kmax, flag);
                    const int zlen = ylen < 0 ? ylen : tie_walk(B, x, znode, zcum, tv, done, n, kmax, flag);
                    if (ylen == -1 || zlen == -1) wait = true;
                    else if (ylen == -2 || zlen == -2) broken = true;
                    else
                        for (int j = 0; j < zlen && dy < 0; j++)
                            for (int i = 0; i < ylen; i++)
                                if (ynode[i] == znode[j]) {
                                    dy = ycum[i];
                                    dz = zcum[j];
                                    break;
                                }
                }
                if (wait || broken) break;
                if (dy < 0) {                     // no common node within TIE_K: the absolute pass numbers
                    dy = tie_chain_pass(B, best, tv, done, n, nodes, flag);
                    dz = tie_chain_pass(B, x, tv, done, n, nodes, flag);
                    if (dy == -1 || dz == -1) {
                        wait = true;
                        break;
                    }
                    if (dy == -2 || dz == -2) {
                        broken = true;
                        break;
                    }
                }
                if (dz < dy || (dz == dy && tie_sigma_less(B, x, best))) best = x;
            }
            if (broken) {
                cs->err |= ERR_TIES;
                break;
            }
            if (wait) {
                open = true;
                continue;
            }
            if (v == -3) B.tparent[c] = best;
            else B.parent[v] = best;
            if (flag && v >= 0) B.dirty[v] = 0;
            for (int b = a; b < n; b++)
                if (tv[b] == v) done[b] = 1;
        }
        if (!open || (cs->err & ERR_TIES)) break;
        if (round == n) cs->err |= ERR_TIES;      // ties that wait on each other (a zero-weight cycle of tight edges)
    }
    if (flag)                                     // leave the flags as the sweep left them
        for (int a = 0; a < n; a++)
            if (tv[a] >= 0) B.dirty[tv[a]] = 0;
}

// Stage 13: walk the parent pointers back from the target; the ORF edges on the path are the calls
// (phanotate.py:65-76: source dropped, then consecutive non-overlapping pairs).  item = contig
PB_HDN void st_backtrack(const Batch& B, i64 c64) {
    if (c64 >= B.nc) return;
    const int c = (int)c64;
    if (contig_chunked(B, c)) return;             // chunk.cuh: st_pj_* trace the path of a long contig in parallel
    CStat* cs = B.cs + c;
    const i32 nb = B.cnode[c], ne = B.cnode[c + 1];
    i32* out = B.call_tmp + call_base(B, c);
    const i32 cap = call_base(B, c + 1) - call_base(B, c);
    i32 n = 0;
    i32 x = B.tparent[c];
    if (x < 0) {
        if (ne > nb) cs->err |= ERR_NOPATH;       // a graph with CDS nodes but no source->target path
        B.call_cnt[c] = 0;
        cs->n_calls = 0;
        return;
    }
    while (x >= 0 && n < cap) {
        i32 e = B.parent[x];                       // entry node of the ORF edge e -> x
        if (e < 0) {
            cs->err |= ERR_INTERNAL;
            break;
        }
        // (the next step's load goes out before this step's store, which the compiler must assume may alias it: two
        // dependent round trips per call instead of three)
        const i32 xn = B.parent[e];                // exit node of the connector into e, or -2 = source
        const i32 orf = ((B.n_kind[x] & 3) == K_RSTART) ? B.n_orf[x] : B.n_orf[e];
        out[n++] = orf;
        x = xn;
    }
    for (i32 a = 0, b = n - 1; a < b; a++, b--) {
        i32 t = out[a];
        out[a] = out[b];
        out[b] = t;
    }
    B.call_cnt[c] = (u32)n;
    cs->n_calls = n;
}
// ORF of call k (calls are numbered contig by contig in path order).  item = call
PB_HDN void st_call_orf(const Batch& B, i64 k) {
    if (k >= B.ncalls) return;
    int lo = 0, hi = B.nc;                 // contig of call k: call_cnt[lo] <= k < call_cnt[lo+1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((i64)B.call_cnt[mid] <= k) lo = mid;
        else hi = mid;
    }
    const int c = lo;
    B.call_orf[k] = B.call_tmp[call_base(B, c) + (k - B.call_cnt[c])];
}
// call table row: entry/exit node positions as phanotate.py:71-75 + locus.py:29-30 produce them.  item = call
PB_HDN void st_gather_calls(const Batch& B, i64 k) {
    if (k >= B.ncalls) return;
    const i32 orf = B.call_orf[k];
    CallRec r;
    if (orf < 0) {                               // a tRNA on the path (functions.py:508: weight -20); strand +-2 marks the gene
        const i32 t = -1 - orf, e = B.nn + 2 * t;
        r.contig = B.t_contig[t] + B.contig_base;
        r.left = B.n_pos[e];
        r.right = B.n_pos[e + 1] + 2;
        r.strand = B.t_start[t] < B.t_stop[t] ? 2 : -2;
        r.weight = dec_from_u64(20);             // -Decimal(20)
        r.weight.neg = 1;
        r.score = -20.0;
        B.calls[k] = r;
        return;
    }
    const int c = B.o_contig[orf];
    const bool rev = B.o_frame[orf] < 0;
    r.contig = (i32)c + B.contig_base;
    r.left = rev ? B.o_stop[orf] : B.o_start[orf];            // left = entry node position
    r.right = (rev ? B.o_start[orf] : B.o_stop[orf]) + 2;     // right = exit node position + 2
    r.strand = rev ? -1 : 1;
    if (B.o_lit[orf] == 1) {
        r.weight = B.o_weight[orf];
        bool ok;
        r.score = dec_to_double(r.weight, &ok);
        if (!ok) PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_RANGE);
    } else {                               // certified score, Decimal weight not materialised (no PB200_CALL_WEIGHTS)
        w_zero(r.weight.c);
        r.weight.e = 0;
        r.weight.neg = 0;
        r.score = B.call_score[k];
    }
    B.calls[k] = r;
}

// ------------------------------------------------------------------------------------------------
// Edge list of get_graph for the API mirror / --dump (functions.py:307-454): every edge including
// the implicit gap edges, grouped by the node that emits it.  fill=false counts.  item = node
PB_HDN void edges_of(const Batch& B, i32 u, bool fill) {
    const int c = contig_of_node(B, u);
    const int kind = B.n_kind[u] & 3;
    const int pu = B.n_pos[u];
    const int L = B.cs[c].L;
    const i32 ne = B.cnode[c + 1];
    u32 cnt = 0;
    EdgeRec* out = fill ? B.edges + B.ed_cnt[u] : (EdgeRec*)0;
#define EMIT(S, D, K, W)                 \
    do {                                 \
        if (fill) {                      \
            EdgeRec r;                   \
            r.contig = c;                \
            r.src = (S);                 \
            r.dst = (D);                 \
            r.kind = (K);                \
            r.weight = (W);              \
            out[cnt] = r;                \
        }                                \
        cnt++;                           \
    } while (0)
    const bool TR = B.nt > 0 && B.ctrna[c + 1] > B.ctrna[c];
    const u32 teb = TR ? B.te_cnt[2 * B.ctrna[c]] : 0, tee = TR ? B.te_cnt[2 * B.ctrna[c + 1]] : 0;
    if (u >= B.nn) {                             // a tRNA node (trna.cuh): its edges are in the explicit list
        if (kind_is_entry(kind) && pu <= 2000) EMIT(-2, u, EK_SOURCE, gap_score(B, c, pu, false));
        for (u32 k = teb; k < tee; k++)
            if (B.te_src[k] == u) {
                if (B.te_dst[k] == u + 1 && kind_is_entry(kind)) {
                    Dec w = dec_from_u64(20);          // -Decimal(20)
                    w.neg = 1;
                    EMIT(u, u + 1, EK_TRNA, w);
                } else {
                    EMIT(u, B.te_dst[k], EK_GAP, gap_score(B, c, B.n_pos[B.te_dst[k]] - pu - 3, false));
                }
            }
        if (!kind_is_entry(kind) && L - pu <= 2000) EMIT(u, -3, EK_TARGET, gap_score(B, c, L - pu, false));
    } else if (kind_is_entry(kind)) {
        if (pu <= 2000) EMIT(-2, u, EK_SOURCE, gap_score(B, c, pu, false));
        if (kind == K_FSTART) {
            EMIT(u, B.n_mate[u], EK_ORF, B.o_weight[B.n_orf[u]]);
        } else {
            const int farpos = B.n_pos[B.n_mate[u]];
            for (i32 j = u + 1; j < ne && B.n_pos[j] <= farpos; j++)
                if ((B.n_kind[j] & 3) == K_RSTART && B.n_mate[j] == u) EMIT(u, j, EK_ORF, B.o_weight[B.n_orf[j]]);
        }
    } else {
        for (i32 j = u + 1; j < ne && B.n_pos[j] - pu < 500; j++) {
            const int kj = B.n_kind[j] & 3;
            const int d = B.n_pos[j] - pu;
            if (d <= 0 || !kind_is_entry(kj)) continue;
            bool diff = (kind == K_FSTOP) ? (kj == K_RSTOP) : (kj == K_FSTART);
            if (kind == K_RSTART && kj == K_FSTART && d <= 2) continue;
            EMIT(u, j, EK_GAP, gap_score(B, c, d - 3, diff));
        }
        for (u32 k = B.ov_cnt[u]; k < B.ov_cnt[u + 1]; k++) EMIT(u, B.ov_dst[k], EK_OVERLAP, B.ov_w[k]);
        for (u32 k = B.br_cnt[B.cnode[c]]; k < B.br_cnt[ne]; k++)
            if (B.br_src[k] == u)
                EMIT(u, B.br_dst[k], EK_BRIDGE, gap_score(B, c, B.n_pos[B.br_dst[k]] - B.n_pos[u] - 3, false));
        if (TR && (B.n_brs[u] & 2))
            for (u32 k = teb; k < tee; k++)
                if (B.te_src[k] == u) EMIT(u, B.te_dst[k], EK_GAP, gap_score(B, c, B.n_pos[B.te_dst[k]] - pu - 3, false));
        if (L - pu <= 2000) EMIT(u, -3, EK_TARGET, gap_score(B, c, L - pu, false));
    }
#undef EMIT
    if (!fill) B.ed_cnt[u] = cnt;
}
PB_HDN void st_edge_count(const Batch& B, i64 u) {
    if (u < (i64)B.nn + 2 * B.nt) edges_of(B, (i32)u, false);
}
PB_HDN void st_edge_fill(const Batch& B, i64 u) {
    if (u < (i64)B.nn + 2 * B.nt) edges_of(B, (i32)u, true);
}

// ------------------------------------------------------------------------------------------------
// fastpathz-compatible solver for an arbitrary edge list (phanotate.py:56-64): literal Bellman-Ford,
// edges in the given order, strict '<', passes until nothing changes.  Sequential by definition of
// its tie-breaking; used only by the API mirror, never on the batch path.
struct BFArgs {
    i32 n_nodes, n_edges, source, target;
    const i32* src;
    const i32* dst;
    const WInt* w;
    WInt* dist;
    i32* parent;
    i32* path;      // [n_nodes]
    i32* path_len;
};
PB_HDN void bf_literal(const BFArgs& a) {
    for (i32 i = 0; i < a.n_nodes; i++) {
        a.dist[i] = wint_inf();
        a.parent[i] = -1;
    }
    *a.path_len = 0;
    if (a.source < 0 || a.source >= a.n_nodes || a.target < 0 || a.target >= a.n_nodes) return;
    w_zero(a.dist[a.source]);
    for (i32 pass = 0; pass <= a.n_nodes; pass++) {
        bool changed = false;
        for (i32 k = 0; k < a.n_edges; k++) {
            const WInt du = a.dist[a.src[k]];
            if (wint_is_inf(du)) continue;
            WInt cand = du;
            w_add(cand, a.w[k]);
            if (wint_less(cand, a.dist[a.dst[k]])) {
                a.dist[a.dst[k]] = cand;
                a.parent[a.dst[k]] = a.src[k];
                changed = true;
            }
        }
        if (!changed) break;
    }
    if (wint_is_inf(a.dist[a.target])) return;
    i32 n = 0;
    for (i32 v = a.target; v != -1 && n < a.n_nodes; v = a.parent[v]) a.path[n++] = v;
    for (i32 x = 0, y = n - 1; x < y; x++, y--) {
        i32 t = a.path[x];
        a.path[x] = a.path[y];
        a.path[y] = t;
    }
    *a.path_len = n;
}

// pack the ABI views
struct OrfRec {
    i32 contig, start, stop, frame, rbs_score, trigger, start_weight, node;
    Dec pstop, weight;
};
struct NodeRec {
    i32 contig, position, kind, frame, mate, orf, other_end, trigger;
};
PB_HDN void pack_orf(const Batch& B, i64 oi, OrfRec* out) {
    if (oi >= B.no) return;
    OrfRec r;
    r.contig = B.o_contig[oi];
    r.start = B.o_start[oi];
    r.stop = B.o_stop[oi];
    r.frame = B.o_frame[oi];
    r.rbs_score = B.o_rbs[oi];
    r.trigger = B.n_trig[B.n_mate[B.o_node[oi]]];
    r.start_weight = B.o_sw[oi];
    r.node = B.o_node[oi];
    r.pstop = B.o_pstop[oi];
    r.weight = B.o_weight[oi];
    out[oi] = r;
}
struct ContigRec {        // = pb200_contig
    i32 length;
    u32 err;
    i32 node_off, n_nodes, orf_off, n_orfs, call_off, n_calls, n_ties, wide;
    Dec pstop, pos_max[4], pos_min[4];
    double background_rbs[28], training_rbs[28];
};
PB_HDN void pack_contig(const Batch& B, i64 c, ContigRec* out) {
    if (c >= B.nc) return;
    const CStat* cs = B.cs + c;
    ContigRec r;
    r.length = cs->L;
    r.err = cs->err;
    r.node_off = B.cnode[c];
    r.n_nodes = B.cnode[c + 1] - B.cnode[c];
    r.orf_off = B.corf[c];
    r.n_orfs = B.corf[c + 1] - B.corf[c];
    r.call_off = (i32)B.call_cnt[c];
    r.n_calls = (i32)(B.call_cnt[c + 1] - B.call_cnt[c]);
    r.n_ties = (i32)cs->n_ties;
    r.wide = contig_is_wide(B, (int)c) ? 1 : 0;
    r.pstop = cs->pstop;
    for (int k = 0; k < 4; k++) {
        r.pos_max[k] = cs->pos_max[k];
        r.pos_min[k] = cs->pos_min[k];
    }
    const double ybg = 28.0 + 2.0 * (double)cs->L, ytr = 28.0 + (double)r.n_orfs;     // functions.py:155-156,180-181,254-255
    for (int k = 0; k < 28; k++) {
        r.background_rbs[k] = (1.0 + (double)cs->hist_bg[k]) / ybg;
        r.training_rbs[k] = (1.0 + (double)cs->hist_tr[k]) / ytr;
    }
    out[c] = r;
}
PB_HDN void pack_node(const Batch& B, i64 ni, NodeRec* out) {
    if (ni >= (i64)B.nn + 2 * B.nt) return;
    if (ni >= B.nn) {                            // a tRNA node: frame +-4 (functions.py:498-505), orf = -1 - hit index
        NodeRec t;
        t.contig = contig_of_node(B, (i32)ni);
        t.position = B.n_pos[ni];
        t.kind = B.n_kind[ni] & 3;
        t.frame = 4;
        t.mate = B.n_mate[ni];
        t.orf = B.n_orf[ni];
        t.other_end = B.n_oth[ni];
        t.trigger = 0;
        out[ni] = t;
        return;
    }
    NodeRec r;
    r.contig = contig_of_node(B, (i32)ni);
    r.position = B.n_pos[ni];
    r.kind = B.n_kind[ni] & 3;
    r.frame = B.n_kind[ni] >> 2;
    r.mate = B.n_mate[ni];
    r.orf = B.n_orf[ni];
    r.other_end = B.n_oth[ni];
    r.trigger = (r.kind == K_FSTOP || r.kind == K_RSTOP) ? B.n_trig[ni] : B.n_trig[B.n_mate[ni]];
    out[ni] = r;
}
