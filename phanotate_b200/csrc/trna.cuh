// tRNA masking (reference functions.py:457-509 add_trnas; the tRNA branch of the connect loop functions.py:388-399).
//
// The reference runs aragorn / tRNAscan-SE on the contig and adds, per hit, a node pair (gene 'tRNA', frame +-4) joined by
// an edge of weight -20; in the connect loop a pair of nodes of which at least one is a tRNA node gets ONE kind of edge:
// exit -> entry when 0 < r-l < 500, scored score_gap(r-l-3, 'same') whatever the strands -- no overlap edges, no
// `r-l > 2` exception.  Source / target edges follow the usual rule (functions.py:444-451).  The tools themselves stay
// on the host (they are external programs); what crosses the boundary is their hit list (pb200_set_trnas), in add_trnas'
// order: [start, stop], start > stop on the reverse strand.
//
// tRNA nodes live BEHIND the regular nodes of the batch (indices nn + 2k: entry, nn + 2k + 1: exit), so nothing about
// the position-sorted CDS node tables changes; their edges are an explicit list (te_*), per contig, like the bridges.
#pragma once
#include "graph.cuh"

#define TRNA_W (-20000ll)      /* trunc(-Decimal(20) * 1000) (functions.py:508, edges.py:22) */

PB_HD bool contig_has_trna(const Batch& B, int c) { return B.nt > 0 && B.ctrna[c + 1] > B.ctrna[c]; }
PB_HD i32 trna_lower_bound(const Batch& B, i32 nb, i32 ne, int pos) {       // first node of [nb, ne) at position >= pos
    while (nb < ne) {
        const i32 mid = nb + ((ne - nb) >> 1);
        if (B.n_pos[mid] < pos) nb = mid + 1;
        else ne = mid;
    }
    return nb;
}
// node pair of tRNA k.  item = tRNA
PB_HDN void st_trna_nodes(const Batch& B, i64 k) {
    if (k >= B.nt) return;
    const int c = B.t_contig[k];
    const int s = B.t_start[k], e = B.t_stop[k];
    const bool fwd = s < e;
    const int pe = fwd ? s : e, px = fwd ? e - 2 : s - 2;         // functions.py:496-507
    if (pe < 1 || px < 1 || pe > B.cs[c].L || px > B.cs[c].L || px <= pe) PB_ATOMIC_OR(&B.cs[c].err, (u32)ERR_RANGE);
    const i32 ni = B.nn + 2 * (i32)k;
    for (int j = 0; j < 2; j++) {
        const i32 n = ni + j;
        const int kind = fwd ? (j ? K_FSTOP : K_FSTART) : (j ? K_RSTART : K_RSTOP);
        const int p = j ? px : pe;
        B.n_pos[n] = p;
        B.n_kind[n] = (u8)kind;                                    // frame bits 0: a tRNA node
        B.n_pk[n] = ((u32)p << 4) | (u32)kind;
        B.n_contig[n] = c;
        B.n_mate[n] = ni + 1 - j;
        B.n_orf[n] = -1 - (i32)k;
        B.n_trig[n] = 0;
        B.n_oth[n] = j ? pe : px;
        B.n_oidx[n] = -1;
        B.n_brs[n] = 0;
    }
}
// Edges of tRNA node j (count / fill).  entry: the -20 edge, and the gap edges INTO it from the CDS exits within 500 bp
// upstream; exit: the gap edges out of it to the CDS and tRNA entries within 500 bp downstream.  item = tRNA node
PB_HDN void trna_edges_of(const Batch& B, i64 j, bool fill) {
    if (j >= 2 * (i64)B.nt) return;
    const i32 ni = B.nn + (i32)j;
    const int c = B.n_contig[ni];
    const i32 nb = B.cnode[c], ne = B.cnode[c + 1];
    const int p = B.n_pos[ni];
    u32 cnt = 0;
    u32 at = fill ? B.te_cnt[j] : 0;
#define TE_EMIT(S, D, W)             \
    do {                             \
        if (fill) {                  \
            B.te_src[at + cnt] = (S); \
            B.te_dst[at + cnt] = (D); \
            B.te_w[at + cnt] = (W);  \
        }                            \
        cnt++;                       \
    } while (0)
    bool o;
    if (!(j & 1)) {
        TE_EMIT(ni, ni + 1, TRNA_W);
        for (i32 u = trna_lower_bound(B, nb, ne, p - 499); u < ne && B.n_pos[u] < p; u++) {
            if (kind_is_entry(B.n_kind[u] & 3)) continue;
            TE_EMIT(u, ni, gap_w64(B, c, p - B.n_pos[u] - 3, false, &o));
            if (fill) B.n_brs[u] |= 2;
        }
    } else {
        for (i32 v = trna_lower_bound(B, nb, ne, p + 1); v < ne && B.n_pos[v] - p < 500; v++) {
            if (!kind_is_entry(B.n_kind[v] & 3)) continue;
            TE_EMIT(ni, v, gap_w64(B, c, B.n_pos[v] - p - 3, false, &o));
        }
        for (i32 k = B.ctrna[c]; k < B.ctrna[c + 1]; k++) {
            const i32 v = B.nn + 2 * k;
            const int d = B.n_pos[v] - p;
            if (d > 0 && d < 500) TE_EMIT(ni, v, gap_w64(B, c, d - 3, false, &o));
        }
    }
#undef TE_EMIT
    if (!fill) B.te_cnt[j] = cnt;
}
PB_HDN void st_trna_count(const Batch& B, i64 j) { trna_edges_of(B, j, false); }
PB_HDN void st_trna_fill(const Batch& B, i64 j) { trna_edges_of(B, j, true); }
// the offset tables behind the regular nodes: a tRNA node has no overlap edges and starts no bridge.  item = tail entry
PB_HDN void st_trna_tails(const Batch& B, i64 j) {
    if (j > 2 * (i64)B.nt) return;
    if (j > 0) B.ov_cnt[B.nn + j] = B.ov_cnt[B.nn];
}
