// Long contigs: the shortest-path solve of ONE contig spread over many warps (BASELINE.json config 5; the same graph and
// solver contract as graph.cuh: functions.py:307-454, phanotate.py:53-65).
//
// A single sweep is a serial chain of ~1 us node visits (10 Mb = 3.6e5 nodes = 0.39 s).  The graph is local -- every
// connector spans < 500 bp, only ORF edges are long -- and shortest-path trees forget their root: a few kb downstream of
// any start every distance is "the same constant + something that does not depend on the start".  So:
//
//   1. the node list of a long contig is cut into chunks of ch_core nodes; every chunk is swept by its own warp from a
//      stand-in source ch_warm nodes (~20 kb) upstream, into private distance arrays (chunk_solve);
//   2. chunk k's distances differ from the true ones by one constant: the difference to chunk k-1 is read off the nodes
//      in front of the cut, which both chunks hold (st_chunk_delta), and summed along the contig (chunk_prefix);
//   3. EVERY node's Bellman equation is then checked against the assembled distances, edge by edge and in parallel
//      (st_lv_*): no in-edge may offer less, and a reachable node needs a tight in-edge.  With positive cycles only,
//      the equations have exactly one solution, so distances that pass ARE the reference's distances -- the speculation
//      in 1-2 is never trusted.  The same pass yields the parents (the tight in-edge) and the exact ties (further tight
//      in-edges) that st_tie_fix settles in the reference's edge order;
//   4. a contig with any violated equation is tried once more with four times the warm-up (st_chunk_retry_mark: sequence
//      with few stops forgets its source over 30-60 kb, not 5-10), and solved by the one-warp sweep if that fails too
//      (solve_fallback).
//
// The back-trace over the parents (st_backtrack: 2 dependent loads per call) becomes pointer jumping (st_pj_*), the
// coverage prefix-maximum behind the bridges (reach_contig) a three-pass scan over the chunks (reach_chunk*).
#pragma once
#include "graph.cuh"

#ifdef __CUDA_ARCH__
#define PB_ATOMIC_CAS(p, c, v) atomicCAS((p), (c), (v))
#else
static inline i32 pb_cas(i32* p, i32 c, i32 v) {
    i32 o = *p;
    if (o == c) *p = v;
    return o;
}
#define PB_ATOMIC_CAS(p, c, v) pb_cas((p), (c), (v))
#endif

PB_HD I128 i128_sub(const I128& a, const I128& b) {
    I128 r;
    r.lo = a.lo - b.lo;
    r.hi = (i64)((u64)a.hi - (u64)b.hi - (a.lo < b.lo ? 1ull : 0ull));
    return r;
}

struct ChunkGeo {
    int c, k;
    i32 nb, ne;       // the contig's nodes
    i32 a, b;         // core: the nodes whose distances the chunk contributes
    i32 s, e;         // swept range: warm-up + core + margin
    i64 slot;         // first entry of the chunk's private arrays
};
PB_HD ChunkGeo chunk_geo(const Batch& B, i32 id) {
    ChunkGeo g;
    g.c = B.ch_contig[id];
    g.k = id - (i32)B.ch_cnt[g.c];
    g.nb = B.cnode[g.c];
    g.ne = B.cnode[g.c + 1];
    g.a = g.nb + g.k * B.ch_core;
    g.b = g.a + B.ch_core < g.ne ? g.a + B.ch_core : g.ne;
    g.s = g.a - B.ch_warm > g.nb ? g.a - B.ch_warm : g.nb;
    g.e = g.b + B.ch_margin < g.ne ? g.b + B.ch_margin : g.ne;
    g.slot = (i64)id * (B.ch_warm + B.ch_core + B.ch_margin);
    return g;
}
PB_HD void chunk_viol(const Batch& B, int c) { PB_ATOMIC_OR(&B.cs[c].chunk_viol, 1u); }

// chunks per contig.  item = contig
PB_HDN void st_chunk_plan(const Batch& B, i64 c) {
    if (c >= B.nc) return;
    const i32 n = B.cnode[c + 1] - B.cnode[c];
    const bool trna = B.nt > 0 && B.ctrna[c + 1] > B.ctrna[c];        // tRNA contigs take the plain sweep (trna.cuh)
    const u32 cnt = (n > B.ch_long && !trna && !(B.flags & PB200_SOLVE_NOCHUNK)) ? (u32)((n + B.ch_core - 1) / B.ch_core) : 0u;
    B.ch_cnt[c] = cnt;
    if (cnt) PB_ATOMIC_ADD(B.lit_cnt + 3, cnt);
}
// contig of every chunk.  item = chunk
PB_HDN void st_chunk_ids(const Batch& B, i64 id) {
    if (id >= B.nch) return;
    int lo = 0, hi = B.nc;                 // ch_cnt[lo] <= id < ch_cnt[lo+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((i64)B.ch_cnt[mid] <= id) lo = mid;
        else hi = mid;
    }
    B.ch_contig[id] = lo;
}
PB_HD SolveRange chunk_range(const Batch& B, const ChunkGeo& g) {
    SolveRange R;
    R.s = g.s;
    R.e = g.e;
    R.dist = B.ch_dist + g.slot - g.s;
    R.dirty = B.ch_dirty + g.slot - g.s;
    R.base = g.s == g.nb ? 0 : (int)(B.n_pk[g.s] >> 4);
    return R;
}
// 1. the sweep of one chunk (host statement; the kernel runs solve_contig_win<32, true>)
PB_HDN void chunk_solve(const Batch& B, i32 id, int lane, int NL) {
    const ChunkGeo g = chunk_geo(B, id);
    if (!chunk_active(B, g.c)) return;
    const SolveRange R = chunk_range(B, g);
    solve_contig_t<D128, true>(B, g.c, lane, NL, &R);
}
// 2a. constant between chunk id and its left neighbour, from the nodes in front of the cut that both hold: the value
// most of (up to) CH_VOTE such nodes agree on.  item = chunk
#define CH_VOTE 12
PB_HDN void st_chunk_delta(const Batch& B, i64 id64) {
    if (id64 >= B.nch) return;
    const i32 id = (i32)id64;
    const ChunkGeo g = chunk_geo(B, id);
    if (!chunk_active(B, g.c)) return;
    I128 zero = D128::from_i64(0);
    B.ch_off[id] = zero;
    if (g.s == g.nb) {
        B.ch_flag[id] = 1;                 // swept from the contig's real source: absolute distances
        return;
    }
    const ChunkGeo gp = chunk_geo(B, id - 1);
    const I128* dk = B.ch_dist + g.slot - g.s;
    const I128* dp = B.ch_dist + gp.slot - gp.s;
    const i32 lo = g.s > gp.s ? g.s : gp.s;
    I128 cand[CH_VOTE];
    int n = 0;
    for (i32 v = g.a - 1; v >= lo && v >= g.a - 256 && n < CH_VOTE; v--) {
        const I128 x = dk[v], y = dp[v];
        if (D128::is_inf(x) || D128::is_inf(y)) continue;
        cand[n++] = i128_sub(y, x);
    }
    if (!n) {
        B.ch_flag[id] = 2;
        chunk_viol(B, g.c);
        return;
    }
    int best = 0, bestn = 0;
    for (int i = 0; i < n; i++) {
        int m = 0;
        for (int j = 0; j < n; j++) m += D128::eq(cand[i], cand[j]) ? 1 : 0;
        if (m > bestn) {
            bestn = m;
            best = i;
        }
    }
    B.ch_off[id] = cand[best];
    B.ch_flag[id] = 0;
}
// 2b. running sum of the deltas along the contig, restarting at every chunk that holds absolute distances.
// item = contig (a warp: each lane sums a run of consecutive chunks, the lanes' totals are combined, then every lane
// writes its run)
PB_HDN void chunk_prefix(const Batch& B, int c, int lane, int NL) {
    if (!chunk_active(B, c)) return;
    const i32 ib = (i32)B.ch_cnt[c], ie = (i32)B.ch_cnt[c + 1];
    const i32 per = (ie - ib + NL - 1) / NL;
    const i32 a = ib + lane * per, b = a + per < ie ? a + per : ie;
    // (sum, reset): what the run adds, and whether it contains a restart (then `sum` counts from the last restart)
    I128 sum = D128::from_i64(0);
    int reset = 0;
    for (i32 id = a; id < b; id++) {
        if (B.ch_flag[id] == 1) {
            sum = D128::from_i64(0);
            reset = 1;
        } else {
            sum = D128::add(sum, B.ch_off[id]);
        }
    }
    I128 carry = D128::from_i64(0);         // value in front of this lane's run
#ifdef __CUDA_ARCH__
    for (int l = 0; l < NL - 1; l++) {      // sequential over 32 lanes: tiny
        const I128 s_l = shfl128<32>(0xFFFFFFFFu, sum, l);
        const int r_l = __shfl_sync(0xFFFFFFFFu, reset, l);
        if (lane > l) carry = r_l ? s_l : D128::add(carry, s_l);
    }
#endif
    I128 acc = carry;
    for (i32 id = a; id < b; id++) {
        if (B.ch_flag[id] == 1) acc = D128::from_i64(0);
        else acc = D128::add(acc, B.ch_off[id]);
        B.ch_off[id] = acc;
    }
}
// 3a. assembled distance of every node of a chunked contig, parents cleared.  item = node
PB_HDN void st_lv_init(const Batch& B, i64 v64) {
    if (v64 >= B.nn) return;
    const i32 v = (i32)v64;
    const int c = B.n_contig[v];
    if (!chunk_active(B, c)) return;
    const i32 id = (i32)B.ch_cnt[c] + (v - B.cnode[c]) / B.ch_core;
    const ChunkGeo g = chunk_geo(B, id);
    const I128 rel = B.ch_dist[g.slot + (v - g.s)];
    B.dist128[v] = D128::is_inf(rel) ? rel : D128::add(rel, B.ch_off[id]);
    B.parent[v] = -1;
    B.dirty[v] = 0;                          // (the one-warp sweep leaves them clear; st_tie_fix borrows them as node flags)
}
// a tight in-edge u -> v beside the parent: the tie record st_tie_fix works from (pad bit 1: made by this stage)
// (v = -3 - c: the contig's target)
PB_HD void lv_tie(const Batch& B, int c, i32 v, i32 from, const I128& cand) {
    PB_ATOMIC_ADD(&B.cs[c].n_ties, 1u);
    TieEv* e = tie_slot(B, c, v, from);
    if (e) {
        e->pad = 3 | (B.ch_round << 4);
        u64* w = (u64*)&e->cand;
        w[0] = cand.lo;
        w[1] = (u64)cand.hi;
    }
}
// 3b. the implicit in-edges of an entry node: source (functions.py:444-447) and the gap edges from the exits within
// 500 bp upstream (functions.py:360-438).  One thread owns parent[v] here; the explicit edges come afterwards.  item = node
PB_HDN void st_lv_node(const Batch& B, i64 v64) {
    if (v64 >= B.nn) return;
    const i32 v = (i32)v64;
    const int c = B.n_contig[v];
    if (!chunk_active(B, c)) return;
    const u32 wv = B.n_pk[v];
    const int kv = (int)(wv & 3), pv = (int)(wv >> 4);
    if (!kind_is_entry(kv)) return;
    const I128 Dv = B.dist128[v];
    i32 par = -1;
    if (pv <= 2000) {
        bool o;
        const I128 cand = D128::from_i64(gap_w64(B, c, pv, false, &o));
        if (D128::less(cand, Dv)) chunk_viol(B, c);
        else if (D128::eq(cand, Dv)) par = -2;
    }
    const i32 nb = B.cnode[c];
    for (i32 j = v - 1; j >= nb; j--) {
        const u32 wj = B.n_pk[j];
        const int kj = (int)(wj & 3), d = pv - (int)(wj >> 4);
        if (d >= 500) break;
        if (d <= 0 || kind_is_entry(kj)) continue;
        if (kj == K_RSTART && kv == K_FSTART && d <= 2) continue;      // functions.py:431
        const I128 Dj = B.dist128[j];
        if (D128::is_inf(Dj)) continue;
        const bool diff = (kj == K_FSTOP) ? (kv == K_RSTOP) : (kv == K_FSTART);
        bool o;
        const I128 cand = D128::add(Dj, D128::from_i64(gap_w64(B, c, d - 3, diff, &o)));
        if (D128::less(cand, Dv)) chunk_viol(B, c);
        else if (D128::eq(cand, Dv)) {
            if (par == -1) par = j;
            else lv_tie(B, c, v, j, cand);
        }
    }
    B.parent[v] = par;
}
// 3c. one explicit edge u -> v of weight w
PB_HD void lv_edge(const Batch& B, int c, i32 u, i32 v, const I128& w) {
    const I128 Du = B.dist128[u];
    if (D128::is_inf(Du)) return;
    const I128 cand = D128::add(Du, w), Dv = B.dist128[v];
    if (D128::less(cand, Dv)) chunk_viol(B, c);
    else if (D128::eq(cand, Dv)) {
        const i32 old = PB_ATOMIC_CAS(&B.parent[v], -1, u);
        if (old != -1 && old != u) lv_tie(B, c, v, u, cand);
    }
}
// ORF edges entry -> exit (functions.py:311-318).  item = ORF
PB_HDN void st_lv_orf(const Batch& B, i64 oi) {
    if (oi >= B.no) return;
    const int c = B.o_contig[oi];
    if (!chunk_active(B, c)) return;
    const i32 sn = B.o_node[oi], kn = B.n_mate[sn];            // start node, stop-key node
    const bool fwd = B.o_frame[oi] > 0;
    lv_edge(B, c, fwd ? sn : kn, fwd ? kn : sn, D128::load_w(B.o_wint + oi));
}
// overlap edges.  item = edge
PB_HDN void st_lv_ov(const Batch& B, i64 k) {
    if (k >= B.nov) return;
    const i32 u = B.ov_src[k];
    const int c = B.n_contig[u];
    if (!chunk_active(B, c)) return;
    const i64 w64 = B.ov_w64[k];
    lv_edge(B, c, u, B.ov_dst[k], w64 != OV_W64_WIDE ? D128::from_i64(w64) : D128::load_w(B.ov_wint + k));
}
// bridges.  item = bridge
PB_HDN void st_lv_br(const Batch& B, i64 k) {
    if (k >= B.nbr) return;
    const i32 u = B.br_src[k];
    const int c = B.n_contig[u];
    if (!chunk_active(B, c)) return;
    lv_edge(B, c, u, B.br_dst[k], D128::load_w(B.br_wint + k));
}
// 3d. a reachable node must have a tight in-edge.  item = node
PB_HDN void st_lv_check(const Batch& B, i64 v64) {
    if (v64 >= B.nn) return;
    const i32 v = (i32)v64;
    const int c = B.n_contig[v];
    if (!chunk_active(B, c)) return;
    if (!D128::is_inf(B.dist128[v]) && B.parent[v] == -1) chunk_viol(B, c);
}
// 3e. exit -> target within 2000 bp of the right end (functions.py:448-451).  item = contig
PB_HDN void st_lv_target(const Batch& B, i64 c64) {
    if (c64 >= B.nc) return;
    const int c = (int)c64;
    if (!chunk_active(B, c)) return;
    const i32 nb = B.cnode[c], ne = B.cnode[c + 1];
    const int L = B.cs[c].L;
    I128 td = D128::inf();
    i32 tp = -1;
    for (int pass = 0; pass < 2; pass++)
        for (i32 u = ne - 1; u >= nb; u--) {
            const u32 wu = B.n_pk[u];
            const int ku = (int)(wu & 3), pu = (int)(wu >> 4);
            if (L - pu > 2000) break;
            if (kind_is_entry(ku)) continue;
            const I128 Du = B.dist128[u];
            if (D128::is_inf(Du)) continue;
            bool o;
            const I128 cand = D128::add(Du, D128::from_i64(gap_w64(B, c, L - pu, false, &o)));
            if (pass == 0) {
                if (!D128::less(td, cand)) {          // cand <= td: the lowest node among equals ends up the parent
                    td = cand;
                    tp = u;
                }
            } else if (u != tp && D128::eq(cand, td)) {
                lv_tie(B, c, -3 - c, u, cand);
            }
        }
    B.tdist[c] = D128::to_wint(td);
    B.tparent[c] = tp;
}
// 3f. between the attempts: contigs whose check failed.  item = contig
PB_HDN void st_chunk_viol_count(const Batch& B, i64 c) {
    if (c >= B.nc) return;
    if (contig_chunked(B, (int)c) && B.cs[c].chunk_viol) PB_ATOMIC_ADD(B.lit_cnt + 6, 1u);
}
// ... they go into the second attempt: a four times longer warm-up forgets the stand-in source where the first did not
// (sequence with few stops -- GC >= 70 % -- coalesces over 30-60 kb instead of 5-10)
PB_HDN void st_chunk_retry_mark(const Batch& B, i64 c) {
    if (c >= B.nc) return;
    CStat* cs = B.cs + c;
    if (contig_chunked(B, (int)c) && cs->chunk_viol) {
        cs->chunk_retry = 1;
        cs->chunk_viol = 0;
        cs->n_ties = 0;
    }
}
// 4. contigs whose assembled distances failed a check: the one-warp sweep (host statement; the kernel runs
// solve_contig_win<32>)
PB_HDN void solve_fallback(const Batch& B, int c, int lane, int NL) {
    if (!contig_chunked(B, c) || !B.cs[c].chunk_viol) return;
    if (lane == 0) {
        B.cs[c].n_ties = 0;
        PB_ATOMIC_ADD(B.lit_cnt + 4, 1u);
    }
    solve_contig_t<D128>(B, c, lane, NL);
}

// ---- back-trace by pointer jumping: jump[v] = 2^r-th ancestor (the root of a tree points at itself), depth[v] = edges
// up to jump[v]; the nodes on the target's chain get marked along the way.  item = node
PB_HDN void st_pj_init(const Batch& B, i64 v64) {
    if (v64 >= B.nn) return;
    const i32 v = (i32)v64;
    const int c = B.n_contig[v];
    if (!contig_chunked(B, c)) return;
    const i32 p = B.parent[v];
    B.pj_jump[v] = p >= 0 ? p : v;
    B.pj_depth[v] = p >= 0 ? 1 : 0;
    B.pj_mark[v] = (v == B.tparent[c]) ? 1 : 0;
}
PB_HDN void st_pj_round(const Batch& B, i64 v64) {
    if (v64 >= B.nn) return;
    const i32 v = (i32)v64;
    const int c = B.n_contig[v];
    if (!contig_chunked(B, c)) return;
    const i32 j = B.pj_jump[v];
    if (B.pj_mark[v]) B.pj_mark[j] = 1;
    B.pj_jump2[v] = B.pj_jump[j];
    B.pj_depth2[v] = B.pj_depth[v] + B.pj_depth[j];
}
// marked exit nodes are the calls: the exit at depth 2k+1 closes call k (phanotate.py:65-76).  item = node
PB_HDN void st_pj_calls(const Batch& B, i64 v64) {
    if (v64 >= B.nn) return;
    const i32 v = (i32)v64;
    const int c = B.n_contig[v];
    if (!contig_chunked(B, c)) return;
    CStat* cs = B.cs + c;
    if (v == B.cnode[c] && B.tparent[c] < 0) {      // (one thread per contig: the unreachable target)
        cs->err |= ERR_NOPATH;
        B.call_cnt[c] = 0;
        cs->n_calls = 0;
    }
    if (!B.pj_mark[v]) return;
    const int kind = (int)(B.n_pk[v] & 3);
    const i32 d = B.pj_depth[v];
    const i32 cap = call_base(B, c + 1) - call_base(B, c);
    const i32 root = B.pj_jump[v];
    if (B.parent[root] != -2) {                     // the chain does not end at the source
        PB_ATOMIC_OR(&cs->err, (u32)ERR_INTERNAL);
        return;
    }
    if (kind_is_entry(kind)) {
        if (d & 1) PB_ATOMIC_OR(&cs->err, (u32)ERR_INTERNAL);
        return;
    }
    if (!(d & 1) || (d >> 1) >= cap) {
        PB_ATOMIC_OR(&cs->err, (u32)ERR_INTERNAL);
        return;
    }
    const i32 e = B.parent[v];
    B.call_tmp[call_base(B, c) + (d >> 1)] = (kind == K_RSTART) ? B.n_orf[v] : B.n_orf[e];
    if (v == B.tparent[c]) {
        B.call_cnt[c] = (u32)((d >> 1) + 1);
        cs->n_calls = (d >> 1) + 1;
    }
}

// ---- coverage prefix maximum of a chunked contig (reach_contig in three passes over its chunks; one warp per chunk /
// per contig).  The tail of n_reach (from nn + 1) holds one value per chunk.
PB_HD void reach_chunk_geo(const Batch& B, i32 id, int& c, i32& a, i32& b) {
    int lo = 0, hi = B.nc;                 // ch_cnt[lo] <= id < ch_cnt[lo+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((i32)B.ch_cnt[mid] <= id) lo = mid;
        else hi = mid;
    }
    c = lo;
    const i32 nb = B.cnode[c], ne = B.cnode[c + 1];
    a = nb + (id - (i32)B.ch_cnt[c]) * B.ch_core;
    b = a + B.ch_core < ne ? a + B.ch_core : ne;
}
// pass 0: maximum over the chunk's core;  pass 2: the nodes' exclusive prefix maxima.  item = chunk (a warp)
PB_HDN void reach_chunk(const Batch& B, i32 id, int pass, int lane, int NL) {
    if (id >= (i32)B.ch_cnt[B.nc]) return;
    int c;
    i32 a, b;
    reach_chunk_geo(B, id, c, a, b);
    i32* slot = B.n_reach + (i64)B.nn + 1 + id;
    if (pass == 0) {
        const int m = reach_range(B, c, a, b, 0, false, lane, NL);
        if (lane == 0) *slot = m;
    } else {
        reach_range(B, c, a, b, *slot, true, lane, NL);
    }
}
// pass 1: exclusive prefix maximum over the chunks of a contig.  item = contig (a warp)
PB_HDN void reach_chunk_prefix(const Batch& B, int c, int lane, int NL) {
    i32* slot = B.n_reach + (i64)B.nn + 1;
    const i32 ib = (i32)B.ch_cnt[c], ie = (i32)B.ch_cnt[c + 1];
    int run = 0;
    for (i32 base = ib; base < ie; base += NL) {
        const i32 i = base + lane;
        const int v = i < ie ? slot[i] : 0;
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
        if (i < ie) slot[i] = excl;
        if (tot > run) run = tot;
    }
}

#ifdef __CUDACC__
// ---- the chunk sweep out of shared memory.  A chunk's sweep is a serial chain of node visits, each a few dependent
// memory round trips: from HBM/L2 that is ~1 us per visit whatever the GPU does beside it.  One warp per block stages
// everything its visits read -- per node: distance, packed word, mate, the weight of the node's own ORF edge, overlap
// range, dirty flag; per overlap edge: weight and target -- in shared memory once (coalesced), sweeps there, and writes
// the distances back.  Same visits, same relaxations as solve_contig_win<32, true>.
#define CHS_WIDE ((i64)0x8000000000000000ll)       /* marker: the ORF weight does not fit 62 bits -> o_wint[] */
PB_HD size_t chunk_smem_bytes(int stride, int ecap) {
    return (size_t)stride * (16 + 8 + 4 + 4 + 4) + 16 + (size_t)ecap * 12 + (size_t)stride + 64;
}
__device__ __forceinline__ void chs_relax(I128* sd, u8* sdirty, i32 v, const I128& cur, const I128& cand) {
    if (D128::less(cand, cur)) {
        sd[v] = cand;
        sdirty[v] = 1;
    }
}
__device__ void solve_chunk_smem(const Batch& B, const ChunkGeo& g, int lane, unsigned char* smem, int stride, int ecap) {
    typedef D128 D;
    typedef I128 T;
    const int c = g.c;
    const i32 s = g.s, n = g.e - g.s;                 // local index = node - s
    T* sd = (T*)smem;
    i64* sw = (i64*)(sd + stride);
    u32* spk = (u32*)(sw + stride);
    i32* smate = (i32*)(spk + stride);
    u32* sov = (u32*)(smate + stride);                // [stride + 1]
    i64* ew = (i64*)(((size_t)(sov + stride + 1) + 15) & ~(size_t)15);
    i32* ed = (i32*)(ew + ecap);
    u8* sdirty = (u8*)(ed + ecap);
    const int base = g.s == g.nb ? 0 : (int)(B.n_pk[g.s] >> 4);
    const u32 ov0 = B.ov_cnt[s];
    // ---- stage
    for (i32 i = lane; i < n; i += 32) {
        const i32 gi = s + i;
        const u32 w = B.n_pk[gi];
        const int kind = (int)(w & 3);
        spk[i] = w;
        smate[i] = B.n_mate[gi];
        sov[i] = B.ov_cnt[gi] - ov0;
        i64 ow = 0;
        if (kind == K_FSTART || kind == K_RSTART) {
            const U4 q = *(const U4*)(B.o_wint + B.n_orf[gi]);
            const i64 lo = (i64)(((u64)q.y << 32) | q.x), hi = (i64)(((u64)q.w << 32) | q.z);
            ow = (hi == (lo >> 63) && lo != CHS_WIDE) ? lo : CHS_WIDE;
        }
        sw[i] = ow;
        T d0 = D::inf();
        u8 f0 = 0;
        if ((int)(w >> 4) - base <= 2000 && kind_is_entry(kind)) {       // source -> entry (functions.py:444-447)
            bool o;
            d0 = D::from_i64(gap_w64(B, c, (int)(w >> 4) - base, false, &o));
            f0 = 1;
        }
        sd[i] = d0;
        sdirty[i] = f0;
    }
    if (lane == 0) sov[n] = B.ov_cnt[s + n] - ov0;
    const u32 ne_ov = B.ov_cnt[s + n] - ov0;
    for (u32 k = lane; k < ne_ov && k < (u32)ecap; k += 32) {
        ew[k] = B.ov_w64[ov0 + k];
        ed[k] = B.ov_dst[ov0 + k];
    }
    __syncwarp();
    const u32 brb = B.br_cnt[g.nb], bre = B.br_cnt[g.ne];
    i32 i = 0;
    int budget = 64 * n + 1024;
    CStat* cs = B.cs + c;
    while (i < n) {
        const i32 j0 = i + lane;
        const bool in = j0 < n;
        u8 dj = 0;
        u32 wj = 0, cj = 0;
        i32 mj = -1;
        T Tj = D::inf();
        if (in) {
            dj = sdirty[j0];
            wj = spk[j0];
            Tj = sd[j0];
            mj = smate[j0];
            cj = sov[j0];
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, dj != 0);
        if (!m) {
            i += 32;
            continue;
        }
        if (--budget < 0) {
            if (lane == 0) atomicOr(&cs->err, (u32)ERR_INTERNAL);
            break;
        }
        const int f = __ffs((int)m) - 1;
        const i32 u = i + f;                              // local
        const u32 wu = __shfl_sync(0xFFFFFFFFu, wj, f);
        const T Du = shfl128<32>(0xFFFFFFFFu, Tj, f);
        const i32 mate_u = __shfl_sync(0xFFFFFFFFu, mj, f) - s;          // local (may lie outside [0, n))
        const int kind = (int)(wu & 3), pu = (int)(wu >> 4);
        if (lane == 0) sdirty[u] = 0;
        i32 rewind = 0x7FFFFFFF;
        const bool behind = in && lane > f;
        if (kind == K_FSTART) {
            if (lane == 0 && mate_u < n) {
                const i64 ow = sw[u];
                const T w = ow != CHS_WIDE ? D::from_i64(ow) : D::load_w(B.o_wint + B.n_orf[s + u]);
                chs_relax(sd, sdirty, mate_u, sd[mate_u], D::add(Du, w));
            }
        } else if (kind == K_RSTOP) {
            // the starts of this reverse family: nodes up to its farthest start (the stop-key node's mate)
            const i32 last = mate_u >= n ? n - 1 : mate_u;
            for (i32 j = u + 1 + lane; j <= last; j += 32) {
                if ((int)(spk[j] & 3) == K_RSTART && smate[j] - s == u) {
                    const i64 ow = sw[j];
                    const T w = ow != CHS_WIDE ? D::from_i64(ow) : D::load_w(B.o_wint + B.n_orf[s + j]);
                    chs_relax(sd, sdirty, j, sd[j], D::add(Du, w));
                }
            }
        } else {
            const u32 ovb = __shfl_sync(0xFFFFFFFFu, cj, f);
            const u32 ove = sov[u + 1];
            // gap edges to entries within 500 bp downstream (functions.py:360-438)
            {
                const int kj = (int)(wj & 3), d = (int)(wj >> 4) - pu;
                if (behind && d > 0 && d < 500 && kind_is_entry(kj) && !(kind == K_RSTART && kj == K_FSTART && d <= 2)) {
                    const bool diff = (kind == K_FSTOP) ? (kj == K_RSTOP) : (kj == K_FSTART);
                    bool o;
                    chs_relax(sd, sdirty, j0, Tj, D::add(Du, D::from_i64(gap_w64(B, c, d - 3, diff, &o))));
                }
            }
            const u32 wlast = __shfl_sync(0xFFFFFFFFu, wj, 31);
            if (i + 32 < n && (int)(wlast >> 4) - pu < 500) {
                for (i32 j = i + 32 + lane; j < n; j += 32) {
                    const u32 w2 = spk[j];
                    const int kj = (int)(w2 & 3), d = (int)(w2 >> 4) - pu;
                    if (d >= 500) break;
                    if (d <= 0 || !kind_is_entry(kj)) continue;
                    const bool diff = (kind == K_FSTOP) ? (kj == K_RSTOP) : (kj == K_FSTART);
                    if (kind == K_RSTART && kj == K_FSTART && d <= 2) continue;      // functions.py:431
                    bool o;
                    chs_relax(sd, sdirty, j, sd[j], D::add(Du, D::from_i64(gap_w64(B, c, d - 3, diff, &o))));
                }
            }
            // overlap edges (backwards)
            if (ove > ovb) {
                for (u32 k = ovb + lane; k < ove; k += 32) {
                    i64 w64;
                    i32 v;
                    if (k < (u32)ecap) {
                        w64 = ew[k];
                        v = ed[k] - s;
                    } else {
                        w64 = B.ov_w64[ov0 + k];
                        v = B.ov_dst[ov0 + k] - s;
                    }
                    if (v < 0) continue;
                    const T cur = sd[v];
                    const T cand = D::add(Du, w64 != OV_W64_WIDE ? D::from_i64(w64) : D::load_w(B.ov_wint + ov0 + k));
                    if (D::less(cand, cur)) {
                        sd[v] = cand;
                        sdirty[v] = 1;
                        if (v < rewind) rewind = v;
                    }
                }
                rewind = (i32)__reduce_min_sync(0xFFFFFFFFu, (unsigned)rewind);
            }
            // bridges (rare: n_brs flags the exit nodes that have any)
            if (bre > brb && (B.n_brs[s + u] & 1)) {
                for (u32 k = brb + lane; k < bre; k += 32) {
                    if (B.br_src[k] != s + u) continue;
                    const i32 v = B.br_dst[k] - s;
                    if (v < 0 || v >= n) continue;
                    chs_relax(sd, sdirty, v, sd[v], D::add(Du, D::load_w(B.br_wint + k)));
                }
            }
        }
        __syncwarp();
        i = (rewind < u) ? rewind : u + 1;
    }
    __syncwarp();
    // ---- the distances back to the chunk's private array
    T* out = B.ch_dist + g.slot;
    for (i32 k = lane; k < n; k += 32) out[k] = sd[k];
}
__global__ void __launch_bounds__(32) k_chunk_solve_smem(const Batch B, int stride, int ecap) {
    extern __shared__ __align__(16) unsigned char chs_smem[];
    const int lane = threadIdx.x;
    for (i64 id = blockIdx.x; id < B.nch; id += gridDim.x) {
        const ChunkGeo g = chunk_geo(B, (i32)id);
        if (!chunk_active(B, g.c)) continue;
        solve_chunk_smem(B, g, lane, chs_smem, stride, ecap);
        __syncwarp();
    }
}
#endif
