// The PHANOTATE hot path as data-parallel stages over a batch of contigs.
//
// Every stage is a per-item function (item = strip of bases, base position, node, ORF, contig ...)
// marked __host__ __device__: kernels.cu wraps each in a grid-stride __global__ kernel; the unit
// tests compile the very same functions for the host (tests/native/hostsim.cpp) so that the stage
// logic can be checked against the oracle on a CPU-only box.  The product only ever runs the CUDA
// build.  Coordinates are 1-based inside a contig, as in the reference; file:line citations are
// into /root/reference.
//
// HBM layout (structure of arrays, contigs concatenated):
//   per base   : seq (input, 1 B), RBS scores (2 B), 26 bit masks (3.25 B; the GC-frame factor class of every codon start is five of them per strand)
//   per 64 bp  : rank_nodes / rank_orfs (exclusive prefix counts -> node / ORF index of a position)
//   per node   : position, kind, mate, ORF id, trigger, other_end, pstop index (sorted by contig, position)
//   per ORF    : start, stop, frame, rbs score, start-codon weight id, pstop, weight (Dec), integer weight
//   per contig : CStat (counts, histograms, pstop, RBS weights, GC-frame exponents, ln g), gap tables
#pragma once
#include "fxpow.cuh"
#include "frepr.cuh"
#include "ddmath.cuh"

// ------------------------------------------------------------------------------------------------
enum { CLS_NONE = 0, CLS_S = 1, CLS_s = 2, CLS_T = 3, CLS_t = 4 };
enum { FLAG_SCAN_REFERENCE = 8 };   // = PB200_SCAN_REFERENCE of the public header (asserted in pb200.cu)
enum { K_FSTART = 0, K_FSTOP = 1, K_RSTOP = 2, K_RSTART = 3 };   // entry: FSTART,RSTOP  exit: FSTOP,RSTART
enum {
    ERR_CHAR = 1,        // letter outside the 15 IUPAC codes -> KeyError (functions.py:20-24,169)
    ERR_RANGE = 2,       // arithmetic outside the supported range
    ERR_PARALLEL = 4,    // ValueError("parallel edges are forbidden") (graphs.py:73-74)
    ERR_OVERFLOW = 8,    // integer distance does not fit the wide type
    ERR_NOPATH = 16,     // target not reachable (behaviour of the absent fastpathz is undefined)
    ERR_INTERNAL = 32,
    ERR_LOOKUP = 64,     // ValueError from Orfs.get_orf (orfs.py:62-69)
    ERR_TIES = 128       // exact ties in the solve that the edge-order tie-break (st_tie_fix) could not settle
};
#define GAPN 303          // gap lengths -2..300 (functions.py:36-46)
#define WN 8              // limbs of the exact distances (256 bit)
typedef Wide<WN> WInt;    // two's complement
#define OV_W64_WIDE ((i64)0x7FFFFFFFFFFFFFFFll)      /* ov_w64 marker: the weight is in ov_wint[] */

#ifdef __CUDACC__
typedef uint4 U4;
#else
struct U4 {
    u32 x, y, z, w;
};
#endif

struct Params {
    u8 codon_cls[64];     // CLS_* with the reference's elif priority applied (functions.py:198-215)
    u8 rev_start[64];     // rev_comp(codon) in start_codons
    signed char sw_fwd[64];   // index into startw for codon / -1
    signed char sw_rev[64];   // index into startw for rev_comp(codon) / -1
    Dec startw[8];
    i32 min_orf_len;
    i32 pad;
};

struct CStat {
    i32 L;
    u32 err;
    u32 nAT, nGC;
    u32 hist_bg[28], hist_tr[28];
    u32 cmax[4], cmin[4];
    u32 n_ties, n_relax;
    i32 tie_head;           // 1 + index of the contig's newest tie event (0: none)
    Dec pstop, g, g100;
    SFx ln_g;
    Dec wrbs[28];
    Dec pos_max[4], pos_min[4];
    Fx fmax[4], fmin[4];
    u8 max_one[4], min_one[4];
    i32 n_calls;
    i32 wide;               // some edge weight of the contig needs more than 110 bits: 256-bit distances in the solve
    i32 fast_ok;            // fe[] and the per-bin RBS weights converted to fixed point without loss of range
    DD fe[6];               // pos_max[im] * pos_min[il] of the six GC-frame factor classes (exponent of 1-pstop per codon)
    i64 gap_hi3, gap_hi4;   // trunc((g**100 + len)*1000) - len*1000 for 3- and 4-digit len (functions.py:40-41)
    i32 huge;               // some ORF weight needs more than 256 bits: 2048-bit distances in the solve (score.cuh: HInt)
    i32 huge_pad;
    u32 chunk_viol;         // chunked solve (chunk.cuh): some node failed the Bellman check -> the contig is solved again by one sweep
    u32 chunk_retry;        // ... after a second attempt with a four times longer warm-up (1: this contig is in it)
};

// an equal-distance relaxation seen by the sweep: edge from -> v offered `cand` when dist[v] was already `cand`
struct TieEv {
    i32 v, from, next, pad;   // v = -3: the target; next = 1 + index of the contig's previous event
    WInt cand;
};

struct CallRec {          // one CDS call (contig, left, right, strand, weight, float score)
    i32 contig, left, right, strand;
    Dec weight;
    double score;
};

struct Call24 {           // = pb200_call24: a call row without the Decimal weight (what crosses PCIe / NVLink when only the printed columns are wanted)
    i32 contig, left, right, strand;
    double score;
};

struct EdgeRec {          // = pb200_edge
    i32 contig, src, dst, kind;
    Dec weight;
};
enum { EK_ORF = 0, EK_GAP = 1, EK_OVERLAP = 2, EK_BRIDGE = 3, EK_SOURCE = 4, EK_TARGET = 5, EK_TRNA = 6 };

struct Batch {
    Params P;
    i32 nc;
    i32 flags;
    i64 nb;
    const u8* seq;
    const i64* coff;
    u64* rank;            // per 64-base block: low 32 = nodes before, high 32 = ORFs before
    CStat* cs;
    Dec* gap_same;
    Dec* gap_diff;
    i64* gapi_same;
    i64* gapi_diff;
    // nodes
    i32 nn;
    i32 no;
    i32* cnode;           // [nc+1]
    i32* corf;            // [nc+1]
    i32* n_pos;
    i32* n_contig;        // [nn] contig of the node
    i32* o_contig;        // [no] contig of the ORF
    u8* n_kind;           // K_* | frame<<2
    i32* n_mate;
    i32* n_orf;
    i32* n_trig;
    i32* n_oth;
    i32* n_oidx;
    u32* n_pk;            // [nn] position << 4 | kind | frame << 2: the one word the solve reads per node
    // ORFs
    i32* o_start;
    i32* o_stop;
    signed char* o_frame;
    u8* o_rbs;
    signed char* o_sw;
    i32* o_node;
    Dec* o_pstop;
    Dec* o_weight;
    WInt* o_wint;
    // explicit connectors (overlaps) in CSR by source exit node, and bridges by contig
    u32* ov_cnt;          // [nn+1] counts then exclusive offsets
    u64* ov_mask;         // [nn] which of the 64 nodes in front of an exit node are its overlap targets (counting pass -> filling pass)
    i32 nov;
    i32 nbr;
    i32* ov_dst;
    Dec* ov_w;
    WInt* ov_wint;
    u32* br_cnt;          // [nn+1] bridges starting at an interval-start node, then offsets
    i32* n_reach;         // [nn] last covered base before the node
    i32* br_src;
    i32* br_dst;
    WInt* br_wint;
    // solve
    WInt* dist;
    Wide<64>* dist_huge;  // [nn] distances of the contigs solved at 2048 bits (allocated only when a batch has one)
    i32 n_huge;           // ORF weights beyond 256 bits in the batch
    struct I128* dist128; // [nn] distances of the contigs solved at 128 bits
    i32* parent;
    u8* dirty;
    WInt* tdist;          // [nc] distance of the target
    i32* tparent;         // [nc]
    TieEv* tie_ev;        // [tie_cap] per-contig lists (CStat.tie_head)
    u32* tie_n;           // [0] events recorded, [1] scratch rows handed out by st_tie_fix
    i32* tie_tv;          // [tie_cap] scratch of st_tie_fix for contigs with more than TIE_MAXN exact ties: node,
    i32* tie_tf;          //           source of the tight edge,
    u8* tie_done;         //           settled flag
    i32 tie_cap;
    // calls
    i32* call_tmp;        // [no] ORF ids on the path, per contig region
    u32* call_cnt;        // [nc+1]
    i32* call_orf;        // compacted
    CallRec* calls;
    i32 ncalls;
    i32 nedges;
    u32* ed_cnt;          // [nn+1] out-degree (all edge kinds) then exclusive offsets
    EdgeRec* edges;
    // bit masks over the concatenated batch, one bit per base (word = 64 bases): codon classes and raw letters
    u64* mS;
    u64* ms;
    u64* mT;
    u64* pre4;            // per 64-base word: letters a | c<<16 | g<<32 | t<<48 counted from the start of the word's group of 32 words
    u64* mt;
    u64* bA;
    u64* bC;
    u64* bG;
    u64* bT;
    u64* mNS;             // bit set where a start node sits
    u64* mNK;             // bit set where a stop-key node sits
    u64* mk[4];           // the same by node type: forward start, reverse start, forward stop key, reverse stop key
    u8* rbsf;             // [nb] score_rbs(dna[i:i+21]) per base (the RBS background, functions.py:168)
    u8* rbsr;             // [nb] score_rbs(rev_comp(dna[i:i+21])) (functions.py:169)
    u64* cF[5];           // bit set where the forward codon starting here has GC-frame factor index k (k = 5: the rest)
    u64* cR[5];           // same for the reverse strand
    // node-parallel fill / ORF scoring split
    i32* role_list;       // [nn] node ids: the start nodes in ORF order, then the stop-key nodes
    u64* n_gpos;          // [nn] (global base position << 1) | role (0 start node, 1 stop-key node)
    Dec* o_hold;          // [no] product over the codons (functions.py:286-298)
    Dec* o_x;             // [no] 1 - pstop
    SFx* o_lnx;           // [no] ln(1 - pstop), Q32.192
    Dec* o_A;             // [no*3] x ** pos_max[im]
    SFx* o_lnA;           // [no*3]
    Dec* o_fac;           // [no*6] the six GC-frame factors of every ORF
    struct HoldFac* o_hf; // [no*6] the same, prepared for the fast multiply
    unsigned short* o_bin;   // [no] codon count bin | 0x8000 if the fast path applies
    u32* len_hist;        // [HOLD_BINS+1]
    u32* len_cursor;      // [HOLD_BINS]
    i32* o_order;         // [no] ORF ids sorted by codon count, longest first
    i32* ov_src;          // [nov]
    u8* ov_diff;          // [nov]
    // certified integer weights (fast.cuh): ORFs / overlap edges whose 28-digit Decimal weight is still owed
    i32 nlit;             // entries of lit_ids the literal chain works on
    i32 lit_all;          // literal chain over every ORF (ids = slot)
    i32* lit_ids;         // [<= no] ORF ids; NULL when lit_all
    u32* lit_cnt;         // [4] device counters: 0 = ORFs sent to the literal chain before the solve, 1 = after, 2 = overlap edges
    u8* o_lit;            // [no] 1 once o_pstop / o_weight / o_wint hold the literal 28-digit results
    U4* o_cnt;            // [no] (#a, #t, #g, length) of the ORF's own strand-oriented sequence (certified runs)
    DD* o_v;              // [no] |weight| * 1000 from the closed form (certified runs)
    double* call_score;   // [ncalls] certified float(weight) of a call whose Decimal weight was not materialised
    DD* sw_dd;            // [9] 1000 * start-codon weight (index 8: no start codon -> 1000)
    DD* wr_dd;            // [nc*28] Decimal(str(weight_rbs)) per contig and RBS bin
    i32* ovlit_ids;       // [<= nov] overlap edges routed to the literal power
    i64* ov_w64;          // [nov] integer weight of the overlap edge, or OV_W64_WIDE -> ov_wint[]
    i32 ov_all;           // literal overlap chain over every edge (slot = edge), ov_w[] indexed by edge
    i32 novlit;
    i32 n_lit_pre, n_lit_post, n_ovlit;   // statistics of the last run
    i32 contig_base;      // added to the contig column of the call rows (a caller that splits a batch over contexts)
    i32 gap_dec;          // the Decimal gap tables gap_same / gap_diff are built
    i32 lit_done;         // every ORF has its literal weight (lazy completion ran)
    i32* wcontig;         // [nb/64] contig of the first base of every 64-base word
    u8* n_brs;            // [nn] bit 0: the exit node is the source of a bridge, bit 1: of an edge into a tRNA node
    // tRNA masking (trna.cuh): hits as add_trnas lists them; nodes nn + 2k (entry), nn + 2k + 1 (exit)
    i32 nt;               // tRNAs in the batch
    const i32* t_contig;  // [nt] sorted by contig
    const i32* t_start;
    const i32* t_stop;    // start > stop: reverse strand
    const i32* ctrna;     // [nc+1] first tRNA of every contig
    u32* te_cnt;          // [2nt+1] edges found by each tRNA node, then offsets
    i32* te_src;
    i32* te_dst;
    i64* te_w;
    i32 nte;
    // chunked solve of long contigs (chunk.cuh)
    u32* ch_cnt;          // [nc+1] chunks per contig, then exclusive offsets (0 chunks: the contig is solved by one sweep)
    i32 nch;              // chunks in the batch
    i32 ch_round;         // 0: first attempt of the chunked solve, 1: second attempt (longer warm-up) of the contigs that failed
    i32 ch_core, ch_warm, ch_margin, ch_long;   // nodes per chunk / of warm-up before it / of margin behind it; contigs above ch_long nodes are chunked
    struct I128* ch_dist; // [nch * (ch_warm + ch_core + ch_margin)] private distances of every chunk (relative to its stand-in source)
    u8* ch_dirty;         // same shape
    struct I128* ch_off;  // [nch] what to add to a chunk's distances (first: per-chunk delta against its left neighbour)
    u8* ch_flag;          // [nch] 1: distances are absolute (warm-up reaches the contig start), 2: no anchor found
    i32* ch_contig;       // [nch]
    i32* pj_jump;         // [nn] pointer jumping over the parents of chunked contigs (parallel back-trace): 2^k-th ancestor
    i32* pj_jump2;
    i32* pj_depth;        // [nn] edges between the node and the source
    i32* pj_depth2;
    u8* pj_mark;          // [nn] node lies on the source -> target path
};

#ifdef __CUDA_ARCH__
#define PB_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define PB_ATOMIC_ADD_RET(p, v) atomicAdd((p), (v))
#define PB_ATOMIC_OR(p, v) atomicOr((p), (v))
#define PB_ATOMIC_EXCH(p, v) atomicExch((p), (v))
#else
#define PB_ATOMIC_ADD(p, v) (*(p) += (v))
static inline u32 pb_fetch_add(u32* p, u32 v) {
    u32 o = *p;
    *p += v;
    return o;
}
#define PB_ATOMIC_ADD_RET(p, v) pb_fetch_add((p), (v))
#define PB_ATOMIC_OR(p, v) (*(p) |= (v))
static inline i32 pb_exch(i32* p, i32 v) {
    i32 o = *p;
    *p = v;
    return o;
}
#define PB_ATOMIC_EXCH(p, v) pb_exch((p), (v))
#endif

// the 24 per-base bit masks of a batch, in allocation order
PB_HD u64* mask_array(const Batch& B, int a) {
    switch (a) {
        case 0: return B.mS;
        case 1: return B.ms;
        case 2: return B.mT;
        case 3: return B.mt;
        case 4: return B.bA;
        case 5: return B.bC;
        case 6: return B.bG;
        case 7: return B.bT;
        case 8: return B.mNS;
        case 23: return B.mNK;
        default: return a < 14 ? B.cF[a - 9] : a < 19 ? B.cR[a - 14] : B.mk[a - 19];
    }
}
// zero the last word below nb (its upper half may stay unwritten) and the two pad words of mask a.  item = mask
PB_HDN void st_zero_tails(const Batch& B, i64 a) {
    if (a >= 24) return;
    const i64 nblk = (B.nb + 63) >> 6;
    u64* m = mask_array(B, (int)a);
    for (i64 w = nblk > 0 ? nblk - 1 : 0; w < nblk + 2; w++) m[w] = 0;
}

// ------------------------------------------------------------------------------------------------
// alphabet (functions.py:19-24,159-163)
PB_HD u8 lower(u8 ch) { return (ch >= 'A' && ch <= 'Z') ? (u8)(ch | 0x20) : ch; }
// 0..3 = a,c,g,t ; 4 = other IUPAC ; 5 = invalid
PB_HD int base_code(u8 ch) {
    switch (ch) {
        case 'a': return 0;
        case 'c': return 1;
        case 'g': return 2;
        case 't': return 3;
        case 'n': case 'r': case 'y': case 's': case 'w': case 'k': case 'm': case 'b': case 'v': case 'd': case 'h':
            return 4;
        default: return 5;
    }
}
PB_HD int gc_flag(u8 ch) { return (ch == 'g' || ch == 'c' || ch == 's' || ch == 'b' || ch == 'v') ? 1 : 0; }
PB_HD int base_code_of(u8 raw) { return TBL(ch_code)[raw] & 7; }      // = base_code(lower(raw)), one table load

PB_HD int contig_of(const Batch& B, i64 g) {
    int lo = 0, hi = B.nc;             // coff[lo] <= g < coff[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (B.coff[mid] <= g) lo = mid;
        else hi = mid;
    }
    return lo;
}
// the same through the per-word table (st_word_contig): one load + a step or two instead of a 14-step binary search per
// candidate word / node (the search was ~15 % of st_mark's and ~8 % of st_fill's samples)
PB_HD int contig_at(const Batch& B, i64 g) {
    int c = B.wcontig[g >> 6];
    while (B.coff[c + 1] <= g) c++;
    return c;
}
// contig of the first base of every 64-base word.  item = word
PB_HDN void st_word_contig(const Batch& B, i64 w) {
    if (w >= ((B.nb + 63) >> 6)) return;
    B.wcontig[w] = contig_of(B, w << 6);
}

// ------------------------------------------------------------------------------------------------
// RBS motif score of one window, scalar form (functions.py:48-138).  s = contig bases, window =
// s[i : i+21] clipped at L.  rev: score_rbs(rev_comp(window)).  Handles truncated windows and
// ambiguity codes (a motif only ever matches plain acgt letters).
PB_HDNI int rbs_score_scalar(const u8* s, int L, int i, bool rev) {
    if (i < 0 || i >= L) return 0;
    int W = L - i;
    if (W > 21) W = 21;
    u8 code[21];
    for (int k = 0; k < W; k++) code[k] = (u8)base_code(lower(s[i + k]));
    int best = 0;
    for (int m = 0; m < PB_NMOTIF; m++) {
        int cls = TBL(rbs_motif)[m][0], len = TBL(rbs_motif)[m][1], mc = TBL(rbs_motif)[m][2];
        for (int a = 3; a <= 15; a++) {
            if (a + len > W) break;
            int grp = (a <= 4) ? 1 : (a <= 10) ? 0 : (a <= 12) ? 2 : 3;
            int sc = TBL(rbs_group_score)[grp][cls];
            if (sc <= best) continue;
            bool hit = true;
            for (int t = 0; t < len && hit; t++) {
                int want = (mc >> (2 * (len - 1 - t))) & 3;          // motif letter t
                int have = rev ? code[a + t] : code[W - 1 - a - t];   // s[a+t] of the reversed / complemented window
                if (have > 3) hit = false;
                else if (rev) hit = (3 - have) == want;
                else hit = have == want;
            }
            if (hit) best = sc;
        }
    }
    return best;
}
PB_HD int rbs_group_max(int g, u32 mask);
// one window through the 6-mer tables when it is complete and unambiguous, else the scalar form
PB_HDNI int rbs_score_window(const u8* s, int L, int i, bool rev) {
    if (i < 0 || i >= L) return 0;
    if (i + 21 > L) return rbs_score_scalar(s, L, i, rev);
    u32 idx = 0;
    u32 gm = 0, gl = 0, gh = 0, gf = 0;
    for (int k = 0; k < 21; k++) {
        int cd = base_code(lower(s[i + k]));
        if (cd > 3) return rbs_score_scalar(s, L, i, rev);
        idx = ((idx << 2) | (u32)cd) & 4095u;
        if (k >= 5) {
            const int j = k - 5;                           // 6-mer starting at window offset j
            if (!rev) {
                const u32 m = TBL(rbs_end_mask)[idx];
                if (j <= 2) gf |= m;
                else if (j <= 4) gh |= m;
                else if (j <= 10) gm |= m;
                else if (j <= 12) gl |= m;
            } else {
                const u32 m = TBL(rbs_start_mask)[idx];
                if (j >= 3 && j <= 4) gl |= m;
                else if (j >= 5 && j <= 10) gm |= m;
                else if (j >= 11 && j <= 12) gh |= m;
                else if (j >= 13 && j <= 15) gf |= m;
            }
        }
    }
    int a = rbs_group_max(0, gm), b = rbs_group_max(1, gl), c = rbs_group_max(2, gh), d = rbs_group_max(3, gf);
    int r = a > b ? a : b;
    if (c > r) r = c;
    if (d > r) r = d;
    return r;
}
PB_HD int rbs_group_max(int g, u32 mask) {
    int lo, hi;
    switch (g) {
        case 0: lo = TBL(rbs_g0_lo)[mask & 31]; hi = TBL(rbs_g0_hi)[mask >> 5]; break;
        case 1: lo = TBL(rbs_g1_lo)[mask & 31]; hi = TBL(rbs_g1_hi)[mask >> 5]; break;
        case 2: lo = TBL(rbs_g2_lo)[mask & 31]; hi = TBL(rbs_g2_hi)[mask >> 5]; break;
        default: lo = TBL(rbs_g3_lo)[mask & 31]; hi = TBL(rbs_g3_hi)[mask >> 5]; break;
    }
    return lo > hi ? lo : hi;
}

// GC-frame ordering code of the triple (a,b,c) = (T(p),T(p+1),T(p+2)): three trits
PB_HD int gc_trits(int a, int b, int c) {
    int ab = (a > b) ? 2 : (a == b) ? 1 : 0;
    int ac = (a > c) ? 2 : (a == c) ? 1 : 0;
    int bc = (b > c) ? 2 : (b == c) ? 1 : 0;
    return ab * 9 + ac * 3 + bc;
}
// (imax,imin) from the trit code; rev = triple reversed (functions.py:272-278,291-297; gc_frame_plot.py:7-28)
PB_HD void gc_class(int code, bool rev, int& imax, int& imin) {
    int ab = code / 9, ac = (code / 3) % 3, bc = code % 3;
    if (!rev) {
        bool agb = ab == 2, agc = ac == 2, bgc = bc == 2;
        imax = agb ? (agc ? 1 : 3) : (bgc ? 2 : 3);
        imin = agb ? (bgc ? 3 : 2) : (agc ? 3 : 1);
    } else {   // x=c, y=b, z=a
        bool xgy = bc == 0, xgz = ac == 0, ygz = ab == 0;   // c>b, c>a, b>a
        imax = xgy ? (xgz ? 1 : 3) : (ygz ? 2 : 3);
        imin = xgy ? (ygz ? 3 : 2) : (xgz ? 3 : 1);
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 1: per-base scan (functions.py:158-171 + codon classes of :196-215 + gc_frame_plot.py)
// item = strip of SCAN_STRIP consecutive bases of the concatenated batch
#define SCAN_STRIP 32
PB_HDN void scan_range(const Batch& B, int c, const u8* s, int L, int i0, int i1, i64 cb, int bitoff, u32* mk) {
    CStat* cs = B.cs + c;
    const int n = i1 - i0;
    // ---- window sums Tz(j) = sum_{k=-19..20} gc0(j+3k), j = i0 .. i1+1
    int tz[SCAN_STRIP + 2];
    for (int r = 0; r < 3; r++) {
        int j = i0 + r;
        if (j > i1 + 1) break;
        int sum = 0;
        for (int k = -19; k <= 20; k++) {
            int q = j + 3 * k;
            if (q >= 0 && q < L) sum += gc_flag(lower(s[q]));
        }
        tz[r] = sum;
        for (int jj = j + 3; jj <= i1 + 1; jj += 3) {
            int qa = jj + 60, qs = jj - 60;
            if (qa >= 0 && qa < L) sum += gc_flag(lower(s[qa]));
            if (qs >= 0 && qs < L) sum -= gc_flag(lower(s[qs]));
            tz[jj - i0] = sum;
        }
    }
    // ---- base codes of [i0, i1+20]
    u8 code[SCAN_STRIP + 21];
    int ncode = n + 20;
    bool bad = false;
    for (int k = 0; k < ncode; k++) {
        int q = i0 + k;
        int cd = 5;
        if (q < L) {
            cd = base_code(lower(s[q]));
            if (cd == 5 && k < n) bad = true;
            if (cd == 5) cd = 4;
        }
        code[k] = (u8)cd;
    }
    if (bad) PB_ATOMIC_OR(&cs->err, (u32)ERR_CHAR);
    // ---- 6-mer motif masks for 6-mers starting at i0+k (valid only when all six bases are acgt), and the
    //      length of the run of unambiguous letters ending at every position (rolling, one pass)
    unsigned short em[SCAN_STRIP + 16], sm[SCAN_STRIP + 16];
    u8 runs[SCAN_STRIP + 21];
    {
        u32 idx = 0;
        int run = 0;
        for (int j = 0; j < ncode; j++) {
            const int cd = code[j];
            run = (cd < 4) ? (run < 255 ? run + 1 : 255) : 0;
            runs[j] = (u8)run;
            idx = ((idx << 2) | (u32)(cd & 3)) & 4095u;
            if (j >= 5 && j - 5 < n + 15) {
                const bool ok6 = run >= 6;
                em[j - 5] = ok6 ? TBL(rbs_end_mask)[idx] : 0;
                sm[j - 5] = ok6 ? TBL(rbs_start_mask)[idx] : 0;
            }
        }
        for (int k = (ncode >= 5 ? ncode - 5 : 0); k < n + 15; k++) em[k] = sm[k] = 0;
    }
    u32 nAT = 0, nGC = 0;
    for (int k = 0; k < n; k++) {
        int i = i0 + k;
        u8 ch = lower(s[i]);
        if (gc_flag(ch)) nGC++;
        else nAT++;
        int cls = CLS_NONE;
        if (i + 2 < L && code[k] < 4 && code[k + 1] < 4 && code[k + 2] < 4)
            cls = B.P.codon_cls[code[k] * 16 + code[k + 1] * 4 + code[k + 2]];
        int tr = gc_trits(tz[k], tz[k + 1], tz[k + 2]);
        // factor index of the codon starting here, forward strand in bits 0-2, reverse strand in bits 3-5
        const int kf = TBL(gc_fac_index)[0][tr], kr = TBL(gc_fac_index)[1][tr];
        const u32 bit = 1u << (bitoff + k);
        if (kf < 5) mk[8 + kf] |= bit;
        if (kr < 5) mk[13 + kr] |= bit;
        if (cls) mk[cls - 1] |= bit;                 // S, s, T, t
        if (code[k] < 4) mk[4 + code[k]] |= bit;     // a, c, g, t
        // RBS background, both strands
        int sf, sr;
        const bool clean = (i + 21 <= L) && runs[k + 20] >= 21;
        if (clean) {
            u32 gm = em[k + 5] | em[k + 6] | em[k + 7] | em[k + 8] | em[k + 9] | em[k + 10];
            u32 gl = em[k + 11] | em[k + 12];
            u32 gh = em[k + 3] | em[k + 4];
            u32 gf = em[k] | em[k + 1] | em[k + 2];
            int a = rbs_group_max(0, gm), b = rbs_group_max(1, gl), cc = rbs_group_max(2, gh), d = rbs_group_max(3, gf);
            sf = a > b ? a : b;
            if (cc > sf) sf = cc;
            if (d > sf) sf = d;
            gm = sm[k + 5] | sm[k + 6] | sm[k + 7] | sm[k + 8] | sm[k + 9] | sm[k + 10];
            gl = sm[k + 3] | sm[k + 4];
            gh = sm[k + 11] | sm[k + 12];
            gf = sm[k + 13] | sm[k + 14] | sm[k + 15];
            a = rbs_group_max(0, gm), b = rbs_group_max(1, gl), cc = rbs_group_max(2, gh), d = rbs_group_max(3, gf);
            sr = a > b ? a : b;
            if (cc > sr) sr = cc;
            if (d > sr) sr = d;
        } else {
            sf = rbs_score_scalar(s, L, i, false);
            sr = rbs_score_scalar(s, L, i, true);
        }
        if (sf) PB_ATOMIC_ADD(&cs->hist_bg[sf], 1u);
        if (sr) PB_ATOMIC_ADD(&cs->hist_bg[sr], 1u);
        B.rbsf[cb + i] = (u8)sf;
        B.rbsr[cb + i] = (u8)sr;
    }
    PB_ATOMIC_ADD(&cs->nAT, nAT);
    PB_ATOMIC_ADD(&cs->nGC, nGC);
}
PB_HDN void st_scan(const Batch& B, i64 strip) {
    i64 g = strip * SCAN_STRIP;
    if (g >= B.nb) return;
    i64 gend = g + SCAN_STRIP;
    if (gend > B.nb) gend = B.nb;
    int c = contig_of(B, g);
    u32 mk[18];
    for (int k = 0; k < 18; k++) mk[k] = 0;
    const i64 g0 = g;
    while (g < gend) {
        while (B.coff[c + 1] <= g) c++;
        i64 cb = B.coff[c];
        int L = (int)(B.coff[c + 1] - cb);
        int i0 = (int)(g - cb);
        i64 e = gend - cb;
        int i1 = (e < L) ? (int)e : L;
        scan_range(B, c, B.seq + cb, L, i0, i1, cb, (int)(g - g0), mk);
        g = cb + i1;
    }
    ((u32*)B.mS)[strip] = mk[0];
    ((u32*)B.ms)[strip] = mk[1];
    ((u32*)B.mT)[strip] = mk[2];
    ((u32*)B.mt)[strip] = mk[3];
    ((u32*)B.bA)[strip] = mk[4];
    ((u32*)B.bC)[strip] = mk[5];
    ((u32*)B.bG)[strip] = mk[6];
    ((u32*)B.bT)[strip] = mk[7];
    for (int k = 0; k < 5; k++) {
        ((u32*)B.cF[k])[strip] = mk[8 + k];
        ((u32*)B.cR[k])[strip] = mk[13 + k];
    }
}

#include "enum_fwd.inc"
// Stage 3: per 64-base block counts (nodes in the low word, ORFs = start nodes in the high word)
PB_HDN void st_count64(const Batch& B, i64 blk) {
    const u64 sf = B.mk[0][blk], sr = B.mk[1][blk], kf = B.mk[2][blk], kr = B.mk[3][blk];
    B.mNS[blk] = sf | sr;
    B.mNK[blk] = kf | kr;
    if ((sf & sr) | (kf & kr)) {             // one position, two nodes of the same role: cannot happen (exclusive codon classes)
        i64 g = blk << 6;
        if (g >= B.nb) g = B.nb - 1;
        PB_ATOMIC_OR(&B.cs[contig_of(B, g)].err, (u32)ERR_INTERNAL);
    }
    u32 ns = (u32)pb_popc64(B.mNS[blk]), nk = (u32)pb_popc64(B.mNK[blk]);
    B.rank[blk] = (u64)(ns + nk) | ((u64)ns << 32);
}
// node index of (global position g, bit) after the exclusive scan of rank[]; start node sorts first
PB_HD i32 node_index(const Batch& B, i64 g, int bit) {
    const i64 blk = g >> 6;
    const u64 below = (1ull << (g & 63)) - 1ull;
    u32 n = (u32)B.rank[blk] + (u32)pb_popc64(B.mNS[blk] & below) + (u32)pb_popc64(B.mNK[blk] & below);
    if (bit == 1) n += (u32)((B.mNS[blk] >> (g & 63)) & 1ull);
    return (i32)n;
}
PB_HD i32 orf_index(const Batch& B, i64 g) {
    const i64 blk = g >> 6;
    const u64 below = (1ull << (g & 63)) - 1ull;
    return (i32)((u32)(B.rank[blk] >> 32) + (u32)pb_popc64(B.mNS[blk] & below));
}
PB_HDN void st_contig_offsets(const Batch& B, i64 c) {
    if (c > B.nc) return;
    i64 g = B.coff[c];
    if (g >= B.nb) {
        B.cnode[c] = B.nn;
        B.corf[c] = B.no;
    } else {
        B.cnode[c] = node_index(B, g, 0);
        B.corf[c] = orf_index(B, g);
    }
}

#include "enum_fill.inc"
