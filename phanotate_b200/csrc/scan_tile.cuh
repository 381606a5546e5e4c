// Stage 1 as a block-cooperative CUDA kernel (the per-strip st_scan in pipeline.cuh is the plain
// statement of the same stage: it stays as the host-side unit-test build and as the device-side
// reference, flag PB200_SCAN_REFERENCE).  functions.py:158-171 + codon classes of :196-215 +
// gc_frame_plot.py:7-74.
//
// A block walks a contiguous range of 2048-base tiles of the concatenated batch.  Per tile and per
// contig segment inside it (one for ordinary contigs, several for tiny ones):
//   A  the tile's letters + 64-base halos arrive in shared memory by a 1-D bulk TMA copy (cp.async.bulk, completion on
//      an mbarrier), issued one tile ahead into the other half of a double buffer; 137 threads turn 16 letters
//      each into base codes and GC / acgt bit planes (positions outside the segment's contig read as "no base")
//   B  6-mer -> Shine-Dalgarno motif-set masks for every position (tables in shared memory), restricted to the
//      motif classes short enough for the run of plain acgt letters at that position: this makes ambiguity
//      codes and the truncated windows at a contig's end exact without a per-letter fallback
//   C  one thread per 8 consecutive bases: the three GC-frame window sums per codon as popcounts of the
//      GC bit plane under a stride-3 mask (118-bit window in registers), codon class, factor class,
//      both RBS background scores from ORs of the motif-set masks; results packed into one meta word and
//      18 mask bytes per thread, RBS histogram in thread-private shared-memory counters (no atomics)
//   D  per-contig histogram / base-count accumulators, flushed with global atomics when the contig changes
// Output words are written once per tile, 8-byte meta stores and 4-byte mask stores, coalesced.
#pragma once
#include "pipeline.cuh"

#ifndef ST_T
#define ST_T 2048               /* bases per tile (256 threads x 8).  1920 (240 threads, one round of phase B) measured 5 % slower */
#endif
#define ST_NT 256
#define ST_HL 64
#define ST_NS (ST_T + 128)

struct ScanSmem {
    unsigned short tab_end[4096];
    unsigned short tab_start[4096];
    unsigned short em[ST_NS + 8];
    unsigned short sm[ST_NS + 8];
    u8 gt4[4][1024];              // best score of a motif-class set per offset group (rbs_gX_lo/hi combined)
    u8 code[ST_NS + 16];
    u32 gcw[ST_NS / 32 + 4];
    u32 okw[ST_NS / 32 + 4];
    u8 hpriv[28][ST_NT];          // thread-private RBS histogram counters (<= 16 per tile and thread)
    u8 maskb[18][ST_NT];
    u8 chlut[256];                // letter -> base code | GC flag << 3
    u8 cls_tab[64];
    u8 facm[32];                  // trit code -> factor index forward | reverse << 3 (= the meta byte)
    unsigned short lenmask[8];    // motif classes of length <= r
    u32 hist[28];
    u32 ngc, nat;
    u64* mptr[18];
    unsigned long long bar[2];    // mbarriers of the two letter buffers
    uint4 raw[2][ST_NS / 16];     // letters of [tg0-64, tg0+2048+64), double buffered
};

// ---- 1-D bulk TMA copy global -> shared with mbarrier completion (PTX ISA: cp.async.bulk, sm_90+)
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, u32 bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one thread: start the copy of a tile's letters into buffer `buf`.  The source range is clipped to [0, nb rounded up to
// 16) -- the library's input buffer is padded by 64 bytes -- and the bytes outside it are never read as bases (phase A
// masks by contig bounds).
__device__ __forceinline__ void scan_issue_tile(const Batch& B, ScanSmem& S, i64 tg0, int buf) {
    i64 g_lo = tg0 - ST_HL, g_hi = tg0 + ST_T + ST_HL;
    const i64 lim = (B.nb + 15) & ~(i64)15;
    if (g_lo < 0) g_lo = 0;
    if (g_hi > lim) g_hi = lim;
    const u32 bytes = (u32)(g_hi - g_lo);
    mbar_expect_tx(&S.bar[buf], bytes);
    tma_load_1d((char*)S.raw[buf] + (g_lo - (tg0 - ST_HL)), B.seq + g_lo, bytes, &S.bar[buf]);
}

__device__ __forceinline__ void scan_flush(const Batch& B, ScanSmem& S, int cur_c, int tid) {
    if (cur_c < 0) return;
    if (tid < 28) {
        const u32 v = S.hist[tid];
        if (v) atomicAdd(&B.cs[cur_c].hist_bg[tid], v);
        S.hist[tid] = 0;
    } else if (tid == 28) {
        if (S.ngc) atomicAdd(&B.cs[cur_c].nGC, S.ngc);
        if (S.nat) atomicAdd(&B.cs[cur_c].nAT, S.nat);
        S.ngc = 0;
        S.nat = 0;
    }
}

// the 16 raw letters of staging chunk `tid` (positions 16*tid .. 16*tid+15 of [tg0-64, tg0+2048+64)) of a tile
__device__ __forceinline__ uint4 scan_load_chunk(const Batch& B, i64 tg0, int tid, bool wide) {
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    const i64 g0 = tg0 - ST_HL + 16 * (i64)tid;
    if (tid >= ST_NS / 16 || g0 + 16 <= 0 || g0 >= B.nb) return raw;
    if (wide && g0 >= 0 && g0 + 16 <= B.nb) return *(const uint4*)(B.seq + g0);
    u32 w[4] = {0u, 0u, 0u, 0u};
    for (int b = 0; b < 16; b++) {
        const i64 g = g0 + b;
        if (g >= 0 && g < B.nb) w[b >> 2] |= (u32)B.seq[g] << (8 * (b & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// use_tma: the letters arrive by bulk TMA copies (the library's own input buffer) or by plain 16-byte loads (a
// caller-owned device buffer without alignment / padding guarantees); a template parameter, so that the path not taken
// costs no registers
template <bool use_tma>
__global__ void __launch_bounds__(ST_NT, 4) k_scan_tiles(const Batch B, i64 ntiles, int tiles_per_block) {
    extern __shared__ __align__(16) unsigned char scan_smem_raw[];
    ScanSmem& S = *reinterpret_cast<ScanSmem*>(scan_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4096; i += ST_NT) {
        S.tab_end[i] = d_rbs_end_mask[i];
        S.tab_start[i] = d_rbs_start_mask[i];
        const int g = i >> 10, m = i & 1023;
        int lo, hi;
        switch (g) {
            case 0: lo = d_rbs_g0_lo[m & 31]; hi = d_rbs_g0_hi[m >> 5]; break;
            case 1: lo = d_rbs_g1_lo[m & 31]; hi = d_rbs_g1_hi[m >> 5]; break;
            case 2: lo = d_rbs_g2_lo[m & 31]; hi = d_rbs_g2_hi[m >> 5]; break;
            default: lo = d_rbs_g3_lo[m & 31]; hi = d_rbs_g3_hi[m >> 5]; break;
        }
        S.gt4[g][m] = (u8)(lo > hi ? lo : hi);
    }
    {
        const u8 ch = lower((u8)tid);
        S.chlut[tid] = (u8)(base_code(ch) | (gc_flag(ch) << 3));
    }
    if (tid < 64) {
        S.cls_tab[tid] = B.P.codon_cls[tid];
        if (tid < 32) S.facm[tid] = (u8)(d_gc_fac_index[0][tid] | (d_gc_fac_index[1][tid] << 3));
        if (tid < 8) {
            u32 m = 0;
            for (int q = 0; q < PB_NMOTIF; q++)
                if ((int)d_rbs_motif[q][1] <= tid) m |= 1u << d_rbs_motif[q][0];
            S.lenmask[tid] = (unsigned short)m;
        }
    }
    for (int b = 0; b < 28; b++) S.hpriv[b][tid] = 0;
    if (tid < 28) S.hist[tid] = 0;
    if (tid == 0) {
        S.ngc = 0;
        S.nat = 0;
        S.mptr[0] = B.mS; S.mptr[1] = B.ms; S.mptr[2] = B.mT; S.mptr[3] = B.mt;
        S.mptr[4] = B.bA; S.mptr[5] = B.bC; S.mptr[6] = B.bG; S.mptr[7] = B.bT;
        for (int k = 0; k < 5; k++) {
            S.mptr[8 + k] = B.cF[k];
            S.mptr[13 + k] = B.cR[k];
        }
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int cur_c = -1;                                  // contig of the block's accumulators
    const i64 t0 = (i64)blockIdx.x * tiles_per_block;
    const i64 t1 = (t0 + tiles_per_block < ntiles) ? t0 + tiles_per_block : ntiles;
    const int pb = ST_HL + 8 * tid;                  // staging index of this thread's first base
    const bool wide = (((size_t)B.seq) & 15) == 0;   // 16-byte loads of the letters (tile starts are multiples of 2048)
    int c = contig_of(B, t0 * ST_T);                 // walking contig cursor with its bounds: tiles are consecutive
    i64 cb = B.coff[c], ce = B.coff[c + 1];
    uint4 raw_next = make_uint4(0u, 0u, 0u, 0u);
    if (use_tma) {
        if (tid == 0 && t0 < t1) scan_issue_tile(B, S, t0 * ST_T, 0);
    } else {
        raw_next = scan_load_chunk(B, t0 * ST_T, tid, wide);
    }
    for (i64 tile = t0; tile < t1; tile++) {
        const i64 tg0 = tile * ST_T;
        const i64 tg1 = (tg0 + ST_T < B.nb) ? tg0 + ST_T : B.nb;
        const i64 gb = tg0 + 8 * tid;                // this thread's bases: gb .. gb+7
        u64 acc_cls = 0, acc_cd = 0, acc_kf = 0, acc_kr = 0, acc_sf = 0, acc_sr = 0;
        uint4 raw = raw_next;
        if (use_tma) {
            const int it = (int)(tile - t0), buf = it & 1;
            // the other buffer was last read in the previous tile's phase A, which every thread left through barriers
            if (tid == 0 && tile + 1 < t1) scan_issue_tile(B, S, tg0 + ST_T, buf ^ 1);   // in flight during this tile's phases
            mbar_wait(&S.bar[buf], (u32)((it >> 1) & 1));
            if (tid < ST_NS / 16) raw = S.raw[buf][tid];
        } else if (tile + 1 < t1) {
            raw_next = scan_load_chunk(B, tg0 + ST_T, tid, wide);
        }
        while (ce <= tg0 && c + 1 < B.nc) {           // (skips empty contigs too)
            c++;
            cb = ce;
            ce = B.coff[c + 1];
        }
        for (;;) {
            const i64 seg_lo = cb > tg0 ? cb : tg0, seg_hi = ce < tg1 ? ce : tg1;
            if (seg_hi > seg_lo) {
            if (c != cur_c) {
                scan_flush(B, S, cur_c, tid);        // phase D of the previous segment ended with a barrier
                cur_c = c;
            }
            __syncthreads();
            // ---- A: codes and bit planes of [tg0-64, tg0+2048+64) from the raw letters, 16 positions per thread
            if (tid < ST_NS / 16 + 8) {
                const i64 g0 = tg0 - ST_HL + 16 * (i64)tid;
                const u32 rw[4] = {raw.x, raw.y, raw.z, raw.w};
                u32 cw[4] = {0u, 0u, 0u, 0u};
                u32 gcbits = 0, okbits = 0;
                bool bad = false;
#pragma unroll
                for (int b = 0; b < 16; b++) {
                    const i64 g = g0 + b;
                    u32 cd = 6;
                    if (tid < ST_NS / 16 && g >= cb && g < ce) {
                        const u32 v = S.chlut[(rw[b >> 2] >> (8 * (b & 3))) & 0xFFu];
                        cd = v & 7u;
                        gcbits |= (v >> 3) << b;
                        if (cd == 5) {
                            if (g >= seg_lo && g < seg_hi) bad = true;
                            cd = 4;
                        }
                        okbits |= (cd < 4 ? 1u : 0u) << b;
                    }
                    cw[b >> 2] |= cd << (8 * (b & 3));
                }
                if (tid <= ST_NS / 16) *(uint4*)(S.code + 16 * tid) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
                ((unsigned short*)S.gcw)[tid] = (unsigned short)gcbits;
                ((unsigned short*)S.okw)[tid] = (unsigned short)okbits;
                if (bad) atomicOr(&B.cs[c].err, (u32)ERR_CHAR);
            }
            __syncthreads();
            // ---- B: motif-set masks of the 6-mer starting at every staged position
            for (int grp = tid; grp < ST_NS / 8; grp += ST_NT) {
                const int p0 = grp * 8;
                const u64 c_lo = *(const u64*)(S.code + p0), c_hi = *(const u64*)(S.code + p0 + 8);
                const u32 okv = __funnelshift_r(S.okw[p0 >> 5], S.okw[(p0 >> 5) + 1], p0 & 31);
                u32 idx = 0;
                u32 e[8], s[8];
#pragma unroll
                for (int j = 0; j < 13; j++) {
                    const u32 cd = (u32)((j < 8 ? (c_lo >> (8 * j)) : (c_hi >> (8 * (j - 8)))) & 3u);
                    idx = ((idx << 2) | cd) & 4095u;
                    if (j >= 5) {
                        const u32 o6 = (okv >> (j - 5)) & 63u;                  // bit t: letter t of the 6-mer is plain acgt
                        const int runf = __ffs((int)(~o6)) - 1;                  // plain letters from its first base on (<= 6)
                        const int runb = __clz((int)(~(o6 << 26)));              // plain letters up to its last base (<= 6)
                        e[j - 5] = S.tab_end[idx] & S.lenmask[runb];
                        s[j - 5] = S.tab_start[idx] & S.lenmask[runf];
                    }
                }
                uint4 ev, sv;
                ev.x = e[0] | (e[1] << 16); ev.y = e[2] | (e[3] << 16); ev.z = e[4] | (e[5] << 16); ev.w = e[6] | (e[7] << 16);
                sv.x = s[0] | (s[1] << 16); sv.y = s[2] | (s[3] << 16); sv.z = s[4] | (s[5] << 16); sv.w = s[6] | (s[7] << 16);
                *(uint4*)(S.em + p0) = ev;
                *(uint4*)(S.sm + p0) = sv;
            }
            __syncthreads();
            // ---- C: this thread's 8 bases
            if (gb < seg_hi && gb + 8 > seg_lo) {
                const int L = (int)(ce - cb);
                int tz[10];
                {
                    const int b0 = pb - 57, w = b0 >> 5, sh = b0 & 31;
                    const u32 q0 = S.gcw[w], q1 = S.gcw[w + 1], q2 = S.gcw[w + 2], q3 = S.gcw[w + 3], q4 = S.gcw[w + 4];
                    const u32 v0 = __funnelshift_r(q0, q1, sh), v1 = __funnelshift_r(q1, q2, sh),
                              v2 = __funnelshift_r(q2, q3, sh), v3 = __funnelshift_r(q3, q4, sh);
#pragma unroll
                    for (int j = 0; j < 10; j++) {
                        const u32 a0 = __funnelshift_r(v0, v1, j), a1 = __funnelshift_r(v1, v2, j),
                                  a2 = __funnelshift_r(v2, v3, j), a3 = v3 >> j;
                        tz[j] = __popc(a0 & 0x49249249u) + __popc(a1 & 0x92492492u) + __popc(a2 & 0x24924924u) +
                                __popc(a3 & 0x00249249u);
                    }
                }
                const u64 c_lo = *(const u64*)(S.code + pb), c_hi = *(const u64*)(S.code + pb + 8);
                // three passes over the thread's 8 bases -- codon / GC-frame classes, forward RBS score, reverse RBS score -- so
                // that the window sums, the forward and the reverse motif-mask words are never live together (the kernel
                // runs at its 64-register cap)
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const i64 g = gb + k;
                    if (g < seg_lo || g >= seg_hi) continue;
                    const u32 cd0 = (u32)(c_lo >> (8 * k)) & 7u;
                    const u32 cd1 = (u32)((k + 1 < 8 ? (c_lo >> (8 * (k + 1))) : (c_hi >> (8 * (k + 1 - 8))))) & 7u;
                    const u32 cd2 = (u32)((k + 2 < 8 ? (c_lo >> (8 * (k + 2))) : (c_hi >> (8 * (k + 2 - 8))))) & 7u;
                    int cls = CLS_NONE;
                    if (cd0 < 4 && cd1 < 4 && cd2 < 4) cls = S.cls_tab[cd0 * 16 + cd1 * 4 + cd2];
                    const int tr = gc_trits(tz[k], tz[k + 1], tz[k + 2]);
                    const u32 mb = S.facm[tr];
                    const u32 kf = mb & 7u, kr = mb >> 3;
                    // one bit per base in byte `value` of a 64-bit word: byte v of accX = the 8-base mask of "X == v"
                    acc_cls |= 1ull << (cls * 8 + k);
                    acc_cd |= 1ull << (cd0 * 8 + k);
                    acc_kf |= 1ull << (kf * 8 + k);
                    acc_kr |= 1ull << (kr * 8 + k);
                }
                u32 e2[22];                           // e2[x] = em[x] | em[x+1]
                {
                    const uint4 ea = *(const uint4*)(S.em + pb), eb = *(const uint4*)(S.em + pb + 8), ec = *(const uint4*)(S.em + pb + 16);
                    const u32 ew[12] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w, ec.x, ec.y, ec.z, ec.w};
                    u32 ee[24];
#pragma unroll
                    for (int x = 0; x < 24; x++) ee[x] = (x & 1) ? (ew[x >> 1] >> 16) : (ew[x >> 1] & 0xFFFFu);
#pragma unroll
                    for (int x = 0; x < 22; x++) e2[x] = ee[x] | ee[x + 1];
                    // groups of three: em[k]|em[k+1]|em[k+2] = e2[k] | e2[k+1]
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const i64 g = gb + k;
                    if (g < seg_lo || g >= seg_hi) continue;
                    int sf;
                    const int i = (int)(g - cb);
                    if (i + 21 <= L) {
                        // forward strand, score_rbs(dna[i:i+21]): 6-mers at window offsets 5..10 (gm), 11..12 (gl), 3..4 (gh), 0..2 (gf)
                        const u32 gm = e2[k + 5] | e2[k + 7] | e2[k + 9], gl = e2[k + 11], gh = e2[k + 3], gf = e2[k] | e2[k + 1];
                        const int a = S.gt4[0][gm], b = S.gt4[1][gl], cc = S.gt4[2][gh], d = S.gt4[3][gf];
                        sf = a > b ? a : b;
                        sf = cc > sf ? cc : sf;
                        sf = d > sf ? d : sf;
                    } else {
                        // window truncated at the contig end (W < 21): the reversed window is scored, so offsets count from
                        // its last base; the motif whose reversed form ends a bases before it must also start inside it
                        const int W = L - i;
                        u32 gq[4] = {0u, 0u, 0u, 0u};
                        for (int a = 3; a <= 15 && a + 3 <= W; a++) {
                            const int room = W - a;
                            const u32 m = S.em[pb + k + W - 1 - a - 5] & S.lenmask[room < 6 ? room : 6];
                            gq[(a <= 4) ? 1 : (a <= 10) ? 0 : (a <= 12) ? 2 : 3] |= m;
                        }
                        const int a = S.gt4[0][gq[0]], b = S.gt4[1][gq[1]], cc = S.gt4[2][gq[2]], d = S.gt4[3][gq[3]];
                        sf = a > b ? a : b;
                        sf = cc > sf ? cc : sf;
                        sf = d > sf ? d : sf;
                    }
                    if (sf) S.hpriv[sf][tid]++;
                    acc_sf |= (u64)sf << (8 * k);
                }
                u32 s2[22];                           // s2[x] = sm[x] | sm[x+1]
                {
                    const uint4 sa = *(const uint4*)(S.sm + pb), sb = *(const uint4*)(S.sm + pb + 8), sc = *(const uint4*)(S.sm + pb + 16);
                    const u32 sw[12] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w, sc.x, sc.y, sc.z, sc.w};
                    u32 ss[24];
#pragma unroll
                    for (int x = 0; x < 24; x++) ss[x] = (x & 1) ? (sw[x >> 1] >> 16) : (sw[x >> 1] & 0xFFFFu);
#pragma unroll
                    for (int x = 0; x < 22; x++) s2[x] = ss[x] | ss[x + 1];
                    // sm[k+13..15] = s2[k+13] | s2[k+14]
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const i64 g = gb + k;
                    if (g < seg_lo || g >= seg_hi) continue;
                    // reverse strand, score_rbs(rev_comp(dna[i:i+21])): motifs start at window offsets 5..10 (gm),
                    // 3..4 (gl), 11..12 (gh), 13..15 (gf); a truncated window only cuts motifs at the contig end
                    const u32 rm = s2[k + 5] | s2[k + 7] | s2[k + 9], rl = s2[k + 3], rh = s2[k + 11], rf = s2[k + 13] | s2[k + 14];
                    const int a = S.gt4[0][rm], b = S.gt4[1][rl], cc = S.gt4[2][rh], d = S.gt4[3][rf];
                    int sr = a > b ? a : b;
                    sr = cc > sr ? cc : sr;
                    sr = d > sr ? d : sr;
                    if (sr) S.hpriv[sr][tid]++;
                    acc_sr |= (u64)sr << (8 * k);
                }
            }
            __syncthreads();
            // ---- D: fold the private counters into the block's per-contig accumulators
            for (int bin = 1 + warp; bin < 28; bin += ST_NT / 32) {
                u64* row = (u64*)S.hpriv[bin];
                const u64 x = row[lane];
                u32 sum = 0;
                if (x) {
                    row[lane] = 0;
                    sum = __dp4a((u32)x, 0x01010101u, 0u) + __dp4a((u32)(x >> 32), 0x01010101u, 0u);
                }
                sum = __reduce_add_sync(0xFFFFFFFFu, sum);
                if (lane == 0 && sum) S.hist[bin] += sum;
            }
            if (warp == ST_NT / 32 - 1) {
                u32 cnt = 0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int wi = ST_HL / 32 + lane + 32 * h;
                    const i64 gw = tg0 + 32 * (i64)(lane + 32 * h);
                    u32 m = 0xFFFFFFFFu;
                    if (gw < seg_lo) m &= (seg_lo - gw >= 32) ? 0u : (0xFFFFFFFFu << (int)(seg_lo - gw));
                    if (gw + 32 > seg_hi) m &= (seg_hi <= gw) ? 0u : (0xFFFFFFFFu >> (int)(gw + 32 - seg_hi));
                    cnt += __popc(S.gcw[wi] & m);
                }
                cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
                if (lane == 0) {
                    S.ngc += cnt;
                    S.nat += (u32)(seg_hi - seg_lo) - cnt;
                }
            }
            __syncthreads();
            }
            if (ce >= tg1 || c + 1 >= B.nc) break;    // the tile ends inside this contig (or it is the last one)
            c++;
            cb = ce;
            ce = B.coff[c + 1];
        }
        // ---- outputs of the tile
        if (8 * tid >= ST_T) {
            // (a thread beyond the tile's last base: nothing to store)
        } else if (gb + 8 <= B.nb) {
            *(u64*)(B.rbsf + gb) = acc_sf;
            *(u64*)(B.rbsr + gb) = acc_sr;
        } else {
            for (int k = 0; k < 8 && gb + k < B.nb; k++) {
                B.rbsf[gb + k] = (u8)(acc_sf >> (8 * k));
                B.rbsr[gb + k] = (u8)(acc_sr >> (8 * k));
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            S.maskb[q][tid] = (u8)(acc_cls >> (8 * (q + 1)));        // CLS_S, CLS_s, CLS_T, CLS_t = 1..4
            S.maskb[4 + q][tid] = (u8)(acc_cd >> (8 * q));           // a, c, g, t
        }
#pragma unroll
        for (int q = 0; q < 5; q++) {
            S.maskb[8 + q][tid] = (u8)(acc_kf >> (8 * q));
            S.maskb[13 + q][tid] = (u8)(acc_kr >> (8 * q));
        }
        __syncthreads();
        for (int idx = tid; idx < 18 * (ST_T / 32); idx += ST_NT) {
            const int m = idx / (ST_T / 32), wd = idx % (ST_T / 32);
            const i64 wi = tg0 / 32 + wd;
            if (wi * 32 < B.nb) ((u32*)S.mptr[m])[wi] = ((const u32*)S.maskb[m])[wd];
        }
        // (the barrier at the top of the next segment orders these reads against the next writes to maskb)
        __syncthreads();
    }
    __syncthreads();
    scan_flush(B, S, cur_c, tid);
}
