/* phanotate_b200 -- C ABI of the B200-native PHANOTATE hot path.
 *
 * One call processes a BATCH of contigs: six-frame ORF scan + RBS / start-codon / GC-frame
 * scoring, ORF/gap/overlap graph, exact shortest path, CDS call table.  This is what a
 * maintainer of the reference binds (ctypes) in place of the per-contig
 * Python calls in /root/reference/phanotate.py:40-76:
 *
 *   functions.get_orfs(locus)          phanotate_modules/functions.py:143-303   -> pb200_run + pb200_get_orfs
 *   functions.get_graph(orfs)          phanotate_modules/functions.py:307-454   -> pb200_build_edges + pb200_get_nodes/edges
 *   fz.empty_graph / add_edge / get_path   phanotate.py:56-64 (third-party fastpathz)
 *                                                                              -> inside pb200_run; pb200_bellman_ford for
 *                                                                                 an arbitrary edge list
 *   path -> CDS features               phanotate.py:65-76, locus.py:29-37       -> pb200_get_calls
 *
 * Plain pointers and sizes only.  Input buffers are caller-owned; result tables are owned by the
 * context and valid until the next pb200_run / pb200_destroy.  Every function returns 0 on
 * success or a negative code; pb200_last_error() gives the text.  A context is bound to one CUDA
 * device and one stream and must not be used from two threads at once (the reference is
 * single-threaded; its solver keeps module-global state, phanotate.py:56).
 */
#ifndef PHANOTATE_B200_H
#define PHANOTATE_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* decimal number sign * c * 10^e, c < 10^38 as four little-endian 32-bit limbs: the reference's
 * Decimal values (prec 28) cross the boundary digit for digit */
typedef struct pb200_dec {
    uint32_t c[4];
    int32_t e;
    int32_t neg;
} pb200_dec;

/* the five scoring parameters of file_handling.get_args (file_handling.py:46-66) */
typedef struct pb200_params {
    int32_t n_start;            /* <= 8 */
    char start_codon[8][4];     /* lower-case acgt, NUL padded */
    pb200_dec start_weight[8];  /* already divided by the max weight (file_handling.py:58-62) */
    int32_t n_stop;             /* <= 8 */
    char stop_codon[8][4];
    int32_t min_orf_len;        /* -l/--minlen, default 90; must be >= 9 */
    int32_t reserved;
} pb200_params;

typedef struct pb200_call {     /* one CDS row of Locus.tabular (locus.py:39-56) */
    int32_t contig;
    int32_t left;               /* entry node position              (phanotate.py:72-74) */
    int32_t right;              /* exit node position + 2           (locus.py:30)        */
    int32_t strand;             /* +1 / -1 = sign(left.frame)                             */
    pb200_dec weight;           /* graph.weight(Edge(left,right)) = Orf.weight (see PB200_CALL_WEIGHTS) */
    double score;               /* float(weight), what '%E' prints  (phanotate.py:75-76) */
} pb200_call;

typedef struct pb200_call24 {   /* the same row without the Decimal weight: the columns Locus.tabular prints (24 bytes instead of 48 over PCIe / NVLink) */
    int32_t contig, left, right, strand;
    double score;
} pb200_call24;

typedef struct pb200_orf {      /* Orf (orfs.py:71-95) */
    int32_t contig, start, stop, frame;   /* start/stop = leftmost base of the codon, frame = +-1..3 */
    int32_t rbs_score;
    int32_t trigger;            /* scan position at which the reference emits the ORF (insertion order) */
    int32_t start_weight;       /* index into start_codon[] or -1 */
    int32_t node;               /* start node id */
    pb200_dec pstop, weight;
} pb200_orf;

typedef struct pb200_node {     /* Node (nodes.py:2-21); ids are global over the batch, sorted by contig, position */
    int32_t contig, position;
    int32_t kind;               /* 0 (start,+f) entry, 1 (stop,+f) exit, 2 (stop,-f) entry, 3 (start,-f) exit */
    int32_t frame;              /* 1..3 */
    int32_t mate;               /* start node: its stop-key node; stop-key node: farthest start node */
    int32_t orf;                /* start node: its ORF; stop-key node: the family's longest ORF */
    int32_t other_end;          /* Orfs.other_end[position] (orfs.py:22-30) */
    int32_t trigger;
} pb200_node;

enum { PB200_EDGE_ORF = 0, PB200_EDGE_GAP = 1, PB200_EDGE_OVERLAP = 2, PB200_EDGE_BRIDGE = 3,
       PB200_EDGE_SOURCE = 4, PB200_EDGE_TARGET = 5, PB200_EDGE_TRNA = 6 /* the -20 edge of a tRNA hit */ };
#define PB200_NODE_SOURCE (-2)
#define PB200_NODE_TARGET (-3)
typedef struct pb200_edge {     /* Edge (edges.py:3-23) */
    int32_t contig, src, dst, kind;
    pb200_dec weight;
} pb200_edge;

typedef struct pb200_contig {
    int32_t length;
    uint32_t err;               /* PB200_ERR_* bits */
    int32_t node_off, n_nodes, orf_off, n_orfs, call_off, n_calls;
    int32_t n_ties;             /* relaxations that found an equal distance (tie-break diagnostics) */
    int32_t wide;               /* 1: some edge weight exceeds 110 bits and the solve ran on 256-bit distances */
    pb200_dec pstop;            /* contig-level P(stop) = pgap (functions.py:178,309) */
    pb200_dec pos_max[4], pos_min[4];   /* GC-frame exponents (functions.py:281-284) */
    double background_rbs[28], training_rbs[28];
} pb200_contig;

enum {
    PB200_ERR_CHAR = 1,      /* letter outside the IUPAC alphabet: the reference raises KeyError */
    PB200_ERR_RANGE = 2,
    PB200_ERR_PARALLEL = 4,  /* the reference raises ValueError("parallel edges are forbidden") */
    PB200_ERR_OVERFLOW = 8,
    PB200_ERR_NOPATH = 16,
    PB200_ERR_INTERNAL = 32,
    PB200_ERR_LOOKUP = 64,
    PB200_ERR_TIES = 128        /* exact ties in the solve that could not be settled in the reference's edge order */
};

enum {
    PB200_INPUT_DEVICE = 1,   /* flags of pb200_run: bases/offsets are device pointers */
    PB200_REUSE_INPUT = 2,    /* the batch uploaded by the previous pb200_run is still resident: skip the copy */
    PB200_INPUT_PACKED4 = 256, /* `bases` holds 4-bit letters, two per byte (pb200_pack4): half the bytes over the host link;
                                 the library expands them on the device.  Results are those of the lower-cased letters. */
    PB200_CALL_WEIGHTS = 16,  /* also fill pb200_call.weight with the 28-digit Decimal of every call (otherwise it is
                                 filled only where the Decimal chain ran anyway and is 0 elsewhere; pb200_call.score,
                                 the float that Locus.tabular prints with '%E', is always exact) */
    PB200_SOLVE_WIDE = 32,    /* testing: 256-bit distances in the solve for every contig */
    PB200_SOLVE_PLAIN = 64,   /* testing: the plain statement of the 128-bit sweep (every operand fetched when needed) instead
                                 of the windowed kernel */
    PB200_SOLVE_NOCHUNK = 128, /* testing: long contigs are solved by one sweep like the others, not in chunks */
    PB200_SCAN_REFERENCE = 8, /* testing: run the per-strip statement of the scan stage instead of the tiled kernel, and the
                                 per-candidate statement of the node marks instead of the 64-positions-at-a-time form */
    PB200_LITERAL = 4         /* replay the reference's Decimal arithmetic for EVERY ORF and overlap edge inside
                                 pb200_run.  Default: the solve uses certified integer weights (exactly
                                 trunc(weight*1000), edges.py:22) and the 28-digit Decimal weights are computed
                                 for the called CDS, and for everything else when pb200_get_orfs /
                                 pb200_build_edges ask for them.  Results are identical either way. */
};

typedef struct pb200_ctx pb200_ctx;

int pb200_create(int device, pb200_ctx** out);
void pb200_destroy(pb200_ctx* ctx);
const char* pb200_last_error(pb200_ctx* ctx);

/* bases: concatenated contig letters (any case, 15 IUPAC codes); offsets[n_contigs+1]. */
int pb200_run(pb200_ctx* ctx, const uint8_t* bases, const int64_t* offsets, int32_t n_contigs,
              const pb200_params* params, uint32_t flags);

/* host->device copy of a batch without running it; run it with pb200_run(ctx, bases, offsets, n, params,
 * PB200_REUSE_INPUT).  With several contexts this lets the copies go one after the other while other contexts compute. */
int pb200_upload(pb200_ctx* ctx, const uint8_t* bases, const int64_t* offsets, int32_t n_contigs);

/* tRNA masking (functions.py:457-509 add_trnas + the tRNA branch of the connect loop, functions.py:388-399).  The
 * reference runs aragorn / tRNAscan-SE itself; here the caller runs them (or anything else) and hands over the hit list
 * for the following runs: hit k lies on contig contig[k] (index inside the batch, ascending) from start[k] to stop[k],
 * 1-based, start > stop on the reverse strand -- the [start, stop] pairs add_trnas collects, in its order.  Every hit
 * becomes a node pair (gene 'tRNA', frame +-4) joined by an edge of weight -20 and connected by 'same'-scored gap edges
 * to every exit / entry node within 500 bp.  A tRNA on the shortest path comes back as a call row with strand +-2 and
 * score -20.  pb200_get_nodes returns the tRNA nodes behind the regular ones (orf = -1 - k).  n = 0 clears the list. */
int pb200_set_trnas(pb200_ctx* ctx, const int32_t* contig, const int32_t* start, const int32_t* stop, int32_t n);

/* 4-bit letters.  pb200_pack4: letters -> codes (a c g t n r y s w k m b v d h, either case; 15 = any other byte, flagged by
 * the run like the letter itself), two per byte, low nibble first, into out[(n + 1) / 2]; host threads, no context.
 * pb200_run(..., PB200_INPUT_PACKED4) takes such a buffer for `bases` (offsets stay in bases);
 * pb200_upload_packed4 = pb200_upload for it, `skip` (0 / 1) = the nibble of packed[0] the batch starts at (a group of
 * contigs cut out of a larger packed batch may start in the middle of a byte). */
int64_t pb200_pack4(const uint8_t* bases, int64_t n, uint8_t* out);
int pb200_upload_packed4(pb200_ctx* ctx, const uint8_t* packed, int32_t skip, const int64_t* offsets, int32_t n_contigs);
/* pb200_upload / pb200_upload_packed4 without blocking the host: the copy is queued on `via`'s copy stream -- contexts that
 * pass the same `via` get their copies one after the other in call order -- and ctx's stream waits for it on the device.
 * skip < 0: one byte per base; 0 / 1: 4-bit letters.  The host buffers stay untouched until the run has finished. */
int pb200_upload_async(pb200_ctx* ctx, pb200_ctx* via, const uint8_t* data, int32_t skip, const int64_t* offsets, int32_t n_contigs);
/* Double buffering across batches: the NEXT batch's 4-bit letters are copied in while the current run is going (the packed
 * buffer is free once the current batch has been expanded); the next pb200_upload_async with the same data / skip / size
 * then sends only the offsets.  Same `via` as the uploads; `data` stays untouched until that run. */
int pb200_prefetch_async(pb200_ctx* ctx, pb200_ctx* via, const uint8_t* data, int32_t skip, int64_t n_bases);

/* number added to the contig column of the call rows of the following runs (default 0): for a caller that cuts one
 * batch into groups for several contexts and wants the rows numbered in the whole batch */
int pb200_set_contig_base(pb200_ctx* ctx, int32_t base);

/* Long contigs (BASELINE config 5: one 10-Mb contig; functions.py:360-438 + phanotate.py:56-64 on a graph of 3.6e5
 * nodes): a contig with more than `long_nodes` graph nodes (~28 bp per node; default 4096) is solved as chunks of `core`
 * nodes (256), one warp each, swept from `warm` nodes upstream (768) to `margin` nodes downstream (64); the chunks'
 * distances are put together and EVERY node's Bellman equation is checked, so the result is the exact solve whatever
 * the geometry (a contig that fails the check is tried once more with four times the warm-up -- only while the library
 * picks the geometry itself -- and then solved again by one sweep).  The geometry only moves the time. */
int pb200_set_chunking(pb200_ctx* ctx, int32_t core, int32_t warm, int32_t margin, int32_t long_nodes);

/* out[0..7] = n_contigs, n_bases, n_nodes, n_orfs, n_overlap_edges, n_bridge_edges, n_calls, n_edges */
int pb200_sizes(pb200_ctx* ctx, int64_t out[8]);
/* out[0] = ORFs whose weight went through the literal Decimal chain before the solve, out[1] = after
 * it (called CDS), out[2] = overlap edges through the literal power, out[3] = chunks the long contigs were solved
 * in, out[4] = long contigs whose chunked solve failed its check and was redone by one sweep, out[5] = tRNA hits, out[6] = ORF weights beyond 256 bits (their contigs solved
 * with 2048-bit distances), out[7] = 1 if some long contig needed the second attempt of the chunked solve (four times the
 * warm-up) */
int pb200_stats(pb200_ctx* ctx, int64_t out[8]);
/* the integer weight the solver used for every ORF edge: trunc(Orf.weight * 1000) (edges.py:22) as
 * 8 little-endian 32-bit limbs, two's complement, per ORF */
int pb200_get_orf_int_weights(pb200_ctx* ctx, uint32_t* out);
/* the same for every overlap edge (n_overlap_edges entries, in edge order): trunc(score_overlap * 1000),
 * INT64_MAX where it needs more than 62 bits (the solver then uses the wide value) */
int pb200_get_overlap_int_weights(pb200_ctx* ctx, int64_t* out);
/* Orf.hold (orfs.py:84; functions.py:286-298: the product over the ORF's codons of ((1-pstop)**pos_max[i])**pos_min[j],
 * before Orf.score() -- orfs.py:122-127 -- inverts it) for every ORF, in pb200_get_orfs order.  Needs a PB200_LITERAL run:
 * a certified run never forms the product. */
int pb200_get_orf_holds(pb200_ctx* ctx, pb200_dec* out);

/* trunc(score_gap(len, 'same'|'diff') * 1000) for len = -2..300 per contig: 303 entries per contig each */
int pb200_get_gap_int_weights(pb200_ctx* ctx, int64_t* same, int64_t* diff);
int pb200_get_calls(pb200_ctx* ctx, pb200_call* out);
int pb200_get_calls24(pb200_ctx* ctx, pb200_call24* out);   /* the same rows, compact */
int pb200_get_contigs(pb200_ctx* ctx, pb200_contig* out);
int pb200_get_orfs(pb200_ctx* ctx, pb200_orf* out);
int pb200_get_nodes(pb200_ctx* ctx, pb200_node* out);
/* materialise every edge of get_graph (gap edges included) for the last batch */
int pb200_build_edges(pb200_ctx* ctx);
int pb200_get_edges(pb200_ctx* ctx, pb200_edge* out);

/* fastpathz-compatible solve of an arbitrary graph: exact integers (8 little-endian 32-bit limbs,
 * two's complement, per edge), edges relaxed in the given order with strict '<' until a pass
 * changes nothing.  path_out receives node ids source..target; *path_len = 0 if unreachable. */
int pb200_bellman_ford(pb200_ctx* ctx, int32_t n_nodes, int32_t n_edges, const int32_t* src,
                       const int32_t* dst, const uint32_t* weight_limbs, int32_t source, int32_t target,
                       int32_t* path_out, int32_t* path_len);

/* The join of the reference's C extension (src/phanotate_connect.c:78-121 `get_connected`, fed by `add_edge` :62-76;
 * the extension is built by setup.py:6-21 but phanotate.py never imports it).  Edge i = (left[i], right[i]) in add_edge
 * order.  *n_rows = number of rows; when out != NULL and cap_rows >= *n_rows, out receives the rows as int32 pairs
 * (right_i, left_j) in the reference's order (right entries outermost, left entries innermost, both in insertion order);
 * the reference's third tuple member is the constant 0.  Row condition: |right_i - left_j| <= 300, right_i != right_j,
 * left_i != left_j (:108-110; `min_distance` is ignored by the reference, :84-92).  Host pointers. */
int pb200_connect(pb200_ctx* ctx, const int32_t* left, const int32_t* right, int32_t n, int32_t* out, int64_t cap_rows,
                  int64_t* n_rows);

/* Host-side text ingest / output for whole batches (no device work): multi-record FASTA -> the packed batch of
 * pb200_run, and the tabular text of a batch (Locus.tabular, locus.py:39-56).
 * pb200_fasta_count: number of records ('>' at a line start).  pb200_fasta_parse: bases must hold n bytes, offsets
 * max_records+1 entries; name_begin/name_end are byte ranges of the first word of every header inside data; returns
 * the number of records or -1.  pb200_format_tabular: names = concatenated record names, name_off[n_contigs+1];
 * returns the number of bytes written, or -(bytes needed) when cap is too small. */
int64_t pb200_fasta_count(const char* data, int64_t n);
int64_t pb200_fasta_parse(const char* data, int64_t n, uint8_t* bases, int64_t* offsets, int64_t* name_begin,
                          int64_t* name_end, int64_t max_records);
int64_t pb200_format_tabular(const pb200_call* calls, const pb200_contig* contigs, int32_t n_contigs, const char* names,
                             const int64_t* name_off, char* out, int64_t cap);

/* device time of the stages of the last pb200_run in milliseconds (CUDA events on the context's
 * stream); names[i] are static strings.  Returns the number of stages written (<= cap). */
int pb200_stage_times(pb200_ctx* ctx, const char** names, float* ms, int cap);
/* device time between the timed stages of the last pb200_run (untimed helpers, host round trips, launch latency):
 * ms[i] = gap in front of stage names[i] */
int pb200_stage_gaps(pb200_ctx* ctx, const char** names, float* ms, int cap);
/* number of kernels launched by the last pb200_run */
/* "%E" of a score exactly as printf (and Python's '%E' % weight, phanotate.py:75-76) rounds it; what pb200_format_tabular
 * writes per row.  out: at least 32 bytes.  Returns the length. */
int pb200_format_score(double x, char* out);
int pb200_launch_count(pb200_ctx* ctx);
/* stopwatch on the context's stream: pb200_mark records event k (0..3) behind everything queued so far; pb200_elapsed_ms
 * waits for mark b and returns the device time between marks a and b (a region of several runs and gathers) */
int pb200_mark(pb200_ctx* ctx, int32_t k);
float pb200_elapsed_ms(pb200_ctx* ctx, int32_t a, int32_t b);
/* ... between mark a of one context and mark b of another context on the same device */
float pb200_elapsed_between_ms(pb200_ctx* from, int32_t a, pb200_ctx* to, int32_t b);

/* ---- multi-GPU: one process per GPU, contigs sharded over the ranks (they are independent:
 * phanotate.py:40-56 is a loop over loci), and ONE collective -- the gather of the ranks' call tables to rank 0.  Raw NCCL
 * bound at run time (dlopen libnccl.so.2, or the path in PB200_NCCL_LIB); nothing else of the library needs NCCL. */
/* rank 0: a fresh NCCL unique id, to be handed to every rank out of band (file, socket, environment) */
int pb200_comm_unique_id(uint8_t out[128]);
/* every rank: join the communicator on the context's device */
int pb200_comm_init(pb200_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t world);
int pb200_comm_destroy(pb200_ctx* ctx);
/* gather of call rows to rank 0 on the context's stream.  parts[i] = DEVICE pointer to part_rows[i] pb200_call records
 * (several contexts of one process: pass their pb200_device_calls); nparts = 0: the context's own table of its last run.
 * counts_out[world] (host, every rank) = rows per rank; rank 0: *rows_dev = device pointer to all rows in rank order
 * (valid until the next gather), *total = their number.  Exact row counts travel (an all-gather of one int64 per rank,
 * then one grouped send/receive), no padding. */
int pb200_comm_gather_calls(pb200_ctx* ctx, const void* const* parts, const int64_t* part_rows, int32_t nparts,
                            int64_t* counts_out, const pb200_call** rows_dev, int64_t* total);
/* the same gather with the rows converted to pb200_call24 on the way (parts still point to pb200_call records): half the
 * bytes over NVLink and over rank 0's host link; the fetches below then deliver pb200_call24 rows */
int pb200_comm_gather_calls24(pb200_ctx* ctx, const void* const* parts, const int64_t* part_rows, int32_t nparts,
                              int64_t* counts_out, const pb200_call24** rows_dev, int64_t* total);
/* rank 0: rows [first, first+n) of the last gather -> host (pb200_call or pb200_call24 records, as gathered) */
int pb200_comm_fetch_gathered(pb200_ctx* ctx, int64_t first, int64_t n, void* out);
/* rank 0: the same copy on its own stream, beside whatever the context runs next: _begin returns at once (page-locked
 * `out`), _wait blocks until the rows are there; the next gather waits for a pending copy by itself */
int pb200_comm_fetch_begin(pb200_ctx* ctx, int64_t first, int64_t n, void* out);
int pb200_comm_fetch_wait(pb200_ctx* ctx);
/* in-place reduction of n <= 64 doubles over the ranks (op 0 sum, 1 max); barrier */
int pb200_comm_allreduce(pb200_ctx* ctx, double* vals, int32_t n, int32_t op);
int pb200_comm_barrier(pb200_ctx* ctx);
/* NCCL version code (e.g. 22703) or -1 when NCCL is not available */
int pb200_comm_nccl_version(void);
/* device time of the whole last pb200_run (events at its first and last operation on the stream) */
float pb200_last_run_ms(pb200_ctx* ctx);
/* device address of the call table of the last run (n_calls rows), for zero-copy hand-off to a
 * collective (the multi-GPU gather of call tables) */
const pb200_call* pb200_device_calls(pb200_ctx* ctx);
/* page-lock / unlock a caller-owned host buffer so that pb200_run's copies are asynchronous DMA */
int pb200_pin_host(void* ptr, size_t bytes);
int pb200_unpin_host(void* ptr);
/* sizeof() of dec, params, call, orf, node, edge, contig -- lets a binding verify its struct layout */
int pb200_struct_sizes(int32_t out[8]);

#ifdef __cplusplus
}
#endif
#endif
