"""Batch engine: one call = the whole hot path for a batch of contigs on one B200.

    eng = Engine()                       # opens cuda:0 through the C ABI; raises without a GPU
    res = eng.run([b"acgt...", ...])     # or eng.run_packed(bases_u8, offsets_i64)
    res.calls                            # numpy structured array, one row per CDS
    res.call_rows(0)                     # [(left, right, '+', '-4.827981E+02'), ...] like Locus.tabular

Everything numeric happens in the CUDA library; this module only packs inputs, unpacks the
result tables and formats text.
"""
from __future__ import annotations

import ctypes
from decimal import Decimal
from typing import Iterable, Sequence

import numpy as np

from . import _native as N

DEFAULT_START_CODONS = "atg:0.85,gtg:0.10,ttg:0.05"     # file_handling.py:50
DEFAULT_STOP_CODONS = "tag,tga,taa"                      # file_handling.py:51
DEFAULT_MIN_ORF_LEN = 90                                 # file_handling.py:52


def parse_start_codons(text: str) -> dict:
    """file_handling.py:57-62: 'atg:0.85,...' -> {codon: Decimal weight / max weight}."""
    w = {}
    for item in text.split(","):
        codon, weight = item.split(":")
        w[codon.lower()] = Decimal(weight)
    m = max(w.values())
    return {k: v / m for k, v in w.items()}


def parse_stop_codons(text: str) -> list:
    """file_handling.py:63-66"""
    return [c.lower() for c in text.split(",")]


class PhanotateError(RuntimeError):
    pass


def make_params(start_codons=None, stop_codons=None, min_orf_len=DEFAULT_MIN_ORF_LEN) -> np.ndarray:
    if start_codons is None:
        start_codons = parse_start_codons(DEFAULT_START_CODONS)
    if stop_codons is None:
        stop_codons = parse_stop_codons(DEFAULT_STOP_CODONS)
    if len(start_codons) > 8 or len(stop_codons) > 8:
        raise ValueError("at most 8 start and 8 stop codons are supported")
    p = np.zeros(1, dtype=N.PARAMS)
    p["n_start"] = len(start_codons)
    for k, (codon, w) in enumerate(start_codons.items()):
        p["start_codon"][0][k] = codon.encode()
        c, e, s = N.decimal_to_dec(Decimal(w))
        p["start_weight"][0][k]["c"] = c
        p["start_weight"][0][k]["e"] = e
        p["start_weight"][0][k]["neg"] = s
    p["n_stop"] = len(stop_codons)
    for k, codon in enumerate(stop_codons):
        p["stop_codon"][0][k] = codon.encode()
    p["min_orf_len"] = int(min_orf_len)
    return p


class Result:
    """Tables of one batch (copied out of the context, so they outlive the next run)."""

    def __init__(self, engine, names=None):
        self._e = engine
        self._gen = engine._generation            # the context's tables belong to this run only until the next one starts
        self.names = names
        sz = np.zeros(8, dtype=np.int64)
        engine._ck(engine.lib.pb200_sizes(engine.ctx, sz.ctypes.data))
        (self.n_contigs, self.n_bases, self.n_nodes, self.n_orfs, self.n_overlaps, self.n_bridges,
         self.n_calls, _) = (int(v) for v in sz)
        self.calls = np.zeros(self.n_calls, dtype=N.CALL)
        engine._ck(engine.lib.pb200_get_calls(engine.ctx, self.calls.ctypes.data))
        self.contigs = np.zeros(self.n_contigs, dtype=N.CONTIG)
        engine._ck(engine.lib.pb200_get_contigs(engine.ctx, self.contigs.ctypes.data))
        self._orfs = self._nodes = self._edges = None
        st = np.zeros(8, dtype=np.int64)
        engine._ck(engine.lib.pb200_stats(engine.ctx, st.ctypes.data))
        self.n_literal_presolve, self.n_literal_postsolve, self.n_literal_overlaps = (int(v) for v in st[:3])
        self.n_chunks, self.n_chunk_fallbacks, self.n_trnas = int(st[3]), int(st[4]), int(st[5])
        self.n_huge_weights, self.chunk_second_attempt = int(st[6]), bool(st[7])
        self.launches = int(engine.lib.pb200_launch_count(engine.ctx))
        self.stage_ms = engine._stage_times()

    # lazily fetched tables -- only valid to request before the engine runs the next batch
    def _live(self, what):
        if self._gen != self._e._generation:
            raise PhanotateError("Result.%s requested after the engine ran another batch: the context holds that batch's "
                                 "tables now (fetch_all() before the next run keeps them)" % what)

    @property
    def orfs(self):
        if self._orfs is None:
            self._live("orfs")
            self._orfs = np.zeros(self.n_orfs, dtype=N.ORF)
            self._e._ck(self._e.lib.pb200_get_orfs(self._e.ctx, self._orfs.ctypes.data))
        return self._orfs

    @property
    def nodes(self):
        if self._nodes is None:
            self._live("nodes")
            self._nodes = np.zeros(self.n_nodes + 2 * self.n_trnas, dtype=N.NODE)     # tRNA node pairs follow the regular nodes
            self._e._ck(self._e.lib.pb200_get_nodes(self._e.ctx, self._nodes.ctypes.data))
        return self._nodes

    @property
    def edges(self):
        if self._edges is None:
            self._live("edges")
            self._e._ck(self._e.lib.pb200_build_edges(self._e.ctx))
            sz = np.zeros(8, dtype=np.int64)
            self._e._ck(self._e.lib.pb200_sizes(self._e.ctx, sz.ctypes.data))
            self._edges = np.zeros(int(sz[7]), dtype=N.EDGE)
            self._e._ck(self._e.lib.pb200_get_edges(self._e.ctx, self._edges.ctypes.data))
        return self._edges

    def orf_holds(self):
        """Orf.hold per ORF (Decimal) -- only after a literal=True run."""
        self._live("orf_holds")
        raw = np.zeros(self.n_orfs, dtype=N.DEC)
        self._e._ck(self._e.lib.pb200_get_orf_holds(self._e.ctx, raw.ctypes.data))
        return [N.dec_to_decimal(r) for r in raw]

    def orf_int_weights(self):
        """trunc(Orf.weight*1000) per ORF as Python ints -- what the solver used (request before .orfs)."""
        self._live("orf_int_weights")
        raw = np.zeros((self.n_orfs, 8), dtype=np.uint32)
        self._e._ck(self._e.lib.pb200_get_orf_int_weights(self._e.ctx, raw.ctypes.data))
        out = []
        for i, row in enumerate(raw):
            if int(row[7]) == 0x7FFFFFFE:        # marker: beyond 256 bits, the solve formed it from the Decimal weight
                w = N.dec_to_decimal(self.orfs[i]["weight"]) * 1000
                out.append(int(w))               # (truncation toward zero, like fastpathz: phanotate.py:55)
                continue
            v = 0
            for k in range(7, -1, -1):
                v = (v << 32) | int(row[k])
            out.append(v - (1 << 256) if v >> 255 else v)
        return out

    def overlap_int_weights(self):
        """trunc(score_overlap*1000) per overlap edge (int64; INT64_MAX marks a wider value)."""
        self._live("overlap_int_weights")
        out = np.zeros(self.n_overlaps, dtype=np.int64)
        self._e._ck(self._e.lib.pb200_get_overlap_int_weights(self._e.ctx, out.ctypes.data))
        return out

    def gap_int_weights(self):
        """(same, diff): trunc(score_gap(len)*1000) for len = -2..300, one row of 303 per contig."""
        self._live("gap_int_weights")
        same = np.zeros((self.n_contigs, 303), dtype=np.int64)
        diff = np.zeros((self.n_contigs, 303), dtype=np.int64)
        self._e._ck(self._e.lib.pb200_get_gap_int_weights(self._e.ctx, same.ctypes.data, diff.ctypes.data))
        return same, diff

    def fetch_all(self):
        self.orfs, self.nodes, self.edges
        return self

    @staticmethod
    def fatal(err: int) -> bool:
        """Does this contig's error word stop its output?  (ERR_NOPATH alone does not: no source->target path = no calls.)"""
        return bool(int(err) & ~N.ERR_NOPATH)

    def check(self, contig: int | None = None):
        """Raise what the reference would have raised for a contig (or for any contig): KeyError for a letter outside
        the IUPAC alphabet (functions.py:20-24), ValueError for parallel edges (graphs.py:73-74) and for the Orfs.get_orf
        lookup (orfs.py:62-69).  Anything else this implementation cannot finish (edge weights beyond 2048 bits, more than 16,384
        exact ties in one contig; DESIGN.md: known limits) raises PhanotateError."""
        cs = self.contigs if contig is None else self.contigs[contig:contig + 1]
        for i, c in enumerate(cs):
            err = int(c["err"])
            if not err:
                continue
            k = i if contig is None else contig
            if err & N.ERR_CHAR:
                raise KeyError("contig %d: letter outside the IUPAC alphabet (functions.rev_comp)" % k)
            if err & N.ERR_PARALLEL:
                raise ValueError("parallel edges are forbidden")
            if err & N.ERR_LOOKUP:
                raise ValueError("contig %d: orf not found (Orfs.get_orf)" % k)
            if err & (N.ERR_OVERFLOW | N.ERR_RANGE):
                raise PhanotateError("contig %d: an edge weight is beyond this build's 2048-bit exact range (~1e560; an ORF "
                                     "of tens of kb of pure A/T) -- the reference has no such limit "
                                     "(device error bits 0x%x)" % (k, err))
            if err & N.ERR_NOPATH and not (err & ~N.ERR_NOPATH):
                continue                        # no source->target path: no calls (undefined in the reference)
            if err & N.ERR_TIES:
                raise PhanotateError("contig %d: more exact ties in the shortest path than this build settles in the "
                                     "reference's edge order (16,384 per contig; an exact repeat over megabases) -- no calls rather "
                                     "than calls that might differ (device error bits 0x%x)" % (k, err))
            raise PhanotateError("contig %d: device error bits 0x%x" % (k, err))

    def call_rows(self, contig: int):
        """[(left, right, strand char, '%E' score)] in path order, the columns Locus.tabular prints."""
        c = self.contigs[contig]
        rows = []
        for r in self.calls[c["call_off"]:c["call_off"] + c["n_calls"]]:
            rows.append((int(r["left"]), int(r["right"]), "+" if r["strand"] > 0 else "-", "%E" % float(r["score"])))
        return rows

    def call_genes(self, contig: int):
        """'CDS' / 'tRNA' per call row (a tRNA hit on the path: strand column +-2)"""
        c = self.contigs[contig]
        return ["tRNA" if abs(int(s)) == 2 else "CDS" for s in self.calls["strand"][c["call_off"]:c["call_off"] + c["n_calls"]]]


class Engine:
    def __init__(self, device: int = 0, lib_path: str | None = None):
        self.lib = N.load(lib_path)
        ctx = ctypes.c_void_p()
        rc = self.lib.pb200_create(int(device), ctypes.byref(ctx))
        if rc != 0 or not ctx.value:
            raise RuntimeError("phanotate_b200: cannot open CUDA device %d (no CPU fallback exists)" % device)
        self.ctx = ctx
        self.device = device
        self._generation = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pb200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PhanotateError(self.lib.pb200_last_error(self.ctx).decode() or "error %d" % rc)

    def _stage_times(self):
        names = (ctypes.c_char_p * 64)()
        ms = (ctypes.c_float * 64)()
        n = self.lib.pb200_stage_times(self.ctx, names, ms, 64)
        out = {}
        for i in range(max(n, 0)):
            k = names[i].decode()
            out[k] = out.get(k, 0.0) + float(ms[i])
        return out

    def set_trnas(self, trnas=None):
        """tRNA hits for the following runs (functions.py:457-509): [(contig index, start, stop)], start > stop on the
        reverse strand, as add_trnas collects them; sorted by contig here (order inside a contig kept).  None / [] clears."""
        rows = sorted(trnas or [], key=lambda r: r[0])
        arr = np.asarray(rows, dtype=np.int32).reshape(-1, 3)
        c, a, b = (np.ascontiguousarray(arr[:, k]) for k in range(3))
        self._ck(self.lib.pb200_set_trnas(self.ctx, c.ctypes.data, a.ctypes.data, b.ctypes.data, len(arr)))

    def pack4(self, bases):
        """letters (uint8) -> 4-bit codes, two per byte (pb200_pack4): the input format that costs half the host link.
        What run_packed(..., packed4=True) takes; a FASTA reader would emit it directly."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        out = np.zeros((len(bases) + 1) // 2, dtype=np.uint8)
        if self.lib.pb200_pack4(bases.ctypes.data, len(bases), out.ctypes.data) < 0:
            raise PhanotateError("pb200_pack4 failed")
        return out

    def run_packed(self, bases, offsets, params=None, names=None, resident=False, fetch=True, literal=False, flags=0,
                   call_weights=False, packed4=False):
        """bases: uint8 array of concatenated contigs (packed4=True: their 4-bit codes from pack4), offsets: int64[n+1].

        resident=True reuses the batch the previous call uploaded (inputs already in HBM).
        fetch=False skips copying the result tables to the host (returns None).
        literal=True replays the reference's Decimal arithmetic for every ORF and overlap edge inside the
        run (PB200_LITERAL); by default the solve runs on certified integer weights and Decimal weights are
        produced lazily for the ORF table / edge dump.  Results are identical.
        call_weights=True also fills the Decimal `weight` column of the call rows (PB200_CALL_WEIGHTS); the
        float `score` column -- what the tabular output prints -- is always exact.
        """
        if params is None:
            params = make_params()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._generation += 1
        self._ck(self.lib.pb200_run(self.ctx, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                    params.ctypes.data,
                                    (N.REUSE_INPUT if resident else 0) | (N.LITERAL if literal else 0) |
                                    (N.CALL_WEIGHTS if call_weights else 0) | (N.INPUT_PACKED4 if packed4 and not resident else 0) |
                                    int(flags)))
        return Result(self, names) if fetch else None

    def set_chunking(self, core=256, warm=768, margin=64, long_nodes=4096):
        """Geometry (in graph nodes) of the chunked solve of long contigs; results do not depend on it.  (Until this is
        called the library picks it: 256/768/64 for contigs above 4096 nodes, 128/384/32 for a run of one or a few genomes.)"""
        self._ck(self.lib.pb200_set_chunking(self.ctx, int(core), int(warm), int(margin), int(long_nodes)))

    def last_run_ms(self) -> float:
        return float(self.lib.pb200_last_run_ms(self.ctx))

    def sizes(self):
        sz = np.zeros(8, dtype=np.int64)
        self._ck(self.lib.pb200_sizes(self.ctx, sz.ctypes.data))
        return [int(v) for v in sz]

    def pin(self, arr: np.ndarray):
        return self.lib.pb200_pin_host(arr.ctypes.data, arr.nbytes) == 0

    def unpin(self, arr: np.ndarray):
        return self.lib.pb200_unpin_host(arr.ctypes.data) == 0

    def run(self, seqs: Sequence[bytes] | Iterable[bytes], params=None, names=None, literal=False, flags=0,
            call_weights=False, trnas=None) -> Result:
        """trnas: [(contig index, start, stop)] hits of aragorn / tRNAscan-SE for THIS batch (see set_trnas)"""
        seqs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        offs = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offs[1:])
        bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
        if trnas:
            self.set_trnas(trnas)
        try:
            return self.run_packed(bases, offs, params, names, literal=literal, flags=flags, call_weights=call_weights)
        finally:
            if trnas:
                self.set_trnas(None)


class MergedResult:
    """Call and contig tables of a batch that ran as several groups (PipelinedEngine): same columns as Result,
    contig ids and table offsets renumbered to the whole batch.  `calls` and `contigs` are views into the
    engine's pinned output buffers: valid until its next run (copy them to keep them)."""

    def __init__(self, calls, contigs, first_contig, sizes, stats, launches):
        self.calls, self.contigs, self.first_contig = calls, contigs, list(first_contig)
        self.n_contigs, self.n_calls = len(contigs), len(calls)
        self.n_bases = sum(z[1] for z in sizes)
        self.n_nodes = sum(z[2] for z in sizes)
        self.n_orfs = sum(z[3] for z in sizes)
        self.n_overlaps = sum(z[4] for z in sizes)
        self.n_bridges = sum(z[5] for z in sizes)
        self.n_literal_presolve = sum(z[0] for z in stats)
        self.n_literal_postsolve = sum(z[1] for z in stats)
        self.n_literal_overlaps = sum(z[2] for z in stats)
        self.launches = launches

    check = Result.check
    call_rows = Result.call_rows


class PipelinedEngine:
    """Several contexts (streams) on one GPU, one host thread each.  A batch is cut into consecutive groups of
    contigs; while one group's kernels run, the next group's letters are being copied in, so a run from
    pinned host buffers costs little more than the kernels alone.  The call and contig tables of all groups
    land in one pinned output buffer.  Contigs are independent (phanotate.py:40-56), so grouping does not
    change any result.  For the ORF / node / edge tables use Engine."""

    def __init__(self, device: int = 0, lanes: int = 4, lib_path: str | None = None):
        from concurrent.futures import ThreadPoolExecutor
        self.engines = [Engine(device, lib_path) for _ in range(max(1, lanes))]
        self.pool = ThreadPoolExecutor(len(self.engines))
        self.device = device
        self._calls = np.zeros(0, dtype=N.CALL)
        self._calls24 = np.zeros(0, dtype=N.CALL24)
        self._contigs = np.zeros(0, dtype=N.CONTIG)

    def close(self):
        for buf in (self._calls, self._calls24, self._contigs):
            if len(buf):
                self.engines[0].unpin(buf)
        self._calls = np.zeros(0, dtype=N.CALL)
        self._calls24 = np.zeros(0, dtype=N.CALL24)
        self._contigs = np.zeros(0, dtype=N.CONTIG)
        for e in self.engines:
            e.close()
        self.pool.shutdown(wait=True)

    def pin(self, arr):
        return self.engines[0].pin(arr)

    def unpin(self, arr):
        return self.engines[0].unpin(arr)

    def _grow(self, name, rows, dtype):
        buf = getattr(self, name)
        if rows > len(buf):
            if len(buf):
                self.engines[0].unpin(buf)
            buf = np.zeros(rows + rows // 4 + 1024, dtype=dtype)
            self.engines[0].pin(buf)                  # page-locked: the device->host copies are plain DMA
            setattr(self, name, buf)
        return buf

    def pack4(self, bases):
        return self.engines[0].pack4(bases)

    def _groups(self, offsets):
        """consecutive groups of contigs, one per lane; the first is half as large as the others so that kernels start
        early -> (cut points, lanes that get contigs)"""
        n = len(offsets) - 1
        lanes = min(len(self.engines), max(n, 1))
        weights = [1.0] + [2.0] * (lanes - 1)
        import os
        if os.environ.get("PB200_LANE_WEIGHTS"):                  # experiments: relative sizes of the groups, e.g. "1,3,4,4"
            w = [float(x) for x in os.environ["PB200_LANE_WEIGHTS"].split(",")]
            weights = (w + [w[-1]] * lanes)[:lanes]
        acc, total, cuts = 0.0, sum(weights), [0]
        for k in range(lanes - 1):
            acc += weights[k]
            c = int(np.searchsorted(offsets, int(offsets[-1] * acc / total), side="left"))
            cuts.append(min(max(c, cuts[-1]), n))
        cuts.append(n)
        return cuts, [k for k in range(lanes) if cuts[k + 1] > cuts[k]]

    def run_packed(self, bases, offsets, params=None, literal=False, call_weights=False, flags=0, resident=False,
                   fetch=True, packed4=False, prefetch=None, compact=False):
        """resident=True: every lane still holds its group of this same batch from the previous call (no copy-in).
        fetch=False: leave the tables on the device (returns None).
        packed4=True: `bases` holds the 4-bit codes of the batch (pack4): half the bytes over the host link.
        prefetch=(bases, offsets) of the NEXT batch (4-bit letters, pinned, untouched until its own run_packed): its
        letters are copied in while this batch computes (pb200_prefetch_async), so that run starts without waiting for
        the host link.
        compact=True: the call rows come back as pb200_call24 records (no Decimal weight column: the columns
        Locus.tabular prints) -- half the bytes from the device."""
        if params is None:
            params = make_params()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        cuts, live = self._groups(offsets)
        nlive = len(live)
        import threading
        sized = [threading.Event() for _ in range(nlive)]        # lane j knows its table sizes
        sizes, stats, launches = [None] * nlive, [None] * nlive, [0] * nlive
        late = [False] * nlive                                   # lane j did not fit the output buffer: fetched afterwards
        ncont_total = sum(cuts[k + 1] - cuts[k] for k in live)
        contigs = self._grow("_contigs", ncont_total, N.CONTIG)[:ncont_total] if fetch else None
        cont_off = np.concatenate(([0], np.cumsum([cuts[k + 1] - cuts[k] for k in live]))).astype(np.int64)
        cname, cdtype = ("_calls24", N.CALL24) if compact else ("_calls", N.CALL)
        getter = "pb200_get_calls24" if compact else "pb200_get_calls"
        capacity = len(getattr(self, cname))                     # fixed during the run: nothing is re-pinned under a copy

        def copy_out(j):
            k = live[j]
            e = self.engines[k]
            call_off = sum(sizes[i][6] for i in range(j))
            node_off = sum(sizes[i][2] for i in range(j))
            orf_off = sum(sizes[i][3] for i in range(j))
            cl = getattr(self, cname)[call_off:call_off + sizes[j][6]]
            ct = contigs[cont_off[j]:cont_off[j + 1]]
            if len(cl):
                e._ck(getattr(e.lib, getter)(e.ctx, cl.ctypes.data))   # (rows already numbered in the whole batch)
            e._ck(e.lib.pb200_get_contigs(e.ctx, ct.ctypes.data))
            ct["call_off"] += call_off
            ct["node_off"] += node_off
            ct["orf_off"] += orf_off

        # The copies of all groups are queued up front on ONE copy stream, in group order, without blocking anybody: the
        # first (small) group's letters arrive first and its kernels start while the others are still on the link; every
        # lane's stream waits for its own letters on the device (pb200_upload_async).
        subs = []
        for j, k in enumerate(live):
            a, b = cuts[k], cuts[k + 1]
            sub_o = np.ascontiguousarray(offsets[a:b + 1] - offsets[a])
            # (4-bit letters: the group's bytes, and which nibble of the first one it starts at)
            sub_b = bases[offsets[a] // 2:(offsets[b] + 1) // 2] if packed4 else bases[offsets[a]:offsets[b]]
            subs.append((sub_b, sub_o))
            e = self.engines[k]
            e._ck(e.lib.pb200_set_contig_base(e.ctx, a))
            if not resident:
                e._ck(e.lib.pb200_upload_async(e.ctx, self.engines[live[0]].ctx, sub_b.ctypes.data,
                                               (int(offsets[a]) & 1) if packed4 else -1, sub_o.ctypes.data, len(sub_o) - 1))
        if prefetch is not None and packed4 and not resident:
            nb_, no_ = prefetch
            nb_ = np.ascontiguousarray(nb_, dtype=np.uint8)
            no_ = np.ascontiguousarray(no_, dtype=np.int64)
            cuts2, live2 = self._groups(no_)
            if live2 and live2[0] == live[0]:                       # (the same copy stream as this batch's uploads)
                for k in live2:
                    a, b = cuts2[k], cuts2[k + 1]
                    sub = nb_[no_[a] // 2:(no_[b] + 1) // 2]
                    e = self.engines[k]
                    e._ck(e.lib.pb200_prefetch_async(e.ctx, self.engines[live[0]].ctx, sub.ctypes.data, int(no_[a]) & 1,
                                                     int(no_[b] - no_[a])))

        def lane(j):
            k = live[j]
            e = self.engines[k]
            try:
                sub_b, sub_o = subs[j]
                e.run_packed(sub_b, sub_o, params, fetch=False, literal=literal, call_weights=call_weights, flags=flags,
                             resident=True)
                st = np.zeros(8, dtype=np.int64)
                e._ck(e.lib.pb200_stats(e.ctx, st.ctypes.data))
                stats[j] = [int(v) for v in st[:3]]
                launches[j] = int(e.lib.pb200_launch_count(e.ctx))
                sizes[j] = e.sizes()
            finally:
                sized[j].set()
            if not fetch:
                return
            # this lane's rows go right behind those of the lanes before it: copy out as soon as their sizes are known,
            # while later lanes are still computing
            for i in range(j):
                if not sized[i].wait(timeout=600) or sizes[i] is None:
                    raise PhanotateError("an earlier group failed")
            if sum(sizes[i][6] for i in range(j + 1)) > capacity:
                late[j] = True                         # first run or a larger batch: after the run, into a grown buffer
                return
            copy_out(j)

        list(self.pool.map(lane, range(nlive)))
        if not fetch:
            return None
        ncalls = sum(z[6] for z in sizes)
        if any(late):
            old = getattr(self, cname)[:min(capacity, ncalls)].copy()
            self._grow(cname, ncalls + ncalls // 4, cdtype)
            getattr(self, cname)[:len(old)] = old
            for j in range(nlive):
                if late[j]:
                    copy_out(j)
        calls = getattr(self, cname)[:ncalls]
        done = [(sizes[j], stats[j], launches[j]) for j in range(nlive)]
        return MergedResult(calls, contigs, [cuts[k] for k in live], sizes, [d[1] for d in done], sum(d[2] for d in done))
