"""Drop-in surface (phanotate_modules / fastpathz / phanotate.py) and the multi-rank host logic.

CPU: the mirror is driven exactly like reference phanotate.py:40-76 drives the reference, on top of the
host build of the stage functions (test-only).  GPU (-m gpu): the same through the CUDA library.
"""
import io
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import INDEX, golden_text, md5, seq_of
from phanotate_b200 import _native as N
from phanotate_b200 import dist as pdist
from phanotate_b200.engine import Engine, parse_start_codons

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HOSTSIM = os.path.join(ROOT, "tests", "native", "pb200_hostsim.so")


class Locus:
    start_codons = parse_start_codons("atg:0.85,gtg:0.10,ttg:0.05")
    stop_codons = ['tag', 'tga', 'taa']
    min_orf_len = 90

    def __init__(self, dna):
        self._d = dna

    def seq(self):
        return self._d

    def length(self):
        return len(self._d)


def drive_like_reference(name):
    """phanotate.py:40-76 with our modules substituted for the reference's."""
    import fastpathz as fz
    from phanotate_modules import functions
    from phanotate_modules.edges import Edge
    from phanotate_modules.nodes import Node  # noqa: F401  (eval)
    from phanotate_modules.file_handling import pairwise
    locus = Locus(seq_of(name))
    orfs = functions.get_orfs(locus)
    graph = functions.get_graph(orfs)
    orf_txt = "".join("%d,%d,%d,%d,%s,%s\n" % (o.start, o.stop, o.frame, o.rbs_score, o.pstop, o.weight)
                      for o in orfs.iter_orfs())
    edge_txt = "".join(str(e) + "\n" for e in graph.iteredges())
    source = "Node('source','source',0,0)"
    target = "Node('target','target',0," + str(locus.length() + 1) + ")"
    fz.empty_graph()
    for e in graph.iteredges():
        fz.add_edge(str(e))
    path = fz.get_path(source=source, target=target)[1:] if len(graph) > 2 else []
    rows = []
    for s, t in pairwise(path):
        left, right = eval(s), eval(t)
        w = graph.weight(Edge(left, right, 0))
        rows.append("%d\t%d\t%s\t%s\n" % (left.position, right.position + 2, '+' if left.frame > 0 else '-', '%E' % w))
    return orf_txt, edge_txt, "".join(rows), orfs, graph


def _check(name):
    orf_txt, edge_txt, calls, orfs, graph = drive_like_reference(name)
    g = INDEX[name]
    assert md5(orf_txt) == g["orfs_md5"] and md5(edge_txt) == g["edges_md5"] and md5(calls) == g["calls_md5"]
    assert len(graph) == g["n_nodes"] and len(orfs) == g["n_families"]
    assert str(orfs.pstop) == g["pstop"]
    o = next(orfs.iter_orfs())
    assert o.has_stop() or o.stop + 2 >= orfs.contig_length - 2 or o.frame < 0
    assert orfs.get_orf(o.start, o.stop) is o and orfs.other_end[o.start] == o.stop
    # Orf.hold (orfs.py:84) comes from the device's literal chain; Orf.score() (orfs.py:122-127) recomputed from it in Python
    # must give the device's weight
    import hashlib
    import json
    from helpers import GOLDEN
    gm = json.load(open(os.path.join(GOLDEN, "mirror.json")))["holds"]
    holds = [str(x.hold) for x in orfs.iter_orfs()]
    if name in gm:                                   # goldens from the reference's own get_orfs (make_mirror_golden.py)
        assert len(holds) == gm[name]["n"] and holds[:3] == gm[name]["first"]
        assert hashlib.md5(",".join(holds).encode()).hexdigest() == gm[name]["md5"]
    for x in orfs.iter_orfs():
        w = x.weight
        x.score()
        assert x.weight == w and str(x.weight) == str(w)


@pytest.fixture()
def sim_engine():
    import fastpathz
    from phanotate_modules import functions
    from helpers import hostsim_path
    e = Engine(0, lib_path=hostsim_path())
    functions.set_engine(e)
    fastpathz._engine = e
    yield e
    functions.set_engine(None)
    fastpathz._engine = None
    e.close()


@pytest.mark.parametrize("name", ["phiX174", "stress13", "stress26", "lambda"])
def test_mirror_driven_like_reference_cpu(sim_engine, name):
    _check(name)


def test_gcframe_class_matches_the_reference():
    """GCframe.add_base / _close / get (gc_frame_plot.py:29-74) against md5s made with the reference's class
    (tests/golden/make_mirror_golden.py), including the second get() that closes and appends again"""
    import hashlib
    import json
    import random
    from helpers import GOLDEN
    from phanotate_modules.gc_frame_plot import GCframe, max_idx, min_idx
    gm = json.load(open(os.path.join(GOLDEN, "mirror.json")))["gcframe"]
    for n, want in gm.items():
        rnd = random.Random(1000 + int(n))
        g = GCframe()
        for ch in "".join(rnd.choice("acgt") for _ in range(int(n))):
            g.add_base(ch)
        first = hashlib.md5(repr(g.get()).encode()).hexdigest()
        assert [first, hashlib.md5(repr(g.get()).encode()).hexdigest()] == want, n
    assert len(gm) >= 10
    assert [max_idx(3, 2, 1), max_idx(1, 3, 2), max_idx(1, 1, 1), min_idx(3, 2, 1), min_idx(1, 1, 1), min_idx(2, 1, 1)] == [1, 2, 3, 3, 1, 2]
    with pytest.raises(KeyError):
        GCframe().add_base("n")                      # the reference's frequency dict has no such key either


def test_cli_tabular_output_cpu(sim_engine, capsys):
    import phanotate
    phanotate.main([os.path.join(ROOT, "tests", "data", "phiX174.fasta")])
    out = capsys.readouterr().out
    want = "#id:\tphiX174\n#START\tSTOP\tFRAME\tCONTIG\tSCORE\n"
    for line in golden_text("phiX174", "calls.tsv").splitlines():
        l, r, s, sc = line.split("\t")
        want += "%s\t%s\t%s\tphiX174\t%s\n" % (l, r, s, sc)
    assert out == want


def test_cli_dump_matches_reference_cpu(sim_engine, capsys):
    import phanotate
    phanotate.main([os.path.join(ROOT, "tests", "data", "phiX174.fasta"), "--dump"])
    assert capsys.readouterr().out == golden_text("phiX174", "edges.txt.gz")


def test_cli_genbank_header_like_readme_cpu(sim_engine, capsys):
    import phanotate
    phanotate.main([os.path.join(ROOT, "tests", "data", "phiX174.fasta"), "-f", "genbank"])
    out = capsys.readouterr().out.splitlines()
    assert out[2:6] == ["     CDS             100..627", "                     /note=score:-4.827981E+02",
                        "     CDS             687..1622", "                     /note=score:-4.857517E+06"]


def test_cli_gff3_and_fasta_formats_cpu(sim_engine, capsys):
    """-f gff3 (plain GFF3; the reference's own writer is in the absent genbank package) and -f fna / faa (README.md:56-68)"""
    import phanotate
    fa = os.path.join(ROOT, "tests", "data", "phiX174.fasta")
    phanotate.main([fa, "-f", "gff3"])
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "##gff-version 3" and out[1] == "##sequence-region phiX174 1 5386"
    assert out[2].split("\t") == ["phiX174", "PHANOTATE", "CDS", "100", "627", "-4.827981E+02", "+", "0", "ID=phiX174_CDS_1"]
    assert len(out) == 2 + 7
    phanotate.main([fa, "-f", "gff"])
    assert capsys.readouterr().out.splitlines() == out
    phanotate.main([fa, "-f", "fna"])
    fna = capsys.readouterr().out.splitlines()
    assert fna[0] == ">phiX174_CDS_[100..627] [note=score:-4.827981E+02]" and fna[1].startswith("atgtttcagacttttatttctcgccataattcaaac")
    phanotate.main([fa, "-f", "faa"])
    faa = capsys.readouterr().out.splitlines()
    assert faa[1].startswith("MFQTFISRHNSNFFSDKLVLTSVTPASSAPVLQTPKATSSTLYFDSLTVNAG") and faa[1].endswith("*")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["phiX174", "stress13", "T4"])
def test_mirror_driven_like_reference_gpu(name):
    import fastpathz
    from phanotate_modules import functions
    functions.set_engine(None)
    fastpathz._engine = None
    _check(name)


def test_lpt_sharding_balances_and_covers():
    rng = np.random.default_rng(3)
    lens = rng.integers(1000, 200000, size=500)
    parts = pdist.shard_contigs(lens, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(500))
    loads = [int(lens[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= int(lens.max())


WORKER = r'''
# two ranks (gloo on the CPU; NCCL inside the library on the GPUs): every rank runs ITS shard of a batch through the host
# build, rank 0 assembles the gathered rows with the product's host logic (dist.rank_slices / dist.unshard_calls) and
# compares with the whole batch run by one process
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, torch.distributed as dist
from phanotate_b200 import _native as N, engine, dist as pdist
from helpers import STRESS, seq_of, hostsim_path
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
names = ["phiX174"] + STRESS[:9] + ["lambda"]
seqs = [seq_of(n).encode() for n in names]
parts = pdist.shard_contigs([len(s) for s in seqs], world)
e = engine.Engine(0, lib_path=hostsim_path())
mine = e.run([seqs[i] for i in parts[rank]]).calls
cnt = torch.tensor([len(mine)], dtype=torch.int64)
allc = [torch.zeros_like(cnt) for _ in range(world)]
dist.all_gather(allc, cnt)
counts = [int(c) for c in allc]
width = max(counts) * N.CALL.itemsize
buf = torch.zeros(width, dtype=torch.uint8)
buf[:mine.nbytes] = torch.from_numpy(mine.view(np.uint8).copy())
out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
dist.gather(buf, out, dst=0)
if rank == 0:
    rows = np.concatenate([np.frombuffer(o.numpy().tobytes()[:c * N.CALL.itemsize], dtype=N.CALL) for o, c in zip(out, counts)])
    assert pdist.rank_slices(counts)[1] == (counts[0], counts[1])
    whole = e.run(seqs).calls
    got = pdist.unshard_calls(rows, counts, parts)
    assert len(got) == len(whole) and np.array_equal(got, whole)
    print("GATHER_OK", counts)
dist.destroy_process_group()
'''


def test_call_table_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % (ROOT, ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert "GATHER_OK" in r.stdout, r.stdout + r.stderr


def test_product_package_does_not_import_torch():
    """north_star: host code is Python over ctypes, no PyTorch -- not even for the multi-GPU gather (csrc/comm.inc is raw NCCL)"""
    code = ("import sys; sys.path.insert(0, %r); import phanotate_b200.engine, phanotate_b200.dist, phanotate_b200.fastio, "
            "phanotate_b200.mirror, phanotate_b200.synth, phanotate_modules.functions, phanotate, fastpathz, phanotate_connect; "
            "assert 'torch' not in sys.modules, 'torch imported'; print('NO_TORCH')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "NO_TORCH" in r.stdout, r.stdout + r.stderr


def test_pipelined_engine_equals_single_context():
    """PipelinedEngine (several contexts, groups of contigs) returns the same call / contig tables as one batch."""
    import numpy as np
    from helpers import STRESS, golden_text, seq_of
    from phanotate_b200 import engine
    from helpers import hostsim_path
    hostsim_path()
    names = ["phiX174"] + STRESS[:12] + ["lambda"] + STRESS[12:24]
    seqs = [seq_of(n).encode() for n in names]
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    pe = engine.PipelinedEngine(0, lanes=3, lib_path=HOSTSIM)
    e = engine.Engine(0, lib_path=HOSTSIM)
    m, s = pe.run_packed(bases, offs), e.run_packed(bases, offs)
    assert np.array_equal(m.calls, s.calls) and np.array_equal(m.contigs, s.contigs)
    for k, n in enumerate(names):
        assert "".join("%d\t%d\t%s\t%s\n" % r for r in m.call_rows(k)) == golden_text(n, "calls.tsv")
    assert int(m.contigs[len(names) - 1]["length"]) == len(seqs[-1])
    m2 = pe.run_packed(bases[:offs[5]], offs[:6])            # the output buffers are reused from run to run
    assert np.array_equal(m2.calls, s.calls[:m2.n_calls]) and m2.n_contigs == 5
    pe.close()
    e.close()


def test_fast_fasta_ingest_equals_line_reader(tmp_path):
    """fastio.read_fasta_packed (vectorised) against phanotate_modules.file.File on a multi-record file with CRLF
    line ends, blank lines, mixed case, text before the first header and a record without sequence; also gzipped."""
    import gzip
    import numpy as np
    from phanotate_b200 import fastio
    from phanotate_modules.file import File
    text = ("junk before any header\n>rec1 first record\r\nACGTacgtnnRY\r\n\r\nacgt \n>empty\n>rec3\tdescription\n" +
            "tttt\ncccc\n" + open(os.path.join(ROOT, "tests", "data", "phiX174.fasta")).read())
    p = tmp_path / "m.fasta"
    p.write_text(text)
    pz = tmp_path / "m.fasta.gz"
    pz.write_bytes(gzip.compress(text.encode()))
    from helpers import hostsim_path
    from phanotate_b200 import _native as N
    lib = N.load(hostsim_path())
    for path in (p, pz):
        names, bases, offs = fastio.read_fasta_packed(str(path), lib)
        loci = list(File(str(path)))
        assert names == [l.name() for l in loci] == ["rec1", "empty", "rec3", "phiX174"]
        assert [bases[offs[k]:offs[k + 1]].tobytes().decode() for k in range(len(loci))] == [l.seq() for l in loci]


def test_cli_tabular_fast_path_equals_per_locus_writer(sim_engine, capsys, tmp_path):
    """The batch tabular writer (CLI default format) produces the bytes of Locus.tabular (locus.py:39-56), reverse-strand
    rows included, on a two-record file."""
    import io
    import phanotate
    from phanotate_modules.file import File
    from phanotate_b200.engine import make_params
    data = os.path.join(ROOT, "tests", "data")
    p = tmp_path / "two.fasta"
    p.write_text(open(os.path.join(data, "NC_001416.1.fasta")).read() + open(os.path.join(data, "phiX174.fasta")).read())
    phanotate.main([str(p)])
    fast = capsys.readouterr().out
    loci = list(File(str(p)))
    res = sim_engine.run([l.seq().encode() for l in loci], make_params())
    slow = io.StringIO()
    for k, locus in enumerate(loci):
        c = res.contigs[k]
        for r in res.calls[c["call_off"]:c["call_off"] + c["n_calls"]]:
            f = locus.add_feature('CDS', 1 if r["strand"] > 0 else -1, [[int(r["left"]), int(r["right"]) - 2]], {})
            f.weight = '%E' % float(r["score"])
        locus.tabular(slow)
    assert fast == slow.getvalue() and "\t-\t" in fast


@pytest.mark.gpu
def test_cli_on_the_gpu_tabular_dump_and_genbank(capsys, tmp_path):
    """The command line itself on the CUDA path (no engine injected): a three-record FASTA (T4, lambda, phiX174, one of them
    gzipped away from the others) -> tabular rows == the reference goldens; --dump == the reference's edge text; the other
    formats carry the same calls."""
    import gzip
    import hashlib
    import fastpathz
    import phanotate
    from phanotate_modules import functions
    functions.set_engine(None)
    fastpathz._engine = None
    data = os.path.join(ROOT, "tests", "data")
    names = [("T4", "NC_000866.1.fasta"), ("lambda", "NC_001416.1.fasta"), ("phiX174", "phiX174.fasta")]
    text = "".join(open(os.path.join(data, f)).read() for _, f in names)
    p = tmp_path / "three.fasta"
    p.write_text(text)
    gz = tmp_path / "three.fasta.gz"
    gz.write_bytes(gzip.compress(text.encode()))
    for path in (p, gz):
        phanotate.main([str(path)])
        out = capsys.readouterr().out
        blocks = out.split("#id:\t")[1:]
        assert len(blocks) == 3
        for (nm, _), b in zip(names, blocks):
            rows = [l.split("\t") for l in b.splitlines()[2:]]
            got = "".join("%s\t%s\t%s\t%s\n" % (((r[0], r[1]) if r[2] == "+" else (r[1], r[0])) + (r[2], r[4])) for r in rows)
            assert got == golden_text(nm, "calls.tsv"), nm
    px = tmp_path / "phix.fasta"
    px.write_text(open(os.path.join(data, "phiX174.fasta")).read())
    phanotate.main([str(px), "--dump"])
    assert hashlib.md5(capsys.readouterr().out.encode()).hexdigest() == INDEX["phiX174"]["edges_md5"]
    phanotate.main([str(p), "-f", "genbank"])
    gb = capsys.readouterr().out
    assert gb.count("LOCUS") == 3 and gb.count("     CDS             ") == sum(INDEX[n]["n_calls"] for n, _ in names)
    functions.set_engine(None)


def _comm_worker_script():
    return r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
from phanotate_b200 import _native as N, engine, dist as pdist
from helpers import STRESS, seq_of
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
names = ["phiX174"] + STRESS[:9] + ["lambda", "T4"]
seqs = [seq_of(n).encode() for n in names]
e = engine.Engine(int(os.environ.get("LOCAL_RANK", "0")))
comm = pdist.Comm(e, rank, world)
parts = pdist.shard_contigs([len(s) for s in seqs], world)
for rep in range(3):                                  # (the gather buffer and the pinned host buffers are reused)
    mine = e.run([seqs[i] for i in parts[rank]])
    counts, total = comm.gather_calls()
    assert counts[rank] == mine.n_calls and total == sum(counts)
    s = comm.allreduce([mine.n_calls, rank], "sum")
    assert int(s[0]) == total and int(s[1]) == world * (world - 1) // 2
    assert comm.allreduce([float(rank)], "max")[0] == world - 1
    if rank == 0:
        rows = np.array(comm.fetch(0, total), copy=True)
        comm.fetch_begin(0, total)
        assert np.array_equal(comm.fetch_wait(), rows)
# the same gather with compact rows (pb200_call24: converted on the device on the way; several segments per rank)
mine = e.run([seqs[i] for i in parts[rank]])
counts24, total24 = comm.gather_calls([e, e], compact=True)
assert total24 == 2 * total and counts24[rank] == 2 * mine.n_calls
if rank == 0:
    rows24 = np.array(comm.fetch(0, total24), copy=True)
    assert rows24.dtype == N.CALL24
    comm.fetch_begin(0, total24)
    assert np.array_equal(comm.fetch_wait(), rows24)
comm.barrier()
if rank == 0:
    whole = e.run(seqs).calls
    got = pdist.unshard_calls(rows, counts, parts)
    assert len(got) == len(whole) and np.array_equal(got, whole)
    at = 0
    for r in range(world):                       # rank r's rows twice (its two segments), column by column
        n = counts[r]
        ref = rows[sum(counts[:r]):sum(counts[:r]) + n]
        for rep in range(2):
            blk = rows24[at:at + n]
            for col in ("contig", "left", "right", "strand", "score"):
                assert np.array_equal(blk[col], ref[col]), (r, rep, col)
            at += n
    print("NCCL_GATHER_OK", world, counts, e.lib.pb200_comm_nccl_version())
comm.close()
e.close()
''' % (ROOT, ROOT)


def _gpu_count():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cuda.cuInit(0) != 0:
            return 0
        cuda.cuDeviceGetCount(ctypes.byref(n))
        return n.value
    except OSError:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
def test_nccl_gather_inside_the_library(tmp_path, world):
    """The library's own communicator (csrc/comm.inc: raw NCCL, no PyTorch): every rank runs its LPT shard of a batch on its
    GPU, the call tables are gathered to rank 0 (exact row counts), copied to the host both ways (blocking, and on the copy
    stream), reductions and barrier work; the re-assembled table equals the whole batch run by one process.  One rank on
    any GPU box; two ranks where the box has two GPUs."""
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "w.py"
    script.write_text(_comm_worker_script())
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29540 + world))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert "NCCL_GATHER_OK %d" % world in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
