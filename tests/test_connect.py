"""the join of the reference's C extension (src/phanotate_connect.c:78-121).

CPU: oracle/connect_oracle.py against goldens generated from the COMPILED reference (oracle/_ref, `make -C oracle`) and,
where oracle/_ref is present, against the compiled reference itself on random edge lists; the item function of the CUDA
kernels in the host build.  GPU: pb200_connect / the drop-in module `phanotate_connect` against the oracle.
"""
import ctypes
import glob
import hashlib
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, hostsim_path

sys.path.insert(0, GOLDEN)
import make_connect_golden as MG          # noqa: E402  (case generator; the reference run itself only in the tests below)
from oracle import connect_oracle as CO    # noqa: E402

GOLD = json.load(open(os.path.join(GOLDEN, "connect.json")))
CASES = MG.cases()
REF_SO = glob.glob(os.path.join(ROOT, "oracle", "_ref", "phanotate_connect*.so"))


def md5_rows(rows):
    return hashlib.md5(np.asarray(rows, dtype=np.int32).reshape(-1, 2).tobytes()).hexdigest()


def check_against_golden(name, rows):
    g = GOLD[name]
    assert len(rows) == g["n_rows"], name
    assert md5_rows(rows) == g["md5"], name
    if g["rows"] is not None:
        assert [list(map(int, r)) for r in rows] == g["rows"], name


def split(edges):
    return [e[0] for e in edges], [e[1] for e in edges]


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_compiled_reference_goldens(name):
    assert GOLD[name]["n_edges"] == len(CASES[name])
    check_against_golden(name, CO.get_connected(*split(CASES[name])))


@pytest.mark.skipif(not REF_SO, reason="oracle/_ref not built (needs /root/reference; `make -C oracle`)")
def test_oracle_matches_compiled_reference_on_random_edges():
    rng = np.random.Generator(np.random.PCG64(7))
    for trial in range(6):
        n = int(rng.integers(1, 700))
        span = int(rng.choice([500, 5000, 100000]))
        l = rng.integers(-span, span, size=n)
        r = l + rng.integers(-400, 3000, size=n)
        edges = [(int(a), int(b)) for a, b in zip(l, r)]
        want = MG.reference_rows(edges)
        got = CO.get_connected(*split(edges))
        assert [(int(a), int(b), 0) for a, b in got] == [tuple(t) for t in want]


@pytest.fixture(scope="module")
def sim():
    from phanotate_b200.engine import Engine
    e = Engine(0, lib_path=hostsim_path())
    yield e
    e.close()


def run_lib(engine, edges):
    import phanotate_connect as pc
    l, r = split(edges)
    return pc.connected_arrays(l, r, engine=engine)


@pytest.mark.parametrize("name", sorted(CASES))
def test_item_function_on_host_matches_goldens(sim, name):
    check_against_golden(name, run_lib(sim, CASES[name]))


def test_capacity_protocol_on_host(sim):
    l, r = (np.asarray(v, dtype=np.int32) for v in split(CASES["readme_like"]))
    rows = ctypes.c_int64(-1)
    small = np.full((2, 2), -7, dtype=np.int32)
    assert sim.lib.pb200_connect(sim.ctx, l.ctypes.data, r.ctypes.data, len(l), small.ctypes.data, 2, ctypes.byref(rows)) == 0
    assert rows.value == 4 and (small == -7).all()          # too small: only the count comes back
    assert sim.lib.pb200_connect(sim.ctx, l.ctypes.data, r.ctypes.data, -1, None, 0, ctypes.byref(rows)) != 0


def test_module_surface_matches_the_extension():
    import phanotate_connect as pc
    pc.clear()
    with pytest.raises(TypeError):
        pc.add_edge(1.5, 2)
    with pytest.raises(OverflowError):
        pc.add_edge(1 << 31, 2)
    with pytest.raises(TypeError):
        pc.add_edge(1)
    pc.clear()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def eng():
    from phanotate_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_join_matches_goldens(eng, name):
    check_against_golden(name, run_lib(eng, CASES[name]))


@pytest.mark.gpu
def test_cuda_join_matches_oracle_on_random_edges(eng):
    rng = np.random.Generator(np.random.PCG64(11))
    for n in (1, 2, 255, 256, 257, 2047, 2048, 2049, 6000):
        l = rng.integers(0, 40 * n + 50, size=n)
        r = l + rng.integers(-300, 3000, size=n)
        edges = [(int(a), int(b)) for a, b in zip(l, r)]
        got = run_lib(eng, edges)
        want = CO.get_connected(*split(edges))
        assert got.shape == want.shape and (got == want).all(), n


@pytest.mark.gpu
def test_cuda_join_drop_in_module_and_size_independent_properties(eng):
    """Module-level state like the extension's; at a size the oracle does not reach: every row satisfies the predicate,
    rows are right-entry-major, and the row count equals a sort-based count of the same predicate."""
    import phanotate_connect as pc
    pc.clear()
    for l, r in CASES["readme_like"]:
        pc.add_edge(l, r)
    assert pc.get_connected() == [(400, 350, 0), (400, 650, 0), (900, 650, 0), (700, 650, 0)]
    assert pc.get_connected(min_distance=10) == pc.get_connected()        # ignored, as in the reference
    pc.clear()
    rng = np.random.Generator(np.random.PCG64(13))
    n = 60000
    l = rng.integers(0, 3_000_000, size=n).astype(np.int32)
    r = (l + rng.integers(90, 3000, size=n)).astype(np.int32)
    rows = pc.connected_arrays(l, r, engine=eng)
    # the same rows from a different algorithm: a window over the sorted left ends, candidates re-ordered by insertion index
    order = np.argsort(l, kind="stable")
    ls = l[order].astype(np.int64)
    lo = np.searchsorted(ls, r.astype(np.int64) - 300, side="left")
    hi = np.searchsorted(ls, r.astype(np.int64) + 300, side="right")
    want = []
    for i in range(n):
        js = np.sort(order[lo[i]:hi[i]])
        js = js[(r[js] != r[i]) & (l[js] != l[i])]
        if len(js):
            want.append(np.stack([np.full(len(js), r[i]), l[js]], axis=1))
    want = np.concatenate(want, axis=0)
    assert rows.shape == want.shape and (rows == want).all()
