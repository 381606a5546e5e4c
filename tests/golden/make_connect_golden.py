"""Generates tests/golden/connect.json from the COMPILED REFERENCE extension (oracle/_ref/phanotate_connect*.so, built by
`make -C oracle` from /root/reference/src/phanotate_connect.c).  The extension keeps global state and has no reset, so
every case runs in its own interpreter.

    python tests/golden/make_connect_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.path.join(ROOT, "oracle", "_ref")

CHILD = r"""
import sys, json
sys.path.insert(0, %r)
import phanotate_connect as pc
edges = json.load(sys.stdin)
for l, r in edges:
    pc.add_edge(l, r)
print(json.dumps(pc.get_connected()))
"""


def reference_rows(edges):
    out = subprocess.run([sys.executable, "-c", CHILD % REF], input=json.dumps(edges), capture_output=True, text=True, check=True)
    return json.loads(out.stdout)


def cases():
    """name -> list of (left, right); the same generator is imported by tests/test_connect.py"""
    out = {"empty": [], "single": [(100, 400)], "self_only": [(5, 5)],
           "readme_like": [(100, 400), (350, 900), (100, 700), (650, 1000)]}
    rng = np.random.Generator(np.random.PCG64(20261017))
    # ORF-like: left < right, lengths 90..3000, positions on a 50-kb contig, with repeated stops (families) and duplicates
    l = rng.integers(1, 50000, size=400)
    r = l + rng.integers(90, 3000, size=400)
    e = [(int(a), int(b)) for a, b in zip(l, r)]
    out["orf_like_400"] = e + e[:25]
    # dense: many equal keys and distances right at the 300 / 301 boundary
    base = rng.integers(0, 2000, size=150)
    out["boundary_150"] = [(int(b), int(b) + int(d)) for b, d in zip(base, rng.choice([0, 1, 299, 300, 301, 600], size=150))]
    out["negative_and_zero"] = [(-500, -250), (-250, 0), (0, 250), (-100, 100), (100, -100), (0, 0)]
    # more than one 2048-entry chunk of the CUDA kernel
    l = rng.integers(0, 400000, size=4500)
    r = l + rng.integers(-200, 2500, size=4500)
    out["chunks_4500"] = [(int(a), int(b)) for a, b in zip(l, r)]
    return out


def main():
    gold = {}
    for name, edges in cases().items():
        rows = reference_rows(edges)
        assert all(t[2] == 0 for t in rows)
        flat = np.asarray([t[:2] for t in rows], dtype=np.int32).reshape(-1, 2)
        gold[name] = {"n_edges": len(edges), "n_rows": len(rows), "md5": hashlib.md5(flat.tobytes()).hexdigest(),
                      "rows": [t[:2] for t in rows] if len(rows) <= 64 else None, "head": [t[:2] for t in rows[:8]]}
    with open(os.path.join(HERE, "connect.json"), "w") as fh:
        json.dump(gold, fh, indent=1)
    print({k: v["n_rows"] for k, v in gold.items()})


if __name__ == "__main__":
    main()
