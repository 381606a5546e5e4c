#!/usr/bin/env python3
"""Goldens for two pieces of the API mirror, from the reference's own code (run in the build container):

* GCframe (gc_frame_plot.py:29-74): md5 of repr(get()) -- and of a second get(), which closes and appends again -- for
  seeded random sequences of several lengths;
* Orf.hold (orfs.py:84, functions.py:286-298): md5 of the holds of every ORF of phiX174 and stress13 in
  Orfs.iter_orfs() order.

    python tests/golden/make_mirror_golden.py   ->  tests/golden/mirror.json
"""
import hashlib
import importlib.util
import json
import os
import random
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

GC_CASES = [(n, 1000 + n) for n in (60, 61, 62, 63, 119, 120, 121, 122, 500, 1001, 5386)]


def gc_sequence(n, seed):
    rnd = random.Random(seed)
    return "".join(rnd.choice("acgt") for _ in range(n))


def main():
    spec = importlib.util.spec_from_file_location("refgc", "/root/reference/phanotate_modules/gc_frame_plot.py")
    refgc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refgc)
    out = {"gcframe": {}, "holds": {}}
    for n, seed in GC_CASES:
        g = refgc.GCframe()
        for ch in gc_sequence(n, seed):
            g.add_base(ch)
        first = hashlib.md5(repr(g.get()).encode()).hexdigest()
        second = hashlib.md5(repr(g.get()).encode()).hexdigest()
        out["gcframe"]["%d" % n] = [first, second]
    sys.path.insert(0, "/root/reference")
    import make_golden as MG
    from helpers import seq_of
    from phanotate_modules import functions            # the reference's (first on sys.path)
    assert functions.__file__.startswith("/root/reference")
    for name in ("phiX174", "stress13"):
        orfs = functions.get_orfs(MG.LocusShim(name, seq_of(name)))
        holds = [str(o.hold) for o in orfs.iter_orfs()]
        out["holds"][name] = {"n": len(holds), "md5": hashlib.md5(",".join(holds).encode()).hexdigest(), "first": holds[:3]}
    with open(os.path.join(HERE, "mirror.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(out["holds"]))


if __name__ == "__main__":
    main()
