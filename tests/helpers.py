"""Shared fixtures for the parity tests: golden index and the named input contigs."""
import gzip
import hashlib
import json
import os

from phanotate_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
DATA = os.path.join(HERE, "data")
FASTA = {"phiX174": "phiX174.fasta", "lambda": "NC_001416.1.fasta", "T4": "NC_000866.1.fasta"}

with open(os.path.join(GOLDEN, "index.json")) as _fh:
    INDEX = json.load(_fh)
_stress = None


def seq_of(name: str) -> str:
    global _stress
    if name in FASTA:
        return synth.read_fasta_bytes(os.path.join(DATA, FASTA[name]))[0][1].decode()
    if name.startswith("synth4_"):
        return synth.synth4_contig(int(name.split("_")[1])).decode()
    if _stress is None:
        _stress = dict(synth.stress_contigs())
    return _stress[name].decode()


def golden_text(name: str, table: str) -> str:
    p = os.path.join(GOLDEN, "%s.%s" % (name, table))
    if p.endswith(".gz"):
        return gzip.open(p, "rb").read().decode()
    return open(p).read()


def md5(text: str) -> str:
    return hashlib.md5(text.encode()).hexdigest()


STRESS = sorted((k for k in INDEX if k.startswith("stress")), key=lambda s: int(s[6:]))
