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


ROOT = os.path.abspath(os.path.join(HERE, ".."))
HOSTSIM = os.path.join(HERE, "native", "pb200_hostsim.so")


def hostsim_path() -> str:
    """Host build of the SAME stage functions the kernels run (tests only; see pb200.cu header); rebuilt when stale."""
    import subprocess
    src = os.path.join(ROOT, "phanotate_b200", "csrc")
    deps = [os.path.join(src, f) for f in os.listdir(src)] + [os.path.join(ROOT, "include", "phanotate_b200.h")]
    if not os.path.exists(HOSTSIM) or any(os.path.getmtime(d) > os.path.getmtime(HOSTSIM) for d in deps):
        subprocess.check_call(["g++", "-O2", "-x", "c++", "-std=c++17", "-DPB_HOSTSIM", "-shared", "-fPIC",
                               "-o", HOSTSIM, os.path.join(src, "pb200.cu")])
    return HOSTSIM
