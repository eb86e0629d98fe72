"""The device parse code (elba_b200/csrc/common.cuh: 2-bit unpack, rolling forward / reverse-complement windows,
canonicalisation, chunking) compiled for the HOST with intrinsic stand-ins and compared with the oracle.
Catches bit-twiddling mistakes without a GPU."""
import ctypes
import os
import subprocess

import numpy as np

from elba_b200.dnabuffer import DnaBuffer
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostparse") / "host_parse.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "tests", "host_shim"),
                    "-o", out, os.path.join(ROOT, "tests", "host_parse_check.cpp")], check=True)
    L = ctypes.CDLL(out)
    L.host_parse.restype = ctypes.c_uint64
    return L


def test_device_parse_on_host(tmp_path_factory):
    L = _lib(tmp_path_factory)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(1)
    lens = [0, 1, 5, 16, 17, 18, 31, 32, 33, 48, 49, 50, 63, 64, 65, 100, 127, 128, 129, 1000, 33, 4097]
    seqs = ["".join("ACGT"[c] for c in rng.integers(0, 4, l)) for l in lens]
    seqs += ["A" * 200, "T" * 77, "ACGT" * 30]
    dna = DnaBuffer.from_strings(seqs)
    buf = np.concatenate([dna.buf, np.zeros(64, np.uint8)])
    for k in (3, 15, 17, 21, 31, 32):
        for stride in (1, 3):
            M = dna.num_kmers(k)
            ok, op, orr = np.zeros(M, np.uint64), np.zeros(M, np.uint32), np.zeros(M, np.uint32)
            m = L.host_parse(p(buf), p(dna.offsets), p(dna.lengths), dna.size(), k, stride, p(ok), p(op), p(orr))
            want_k, want_p, want_r = [], [], []
            for i in range(dna.size()):
                ks = O.rep_kmers(dna.buf[int(dna.offsets[i]):], int(dna.lengths[i]), k)
                sel = np.arange(len(ks))[::stride]
                want_k.append(ks[sel]); want_p.append(sel); want_r.append(np.full(len(sel), i))
            want_k, want_p, want_r = np.concatenate(want_k), np.concatenate(want_p), np.concatenate(want_r)
            assert m == len(want_k), (k, stride)
            assert np.array_equal(ok[:m], want_k) and np.array_equal(op[:m], want_p) and np.array_equal(orr[:m], want_r), (k, stride)
