"""The oracle (oracle/elba_oracle.cpp) against the golden vectors produced by the REFERENCE'S OWN code
(tests/golden/golden.json, written by tests/golden/make_golden.py from oracle/_ref).  CPU only."""
import numpy as np
import pytest

from common import check_seeds_valid, digest, oracle_digests
from oracle import oracle as O


def test_kmer_value_goldens(golden):
    """SURVEY.md §8c golden vectors: forward / twin / canonical value, GetHash, GetKmerOwner."""
    for s, g in golden["values"]["kmers"].items():
        fwd, twin, rep, h = O.kmer_info(s)
        assert (hex(fwd), hex(twin), hex(rep), hex(h)) == (g["fwd"], g["twin"], g["rep"], g["hash"]), s
        assert O.owner(rep, 8) == g["owner8"] and O.owner(rep, 4) == g["owner4"], s
    # the survey's own table
    assert O.kmer_info("ACGTACGTACGTACGTA") == (0x1b1b1b1b00000000, 0xc6c6c6c6c0000000, 0x1b1b1b1b00000000, 0x2614c45a50619313)
    assert O.kmer_info("TTTTTTTTTTTTTTTTT")[2:] == (0x0, 0x864ba144df098483)
    assert O.kmer_info("GATTACAGATTACAGAT")[1:] == (0x37b0dec340000000, 0x37b0dec340000000, 0x90d629b9ebbd8e01)
    assert O.kmer_info("ACGTACGTACGTACGTACGTACGTACGTACG")[3] == 0x7b2460db6e800453
    assert O.owner(0x1b1b1b1b00000000, 8) == 1 and O.owner(0x1b1b1b1b00000000, 4) == 0 and O.owner(0, 8) == 4


def test_hash_goldens(golden):
    for x, h in golden["values"]["hash64"].items():
        assert hex(O.kmer_hash(int(x, 16))) == h


def test_bloom_goldens(golden):
    g = golden["values"]["bloom"]
    assert O.bloom_size(g["entries"], g["error"]) == (g["bits"], g["hashes"])
    assert O.bloom_size(283870) == (1769993, 5) and O.bloom_size(2450167) == (15277340, 5)
    assert O.bloom_ab(0x1b1b1b1b00000000) == (0x7bea1f538296949f, 0xadf6fcb66b1b7b43)
    keys = np.array([int(x, 16) for x in g["keys"]], np.uint64)
    assert digest(O.bloom_fill(g["entries"], g["error"], keys)) == g["bf_sha256"]


def test_hll_goldens(golden, fixtures):
    for name, k in (("reads_fa", 17), ("reads_fa", 31), ("example_medium", 17)):
        g = golden["values"][f"hll_{name}_k{k}"]
        est, regs = O.hll(fixtures(name), k)
        assert est == g["estimate"] and digest(regs) == g["registers_sha256"], (name, k)
    assert abs(golden["values"]["hll_reads_fa_k17"]["estimate"] - 283870.0) < 0.05
    assert abs(golden["values"]["hll_example_medium_k17"]["estimate"] - 2450166.4) < 0.05


@pytest.mark.parametrize("key", ["reads_fa_k17_l2_u8", "reads_fa_first135_k17_l2_u8", "reads_fa_k31_l2_u4", "reads_fa_k31_l15_u35",
                                 "example_medium_k17_l2_u8", "example_medium_k31_l2_u4", "example_medium_k31_l15_u35"])
def test_whole_path_digests(golden, fixtures, key):
    """Tier 1 of the whole path == what the reference's KmerOps.cpp + SharedSeeds.cpp produced (np=1 and np=4)."""
    g = golden["configs"][key]
    dna = fixtures(g["fixture"])
    assert dna.size() == g["N"] and dna.num_kmers(g["k"]) == g["M"]
    r = O.run(dna, g["k"], g["lower"], g["upper"], threads=4)
    assert (r.R, r.nnzA, r.nnzB_pre, r.nnzB) == (g["R"], g["nnzA"], g["nnzB_pre"], g["nnzB"])
    assert oracle_digests(r) == g["digests"]
    assert int(r.b_num.sum()) == g["numshared_sum"]
    assert check_seeds_valid(dna, g["k"], r.b_rowptr, r.b_col, r.b_seeds, max_checks=1500) == 0


def test_survey_size_table(golden):
    """The rows of SURVEY.md §8 / BASELINE.md §2."""
    c = golden["configs"]
    t = c["reads_fa_k17_l2_u8"]
    assert (t["N"], t["M"], t["R"], t["nnzA"], t["nnzB_pre"], t["nnzB"], t["diag"], t["strict_upper"]) == (227, 3321268, 14751, 50953, 2550, 2477, 219, 1129)
    t = c["example_medium_k17_l2_u8"]
    assert (t["N"], t["M"], t["R"], t["nnzA"], t["nnzB_pre"], t["nnzB"], t["diag"], t["strict_upper"]) == (1989, 28872543, 118856, 285499, 23383, 22404, 1918, 10243)
    t = c["example_medium_k31_l15_u35"]
    assert (t["R"], t["nnzA"], t["nnzB_pre"], t["nnzB"], t["diag"], t["strict_upper"]) == (914431, 23715519, 116802, 115612, 1988, 56812)
    t = c["example_medium_k31_l2_u4"]
    assert (t["R"], t["nnzA"], t["nnzB_pre"], t["nnzB"]) == (176084, 391755, 18597, 18270)


def test_bloom_gated_pass1_is_result_neutral(fixtures, golden):
    """with LOWER >= 2 the Bloom-gated first pass only decides which singletons enter the map;
    the number of keys after pass 1 matches the reference's own run."""
    dna = fixtures("reads_fa")
    est, _ = O.hll(dna, 17)
    assert O.pass1_keys(dna, 17, int(np.ceil(est))) == golden["configs"]["reads_fa_k17_l2_u8"]["keys_after_pass1"]
