"""Input data formats: read_fasta_file (suffix-array-searching/src/util.rs:144-169) and the --human k-mer
keys (static-search-tree/src/bin/bench.rs:60-76).  CPU: oracle restatements on hand-made cases.
GPU: the CUDA encoders against them, and the whole chain FASTA -> codes -> keys -> tree -> queries."""
import numpy as np
import pytest

from util import MAX

CASES = [
    b"",
    b">only header\n",
    b">r1\nACGT\n",
    b">r1\nACGT",                                   # no trailing newline
    b">r1 desc\r\nAC\r\nGT\r\n>r2\r\nTTaacgNNx>\r\n",  # CRLF, lower case, N and junk -> 0, '>' inside a line
    b"\n\n>r\n\nAC\n\nGT\n\n",                      # empty lines
    b"ACGT\n>r\nGG\n",                              # sequence before any header (no '>' at its line start)
    b">a\n" + b"ACGTTGCA" * 1000 + b"\n>b\n" + b"\n".join([b"GATTACA" * 10] * 50) + b"\n",
]


def test_oracle_fasta_known_answers(oracle):
    assert oracle.read_fasta(b">r1\nACGT\n").tolist() == [0, 1, 2, 3]
    assert oracle.read_fasta(CASES[4]).tolist() == [0, 1, 2, 3, 3, 3, 0, 0, 1, 2, 0, 0, 0, 0]
    assert oracle.read_fasta(b">x\n").size == 0 and oracle.read_fasta(b"").size == 0
    assert oracle.read_fasta(CASES[6]).tolist() == [0, 1, 2, 3, 2, 2]


FASTQ_CASES = [
    b"@r1\nACGT\n+\nIIII\n",
    b"@r1 desc\nACGT\n+r1\n@III\n@r2\nggta\n+\n>>>>\n",     # '@' and '>' inside quality lines, lower case
    b"@r1\r\nACNT\r\n+\r\nIIII\r\n@r2\r\nTT\r\n+\r\nII",    # CRLF, N -> 0, no trailing newline
    b"@empty\n\n+\n\n@r\nCC\n+\nII\n",                      # an empty read
]


def test_oracle_fastq_known_answers(oracle):
    """needletail::parse_fastx_file (util.rs:161) also reads FASTQ: four-line records, the second line is the sequence."""
    assert oracle.read_fasta(FASTQ_CASES[0]).tolist() == [0, 1, 2, 3]
    assert oracle.read_fasta(FASTQ_CASES[1]).tolist() == [0, 1, 2, 3, 2, 2, 3, 0]
    assert oracle.read_fasta(FASTQ_CASES[2]).tolist() == [0, 1, 0, 3, 3, 3]
    assert oracle.read_fasta(FASTQ_CASES[3]).tolist() == [1, 1]


def test_oracle_kmer_keys_known_answers(oracle):
    codes = np.array([0, 1, 2, 3, 3, 2, 1, 0], np.uint8)
    k3 = oracle.kmer_keys(codes, k=3, sort=False)
    # windows 012 123 233 332 321 210 -> base-4 values; key 0 replaced by MAX (bench.rs:74)
    assert k3.tolist() == [MAX, 27, 47, 62, 57, 36]
    assert oracle.kmer_keys(codes, k=3, sort=True).tolist() == sorted(k3.tolist())
    assert oracle.kmer_keys(codes, k=9).size == 0
    assert oracle.kmer_keys(codes, k=3, max_keys=2, sort=False).tolist() == [MAX, 27]
    # k = 16 uses all 32 bits before the i32::MAX mask (bench.rs:71-72)
    c16 = np.full(17, 3, np.uint8)
    assert oracle.kmer_keys(c16, k=16, sort=False).tolist() == [MAX, 0xFFFFFFFF & MAX]


@pytest.mark.gpu
def test_gpu_fasta_matches_oracle(gpu, oracle):
    sst = gpu
    rng = np.random.default_rng(5)
    big = b"".join(b">rec%d some text\n" % i + b"\n".join(bytes(rng.choice(list(b"ACGTacgtN"), 61).tolist()) for _ in range(int(rng.integers(1, 40)))) + b"\n"
                   for i in range(200))
    bigq = b"".join(b"@read%d\n" % i + bytes(rng.choice(list(b"ACGTacgtN"), n).tolist()) + b"\n+\n" + bytes(rng.choice(list(b"@>+IJ#"), n).tolist()) + b"\n"
                    for i, n in enumerate(rng.integers(1, 300, 500)))
    for data in CASES + [big] + FASTQ_CASES + [bigq]:
        got = sst.read_fasta(data)
        assert np.array_equal(got, oracle.read_fasta(data)), data[:40]


@pytest.mark.gpu
def test_gpu_kmer_keys_match_oracle(gpu, oracle):
    sst = gpu
    rng = np.random.default_rng(6)
    codes = rng.integers(0, 4, 50_000, dtype=np.uint8)
    for k in (1, 3, 15, 16, 17, 32):
        for sort in (False, True):
            assert np.array_equal(sst.kmer_keys(codes, k=k, sort=sort), oracle.kmer_keys(codes, k=k, sort=sort)), (k, sort)
    assert np.array_equal(sst.kmer_keys(codes, k=16, max_keys=1000), oracle.kmer_keys(codes, k=16, max_keys=1000))
    assert sst.kmer_keys(codes[:10], k=16).size == 0


@pytest.mark.gpu
def test_gpu_fasta_to_tree_pipeline(gpu, oracle):
    """bench.rs --human: FASTA -> 16-mer keys -> sort -> STree16 -> lower_bound, and the SA path on the same text."""
    sst = gpu
    rng = np.random.default_rng(7)
    seq = bytes(rng.choice(list(b"ACGT"), 300_000).tolist())
    fasta = b">chrTest\n" + b"\n".join(seq[i : i + 80] for i in range(0, len(seq), 80)) + b"\n"
    codes = sst.read_fasta(fasta)
    assert codes.size == len(seq)
    keys = sst.kmer_keys(codes, k=16, sort=True)
    assert keys.size == len(seq) - 15 and keys[-1] == MAX and bool((np.diff(keys.astype(np.int64)) >= 0).all())
    qs = rng.integers(0, MAX, 100_000, dtype=np.uint32)
    ev, ei = oracle.lower_bound(keys, qs)
    v, i = sst.STree16.new_params(keys, True, False, False).query(qs, want_index=True)
    assert np.array_equal(v, ev) and np.array_equal(i, ei)
    sa = sst.SaNaive.build(codes)
    assert sa.check() == 0
    pat = codes[12345 : 12345 + 40].tobytes()
    assert codes[sa.binary_search(pat) :][:40].tobytes() == pat
