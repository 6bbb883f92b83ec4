"""GPU parity tests (-m gpu): suffix array construction and pattern search vs the CPU oracle."""
import numpy as np
import pytest

from util import random_patterns, random_text

pytestmark = pytest.mark.gpu


def _check_search(sst, oracle, sa_obj, text, sa, pats):
    flat, off = sst.pack_patterns(pats)
    oflat, ooff = oracle.pack_patterns(pats)
    elo, ehi, epos, _ = oracle.sa_search(text, sa, oflat, ooff)
    for mode in (sst.SA_BINARY, sst.SA_MLR):
        lo, hi, pos = sa_obj.search(flat, off, mode)
        assert np.array_equal(lo, elo), mode
        assert np.array_equal(hi, ehi), mode
        assert np.array_equal(pos, epos), mode
    lo, hi, pos = sa_obj.search(flat, off, sst.SA_BINARY, want_hi=False)
    assert hi is None and np.array_equal(lo, elo) and np.array_equal(pos, epos)


@pytest.mark.parametrize("n,sigma", [(1, 4), (2, 4), (7, 4), (8, 2), (100, 4), (5000, 4), (60_000, 4), (60_000, 256), (30_000, 1), (40_000, 2)])
def test_sa_build_matches_oracle(gpu, oracle, n, sigma):
    """The suffix array is unique: GPU prefix doubling == oracle sort; strict order holds (sa_search.rs:36-38)."""
    sst = gpu
    text = random_text(n, seed=n + sigma, sigma=sigma)
    s = sst.SaNaive.build(text)
    sa = s.sa
    assert np.array_equal(sa, oracle.sa_build(text))
    assert s.check() == 0 and oracle.sa_check(text, sa) == 0


def test_sa_build_repetitive(gpu, oracle):
    sst = gpu
    unit = random_text(37, seed=1)
    text = np.concatenate([np.tile(unit, 300), random_text(500, seed=2), np.tile(unit, 100)])
    s = sst.SaNaive.build(text)
    assert np.array_equal(s.sa, oracle.sa_build(text))
    assert s.check() == 0


def test_sa_check_detects_corruption(gpu, oracle):
    sst = gpu
    text = random_text(10_000, seed=4)
    sa = oracle.sa_build(text)
    bad = sa.copy()
    bad[[100, 101]] = bad[[101, 100]]
    assert sst.SaNaive.from_parts(text, sa).check() == 0
    with pytest.raises(sst.SstError):  # an uploaded array is validated like the reference's build asserts (sa_search.rs:36-38)
        sst.SaNaive.from_parts(text, bad)
    sst.set_option("SA_VALIDATE", 0)   # trusted caller: no check at upload, the handle's own check still reports it
    assert sst.SaNaive.from_parts(text, bad).check() > 0


@pytest.mark.parametrize("lanes,pivot_levels", [("1", "20"), ("1", "6"), ("1", "0"), ("32", "20"), ("8", "20"), ("4", "20"), ("2", "20")])
def test_sa_search_parity(gpu, oracle, lanes, pivot_levels, monkeypatch):
    """lanes=1: thread-per-pattern kernel with the pivot-prefix table (all / some / no levels from the table);
    lanes>1: sub-warp kernels (32 = warp per pattern)."""
    sst = gpu
    gpu.set_option("SA_LANES", int(lanes))
    gpu.set_option("SA_PIVOT_LEVELS", int(pivot_levels))
    gpu.set_option("SA_KMER", 0)  # the pivot-prefix table serves every level (as for texts over a byte alphabet)
    text = random_text(200_000, seed=11)
    s = sst.SaNaive.build(text)
    sa = s.sa
    rng = np.random.default_rng(12)
    pats = random_patterns(text, 3000, seed=13)                      # util.rs:18-26
    pats += [text[i : i + 32].tobytes() for i in rng.integers(0, text.size - 200, 500)]  # config C3: 32-mers
    pats += [bytes(rng.integers(0, 4, int(l), dtype=np.uint8)) for l in rng.integers(1, 40, 500)]  # mostly absent
    pats += [b"", bytes([3] * 64), bytes([0] * 64), bytes([0]), bytes([3]), bytes([4]), bytes([255] * 3),
             text[-1:].tobytes(), text[-7:].tobytes(), text[-40:].tobytes(), text[-40:].tobytes() + b"\x00",
             text[:150].tobytes(), text.tobytes()[:1000]]
    _check_search(sst, oracle, s, text, sa, pats)


@pytest.mark.parametrize("pivot_levels", ["20", "4"])
def test_sa_search_repetitive_text(gpu, oracle, pivot_levels, monkeypatch):
    """Long LCPs: all-equal text and tandem repeats (many occurrences -> hi - lo large)."""
    sst = gpu
    gpu.set_option("SA_PIVOT_LEVELS", int(pivot_levels))
    gpu.set_option("SA_KMER", int("0" if pivot_levels == "20" else "1"))  # once through the pivot table, once through the k-mer table
    for text in (np.zeros(5000, np.uint8), np.tile(random_text(13, seed=5), 700)):
        s = sst.SaNaive.build(text)
        sa = s.sa
        assert np.array_equal(sa, oracle.sa_build(text))
        pats = [text[i : i + l].tobytes() for i, l in [(0, 1), (0, 13), (5, 64), (100, 300), (4000, 999), (4990, 10), (4990, 11)]]
        pats += [bytes([1]), bytes([0] * 6000)]
        _check_search(sst, oracle, s, text, sa, pats)


def test_sa_zero_bytes_and_short_suffixes(gpu, oracle):
    """Text full of zero bytes and patterns that end in zeros: the pivot table's zero padding must never decide wrongly."""
    sst = gpu
    rng = np.random.default_rng(77)
    text = (rng.integers(0, 3, 30_000) * (rng.random(30_000) < 0.5)).astype(np.uint8)
    text[-20:] = 0
    s = sst.SaNaive.build(text)
    sa = s.sa
    assert np.array_equal(sa, oracle.sa_build(text))
    tail = text.tobytes()
    pats = [tail[-k:] for k in range(1, 40)] + [tail[-k:] + b"\x00" * z for k in (1, 5, 17) for z in (1, 3, 16, 20)]
    pats += [tail[-k:] + b"\x01" for k in (1, 2, 15, 16, 17)] + [b"\x00" * k for k in (1, 2, 15, 16, 17, 19, 20, 21, 33)]
    pats += random_patterns(text, 1000, seed=78, lo=1, hi=50)
    _check_search(sst, oracle, s, text, sa, pats)


def test_sa_host_pipeline_chunks(gpu, oracle, monkeypatch):
    """Host path in many small chunks (ring of three staging buffers, shifted pattern base per chunk)."""
    sst = gpu
    gpu.set_option("SA_CHUNK", 37)
    text = random_text(80_000, seed=41)
    s = sst.SaNaive.build(text)
    pats = random_patterns(text, 1000, seed=42, lo=1, hi=90) + [b"", b"", text[-3:].tobytes()]
    _check_search(sst, oracle, s, text, s.sa, pats)
    gpu.set_option("SA_CHUNK", 1000000)
    _check_search(sst, oracle, s, text, s.sa, pats)


def test_sa_byte_alphabet(gpu, oracle):
    sst = gpu
    text = random_text(50_000, seed=21, sigma=256)
    s = sst.SaNaive.build(text)
    pats = random_patterns(text, 2000, seed=22, lo=1, hi=20) + [bytes([255, 255]), bytes([0, 0]), bytes([128])]
    _check_search(sst, oracle, s, text, s.sa, pats)


def test_sa_reference_return_value(gpu, oracle):
    """binary_search returns sa[l] (sa_search.rs:111): the text position of the lower-bound suffix."""
    sst = gpu
    text = random_text(100_000, seed=31)
    s = sst.SaNaive.build(text)
    q = text[5000:5040].tobytes()
    pos = s.binary_search(q)
    assert text[pos : pos + 40].tobytes() == q


@pytest.mark.parametrize("levels", ["3", "9", "18"])
def test_sa_search_sorted_order(gpu, oracle, levels, monkeypatch):
    """Opt-in reordered batch (SST_SA_SORT_MIN): coarse pass -> radix sort -> search in sorted order; identical outputs,
    including duplicate patterns, absent patterns and patterns of very different lengths."""
    sst = gpu
    gpu.set_option("SA_SORT_MIN", 1)
    gpu.set_option("SA_SORT_LEVELS", int(levels))
    gpu.set_option("SA_KMER", 0)
    text = random_text(300_000, seed=21)
    sa = oracle.sa_build(text)
    s = sst.SaNaive.from_parts(text, sa)
    pats = random_patterns(text, 5000, seed=22, lo=1, hi=120)
    rng = np.random.default_rng(23)
    pats += [rng.integers(0, 4, int(rng.integers(1, 60)), dtype=np.uint8).tobytes() for _ in range(1500)]  # mostly absent
    pats += pats[:700]  # duplicates
    pats += [bytes([3] * 40), bytes([0]), bytes([3]), text[-5:].tobytes(), text[:50].tobytes()]
    _check_search(sst, oracle, s, text, sa, pats)


@pytest.mark.parametrize("inline_bases", ["32", "15"])
@pytest.mark.parametrize("n,k", [(200_000, "15"), (200_000, "4"), (5000, "15"), (70_000, "7"), (120_000, "force12"), (90_000, "force16")])
def test_sa_search_kmer_table(gpu, oracle, n, k, inline_bases, monkeypatch):
    """Texts over {0,1,2,3} get a k-mer table (kmer[x] = lower bound of the k-base string x): the search starts in
    [kmer[x], kmer[x+1]).  Patterns shorter than k, patterns with a byte outside the alphabet (fall back to the pivot table),
    patterns made of the text's tail (proper prefixes of padded k-mers), absent patterns, the empty pattern."""
    sst = gpu
    # 8-byte {sa, 15 bases} entries (what an index takes when memory is short) or 16-byte {sa, 48 bases}
    gpu.set_option("SA_INLINE", int(inline_bases))
    if k.startswith("force"):  # deeper than one suffix per cell; 16 = the 3 Gbp configuration's depth (2^32 + 1 cells, 64-bit cell index)
        gpu.set_option("SA_KMER_FORCE", int(k[5:]))
    else:
        gpu.set_option("SA_KMER_K", int(k))
    text = random_text(n, seed=n + 5)
    s = sst.SaNaive.build(text)
    sa = s.sa
    assert np.array_equal(sa, oracle.sa_build(text))
    rng = np.random.default_rng(n)
    tail = text.tobytes()
    pats = random_patterns(text, 3000, seed=n + 6, lo=1, hi=100)
    pats += [bytes(rng.integers(0, 4, int(l), dtype=np.uint8)) for l in rng.integers(1, 40, 1500)]      # mostly absent
    pats += [tail[-j:] for j in range(1, 40)] + [tail[-j:] + b"\x00" * z for j in (1, 3, 9) for z in (1, 2, 7)]
    pats += [b"", bytes([0]), bytes([3]), bytes([3] * 64), bytes([0] * 64), bytes([4]), bytes([1, 2, 200, 3]), bytes([255] * 3),
             bytes([3] * 3) + bytes([4]), tail[:150], bytes([0, 1, 2, 3] * 4)]
    # inlined bases ({sa, the 15 bases after the first k}): patterns with a byte outside the alphabet after the first k
    # (no inline compare for them), patterns of exactly k + 14 / k + 15 / k + 16 bases, long matches (equal codes -> text)
    head = text.tobytes()
    for at in (10, 17, 25, 31, 40):
        pats += [head[100:100 + at] + bytes([7]) + head[101 + at:160], head[3000:3000 + at] + bytes([4])]
    for ln in range(16, 56):
        pats += [head[700:700 + ln], head[701:701 + ln - 1] + bytes([(head[701 + ln - 1] + 1) & 3])]
    for at in (33, 36, 44, 47, 48, 50):  # a byte outside the alphabet in the third 16-byte window
        pats += [head[900:900 + at] + bytes([9]) + head[901 + at:970], head[900:900 + at] + bytes([200])]
    # long patterns (the text compare behind the inlined bases): matches and near-misses of 60 .. 130 bases, a byte outside the
    # alphabet in the fourth .. eighth 16-byte window, patterns that run into the text's end
    for ln in list(range(60, 131, 3)) + [kk + d for kk in (4, 7, 12, 15, 16) for d in (94, 95, 96, 97)]:
        if ln + 1200 < n:
            pats += [head[1200:1200 + ln], head[1201:1201 + ln - 1] + bytes([(head[1201 + ln - 1] + 1) & 3]),
                     head[1202:1202 + ln // 2] + bytes([(head[1202 + ln // 2] + 2) & 3]) + head[1203 + ln // 2:1202 + ln]]
    for at in (52, 63, 64, 79, 95, 96, 100, 111, 112, 120):
        pats += [head[1500:1500 + at] + bytes([9]) + head[1501 + at:1640], head[1500:1500 + at] + bytes([200])]
    pats += [tail[-j:] for j in range(90, 135, 4)] + [tail[-j:-3] for j in range(90, 135, 7)]
    _check_search(sst, oracle, s, text, sa, pats)
    gpu.set_option("SA_USE_PACKED_TEXT", 0)  # compares behind the inlined bases on the byte text instead of the 2-bit packed one
    _check_search(sst, oracle, s, text, sa, pats)
    gpu.set_option("SA_USE_PACKED_TEXT", 1)
    gpu.set_option("SA_USE_CELLS", 0)   # without the packed 64-byte cells (range + first five entries in one line)
    _check_search(sst, oracle, s, text, sa, pats[:3000] + pats[-200:])
    gpu.set_option("SA_USE_CELLS", 1)
    gpu.set_option("SA_USE_INLINE", 0)  # k-mer table, probes on the text
    _check_search(sst, oracle, s, text, sa, pats[:3000] + pats[-60:])
    gpu.set_option("SA_USE_INLINE", 1)
    # the same through the pivot table only
    gpu.set_option("SA_USE_KMER", 0)
    _check_search(sst, oracle, s, text, sa, pats[:2000])


@pytest.mark.parametrize("n,k", [(300_000, 7), (300_000, 9), (50_000, 8), (2_000_000, 10)])
def test_sa_packed_cells(gpu, oracle, n, k):
    """Packed k-mer cells: patterns of k .. k + 31 bases are answered from one 64-byte line {start, 5 x {sa, 32 bases}}.  Shallow
    tables (many suffixes per cell: most cells flagged), deep ones (most cells with 0..5 suffixes), repeats (cells with many
    equal entries: hi - lo up to 5), absent patterns that fall behind every entry of their cell, the text's tail."""
    sst = gpu
    gpu.set_option("SA_KMER_K", k)
    rng = np.random.default_rng(n + k)
    text = random_text(n, seed=n + k)
    text[1000:1200] = np.tile(text[1000:1050], 4)  # a few repeated 50-mers
    s = sst.SaNaive.build(text)
    sa = s.sa
    head = text.tobytes()
    pats = [head[i:i + int(l)] for i, l in zip(rng.integers(0, n - 100, 4000), rng.integers(k, k + 34, 4000))]          # occur
    pats += [bytes(rng.integers(0, 4, int(l), dtype=np.uint8)) for l in rng.integers(k, k + 33, 3000)]                     # mostly absent
    pats += [head[i:i + k] + bytes([3] * int(l)) for i, l in zip(rng.integers(0, n - 100, 500), rng.integers(0, 30, 500))]  # behind their cell's entries
    pats += [head[i:i + k] + bytes([0] * int(l)) for i, l in zip(rng.integers(0, n - 100, 500), rng.integers(0, 30, 500))]  # before them
    pats += [head[1000:1000 + l] for l in range(k, 50)] + [head[-j:] for j in range(k, k + 40)]
    _check_search(sst, oracle, s, text, sa, pats)
    gpu.set_option("SA_USE_CELLS", 0)
    flat, off = sst.pack_patterns(pats)
    a = s.search(flat, off)
    gpu.set_option("SA_USE_CELLS", 1)
    b = s.search(flat, off)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_sa_search_capped_grid(gpu, oracle):
    """The search kernel makes one block per 128 patterns by default; with a capped grid (option SA_GRID = blocks per SM) every
    thread walks several patterns with a grid stride.  Same results either way, with and without the packed cells."""
    sst = gpu
    text = random_text(150_000, seed=91)
    s = sst.SaNaive.build(text)
    pats = random_patterns(text, 40_000, seed=92, lo=1, hi=110)
    rng = np.random.default_rng(93)
    pats += [bytes(rng.integers(0, 4, int(l), dtype=np.uint8)) for l in rng.integers(1, 60, 5000)]
    flat, off = sst.pack_patterns(pats)
    want = s.search(flat, off)
    _check_search(sst, oracle, s, text, s.sa, pats[:3000])
    for cap in (1, 3):
        gpu.set_option("SA_GRID", cap)
        got = s.search(flat, off)
        assert all(np.array_equal(a, b) for a, b in zip(want, got)), cap
        got = s.search(flat, off, sst.SA_MLR)
        assert all(np.array_equal(a, b) for a, b in zip(want, got)), cap
    gpu.set_option("SA_GRID", 0)
