"""GPU parity tests (-m gpu) of the multi-GPU entry points on ONE device listed several times: the replica / shard logic
is the same as on several devices (index replicated, batch sharded by chunk = ceil(n / G), static-search-tree/src/bin/bench.rs:558-573).
Covers sst_multi_sa_* (replicas copied device to device), sst_multi_query_device, the tree replicas built from peer-copied
keys, the probe counter of sa_search.rs:98-112, the option table and sst_query_reserve."""
import numpy as np
import pytest

from util import MAX, gen_queries, gen_vals, random_patterns, random_text

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0, 0, 0]])
def test_multi_sa_search_matches_oracle(gpu, oracle, devices):
    sst = gpu
    text = random_text(80_000, seed=len(devices))
    pats = random_patterns(text, 2003, seed=9, lo=1, hi=90) + [b"", bytes([3] * 20), bytes([200] * 5)]
    flat, off = sst.pack_patterns(pats)
    of, oo = oracle.pack_patterns(pats)
    sa_ref = oracle.sa_build(text)
    elo, ehi, epos, _ = oracle.sa_search(text, sa_ref, of, oo)
    for m in (sst.MultiSa.build(text, devices), sst.MultiSa.from_parts(text, sa_ref, devices)):
        assert m.n_devices == len(devices)
        for mode in (sst.SA_BINARY, sst.SA_MLR):
            lo, hi, pos = m.search(flat, off, mode)
            assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(pos, epos), (len(devices), mode)
        lo, hi, pos = m.search(flat, off, sst.SA_BINARY, want_hi=False)
        assert hi is None and np.array_equal(lo, elo) and np.array_equal(pos, epos)


def test_multi_sa_fewer_patterns_than_devices(gpu, oracle):
    sst = gpu
    text = random_text(5000, seed=3)
    m = sst.MultiSa.build(text, [0, 0, 0, 0])
    pats = random_patterns(text, 2, seed=1)
    flat, off = sst.pack_patterns(pats)
    of, oo = oracle.pack_patterns(pats)
    elo, ehi, epos, _ = oracle.sa_search(text, oracle.sa_build(text), of, oo)
    lo, hi, pos = m.search(flat, off)
    assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(pos, epos)
    flat0, off0 = sst.pack_patterns([])
    lo, hi, pos = m.search(flat0, off0)
    assert lo.size == 0


def test_sa_from_parts_rejects_a_wrong_suffix_array(gpu, oracle):
    """The reference asserts strict suffix order when it builds (sa_search.rs:36-38); an uploaded array is checked the same way."""
    sst = gpu
    text = random_text(10_000, seed=4)
    sa = oracle.sa_build(text)
    bad = sa.copy()
    bad[[100, 101]] = bad[[101, 100]]
    with pytest.raises(sst.SstError) as e:
        sst.SaNaive.from_parts(text, bad)
    assert e.value.status == sst.ERR_ARG
    oob = sa.copy()
    oob[7] = text.size + 5
    with pytest.raises(sst.SstError):
        sst.SaNaive.from_parts(text, oob)
    sst.set_option("SA_VALIDATE", 0)  # trusted caller: no check (the handle's own check() still reports it)
    assert sst.SaNaive.from_parts(text, bad).check() > 0


def test_sa_probe_counter_is_the_reference_loop(gpu, oracle):
    """sst_sa_search_probes: sa[l] and the number of iterations of `while l < r` (the reference's cnt, sa_search.rs:98-112)."""
    sst = gpu
    text = random_text(30_000, seed=6)
    s = sst.SaNaive.build(text)
    sa = s.sa
    pats = random_patterns(text, 300, seed=2, lo=1, hi=60) + [bytes([0]), bytes([3] * 50)]
    flat, off = sst.pack_patterns(pats)
    pos, probes = s.search_probes(flat, off)
    n = text.size
    for i, q in enumerate(pats):
        l, r, cnt = 0, n, 0
        while l < r:  # sa_search.rs:101-110
            cnt += 1
            m = (l + r) // 2
            if text[sa[m]:].tobytes() < q:
                l = m + 1
            else:
                r = m
        assert probes[i] == cnt, i
        assert pos[i] == (sa[l] if l < n else 0xFFFFFFFF), i


@pytest.mark.parametrize("variant", ["plain", "map"])
def test_multi_tree_replicas_and_device_shards(gpu, oracle, variant):
    import torch

    sst = gpu
    vals = gen_vals(300_000, seed=21)
    qs = gen_queries(100_003, seed=22, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    devices = [0, 0, 0]
    m = sst.MultiIndex.stree(vals, devices, left_max=True) if variant == "plain" else sst.MultiIndex.pstree(vals, 12, sst.MAP, devices)
    v, i = m.query(qs, want_index=True)
    assert np.array_equal(v, ev) and np.array_equal(i, ei)
    # the same shards, already resident on the replicas' device
    G = len(devices)
    chunk = -(-qs.size // G)
    shards = [torch.from_numpy(qs[k * chunk:(k + 1) * chunk].view(np.int32).copy()).cuda() for k in range(G)]
    vs, is_ = m.query_device(shards, want_index=True)
    gv = np.concatenate([x.cpu().numpy().view(np.uint32) for x in vs])
    gi = np.concatenate([x.cpu().numpy().astype(np.uint64) for x in is_])
    assert np.array_equal(gv, ev) and np.array_equal(gi, ei)


def test_options_table(gpu):
    sst = gpu
    names = sst.option_names()
    assert "BK_R" in names and "SA_CHUNK" in names and len(set(names)) == len(names)
    assert sst.get_option("SST_CHUNK") == sst.get_option("CHUNK")
    with pytest.raises(sst.SstError):
        sst.set_option("THREADS", 4096)  # out of range: rejected, not applied at the next launch
    with pytest.raises(sst.SstError):
        sst.set_option("NO_SUCH_OPTION", 1)
    with sst.options(CHUNK=4096):
        assert sst.get_option("CHUNK") == 4096
    assert sst.get_option("CHUNK") == 1 << 22


def test_reserve_makes_the_pipeline_allocation_free(gpu, oracle):
    """After sst_query_reserve a pipeline call allocates nothing: it can be captured into a CUDA graph and replayed."""
    import torch

    sst = gpu
    sst.set_option("BK_MIN_N", 0)
    sst.set_option("BK_R", 256)
    vals = gen_vals(400_000, seed=31)
    t = sst.STree16.new_params(vals, True, False, False)
    qs = gen_queries(150_000, seed=32, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    d = torch.from_numpy(qs.view(np.int32)).cuda()
    t.reserve(d.numel(), want_index=True)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        v, i = t.query(d, sst.SCHEME_BUCKETED, want_index=True)  # warm-up on the capture stream (function attributes)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            v, i = t.query(d, sst.SCHEME_BUCKETED, want_index=True)
        v.zero_(); i.zero_()
        g.replay()
        stream.synchronize()
    assert np.array_equal(v.cpu().numpy().view(np.uint32), ev) and np.array_equal(i.cpu().numpy().astype(np.uint64), ei)
    # (a cudaMalloc or cudaFree inside the captured call would have invalidated the capture: torch captures in the mode that
    # forbids them, so a successful capture + replay IS the proof that the call allocated nothing)
