// test_host.cpp -- the reference's own tests, restated against the C++ host mirror:
//   s_tree.rs:861-895 (known answers), test.rs:142-260 (every index/scheme equals binary search),
//   lib.rs:85-92 (batched asserts on leftovers), plus the SA strict-order property (sa_search.rs:36-38).
// Needs a B200; exits non-zero on the first mismatch.  Run by tests/test_gpu_host.py.
#include <algorithm>
#include <cstdio>
#include <random>
#include <thread>

#include "sst.hpp"

using namespace sst;

static int failures = 0;
#define EXPECT(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)

// SortedVec::binary_search (binary_search.rs:36-49): the expectation of test.rs:200
static uint32_t sorted_vec_binary_search(const std::vector<uint32_t>& vals, uint32_t q) {
    size_t l = 0, r = vals.size();
    while (l < r) { size_t m = (l + r) / 2; if (vals[m] < q) l = m + 1; else r = m; }
    return vals[l];
}

static std::vector<uint32_t> gen_vals(size_t n, std::mt19937_64& rng) {  // util.rs:31-42
    std::vector<uint32_t> v(n);
    for (auto& x : v) x = (uint32_t)(rng() % MAX);
    v[0] = MAX;
    std::sort(v.begin(), v.end());
    return v;
}

template <class I>
static void expect_equal_to_binary_search(const I& index, const std::vector<uint32_t>& vals, const std::vector<uint32_t>& qs, const char* what) {
    auto got = index.query(qs);
    for (size_t i = 0; i < qs.size(); i++)
        if (got[i] != sorted_vec_binary_search(vals, qs[i])) { fprintf(stderr, "FAIL %s: q=%u got %u\n", what, qs[i], got[i]); failures++; return; }
}

int main() {
    if (sst_device_count() < 1) { fprintf(stderr, "no sm_100 device (no CPU fallback)\n"); return 2; }
    {   // test_bptree_search_bottom_layer / test_bptree_search_top_node
        std::vector<uint32_t> vals;
        for (uint32_t i = 1; i < 2000; i++) vals.push_back(i);
        vals.push_back(MAX);
        auto t = STree16::new_(vals);
        EXPECT(t.search(452) == 452);
        EXPECT(t.search(289) == 289);
        EXPECT(t.layers() == 3);
    }
    std::mt19937_64 rng(7);
    for (size_t size : {size_t(64), size_t(80), size_t(1 << 12), size_t(7 << 14), size_t(5 << 18)}) {  // sizes in bytes as test.rs:146-153
        auto vals = gen_vals(size / 4, rng);
        std::vector<uint32_t> qs(1024);  // 1000.next_multiple_of(128)
        for (auto& q : qs) q = (uint32_t)(rng() % MAX);
        expect_equal_to_binary_search(STree16::new_(vals), vals, qs, "STree16");
        expect_equal_to_binary_search(STree15::new_(vals), vals, qs, "STree15");
        expect_equal_to_binary_search(STree16::new_params(vals, true, false, false), vals, qs, "STree16 left_max");
        expect_equal_to_binary_search(STree16::new_params(vals, true, false, true), vals, qs, "STree16 left_max full");
        for (uint32_t b : {0u, 4u, 8u, 16u, 20u}) {
            expect_equal_to_binary_search(PartitionedSTree16::new_(vals, b), vals, qs, "psp");
            expect_equal_to_binary_search(PartitionedSTree16C::new_(vals, b), vals, qs, "pspc");
            expect_equal_to_binary_search(PartitionedSTree16L::new_(vals, b), vals, qs, "pspl");
            expect_equal_to_binary_search(PartitionedSTree16O::new_(vals, b), vals, qs, "pspo");
            expect_equal_to_binary_search(PartitionedSTree16M::new_(vals, b), vals, qs, "pspm");
        }
        // SearchScheme adapters: batched::<128>(STree16::batch_final) and full(batch_interleave_all_128)
        auto idx = STree16::new_params(vals, true, false, false);
        auto s1 = batched<128, STree16>([](const STree16& i, const std::array<uint32_t, 128>& qb) { return i.batch<128>(qb); });
        auto s2 = full<STree16>([](const STree16& i, const std::vector<uint32_t>& q) { return i.batch_interleave_all_128(q); });
        EXPECT(s1.query(idx, qs) == s2.query(idx, qs));
        bool threw = false;
        try { s1.query(idx, std::vector<uint32_t>(100, 1)); } catch (const Panic&) { threw = true; }
        EXPECT(threw);
        // `Sync`: concurrent queries on one index from several host threads (bench.rs:558-573)
        std::vector<std::vector<uint32_t>> outs(4);
        std::vector<std::thread> th;
        for (int k = 0; k < 4; k++) th.emplace_back([&, k] { outs[k] = idx.query(qs); });
        for (auto& x : th) x.join();
        for (int k = 1; k < 4; k++) EXPECT(outs[k] == outs[0]);
        EXPECT(MultiIndex::stree(vals, {0, 0}, true).query(qs) == outs[0]);
    }
    {   // eytzinger.rs:199-229 golden vectors
        std::vector<uint32_t> in15, in10;
        for (uint32_t i = 1; i <= 15; i++) in15.push_back(i);
        for (uint32_t i = 0; i < 10; i++) in10.push_back(i);
        const uint32_t M = 0xffffffffu;
        EXPECT((Eytzinger::new_(in15).vals() == std::vector<uint32_t>{M, 8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15}));
        auto e = Eytzinger::new_(in10);
        EXPECT((e.vals() == std::vector<uint32_t>{M, 6, 3, 8, 1, 5, 7, 9, 0, 2, 4}));
        EXPECT(e.search(3) == 3 && e.search(12) == M);
    }
    {   // read_fasta_file + --human keys feeding the tree
        auto codes = read_fasta(">r1 x\nACGT\nacgtN\n>r2\nTT\n");
        EXPECT((codes == std::vector<uint8_t>{0, 1, 2, 3, 0, 1, 2, 3, 0, 3, 3}));
        auto keys = kmer_keys(codes, 3, true);
        EXPECT(keys.size() == 9 && std::is_sorted(keys.begin(), keys.end()) && keys.back() == MAX);
        EXPECT(STree16::new_(keys).search(0) == keys[0]);
    }
    {   // panics of the reference become exceptions
        bool threw = false;
        try { STree16::new_({3, 2, 1}); } catch (const Panic&) { threw = true; }
        EXPECT(threw);
        threw = false;
        try { STree16::new_({1, 2, 0x80000000u}); } catch (const Panic&) { threw = true; }  // s_tree.rs:87-89
        EXPECT(threw);
    }
    {   // suffix array: strict order (sa_search.rs:36-38) and sa[l] of binary_search (sa_search.rs:98-112)
        std::vector<uint8_t> t(100000);
        for (auto& c : t) c = (uint8_t)(rng() % 4);  // util.rs:9-15
        auto sa = SaNaive::build(t);
        EXPECT(sa.check_order() == 0);
        std::vector<uint8_t> q(t.begin() + 777, t.begin() + 777 + 40);
        size_t pos = binary_search(sa, q);
        EXPECT(std::equal(q.begin(), q.end(), t.begin() + pos));
        auto hits = sa.search({q}, SST_SA_MLR);
        EXPECT(hits[0].pos == pos && hits[0].hi > hits[0].lo);
        // the reference's probe counter: iterations of `while l < r` over [0, n) = floor(log2 n) or one more (sa_search.rs:98-112)
        size_t cnt = 0;
        EXPECT(binary_search(sa, q, &cnt) == pos && cnt >= 16 && cnt <= 17);
        // replicas (one device listed three times), patterns sharded by chunk = ceil(npat / G)
        auto msa = MultiSa::build(t, {0, 0, 0});
        std::vector<std::vector<uint8_t>> pats;
        for (int k = 0; k < 50; k++) pats.emplace_back(t.begin() + 1000 * k, t.begin() + 1000 * k + 20 + k);
        auto one = sa.search(pats), many = msa.search(pats, SST_SA_MLR);
        EXPECT(msa.devices() == 3 && one.size() == many.size());
        for (size_t k = 0; k < one.size(); k++) EXPECT(one[k].lo == many[k].lo && one[k].hi == many[k].hi && one[k].pos == many[k].pos);
    }
    {   // options are validated; calibrate is a no-op on an index the pipeline does not serve
        bool threw = false;
        try { set_option("SA_LANES", 0); } catch (const Panic&) { threw = true; }
        EXPECT(threw && get_option("SA_LANES") == 1);
        auto small = STree16::new_({1, 5, 9, MAX});
        EXPECT(small.calibrate() == 0);
        small.reserve(1000);
    }
    if (failures) { fprintf(stderr, "%d failure(s)\n", failures); return 1; }
    printf("host mirror OK\n");
    return 0;
}
