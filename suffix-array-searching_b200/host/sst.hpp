// sst.hpp -- C++ host-side mirror of the reference's operator API, header-only over the C ABI
// (include/sst_b200.h).  The reference is Rust and its toolchain is absent from the build image,
// so this header is the compiled-language stand-in for the Rust shim in ../rust/: same names,
// argument meaning and error behaviour as
//   trait SearchIndex / SearchScheme, batched, full      static-search-tree/src/lib.rs:30-107
//   STree<B,16>::new / new_params / search               static-search-tree/src/s_tree.rs:47-59,72,196
//   PartitionedSTree<16,16,Tp>::new / try_new / search   static-search-tree/src/partitioned_s_tree.rs:231-241,354-364
//   SaNaive::build, binary_search                        suffix-array-searching/src/sa_search.rs:30,98
// Reference errors are panics; here they are sst::Panic exceptions.  `try_new` returns
// std::nullopt where the reference returns None.
#pragma once

#include <array>
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "sst_b200.h"

namespace sst {

struct Panic : std::runtime_error {
    int status;
    Panic(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};
inline void panic_last() { throw Panic(sst_last_status(), std::string("sst_b200: ") + sst_last_error()); }
inline void check(int rc) { if (rc != SST_OK) panic_last(); }

constexpr uint32_t MAX = SST_MAX;  // node.rs:5

/// CPU affinity of the calling thread := the CPUs local to `device` (NUMA); the size of the set, 0 if the topology is hidden.
inline int bind_thread_to_device(int device) { return sst_bind_thread_to_device(device); }

/// trait SearchIndex (lib.rs:30-48)
class SearchIndex {
  public:
    size_t size() const { return sst_index_size_bytes(h_.get()); }
    size_t layers() const { return sst_index_layers(h_.get()); }
    size_t len() const { return sst_index_len(h_.get()); }
    /// SearchScheme::query through the default (best) kernel: values of the first key >= q.
    std::vector<uint32_t> query(const std::vector<uint32_t>& qs, int scheme = SST_SCHEME_AUTO) const {
        std::vector<uint32_t> out(qs.size());
        check(sst_query(h_.get(), qs.data(), qs.size(), out.data(), nullptr, scheme));
        return out;
    }
    /// values and sorted-array indices (the `l` of binary_search.rs:36-49)
    std::pair<std::vector<uint32_t>, std::vector<uint64_t>> query_with_index(const std::vector<uint32_t>& qs, int scheme = SST_SCHEME_AUTO) const {
        std::vector<uint32_t> v(qs.size());
        std::vector<uint64_t> i(qs.size());
        check(sst_query(h_.get(), qs.data(), qs.size(), v.data(), i.data(), scheme));
        return {std::move(v), std::move(i)};
    }
    uint32_t query_one(uint32_t q) const { return query({q})[0]; }
    uint32_t search(uint32_t q) const { return query_one(q); }  // s_tree.rs:196
    const sst_index_t* raw() const { return h_.get(); }
    /// Pre-sizes the calling thread's pipeline scratch: later device-side queries of up to nq allocate nothing.
    void reserve(size_t nq, bool want_index = false) const { check(sst_query_reserve(h_.get(), nq, want_index ? 1 : 0)); }
    /// Measures on this device from which batch size on the reordered-batch pipeline wins and makes SST_SCHEME_AUTO use it.
    size_t calibrate(size_t max_nq = (size_t)1 << 25) { size_t v = 0; check(sst_query_calibrate(h_.get(), max_nq, &v)); return v; }

  protected:
    explicit SearchIndex(sst_index_t* h) : h_(h, sst_index_free) { if (!h) panic_last(); }
    std::shared_ptr<sst_index_t> h_;  // immutable after build: shared freely across threads (`Sync`)
};

/// STree<B,16> (s_tree.rs:14-20)
template <uint32_t B>
class STree : public SearchIndex {
  public:
    static STree new_(const std::vector<uint32_t>& vals, int device = 0) { return new_params(vals, false, false, false, device); }
    static STree new_params(const std::vector<uint32_t>& vals, bool left_max, bool reverse_storage, bool full_array, int device = 0) {
        const uint32_t flags = (left_max ? SST_LEFT_MAX : 0) | (reverse_storage ? SST_REVERSE_STORAGE : 0) | (full_array ? SST_FULL_ARRAY : 0);
        return STree(sst_stree_build(vals.data(), vals.size(), B, flags, device));
    }
    /// batch::<P> .. batch_final::<P> (s_tree.rs:208-326): one fixed-size batch
    template <size_t P>
    std::array<uint32_t, P> batch(const std::array<uint32_t, P>& qb) const {
        std::array<uint32_t, P> out;
        check(sst_query(h_.get(), qb.data(), P, out.data(), nullptr, SST_SCHEME_AUTO));
        return out;
    }
    /// batch_interleave_all_128 (s_tree.rs:684-696): whole slice at once
    std::vector<uint32_t> batch_interleave_all_128(const std::vector<uint32_t>& qs) const { return query(qs); }

  private:
    explicit STree(sst_index_t* h) : SearchIndex(h) {}
};
using STree16 = STree<16>;
using STree15 = STree<15>;

/// PartitionedSTree<16,16,Tp> (partitioned_s_tree.rs:19-98)
template <int VARIANT>
class PartitionedSTree : public SearchIndex {
  public:
    static PartitionedSTree new_(const std::vector<uint32_t>& vals, uint32_t b, int device = 0) {
        return PartitionedSTree(sst_pstree_build(vals.data(), vals.size(), b, VARIANT, device));
    }
    static std::optional<PartitionedSTree> try_new(const std::vector<uint32_t>& vals, uint32_t b, int device = 0) {
        sst_index_t* h = sst_pstree_build(vals.data(), vals.size(), b, VARIANT, device);
        if (!h && sst_last_status() == SST_ERR_CAPACITY) return std::nullopt;  // reference: None
        return PartitionedSTree(h);
    }

  private:
    explicit PartitionedSTree(sst_index_t* h) : SearchIndex(h) {}
};
using PartitionedSTree16 = PartitionedSTree<SST_SIMPLE>;
using PartitionedSTree16C = PartitionedSTree<SST_COMPACT>;
using PartitionedSTree16L = PartitionedSTree<SST_L1>;
using PartitionedSTree16O = PartitionedSTree<SST_OVERLAPPING>;
using PartitionedSTree16M = PartitionedSTree<SST_MAP>;

/// Eytzinger (eytzinger.rs:9-89): baseline layout; unsigned compares; 0xffffffff when q is above every key
class Eytzinger : public SearchIndex {
  public:
    static Eytzinger new_(const std::vector<uint32_t>& vals, int device = 0) { return Eytzinger(sst_eytzinger_build(vals.data(), vals.size(), device)); }
    std::vector<uint32_t> vals() const {  // the `vals` field: n + 1 entries, entry 0 = u32::MAX
        std::vector<uint32_t> out(sst_index_image_words(h_.get()));
        check(sst_index_image(h_.get(), out.data()));
        return out;
    }

  private:
    explicit Eytzinger(sst_index_t* h) : SearchIndex(h) {}
};

/// read_fasta_file's decoding (sas/util.rs:144-169) and the --human k-mer keys (sst/bin/bench.rs:60-76)
inline std::vector<uint8_t> read_fasta(const std::string& fasta, int device = 0) {
    std::vector<uint8_t> out(fasta.size() ? fasta.size() : 1);
    size_t n = 0;
    check(sst_fasta_encode(fasta.data(), fasta.size(), out.data(), &n, device));
    out.resize(n);
    return out;
}
inline std::vector<uint32_t> kmer_keys(const std::vector<uint8_t>& codes, uint32_t k = 16, bool sort = true, int device = 0) {
    std::vector<uint32_t> out(codes.size() >= k ? codes.size() - k + 1 : 1);
    size_t n = 0;
    check(sst_kmer_keys(codes.data(), codes.size(), k, out.size(), out.data(), &n, sort ? 1 : 0, device));
    out.resize(n);
    return out;
}

/// trait SearchScheme<I> + adapters (lib.rs:51-107)
template <class I>
struct SearchScheme {
    std::function<std::vector<uint32_t>(const I&, const std::vector<uint32_t>&)> query;
    uint32_t query_one(const I& index, uint32_t q) const { return query(index, {q})[0]; }
};
/// batched::<P>(f): asserts no remainder exactly like lib.rs:85-92
template <size_t P, class I, class F>
SearchScheme<I> batched(F f) {
    return {[f](const I& index, const std::vector<uint32_t>& qs) {
        if (qs.size() % P != 0) throw Panic(SST_ERR_ARG, "For now, batched queries cannot handle leftovers");
        std::vector<uint32_t> out;
        out.reserve(qs.size());
        for (size_t i = 0; i < qs.size(); i += P) {
            std::array<uint32_t, P> qb;
            std::copy(qs.begin() + i, qs.begin() + i + P, qb.begin());
            auto r = f(index, qb);
            out.insert(out.end(), r.begin(), r.end());
        }
        return out;
    }};
}
/// full(f) (lib.rs:97-107)
template <class I, class F>
SearchScheme<I> full(F f) { return {[f](const I& index, const std::vector<uint32_t>& qs) { return f(index, qs); }}; }

/// Replicas on several GPUs, queries sharded contiguously (bench.rs:558-573)
class MultiIndex {
  public:
    static MultiIndex stree(const std::vector<uint32_t>& vals, const std::vector<int>& devices, bool left_max = false) {
        return MultiIndex(sst_multi_stree_build(vals.data(), vals.size(), 16, left_max ? SST_LEFT_MAX : 0, devices.data(), (int)devices.size()));
    }
    std::vector<uint32_t> query(const std::vector<uint32_t>& qs) const {
        std::vector<uint32_t> out(qs.size());
        check(sst_multi_query(h_.get(), qs.data(), qs.size(), out.data(), nullptr, SST_SCHEME_AUTO));
        return out;
    }

  private:
    explicit MultiIndex(sst_multi_t* h) : h_(h, sst_multi_free) { if (!h) panic_last(); }
    std::shared_ptr<sst_multi_t> h_;
};

/// SaNaive (sa_search.rs:11-57) and the free search functions (sa_search.rs:98-112)
class SaNaive {
  public:
    static SaNaive build(const std::vector<uint8_t>& t, int device = 0) { return SaNaive(sst_sa_build(t.data(), t.size(), device), t.size()); }
    size_t len() const { return n_; }
    /// number of adjacent suffix pairs out of order (the assertion of sa_search.rs:36-38 counts 0)
    uint64_t check_order() const { uint64_t v = 0; check(sst_sa_check(h_.get(), &v)); return v; }
    std::vector<uint32_t> sa() const { std::vector<uint32_t> out(n_); check(sst_sa_get(h_.get(), out.data())); return out; }
    struct Hit { uint32_t lo, hi, pos; };
    std::vector<Hit> search(const std::vector<std::vector<uint8_t>>& pats, int mode = SST_SA_BINARY) const {
        std::vector<uint8_t> flat;
        std::vector<uint64_t> off{0};
        for (auto& p : pats) { flat.insert(flat.end(), p.begin(), p.end()); off.push_back(flat.size()); }
        std::vector<uint32_t> lo(pats.size()), hi(pats.size()), pos(pats.size());
        check(sst_sa_search(h_.get(), flat.data(), off.data(), pats.size(), mode, lo.data(), hi.data(), pos.data()));
        std::vector<Hit> out(pats.size());
        for (size_t i = 0; i < pats.size(); i++) out[i] = {lo[i], hi[i], pos[i]};
        return out;
    }
    const sst_sa_t* raw() const { return h_.get(); }

  private:
    SaNaive(sst_sa_t* h, size_t n) : h_(h, sst_sa_free), n_(n) { if (!h) panic_last(); }
    std::shared_ptr<sst_sa_t> h_;
    size_t n_;
};
/// binary_search(sa, q, cnt) -> sa[l]  (sa_search.rs:98-112)
inline size_t binary_search(const SaNaive& sa, const std::vector<uint8_t>& q) { return sa.search({q})[0].pos; }
/// the same with the reference's probe counter: *cnt advances by the iterations of `while l < r` (sa_search.rs:101-110)
inline size_t binary_search(const SaNaive& sa, const std::vector<uint8_t>& q, size_t* cnt) {
    const uint64_t off[2] = {0, q.size()};
    uint32_t pos = 0, probes = 0;
    check(sst_sa_search_probes(sa.raw(), q.data(), off, 1, &pos, &probes));
    if (cnt) *cnt += probes;
    return pos;
}

/// Text + suffix array replicated on several GPUs, the pattern batch sharded contiguously (chunk = ceil(npat / G)): the
/// serial callers of sa_search.rs:423-451 in the harness shape of bench.rs:558-573.
class MultiSa {
  public:
    static MultiSa build(const std::vector<uint8_t>& t, const std::vector<int>& devices) {
        return MultiSa(sst_multi_sa_build(t.data(), t.size(), devices.data(), (int)devices.size()));
    }
    static MultiSa from_parts(const std::vector<uint8_t>& t, const std::vector<uint32_t>& sa, const std::vector<int>& devices) {
        return MultiSa(sst_multi_sa_from_parts(t.data(), t.size(), sa.data(), devices.data(), (int)devices.size()));
    }
    int devices() const { return sst_multi_sa_devices(h_.get()); }
    std::vector<SaNaive::Hit> search(const std::vector<std::vector<uint8_t>>& pats, int mode = SST_SA_BINARY) const {
        std::vector<uint8_t> flat;
        std::vector<uint64_t> off{0};
        for (auto& p : pats) { flat.insert(flat.end(), p.begin(), p.end()); off.push_back(flat.size()); }
        std::vector<uint32_t> lo(pats.size()), hi(pats.size()), pos(pats.size());
        check(sst_multi_sa_search(h_.get(), flat.data(), off.data(), pats.size(), mode, lo.data(), hi.data(), pos.data()));
        std::vector<SaNaive::Hit> out(pats.size());
        for (size_t i = 0; i < pats.size(); i++) out[i] = {lo[i], hi[i], pos[i]};
        return out;
    }

  private:
    explicit MultiSa(sst_multi_sa_t* h) : h_(h, sst_multi_sa_free) { if (!h) panic_last(); }
    std::shared_ptr<sst_multi_sa_t> h_;
};

/// Library options (the table of csrc/common.cuh; read from SST_<NAME> once at load time, then only through these)
inline void set_option(const std::string& name, long long v) { check(sst_set_option(name.c_str(), v)); }
inline long long get_option(const std::string& name) { long long v = 0; check(sst_get_option(name.c_str(), &v)); return v; }

}  // namespace sst
