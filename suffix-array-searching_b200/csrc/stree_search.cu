// stree_search.cu -- batched lower_bound kernels over the S+-tree image (the hot path).
//
// Replaces (reference paths relative to static-search-tree/src):
//   BTreeNode::find_popcnt / find_splat / find_splat64    node.rs:93-138
//   STree::search, batch_*, batch_final, batch_interleave_*   s_tree.rs:196-832
//   PartitionedSTree::<..>::search (all five layouts)     partitioned_s_tree.rs:654-880
//
// The reference hides DRAM latency by keeping 128 queries in flight per CPU thread with
// software prefetch.  On B200 the same descent is mapped as follows:
//   * A node (64 B = 2 sectors) is fetched by a G-lane subgroup in ONE load instruction so that
//     the L1TEX tag stage sees one line per query per level: G=4 lanes x 16 B (LDG.128),
//     G=2 lanes x 32 B (LDG.256, new on sm_100) or G=16 lanes x 4 B (the ballot/popc mapping).
//   * Each lane group keeps G*T independent descents in flight (ILP); with 32 resident warps per
//     SM that is >= 1000 outstanding 64-B requests per SM, enough for HBM latency.
//   * The top levels of the plain tree are replaced by a rank table (bucket table over the top 15
//     key bits + low 16 bits of every separator of the first global level), staged into shared
//     memory once per CTA by 1-D TMA bulk copies (cp.async.bulk + mbarrier).  Copying the top
//     NODES into shared memory instead was measured slower (profiles/r1_tree_sweep1.log).
//   * Lower internal levels are loaded with an L2 evict_last policy, the leaf level (the only
//     HBM-resident one at 2^28 keys) with evict_first and no L1 allocation, so that leaf
//     traffic does not push the last internal level out of the 126 MB L2.
//   * Queries and results move as one coalesced 128-B line per warp.
// All arithmetic is integer; results are bit-exact with the reference (signed compares as in
// node.rs:91-108, flat leaf read that may spill into the next node as in s_tree.rs:322-325).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <vector>

#include "common.cuh"

namespace sst {
namespace {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// Per-lane slice of a node: W = 16 / G consecutive keys.
template <int W>
struct Keys {
    uint32_t k[W];
};

// Node loads carry the L2::64B prefetch-size qualifier (SASS ...LTC64B): without it B200 fills L2 from
// HBM in 128-byte units, i.e. every random 64-byte node drags its neighbour along (measured with the
// gather probe: 11.7 GB vs 6.1 GB of DRAM reads per 10^8 gathers, profiles/r1_ncu_probe_pf.csv).
template <int W, bool NO_L1>
__device__ __forceinline__ Keys<W> ldg_keys(const uint32_t* p, uint64_t pol) {
    Keys<W> r;
    if constexpr (W == 1) {
        if constexpr (NO_L1)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.u32 %0, [%1], %2;" : "=r"(r.k[0]) : "l"(p), "l"(pol));
        else
            asm volatile("ld.global.nc.L2::cache_hint.L2::64B.u32 %0, [%1], %2;" : "=r"(r.k[0]) : "l"(p), "l"(pol));
    } else if constexpr (W == 2) {
        if constexpr (NO_L1)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.k[0]), "=r"(r.k[1]) : "l"(p), "l"(pol));
        else
            asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.k[0]), "=r"(r.k[1]) : "l"(p), "l"(pol));
    } else if constexpr (W == 4) {
        if constexpr (NO_L1)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(r.k[0]), "=r"(r.k[1]), "=r"(r.k[2]), "=r"(r.k[3]) : "l"(p), "l"(pol));
        else
            asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(r.k[0]), "=r"(r.k[1]), "=r"(r.k[2]), "=r"(r.k[3]) : "l"(p), "l"(pol));
    } else {
        static_assert(W == 8, "unsupported slice width");
        if constexpr (NO_L1)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                         : "=r"(r.k[0]), "=r"(r.k[1]), "=r"(r.k[2]), "=r"(r.k[3]), "=r"(r.k[4]), "=r"(r.k[5]), "=r"(r.k[6]), "=r"(r.k[7])
                         : "l"(p), "l"(pol));
        else
            asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                         : "=r"(r.k[0]), "=r"(r.k[1]), "=r"(r.k[2]), "=r"(r.k[3]), "=r"(r.k[4]), "=r"(r.k[5]), "=r"(r.k[6]), "=r"(r.k[7])
                         : "l"(p), "l"(pol));
    }
    return r;
}

// 32-byte load with an explicit L2 prefetch-size qualifier (SASS: LDG.E...LTC64B/LTC128B/LTC256B).
template <int PF>
__device__ __forceinline__ Keys<8> ldg_keys8_pf(const uint32_t* p, uint64_t pol) {
    Keys<8> r;
#define SST_LD8(Q)                                                                                                                        \
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint" Q ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"                          \
                 : "=r"(r.k[0]), "=r"(r.k[1]), "=r"(r.k[2]), "=r"(r.k[3]), "=r"(r.k[4]), "=r"(r.k[5]), "=r"(r.k[6]), "=r"(r.k[7])       \
                 : "l"(p), "l"(pol))
    if constexpr (PF == 64) SST_LD8(".L2::64B");
    else if constexpr (PF == 128) SST_LD8(".L2::128B");
    else if constexpr (PF == 256) SST_LD8(".L2::256B");
    else SST_LD8("");
#undef SST_LD8
    return r;
}

// Plain read-only load without an L2 cache hint (so that an access-policy window, if any, decides).
template <int W>
__device__ __forceinline__ Keys<W> ldg_keys_plain(const uint32_t* p) {
    Keys<W> r;
    if constexpr (W == 1) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r.k[0]) : "l"(p));
    else if constexpr (W == 2) asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.k[0]), "=r"(r.k[1]) : "l"(p));
    else if constexpr (W == 4)
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.k[0]), "=r"(r.k[1]), "=r"(r.k[2]), "=r"(r.k[3]) : "l"(p));
    else
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.k[0]), "=r"(r.k[1]), "=r"(r.k[2]), "=r"(r.k[3]), "=r"(r.k[4]), "=r"(r.k[5]), "=r"(r.k[6]), "=r"(r.k[7])
                     : "l"(p));
    return r;
}

// node.rs:93-109: number of keys < q (signed), reduced over the G lanes of the group.
template <int G>
__device__ __forceinline__ unsigned group_count(const Keys<16 / G>& ks, uint32_t q, unsigned gshift) {
    constexpr int W = 16 / G;
    if constexpr (G == 16) {
        const unsigned b = __ballot_sync(kFull, (int)ks.k[0] < (int)q);
        return __popc((b >> gshift) & 0xffffu);
    } else {
        unsigned c = 0;
#pragma unroll
        for (int i = 0; i < W; i++) c += ((int)ks.k[i] < (int)q) ? 1u : 0u;
#pragma unroll
        for (int m = 1; m < G; m <<= 1) c += __shfl_xor_sync(kFull, c, m);
        return c;
    }
}

// Key at position e (0..W-1) of this lane's slice, without dynamic register indexing.
template <int W>
__device__ __forceinline__ uint32_t pick(const Keys<W>& ks, unsigned e) {
    uint32_t v = ks.k[0];
#pragma unroll
    for (int i = 1; i < W; i++) v = (e == (unsigned)i) ? ks.k[i] : v;
    return v;
}

// ------------------------------------------------------------------------------------------------
// Fast kernel: plain S+-tree with B = 16 (any of left_max / reverse_storage / full_array).
//
// TOP = true replaces the walk through levels [0, top_level) by a rank query in shared memory.
// The in-order sequence of all keys stored in levels [0, top_level) is exactly the sorted list of
// separators between consecutive nodes of level `top_level`, and the node the reference's
// descent (s_tree.rs:196-203) reaches at that level is the number of separators < q.  The
// builder stores, per index, a bucket table over the top 15 key bits (`top_table[b]` = number of
// separators with key >> 16 < b) and the low 16 bits of every separator (`top_low`); both are
// staged into shared memory once per CTA by 1-D TMA bulk copies.  One lane answers one rank
// query with 2 + ~2 two-byte shared loads, which costs ~0.1 L1 data-pipe wavefronts per load
// instead of the 1 wavefront per 64-byte node of the node walk (ncu: the node walk ran at 84 %
// of the L1TEX data-pipe peak, profiles/r1_*).
// ------------------------------------------------------------------------------------------------
struct FastParams {
    const uint32_t* tree;
    unsigned long long level_slot[kMaxLevels];
    int levels;
    int top_level;                    // first level read from global memory (0 without TOP)
    const uint16_t* top_table;        // [2^15 + 1] (+ padding to 16 B)
    const uint16_t* top_low;          // [top_nbound] (+ padding to 16 B)
    unsigned top_nbound;
    unsigned long long leaf_slots;
    unsigned long long n;
    int hints;                        // bit0: L2 evict_last on inner levels, bit1: evict_first on leaf
    unsigned l1_levels;               // bit h set: level h is small enough to live in L1 (allocate there)
    const uint16_t* c5;               // 16-bit compressed copy of level `levels - 2` (or null)
    const uint32_t* h5;               // per node: base (first separator), 0xffffffff = read the exact node
};

constexpr unsigned kTopBuckets = 1u << 15;
constexpr unsigned kTopTableBytes = ((kTopBuckets + 1) * 2 + 15) & ~15u;

__device__ __forceinline__ unsigned top_rank(const uint16_t* __restrict__ tab, const uint16_t* __restrict__ low, uint32_t q) {
    if (q > kMax) return 0;  // signed compare of node.rs:91-108: such a q is below every key
    const unsigned b = q >> 16, ql = q & 0xffffu;
    unsigned lo = tab[b], hi = tab[b + 1];
    if (hi - lo > 8u) {  // skewed keys: many separators share the bucket
        while (lo < hi) {
            const unsigned m = (lo + hi) >> 1;
            if (low[m] < ql) lo = m + 1; else hi = m;
        }
    } else {
        while (lo < hi && low[lo] < ql) lo++;
    }
    return lo;
}

template <int G, int T, bool TOP>
__global__ void __launch_bounds__(1024, 1)
stree_search_fast(const __grid_constant__ FastParams p, const uint32_t* __restrict__ qs, size_t nq,
                  uint32_t* __restrict__ out_vals, unsigned long long* __restrict__ out_idx) {
    constexpr int W = 16 / G;   // keys per lane
    constexpr int D = G * T;    // descents in flight per lane group
    extern __shared__ __align__(128) uint16_t smem16[];
    __shared__ __align__(8) uint64_t bar;
    const uint16_t* s_tab = smem16;
    const uint16_t* s_low = smem16 + kTopTableBytes / 2;

    if constexpr (TOP) {  // stage the rank table: one elected thread issues 1-D TMA bulk copies
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned low_bytes = (p.top_nbound * 2u + 15u) & ~15u;
            mbar_expect_tx(&bar, kTopTableBytes + low_bytes);
            for (unsigned off = 0; off < kTopTableBytes; off += 32768u)
                tma_bulk_g2s((char*)smem16 + off, (const char*)p.top_table + off, min(32768u, kTopTableBytes - off), &bar);
            for (unsigned off = 0; off < low_bytes; off += 32768u)
                tma_bulk_g2s((char*)smem16 + kTopTableBytes + off, (const char*)p.top_low + off, min(32768u, low_bytes - off), &bar);
        }
        mbar_wait(&bar, 0);
    }

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const unsigned sub = lane & (G - 1), gbase = lane & ~(unsigned)(G - 1);
    const uint64_t pol_inner = (p.hints & 1) ? policy_evict_last() : policy_evict_normal();
    const uint64_t pol_leaf = (p.hints & 2) ? policy_evict_first() : policy_evict_normal();
    const int L = p.levels;
    const int h0 = TOP ? p.top_level : 0;

    for (size_t base = ((size_t)blockIdx.x * warps + warp) * (32 * T); base < nq; base += (size_t)gridDim.x * warps * (32 * T)) {
        // one coalesced line of queries per tile; lane l owns query base + t*32 + l
        uint32_t qown[T], kown[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            const size_t i = base + (size_t)t * 32 + lane;
            qown[t] = i < nq ? __ldcs(qs + i) : 0u;
            kown[t] = TOP ? top_rank(s_tab, s_low, qown[t]) : 0u;
        }
        uint32_t q[D], k[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            q[d] = __shfl_sync(kFull, qown[d / G], gbase + (d % G));
            k[d] = TOP ? __shfl_sync(kFull, kown[d / G], gbase + (d % G)) : 0u;
        }
        // ---- internal levels read from L1/L2 ----
        for (int h = h0; h + 1 < L; h++) {
            if constexpr (G == 2) {
                if (p.c5 != nullptr && h + 2 == L) {
                    // Last internal level through its 16-bit copy: 32 B per node instead of 64, so the
                    // level stays L2-resident next to the leaf stream (separator = base + 16-bit delta).
                    // Nodes spanning >= 2^16 in value (and q >= 2^31) take the exact node.
                    uint4 cw[D];
                    unsigned hh[D];
#pragma unroll
                    for (int d = 0; d < D; d++) {
                        hh[d] = __ldg(p.h5 + k[d]);
                        const uint4* cp = reinterpret_cast<const uint4*>(p.c5 + (size_t)k[d] * 16u) + sub;
                        asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                     : "=r"(cw[d].x), "=r"(cw[d].y), "=r"(cw[d].z), "=r"(cw[d].w) : "l"(cp), "l"(pol_inner));
                    }
#pragma unroll
                    for (int d = 0; d < D; d++) {
                        unsigned c;
                        if (hh[d] == 0xffffffffu || q[d] > kMax) {  // group-uniform, rare
                            const unsigned gm = 3u << gbase;
                            const Keys<W> ex = ldg_keys<W, true>(p.tree + p.level_slot[h] + sub * W + (size_t)k[d] * 16u, pol_inner);
                            c = 0;
#pragma unroll
                            for (int i = 0; i < W; i++) c += ((int)ex.k[i] < (int)q[d]) ? 1u : 0u;
                            c += __shfl_xor_sync(gm, c, 1);
                        } else {
                            // separator_i < q  <=>  delta_i < q - base  (0 deltas qualify when q <= base, all when q - base > 0xffff)
                            const uint32_t x = q[d] > hh[d] ? min(q[d] - hh[d], 0x10000u) : 0u;
                            const uint32_t w0 = cw[d].x, w1 = cw[d].y, w2 = cw[d].z, w3 = cw[d].w;
                            const unsigned cnt = ((w0 & 0xffffu) < x) + ((w0 >> 16) < x) + ((w1 & 0xffffu) < x) + ((w1 >> 16) < x) +
                                                 ((w2 & 0xffffu) < x) + ((w2 >> 16) < x) + ((w3 & 0xffffu) < x) + ((w3 >> 16) < x);
                            c = cnt + __shfl_xor_sync(3u << gbase, cnt, 1);
                        }
                        k[d] = k[d] * 17u + c;
                    }
                    continue;
                }
            }
            Keys<W> ks[D];
            const uint32_t* gl = p.tree + p.level_slot[h] + sub * W;
            if (p.hints & 4) {  // experiment: un-hinted loads, the launch's access-policy window decides
#pragma unroll
                for (int d = 0; d < D; d++) ks[d] = ldg_keys_plain<W>(gl + (size_t)k[d] * 16u);
            } else if ((p.l1_levels >> h) & 1u) {
#pragma unroll
                for (int d = 0; d < D; d++) ks[d] = ldg_keys<W, false>(gl + (size_t)k[d] * 16u, pol_inner);
            } else {  // larger than L1: do not allocate there
#pragma unroll
                for (int d = 0; d < D; d++) ks[d] = ldg_keys<W, true>(gl + (size_t)k[d] * 16u, pol_inner);
            }
#pragma unroll
            for (int d = 0; d < D; d++) k[d] = k[d] * 17u + group_count<G>(ks[d], q[d], gbase);
        }
        // ---- leaf level: s_tree.rs:322-325 ----
        {
            const uint32_t* gl = p.tree + p.level_slot[L - 1];
            Keys<W> ks[D];
            if ((p.l1_levels >> (L - 1)) & 1u) {  // small tree: the leaf level itself is L1-resident
#pragma unroll
                for (int d = 0; d < D; d++) ks[d] = ldg_keys<W, false>(gl + (size_t)k[d] * 16u + sub * W, pol_inner);
            } else {
#pragma unroll
                for (int d = 0; d < D; d++) ks[d] = ldg_keys<W, true>(gl + (size_t)k[d] * 16u + sub * W, pol_leaf);
            }
            uint32_t myval[T];
            unsigned long long myidx[T];
#pragma unroll
            for (int d = 0; d < D; d++) {
                const unsigned c = group_count<G>(ks[d], q[d], gbase);  // 0..16, uniform in the group
                // the answer sits in lane c / W of the group, element c % W
                const uint32_t cand = pick<W>(ks[d], c % W);
                uint32_t v = __shfl_sync(kFull, cand, gbase + (c < 16u ? c / W : 0u));
                const unsigned long long pos = (unsigned long long)k[d] * 16ull + c;
                if (c == 16u) v = pos < p.leaf_slots ? __ldg(gl + pos) : kMax;  // flat read spills into the next node
                if (sub == (unsigned)(d % G)) {
                    myval[d / G] = v;
                    myidx[d / G] = pos < p.n ? pos : p.n;
                }
            }
#pragma unroll
            for (int t = 0; t < T; t++) {
                const size_t i = base + (size_t)t * 32 + lane;
                if (i < nq) {
                    __stcs(out_vals + i, myval[t]);
                    if (out_idx) __stcs(out_idx + i, myidx[t]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic kernel: one thread per query, any layout (all partitioned variants, B = 15, ...).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned count4(const uint4& v, uint32_t q) {
    return ((int)v.x < (int)q) + ((int)v.y < (int)q) + ((int)v.z < (int)q) + ((int)v.w < (int)q);
}

// Number of keys < q (signed) in the 16-slot window starting at p (4-byte aligned).
__device__ __forceinline__ unsigned count16(const uint32_t* __restrict__ p, uint32_t q) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 31u) == 0) {  // node-aligned: two 256-bit loads (LDG.E.256)
        const Keys<8> lo = ldg_keys_plain<8>(p), hi = ldg_keys_plain<8>(p + 8);
        unsigned c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) c += ((int)lo.k[i] < (int)q) + ((int)hi.k[i] < (int)q);
        return c;
    }
    // unaligned window (partitioned_s_tree.rs:807,857,876 read_unaligned): five aligned 16-byte
    // chunks cover it; slots outside [p, p+16) are masked out.
    const unsigned sh = (unsigned)((a >> 2) & 3u);  // window starts at slot `sh` of the first chunk
    const uint4* c0 = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    unsigned c = 0;
    const uint4 f = __ldg(c0);
    c += (sh <= 0 && (int)f.x < (int)q) + (sh <= 1 && (int)f.y < (int)q) + (sh <= 2 && (int)f.z < (int)q) + ((int)f.w < (int)q);
#pragma unroll
    for (int i = 1; i < 4; i++) c += count4(__ldg(c0 + i), q);
    if (sh) {
        const uint4 l = __ldg(c0 + 4);
        c += (sh > 0 && (int)l.x < (int)q) + (sh > 1 && (int)l.y < (int)q) + (sh > 2 && (int)l.z < (int)q);
    }
    return c;
}

__global__ void __launch_bounds__(256)
stree_search_generic(const __grid_constant__ SstTreeView v, const uint32_t* __restrict__ qs, size_t nq,
                     uint32_t* __restrict__ out_vals, unsigned long long* __restrict__ out_idx) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = qs[i];
        const unsigned long long part = (unsigned long long)(q >> v.shift);
        uint32_t val = kMax;
        unsigned long long index = v.n;
        if (v.variant == SST_PLAIN || part < v.parts) {
            unsigned long long s = 0, pb = 0;
            switch (v.variant) {
                case SST_COMPACT: pb = part * v.part_stride; break;
                case SST_MAP: s = v.prefix_map[part]; break;
                case SST_PLAIN: break;
                default: s = part * v.start_mul; break;
            }
            const int L = v.levels;
            const bool dense_upper = v.variant == SST_COMPACT && v.upper != nullptr;
            const uint32_t* ub = dense_upper ? v.upper + part * v.upper_stride : v.tree + pb;
            for (int h = 0; h + 1 < L; h++) {
                const unsigned c = count16(ub + v.level_slot[h] + s, q);
                s = s * v.mult[h] + 16ull * c;
            }
            const uint32_t* leaf = v.tree + v.level_slot[L - 1] + pb;
            const unsigned c = count16(leaf + s, q);
            const unsigned long long pos = s + c;  // flat slot inside the leaf level (of this part for COMPACT)
            // The read below may spill into the next node (idx == 16, s_tree.rs:322-325).  Past the
            // leaf level the reference reads allocation slack; defined as MAX.
            val = pos < v.leaf_slots ? __ldg(leaf + pos) : kMax;
            if (out_idx) switch (v.variant) {
                case SST_PLAIN: index = (s >> 4) * v.node_b + c; break;
                case SST_MAP: index = pos; break;
                case SST_COMPACT: {
                    const unsigned long long st = v.part_start[part], cnt = v.part_start[part + 1] - st;
                    index = st + (pos < cnt ? pos : cnt);
                    break;
                }
                default: {
                    const unsigned long long st = v.part_start[part], cnt = v.part_start[part + 1] - st, pp = v.part_pos[part];
                    const unsigned long long off = pos > pp ? pos - pp : 0;
                    index = st + (off < cnt ? off : cnt);
                    break;
                }
            }
            if (index > v.n) index = v.n;
        }
        out_vals[i] = val;
        if (out_idx) out_idx[i] = index;
    }
}


// ------------------------------------------------------------------------------------------------
// Baseline: SortedVec::binary_search (binary_search.rs:36-49) on the GPU, one thread per query over
// the leaf level of a plain B=16 tree (which is the sorted array itself).  This is the comparison
// the reference's headline is about ("S-tree vs binary search"); not a production path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
binary_search_kernel(const uint32_t* __restrict__ vals, unsigned long long n, const uint32_t* __restrict__ qs, size_t nq,
                     uint32_t* __restrict__ out_vals, unsigned long long* __restrict__ out_idx) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = qs[i];
        unsigned long long l = 0, r = n;
        while (l < r) {
            const unsigned long long m = (l + r) >> 1;
            if (__ldg(vals + m) < q) l = m + 1; else r = m;
        }
        out_vals[i] = l < n ? __ldg(vals + l) : kMax;
        if (out_idx) out_idx[i] = l;
    }
}

// ------------------------------------------------------------------------------------------------
// Baseline: Eytzinger::search (eytzinger.rs:82-89), one thread per query, unsigned compares.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eytzinger_search_kernel(const uint32_t* __restrict__ e, unsigned long long len, int H, const uint32_t* __restrict__ qs, size_t nq,
                        uint32_t* __restrict__ out_vals, unsigned long long* __restrict__ out_idx) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = qs[i];
        unsigned long long idx = 1;
        while (idx < len) idx = 2 * idx + (q > __ldg(e + idx) ? 1ull : 0ull);
        idx >>= (__ffsll(~(long long)idx));  // search_result_to_index: drop the trailing ones and one more bit
        out_vals[i] = __ldg(e + idx);
        if (out_idx) out_idx[i] = idx ? eytz_rank(idx, len - 1, H) : len - 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Lane-group kernel for the partitioned layouts (Simple, Compact, L1, Overlapping, Map; B = 16):
// the descent of partitioned_s_tree.rs:654-880 with 2 lanes x 32 B per node like the fast kernel.
// Positions are kept in slots (4 B) so that the unaligned root windows of Overlapping / Map
// (read_unaligned, :807,857,876) fit the same loop: s' = s * mult[h] + 16 * count.
// ------------------------------------------------------------------------------------------------
// 8 consecutive keys starting at p (4-byte aligned): one LDG.256 when 32-byte aligned, otherwise three
// aligned 16-byte chunks and a word-select network.
template <bool LEAF>
__device__ __forceinline__ Keys<8> ldg_window8(const uint32_t* p, uint64_t pol) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 31u) == 0) return ldg_keys<8, LEAF>(p, pol);  // same L2 policies as the fast kernel
    const uint4* c = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint4 A = __ldg(c), B = __ldg(c + 1);
    const unsigned w = (unsigned)((a >> 2) & 3u);
    uint4 Cc = make_uint4(0, 0, 0, 0);
    if (w) Cc = __ldg(c + 2);
    const bool s2 = (w & 2u) != 0, s1 = (w & 1u) != 0;
    const uint32_t W0 = A.x, W1 = A.y, W2 = A.z, W3 = A.w, W4 = B.x, W5 = B.y, W6 = B.z, W7 = B.w, W8 = Cc.x, W9 = Cc.y, W10 = Cc.z;
    const uint32_t u0 = s2 ? W2 : W0, u1 = s2 ? W3 : W1, u2 = s2 ? W4 : W2, u3 = s2 ? W5 : W3, u4 = s2 ? W6 : W4, u5 = s2 ? W7 : W5,
                   u6 = s2 ? W8 : W6, u7 = s2 ? W9 : W7, u8 = s2 ? W10 : W8;
    Keys<8> r;
    r.k[0] = s1 ? u1 : u0; r.k[1] = s1 ? u2 : u1; r.k[2] = s1 ? u3 : u2; r.k[3] = s1 ? u4 : u3;
    r.k[4] = s1 ? u5 : u4; r.k[5] = s1 ? u6 : u5; r.k[6] = s1 ? u7 : u6; r.k[7] = s1 ? u8 : u7;
    return r;
}

template <int T>
__global__ void __launch_bounds__(512, 2)
pstree_search_group(const __grid_constant__ SstTreeView v, const uint32_t* __restrict__ qs, size_t nq,
                    uint32_t* __restrict__ out_vals, unsigned long long* __restrict__ out_idx) {
    constexpr int G = 2, W = 8, D = G * T;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const unsigned sub = lane & 1u, gbase = lane & ~1u;
    const int L = v.levels;
    const uint64_t pol_inner = policy_evict_last(), pol_leaf = policy_evict_first();
    for (size_t base = ((size_t)blockIdx.x * warps + warp) * (32 * T); base < nq; base += (size_t)gridDim.x * warps * (32 * T)) {
        uint32_t qown[T];
        unsigned long long sown[T];  // first slot; ~0 marks "prefix beyond the last part"
#pragma unroll
        for (int t = 0; t < T; t++) {
            const size_t i = base + (size_t)t * 32 + lane;
            const uint32_t q = i < nq ? __ldcs(qs + i) : 0u;
            qown[t] = q;
            const unsigned long long part = (unsigned long long)(q >> v.shift);
            unsigned long long s0;
            if (part >= v.parts) s0 = ~0ull;
            else if (v.variant == SST_MAP) s0 = __ldg(v.prefix_map + part);
            else if (v.variant == SST_COMPACT) s0 = 0;
            else s0 = part * v.start_mul;
            sown[t] = s0;
        }
        uint32_t q[D], cpart[D];  // cpart: part index for COMPACT (0 otherwise)
        unsigned long long s[D];
        bool live[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            q[d] = __shfl_sync(kFull, qown[d / G], gbase + (d % G));
            const unsigned long long s0 = __shfl_sync(kFull, sown[d / G], gbase + (d % G));
            live[d] = s0 != ~0ull;
            s[d] = live[d] ? s0 : 0ull;
            cpart[d] = (v.variant == SST_COMPACT && live[d]) ? (q[d] >> v.shift) : 0u;
        }
        // COMPACT reads its non-leaf levels from the dense copy (a few pages instead of one page per part)
        const bool dense_upper = v.variant == SST_COMPACT && v.upper != nullptr;
        const unsigned long long ustride = dense_upper ? v.upper_stride : v.part_stride;
        for (int h = 0; h + 1 < L; h++) {
            Keys<W> ks[D];
            const uint32_t* gl = (dense_upper ? v.upper : v.tree) + v.level_slot[h] + sub * W;
#pragma unroll
            for (int d = 0; d < D; d++) ks[d] = ldg_window8<false>(gl + (unsigned long long)cpart[d] * ustride + s[d], pol_inner);
#pragma unroll
            for (int d = 0; d < D; d++) s[d] = s[d] * v.mult[h] + 16ull * group_count<G>(ks[d], q[d], gbase);
        }
        const uint32_t* gl = v.tree + v.level_slot[L - 1];
        Keys<W> ks[D];
#pragma unroll
        for (int d = 0; d < D; d++) ks[d] = ldg_window8<true>(gl + (unsigned long long)cpart[d] * v.part_stride + s[d] + sub * W, pol_leaf);
        uint32_t myval[T];
        unsigned long long mypos[T];
#pragma unroll
        for (int d = 0; d < D; d++) {
            const unsigned c = group_count<G>(ks[d], q[d], gbase);
            const uint32_t cand = pick<W>(ks[d], c % W);
            uint32_t val = __shfl_sync(kFull, cand, gbase + (c < 16u ? c / W : 0u));
            const unsigned long long pos = s[d] + c;  // flat slot in the leaf level (of this part for Compact)
            if (c == 16u) val = pos < v.leaf_slots ? __ldg(gl + (unsigned long long)cpart[d] * v.part_stride + pos) : kMax;
            if (!live[d]) val = kMax;
            if (sub == (unsigned)(d % G)) { myval[d / G] = val; mypos[d / G] = pos; }
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
            const size_t i = base + (size_t)t * 32 + lane;
            if (i >= nq) continue;
            __stcs(out_vals + i, myval[t]);
            if (out_idx) {  // sorted-array index of record through the per-part tables
                unsigned long long index = v.n;
                const unsigned long long part = (unsigned long long)(qown[t] >> v.shift), pos = mypos[t];
                if (part < v.parts) {
                    if (v.variant == SST_MAP) index = pos;
                    else {
                        const unsigned long long st = v.part_start[part], cnt = v.part_start[part + 1] - st;
                        const unsigned long long pp = v.variant == SST_COMPACT ? 0ull : v.part_pos[part];
                        const unsigned long long off = pos > pp ? pos - pp : 0ull;
                        index = st + (off < cnt ? off : cnt);
                    }
                    if (index > v.n) index = v.n;
                }
                __stcs(out_idx + i, index);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Random 64-byte gather probe: the access pattern of ONE tree level without the dependent chain.
// Gives the practical ceiling of "one random node per query" on this GPU for each lane mapping.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

template <int PF>
__global__ void __launch_bounds__(1024, 1)
gather_probe_pf_kernel(const uint32_t* __restrict__ buf, unsigned long long nodes, size_t n_gathers, uint32_t* __restrict__ sink,
                       unsigned long long seed) {
    constexpr int D = 4;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const unsigned sub = lane & 1u, grp = lane / 2;
    const uint64_t pol = policy_evict_first();
    unsigned acc = 0;
    const size_t per_warp = 16 * D;
    for (size_t base = ((size_t)blockIdx.x * warps + warp) * per_warp; base < n_gathers; base += (size_t)gridDim.x * warps * per_warp) {
        Keys<8> ks[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint64_t id = ((mix64(seed + base + (size_t)grp * D + d) >> 32) * nodes) >> 32;
            ks[d] = ldg_keys8_pf<PF>(buf + id * 16 + sub * 8, pol);
        }
#pragma unroll
        for (int d = 0; d < D; d++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc += ks[d].k[i] < 0x40000000u;
    }
    if (acc == 0xffffffffu) sink[0] = acc;
}

template <int G, int D>
__global__ void __launch_bounds__(1024, 1)
gather_probe_kernel(const uint32_t* __restrict__ buf, unsigned long long nodes, size_t n_gathers, uint32_t* __restrict__ sink,
                    unsigned long long seed) {
    constexpr int W = 16 / G;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const unsigned sub = lane & (G - 1), grp = lane / G;
    constexpr unsigned GROUPS = 32 / G;
    const uint64_t pol = policy_evict_first();
    unsigned acc = 0;
    const size_t per_warp = (size_t)GROUPS * D;
    for (size_t base = ((size_t)blockIdx.x * warps + warp) * per_warp; base < n_gathers; base += (size_t)gridDim.x * warps * per_warp) {
        Keys<W> ks[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint64_t id = ((mix64(seed + base + (size_t)grp * D + d) >> 32) * nodes) >> 32;  // multiply-shift range reduction
            ks[d] = ldg_keys<W, true>(buf + id * 16 + sub * W, pol);
        }
#pragma unroll
        for (int d = 0; d < D; d++)
#pragma unroll
            for (int i = 0; i < W; i++) acc += ks[d].k[i] < 0x40000000u;
    }
    if (acc == 0xffffffffu) sink[0] = acc;  // never true; keeps the loads alive
}

// ------------------------------------------------------------------------------------------------
// Launch planning
// ------------------------------------------------------------------------------------------------
thread_local int g_host_grid_cap = 0;  // 0 = no cap (device-resident batches)

bool fast_eligible(const sst_index* idx) { return idx->variant == SST_PLAIN && idx->node_b == 16; }
bool top_eligible(const sst_index* idx) { return fast_eligible(idx) && idx->d_top_table != nullptr; }

template <int G, int T, bool TOP>
int launch_fast(const sst_index* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_vals, unsigned long long* d_idx,
                cudaStream_t st) {
    FastParams fp{};
    fp.tree = idx->d_tree;
    fp.levels = idx->levels;
    fp.leaf_slots = (unsigned long long)idx->layer_sizes[idx->levels - 1] * 16;
    fp.n = idx->n;
    fp.hints = (int)opt(OPT_HINTS);
    for (int h = 0; h < idx->levels; h++) {
        fp.level_slot[h] = (unsigned long long)idx->offsets[h] * 16;
        if (idx->layer_sizes[h] * 64 <= (size_t)opt(OPT_L1_LEVEL_KB) * 1024) fp.l1_levels |= 1u << h;
    }
    if (G == 2 && TOP && idx->d_c5 && opt(OPT_USE_C5)) { fp.c5 = idx->d_c5; fp.h5 = idx->d_h5; }
    size_t smem_bytes = 0;
    if (TOP) {
        fp.top_level = idx->top_level;
        fp.top_table = idx->d_top_table;
        fp.top_low = idx->d_top_low;
        fp.top_nbound = (unsigned)idx->top_nbound;
        smem_bytes = kTopTableBytes + ((idx->top_nbound * 2 + 15) & ~(size_t)15);
    }
    const int threads = std::max(32, (int)opt(OPT_THREADS) & ~31);
    const int sms = sm_count(idx->device);
    const size_t per_cta = (size_t)(threads / 32) * 32 * T;
    int grid = (int)std::min<size_t>(div_ceil(nq, per_cta), (size_t)sms * (size_t)opt(OPT_WAVES));
    // Host-buffer path: sst_query caps the grid (g_host_grid_cap) so that the kernel runs through a chunk's whole copy period
    // at low DRAM intensity instead of saturating the random-access rate in bursts next to the copy engines.
    if (const int cap = opt(OPT_GRID_CAP) > 0 ? (int)opt(OPT_GRID_CAP) : g_host_grid_cap; cap > 0 && grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    auto kern = stree_search_fast<G, T, TOP>;
    if (smem_bytes > 48 * 1024 &&
        !SST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)))
        return SST_ERR_CUDA;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    // Optional: pin the last internal level (the one that just fits L2) with a persisting
    // access-policy window for this launch only; the leaf level streams through.
    const int persist = (int)opt(OPT_PERSIST);
    if (persist && idx->levels >= 2 && idx->persist_ok) {
        const int h = idx->levels - 2;
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = (void*)(idx->d_tree + idx->offsets[h] * 16);
        attr[0].val.accessPolicyWindow.num_bytes = std::min<size_t>(idx->layer_sizes[h] * 64, idx->persist_window_max);
        attr[0].val.accessPolicyWindow.hitRatio = persist >= 100 ? 1.0f : persist / 100.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
        if (persist >= 200) {  // experiment: set it on the stream instead of the launch
            cudaStreamAttrValue sv{};
            sv.accessPolicyWindow = attr[0].val.accessPolicyWindow;
            sv.accessPolicyWindow.hitRatio = 1.0f;
            if (!SST_CUDA_OK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &sv))) return SST_ERR_CUDA;
            nattr = 0;
        }
        if (opt(OPT_DEBUG))
            fprintf(stderr, "[sst] access window base=%p bytes=%zu hit=%.2f mode=%s\n", attr[0].val.accessPolicyWindow.base_ptr,
                    (size_t)attr[0].val.accessPolicyWindow.num_bytes, attr[0].val.accessPolicyWindow.hitRatio, nattr ? "launch" : "stream");
    }
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
    return SST_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, fp, d_qs, nq, d_vals, d_idx)) ? SST_OK : SST_ERR_CUDA;
}

template <int G, int T>
int launch_fast_top(const sst_index* idx, bool top, const uint32_t* d_qs, size_t nq, uint32_t* d_vals,
                    unsigned long long* d_idx, cudaStream_t st) {
    if (top) return launch_fast<G, T, true>(idx, d_qs, nq, d_vals, d_idx, st);
    return launch_fast<G, T, false>(idx, d_qs, nq, d_vals, d_idx, st);
}


// The kernel AUTO picks for this index and batch size.  Measured on B200 (tools/batch_sweep.py, tools/bucketed_once.py):
// the thread-per-query kernel has the shortest latency (6-10 us up to 2^16 queries), the rank-table kernel wins from 2^17
// queries, and for large batches over >= 2^25 keys the reordered-batch pipeline does (59.7 vs 52.8 Gq/s at 2^25 keys,
// 55 vs 34 Gq/s at 2^28 keys for 10^8 queries); the table-less group kernel never wins.
// Synthetic queries for sst_query_calibrate: a uniformly chosen bucket of the pipeline's splitters, a uniform value inside it
// (so they follow the key distribution the index was built on); counter-based hash, no state.
__global__ void calib_queries_kernel(const uint32_t* __restrict__ split, unsigned nb, uint32_t* __restrict__ q, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        const unsigned b = (unsigned)((x >> 40) % nb);
        const uint32_t lo = split[b], hi = split[b + 1];
        q[i] = lo + (uint32_t)((x & 0xffffffffull) % ((unsigned long long)(hi - lo) + 1ull));
    }
}

int resolve_scheme(const sst_index* idx, int scheme, size_t nq) {
    if (scheme != SST_SCHEME_AUTO) return scheme;
    if (idx->variant == SST_EYTZINGER) return SST_SCHEME_GENERIC;
    const int forced = (int)opt(OPT_SCHEME);  // A/B runs: what AUTO resolves to (-1 = the measured rule below)
    if (bucketed_eligible(idx) && idx->n >= (size_t)opt(OPT_BK_AUTO_MIN_N)) {  // plain and Map-partitioned trees
        // Crossover between the rank-table kernel and the pipeline, in queries per batch.  Default rule from the size x batch sweep
        // of profiles/r2_s3_size_batch_sweep.jsonl (one B200: the pipeline wins from 2^24 queries on at 2^25..2^29 keys and from
        // 2^23 on at 2^30; at 2^24 keys and below the tree is L2-resident and the direct kernel always wins); an index that
        // was calibrated on its own device and key distribution (sst_query_calibrate) carries its measured value instead.
        const size_t rule = idx->n >= ((size_t)1 << 30) ? (size_t)1 << 23 : (size_t)1 << 24;
        const size_t min_nq = idx->auto_min_nq ? idx->auto_min_nq : rule;
        if (nq >= (opt(OPT_BK_AUTO_MIN_NQ) >= 0 ? (size_t)opt(OPT_BK_AUTO_MIN_NQ) : min_nq)) return forced >= 0 ? forced : SST_SCHEME_BUCKETED;
    }
    if (idx->variant != SST_PLAIN && idx->node_b == 16 && opt(OPT_PGROUP)) return SST_SCHEME_AUTO;  // lane-group kernel
    if (!fast_eligible(idx)) return SST_SCHEME_GENERIC;
    if (nq < (size_t)opt(OPT_TABLE_MIN_NQ)) return forced >= 0 ? forced : SST_SCHEME_GENERIC;
    return forced >= 0 ? forced : (top_eligible(idx) ? SST_SCHEME_TABLE : SST_SCHEME_GROUP2);
}

}  // namespace

// Kernel launches of one sst_query_device call.  The reordered-batch pipeline runs 4 kernels per run of up to 2^30 queries (BK_SUB_LOG2)
// (partition, work items, search, un-permute) and one more un-permute for the index output; the round-1
// pipeline ran 7 (rank, column sums, plan, offsets, scatter, search, gather).
int query_launch_count(const sst_index* idx, int scheme, size_t nq, bool want_idx) {
    if (nq == 0) return 0;
    if (resolve_scheme(idx, scheme, nq) != SST_SCHEME_BUCKETED) return 1;
    const int base = 4;
    return (int)(div_ceil(nq, bucketed_sub_batch()) * (base + (want_idx ? 1 : 0) + (idx->variant == SST_PLAIN ? 0 : 1) +
                                                   (want_idx && idx->variant != SST_PLAIN && idx->variant != SST_MAP && idx->variant != SST_COMPACT ? 1 : 0)));  // partitioned: + the q > MAX fix-up (+ flat -> sorted index)
}

int launch_query(const sst_index* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_vals, unsigned long long* d_idx,
                 int scheme, cudaStream_t st) {
    if (nq == 0) return SST_OK;
    if (idx->variant == SST_EYTZINGER) {
        if (scheme != SST_SCHEME_AUTO && scheme != SST_SCHEME_GENERIC) {
            set_error(SST_ERR_UNSUPPORTED, "the Eytzinger baseline has one kernel; use SST_SCHEME_AUTO");
            return SST_ERR_UNSUPPORTED;
        }
        const int grid = (int)std::min<size_t>(div_ceil(nq, 256), (size_t)sm_count(idx->device) * 8);
        eytzinger_search_kernel<<<grid, 256, 0, st>>>(idx->d_tree, idx->eytz_words, idx->eytz_h, d_qs, nq, d_vals, d_idx);
        return SST_CUDA_OK(cudaGetLastError()) ? SST_OK : SST_ERR_CUDA;
    }
    auto launch_pgroup = [&]() {  // lane-group kernel of the partitioned layouts
        const int sms = sm_count(idx->device);
        const int T = (int)opt(OPT_PT);  // measured: T=1 27.7 vs T=2 23.9 Gq/s (Simple, 2^28 keys)
        const int grid = (int)std::min<size_t>(div_ceil(nq, (size_t)16 * 32 * T), (size_t)sms * 2);
        if (T == 1) pstree_search_group<1><<<grid, 512, 0, st>>>(idx->view, d_qs, nq, d_vals, d_idx);
        else pstree_search_group<2><<<grid, 512, 0, st>>>(idx->view, d_qs, nq, d_vals, d_idx);
        return SST_CUDA_OK(cudaGetLastError()) ? SST_OK : SST_ERR_CUDA;
    };
    const bool pgroup = idx->variant != SST_PLAIN && idx->node_b == 16 && opt(OPT_PGROUP);
    if (scheme == SST_SCHEME_AUTO && pgroup && resolve_scheme(idx, scheme, nq) != SST_SCHEME_BUCKETED) return launch_pgroup();
    const bool was_auto = scheme == SST_SCHEME_AUTO;
    scheme = resolve_scheme(idx, scheme, nq);
    if (was_auto && scheme == SST_SCHEME_BUCKETED) {
        // the pipeline needs ~10-14 bytes of device scratch per query: when that cannot be had, AUTO falls back to the direct kernel
        const int rc = launch_bucketed(idx, d_qs, nq, d_vals, d_idx, st);
        if (rc != SST_ERR_CAPACITY) return rc;
        clear_error();
        if (idx->variant != SST_PLAIN) return pgroup ? launch_pgroup() : launch_query(idx, d_qs, nq, d_vals, d_idx, SST_SCHEME_GENERIC, st);  // Map tree
        scheme = top_eligible(idx) ? SST_SCHEME_TABLE : SST_SCHEME_GROUP2;
    }
    if (scheme == SST_SCHEME_BUCKETED && bucketed_eligible(idx)) return launch_bucketed(idx, d_qs, nq, d_vals, d_idx, st);
    if (scheme != SST_SCHEME_GENERIC && !fast_eligible(idx)) {  // (BINSEARCH included: plain B=16 only)
        set_error(SST_ERR_UNSUPPORTED, "the group/table kernels serve the plain B=16 tree; use SST_SCHEME_AUTO or SST_SCHEME_GENERIC");
        return SST_ERR_UNSUPPORTED;
    }
    const int T = (int)opt(OPT_T);
    switch (scheme) {
        case SST_SCHEME_BUCKETED:
            return launch_bucketed(idx, d_qs, nq, d_vals, d_idx, st);
        case SST_SCHEME_TABLE: {
            const bool top = top_eligible(idx);  // trees of height 1 have nothing above the leaf
            if (opt(OPT_TABLE_G) == 4) {
                if (T == 1) return launch_fast_top<4, 1>(idx, top, d_qs, nq, d_vals, d_idx, st);
                return launch_fast_top<4, 2>(idx, top, d_qs, nq, d_vals, d_idx, st);
            }
            if (T == 1) return launch_fast_top<2, 1>(idx, top, d_qs, nq, d_vals, d_idx, st);
            return launch_fast_top<2, 2>(idx, top, d_qs, nq, d_vals, d_idx, st);
        }
        case SST_SCHEME_GROUP4:
            if (T == 1) return launch_fast<4, 1, false>(idx, d_qs, nq, d_vals, d_idx, st);
            return launch_fast<4, 2, false>(idx, d_qs, nq, d_vals, d_idx, st);
        case SST_SCHEME_GROUP16:
            return launch_fast<16, 1, false>(idx, d_qs, nq, d_vals, d_idx, st);
        case SST_SCHEME_GROUP2:
            if (T == 1) return launch_fast<2, 1, false>(idx, d_qs, nq, d_vals, d_idx, st);
            return launch_fast<2, 2, false>(idx, d_qs, nq, d_vals, d_idx, st);
        case SST_SCHEME_BINSEARCH: {
            // every plain B=16 layout stores the sorted keys verbatim at the start of its leaf level
            const int sms = sm_count(idx->device);
            const int grid = (int)std::min<size_t>(div_ceil(nq, 256), (size_t)sms * 8);
            binary_search_kernel<<<grid, 256, 0, st>>>(idx->d_tree + idx->offsets[idx->levels - 1] * 16, idx->n, d_qs, nq, d_vals, d_idx);
            return SST_CUDA_OK(cudaGetLastError()) ? SST_OK : SST_ERR_CUDA;
        }
        case SST_SCHEME_GENERIC: {
            const int sms = sm_count(idx->device);
            const int grid = (int)std::min<size_t>(div_ceil(nq, 256), (size_t)sms * 32);
            stree_search_generic<<<grid, 256, 0, st>>>(idx->view, d_qs, nq, d_vals, d_idx);
            return SST_CUDA_OK(cudaGetLastError()) ? SST_OK : SST_ERR_CUDA;
        }
        default:
            set_error(SST_ERR_ARG, "unknown scheme");
            return SST_ERR_ARG;
    }
}

}  // namespace sst

// =================================================================================================
// C ABI
// =================================================================================================
using namespace sst;

extern "C" {

int sst_query_launches(const sst_index_t* idx, int scheme) { return idx ? query_launch_count(idx, scheme, 1, false) : 0; }

int sst_query_plan(const sst_index_t* idx, size_t nq, int scheme, int want_index, int* out_scheme, int* out_launches) {
    clear_error();
    if (!idx) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (out_scheme) *out_scheme = resolve_scheme(idx, scheme, nq);
    if (out_launches) *out_launches = query_launch_count(idx, scheme, nq, want_index != 0);
    return SST_OK;
}

int sst_last_stage_ms(double* out, int n) { return last_stage_ms(out, n); }

int sst_query_reserve(const sst_index_t* idx, size_t nq, int want_index) {
    clear_error();
    if (!idx) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    DeviceGuard g(idx->device);
    if (!g.ok) return SST_ERR_CUDA;
    return reserve_bucketed(idx, nq, want_index != 0);
}
void sst_query_release(void) { release_bucketed_scratch(); }

int sst_query_device(const sst_index_t* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_out_vals, uint64_t* d_out_idx,
                     int scheme, void* stream) {
    clear_error();
    if (!idx || (nq && (!d_qs || !d_out_vals))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    DeviceGuard g(idx->device);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;  // NULL == the CUDA legacy default stream
    return launch_query(idx, d_qs, nq, d_out_vals, (unsigned long long*)d_out_idx, scheme, st);
}

// Per (host thread, device) staging ring for the host-buffer path: allocated once and reused, so
// that a call costs no cudaMalloc/cudaFree (which synchronise the device).
namespace {
constexpr int kRing = 3;  // device-side ring: three chunks in flight (H2D | kernel | D2H)
struct Staging {
    size_t cap = 0, cap_idx = 0;
    uint32_t* q[kRing] = {};
    uint32_t* v[kRing] = {};
    unsigned long long* i[kRing] = {};
    cudaEvent_t e_in[kRing] = {}, e_k[kRing] = {}, e_out[kRing] = {};
    bool events = false;
    int device = -1;
    ~Staging() {  // host thread exits: give the device memory back (errors at process teardown are ignored)
        if (device < 0 || (!cap && !cap_idx && !events)) return;
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return; }
        for (int b = 0; b < kRing; b++) {
            cudaFree(q[b]); cudaFree(v[b]); cudaFree(i[b]);
            if (events) { cudaEventDestroy(e_in[b]); cudaEventDestroy(e_k[b]); cudaEventDestroy(e_out[b]); }
        }
        (void)cudaGetLastError();
        if (prev >= 0) cudaSetDevice(prev);
    }
};
thread_local Staging g_staging[64];

bool staging_ensure(Staging& s, size_t cap, bool want_idx) {
    if (!s.events) {
        for (int b = 0; b < kRing; b++)
            if (!SST_CUDA_OK(cudaEventCreateWithFlags(&s.e_in[b], cudaEventDisableTiming)) ||
                !SST_CUDA_OK(cudaEventCreateWithFlags(&s.e_k[b], cudaEventDisableTiming)) ||
                !SST_CUDA_OK(cudaEventCreateWithFlags(&s.e_out[b], cudaEventDisableTiming)))
                return false;
        s.events = true;
    }
    if (cap > s.cap) {
        for (int b = 0; b < kRing; b++) {
            cudaFree(s.q[b]); cudaFree(s.v[b]);
            s.q[b] = s.v[b] = nullptr;
            if (!SST_CUDA_OK(cudaMalloc(&s.q[b], cap * 4)) || !SST_CUDA_OK(cudaMalloc(&s.v[b], cap * 4))) { s.cap = 0; return false; }
        }
        s.cap = cap;
    }
    if (want_idx && cap > s.cap_idx) {
        for (int b = 0; b < kRing; b++) {
            cudaFree(s.i[b]);
            s.i[b] = nullptr;
            if (!SST_CUDA_OK(cudaMalloc(&s.i[b], cap * 8))) { s.cap_idx = 0; return false; }
        }
        s.cap_idx = cap;
    }
    return true;
}
}  // namespace

// Host buffers: chunked three-stage pipeline (H2D | kernel | D2H) over two copy streams and the
// compute stream, so that PCIe traffic in both directions overlaps the kernel.  Full speed needs
// page-locked host buffers (cudaMemcpyAsync from pageable memory is staged by the driver).
int sst_query(const sst_index_t* idx, const uint32_t* qs, size_t nq, uint32_t* out_vals, uint64_t* out_idx, int scheme) {
    clear_error();
    if (!idx || (nq && (!qs || !out_vals))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (nq == 0) return SST_OK;
    DeviceGuard g(idx->device);
    if (!g.ok) return SST_ERR_CUDA;
    const int dev = idx->device;
    if (dev < 0 || dev >= 64) { set_error(SST_ERR_ARG, "device index out of range"); return SST_ERR_ARG; }
    cudaStream_t s_in = thread_copy_stream(dev, 0), s_k = thread_stream(dev), s_out = thread_copy_stream(dev, 1);
    if (!s_in || !s_k || !s_out) return SST_ERR_CUDA;
    const size_t chunk = (size_t)opt(OPT_CHUNK);
    const size_t nchunks = div_ceil(nq, chunk);
    const int NB = kRing;
    Staging& sg = g_staging[dev];
    sg.device = dev;
    if (!staging_ensure(sg, std::min(chunk, nq), out_idx != nullptr)) return SST_ERR_CUDA;
    // The kernel next to the copies.  At full width the search kernel saturates the DRAM random-access rate for a third of
    // every chunk period and the copy engines lose ~40 % of their rate meanwhile (2^28 keys, 10^8 queries: copies alone
    // 8.4-8.7 ms, with the kernel 9.6-10.1 ms).  A kernel that only just keeps up with PCIe runs through the whole period at
    // low intensity: on a quarter of the SMs per HBM-resident level (37 CTAs at 2^28 keys, ~0.33 Gq/s each = 12.2 Gq/s, the
    // PCIe ceiling) the call takes 8.7-9.2 ms on four of five boxes measured, 9.7 on the fifth (profiles/r1_s3_e2e_*.log).
    // Narrower is kernel-bound (30 CTAs: 10.3 ms), so the rule errs on the wide side: a level counts as HBM-resident as soon
    // as the levels down to it outgrow L2.  A feedback controller on the kernel stream's utilisation was tried and dropped:
    // the timing events and the per-chunk host wait it needs cost more (1.0-1.2 ms) than the pacing gains.
    // SST_HOST_GRID_CAP: 0 = this rule, < 0 = full width, > 0 = CTAs.
    struct CapGuard { ~CapGuard() { g_host_grid_cap = 0; } } cap_guard;
    {
        int cap = (int)opt(OPT_HOST_GRID_CAP);
        if (cap == 0 && idx->variant == SST_PLAIN && idx->node_b == 16 && nchunks >= 3) {  // (fewer chunks: nothing to overlap with)
            int l2 = 0;
            cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
            size_t cum = 0;
            int hbm_levels = 0;
            for (int h = 0; h < idx->levels; h++) {
                cum += (size_t)idx->layer_blocks[h] * 64;
                if (cum > (size_t)l2) hbm_levels++;
            }
            cap = hbm_levels ? sm_count(dev) / 4 * hbm_levels : 0;
        }
        g_host_grid_cap = cap > 0 ? cap : 0;
    }
    int rc = SST_OK;
    for (size_t c = 0; c < nchunks && rc == SST_OK; c++) {
        const int b = (int)(c % NB);
        const size_t off = c * chunk, cnt = std::min(chunk, nq - off);
        // buffer b is free once the D2H of chunk c-NB has finished
        if (c >= (size_t)NB && !SST_CUDA_OK(cudaStreamWaitEvent(s_in, sg.e_out[b], 0))) { rc = SST_ERR_CUDA; break; }
        if (!SST_CUDA_OK(cudaMemcpyAsync(sg.q[b], qs + off, cnt * 4, cudaMemcpyHostToDevice, s_in)) ||
            !SST_CUDA_OK(cudaEventRecord(sg.e_in[b], s_in)) || !SST_CUDA_OK(cudaStreamWaitEvent(s_k, sg.e_in[b], 0))) { rc = SST_ERR_CUDA; break; }
        rc = launch_query(idx, sg.q[b], cnt, sg.v[b], out_idx ? sg.i[b] : nullptr, scheme, s_k);
        if (rc != SST_OK) break;
        if (!SST_CUDA_OK(cudaEventRecord(sg.e_k[b], s_k)) || !SST_CUDA_OK(cudaStreamWaitEvent(s_out, sg.e_k[b], 0)) ||
            !SST_CUDA_OK(cudaMemcpyAsync(out_vals + off, sg.v[b], cnt * 4, cudaMemcpyDeviceToHost, s_out)) ||
            (out_idx && !SST_CUDA_OK(cudaMemcpyAsync(out_idx + off, sg.i[b], cnt * 8, cudaMemcpyDeviceToHost, s_out))) ||
            !SST_CUDA_OK(cudaEventRecord(sg.e_out[b], s_out))) { rc = SST_ERR_CUDA; break; }
    }
    if (!SST_CUDA_OK(cudaStreamSynchronize(s_in)) || !SST_CUDA_OK(cudaStreamSynchronize(s_k)) ||
        !SST_CUDA_OK(cudaStreamSynchronize(s_out)))
        rc = rc == SST_OK ? SST_ERR_CUDA : rc;
    return rc;
}

double sst_time_query_device(const sst_index_t* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_out_vals,
                             uint64_t* d_out_idx, int scheme, int warmup, int iters) {
    clear_error();
    if (!idx || iters < 1) { set_error(SST_ERR_ARG, "bad argument"); return -1.0; }
    DeviceGuard g(idx->device);
    if (!g.ok) return -1.0;
    cudaStream_t st = thread_stream(idx->device);
    if (!SST_CUDA_OK(cudaDeviceSynchronize())) return -1.0;  // inputs may come from another stream
    cudaEvent_t a, b;
    if (!SST_CUDA_OK(cudaEventCreate(&a)) || !SST_CUDA_OK(cudaEventCreate(&b))) return -1.0;
    for (int i = 0; i < warmup; i++)
        if (launch_query(idx, d_qs, nq, d_out_vals, (unsigned long long*)d_out_idx, scheme, st) != SST_OK) return -1.0;
    cudaEventRecord(a, st);
    for (int i = 0; i < iters; i++)
        if (launch_query(idx, d_qs, nq, d_out_vals, (unsigned long long*)d_out_idx, scheme, st) != SST_OK) return -1.0;
    cudaEventRecord(b, st);
    float ms = 0;
    bool ok = SST_CUDA_OK(cudaEventSynchronize(b)) && SST_CUDA_OK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ok ? (double)ms / iters : -1.0;
}

// Measures, on this index and this device, from which batch size on the reordered-batch pipeline beats the direct kernel, and
// makes SST_SCHEME_AUTO use that crossover for this index (instead of the rule taken from one B200).  Synthetic queries that
// follow the key distribution (a uniformly chosen bucket of the pipeline's splitters, a uniform value inside it), batches of
// 2^20 .. max_nq queries (max_nq is rounded down to a power of two, at most 2^26).  Costs a few tens of milliseconds and
// 12 bytes of device memory per query of the largest batch while it runs.  *out_min_nq (nullable) receives the crossover
// (SIZE_MAX: the pipeline never won).  No-op (SST_OK, *out_min_nq = 0) for an index the pipeline does not serve.
int sst_query_calibrate(sst_index_t* idx, size_t max_nq, size_t* out_min_nq) {
    clear_error();
    if (!idx) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (out_min_nq) *out_min_nq = 0;
    if (!bucketed_eligible(idx)) return SST_OK;
    DeviceGuard g(idx->device);
    if (!g.ok) return SST_ERR_CUDA;
    int top = 20;
    while (top < 26 && ((size_t)2 << top) <= max_nq) top++;
    const size_t cap = (size_t)1 << top;
    if (cap > max_nq) { set_error(SST_ERR_ARG, "max_nq must be at least 2^20"); return SST_ERR_ARG; }
    uint32_t *d_q = nullptr, *d_v = nullptr;
    if (!SST_CUDA_OK(cudaMalloc(&d_q, cap * 4)) || !SST_CUDA_OK(cudaMalloc(&d_v, cap * 4))) { cudaFree(d_q); return SST_ERR_CUDA; }
    cudaStream_t st = thread_stream(idx->device);
    calib_queries_kernel<<<cur_sms() * 8, 256, 0, st>>>(idx->bk.d_split, idx->bk.nb, d_q, cap);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int direct = idx->variant != SST_PLAIN ? SST_SCHEME_AUTO : top_eligible(idx) ? SST_SCHEME_TABLE : SST_SCHEME_GROUP2;
    const size_t saved = idx->auto_min_nq;
    idx->auto_min_nq = ~(size_t)0;  // AUTO = the direct kernel while the partitioned layouts are timed through it
    auto time_ms = [&](int scheme, size_t nq) -> double {
        for (int it = 0; it < 4; it++) {
            if (it == 1) cudaEventRecord(a, st);
            if (launch_query(idx, d_q, nq, d_v, nullptr, scheme, st) != SST_OK) return -1.0;
        }
        cudaEventRecord(b, st);
        float ms = 0;
        if (!SST_CUDA_OK(cudaEventSynchronize(b)) || !SST_CUDA_OK(cudaEventElapsedTime(&ms, a, b))) return -1.0;
        return ms / 3.0;
    };
    size_t found = ~(size_t)0;
    bool ok = true;
    for (int lg = top; lg >= 20; lg--) {  // from the largest batch down: the crossover is the smallest size of the winning run
        const size_t nq = (size_t)1 << lg;
        const double td = time_ms(direct, nq), tb = time_ms(SST_SCHEME_BUCKETED, nq);
        if (td < 0 || tb < 0) { ok = false; break; }
        if (tb <= td) found = nq; else break;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_q);
    cudaFree(d_v);
    if (!ok) { idx->auto_min_nq = saved; return sst_last_status() != SST_OK ? sst_last_status() : SST_ERR_CUDA; }
    idx->auto_min_nq = found;
    if (out_min_nq) *out_min_nq = found;
    return SST_OK;
}

double sst_probe_gather64(int device, size_t bytes, size_t n_gathers, int lanes_per_node, int iters) {
    clear_error();
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU"); return -1.0; }
    if (bytes < 64 || iters < 1) { set_error(SST_ERR_ARG, "bad argument"); return -1.0; }
    DeviceGuard g(device);
    if (!g.ok) return -1.0;
    configure_l2_fetch(device);
    cudaStream_t st = thread_stream(device);
    uint32_t *buf = nullptr, *sink = nullptr;
    const unsigned long long nodes = bytes / 64;  // < 2^32 nodes (256 GiB)
    if (!SST_CUDA_OK(cudaMalloc(&buf, nodes * 64)) || !SST_CUDA_OK(cudaMalloc(&sink, 64))) { cudaFree(buf); return -1.0; }
    cudaMemsetAsync(buf, 0x55, nodes * 64, st);
    const int grid = sm_count(device), threads = 1024;
    auto launch = [&](unsigned long long seed) {
        switch (lanes_per_node) {
            case 16: gather_probe_kernel<16, 16><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;
            case 8: gather_probe_kernel<8, 8><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;
            case 4: gather_probe_kernel<4, 8><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;
            case 2: gather_probe_kernel<2, 4><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;
            case 1064: gather_probe_pf_kernel<64><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;   // experiment:
            case 1128: gather_probe_pf_kernel<128><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;  // explicit L2
            case 1256: gather_probe_pf_kernel<256><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;  // prefetch size
            case 1000: gather_probe_pf_kernel<0><<<grid, threads, 0, st>>>(buf, nodes, n_gathers, sink, seed); break;
            default: return false;
        }
        return true;
    };
    double result = -1.0;
    if (!launch(1)) set_error(SST_ERR_ARG, "lanes_per_node must be 2, 4, 8 or 16");
    else {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
        for (int i = 0; i < iters; i++) launch(1000003ull * (i + 2));
        cudaEventRecord(b, st);
        float ms = 0;
        if (SST_CUDA_OK(cudaEventSynchronize(b)) && SST_CUDA_OK(cudaEventElapsedTime(&ms, a, b)) && SST_CUDA_OK(cudaGetLastError()))
            result = (double)n_gathers * 64.0 * iters / (ms * 1e-3) / 1e9;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    cudaFree(buf);
    cudaFree(sink);
    return result;
}

}  // extern "C"
