// bucketed.cu -- reordered-batch lower_bound: sort every tile of the query batch by key range, answer each range
// from shared memory + one leaf sector per query, put the answers back in the caller's order.  Same results as every other scheme.
//
// Replaces (static-search-tree/src): the batched/interleaved searches of s_tree.rs:208-832 for LARGE
// batches over LARGE trees.  The reference keeps 128 independent queries in flight and lets each one
// miss the cache once per level (s_tree.rs:303-326); on B200 that design is bound by the number of
// random DRAM accesses per second (~43 G/s), one per query for the leaf level.
// Here the batch is first reordered so that all queries that fall into the same 1-2 MB window of the
// leaf level are answered together by one CTA (4 launches per run of up to 2^30 queries):
//
//   partition  (bk_part_kernel)     per 16384-query tile: TMA bulk load (next tile prefetched) -> bucket id per query (ONE shared
//                                   load: a table over the top 13 key bits that packs the bucket count, a flag and the next
//                                   splitter's low bits) -> rank inside the tile by per-warp counters with claim / ballot steps
//                                   (no atomics) -> scan -> sorted tile in shared memory -> TMA bulk store; plus the 16-bit
//                                   position map and one run descriptor {start, count} per (bucket, tile), bucket-major
//   plan       (bk_items_kernel)    work items (bucket, tile range) of at most 32768 queries
//   search     (bk_search2_kernel)  per work item: the bucket's separators (last key of every 8-key half node: 32768 x 32 bits, or
//                                   above 2^28 slots 65536 x 16-bit offsets inside their jump cell) and its jump table (192 KB in
//                                   all) are staged by 1-D TMA bulk copies; a warp walks 32 runs at a time, a query is ranked
//                                   among the separators with the jump cell + three separator loads and finished with ONE
//                                   32-byte leaf load (LDG.256); the answer overwrites the query IN PLACE
//   un-permute (bk_unperm_kernel)   tile of answers by TMA bulk load (double buffered) -> caller's order through the position
//                                   map in shared memory -> 16-byte stores
//
// DRAM traffic per 10^8 queries over 2^28 keys: the leaf level once (1 GiB, instead of 6.4-12.8 GB of
// random sectors), 128 MB of separators, and ~3 GB of query/result/position streams (4.26 GB in all, ncu).  The tree image is
// untouched: the leaf level is read in place and the separators are a GPU-only auxiliary array like
// the rank table of stree_search.cu.  Served: plain B = 16 trees (any new_params flags) and the Map, Simple, L1 and
// Overlapping partitioned layouts (their leaf level is one sorted flat array) of 2^22 .. 2^30 leaf slots; Compact through a
// dense copy of its keys.
#include <algorithm>
#include <type_traits>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace sst {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kTile = 16384;                 // queries per partition tile (64 KB of shared memory)
constexpr int kThreads = 512;
// jump table: one cell per separator of a bucket (cells = r), r + 8 u16 entries per bucket (a multiple of 16 bytes)
constexpr int kBtShift = 18;                 // bucket table over the top 13 bits of a 31-bit key
constexpr int kBtCells = 1 << (31 - kBtShift);
constexpr int kBtStride = kBtCells + 8;
constexpr unsigned kMinChunk = 16384;        // queries per search work item: at least this many (scratch sizing)

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
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Exclusive prefix sum of one value per thread over the CTA (blockDim.x = 32 * nwarps <= 1024).
// s_warp needs nwarps + 1 entries; the caller synchronises before reusing it.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* s_warp, unsigned* total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(kFull, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const unsigned w = lane < nwarps ? s_warp[lane] : 0u;
        unsigned ws = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(kFull, ws, o);
            if (lane >= (unsigned)o) ws += y;
        }
        if (lane < nwarps) s_warp[lane] = ws - w;
        if (lane == 31) s_warp[nwarps] = ws;
    }
    __syncthreads();
    *total = s_warp[nwarps];
    return s_warp[warp] + x - v;
}

// Signed node compare of node.rs:91-108: a query above MAX is below every key, i.e. behaves like 0.
__device__ __forceinline__ uint32_t canonical(uint32_t q) { return q > kMax ? 0u : q; }

// bucket(q) = number of splitters < q; split[i], i in [1, nb), are the splitters (split[0] = 0, split[nb] = MAX).
// bt[c] & 0x7fff = number of splitters whose top 13 bits are < c; bit 15 is set when two or more splitters share
// the prefix c (skewed keys), otherwise one compare against the next splitter settles it: a splitter in a later
// cell is above q anyway.
__device__ __forceinline__ unsigned bk_bucket(const uint16_t* __restrict__ bt, const uint32_t* __restrict__ split, uint32_t q) {
    const unsigned c = q >> kBtShift;
    const unsigned e = bt[c];
    unsigned lo = e & 0x7fffu;
    if (e & 0x8000u) {
        unsigned hi = bt[c + 1] & 0x7fffu;
        while (lo < hi) {
            const unsigned m = (lo + hi) >> 1;
            if (split[m + 1] < q) lo = m + 1; else hi = m;
        }
        return lo;
    }
    return lo + (split[lo + 1] < q ? 1u : 0u);
}

// The same in ONE shared load per query: pk[c] = next << 12 | flag << 11 | lo with lo = bt[c] & 0x7fff, flag = bit 15 of
// bt[c] and next = the low 18 key bits of splitter lo + 1 when that splitter lies in cell c, else 0x3ffff (no key of the
// cell is above it).  Flagged cells (two or more splitters with the same 13-bit prefix: skewed keys) take bk_bucket
// through the global copies of the tables (the lookup itself is written out in part_tile).
__device__ __forceinline__ uint32_t bk_pack_cell(const uint16_t* __restrict__ bt, const uint32_t* __restrict__ split, unsigned nb, unsigned c) {
    const unsigned e = bt[c], lo = e & 0x7fffu;
    const uint32_t nx = lo + 1u <= nb ? split[lo + 1u] : kMax;
    const uint32_t thr = (nx >> kBtShift) == c ? (nx & ((1u << kBtShift) - 1u)) : ((1u << kBtShift) - 1u);
    return (thr << 12) | ((e & 0x8000u) ? 0x800u : 0u) | lo;
}
struct BkView {
    const uint16_t* bt;
    const uint32_t* split;
    unsigned nb, nbp, bpt;  // buckets, padded to a multiple of kThreads, buckets per thread
    unsigned* above;        // Map-partitioned trees: set to 1 when some query is above MAX (null for plain trees)
};

// ------------------------------------------------------------------------------------------------
// rank: bucket id, stable tile-local position and tile x bucket counts
// ------------------------------------------------------------------------------------------------
// peers &= (bit BIT of b set) ? ballot(bit set) : ~ballot(bit set), for BIT in [BIT, BITS): the lanes whose bucket id
// equals this lane's.  Spelled out in PTX so that a bit costs 4 instructions (LOP3->P, VOTE, SEL, LOP3).
template <int BIT, int BITS>
__device__ __forceinline__ void ballot_bits(unsigned& peers, unsigned b) {
    if constexpr (BIT < BITS) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b32 t, v, m;\n"
            "and.b32 t, %1, %2;\n"
            "setp.ne.u32 p, t, 0;\n"
            "vote.sync.ballot.b32 v, p, 0xffffffff;\n"
            "selp.b32 m, 0, 0xffffffff, p;\n"
            "lop3.b32 %0, %0, v, m, 0x60;\n"  // a & (b ^ c)
            "}\n"
            : "+r"(peers)
            : "r"(b), "n"(1u << BIT));
        ballot_bits<BIT + 1, BITS>(peers, b);
    }
}

// Ranks the ITEMS queries every lane of a warp holds (pk[r] = bucket id, 0xffffffff = no query) among the warp's queries of
// the same bucket, through the warp's private counters cntw[bucket] (count in bits 0..10, claim tag above): on return
// pk[r] = bucket | rank << 16 and cntw[b] & 0x7ff = the warp's number of queries in bucket b.
template <int BITS, bool FULL, int HYBRID, int ITEMS>
__device__ __forceinline__ void rank_items(uint16_t* cntw, uint32_t (&pk)[ITEMS], unsigned lane, unsigned lt_mask) {
        // rank inside the warp's queries: lanes with the same bucket find each other by ballots over
            // the bucket bits; the lowest of them bumps the warp's private counter
    #pragma unroll
            for (int r = 0; r < ITEMS; r++) {
                const bool valid = FULL || pk[r] != 0xffffffffu;
                const unsigned b = valid ? pk[r] : 0u;
                // Two ways to rank, mixed step by step so that the work is split between the ALU pipe (ballots) and
                // the shared-memory pipe (claims) -- each alone is bound by its pipe (0.61 / 0.55 ms per 10^8 queries):
                //  claim:   every lane writes count + 1 tagged with its lane id (5 tag bits above the 11 count bits); the lane
                //           whose tag sticks takes rank = count; if any lane lost (two queries of one bucket in the same step,
                //           ~1 step in 5 for uniform queries) the losers are settled by ballots, so the cost stays bounded
                //           when every query hits the same bucket
                //  ballots: lanes with the same bucket find each other by ballots over the bucket bits; the lowest of them
                //           bumps the counter by the group size
                // HYBRID = 0: ballots only; k > 0: claim on every step with r % k != 0 (2: every other step); k < 0: claim on r % -k == 0
                const bool claim_step = HYBRID > 0 ? (r % (HYBRID > 0 ? HYBRID : 1)) != 0 : HYBRID < 0 ? (r % (HYBRID < 0 ? -HYBRID : 1)) == 0 : false;
                if (claim_step) {
                    const unsigned w = valid ? cntw[b] : 0u;
                    __syncwarp();
                    if (valid) cntw[b] = (uint16_t)(((w & 0x7ffu) + 1u) | (lane << 11));
                    __syncwarp();
                    const bool lost = valid && (unsigned)(cntw[b] >> 11) != lane;
                    unsigned rank = w & 0x7ffu;
                    unsigned lostm = __ballot_sync(kFull, lost);
                    if (lostm) {  // (warp-uniform, 2 steps in 5) lanes that share a bucket with a winner
                        // Usually ONE bucket is contested (a pair of lanes; or every lane when the queries are all alike): the
                        // losers of the lowest contested lane's bucket find each other with one shuffle and one ballot and are
                        // settled together; twice over, then whatever is left takes the ballots over the bucket bits, so the
                        // cost stays bounded for any input (the ballots alone cost 45 instructions whenever a step had a loser;
                        // letting the losers claim again instead was measured slower: rank 0.410 -> 0.431 ms).
                        bool mine = lost;
#pragma unroll
                        for (int it = 0; it < 2; it++) {
                            if (lostm) {
                                const unsigned bl = __shfl_sync(kFull, b, __ffs(lostm) - 1);
                                const bool hit = mine && b == bl;
                                const unsigned same = __ballot_sync(kFull, hit);
                                const unsigned old = hit ? (cntw[b] & 0x7ffu) : 0u;  // includes the winner's +1
                                __syncwarp();
                                if (hit && (same & lt_mask) == 0u) cntw[b] = (uint16_t)(old + __popc(same));
                                if (hit) { rank = old + __popc(same & lt_mask); mine = false; }
                                lostm &= ~same;
                            }
                        }
                        if (lostm) {
                            unsigned peers = lostm;
                            ballot_bits<0, BITS>(peers, b);
                            const unsigned before = peers & lt_mask;
                            const unsigned old = mine ? (cntw[b] & 0x7ffu) : 0u;
                            __syncwarp();
                            if (mine && before == 0u) cntw[b] = (uint16_t)(old + __popc(peers));
                            if (mine) rank = old + __popc(before);
                        }
                        __syncwarp();
                    }
                    if (valid) pk[r] = b | (rank << 16);
                } else {
                    unsigned peers = FULL ? kFull : __ballot_sync(kFull, valid);
                    ballot_bits<0, BITS>(peers, b);
                    const unsigned before = peers & lt_mask;
                    const unsigned old = valid ? (cntw[b] & 0x7ffu) : 0u;
                    __syncwarp();
                    if (valid && before == 0u) cntw[b] = (uint16_t)(old + __popc(peers));
                    __syncwarp();
                    if (valid) pk[r] = b | ((old + __popc(before)) << 16);
                }
            }
}

// ------------------------------------------------------------------------------------------------
// search: one work item = (bucket, chunk of its queries)
// ------------------------------------------------------------------------------------------------
struct BkSearchParams {
    const uint32_t* sep;     // [nb * r]
    const uint16_t* jump;    // [nb][r + 8]
    const uint2* meta;       // [nb] {lo, shift}
    const uint32_t* leaf;    // sorted keys (leaf level of the image, MAX-padded)
    unsigned r;              // half nodes (separators) per bucket
    const uint16_t* sep16;   // 16-bit mode: [nb * r] offsets of the separators inside their jump cells (sep is unused then)
    unsigned cells;          // jump cells per bucket (r, or r / 2 in 16-bit mode): jump has cells + 8 entries per bucket
    unsigned long long m8;   // blocks of G keys (half nodes / nodes) that hold keys
    unsigned long long n;
};


// ================================================================================================
// The pipeline: tile-local partition -> plan -> search over runs, in place -> streaming un-permute
//
// The round-1 pipeline (git history) gathered every bucket into one contiguous array: rank kernel, three plan kernels over the
// tiles x buckets count matrix, a scatter kernel that read the queries a second time, and a gather kernel that collected
// ~16-query runs back.  Here a tile is only sorted LOCALLY: the partition kernel reads a tile once (TMA bulk load,
// prefetched one tile ahead), ranks it, and writes it back as one contiguous 64 KB block in bucket order (TMA bulk store)
// together with the 16-bit position map and one descriptor {start, count} per (bucket, tile), stored bucket-major.  A
// bucket is then the list of its ~16-query runs, one per tile, 64 KB apart; the search kernel walks those runs (a warp
// owns 32 runs at a time: warp scan of the counts, lanes find their run with five shuffles) and overwrites every query
// with its answer IN PLACE, so the sectors it writes are the ones it has just read.  The last kernel streams each tile's
// answers back (TMA bulk load, double buffered) and applies the position map in shared memory: all of its global traffic
// is contiguous.  Per step: the queries are read once, nothing waits on another CTA (no look-back), 5 launches.
// ================================================================================================
constexpr int kPThreads = 1024;                // partition / un-permute CTA
#ifndef SST_BK_STHREADS
#define SST_BK_STHREADS 1024
#endif
constexpr int kSThreads = SST_BK_STHREADS;     // search CTA
constexpr int kPWarps = kPThreads / 32;
constexpr int kPItems = kTile / kPThreads;     // 16 queries per thread
constexpr unsigned kCntPad = 16;               // u16 of padding per counter row of the partition kernel (32 bytes = 8 banks)
constexpr unsigned kRunShift = 15;             // run descriptor = start | count << 15 (both <= kTile = 2^14: an empty bucket behind the last query starts AT kTile)
static_assert(kTile < (1 << kRunShift), "run descriptors hold a start of up to kTile in 15 bits");

__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct PartParams {
    BkView v;              // bucket lookup tables; v.bpt = buckets per thread of a 1024-thread CTA
    const uint32_t* qs;    // the caller's queries
    size_t nq;
    unsigned ntiles, ntp;  // tiles, and the row pitch of `runs` (tiles rounded up to a multiple of 32)
    uint32_t* qsort;       // [ntiles * kTile] every tile in bucket order (canonical queries)
    uint16_t* lpos;        // [nq] position of every query inside its sorted tile
    uint32_t* runs;        // [nbp][ntp] start | count << 15 of (bucket, tile)
    uint32_t* tot;         // [nbp] queries per bucket (zeroed before the launch)
    int tma_ok;            // qs is 16-byte aligned: full tiles move by bulk copies
};

struct PartCtx {
    uint32_t *s_in, *s_pk;
    uint16_t *cnt, *cntw;
    uint32_t* s_tile;
    uint64_t* bar_in;
    unsigned* s_warp;
    unsigned tid, lane, lt_mask, pitch16, i0;
};

// (thread 0) tile t of a 16-byte aligned batch -> the landing buffer, by two bulk copies
__device__ __forceinline__ void part_issue_load(const PartParams& p, uint32_t* s_in, uint64_t* bar_in, unsigned t) {
    if (p.tma_ok && t < p.ntiles && (size_t)(t + 1) * kTile <= p.nq) {
        mbar_expect_tx(bar_in, kTile * 4u);
        const char* src = reinterpret_cast<const char*>(p.qs + (size_t)t * kTile);
        tma_bulk_g2s(s_in, src, 32768u, bar_in);
        tma_bulk_g2s(reinterpret_cast<char*>(s_in) + 32768, src + 32768, 32768u, bar_in);
    }
}

// One tile of the partition kernel.  Compiled twice, for full tiles and for the partial last one: with `full` a compile-time
// constant the per-query validity tests vanish from the full-tile code (they were 11 of the 19 instructions per query of the
// position phase).
template <int BITS, int HYBRID, bool full>
__device__ __forceinline__ void part_tile(const PartParams& p, const PartCtx& c, const unsigned tile, unsigned& phase, unsigned (&acc_tot)[2]) {
    const BkView& v = p.v;
    uint32_t* const s_in = c.s_in;
    uint32_t* const s_pk = c.s_pk;
    uint16_t* const cnt = c.cnt;
    uint16_t* const cntw = c.cntw;
    uint32_t* const s_tile = c.s_tile;
    unsigned* const s_warp = c.s_warp;
    const unsigned tid = c.tid, lane = c.lane, lt_mask = c.lt_mask, pitch16 = c.pitch16, i0 = c.i0;
    const size_t tile_base = (size_t)tile * kTile;
    const unsigned tile_n = full ? (unsigned)kTile : (unsigned)(p.nq - tile_base);
    const bool via_tma = p.tma_ok && full;  // (block-uniform)
    uint32_t q[kPItems], pk[kPItems];
    if (via_tma) {
        mbar_wait(c.bar_in, phase);
        phase ^= 1u;
#pragma unroll
        for (int r = 0; r < kPItems; r++) q[r] = s_in[i0 + r * 32];
    } else {
#pragma unroll
        for (int r = 0; r < kPItems; r++) q[r] = i0 + r * 32 < tile_n ? __ldcs(p.qs + tile_base + i0 + r * 32) : 0u;
    }
    __syncthreads();  // s_in has been read by everyone
    if (tid == 0) {
        part_issue_load(p, c.s_in, c.bar_in, tile + gridDim.x);  // lands during the ranking below
        tma_store_wait_read();         // the previous tile's bulk store has read s_tile (= the counters zeroed next)
    }
    if (v.above) {  // (uniform) the partitioned layouts answer q > MAX with (MAX, n), not with the signed compare: note it
        uint32_t acc = 0;
#pragma unroll
        for (int r = 0; r < kPItems; r++) acc |= q[r];
        if (__any_sync(kFull, acc > kMax) && lane == 0) atomicOr(v.above, 1u);
    }
    {   // bucket of every query: ONE shared load and a compare (bk_pack_cell); a cell that two or more splitters share (skewed
        // keys) is flagged and settled afterwards through the global tables, outside the straight-line path (the branchy
        // form cost 30 instructions per query, a quarter of the kernel)
        uint32_t flags = 0;
#pragma unroll
        for (int r = 0; r < kPItems; r++) {
            q[r] = canonical(q[r]);
            const uint32_t e = s_pk[q[r] >> kBtShift];
            flags |= e;
            pk[r] = (e & 0x7ffu) + ((q[r] & ((1u << kBtShift) - 1u)) > (e >> 12) ? 1u : 0u);
        }
        if (flags & 0x800u) {  // (unrolled: a rolled loop would index q[] and pk[] dynamically and push both into local memory)
#pragma unroll
            for (int r = 0; r < kPItems; r++)
                if (s_pk[q[r] >> kBtShift] & 0x800u) pk[r] = bk_bucket(v.bt, v.split, q[r]);
        }
        if (!full) {
#pragma unroll
            for (int r = 0; r < kPItems; r++)
                if (i0 + r * 32 >= tile_n) pk[r] = 0xffffffffu;
        }
    }
    __syncthreads();  // thread 0 has seen the store's reads complete
    {   // zero the per-warp counters
        uint4* c4 = reinterpret_cast<uint4*>(cnt);
        const unsigned n16 = kPWarps * pitch16 / 8;
        for (unsigned i = tid; i < n16; i += kPThreads) c4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    rank_items<BITS, full, HYBRID, kPItems>(cntw, pk, lane, lt_mask);
    __syncthreads();
    // Per bucket: exclusive scan of the counters over the warps, the bucket's run {start, count} of this tile, and the
    // bucket's start folded into the per-warp bases.  Four adjacent lanes share a group of four adjacent buckets (one
    // 8-byte access = the four 16-bit counters of one warp) and take eight warps each (lane j: warps j, j + 4, ...), so
    // the 64 KB counter matrix is read twice and written once in 8-byte accesses (it was 4 x 32 two-byte accesses per
    // thread).  Rows are padded by 32 bytes: the four lanes of a group then hit banks 8 apart and a half-warp's sixteen
    // 8-byte accesses cover all 32 banks once (unpadded rows put the four lanes on the same banks: measured +7 %).
    {
        const unsigned j = tid & 3u, pitch = pitch16 / 4;  // quarter of the warps; row pitch in uint2
        // buckets are laid out in the order b = tid (k = 0), then b = tid + 1024 (k = 1): one k after the other, each with its
        // own block scan, so that only one set of partial sums is live next to the 32 query / bucket registers
        unsigned run_total = 0;
#pragma unroll
        for (unsigned k = 0; k < 2; k++)
            if (k < v.bpt) {
                uint2* c2 = reinterpret_cast<uint2*>(cnt) + (tid >> 2) + k * (kPThreads / 4) + (size_t)j * pitch;
                unsigned rx = 0, ry = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint2 c = c2[(size_t)(4 * i) * pitch];
                    rx += c.x & 0x07ff07ffu;  // (claim tags masked off; 16-bit halves cannot carry: every sum is <= kTile)
                    ry += c.y & 0x07ff07ffu;
                }
                unsigned px = rx, py = ry;  // inclusive scan over the four quarters
                unsigned yx = __shfl_up_sync(kFull, px, 1), yy = __shfl_up_sync(kFull, py, 1);
                if (j >= 1) { px += yx; py += yy; }
                yx = __shfl_up_sync(kFull, px, 2); yy = __shfl_up_sync(kFull, py, 2);
                if (j >= 2) { px += yx; py += yy; }
                const unsigned ex_x = px - rx, ex_y = py - ry;
                const unsigned tx = __shfl_sync(kFull, px, lane | 3u), ty = __shfl_sync(kFull, py, lane | 3u);  // the group's four totals
                const unsigned gtot = (tx & 0xffffu) + (tx >> 16) + (ty & 0xffffu) + (ty >> 16);
                unsigned total;
                if (k) __syncthreads();  // s_warp of the previous scan has been read
                const unsigned base = run_total + block_excl_scan(j == 0 ? gtot : 0u, s_warp, &total);
                run_total += total;
                const unsigned s0 = __shfl_sync(kFull, base, lane & ~3u);  // start of the group's first bucket
                const unsigned s1 = s0 + (tx & 0xffffu), s2 = s1 + (tx >> 16), s3 = s2 + (ty & 0xffffu);
                const unsigned st = j == 0 ? s0 : j == 1 ? s1 : j == 2 ? s2 : s3;  // this thread's bucket: tid + 1024 k
                const unsigned tt = j == 0 ? (tx & 0xffffu) : j == 1 ? (tx >> 16) : j == 2 ? (ty & 0xffffu) : (ty >> 16);
                p.runs[(size_t)(tid + k * kPThreads) * p.ntp + tile] = st | (tt << kRunShift);
                acc_tot[k] += tt;
                rx = (s0 | (s1 << 16)) + ex_x;  // bucket start + queries of the bucket in earlier warps
                ry = (s2 | (s3 << 16)) + ex_y;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint2 c = c2[(size_t)(4 * i) * pitch];
                    c2[(size_t)(4 * i) * pitch] = make_uint2(rx, ry);
                    rx += c.x & 0x07ff07ffu;
                    ry += c.y & 0x07ff07ffu;
                }
            }
    }
    __syncthreads();
    // final position of every query inside the sorted tile; the position map goes out as it is computed
    uint16_t* tl = p.lpos + tile_base + i0;
#pragma unroll
    for (int r = 0; r < kPItems; r++) {
        const bool valid = full || i0 + r * 32 < tile_n;
        const unsigned pos = valid ? (unsigned)cntw[pk[r] & 0xffffu] + (pk[r] >> 16) : 0u;
        pk[r] = pos;
        if (valid) tl[r * 32] = (uint16_t)pos;
    }
    __syncthreads();  // every counter has been read: the sorted tile may overwrite them
#pragma unroll
    for (int r = 0; r < kPItems; r++)
        if (full || i0 + r * 32 < tile_n) s_tile[pk[r]] = q[r];
    if (via_tma) {
        fence_proxy_async();  // generic-proxy writes to shared memory -> visible to the bulk-copy engine
        __syncthreads();
        if (tid == 0) {
            char* dst = reinterpret_cast<char*>(p.qsort + tile_base);
            tma_bulk_s2g(dst, s_tile, 32768u);
            tma_bulk_s2g(dst + 32768, reinterpret_cast<char*>(s_tile) + 32768, 32768u);
            tma_store_commit();
        }
    } else {
        __syncthreads();
        for (unsigned i = tid; i < tile_n; i += kPThreads) p.qsort[tile_base + i] = s_tile[i];
    }
}

template <int BITS, int HYBRID>
__global__ void __launch_bounds__(kPThreads, 1)
bk_part_kernel(const PartParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* s_in = reinterpret_cast<uint32_t*>(smem_raw);            // [kTile] landing buffer of the next tile (TMA)
    uint32_t* s_pk = s_in + kTile;                                      // [kBtCells] packed bucket table
    uint16_t* cnt = reinterpret_cast<uint16_t*>(s_pk + kBtCells);      // [kPWarps][nbp + 16] per-warp counters (rows padded by 32 bytes) ...
    uint32_t* s_tile = reinterpret_cast<uint32_t*>(cnt);               // ... reused as the sorted tile [kTile]
    __shared__ __align__(8) uint64_t bar_in;
    __shared__ unsigned s_warp[kPWarps + 1];
    const BkView& v = p.v;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned pitch16 = v.nbp + kCntPad;  // counter row pitch in u16: 32 bytes of padding shift every row by 8 banks (see the scan below)
    uint16_t* cntw = cnt + (size_t)warp * pitch16;
    for (unsigned i = tid; i < (unsigned)kBtCells; i += kPThreads) s_pk[i] = bk_pack_cell(v.bt, v.split, v.nb, i);
    if (tid == 0) {
        mbar_init(&bar_in, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) part_issue_load(p, s_in, &bar_in, blockIdx.x);
    unsigned phase = 0;
    unsigned acc_tot[2] = {0, 0};  // this CTA's queries in buckets tid and tid + 1024, over all of its tiles
    const unsigned i0 = warp * (kPItems * 32) + lane;  // this thread's queries: i0 + 32 r
    const PartCtx ctx{s_in, s_pk, cnt, cntw, s_tile, &bar_in, s_warp, tid, lane, lt_mask, pitch16, i0};
    for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        if ((size_t)(tile + 1) * kTile <= p.nq) part_tile<BITS, HYBRID, true>(p, ctx, tile, phase, acc_tot);
        else part_tile<BITS, HYBRID, false>(p, ctx, tile, phase, acc_tot);
    }
#pragma unroll
    for (unsigned k = 0; k < 2; k++)
        if (k < v.bpt && acc_tot[k]) atomicAdd(p.tot + tid + k * kPThreads, acc_tot[k]);
    if (tid == 0) tma_store_wait_all();  // shared memory must outlive the last bulk store
}

// ---- plan: the work items of the search kernel ------------------------------------------------------
// Work items {bucket, first tile, end tile, 0}: the tiles of a bucket whose exclusive query prefix falls into the same
// multiple of `chunk` form one item (a run holds at most kTile <= chunk queries, so no multiple is skipped); a bucket of
// tot queries owns tot / chunk + 1 consecutive item slots (0 when empty), a slot no tile maps to stays empty.
// ctrl[0] = work counter, ctrl[1] = number of items.  One CTA per bucket: thread i owns a contiguous span of the bucket's
// tiles (all of its descriptor loads are in flight at once), a block scan gives the span's query prefix.
constexpr int kPlanThreads = 256;
__global__ void __launch_bounds__(kPlanThreads, 8)
bk_items_kernel(const uint32_t* __restrict__ runs, const uint32_t* __restrict__ tot, unsigned ntiles, unsigned ntp, unsigned nb,
                unsigned chunk_log2, uint4* __restrict__ items, unsigned* __restrict__ ctrl) {
    __shared__ unsigned s_warp[kPlanThreads / 32 + 1];
    const unsigned b = blockIdx.x, tid = threadIdx.x;
    // first item slot of this bucket: the slots of the buckets before it
    unsigned before = 0;
    for (unsigned j = tid; j < b; j += kPlanThreads) { const unsigned t = __ldg(tot + j); before += t ? (t >> chunk_log2) + 1u : 0u; }
    unsigned ibase_total;
    (void)block_excl_scan(before, s_warp, &ibase_total);
    const unsigned ibase = ibase_total;
    const unsigned mine = __ldg(tot + b), nmine = mine ? (mine >> chunk_log2) + 1u : 0u;
    if (b == nb - 1 && tid == 0) { ctrl[0] = 0; ctrl[1] = ibase + nmine; }
    if (!nmine) return;
    for (unsigned k = tid; k < nmine; k += kPlanThreads) items[ibase + k] = make_uint4(b, 0, 0, 0);
    __syncthreads();  // (also: s_warp may be reused)
    const uint32_t* row = runs + (size_t)b * ntp;
    const unsigned span = (ntiles + kPlanThreads - 1) / kPlanThreads;  // tiles per thread: 24 at 10^8 queries, 256 at 2^30
    const unsigned t_begin = min(ntiles, tid * span), t_end = min(ntiles, t_begin + span);
    // (two passes over the span: the counts are read again after the scan -- L1 / L2 hits -- instead of being held in
    // registers, which had limited the kernel to two CTAs per SM: 3.5 waves of CTAs, 27 us; now one wave)
    unsigned sum = 0;
#pragma unroll 8
    for (unsigned t = t_begin; t < t_end; t++) sum += __ldg(row + t) >> kRunShift;
    unsigned total;
    unsigned e = block_excl_scan(sum, s_warp, &total);  // queries of the bucket in earlier tiles
    // the item of the tile before this span (0xffffffff at the very start): the last tile of the previous span with the same rule
    unsigned prev = t_begin == 0 ? 0xffffffffu : 0u;
    if (t_begin > 0 && t_begin < ntiles) {
        // exclusive prefix of tile t_begin - 1 = e - count(t_begin - 1)
        prev = (e - (__ldg(row + t_begin - 1) >> kRunShift)) >> chunk_log2;
    }
#pragma unroll 4
    for (unsigned t = t_begin; t < t_end; t++) {
        const unsigned item = e >> chunk_log2;
        if (item != prev) {
            items[ibase + item].y = t;
            if (prev != 0xffffffffu) items[ibase + prev].z = t;
        }
        prev = item;
        e += __ldg(row + t) >> kRunShift;
        if (t == ntiles - 1) items[ibase + item].z = ntiles;
    }
}

// ---- search over runs, in place --------------------------------------------------------------------
// One work item = (bucket, tile range).  The bucket's separators and jump table are staged in shared memory by 1-D TMA bulk
// copies; the item's tiles are taken 32 at a time (a "group": lane i holds run i's {start, count}, a warp scan gives every
// run its offset in the group's flattened query sequence, the lane that handles flattened query k finds its run among the
// next few offsets by broadcast shuffles), warp w takes groups w, w + 32, ...  Each answer overwrites its query.
//
// Everything that does not need the staged tables runs ahead of them, because nothing else hides a DRAM round trip with
// 32 warps per SM (ncu before: 10 % of the warp samples sat in the group set-up, 12 % at the item boundary):
//   * the queries of round k + 1 are loaded while round k is answered -- also across a group boundary: the last round of a
//     group sets up the warp's next group (its descriptors were loaded one group earlier) and loads its first queries;
//   * the first warp that finishes an item fetches the next work item (atomic counter, item record, bucket record) into
//     shared memory, so that after the barrier the bulk copies start at once, and every warp sets up its first group and
//     loads its first queries of the new item BEFORE it waits for the copies.
template <bool WANT_IDX, int G, bool S16>
__global__ void __launch_bounds__(kSThreads, 1)
bk_search2_kernel(const BkSearchParams p, uint32_t* __restrict__ qsort, uint32_t* __restrict__ isort, const uint32_t* __restrict__ runs,
                  unsigned ntp, const uint4* __restrict__ items, unsigned* __restrict__ ctrl) {
    constexpr int U = (G == 8 ? 4 : 2) * (1024 / kSThreads);
    constexpr unsigned kSWarps = kSThreads / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // separators: [r] u32, or in 16-bit mode [r] u16 (read as pairs); then the jump table [cells + 8] u16
    uint32_t* s_sep = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* s_jump = reinterpret_cast<uint16_t*>(smem_raw + (size_t)p.r * (S16 ? 2u : 4u));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint4 s_it[2];      // the work item of this / the next iteration: {bucket, first tile, end tile, item number}
    __shared__ uint2 s_meta[2];    // its bucket record {lo, shift}
    __shared__ unsigned s_done;    // warps that have finished the current item
    uint64_t keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    // The bucket's leaf window is read ~3 times per sector (by this and the neighbouring work items): keep it in L2.  Measured:
    // evict_last on the leaf 0.927 -> 0.910 ms; evict-first run loads/stores 0.959 vs 0.910; a bulk L2 prefetch of the window at
    // the start of an item changes nothing (profiles/r2_search2_leaf_l2_prefetch_ab.log).
    auto ldleaf = [&](const uint32_t* a, uint32_t (&k)[8]) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=r"(k[0]), "=r"(k[1]), "=r"(k[2]), "=r"(k[3]), "=r"(k[4]), "=r"(k[5]), "=r"(k[6]), "=r"(k[7])
                     : "l"(a), "l"(keep));
    };
    // (a query is read once, before its own lane overwrites it with the answer: plain weak accesses are enough)
    auto ldq = [&](const uint32_t* a) -> uint32_t { uint32_t v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(a)); return v; };
    auto stq = [&](uint32_t* a, uint32_t v) { *a = v; };
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned nitems = ctrl[1];
    // (one thread) the next work item that holds tiles -> s_it[slot], s_meta[slot]; item number >= nitems: none left
    auto fetch_item = [&](unsigned slot) {
        uint4 it = make_uint4(0, 0, 0, 0xffffffffu);
        while (true) {
            const unsigned item = atomicAdd(&ctrl[0], 1u);
            if (item >= nitems) { it.w = 0xffffffffu; break; }
            it = __ldg(items + item);
            it.w = item;
            if (it.y < it.z) break;  // (an item slot no tile mapped to is skipped)
        }
        s_it[slot] = it;
        if (it.w != 0xffffffffu) s_meta[slot] = __ldg(p.meta + it.x);
    };
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_done = 0;
        fetch_item(0);
    }
    unsigned phase = 0, cur_b = 0xffffffffu, slot = 0;
    while (true) {
        __syncthreads();  // the previous item is finished (s_sep, s_jump may be overwritten) and s_it[slot] is published
        const uint4 it = s_it[slot];
        if (it.w == 0xffffffffu) break;
        const unsigned b = it.x, t0 = it.y, t1 = it.z;
        const bool stage = b != cur_b;  // (block-uniform)
        if (tid == 0) {
            s_done = 0;  // (every warp passes the barrier above before it can finish this item)
            if (stage) {  // 1-D TMA bulk copies (SASS UBLKCP), completion on the mbarrier
                const unsigned sep_bytes = p.r * (S16 ? 2u : 4u), jump_bytes = (p.cells + 8u) * 2u;
                const char* sep_src = S16 ? (const char*)(p.sep16 + (size_t)b * p.r) : (const char*)(p.sep + (size_t)b * p.r);
                mbar_expect_tx(&bar, sep_bytes + jump_bytes);
                for (unsigned off = 0; off < sep_bytes; off += 32768u)
                    tma_bulk_g2s((char*)s_sep + off, sep_src + off, min(32768u, sep_bytes - off), &bar);
                for (unsigned off = 0; off < jump_bytes; off += 32768u)
                    tma_bulk_g2s((char*)s_jump + off, (const char*)(p.jump + (size_t)b * (p.cells + 8u)) + off, min(32768u, jump_bytes - off), &bar);
            }
        }
        const uint2 mt = s_meta[slot];
        const uint32_t lo = mt.x;
        const unsigned sh = mt.y;
        const unsigned hbase32 = b * p.r, m8m1 = (unsigned)(p.m8 - 1);  // block numbers fit 32 bits: at most 2^30 / 8 blocks
        const uint32_t* row = runs + (size_t)b * ntp;
        const unsigned ngroups = (t1 - t0 + 31u) >> 5;
        // ---- this warp's groups: g_next = the next one to set up, d_next = this lane's descriptor of it (already loaded) ----
        auto load_desc = [&](unsigned g) -> uint32_t {
            const unsigned t = t0 + g * 32u + lane;
            return g < ngroups && t < t1 ? __ldg(row + t) : 0u;
        };
        unsigned g_next = warp;
        uint32_t d_next = load_desc(g_next);
        // state of the current group
        unsigned excl = 0, T = 0, rcur = 0;  // this lane's run starts at flattened query excl; T queries in the group's 32 runs
        uint32_t dlt = 0;                    // address of flattened query k of this lane's run = dlt + k
        // Sets up the warp's next non-empty group; false when it has none left.  (warp-uniform)
        auto advance = [&]() -> bool {
            while (g_next < ngroups) {
                const uint32_t d = d_next;
                const unsigned t = t0 + g_next * 32u + lane;
                g_next += kSWarps;
                d_next = load_desc(g_next);  // consumed one group later
                const unsigned c = d >> kRunShift;
                unsigned incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(kFull, incl, o); if (lane >= (unsigned)o) incl += y; }
                T = __shfl_sync(kFull, incl, 31);
                if (T == 0) continue;
                excl = incl - c;
                dlt = t * (unsigned)kTile + (d & ((1u << kRunShift) - 1u)) - excl;
                rcur = 0;
                return true;
            }
            return false;
        };
        // Address of flattened query k (k < T): the last run whose offset is <= k holds it (empty runs tie with their
        // successor).  Slots are visited in order, so the run of a slot's first query is carried along (`rcur`, warp-uniform)
        // and a slot of 32 queries usually crosses at most three run boundaries: the offsets of the next four runs come
        // from four INDEPENDENT broadcast shuffles (uniform source lane) instead of a five-deep chain of per-lane ones;
        // a slot that crosses more (short runs) takes the general five-step search.  (One REDUX.OR over "my run starts at
        // lane j of this slot" + popc + a shared-memory table of the non-empty runs was measured slower: 0.943 vs 0.883 ms.)
        auto locate = [&](unsigned kbase) -> uint32_t {  // kbase = the slot's first query (warp-uniform); this lane's is kbase + lane
            const unsigned k = min(kbase + lane, T - 1u), klast = min(kbase + 31u, T - 1u);
            // (LB = boundaries the fast path handles; 7 instead of 3 for the ~8-query runs of 2048 buckets was measured: search 1.768
            // vs 1.735 ms at 2^30 keys, no gain)
            constexpr int LB = 3;
            unsigned e[LB + 1];
#pragma unroll
            for (int i = 0; i <= LB; i++) e[i] = __shfl_sync(kFull, excl, min(rcur + 1u + (unsigned)i, 31u));
            unsigned r;
            if (rcur + (unsigned)LB + 1u > 31u || e[LB] > klast) {  // (warp-uniform) at most LB boundaries inside the slot
                r = rcur;
                unsigned adv = 0;
#pragma unroll
                for (int i = 0; i < LB; i++) {
                    const bool have = rcur + 1u + (unsigned)i <= 31u;
                    r += (have && e[i] <= k) ? 1u : 0u;
                    adv += (have && e[i] <= klast) ? 1u : 0u;
                }
                rcur += adv;
            } else {
                r = 0;
#pragma unroll
                for (unsigned s = 16; s; s >>= 1) {
                    const unsigned e = __shfl_sync(kFull, excl, r + s);
                    if (e <= k) r += s;
                }
                rcur = __shfl_sync(kFull, r, 31);
            }
            return __shfl_sync(kFull, dlt, r) + k;
        };
        uint32_t an[U], qn[U];  // next round: addresses and queries (their loads stay in flight while this round is answered)
        unsigned vn = 0;        // bit u = slot u of the next round holds a query for this lane (carried along with the round: recomputing
                                // it from k0 and T at the top of the round cost 0.03 ms of the stage, profiles/r2_search2_units_ab.log)
        auto prefetch = [&](unsigned k0) {  // the round that starts at flattened query k0 of the current group
            vn = 0;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned kb = k0 + u * 32u;
                an[u] = kb < T ? locate(kb) : 0u;
                const bool v = kb + lane < T;
                qn[u] = v ? ldq(qsort + an[u]) : lo;
                vn |= (v ? 1u : 0u) << u;
            }
        };
        bool more = advance();
        if (more) prefetch(0);  // (before the wait below: the first queries of the item travel while the tables are staged)
        if (stage) {
            mbar_wait(&bar, phase);
            phase ^= 1u;
            cur_b = b;
        }
        unsigned k0 = 0;
        while (more) {
            uint32_t q[U], ad[U];
            unsigned a[U];
            const unsigned vm = vn;  // bit u = slot u of this round holds a query for this lane
#pragma unroll
            for (int u = 0; u < U; u++) {
                q[u] = qn[u];
                ad[u] = an[u];
            }
            // next round: of this group, or the first one of the warp's next group
            k0 += 32u * U;
            if (k0 >= T) { more = advance(); k0 = 0; }
            if (more) prefetch(k0);
            // rank among the bucket's separators: the jump cell gives the first candidate; later separators are probed only
            // while they are below the query (37 % / 10 % / 2 % of the lanes for uniform keys).  No upper end is needed:
            // every separator of a later cell is above q and the bucket's last separator is >= q.
            if constexpr (S16) {
                // 16-bit mode (leaf levels above 2^28 slots): 65536 separators per bucket as 16-bit offsets inside their jump cell
                // (a cell is at most 2^16 keys wide: 32768 cells over a range below 2^31), two per cell on average.  A cell's
                // separators are the only ones whose offsets compare with the query's, so both ends of the cell are read; six
                // offsets from the cell's start on come in three aligned 32-bit loads, more are needed by ~0.4 % of the queries.
                const uint32_t* s_sep32 = s_sep;  // (pairs of u16)
                const uint16_t* s_sep16 = reinterpret_cast<const uint16_t*>(s_sep);
                const unsigned rw = p.r / 2u - 1u;
                unsigned l[U], h[U];
                uint32_t qr[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t d = q[u] - lo;
                    const unsigned x = d >> sh;
                    qr[u] = d & ((1u << sh) - 1u);
                    l[u] = s_jump[x];
                    h[u] = s_jump[x + 1u];
                }
                uint32_t w0[U], w1[U], w2[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const unsigned wl = l[u] >> 1;
                    w0[u] = s_sep32[min(wl, rw)];
                    w1[u] = s_sep32[min(wl + 1u, rw)];
                    w2[u] = s_sep32[min(wl + 2u, rw)];
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const unsigned base = l[u] & ~1u;
                    const uint32_t v[6] = {w0[u] & 0xffffu, w0[u] >> 16, w1[u] & 0xffffu, w1[u] >> 16, w2[u] & 0xffffu, w2[u] >> 16};
                    unsigned pos = l[u];
#pragma unroll
                    for (unsigned j = 0; j < 6u; j++)  // (offsets are sorted inside a cell: the ones below the query's come first)
                        pos += (base + j >= l[u] && base + j < h[u] && v[j] < qr[u]) ? 1u : 0u;
                    if (pos == base + 6u && pos < h[u]) {  // rare: more than five separators of the cell below the query
                        unsigned hh = h[u];
                        while (pos < hh) {
                            const unsigned m = (pos + hh) >> 1;
                            if (s_sep16[m] < qr[u]) pos = m + 1; else hh = m;
                        }
                    }
                    a[u] = pos;
                }
            } else {
            unsigned l[U];
            uint32_t s0[U], s1[U], s2[U];
#pragma unroll
            for (int u = 0; u < U; u++) l[u] = s_jump[(q[u] - lo) >> sh];
            // three separators from the jump cell on, loaded for every query at once (3 x U independent loads in flight; with
            // one branch per probe the U chains ran one after the other); a fourth is needed by ~1 % of the queries
#pragma unroll
            for (int u = 0; u < U; u++) s0[u] = s_sep[min(l[u], p.r - 1u)];
#pragma unroll
            for (int u = 0; u < U; u++) s1[u] = s_sep[min(l[u] + 1u, p.r - 1u)];
#pragma unroll
            for (int u = 0; u < U; u++) s2[u] = s_sep[min(l[u] + 2u, p.r - 1u)];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool c0 = s0[u] < q[u], c1 = c0 && s1[u] < q[u], c2 = c1 && s2[u] < q[u];
                unsigned pos = l[u] + (c0 ? 1u : 0u) + (c1 ? 1u : 0u) + (c2 ? 1u : 0u);
                if (c2) {
                    unsigned hh = s_jump[((q[u] - lo) >> sh) + 1u];
                    if (hh > pos + 8u) {
                        while (pos < hh) {
                            const unsigned m = (pos + hh) >> 1;
                            if (s_sep[m] < q[u]) pos = m + 1; else hh = m;
                        }
                    } else {
                        while (pos < hh && s_sep[pos] < q[u]) pos++;
                    }
                }
                a[u] = pos;
            }
            }
            // the block of G keys that holds the answer: one 32-byte sector (two for G = 16).  Block numbers fit 32 bits
            // (at most 2^30 / 8 blocks), so the address is one 32 x 32 -> 64 multiply-add.
            uint32_t ks[U][G];
            unsigned hn[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                hn[u] = hbase32 + a[u];
                const unsigned hc = min(hn[u], m8m1);
                const uint32_t* src = p.leaf + (size_t)hc * (unsigned)G;
                ldleaf(src, reinterpret_cast<uint32_t(&)[8]>(ks[u][0]));
                if constexpr (G == 16) ldleaf(src + 8, reinterpret_cast<uint32_t(&)[8]>(ks[u][8]));
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                // The keys are sorted and at most MAX = 2^31 - 1, the query is canonical (<= MAX): the first key >= q is the
                // one with the smallest difference key - q among the non-negative ones, and a negative difference wraps to
                // >= 2^31 + 1.  One subtract and one unsigned min per key (it was a compare, a conditional increment and a
                // compare + select per key: 45 -> 17 instructions per query).
                uint32_t md = ks[u][0] - q[u];
#pragma unroll
                for (int e = 1; e < G; e++) md = min(md, ks[u][e] - q[u]);
                const bool none = md > 0x7fffffffu || hn[u] > m8m1;  // above every key of the block (only past the last key) / past the end
                const uint32_t val = none ? kMax : q[u] + md;
                if (vm >> u & 1u) {
                    stq(qsort + ad[u], val);
                    if constexpr (WANT_IDX) {
                        unsigned cc = 0;
#pragma unroll
                        for (int e = 0; e < G; e++) cc += ks[u][e] < q[u] ? 1u : 0u;
                        unsigned long long pos = (unsigned long long)hn[u] * (unsigned)G + cc;
                        if (none || pos > p.n) pos = p.n;
                        stq(isort + ad[u], (uint32_t)pos);
                    }
                }
            }
        }
        // the first warp to get here has time to spare: it fetches the next work item for everybody
        if (lane == 0 && atomicAdd(&s_done, 1u) == 0u) fetch_item(slot ^ 1u);
        slot ^= 1u;
    }
}

// ---- un-permute: answers of a tile (bucket order) -> the caller's order ------------------------------
// Every global access is contiguous: the tile's 64 KB of answers arrive by TMA bulk copies (two buffers: the next tile
// lands while this one is permuted), the position map is read with 8-byte loads one tile ahead, the output leaves in
// 16-byte stores; the permutation itself is a random read of shared memory.
template <typename OutT>
__global__ void __launch_bounds__(kPThreads, 1)
bk_unperm_kernel(const uint32_t* __restrict__ res, const uint16_t* __restrict__ lpos, size_t nq, unsigned ntiles, int tma_ok,
                 OutT* __restrict__ dst) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* s_buf = reinterpret_cast<uint32_t*>(smem_raw);  // [2][kTile]
    __shared__ __align__(8) uint64_t bars[2];
    const unsigned tid = threadIdx.x;
    constexpr int kVec = kPItems / 4;  // groups of 4 consecutive outputs per thread
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto full_tile = [&](unsigned t) { return t < ntiles && (size_t)(t + 1) * kTile <= nq; };
    auto issue_load = [&](unsigned t, unsigned buf) {  // (thread 0)
        if (tma_ok && full_tile(t)) {
            mbar_expect_tx(&bars[buf], kTile * 4u);
            const char* src = reinterpret_cast<const char*>(res + (size_t)t * kTile);
            char* d = reinterpret_cast<char*>(s_buf + (size_t)buf * kTile);
            tma_bulk_g2s(d, src, 32768u, &bars[buf]);
            tma_bulk_g2s(d + 32768, src + 32768, 32768u, &bars[buf]);
        }
    };
    uint2 ln[kVec];
    auto load_map = [&](unsigned t) {
        if (tma_ok && full_tile(t)) {
            const uint2* l2 = reinterpret_cast<const uint2*>(lpos + (size_t)t * kTile);
#pragma unroll
            for (int r = 0; r < kVec; r++) ln[r] = __ldcs(l2 + r * kPThreads + tid);
        }
    };
    if (tid == 0) issue_load(blockIdx.x, 0);
    load_map(blockIdx.x);
    unsigned ph[2] = {0, 0}, buf = 0;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1u) {
        const size_t tile_base = (size_t)tile * kTile;
        const bool full = (size_t)(tile + 1) * kTile <= nq;
        uint32_t* sb = s_buf + (size_t)buf * kTile;
        if (tid == 0) issue_load(tile + gridDim.x, buf ^ 1u);  // that buffer was released by the barrier that ended the previous tile
        if (tma_ok && full) {
            uint2 l[kVec];
#pragma unroll
            for (int r = 0; r < kVec; r++) l[r] = ln[r];
            load_map(tile + gridDim.x);
            mbar_wait(&bars[buf], ph[buf]);
            ph[buf] ^= 1u;
#pragma unroll
            for (int r = 0; r < kVec; r++) {
                const uint32_t v0 = sb[l[r].x & 0xffffu], v1 = sb[l[r].x >> 16], v2 = sb[l[r].y & 0xffffu], v3 = sb[l[r].y >> 16];
                OutT* d = dst + tile_base + (size_t)(r * kPThreads + tid) * 4;
                if constexpr (sizeof(OutT) == 4) {
                    __stcs(reinterpret_cast<uint4*>(d), make_uint4(v0, v1, v2, v3));
                } else {
                    __stcs(reinterpret_cast<ulonglong2*>(d), make_ulonglong2(v0, v1));
                    __stcs(reinterpret_cast<ulonglong2*>(d) + 1, make_ulonglong2(v2, v3));
                }
            }
        } else {  // partial last tile, or buffers that are not 16-byte aligned
            const unsigned tile_n = (unsigned)min((size_t)kTile, nq - tile_base);
            for (unsigned i = tid; i < tile_n; i += kPThreads) sb[i] = __ldcs(res + tile_base + i);
            __syncthreads();
            for (unsigned i = tid; i < tile_n; i += kPThreads) __stcs(dst + tile_base + i, (OutT)sb[lpos[tile_base + i]]);
        }
        __syncthreads();  // sb is free for the bulk load of tile + 2 * gridDim.x
    }
}

// ------------------------------------------------------------------------------------------------
// auxiliary arrays (index build time)
// ------------------------------------------------------------------------------------------------
__global__ void bk_pad_kernel(uint32_t* __restrict__ a, size_t from, size_t to) {
    for (size_t i = from + threadIdx.x; i < to; i += blockDim.x) a[i] = kMax;
}

// sep[m] = last key of block m of g keys (leaf slot g*m + g - 1) for m < m8, 0xffffffff beyond
__global__ void bk_sep_kernel(const uint32_t* __restrict__ leaf, unsigned long long m8, unsigned long long total, unsigned g,
                              uint32_t* __restrict__ sep) {
    for (unsigned long long m = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; m < total;
         m += (unsigned long long)gridDim.x * blockDim.x)
        sep[m] = m < m8 ? leaf[m * g + g - 1] : 0xffffffffu;
}
// split[0] = 0, split[b] = sep[b*r - 1] (last key before bucket b), split[nb] = MAX
__global__ void bk_split_kernel(const uint32_t* __restrict__ sep, unsigned nb, unsigned r, uint32_t* __restrict__ split) {
    for (unsigned b = blockIdx.x * blockDim.x + threadIdx.x; b <= nb; b += gridDim.x * blockDim.x)
        split[b] = b == 0 ? 0u : (b == nb ? kMax : sep[(size_t)b * r - 1]);
}
// bt[c] = number of splitters split[1..nb-1] whose top bits are < c, c in [0, kBtCells]; bit 15: >= 2 splitters in cell c
__global__ void bk_bt_kernel(const uint32_t* __restrict__ split, unsigned nb, uint16_t* __restrict__ bt) {
    for (unsigned c = blockIdx.x * blockDim.x + threadIdx.x; c < (unsigned)kBtStride; c += gridDim.x * blockDim.x) {
        unsigned cnt[2];
        for (int k = 0; k < 2; k++) {
            unsigned l = 0, h = nb - 1;  // count over i in [1, nb): index i = l + 1
            while (l < h) {
                const unsigned m = (l + h) >> 1;
                if ((split[m + 1] >> kBtShift) < c + k) l = m + 1; else h = m;
            }
            cnt[k] = l;
        }
        bt[c] = c <= (unsigned)kBtCells ? (uint16_t)(cnt[0] | (cnt[1] - cnt[0] >= 2u ? 0x8000u : 0u)) : 0;
    }
}
// Per bucket: lo = split[b], shift = smallest s with (split[b+1] - lo) >> s < r, and
// jump[x] = number of the bucket's separators with (sep - lo) >> s < x, x in [0, r].
__global__ void __launch_bounds__(256)
bk_jump_kernel(const uint32_t* __restrict__ sep, const uint32_t* __restrict__ split, unsigned r, unsigned long long m8,
               uint16_t* __restrict__ jump, uint2* __restrict__ meta) {
    const unsigned b = blockIdx.x;
    const uint32_t lo = split[b], hi = split[b + 1];
    unsigned s = 0;
    while (((hi - lo) >> s) >= r) s++;
    if (threadIdx.x == 0) meta[b] = make_uint2(lo, s);
    const uint32_t* sb = sep + (size_t)b * r;
    const unsigned long long first = (unsigned long long)b * r;
    const unsigned valid = (unsigned)min((unsigned long long)r, m8 > first ? m8 - first : 0ull);
    for (unsigned x = threadIdx.x; x < r + 8u; x += blockDim.x) {
        unsigned l = 0, h = valid;
        while (l < h) {
            const unsigned m = (l + h) >> 1;
            if (((sb[m] - lo) >> s) < x) l = m + 1; else h = m;
        }
        jump[(size_t)b * (r + 8u) + x] = x <= r ? (uint16_t)l : 0;
    }
}

// 16-bit mode.  Per bucket: lo = split[b], shift = smallest s with (split[b+1] - lo) >> s < cells (<= 16, since cells = 32768 and
// the range is below 2^31), jump[x] = number of the bucket's separators with (sep - lo) >> s < x for x in [0, cells], saturated
// at 65535 (a query's own cell always starts below that: the bucket's last separator is >= every query of the bucket); and
// sep16[m] = the low s bits of sep[m] - lo, the separator's offset inside its cell.
__global__ void __launch_bounds__(256)
bk_jump16_kernel(const uint32_t* __restrict__ sep, const uint32_t* __restrict__ split, unsigned r, unsigned cells, unsigned long long m8,
                 uint16_t* __restrict__ jump, uint2* __restrict__ meta, uint16_t* __restrict__ sep16) {
    const unsigned b = blockIdx.x;
    const uint32_t lo = split[b], hi = split[b + 1];
    unsigned s = 0;
    while (((hi - lo) >> s) >= cells) s++;
    if (threadIdx.x == 0) meta[b] = make_uint2(lo, s);
    const uint32_t* sb = sep + (size_t)b * r;
    const unsigned long long first = (unsigned long long)b * r;
    const unsigned valid = (unsigned)min((unsigned long long)r, m8 > first ? m8 - first : 0ull);
    for (unsigned x = threadIdx.x; x < cells + 8u; x += blockDim.x) {
        unsigned l = 0, h = valid;
        while (l < h) {
            const unsigned m = (l + h) >> 1;
            if (((sb[m] - lo) >> s) < x) l = m + 1; else h = m;
        }
        jump[(size_t)b * (cells + 8u) + x] = x <= cells ? (uint16_t)min(l, 65535u) : 0;
    }
    for (unsigned m = threadIdx.x; m < r; m += blockDim.x) sep16[(size_t)b * r + m] = m < valid ? (uint16_t)((sb[m] - lo) & ((1u << s) - 1u)) : 0xffffu;
}

// Map-partitioned trees: a query above MAX has no part (partitioned_s_tree.rs:844: the prefix map has no such entry) and is
// answered with (MAX, n) by every kernel of this library; the pipeline canonicalises such queries like the plain tree's
// signed compare, so they are rewritten here.  Exits at once unless the rank stage saw one.
__global__ void bk_above_kernel(const unsigned* __restrict__ flag, const uint32_t* __restrict__ qs, size_t nq, uint32_t* __restrict__ vals,
                                unsigned long long* __restrict__ idx, unsigned long long n) {
    if (!*flag) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (size_t)gridDim.x * blockDim.x)
        if (qs[i] > kMax) {
            vals[i] = kMax;
            if (idx) idx[i] = n;
        }
}

// Simple / L1 / Overlapping: position in the flat leaf level -> index in the sorted array, as the layouts' own kernels
// report it (stree_search.cu): the keys of part p start at slot part_pos[p] and are part_start[p] .. part_start[p+1] - 1.
__global__ void bk_flat_index_kernel(const uint32_t* __restrict__ qs, size_t nq, unsigned long long* __restrict__ idx, unsigned shift,
                                     unsigned long long parts, const uint32_t* __restrict__ part_start,
                                     const unsigned long long* __restrict__ part_pos, unsigned long long n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = qs[i];
        const unsigned long long part = (unsigned long long)(q >> shift);
        unsigned long long index = n;
        if (part < parts) {
            const unsigned long long pos = idx[i], st = part_start[part], cnt = part_start[part + 1] - st, pp = part_pos[part];
            const unsigned long long off = pos > pp ? pos - pp : 0;
            index = st + (off < cnt ? off : cnt);
            if (index > n) index = n;
        }
        idx[i] = index;
    }
}

// ------------------------------------------------------------------------------------------------
// scratch buffers: one set per (host thread, device), grown on demand, released when the thread exits
// ------------------------------------------------------------------------------------------------
thread_local double g_stage_ms[5] = {-1, -1, -1, -1, -1};

template <typename T>
bool regrow(T*& p, size_t count) {
    cudaFree(p);
    p = nullptr;
    return SST_CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
}

}  // namespace

void free_bucket_aux(sst_index* idx) {
    cudaFree(idx->bk.d_dense); cudaFree(idx->bk.d_sep16); cudaFree(idx->bk.d_sep); cudaFree(idx->bk.d_split); cudaFree(idx->bk.d_bt); cudaFree(idx->bk.d_jump); cudaFree(idx->bk.d_meta);
    idx->bk = BkAux{};
}

// Builds the auxiliary arrays of the reordered-batch pipeline for a plain B=16 tree.  Returns false only
// on a CUDA error; an index the pipeline does not serve simply has bk.nb == 0.
bool build_bucket_aux(sst_index* idx, const uint32_t* d_sorted) {
    // plain B=16 trees, and the Map-partitioned layout: its leaf level is the sorted array itself, without per-part gaps
    // (partitioned_s_tree.rs:503), which is all the pipeline reads of an image
    // Simple / L1 / Overlapping: their leaf level is ONE non-decreasing flat array too -- the slots between two parts hold the
    // first key of the next non-empty part (:502-515), the tail MAX -- so the lower bound in it has the right value; its
    // position is turned into the sorted-array index afterwards (bk_flat_index_kernel).
    // Compact interleaves the levels of every part and pads each part's leaves with MAX (:283-307), so its image holds no flat
    // sorted leaf level: the pipeline reads a dense GPU-only copy of the keys instead (4 bytes per key on top of the image, next
    // to the dense copy of the upper levels the lane-group kernel already uses); when memory is short the layout simply stays on
    // the lane-group kernel.
    if (idx->variant == SST_EYTZINGER || idx->node_b != 16) return true;
    if (idx->variant == SST_COMPACT && (!d_sorted || !opt(OPT_BK_COMPACT))) return true;
    if (idx->n < (size_t)opt(OPT_BK_MIN_N)) return true;  // small trees are L2-resident: nothing to gain
    const bool flat_parts = idx->variant == SST_SIMPLE || idx->variant == SST_L1 || idx->variant == SST_OVERLAPPING;
    const size_t n_flat = flat_parts ? (size_t)idx->layer_blocks[idx->levels - 1] * 16 : idx->n;
    // separators per bucket: 16384 (two search CTAs per SM) up to 2^27 keys, 32768 (one 1024-thread CTA) above, so that the
    // partition has at most 1024 buckets up to 2^28 keys -- fewer buckets = longer runs and fewer ballot bits
    // keys per separator: 8 (one leaf sector per query) up to 2^29 keys, 16 (a whole node) up to 2^30
    // Above 2^28 slots 32768 32-bit separators per bucket would need more than 1024 buckets (8-query runs: partition and search
    // both suffer) or 16 keys per separator (two leaf sectors per query, half as many queries in flight per thread: the search took
    // 2.04 ms per 2^27 queries at 2^30 keys against 1.28 at 2^29).  The 16-bit mode keeps 8 keys per separator with 65536
    // separators per bucket in the same 192 KB: the separators are stored as 16-bit offsets inside their jump cell (bk_jump16_kernel).
    // Above 2^30 slots the 16-bit mode takes 16 keys per separator (two leaf sectors per query): 2048 buckets of 65536 separators
    // then cover 2^31 slots, which is every u32 tree this library can hold (n < 2^32 and keys <= MAX).
    const bool s16 = opt(OPT_BK_SEP16) == 1 || (opt(OPT_BK_SEP16) < 0 && opt(OPT_BK_G) <= 0 && opt(OPT_BK_R) <= 0 && div_ceil(n_flat, (size_t)8) > 32768ull * 1024ull);
    unsigned g = s16 ? ((opt(OPT_BK_G) == 16 || div_ceil(n_flat, (size_t)8) > 65536ull * 2048ull) ? 16u : 8u)
                     : opt(OPT_BK_G) > 0 ? (unsigned)opt(OPT_BK_G) : (div_ceil(n_flat, (size_t)8) > 32768ull * 2048ull ? 16u : 8u);
    if (g != 8 && g != 16) g = 8;
    const unsigned long long m8 = div_ceil(n_flat, (size_t)g);
    unsigned r = opt(OPT_BK_R) > 0 ? (unsigned)opt(OPT_BK_R) : (m8 > 16384ull * 1024ull ? 32768u : 16384u);
    if (r < 64 || r > 32768 || (r & (r - 1))) r = 16384;
    if (s16) r = 65536u;  // (fixed: 16-bit offsets need cells = r / 2 >= 2^31 / 2^16 whatever the bucket's key range)
    const unsigned cells = s16 ? r / 2u : r;
    const unsigned long long nb64 = div_ceil((size_t)m8, (size_t)r);
    if (nb64 > 2048) return true;  // (more than 2^31 slots: served by the rank-table kernel)
    BkAux& a = idx->bk;
    const unsigned nb = (unsigned)nb64;
    const unsigned nbp = (unsigned)(div_ceil((size_t)nb, (size_t)kThreads) * kThreads);
    unsigned bits = 0;
    while ((1u << bits) < nb) bits++;
    cudaStream_t st = thread_stream(idx->device);
    const uint32_t* leaf = idx->d_tree + idx->offsets[idx->levels - 1] * 16;
    if (idx->variant == SST_COMPACT) {
        // room for whole blocks of g keys: the tail reads as MAX like the padded leaf level of the other layouts
        const size_t words = (size_t)m8 * g;
        if (cudaMalloc(&a.d_dense, words * 4) != cudaSuccess) { (void)cudaGetLastError(); a.d_dense = nullptr; return true; }  // no room: lane-group kernel
        bk_pad_kernel<<<1, 32, 0, st>>>(a.d_dense, idx->n, words);  // the tail of the last block reads as MAX, like a padded leaf level
        if (!SST_CUDA_OK(cudaMemcpyAsync(a.d_dense, d_sorted, idx->n * 4, cudaMemcpyDeviceToDevice, st))) { free_bucket_aux(idx); return false; }
        leaf = a.d_dense;
    }
    const unsigned long long total = (unsigned long long)nb * r;
    bool ok = SST_CUDA_OK(cudaMalloc(&a.d_sep, total * 4)) && SST_CUDA_OK(cudaMalloc(&a.d_split, ((size_t)nb + 1) * 4)) &&
              SST_CUDA_OK(cudaMalloc(&a.d_bt, (size_t)kBtStride * 2)) && SST_CUDA_OK(cudaMalloc(&a.d_jump, (size_t)nb * (cells + 8) * 2)) &&
              SST_CUDA_OK(cudaMalloc(&a.d_meta, (size_t)nb * sizeof(uint2))) && (!s16 || SST_CUDA_OK(cudaMalloc(&a.d_sep16, total * 2)));
    if (ok) {
        bk_sep_kernel<<<(unsigned)std::min<size_t>(div_ceil((size_t)total, (size_t)256), (size_t)cur_sms() * 16), 256, 0, st>>>(leaf, m8, total, g, a.d_sep);
        bk_split_kernel<<<(unsigned)div_ceil((size_t)nb + 1, (size_t)256), 256, 0, st>>>(a.d_sep, nb, r, a.d_split);
        bk_bt_kernel<<<(unsigned)div_ceil((size_t)kBtStride, (size_t)256), 256, 0, st>>>(a.d_split, nb, a.d_bt);
        if (s16) bk_jump16_kernel<<<nb, 256, 0, st>>>(a.d_sep, a.d_split, r, cells, m8, a.d_jump, a.d_meta, a.d_sep16);
        else bk_jump_kernel<<<nb, 256, 0, st>>>(a.d_sep, a.d_split, r, m8, a.d_jump, a.d_meta);
        ok = SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (!ok) { free_bucket_aux(idx); return false; }
    if (s16) { cudaFree(a.d_sep); a.d_sep = nullptr; }  // (the 32-bit separators were only needed to build the 16-bit form)
    a.cells = cells;
    a.nb = nb; a.nbp = nbp; a.r = r; a.bits = bits; a.m8 = m8; a.g = g; a.n_flat = n_flat;
    return true;
}

bool bucketed_eligible(const sst_index* idx) { return idx->bk.nb != 0; }

// rank, plan, scatter, search, gather (ms) of this thread's last pipeline run under SST_BK_TIMING; returns the count
int last_stage_ms(double* out, int n) {
    int k = 0;
    for (; k < n && k < 5; k++) out[k] = g_stage_ms[k];
    return g_stage_ms[0] < 0 ? 0 : k;
}


// ------------------------------------------------------------------------------------------------
// V2 scratch: one set per (host thread, device); sized by sst_query_reserve or grown on first use
// ------------------------------------------------------------------------------------------------
namespace {
struct Scratch2 {
    int device = -1;
    size_t cap_q = 0, cap_idx = 0, cap_runs = 0, cap_items = 0;
    uint32_t *qsort = nullptr, *isort = nullptr, *runs = nullptr, *tot = nullptr;
    uint16_t* lpos = nullptr;
    uint4* items = nullptr;
    unsigned* ctrl = nullptr;
    cudaEvent_t done = nullptr;
    void release() {
        if (device < 0) return;
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return; }
        cudaFree(qsort); cudaFree(isort); cudaFree(runs); cudaFree(tot); cudaFree(lpos); cudaFree(items); cudaFree(ctrl);
        if (done) cudaEventDestroy(done);
        (void)cudaGetLastError();
        if (prev >= 0) cudaSetDevice(prev);
        *this = Scratch2{};
    }
    ~Scratch2() { release(); }
};
thread_local Scratch2 g_scratch2[64];

struct Shape2 { size_t sub; unsigned ntiles, ntp, nbp2; size_t runs, items; };
// Queries per pipeline run.  A larger run puts more queries on every leaf sector (at 2^30 keys a run of 2^27 queries touches every
// sector once: the whole 4 GB leaf level per run) but needs 6-10 bytes of scratch per query: up to 2^30 (BK_SUB_LOG2) when the
// scratch can be had, halved down to 2^27 when it cannot.
size_t sub_batch_max() { return (size_t)1 << opt(OPT_BK_SUB_LOG2); }
Shape2 shape2(const BkAux& a, size_t nq, size_t sub_max) {
    Shape2 h;
    h.sub = std::min(nq, sub_max);
    h.ntiles = (unsigned)div_ceil(h.sub, (size_t)kTile);
    h.ntp = (h.ntiles + 31u) & ~31u;
    h.nbp2 = (unsigned)(div_ceil((size_t)a.nb, (size_t)kPThreads) * kPThreads);
    h.runs = (size_t)h.nbp2 * h.ntp;
    h.items = h.sub / kMinChunk + a.nb + 2;
    return h;
}

// (needs_alloc != nullptr: only report whether anything would have to be allocated)
bool scratch2_ensure(Scratch2& s, int device, const Shape2& h, bool want_idx, bool* needs_alloc = nullptr) {
    const size_t nslots = (size_t)h.ntiles * kTile;
    const bool grow = !s.ctrl || nslots > s.cap_q || (want_idx && nslots > s.cap_idx) || h.runs > s.cap_runs || h.items > s.cap_items;
    if (needs_alloc) { *needs_alloc = grow; return true; }
    if (!grow) return true;
    s.device = device;
    if (!s.done && !SST_CUDA_OK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming))) return false;
    if (!s.ctrl && (!regrow(s.ctrl, 8) || !regrow(s.tot, 4096))) return false;
    if (nslots > s.cap_q) {
        s.cap_q = 0;
        if (!regrow(s.qsort, nslots) || !regrow(s.lpos, nslots)) return false;
        s.cap_q = nslots;
    }
    if (want_idx && nslots > s.cap_idx) {
        s.cap_idx = 0;
        if (!regrow(s.isort, nslots)) return false;
        s.cap_idx = nslots;
    }
    if (h.runs > s.cap_runs) {
        s.cap_runs = 0;
        if (!regrow(s.runs, h.runs)) return false;
        s.cap_runs = h.runs;
    }
    if (h.items > s.cap_items) {
        s.cap_items = 0;
        if (!regrow(s.items, h.items)) return false;
        s.cap_items = h.items;
    }
    return true;
}

template <int BITS>
void launch_part(const PartParams& pp, int sms, size_t smem, cudaStream_t st) {
    const unsigned grid = (unsigned)std::min<size_t>(pp.ntiles, (size_t)sms);
    if (BITS > 0 && opt(OPT_BK_HYBRID) != 0) {
        auto kern = bk_part_kernel<BITS, 32>;  // claims on every step but the first (see launch_rank)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kPThreads, smem, st>>>(pp);
    } else {
        auto kern = bk_part_kernel<BITS, 0>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kPThreads, smem, st>>>(pp);
    }
}

template <typename OutT>
void launch_unperm(int sms, cudaStream_t st, const uint32_t* res, const uint16_t* lpos, size_t nq, unsigned ntiles, OutT* dst) {
    const int tma_ok = ((uintptr_t)dst & 15u) == 0;
    auto kern = bk_unperm_kernel<OutT>;
    const size_t smem = (size_t)2 * kTile * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)std::min<size_t>(ntiles, (size_t)sms), kPThreads, smem, st>>>(res, lpos, nq, ntiles, tma_ok, dst);
}
}  // namespace

// Pre-sizes the calling thread's pipeline scratch for batches of up to nq queries on this index, so that later
// sst_query_device calls allocate nothing (asynchronous, graph-capturable).  No-op for an index the pipeline does not serve.
int reserve_bucketed(const sst_index* idx, size_t nq, bool want_idx) {
    const BkAux& a = idx->bk;
    if (!a.nb || nq == 0) return SST_OK;
    const int dev = idx->device;
    if (dev < 0 || dev >= 64) { set_error(SST_ERR_ARG, "device index out of range"); return SST_ERR_ARG; }
    if (!scratch2_ensure(g_scratch2[dev], dev, shape2(a, nq, sub_batch_max()), want_idx)) {
        set_error(SST_ERR_CAPACITY, "not enough device memory for the reordered-batch scratch buffers (6-10 bytes per query)");
        return SST_ERR_CAPACITY;
    }
    return SST_OK;
}
size_t bucketed_sub_batch() { return sub_batch_max(); }
void release_bucketed_scratch() {
    for (auto& s : g_scratch2) s.release();
}

static int launch_bucketed_v2(const sst_index* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_vals, unsigned long long* d_idx, cudaStream_t st) {
    const BkAux& a = idx->bk;
    const int dev = idx->device;
    if (dev < 0 || dev >= 64) { set_error(SST_ERR_ARG, "device index out of range"); return SST_ERR_ARG; }
    Scratch2& s = g_scratch2[dev];
    Shape2 h = shape2(a, nq, sub_batch_max());
    while (!scratch2_ensure(s, dev, h, d_idx != nullptr)) {  // out of device memory for the scratch buffers: smaller runs, down to 2^27 queries
        (void)cudaGetLastError();
        if (h.sub <= ((size_t)1 << 27)) {
            set_error(SST_ERR_CAPACITY, "not enough device memory for the reordered-batch scratch buffers (6-10 bytes per query)");
            return SST_ERR_CAPACITY;
        }
        clear_error();
        h = shape2(a, nq, h.sub / 2);
    }
    // scratch reuse across this thread's streams: a call waits for the previous one.  Not while `st` is being captured into a
    // CUDA graph: an event recorded outside the capture cannot be waited on there, and one recorded inside it would tie later
    // calls to the capture; the graph's own launches are ordered by the stream, replays are the caller's to order.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (!SST_CUDA_OK(cudaStreamIsCapturing(st, &cap))) return SST_ERR_CUDA;
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    if (!capturing && !SST_CUDA_OK(cudaStreamWaitEvent(st, s.done, 0))) return SST_ERR_CUDA;
    const int sms = sm_count(dev);
    const bool flat_parts = idx->variant == SST_SIMPLE || idx->variant == SST_L1 || idx->variant == SST_OVERLAPPING;
    const bool map_tree = idx->variant == SST_MAP || idx->variant == SST_COMPACT || flat_parts;  // partitioned: a query above MAX has no part -> (MAX, n)
    if (map_tree && !SST_CUDA_OK(cudaMemsetAsync(s.ctrl + 3, 0, 4, st))) return SST_ERR_CUDA;
    const size_t smem_part = (size_t)kTile * 4 + (size_t)kBtCells * 4 + std::max((size_t)kTile * 4, (size_t)kPWarps * (h.nbp2 + kCntPad) * 2);
    const size_t smem_search = (size_t)a.r * (a.d_sep16 ? 2 : 4) + ((size_t)a.cells + 8) * 2;
    const unsigned chunk_log2 = (unsigned)opt(OPT_BK_CHUNK2_LOG2);
    BkSearchParams sp{a.d_sep, a.d_jump, a.d_meta, a.d_dense ? a.d_dense : idx->d_tree + idx->offsets[idx->levels - 1] * 16, a.r, a.d_sep16, a.cells, a.m8, a.n_flat};
    const bool timing = opt(OPT_BK_TIMING) != 0;
    cudaEvent_t ev[8] = {};
    int nev = 0;
    auto mark = [&]() {
        if (!timing || nev >= 8) return;
        cudaEventCreate(&ev[nev]);
        cudaEventRecord(ev[nev++], st);
    };
    for (size_t off = 0; off < nq; off += h.sub) {
        const size_t cnt = std::min(h.sub, nq - off);
        const unsigned ntiles = (unsigned)div_ceil(cnt, (size_t)kTile);
        const uint32_t* qs = d_qs + off;
        PartParams pp{};
        pp.v = BkView{a.d_bt, a.d_split, a.nb, h.nbp2, h.nbp2 / kPThreads, map_tree ? s.ctrl + 3 : nullptr};
        pp.qs = qs; pp.nq = cnt; pp.ntiles = ntiles; pp.ntp = h.ntp;
        pp.qsort = s.qsort; pp.lpos = s.lpos; pp.runs = s.runs; pp.tot = s.tot;
        if (!SST_CUDA_OK(cudaMemsetAsync(s.tot, 0, (size_t)h.nbp2 * 4, st))) return SST_ERR_CUDA;
        pp.tma_ok = ((uintptr_t)qs & 15u) == 0;
        nev = 0;
        mark();
        switch (a.bits) {
#define SST_BK_PART(B) case B: launch_part<B>(pp, sms, smem_part, st); break;
            SST_BK_PART(0) SST_BK_PART(1) SST_BK_PART(2) SST_BK_PART(3) SST_BK_PART(4) SST_BK_PART(5)
            SST_BK_PART(6) SST_BK_PART(7) SST_BK_PART(8) SST_BK_PART(9) SST_BK_PART(10) SST_BK_PART(11)
#undef SST_BK_PART
            default: set_error(SST_ERR_UNSUPPORTED, "too many buckets"); return SST_ERR_UNSUPPORTED;
        }
        mark();
        bk_items_kernel<<<a.nb, kPlanThreads, 0, st>>>(s.runs, s.tot, ntiles, h.ntp, a.nb, chunk_log2, s.items, s.ctrl);
        mark();
        mark();  // (no scatter stage: kept so that the five stage times line up with the V1 report)
        {
            void (*kern)(const BkSearchParams, uint32_t*, uint32_t*, const uint32_t*, unsigned, const uint4*, unsigned*) =
                a.d_sep16 ? (a.g == 16 ? (d_idx ? bk_search2_kernel<true, 16, true> : bk_search2_kernel<false, 16, true>)
                                       : (d_idx ? bk_search2_kernel<true, 8, true> : bk_search2_kernel<false, 8, true>))
                : a.g == 16 ? (d_idx ? bk_search2_kernel<true, 16, false> : bk_search2_kernel<false, 16, false>) : (d_idx ? bk_search2_kernel<true, 8, false> : bk_search2_kernel<false, 8, false>);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_search);
            kern<<<sms, kSThreads, smem_search, st>>>(sp, s.qsort, s.isort, s.runs, h.ntp, s.items, s.ctrl);
        }
        mark();
        launch_unperm<uint32_t>(sms, st, s.qsort, s.lpos, cnt, ntiles, d_vals + off);
        if (d_idx) launch_unperm<unsigned long long>(sms, st, s.isort, s.lpos, cnt, ntiles, d_idx + off);
        if (flat_parts && d_idx)
            bk_flat_index_kernel<<<sms * 8, 256, 0, st>>>(qs, cnt, d_idx + off, (unsigned)idx->shift, (unsigned long long)idx->parts, idx->d_part_start,
                                                        idx->d_part_pos, (unsigned long long)idx->n);
        if (map_tree) bk_above_kernel<<<sms * 4, 256, 0, st>>>(s.ctrl + 3, qs, cnt, d_vals + off, d_idx ? d_idx + off : nullptr, (unsigned long long)idx->n);
        mark();
        if (timing && nev == 6 && cudaEventSynchronize(ev[5]) == cudaSuccess) {
            float t[5];
            for (int i = 0; i < 5; i++) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
            for (int i = 0; i < 5; i++) g_stage_ms[i] = t[i];
            if (opt(OPT_BK_TIMING) > 1) fprintf(stderr, "[sst] bucketed v2 nq=%zu nb=%u: partition %.3f plan %.3f - search %.3f unpermute %.3f ms\n", cnt, a.nb, t[0], t[1], t[3], t[4]);
        }
        for (int i = 0; i < nev; i++) cudaEventDestroy(ev[i]);
    }
    if (!SST_CUDA_OK(cudaGetLastError()) || (!capturing && !SST_CUDA_OK(cudaEventRecord(s.done, st)))) return SST_ERR_CUDA;
    return SST_OK;
}

int launch_bucketed(const sst_index* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_vals, unsigned long long* d_idx, cudaStream_t st) {
    if (!idx->bk.nb) { set_error(SST_ERR_UNSUPPORTED, "the reordered-batch pipeline serves plain and Map-partitioned B=16 trees of 2^22..2^30 keys"); return SST_ERR_UNSUPPORTED; }
    return launch_bucketed_v2(idx, d_qs, nq, d_vals, d_idx, st);
}

}  // namespace sst
