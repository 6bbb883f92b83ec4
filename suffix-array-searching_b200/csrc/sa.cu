// sa.cu -- suffix array construction, validation and pattern search on the GPU.
//
// Replaces (reference paths relative to suffix-array-searching/src):
//   SaNaive::build / SA::build        sa_search.rs:30-57, experiments.rs:19-38
//     (the libsais call at sa_search.rs:33 / experiments.rs:26 and the strict-order assertion
//      at sa_search.rs:36-38)
//   binary_search                      sa_search.rs:98-112, experiments.rs:51-64
//   cmp (16-byte SIMD compare)         sa_search.rs:346-374
//   the LCP-accelerated search that the reference only sketches (TODO at sa_search.rs:344-345)
//
// Search, two mappings (both bit-identical to the reference's binary_search):
//   * sa_search_thread_kernel (default): one thread per pattern plus a pivot-prefix table that
//     answers all but the last levels of the search with one 16-byte load per probe (see below);
//   * sa_search_kernel<PL>: a sub-warp of PL lanes per pattern; every probe loads sa[m] once
//     (broadcast within the group) and compares 4*PL-byte windows, each lane one (unaligned)
//     4-byte word, with __ballot_sync + ffs to find the first mismatching lane.  PL = 32 is the
//     warp-per-pattern mapping of the north star; it is the slowest one measured, because this
//     path is bound by the number of independent miss chains in flight, not by compare width.
// The `mlr` mode carries lcp(q, suffix(l-1)) and lcp(q, suffix(r)) and starts each comparison at
// their minimum.
// Texts over {0,1,2,3} (every BASELINE configuration) get two more GPU-only accelerators, used by the
// thread kernel: a k-mer table (the suffix-array range of every k-base prefix: SaNaive's `table`,
// sa_search.rs:59-85, which the reference hard-wires to p = 0) that replaces the first ~2k probes by
// one load and bounds the search for `hi`, and the next 15 bases of every suffix inlined next to its
// suffix-array entry ("Inlining values", todo.org:18-19), so that a probe inside a k-mer cell touches
// the text only for the suffix that matches.  Every path returns the same [lo, hi) and sa[lo].
//
// Construction: prefix doubling.  Suffixes are ranked by their first 7 symbols (9 bits each,
// 0 = past the end, so that a proper prefix sorts first exactly as Rust's slice ordering),
// then by (rank[i], rank[i+h]) pairs with h = 7, 14, 28, ... until all ranks are distinct.
// The pair sort uses cub::DeviceRadixSort (a sort primitive, like the reference's use of
// libsais); everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

struct sst_sa {
    int device = 0;
    size_t n = 0;
    uint8_t* d_text = nullptr;  // n bytes + 64 zero bytes
    uint32_t* d_sa = nullptr;   // n entries
    // Pivot-prefix table: node j (heap order, root = 1) of the implicit binary-search tree over
    // [0, n) holds the first 16 bytes (zero padded) of suffix(sa[m_j]).  GPU-only auxiliary.
    uint4* d_pivots = nullptr;
    int pivot_levels = 0;
    // k-mer table (texts over {0,1,2,3} only): kmer[x] = lower bound in the suffix array of the k-base string with 2-bit
    // code x, x in [0, 4^k]; the search of a pattern starts in [kmer[x], kmer[x+1]) instead of [0, n).  The reference has
    // the same idea as SaNaive's prefix `table` (sa_search.rs:59-85, hard-wired to p = 0 there) and packs bases with
    // string_value (util.rs:76-117).  GPU-only auxiliary.
    uint32_t* d_kmer = nullptr;
    int kmer_k = 0;
    // Inlined bases (with the k-mer table only): sax[i] = {sa[i], nx[i]}, nx = the 15 bases that follow the first k of
    // suffix(sa[i]) packed 2 bits each, first base most significant (string_value order, util.rs:76-117), bit 31 set when the
    // suffix has all of them.  Inside a k-mer cell every suffix shares the first k bases with the pattern, so a probe whose
    // next 15 bases differ from the pattern's is decided without touching the text: one sector instead of two or three
    // fills.  The reference has the idea as a TODO ("Inlining values", todo.org:18-19; btree_legacy.rs:4-131).  GPU-only auxiliary.
    uint2* d_sax = nullptr;
    // The same with 48 bases per suffix in 16-byte entries {sa, three words of 16 bases}, built instead of d_sax when memory
    // allows (16 bytes per suffix): a pattern of up to k + 48 bases is then answered without the text.  A suffix with fewer
    // than k + 48 bases has zero words; the reader tells from sa itself (the entries held 32 bases and a "32 present" word
    // until the third session of round 2: C5 9.66 -> 10.0 Gpat/s).
    uint4* d_saw = nullptr;
    // The text at 2 bits per base (with the k-mer table, i.e. for texts over {0,1,2,3}): word w = bases 16 w .. 16 w + 15, first base
    // most significant, zero padded.  When the inlined bases of an entry equal the pattern's and the pattern goes on, the rest
    // is compared here instead of in the byte text: the 52 bytes a 100-base pattern still needs shrink to 13, i.e. one 64-byte
    // fill instead of (usually) two.  n / 4 bytes; GPU-only auxiliary.
    uint32_t* d_text2 = nullptr;
    // Packed k-mer cells (with d_saw, k <= 14, n < 2^31, memory permitting): 128 bytes per cell = {start | overflow << 31, entries
    // 0..4} {-, entries 5..9}, an entry = {sa, code hi, code lo} -- the cell's suffix-array range AND its first ten {sa, 32 bases}
    // entries in ONE line, so that a pattern of k .. k + 31 bases is answered by a single random DRAM access instead of two (k-mer
    // cell, then the entries); the second half is read (an L2 hit) only when the first five entries do not decide.  Unused
    // entries have sa = 0xffffffff; a cell with more than ten suffixes, or with a suffix too close to the text's end to have
    // all 32 bases, is flagged and takes the ordinary path.  GPU-only auxiliary; no result depends on it.
    uint4* d_cells = nullptr;
};

namespace sst {
namespace {

constexpr int kThreads = 256;
#ifndef SST_SA_SEARCH_THREADS
#define SST_SA_SEARCH_THREADS 128   // (one pattern per thread: smaller blocks retire sooner; C5 256 / 128 / 64 threads 8.78 / 8.97 / 9.02 Gpat/s, C3 19.1 / 19.0 / 18.8)
#endif
constexpr int kSearchThreads = SST_SA_SEARCH_THREADS;  // block of sa_search_thread_kernel (MINB counts blocks of 256 threads' worth)
#ifndef SST_SA_MIN_BLOCKS
#define SST_SA_MIN_BLOCKS 5
#endif

inline unsigned grid_for(size_t work) { return (unsigned)std::min<size_t>(div_ceil(work, (size_t)kThreads), (size_t)cur_sms() * 32); }

// ---- construction --------------------------------------------------------------------------
__global__ void sa_init_keys(const uint8_t* __restrict__ t, size_t n, unsigned long long* __restrict__ keys,
                             uint32_t* __restrict__ vals) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long k = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const size_t p = i + j;
            const unsigned long long sym = p < n ? (unsigned long long)t[p] + 1ull : 0ull;
            k = (k << 9) | sym;
        }
        keys[i] = k;
        vals[i] = (uint32_t)i;
    }
}

// flags[j] = 1 when sorted key j starts a new group.
__global__ void sa_flag_heads(const unsigned long long* __restrict__ keys, size_t n, uint32_t* __restrict__ flags) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
        flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

// rank[sa[j]] = (number of group heads up to j) - 1
__global__ void sa_scatter_rank(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ scan, size_t n,
                                uint32_t* __restrict__ rank) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
        rank[sa[j]] = scan[j] - 1u;
}

// key for suffix i after knowing ranks by the first h symbols: (rank[i], rank[i+h] + 1 | 0)
__global__ void sa_pair_keys(const uint32_t* __restrict__ rank, size_t n, size_t h, unsigned long long* __restrict__ keys,
                             uint32_t* __restrict__ vals) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long hi = rank[i];
        const unsigned long long lo = i + h < n ? (unsigned long long)rank[i + h] + 1ull : 0ull;
        keys[i] = (hi << 32) | lo;
        vals[i] = (uint32_t)i;
    }
}

int bits_needed(unsigned long long max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b)) b++;
    return b;
}

bool build_sa_device(const uint8_t* d_text, size_t n, uint32_t* d_sa, int device) {
    cudaStream_t st = thread_stream(device);
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    uint32_t *v0 = nullptr, *v1 = nullptr, *rank = nullptr, *flags = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0, scan_bytes = 0;
    bool ok = SST_CUDA_OK(cudaMalloc(&k0, n * 8)) && SST_CUDA_OK(cudaMalloc(&k1, n * 8)) && SST_CUDA_OK(cudaMalloc(&v0, n * 4)) &&
              SST_CUDA_OK(cudaMalloc(&v1, n * 4)) && SST_CUDA_OK(cudaMalloc(&rank, n * 4)) && SST_CUDA_OK(cudaMalloc(&flags, n * 4));
    if (ok) {
        cub::DoubleBuffer<unsigned long long> dk(k0, k1);
        cub::DoubleBuffer<uint32_t> dv(v0, v1);
        ok = SST_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (unsigned long long)n, 0, 64, st)) &&
             SST_CUDA_OK(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, flags, flags, (unsigned long long)n, st));
        tmp_bytes = std::max(tmp_bytes, scan_bytes);
        ok = ok && SST_CUDA_OK(cudaMalloc(&tmp, tmp_bytes));
    }
    if (ok) {
        sa_init_keys<<<grid_for(n), kThreads, 0, st>>>(d_text, n, k0, v0);
        int end_bit = 63;
        size_t h = 7;
        for (int round = 0; ok; round++) {
            cub::DoubleBuffer<unsigned long long> dk(k0, k1);
            cub::DoubleBuffer<uint32_t> dv(v0, v1);
            size_t tb = tmp_bytes;
            ok = SST_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (unsigned long long)n, 0, end_bit, st));
            if (!ok) break;
            const unsigned long long* sk = dk.Current();
            const uint32_t* sv = dv.Current();
            sa_flag_heads<<<grid_for(n), kThreads, 0, st>>>(sk, n, flags);
            tb = tmp_bytes;
            ok = SST_CUDA_OK(cub::DeviceScan::InclusiveSum(tmp, tb, flags, flags, (unsigned long long)n, st));
            if (!ok) break;
            uint32_t groups = 0;
            ok = SST_CUDA_OK(cudaMemcpyAsync(&groups, flags + (n - 1), 4, cudaMemcpyDeviceToHost, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
            if (!ok) break;
            if (groups == n || h >= n) {
                ok = SST_CUDA_OK(cudaMemcpyAsync(d_sa, sv, n * 4, cudaMemcpyDeviceToDevice, st));
                break;
            }
            sa_scatter_rank<<<grid_for(n), kThreads, 0, st>>>(sv, flags, n, rank);
            sa_pair_keys<<<grid_for(n), kThreads, 0, st>>>(rank, n, h, k0, v0);
            end_bit = 32 + bits_needed(groups);
            h *= 2;
        }
        ok = ok && SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(rank); cudaFree(flags); cudaFree(tmp);
    return ok;
}

// ---- strict-order check: sa_search.rs:36-38 --------------------------------------------------
__global__ void sa_check_kernel(const uint8_t* __restrict__ t, size_t n, const uint32_t* __restrict__ sa,
                                unsigned long long* __restrict__ bad) {
    unsigned long long local = 0;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x + 1; j < n; j += (size_t)gridDim.x * blockDim.x) {
        const size_t x = sa[j - 1], y = sa[j];
        if (x >= n || y >= n) { local++; continue; }
        const size_t lx = n - x, ly = n - y, m = lx < ly ? lx : ly;
        size_t i = 0;
        while (i < m && t[x + i] == t[y + i]) i++;
        const bool less = i < m ? t[x + i] < t[y + i] : lx < ly;
        if (!less) local++;
    }
    if (local) atomicAdd(bad, local);
}

// ---- search ----------------------------------------------------------------------------------
// 4 bytes starting at byte address base+pos (little endian), built from aligned words; words at
// or beyond `end` read as zero, so nothing outside [base_aligned, end) is touched.
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* __restrict__ base, unsigned long long pos,
                                                       unsigned long long end) {
    // the ADDRESS is aligned down (a chunk of the host-buffer path starts at an arbitrary byte of the packed patterns)
    const uint8_t* addr = base + pos;
    const uint8_t* lim = base + end;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(addr) & ~(uintptr_t)3);
    const uint32_t w0 = reinterpret_cast<const uint8_t*>(w) < lim ? __ldg(w) : 0u;
    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(addr) & 3u);
    if (sh == 0) return w0;
    const uint32_t w1 = reinterpret_cast<const uint8_t*>(w + 1) < lim ? __ldg(w + 1) : 0u;
    return __funnelshift_r(w0, w1, sh * 8);
}

struct SaParams {
    const uint8_t* text;
    const uint32_t* sa;
    unsigned long long n;
    const uint8_t* pats;
    const unsigned long long* pat_off;
    unsigned long long pats_bytes;  // pat_off[npat]
    unsigned long long npat;
    uint32_t* out_lo;
    uint32_t* out_hi;
    uint32_t* out_pos;
    const uint4* pivots;
    int pivot_levels;
    // reordered batch: the coarse pass writes the lower bound after `coarse_levels` table levels (sort key) and the
    // identity (sort value); the main pass visits the patterns in the sorted order `perm`
    const uint32_t* perm;
    uint32_t* keys;
    uint32_t* ident;
    int coarse_levels;
    const uint32_t* kmer;  // k-mer table (or null)
    int kmer_k;
    const uint2* sax;      // {sa, next 15 bases} entries (or null)
    const uint4* saw;      // {sa, next 48 bases in three words} entries (or null; the WIDE kernels)
    const uint4* cells;    // packed k-mer cells, 4 x uint4 each (or null)
    const uint32_t* text2; // the text at 2 bits per base (or null)
    uint32_t* out_probes;  // sa_search_kernel only: iterations of the reference's loop (its `cnt`, sa_search.rs:98-112), or null
};

// Compares suffix(spos) with the pattern from byte `start` on.  Returns lcp (group-uniform) and
// sets less = suffix < pattern under Rust slice ordering (proper prefix is smaller).
template <int PL>
__device__ __forceinline__ uint32_t group_compare(const SaParams& p, unsigned long long spos, const uint8_t* pat_base,
                                                  unsigned long long pat_pos, uint32_t ql, uint32_t start, unsigned sub,
                                                  unsigned gbase, unsigned gmask, bool& less) {
    const unsigned long long sl64 = p.n - spos;
    const uint32_t sl = sl64 > 0xffffffffull ? 0xffffffffu : (uint32_t)sl64;
    const uint32_t lim = sl < ql ? sl : ql;
    uint32_t off = start;
    while (true) {
        if (off >= lim) { less = sl < ql; return lim; }
        const uint32_t my = off + 4u * sub;
        uint32_t diff = 0, tw = 0, pw = 0;
        if (my < lim) {
            tw = load_u32_unaligned(p.text, spos + my, p.n);
            pw = load_u32_unaligned(pat_base, pat_pos + my, p.pats_bytes);
            diff = tw ^ pw;
            const uint32_t v = lim - my;
            if (v < 4u) diff &= (1u << (8u * v)) - 1u;
        }
        const unsigned b = (__ballot_sync(gmask, diff != 0u) >> gbase) & (PL == 32 ? 0xffffffffu : ((1u << PL) - 1u));
        if (b) {
            const unsigned f = __ffs(b) - 1;
            const unsigned bytepos = (__ffs(diff) - 1) >> 3;  // meaningful on lane f only
            const uint32_t my_lcp = my + bytepos;
            const int my_less = ((tw >> (8u * bytepos)) & 0xffu) < ((pw >> (8u * bytepos)) & 0xffu);
            const uint32_t lcp = __shfl_sync(gmask, my_lcp, gbase + f);
            less = __shfl_sync(gmask, my_less, gbase + f) != 0;
            return lcp;
        }
        off += 4u * PL;
    }
}

template <int PL, bool MLR>
__global__ void __launch_bounds__(kThreads)
sa_search_kernel(const __grid_constant__ SaParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned sub = lane & (PL - 1), gbase = lane & ~(unsigned)(PL - 1);
    const unsigned gmask = PL == 32 ? 0xffffffffu : (((1u << PL) - 1u) << gbase);
    const unsigned long long groups_per_block = kThreads / PL;
    for (unsigned long long i = (unsigned long long)blockIdx.x * groups_per_block + threadIdx.x / PL; i < p.npat;
         i += (unsigned long long)gridDim.x * groups_per_block) {
        const unsigned long long po = p.pat_off[i];
        const uint32_t ql = (uint32_t)(p.pat_off[i + 1] - po);
        // ---- lower bound: sa_search.rs:98-112 ----
        unsigned long long l = 0, r = p.n;
        uint32_t lcp_l = 0, lcp_r = 0, probes = 0;
        while (l < r) {
            probes++;  // the reference's `*cnt += 1` (sa_search.rs:104)
            const unsigned long long m = (l + r) >> 1;
            const unsigned long long spos = __ldg(p.sa + m);
            bool less;
            const uint32_t start = MLR ? (lcp_l < lcp_r ? lcp_l : lcp_r) : 0u;
            const uint32_t lcp = group_compare<PL>(p, spos, p.pats, po, ql, start, sub, gbase, gmask, less);
            if (less) { l = m + 1; lcp_l = lcp; } else { r = m; lcp_r = lcp; }
        }
        const unsigned long long lo = l;
        // ---- hi: first index >= lo whose suffix does not start with the pattern (gallop + bisect) ----
        unsigned long long hi = lo;
        if (p.out_hi) {
            // Suffixes starting with q are contiguous from lo: gallop to bracket the end, then bisect.
            unsigned long long a = lo, b = p.n, step = 1;  // invariant: all of [lo, a) start with q
            while (true) {
                const unsigned long long pr = a + step - 1;
                if (pr >= b) break;
                bool less;
                const uint32_t lcp = group_compare<PL>(p, __ldg(p.sa + pr), p.pats, po, ql, 0u, sub, gbase, gmask, less);
                if (lcp >= ql) { a = pr + 1; step <<= 1; } else { b = pr; break; }
            }
            // bisect in [a, b): everything before a matches, b does not (or b == n)
            while (a < b) {
                const unsigned long long m = (a + b) >> 1;
                bool less;
                const uint32_t lcp = group_compare<PL>(p, __ldg(p.sa + m), p.pats, po, ql, 0u, sub, gbase, gmask, less);
                if (lcp >= ql) a = m + 1; else b = m;
            }
            hi = a;
        }
        if (sub == 0) {
            if (p.out_probes) p.out_probes[i] = probes;
            p.out_lo[i] = (uint32_t)lo;
            if (p.out_hi) p.out_hi[i] = (uint32_t)hi;
            if (p.out_pos) p.out_pos[i] = lo < p.n ? __ldg(p.sa + lo) : 0xffffffffu;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Thread-per-pattern search with a pivot-prefix table (default).
//
// The binary search of sa_search.rs:98-112 visits a fixed implicit tree of midpoints.  For its top
// `pivot_levels` levels the first 16 bytes of every pivot suffix are stored in 128-byte blocks of
// three levels each (pivot_entry), so a probe there is ONE 16-byte load, three probes share one
// L2/DRAM fill, and the two dependent misses sa[m] -> text[sa[m]..] are only paid on a 16-byte tie.
// By default the table covers the whole search when it fits the memory budget (16 B per suffix,
// 2.4 GB for a 10^8 text).  Below the table every probe loads sa[m] and compares 16-byte windows
// (two aligned LDG.128 + funnel shifts).  One thread per pattern maximises the number of independent miss chains per SM, which
// is what bounds this path (measured: 8 lanes/pattern 0.57, 4: 0.82, 2: 1.04 Gpat/s).
// ------------------------------------------------------------------------------------------------
// Blocked pivot table: three consecutive levels of the implicit search tree (1 + 2 + 4 pivots, 16 B each)
// share one 128-byte block, so one L2/DRAM fill serves three probes.  Node j (heap index, root = 1) at
// depth d lives in triple t = d / 3; its block is rooted at j >> (d % 3); blocks of triple t start at
// entry 8 * (8^t - 1) / 7.  Entry 0 of every block is unused.
__host__ __device__ __forceinline__ unsigned long long pivot_entry(unsigned long long j, int d) {
    const int t = d / 3, e = d - 3 * t;
    const unsigned long long root = j >> e;                          // heap index of the block's root
    const unsigned long long block = root - (1ull << (3 * t));       // ordinal within the triple
    const unsigned long long base = ((1ull << (3 * t)) - 1ull) / 7ull;  // blocks in earlier triples
    const unsigned long long within = (1ull << e) | (j & ((1ull << e) - 1ull));
    return (base + block) * 8ull + within;
}
__host__ __device__ __forceinline__ unsigned long long pivot_table_entries(int levels) {  // levels is a multiple of 3
    return (((1ull << levels) - 1ull) / 7ull) * 8ull;
}

// Random loads of the search (k-mer cells, {sa, bases} entries, sa[m], text windows) carry the L2::64B prefetch-size
// qualifier: by default B200 fills L2 from HBM in 128-byte units, so every isolated 16-32-byte read dragged 128 bytes
// in (ncu, round 1: 299 B of DRAM reads per C3 pattern for ~2.3 fills).  -DSST_SA_PF64=0 builds the plain __ldg form (A/B).
#ifndef SST_SA_PF64
#define SST_SA_PF64 1
#endif
__device__ __forceinline__ uint32_t ldr(const uint32_t* p) {
#if SST_SA_PF64
    uint32_t v;
    asm("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint2 ldr(const uint2* p) {
#if SST_SA_PF64
    uint2 v;
    asm("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint4 ldr(const uint4* p) {
#if SST_SA_PF64
    uint4 v;
    asm("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// (the packed k-mer cells are 128 bytes: ask L2 for the whole line, the second half is read only by crowded cells;
// 32 bytes of a cell in one request: sm_100 has 256-bit global loads)
__device__ __forceinline__ void ldc256(const uint4* p, uint4& a, uint4& b) {
    asm("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

struct W4 { uint32_t w[4]; };

// 16 bytes starting at byte address `addr` (little endian words); aligned 16-byte chunks starting
// at or beyond `end` read as zero.
template <bool GUARD, bool RANDOM = false>
__device__ __forceinline__ W4 load16_unaligned(const uint8_t* addr, const uint8_t* end) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(addr) & ~(uintptr_t)15;
    const uint4* c = reinterpret_cast<const uint4*>(a);
    uint4 A = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0, 0);
    if (!GUARD || reinterpret_cast<const uint8_t*>(c) < end) A = RANDOM ? ldr(c) : __ldg(c);
    if (!GUARD || reinterpret_cast<const uint8_t*>(c + 1) < end) B = RANDOM ? ldr(c + 1) : __ldg(c + 1);
    const unsigned s = (unsigned)(reinterpret_cast<uintptr_t>(addr) & 15u);
    const bool s2 = (s & 8u) != 0, s1 = (s & 4u) != 0;
    const unsigned bs = (s & 3u) * 8u;
    // word rotate by s/4 without dynamic register indexing
    const uint32_t u0 = s2 ? A.z : A.x, u1 = s2 ? A.w : A.y, u2 = s2 ? B.x : A.z, u3 = s2 ? B.y : A.w, u4 = s2 ? B.z : B.x, u5 = s2 ? B.w : B.y;
    const uint32_t v0 = s1 ? u1 : u0, v1 = s1 ? u2 : u1, v2 = s1 ? u3 : u2, v3 = s1 ? u4 : u3, v4 = s1 ? u5 : u4;
    W4 r;
    r.w[0] = __funnelshift_r(v0, v1, bs);
    r.w[1] = __funnelshift_r(v1, v2, bs);
    r.w[2] = __funnelshift_r(v2, v3, bs);
    r.w[3] = __funnelshift_r(v3, v4, bs);
    return r;
}

// First mismatching byte (0..15, or 16 if none) between two 16-byte windows; tb/pb = the bytes there.
// Branch-free: a select chain picks the first differing word.
__device__ __forceinline__ unsigned first_mismatch16(const W4& t, const W4& q, unsigned& tb, unsigned& pb) {
    const uint32_t x0 = t.w[0] ^ q.w[0], x1 = t.w[1] ^ q.w[1], x2 = t.w[2] ^ q.w[2], x3 = t.w[3] ^ q.w[3];
    const bool n0 = x0 != 0, n1 = x1 != 0, n2 = x2 != 0;
    const uint32_t x = n0 ? x0 : n1 ? x1 : n2 ? x2 : x3;
    const uint32_t tw = n0 ? t.w[0] : n1 ? t.w[1] : n2 ? t.w[2] : t.w[3];
    const uint32_t pw = n0 ? q.w[0] : n1 ? q.w[1] : n2 ? q.w[2] : q.w[3];
    const unsigned wi = n0 ? 0u : n1 ? 1u : n2 ? 2u : 3u;
    const unsigned sh = (__ffs(x) - 1) & 24u;  // bit offset of the first differing byte (x == 0 -> 24, masked below)
    tb = (tw >> sh) & 0xffu;
    pb = (pw >> sh) & 0xffu;
    return x ? 4u * wi + (sh >> 3) : 16u;
}

// Four bases (one per byte, values 0..3, first byte first) -> 8 bits, first base most significant: the four 2-bit fields
// land in bits 24..31 of the product without overlapping any cross term.
__device__ __forceinline__ uint32_t pack4(uint32_t w) { return ((w & 0x03030303u) * 0x40100401u) >> 24; }
// 16 bases of a window -> 32 bits, first base most significant (string_value order, util.rs:76-117)
__device__ __forceinline__ uint32_t pack16(const W4& w) { return (pack4(w.w[0]) << 24) | (pack4(w.w[1]) << 16) | (pack4(w.w[2]) << 8) | pack4(w.w[3]); }

// suffix(spos) vs pattern from byte `start` (a multiple of 16) on; returns lcp, sets less.
__device__ __forceinline__ uint32_t thread_compare(const SaParams& p, unsigned long long spos, const W4& p0, const W4& p1,
                                                   const uint8_t* pat, uint32_t ql, uint32_t start, bool& less) {
    const unsigned long long sl64 = p.n - spos;
    const uint32_t sl = sl64 > 0xffffffffull ? 0xffffffffu : (uint32_t)sl64;
    const uint32_t lim = sl < ql ? sl : ql;
    const uint8_t* tbase = p.text + spos;
    const uint8_t* tend = p.text + p.n + 64;   // the text is followed by 64 zero bytes
    const uint8_t* pend = p.pats + p.pats_bytes;
    if (sl < start) start = 0u;  // (a suffix shorter than the prefix the caller assumes equal: see the k-mer cell's LCP seed)
    for (uint32_t off = start;; off += 16u) {
        if (off >= lim) { less = sl < ql; return lim; }
        const W4 tw = load16_unaligned<false, true>(tbase + off, tend);  // the text has 64 bytes of zero padding
        W4 pw;
        if (off == 0u) pw = p0;
        else if (off == 16u) pw = p1;
        else pw = load16_unaligned<true>(pat + off, pend);
        unsigned tb, pb;
        const unsigned mp = first_mismatch16(tw, pw, tb, pb);
        const uint32_t valid = lim - off;  // >= 1
        if (mp < 16u && mp < valid) { less = tb < pb; return off + mp; }
    }
}

// Mask of the bases a pattern has among bases 16 j .. 16 j + 15 after the first k, when it has nb of them in all (2 bits per
// base, first base most significant).
__device__ __forceinline__ uint32_t mask16(uint32_t nb, int j) {
    const uint32_t from = 16u * (uint32_t)j;
    return nb >= from + 16u ? 0xffffffffu : nb <= from ? 0u : 0xffffffffu << (2u * (16u - (nb - from)));
}

// suffix(spos) vs the pattern from base `from` (a multiple of 16) on, through the 2-bit packed text; the caller knows that the
// bases before `from` are equal.  Same contract as thread_compare (returns lcp, sets less); a pattern byte outside the alphabet
// or a suffix shorter than `from` hands over to it.
__device__ __forceinline__ uint32_t packed_compare(const SaParams& p, uint32_t spos, const W4& p0, const W4& p1, const uint8_t* pat, uint32_t ql,
                                                   uint32_t from, bool& less) {
    const unsigned long long sl64 = p.n - spos;
    const uint32_t sl = sl64 > 0xffffffffull ? 0xffffffffu : (uint32_t)sl64;
    if (sl < from) return thread_compare(p, spos, p0, p1, pat, ql, from, less);
    const uint32_t lim = sl < ql ? sl : ql;
    const uint8_t* pend = p.pats + p.pats_bytes;
    const unsigned long long b0 = (unsigned long long)spos + from;  // first base to compare
    const uint32_t* tw = p.text2 + (b0 >> 4);
    const unsigned sh = (unsigned)(b0 & 15ull) * 2u;
    uint32_t hi = ldr(tw);
    for (uint32_t off = from;; off += 16u) {
        if (off >= lim) { less = sl < ql; return lim; }
        const uint32_t lo = ldr(++tw);
        const uint32_t t16 = __funnelshift_l(lo, hi, sh);  // text bases spos + off .. + 15
        hi = lo;
        const W4 pw = load16_unaligned<true>(pat + off, pend);
        const uint32_t valid = lim - off;  // >= 1
        uint32_t bad = 0;
#pragma unroll
        for (int wi = 0; wi < 4; wi++) {
            const uint32_t have = valid >= 4u * wi + 4u ? 0xffffffffu : valid <= 4u * wi ? 0u : (1u << (8u * (valid - 4u * wi))) - 1u;
            bad |= pw.w[wi] & 0xfcfcfcfcu & have;
        }
        if (bad) return thread_compare(p, spos, p0, p1, pat, ql, off, less);
        const uint32_t mask = valid >= 16u ? 0xffffffffu : 0xffffffffu << (2u * (16u - valid));
        const uint32_t a = t16 & mask, b = pack16(pw) & mask;
        if (a != b) { less = a < b; return off + ((uint32_t)__clz((int)(a ^ b)) >> 1); }
    }
}

// The same for a pattern whose bases 48 .. 111 are already packed (cw[c] = bases 48 + 16 c .. 63 + 16 c, checked to lie in the
// alphabet): the 48-base entries leave exactly these chunks to compare for patterns of up to 112 bases, and a pattern is probed more
// than once (lower bound, then hi), so packing them once per pattern instead of once per probe and chunk removes ~90 of the ~105
// instructions of a chunk.  Bases 0 .. 47 are known equal; a longer pattern goes on in packed_compare.
__device__ __forceinline__ uint32_t packed_compare48(const SaParams& p, uint32_t spos, const W4& p0, const W4& p1, const uint8_t* pat, uint32_t ql,
                                                     uint32_t from, const uint32_t (&cw)[4], bool& less) {
    const unsigned long long sl64 = p.n - spos;
    const uint32_t sl = sl64 > 0xffffffffull ? 0xffffffffu : (uint32_t)sl64;
    if (sl < from) return thread_compare(p, spos, p0, p1, pat, ql, from, less);
    if (from >= 112u) return packed_compare(p, spos, p0, p1, pat, ql, from, less);  // (mlr on long patterns: the packed chunks are known equal)
    const uint32_t lim = sl < ql ? sl : ql;
    const unsigned long long b0 = (unsigned long long)spos + 48ull;
    const uint32_t* tw = p.text2 + (b0 >> 4);
    const unsigned sh = (unsigned)(b0 & 15ull) * 2u;
    const uint32_t c0 = from >= 64u ? 1u : 0u;  // (k = 16: the entry covers bases 16 .. 63, the first chunk is known equal too)
    uint32_t hi = ldr(tw + c0);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t off = 48u + 16u * (uint32_t)c;
        if ((uint32_t)c < c0) continue;
        if (off >= lim) { less = sl < ql; return lim; }
        const uint32_t lo = ldr(tw + c + 1);
        const uint32_t t16 = __funnelshift_l(lo, hi, sh);  // text bases spos + off .. + 15
        hi = lo;
        const uint32_t valid = lim - off;  // >= 1
        const uint32_t mask = valid >= 16u ? 0xffffffffu : 0xffffffffu << (2u * (16u - valid));
        const uint32_t a = t16 & mask, b = cw[c] & mask;
        if (a != b) { less = a < b; return off + ((uint32_t)__clz((int)(a ^ b)) >> 1); }
    }
    if (112u >= lim) { less = sl < ql; return lim; }
    return packed_compare(p, spos, p0, p1, pat, ql, from > 112u ? from : 112u, less);
}

// PHASE 0: patterns in the caller's order.  PHASE 1 (coarse): only the first coarse_levels table levels; writes the lower
// bound reached (a monotone function of the pattern: the sort key) and the pattern's index.  PHASE 2: patterns in the
// order of `perm`, i.e. sorted by that key: the lanes of a warp then walk (almost) the same path, so their table and
// suffix-array loads fall into the same lines instead of 32 different ones.
// MINB: resident blocks of 256 threads' worth per SM the register budget is set for: 4 (62 registers, no spills) for the 48-base
// entries, 5 (48 registers) for the other paths.  One pattern per thread: the launch makes one block per kSearchThreads patterns.
// WIDE: the inlined entries the index holds: 0 = {sa, 15 bases} (8 bytes), 1 = {sa, 48 bases} (16 bytes).
template <bool MLR, int PHASE, int WIDE = 0, int MINB = SST_SA_MIN_BLOCKS>
__global__ void __launch_bounds__(kSearchThreads, MINB * (256 / kSearchThreads))
sa_search_thread_kernel(const __grid_constant__ SaParams p) {
    for (unsigned long long slot = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; slot < p.npat;
         slot += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = PHASE == 2 ? (unsigned long long)p.perm[slot] : slot;
        const unsigned long long po = p.pat_off[i];
        const uint32_t ql = (uint32_t)(p.pat_off[i + 1] - po);
        const uint8_t* pat = p.pats + po;
        const uint8_t* pend = p.pats + p.pats_bytes;
        const W4 p0 = load16_unaligned<true>(pat, pend), p1 = load16_unaligned<true>(pat + 16, pend);
        // ---- k-mer table: the first k bases select the suffix-array range directly (one load instead of ~2k probes) ----
        bool have_range = false;
        uint32_t range_end = (uint32_t)p.n;  // no suffix from here on starts with q (bounds the search for hi)
        uint32_t kx = 0;                     // 2-bit code of the pattern's first k bases (padded with the smallest base)
        if (PHASE != 1 && p.kmer_k) {
            const uint32_t k = (uint32_t)p.kmer_k;  // <= 16: the bases sit in p0
            const uint32_t nk = ql < k ? ql : k;     // bases of the pattern among the first k
            uint32_t bad_k = 0;
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                const uint32_t have = nk >= 4u * wi + 4u ? 0xffffffffu : nk <= 4u * wi ? 0u : (1u << (8u * (nk - 4u * wi))) - 1u;
                bad_k |= p0.w[wi] & 0xfcfcfcfcu & have;
            }
            // the first 16 bases at 2 bits each, first base most significant; the ones the pattern lacks read as the smallest base
            // (the loop over the k bytes this replaces cost ~100 of the kernel's ~1800 instructions per pattern)
            const uint32_t c16 = pack16(p0) & (nk >= 16u ? 0xffffffffu : nk ? 0xffffffffu << (2u * (16u - nk)) : 0u);
            have_range = bad_k == 0u;
            kx = c16 >> (32u - 2u * k);
        }
        // ---- inlined bases: inside the k-mer cell a probe reads {sa[m], the 15 bases after the first k} and goes to the text
        // only when those 15 bases equal the pattern's (the suffix that matches, if any) ----
        bool inl = false;
        using Code = typename std::conditional<WIDE != 0, unsigned long long, uint32_t>::type;
        constexpr uint32_t NB = WIDE == 1 ? 48u : 15u;  // bases inlined per suffix
        Code pq = 0, pmask = 0;      // the pattern's bases k .. k+14 (k+31 for the packed cells), those it has, and the mask of the ones it has
        uint32_t pq3[WIDE == 1 ? 3 : 1] = {};  // WIDE: the pattern's bases k .. k+47 in three words of 16 (masked to the nbp it has)
        uint32_t cwk[4] = {};                  // WIDE: the pattern's bases 48 .. 111 in four words of 16 (for packed_compare48)
        uint32_t nbp = 0;
        bool pat_ends = false;       // the pattern ends within those bases: equal bases = the suffix starts with the pattern
        if (have_range && (WIDE == 1 ? (const void*)p.saw : (const void*)p.sax) && ql >= (uint32_t)p.kmer_k) {
            uint32_t bad = 0;  // a byte outside the alphabet among the first min(ql, 32 or 112) bytes: no inline compare for this pattern
#pragma unroll
            for (int wi = 0; wi < 8; wi++) {
                const uint32_t w = wi < 4 ? p0.w[wi] : p1.w[wi - 4];
                const uint32_t have = ql >= 4u * wi + 4u ? 0xffffffffu : ql <= 4u * wi ? 0u : (1u << (8u * (ql - 4u * wi))) - 1u;
                bad |= w & 0xfcfcfcfcu & have;
            }
            const unsigned long long code = ((unsigned long long)pack16(p0) << 32) | pack16(p1);  // bases 0..31
            const uint32_t nb = ql - (uint32_t)p.kmer_k < NB ? ql - (uint32_t)p.kmer_k : NB;  // bases of the pattern after the first k
            pat_ends = nb < NB;
            if constexpr (WIDE == 1) {
                uint32_t cw[7] = {(uint32_t)(code >> 32), (uint32_t)code, 0u, 0u, 0u, 0u, 0u};  // bases 0..15, 16..31, .. 96..111
#pragma unroll
                for (int j = 2; j < 7; j++)
                    if (ql > 16u * j) {
                        const W4 pj = load16_unaligned<true>(pat + 16 * j, pend);
#pragma unroll
                        for (int wi = 0; wi < 4; wi++) {
                            const uint32_t at = 16u * j + 4u * wi;
                            const uint32_t have = ql >= at + 4u ? 0xffffffffu : ql <= at ? 0u : (1u << (8u * (ql - at))) - 1u;
                            bad |= pj.w[wi] & 0xfcfcfcfcu & have;
                        }
                        cw[j] = pack16(pj);
                    }
                const unsigned sh = 2u * (unsigned)p.kmer_k;  // 2 .. 32: the k bases the cell fixes are shifted out
                nbp = nb;
#pragma unroll
                for (int j = 0; j < 3; j++) pq3[j] = (sh >= 32u ? cw[j + 1] : ((cw[j] << sh) | (cw[j + 1] >> (32u - sh)))) & mask16(nb, j);
#pragma unroll
                for (int j = 0; j < 4; j++) cwk[j] = cw[3 + j];
                // (the packed cells hold 32 bases per entry: they answer a pattern that ends within those)
                pmask = nb >= 32u ? ~0ull : nb ? ~0ull << (2u * (32u - nb)) : 0ull;
                pq = ((unsigned long long)pq3[0] << 32) | pq3[1];
            } else {
                pmask = nb ? (0x3fffffffu >> (2u * (15u - nb))) << (2u * (15u - nb)) : 0u;
                pq = (uint32_t)(code >> (2 * (32 - p.kmer_k - 15))) & pmask;
            }
            inl = bad == 0u;
        }
        // ---- packed cell: range and entries of the pattern's k-mer in one 64-byte line; a pattern that ends within the inlined
        // bases is answered from it alone when the cell is not flagged (one random DRAM access per pattern instead of two) ----
        if constexpr (WIDE == 1 && PHASE == 0) {
            if (p.cells && inl && nbp < 32u) {
                const uint4* c = p.cells + (size_t)kx * 8;
                uint4 c0, c1, c2, c3;  // first half: start + entries 0..4
                ldc256(c, c0, c1); ldc256(c + 2, c2, c3);
                if (!(c0.x & 0x80000000u)) {
                    const uint32_t start = c0.x;
                    unsigned below = 0, equal = 0, cnt = 0;
                    uint32_t sp = 0xffffffffu;  // sa of entry number `below` (the lower bound), once seen
                    // five entries {sa, code hi, code lo} in words 1..15 of a half; an entry with sa = 0xffffffff is unused
                    auto half = [&](unsigned base) {
                        const uint32_t es[5] = {c0.y, c1.x, c1.w, c2.z, c3.y};
                        const unsigned long long ec[5] = {((unsigned long long)c0.z << 32) | c0.w, ((unsigned long long)c1.y << 32) | c1.z,
                                                          ((unsigned long long)c2.x << 32) | c2.y, ((unsigned long long)c2.w << 32) | c3.x,
                                                          ((unsigned long long)c3.z << 32) | c3.w};
                        const unsigned below0 = below;
#pragma unroll
                        for (int e = 0; e < 5; e++) {
                            const bool valid = es[e] != 0xffffffffu;
                            const unsigned long long cm = ec[e] & pmask;
                            cnt += valid ? 1u : 0u;
                            below += valid && cm < pq ? 1u : 0u;   // shares the first k bases, the next ones are smaller: suffix < q
                            equal += valid && cm == pq ? 1u : 0u;  // every base of the pattern matches: the suffix starts with q
                        }
#pragma unroll
                        for (int e = 0; e < 5; e++)  // (entries are sorted: `below` only grows while it equals the entries seen)
                            if (below == base + (unsigned)e && below0 <= base + (unsigned)e) sp = es[e];
                    };
                    half(0u);
                    if (cnt == 5u && below + equal == 5u) {  // all five are < q or start with q: the cell may hold more of either kind
                        ldc256(c + 4, c0, c1); ldc256(c + 6, c2, c3);  // entries 5..9 (word 0 of this half is unused)
                        half(5u);
                    }
                    const uint32_t lo = start + below;
                    p.out_lo[i] = lo;
                    if (p.out_hi) p.out_hi[i] = lo + equal;
                    if (p.out_pos) {
                        if (below >= cnt) sp = lo < p.n ? ldr(p.sa + lo) : 0xffffffffu;  // first suffix of a later cell
                        p.out_pos[i] = sp;
                    }
                    continue;
                }
            }
        }
        uint32_t l = 0, r = (uint32_t)p.n;  // n < 2^32 - 16: all search state fits 32 bits
        uint32_t lcp_l = 0, lcp_r = 0;
        bool lcp_r_exact = false;  // lcp_r == lcp(q, suffix(r)) exactly (not a conservative bound)
        if (have_range) {
            const int k = p.kmer_k;
            if (ql >= (uint32_t)k) {
                l = ldr(p.kmer + kx);
                r = ldr(p.kmer + (size_t)kx + 1);
                range_end = r;  // every suffix that starts with q starts with its first k bases
                // every suffix of the cell that has k bases shares them with the pattern: both LCP bounds start at k instead of
                // 0 (lower bounds, not exact values: lcp_r_exact stays false).  One of the text's last < k suffixes can sit in
                // the cell without sharing them (AAC lies between AAAT and AACA); thread_compare starts such a suffix,
                // which is shorter than the assumed common prefix, at byte 0.
                if (MLR) lcp_l = lcp_r = (uint32_t)k;
            } else {
                // every suffix from kmer[x] on is >= q000.. >= q; the only suffixes below it that are >= q are proper
                // prefixes of q000.. (the last < k suffixes of the text): search the k positions before it
                r = ldr(p.kmer + kx);
                l = r > (uint32_t)k ? r - (uint32_t)k : 0u;
            }
        }
        // suffix(sa[m]) vs the pattern: lcp and order, through the inlined bases where they decide
        auto probe = [&](uint32_t m, uint32_t start, bool& less) -> uint32_t {
            if (inl) {
                if constexpr (WIDE == 1) {
                    const uint4 e = ldr(p.saw + m);  // {sa, bases k .. k+15, k+16 .. k+31, k+32 .. k+47}
                    if ((unsigned long long)e.x + (unsigned)p.kmer_k + 48ull <= p.n) {  // the suffix has all 48 bases
                        const uint32_t ew[3] = {e.y, e.z, e.w};
                        uint32_t x = 0, ev = 0, pv = 0, at = 0;  // first word (16 bases) that differs among the pattern's bases
#pragma unroll
                        for (int j = 2; j >= 0; j--) {
                            const uint32_t a = ew[j] & mask16(nbp, j), d = a ^ pq3[j];
                            if (d) { x = d; ev = a; pv = pq3[j]; at = 16u * (uint32_t)j; }
                        }
                        if (x) {
                            less = ev < pv;
                            return (uint32_t)p.kmer_k + at + ((uint32_t)__clz((int)x) >> 1);
                        }
                        if (pat_ends) { less = false; return ql; }  // every base of the pattern matched: no text access at all
                        const uint32_t known = ((uint32_t)p.kmer_k + NB) & ~15u;  // bytes known to be equal, rounded down to a window
                        if (p.text2) return packed_compare48(p, e.x, p0, p1, pat, ql, start > known ? start : known, cwk, less);  // (known = 48)
                        return thread_compare(p, e.x, p0, p1, pat, ql, start > known ? start : known, less);
                    }
                    return thread_compare(p, e.x, p0, p1, pat, ql, start, less);
                } else {
                    const uint2 e = ldr(p.sax + m);
                    const uint32_t spos = e.x, code = e.y & pmask;
                    if ((e.y >> 31) != 0u) {  // the suffix has all the inlined bases
                        if (code != pq) {
                            less = code < pq;
                            return (uint32_t)p.kmer_k + (((uint32_t)__clz((int)(code ^ pq)) - 2u) >> 1);
                        }
                        if (pat_ends) { less = false; return ql; }  // every base of the pattern matched: no text access at all
                        const uint32_t known = ((uint32_t)p.kmer_k + NB) & ~15u;  // bytes known to be equal, rounded down to a window
                        if (p.text2) return packed_compare(p, spos, p0, p1, pat, ql, start > known ? start : known, less);
                        return thread_compare(p, spos, p0, p1, pat, ql, start > known ? start : known, less);
                    }
                    return thread_compare(p, spos, p0, p1, pat, ql, start, less);
                }
            }
            return thread_compare(p, ldr(p.sa + m), p0, p1, pat, ql, start, less);
        };
        // ---- table levels: one 16-byte load per probe ----
        if (!have_range) {
            const uint32_t c = ql < 16u ? ql : 16u;
            // Incremental form of pivot_entry(j, d): the block of triple t rooted at heap node `root`
            // is number root + off_t with off_0 = -1, off_{t+1} = 8 * off_t + 1 (<= 30 levels: 32-bit).
            uint32_t j = 1, block8 = 0;
            int off_t = -1;
            unsigned e3 = 0;  // depth within the current triple
            const int table_levels = PHASE == 1 ? p.coarse_levels : p.pivot_levels;
            for (int d = 0; d < table_levels && l < r; d++) {
                const uint32_t m = l + ((r - l) >> 1);
                if (e3 == 0) block8 = (uint32_t)((int)j + off_t) * 8u;
                const uint32_t entry = block8 + ((1u << e3) | (j & ((1u << e3) - 1u)));
                if (++e3 == 3) { e3 = 0; off_t = off_t * 8 + 1; }
                const uint4 e = __ldg(p.pivots + entry);
                W4 ew;
                ew.w[0] = e.x; ew.w[1] = e.y; ew.w[2] = e.z; ew.w[3] = e.w;
                unsigned tb, pb;
                const unsigned mp = first_mismatch16(ew, p0, tb, pb);
                bool less, exact;
                uint32_t lcp;
                if (mp < c) {
                    less = tb < pb;
                    exact = tb != 0u;          // a zero may be end-of-text padding: keep the lcp bound conservative
                    lcp = exact ? mp : 0u;
                } else {  // tie on the stored prefix: decide on the text
                    const uint32_t start = MLR ? ((lcp_l < lcp_r ? lcp_l : lcp_r) & ~15u) : 0u;
                    lcp = thread_compare(p, ldr(p.sa + m), p0, p1, pat, ql, start, less);
                    exact = true;
                }
                if (less) { l = m + 1; lcp_l = lcp; j = 2 * j + 1; } else { r = m; lcp_r = lcp; lcp_r_exact = exact; j = 2 * j; }
            }
        }
        if constexpr (PHASE == 1) {
            p.keys[i] = l;
            p.ident[i] = (uint32_t)i;
        } else {
        // ---- remaining levels: sa[m] then text ----
        while (l < r) {
            const uint32_t m = l + ((r - l) >> 1);
            const uint32_t start = MLR ? ((lcp_l < lcp_r ? lcp_l : lcp_r) & ~15u) : 0u;
            bool less;
            const uint32_t lcp = probe(m, start, less);
            if (less) { l = m + 1; lcp_l = lcp; } else { r = m; lcp_r = lcp; lcp_r_exact = true; }
        }
        const unsigned long long lo = l;
        p.out_lo[i] = (uint32_t)lo;
        if (p.out_pos) p.out_pos[i] = lo < p.n ? (inl ? (WIDE == 1 ? ldr(p.saw + lo).x : ldr(p.sax + lo).x) : ldr(p.sa + lo)) : 0xffffffffu;  // (inl: the line the probes read)
        if (p.out_hi) {
            // Suffixes starting with q are contiguous from lo: gallop to bracket the end, then bisect.
            unsigned long long a = lo, b = range_end, step = 1;
            // The lower-bound search ended with r == lo; if r moved at all, lcp_r is lcp(q, suffix(lo))
            // and the first probe of the gallop is already answered.
            if (lo < p.n && r == lo && lcp_r_exact) {
                if (lcp_r >= ql) { a = lo + 1; step = 2; } else { b = lo; step = 0; }
            }
            // LCP-accelerated form: every suffix of [a, b) lies between one that starts with q (lcp = |q|) and suffix(b), so it
            // shares at least lcp(q, suffix(b)) bytes with q: once a probe has failed, later probes start there (rounded to 16).
            uint32_t lcp_b = 0;  // lcp(q, suffix(b)) once b comes from a failed probe
            if (MLR && step == 0) lcp_b = lcp_r;
            while (step) {
                const unsigned long long pr = a + step - 1;
                if (pr >= b) break;
                bool less;
                const uint32_t lcp = probe((uint32_t)pr, 0u, less);
                if (lcp >= ql) { a = pr + 1; step <<= 1; } else { b = pr; lcp_b = lcp; break; }
            }
            while (a < b) {
                const unsigned long long m = (a + b) >> 1;
                bool less;
                const uint32_t lcp = probe((uint32_t)m, MLR ? (lcp_b & ~15u) : 0u, less);
                if (lcp >= ql) a = m + 1; else { b = m; lcp_b = lcp; }
            }
            p.out_hi[i] = (uint32_t)a;
        }
        }  // PHASE != 1
    }
}

// Pivot table builder: node j of the implicit search tree -> 16 bytes of its pivot suffix.
__global__ void sa_pivot_kernel(const uint8_t* __restrict__ t, const uint32_t* __restrict__ sa, unsigned long long n, int levels,
                                uint4* __restrict__ table) {
    const unsigned long long total = 1ull << levels;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1; j < total;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long l = 0, r = n;
        const int depth = 63 - __clzll((long long)j);
        for (int b = depth - 1; b >= 0 && l < r; b--) {
            const unsigned long long m = (l + r) >> 1;
            if ((j >> b) & 1ull) l = m + 1; else r = m;
        }
        uint32_t w[4] = {0, 0, 0, 0};
        if (l < r) {
            const unsigned long long pos = sa[(l + r) >> 1];
            for (int k = 0; k < 16; k++)
                if (pos + k < n) w[k >> 2] |= (uint32_t)t[pos + k] << (8 * (k & 3));
        }
        table[pivot_entry(j, depth)] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

bool build_pivots(sst_sa* s) {
    int need = 1;
    while ((1ull << need) < s->n + 1) need++;
    // Default: all but the last ~3 levels of the search (measured on a 10^8 text: 24 levels 4.16,
    // 27 = full depth 3.74, 21 levels 3.97 Gpat/s: at the bottom every pattern ties with its own
    // suffix and needs sa[lo] anyway), within the memory budget (16 B per heap slot * 8/7).  The
    // levels below the table are plain probes (sa[m], then the text: two fills each).
    int levels = opt(OPT_SA_PIVOT_LEVELS) >= 0 ? (int)opt(OPT_SA_PIVOT_LEVELS) : 3 * ((need - 2) / 3);  // 10^8 text: 24 of 27; 3x10^9 text: 30 of 32 (1.98 vs 1.76 Gpat/s at 27)
    levels = 3 * (levels / 3);
    if (levels > 3 * ((need + 2) / 3)) levels = 3 * ((need + 2) / 3);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const double budget = opt(OPT_SA_TABLE_GB) >= 0 ? (double)opt(OPT_SA_TABLE_GB) * 1e9 : std::min(0.5 * (double)free_b, 64e9);
    if (levels > 30) levels = 30;  // the search kernel keeps heap indices and table offsets in 32 bits
    while (levels >= 3 && (double)pivot_table_entries(levels) * 16.0 > budget) levels -= 3;
    if (levels < 3) { s->pivot_levels = 0; return true; }
    cudaStream_t st = thread_stream(s->device);
    const unsigned long long entries = pivot_table_entries(levels);
    if (!SST_CUDA_OK(cudaMalloc(&s->d_pivots, entries * sizeof(uint4)))) return false;
    if (!SST_CUDA_OK(cudaMemsetAsync(s->d_pivots, 0, entries * sizeof(uint4), st))) return false;
    sa_pivot_kernel<<<grid_for((size_t)1 << std::min(levels, 30)), kThreads, 0, st>>>(s->d_text, s->d_sa, s->n, levels, s->d_pivots);
    if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaStreamSynchronize(st))) return false;
    s->pivot_levels = levels;
    return true;
}

template <int PL>
void launch_search(const SaParams& p, int mode, cudaStream_t st, int device) {
    const unsigned long long gpb = kThreads / PL;
    const unsigned grid = (unsigned)std::min<unsigned long long>((p.npat + gpb - 1) / gpb, (unsigned long long)sm_count(device) * 8);
    if (mode == SST_SA_MLR) sa_search_kernel<PL, true><<<grid, kThreads, 0, st>>>(p);
    else sa_search_kernel<PL, false><<<grid, kThreads, 0, st>>>(p);
}

}  // namespace
}  // namespace sst

using namespace sst;

static bool build_kmer(sst_sa* s);  // defined after sa_search_launch, which it uses to fill the table
static bool build_sax(sst_sa* s);
static bool build_cells(sst_sa* s);

extern "C" {

sst_sa_t* sst_sa_build_device(const uint8_t* d_text, size_t n, int device) {
    clear_error();
    if (n == 0 || !d_text) { set_error(SST_ERR_ARG, "empty text"); return nullptr; }
    if (n >= 0xfffffff0ull) { set_error(SST_ERR_UNSUPPORTED, "text must be shorter than 2^32 - 16 (u32 suffix array, sa_search.rs:35)"); return nullptr; }
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard g(device);
    if (!g.ok || !SST_CUDA_OK(cudaDeviceSynchronize())) return nullptr;  // the text may come from another stream
    auto* s = new sst_sa();
    s->device = device;
    s->n = n;
    bool ok = SST_CUDA_OK(cudaMalloc(&s->d_text, n + 64)) && SST_CUDA_OK(cudaMalloc(&s->d_sa, n * 4));
    cudaStream_t st = thread_stream(device);
    ok = ok && SST_CUDA_OK(cudaMemsetAsync(s->d_text + n, 0, 64, st)) &&
         SST_CUDA_OK(cudaMemcpyAsync(s->d_text, d_text, n, cudaMemcpyDeviceToDevice, st));
    ok = ok && build_sa_device(s->d_text, n, s->d_sa, device) && build_pivots(s) && build_kmer(s) && build_sax(s) && build_cells(s);
    if (!ok) { cudaFree(s->d_text); cudaFree(s->d_sa); cudaFree(s->d_pivots); cudaFree(s->d_kmer); cudaFree(s->d_sax); cudaFree(s->d_saw); cudaFree(s->d_text2); cudaFree(s->d_cells); delete s; return nullptr; }
    return s;
}

sst_sa_t* sst_sa_build(const uint8_t* text, size_t n, int device) {
    clear_error();
    if (n == 0 || !text) { set_error(SST_ERR_ARG, "empty text"); return nullptr; }
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard g(device);
    if (!g.ok) return nullptr;
    uint8_t* d = nullptr;
    cudaStream_t st = thread_stream(device);
    if (!SST_CUDA_OK(cudaMalloc(&d, n)) || !SST_CUDA_OK(cudaMemcpyAsync(d, text, n, cudaMemcpyHostToDevice, st)) ||
        !SST_CUDA_OK(cudaStreamSynchronize(st))) { cudaFree(d); return nullptr; }
    sst_sa_t* s = sst_sa_build_device(d, n, device);
    cudaFree(d);
    return s;
}

sst_sa_t* sst_sa_from_parts(const uint8_t* text, size_t n, const uint32_t* sa, int device) {
    clear_error();
    if (n == 0 || !text || !sa) { set_error(SST_ERR_ARG, "empty text"); return nullptr; }
    if (n >= 0xfffffff0ull) { set_error(SST_ERR_UNSUPPORTED, "text too long for a u32 suffix array"); return nullptr; }
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard g(device);
    if (!g.ok) return nullptr;
    auto* s = new sst_sa();
    s->device = device;
    s->n = n;
    cudaStream_t st = thread_stream(device);  // stream-ordered with build_pivots (see upload_keys in stree_build.cu)
    bool ok = SST_CUDA_OK(cudaMalloc(&s->d_text, n + 64)) && SST_CUDA_OK(cudaMalloc(&s->d_sa, n * 4)) &&
              SST_CUDA_OK(cudaMemsetAsync(s->d_text + n, 0, 64, st)) && SST_CUDA_OK(cudaMemcpyAsync(s->d_text, text, n, cudaMemcpyHostToDevice, st)) &&
              SST_CUDA_OK(cudaMemcpyAsync(s->d_sa, sa, n * 4, cudaMemcpyHostToDevice, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    // The reference asserts strict suffix order when it builds (sa_search.rs:36-38); an entry >= n would make the search
    // kernels read the text out of bounds and an unsorted array silently returns wrong bounds, so a caller's array is checked
    // the same way (strictly increasing suffixes with every entry < n are a permutation).  SA_VALIDATE=0: trusted callers.
    if (ok && opt(OPT_SA_VALIDATE)) {
        uint64_t bad = 0;
        ok = sst_sa_check(s, &bad) == SST_OK;
        if (ok && bad) {
            set_error(SST_ERR_ARG, "not a suffix array of the text: " + std::to_string(bad) + " adjacent pairs out of order or entries >= n (sa_search.rs:36-38)");
            ok = false;
        }
    }
    ok = ok && build_pivots(s) && build_kmer(s) && build_sax(s) && build_cells(s);
    if (!ok) { cudaFree(s->d_text); cudaFree(s->d_sa); cudaFree(s->d_pivots); cudaFree(s->d_kmer); cudaFree(s->d_sax); cudaFree(s->d_saw); cudaFree(s->d_text2); cudaFree(s->d_cells); delete s; return nullptr; }
    return s;
}

void sst_sa_free(sst_sa_t* s) {
    if (!s) return;
    DeviceGuard g(s->device);
    cudaFree(s->d_text);
    cudaFree(s->d_sa);
    cudaFree(s->d_pivots);
    cudaFree(s->d_kmer);
    cudaFree(s->d_sax);
    cudaFree(s->d_saw);
    cudaFree(s->d_text2);
    cudaFree(s->d_cells);
    delete s;
}

size_t sst_sa_len(const sst_sa_t* s) { return s ? s->n : 0; }

int sst_sa_get(const sst_sa_t* s, uint32_t* out_sa) {
    clear_error();
    if (!s || !out_sa) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    DeviceGuard g(s->device);
    if (!g.ok) return SST_ERR_CUDA;
    return SST_CUDA_OK(cudaMemcpy(out_sa, s->d_sa, s->n * 4, cudaMemcpyDeviceToHost)) ? SST_OK : SST_ERR_CUDA;
}

// out_sa[i] = sa[positions[i]] (0xffffffff for a position >= n): the occurrences sa[lo .. hi) of a pattern, or the few
// entries a host-side check needs, without copying the whole array (12 GB at 3x10^9).  Host buffers.
namespace sst { namespace {
__global__ void sa_gather_kernel(const uint32_t* __restrict__ sa, unsigned long long n, const unsigned long long* __restrict__ pos, size_t count,
                                 uint32_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = pos[i] < n ? sa[pos[i]] : 0xffffffffu;
}
} }
int sst_sa_gather(const sst_sa_t* s, const uint64_t* positions, size_t count, uint32_t* out_sa) {
    clear_error();
    if (!s || (count && (!positions || !out_sa))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (count == 0) return SST_OK;
    DeviceGuard g(s->device);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t st = thread_stream(s->device);
    unsigned long long* d_p = nullptr;
    uint32_t* d_o = nullptr;
    bool ok = SST_CUDA_OK(cudaMalloc(&d_p, count * 8)) && SST_CUDA_OK(cudaMalloc(&d_o, count * 4)) &&
              SST_CUDA_OK(cudaMemcpyAsync(d_p, positions, count * 8, cudaMemcpyHostToDevice, st));
    if (ok) {
        sa_gather_kernel<<<grid_for(count), kThreads, 0, st>>>(s->d_sa, s->n, d_p, count, d_o);
        ok = SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaMemcpyAsync(out_sa, d_o, count * 4, cudaMemcpyDeviceToHost, st)) &&
             SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(d_p); cudaFree(d_o);
    return ok ? SST_OK : SST_ERR_CUDA;
}

int sst_sa_check(const sst_sa_t* s, uint64_t* out_violations) {
    clear_error();
    if (!s || !out_violations) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    DeviceGuard g(s->device);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t st = thread_stream(s->device);
    unsigned long long* d_bad = nullptr;
    if (!SST_CUDA_OK(cudaMalloc(&d_bad, 8))) return SST_ERR_CUDA;
    unsigned long long bad = 0;
    bool ok = SST_CUDA_OK(cudaMemsetAsync(d_bad, 0, 8, st));
    if (ok) {
        sa_check_kernel<<<grid_for(s->n), kThreads, 0, st>>>(s->d_text, s->n, s->d_sa, d_bad);
        ok = SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st)) &&
             SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(d_bad);
    *out_violations = bad;
    return ok ? SST_OK : SST_ERR_CUDA;
}

// Sort buffers of the reordered SA batch: one set per (host thread, device), grown on demand.
namespace {
struct SaSortScratch {
    int device = -1;
    size_t cap = 0, tmp_bytes = 0;
    uint32_t *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr;
    void* tmp = nullptr;
    cudaEvent_t done = nullptr;
    ~SaSortScratch() {
        if (device < 0) return;
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return; }
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
        if (done) cudaEventDestroy(done);
        (void)cudaGetLastError();
        if (prev >= 0) cudaSetDevice(prev);
    }
    bool ensure(int dev, size_t npat, int begin_bit, int end_bit) {
        device = dev;
        if (!done && !SST_CUDA_OK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming))) return false;
        if (npat <= cap) return true;
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
        k0 = k1 = v0 = v1 = nullptr; tmp = nullptr; cap = 0;
        const size_t want = npat + npat / 8;
        cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
        size_t tb = 0;
        if (!SST_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (unsigned long long)want, 0, 32))) return false;
        (void)begin_bit; (void)end_bit;
        if (!SST_CUDA_OK(cudaMalloc(&k0, want * 4)) || !SST_CUDA_OK(cudaMalloc(&k1, want * 4)) || !SST_CUDA_OK(cudaMalloc(&v0, want * 4)) ||
            !SST_CUDA_OK(cudaMalloc(&v1, want * 4)) || !SST_CUDA_OK(cudaMalloc(&tmp, tb ? tb : 16)))
            return false;
        tmp_bytes = tb;
        cap = want;
        return true;
    }
};
thread_local SaSortScratch g_sa_sort[64];
}  // namespace

static int sa_search_launch(const sst_sa_t* s, const uint8_t* d_pats, const uint64_t* d_pat_off, unsigned long long pats_end,
                            size_t npat, int mode, uint32_t* d_out_lo, uint32_t* d_out_hi, uint32_t* d_out_pos, cudaStream_t st) {
    SaParams p{};
    p.text = s->d_text; p.sa = s->d_sa; p.n = s->n;
    p.pats = d_pats; p.pat_off = (const unsigned long long*)d_pat_off; p.npat = npat;
    p.out_lo = d_out_lo; p.out_hi = d_out_hi; p.out_pos = d_out_pos;
    p.pats_bytes = pats_end;  // end offset of the packed patterns (bounds the aligned 16-byte loads)
    p.pivots = s->d_pivots;
    p.pivot_levels = s->d_pivots ? std::min(s->pivot_levels, (int)opt(OPT_SA_USE_LEVELS)) : 0;
    p.kmer = s->d_kmer;
    p.kmer_k = s->d_kmer && opt(OPT_SA_USE_KMER) ? s->kmer_k : 0;
    p.sax = p.kmer_k && opt(OPT_SA_USE_INLINE) ? s->d_sax : nullptr;
    p.saw = p.kmer_k && opt(OPT_SA_USE_INLINE) ? s->d_saw : nullptr;
    p.cells = p.saw && opt(OPT_SA_USE_CELLS) ? s->d_cells : nullptr;
    p.text2 = p.kmer_k && opt(OPT_SA_USE_PACKED_TEXT) ? s->d_text2 : nullptr;
    const int lanes = (int)opt(OPT_SA_LANES);
    if (lanes <= 1) {
        const unsigned grid = (unsigned)std::min<unsigned long long>((npat + kSearchThreads - 1) / kSearchThreads, (unsigned long long)sm_count(s->device) * 8);
        // Opt-in (SST_SA_SORT_MIN=<patterns>): search in sorted order (see PHASE above): coarse pass over the cache-resident
        // top of the table, radix sort of (lower bound so far, index) on the bits the coarse pass has decided, main pass
        // through perm.  Measured: C3 3.28 vs 4.61 Gpat/s (a loss), C5 2.12 vs 2.03 Gpat/s at 12 coarse levels, worse with
        // more levels -- the DRAM fills are the text/SA probes of the last levels and of `hi` (777 B per pattern with or
        // without the sort, ncu), not the table, so reordering cannot pay for its two extra passes; off by default.
        const int coarse = std::min(p.pivot_levels, 3 * ((int)opt(OPT_SA_SORT_LEVELS) / 3));
        const unsigned long long sort_min = opt(OPT_SA_SORT_MIN) >= 0 ? (unsigned long long)opt(OPT_SA_SORT_MIN) : ~0ull;
        if (coarse >= 3 && npat >= sort_min && s->device >= 0 && s->device < 64) {
            SaSortScratch& sc = g_sa_sort[s->device];
            int nbits = 1;
            while (nbits < 32 && (1ull << nbits) <= s->n) nbits++;
            const int begin_bit = std::max(0, nbits - coarse - 1);
            if (!sc.ensure(s->device, npat, begin_bit, nbits)) return SST_ERR_CUDA;
            if (!SST_CUDA_OK(cudaStreamWaitEvent(st, sc.done, 0))) return SST_ERR_CUDA;
            p.keys = sc.k0; p.ident = sc.v0; p.coarse_levels = coarse;
            if (mode == SST_SA_MLR) sa_search_thread_kernel<true, 1><<<grid, kSearchThreads, 0, st>>>(p);
            else sa_search_thread_kernel<false, 1><<<grid, kSearchThreads, 0, st>>>(p);
            cub::DoubleBuffer<uint32_t> dk(sc.k0, sc.k1), dv(sc.v0, sc.v1);
            size_t tb = sc.tmp_bytes;
            if (!SST_CUDA_OK(cub::DeviceRadixSort::SortPairs(sc.tmp, tb, dk, dv, (unsigned long long)npat, begin_bit, nbits, st))) return SST_ERR_CUDA;
            p.perm = dv.Current();
            if (mode == SST_SA_MLR) sa_search_thread_kernel<true, 2><<<grid, kSearchThreads, 0, st>>>(p);
            else sa_search_thread_kernel<false, 2><<<grid, kSearchThreads, 0, st>>>(p);
            if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaEventRecord(sc.done, st))) return SST_ERR_CUDA;
            return SST_OK;
        }
        // One block of 256 threads per 256 patterns, however many that makes (the hardware block scheduler back-fills an SM as soon
        // as a block retires).  The grid used to be capped at 8 blocks per SM with a grid-stride loop: every resident thread then
        // walked the batch in lockstep (all of them at their k-mer load, then all at their entry) and the rate depended on how the
        // cap divided by the residency -- 5 or 10 blocks per SM on 5 resident ones were the slowest.  C3 16.3 -> 18.8 Gpat/s, C5
        // 6.96 -> 8.4 (profiles/r2_sa_grid_sweep.log; option SA_GRID = blocks per SM restores a cap).
        auto go = [&](auto kern) {
            unsigned long long blocks = (npat + kSearchThreads - 1) / kSearchThreads;
            if (opt(OPT_SA_GRID) > 0) blocks = std::min<unsigned long long>(blocks, (unsigned long long)sm_count(s->device) * (unsigned long long)opt(OPT_SA_GRID));
            kern<<<(unsigned)std::min<unsigned long long>(blocks, 0x7fffffffull), kSearchThreads, 0, st>>>(p);
        };
        const bool mlr = mode == SST_SA_MLR;
        if (p.saw) {
            // Register budget of four resident blocks of 256 threads' worth (62 registers, no spills) unless the option asks for five
            // (48 registers, a few spills).  With the capped grid five had looked better for long patterns (C5 7.7 vs 6.9) and for
            // the packed cells; with one pattern per thread the spill-free build wins everywhere: C3 19.3 vs 18.8, C5 8.7 vs 8.3 Gpat/s.
            if (opt(OPT_SA_MINB) == 5) { if (mlr) go(sa_search_thread_kernel<true, 0, 1, 5>); else go(sa_search_thread_kernel<false, 0, 1, 5>); }
            else if (mlr) go(sa_search_thread_kernel<true, 0, 1, 4>);
            else go(sa_search_thread_kernel<false, 0, 1, 4>);
        } else if (mlr) go(sa_search_thread_kernel<true, 0>);
        else go(sa_search_thread_kernel<false, 0>);
        return SST_CUDA_OK(cudaGetLastError()) ? SST_OK : SST_ERR_CUDA;
    }
    switch (lanes) {
        case 32: launch_search<32>(p, mode, st, s->device); break;
        case 16: launch_search<16>(p, mode, st, s->device); break;
        case 4: launch_search<4>(p, mode, st, s->device); break;
        case 2: launch_search<2>(p, mode, st, s->device); break;
        default: launch_search<8>(p, mode, st, s->device); break;
    }
    return SST_CUDA_OK(cudaGetLastError()) ? SST_OK : SST_ERR_CUDA;
}

}  // extern "C"

// ---- k-mer table builder: kmer[x] = lower bound of the k-base string x, computed by the search kernel itself ----
namespace sst {
namespace {
__global__ void kmer_max_byte_kernel(const uint8_t* __restrict__ t, size_t n, unsigned* __restrict__ out) {
    unsigned m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = max(m, (unsigned)t[i]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31u) == 0 && m) atomicMax(out, m);
}
// patterns x0 .. x0+count-1 as k bytes each (most significant base first), offsets relative to the chunk
__global__ void kmer_patterns_kernel(unsigned long long x0, unsigned count, int k, uint8_t* __restrict__ pats, unsigned long long* __restrict__ off) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i <= count; i += gridDim.x * blockDim.x) {
        off[i] = (unsigned long long)i * k;
        if (i < count) {
            const unsigned long long x = x0 + i;
            for (int j = 0; j < k; j++) pats[(size_t)i * k + j] = (uint8_t)((x >> (2 * (k - 1 - j))) & 3ull);
        }
    }
}
}  // namespace
}  // namespace sst

// sax[i] = {sa[i], the 15 bases after the first k of suffix(sa[i]) | complete << 31}
namespace sst {
namespace {
__global__ void sax_kernel(const uint8_t* __restrict__ t, const uint32_t* __restrict__ sa, unsigned long long n, int k, uint2* __restrict__ sax) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t pos = sa[i];
        const unsigned long long q = (unsigned long long)pos + (unsigned)k;
        uint32_t nx = 0;
        if (q + 15ull <= n) {  // (the text is followed by 64 zero bytes: the 16-byte window may run into them)
            const W4 w = load16_unaligned<false>(t + q, t + n + 64);
            nx = 0x80000000u | (pack16(w) >> 2);  // 16 bases packed, the last one dropped
        }
        sax[i] = make_uint2(pos, nx);
    }
}
// out[w] = bases 16 w .. 16 w + 15 of the text at 2 bits each, first base most significant (the text is followed by 64 zero bytes)
__global__ void pack_text_kernel(const uint8_t* __restrict__ t, unsigned long long n, unsigned long long words, uint32_t* __restrict__ out) {
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        if (16ull * w < n) {
            const uint4 x = *reinterpret_cast<const uint4*>(t + 16ull * w);
            W4 b;
            b.w[0] = x.x; b.w[1] = x.y; b.w[2] = x.z; b.w[3] = x.w;
            v = pack16(b);
        }
        out[w] = v;
    }
}
// saw[i] = {sa[i], the 48 bases after the first k of suffix(sa[i]) in three words of 16, first base most significant}; zero words
// when the suffix has fewer (sa[i] + k + 48 > n: the reader derives that from sa[i])
__global__ void saw_kernel(const uint8_t* __restrict__ t, const uint32_t* __restrict__ sa, unsigned long long n, int k, uint4* __restrict__ saw) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t pos = sa[i];
        const unsigned long long q = (unsigned long long)pos + (unsigned)k;
        uint4 e = make_uint4(pos, 0u, 0u, 0u);
        if (q + 48ull <= n) {
            e.y = pack16(load16_unaligned<false>(t + q, t + n + 64));
            e.z = pack16(load16_unaligned<false>(t + q + 16, t + n + 64));
            e.w = pack16(load16_unaligned<false>(t + q + 32, t + n + 64));
        }
        saw[i] = e;
    }
}
}  // namespace
}  // namespace sst

namespace sst {
namespace {
// cells[x] (128 bytes) = {start | overflow << 31, entries 0..4} {unused word, entries 5..9}, an entry = {sa, code hi, code lo},
// from the k-mer table and the first 32 bases of the 48-base entries (see sst_sa::d_cells)
__global__ void cells_kernel(const uint32_t* __restrict__ kmer, const uint4* __restrict__ saw, unsigned long long ncells, unsigned long long n, int k,
                             uint4* __restrict__ cells) {
    for (unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; x < ncells; x += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t start = kmer[x], cnt = kmer[x + 1] - start;
        bool overflow = cnt > 10u;
        uint4* c = cells + x * 8;
        for (uint32_t h = 0; h < 2u; h++) {
            uint32_t w[16];
            for (int i = 0; i < 16; i++) w[i] = 0xffffffffu;
            for (uint32_t i = 0; i < 5u && 5u * h + i < cnt; i++) {
                const uint4 e = saw[(size_t)start + 5u * h + i];
                overflow = overflow || (unsigned long long)e.x + (unsigned)k + 48ull > n;  // a suffix without its inlined bases (the text's last k + 47): ordinary path
                w[1 + 3 * i] = e.x; w[2 + 3 * i] = e.y; w[3 + 3 * i] = e.z;  // (the first 32 of the entry's 48 bases)
            }
            if (h == 0) w[0] = start;
            c[4 * h + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            c[4 * h + 2] = make_uint4(w[8], w[9], w[10], w[11]);
            c[4 * h + 3] = make_uint4(w[12], w[13], w[14], w[15]);
            if (h == 1) c[4] = make_uint4(w[0], w[1], w[2], w[3]);
            else c[0] = make_uint4(w[0], w[1], w[2], w[3]);  // (the flag is added below, once the second half has been seen)
        }
        if (overflow) reinterpret_cast<uint32_t*>(c)[0] = start | 0x80000000u;
    }
}
}  // namespace
}  // namespace sst

// Packed k-mer cells: only next to the 16-byte entries, for k <= 14 (128 bytes x 4^k: 8.6 GB at k = 13), texts below 2^31 bytes
// (bit 31 of the start is the flag) and within a quarter of the free memory; an optional accelerator like the others.
static bool build_cells(sst_sa* s) {
    if (!s->d_saw || !s->kmer_k || s->kmer_k > 14 || s->n >= (1ull << 31) || !opt(OPT_SA_CELLS)) return true;
    const unsigned long long ncells = 1ull << (2 * s->kmer_k);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (ncells * 128ull > free_b / 4) return true;
    if (cudaMalloc(&s->d_cells, ncells * 128ull) != cudaSuccess) { s->d_cells = nullptr; (void)cudaGetLastError(); return true; }
    cudaStream_t st = thread_stream(s->device);
    cells_kernel<<<sm_count(s->device) * 16, 256, 0, st>>>(s->d_kmer, s->d_saw, ncells, s->n, s->kmer_k, s->d_cells);
    if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaStreamSynchronize(st))) { cudaFree(s->d_cells); s->d_cells = nullptr; return false; }
    return true;
}

static bool build_sax(sst_sa* s) {
    if (!s->kmer_k || !opt(OPT_SA_INLINE)) return true;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    cudaStream_t st0 = thread_stream(s->device);
    // the 2-bit packed text (n / 4 bytes) for the compares behind the inlined bases
    if (opt(OPT_SA_PACKED_TEXT) && s->n / 4 <= free_b / 8) {
        const unsigned long long words = s->n / 16 + 16;
        if (cudaMalloc(&s->d_text2, words * 4) == cudaSuccess) {
            pack_text_kernel<<<sm_count(s->device) * 16, 256, 0, st0>>>(s->d_text, s->n, words, s->d_text2);
            if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaStreamSynchronize(st0))) { cudaFree(s->d_text2); s->d_text2 = nullptr; return false; }
        } else { s->d_text2 = nullptr; (void)cudaGetLastError(); }
    }
    // 48 bases per suffix (16-byte entries) when half of the free memory holds them (3x10^9 text: 48 GB of the ~130 GB left
    // on a 180 GB part: 7.7 vs 6.3 Gpat/s), else 15 bases (8-byte entries) within a third of it
    if (opt(OPT_SA_INLINE) != 15 && s->n * 16ull <= free_b / (size_t)opt(OPT_SA_INLINE_DIV) && SST_CUDA_OK(cudaMalloc(&s->d_saw, s->n * sizeof(uint4)))) {
        saw_kernel<<<sm_count(s->device) * 16, 256, 0, st0>>>(s->d_text, s->d_sa, s->n, s->kmer_k, s->d_saw);
        if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaStreamSynchronize(st0))) { cudaFree(s->d_saw); s->d_saw = nullptr; return false; }
        return true;
    }
    s->d_saw = nullptr;
    (void)cudaGetLastError();
    if (s->n * 8ull > free_b / 3) return true;  // an optional accelerator: never at the cost of the caller's memory
    if (!SST_CUDA_OK(cudaMalloc(&s->d_sax, s->n * sizeof(uint2)))) { s->d_sax = nullptr; (void)cudaGetLastError(); return true; }
    cudaStream_t st = thread_stream(s->device);
    sax_kernel<<<sm_count(s->device) * 16, 256, 0, st>>>(s->d_text, s->d_sa, s->n, s->kmer_k, s->d_sax);
    if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaStreamSynchronize(st))) { cudaFree(s->d_sax); s->d_sax = nullptr; return false; }
    return true;
}

static bool build_kmer(sst_sa* s) {
    if (!opt(OPT_SA_KMER) || s->n < 4096) return true;
    cudaStream_t st = thread_stream(s->device);
    unsigned* d_max = nullptr;
    unsigned h_max = 0;
    if (!SST_CUDA_OK(cudaMalloc(&d_max, 4)) || !SST_CUDA_OK(cudaMemsetAsync(d_max, 0, 4, st))) { cudaFree(d_max); return false; }
    kmer_max_byte_kernel<<<sm_count(s->device) * 8, 256, 0, st>>>(s->d_text, s->n, d_max);
    const bool ok0 = SST_CUDA_OK(cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(d_max);
    if (!ok0) return false;
    if (h_max > 3) return true;  // not a 2-bit alphabet: the pivot-prefix table serves every level
    int k = 1;
    while (k < 16 && (1ull << (2 * k + 1)) <= s->n) k++;  // 4^k nearest to n (in ratio): about one suffix per table cell
    k = std::min(k, (int)opt(OPT_SA_KMER_K));
    if (const int force = (int)opt(OPT_SA_KMER_FORCE); force >= 4 && force <= 16) k = force;  // tests: a table deeper than the text needs
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    while (k > 4 && ((1ull << (2 * k)) + 1) * 4ull > free_b / 4) k--;
    if (k < 4) return true;
    const unsigned long long cells = 1ull << (2 * k);
    const unsigned chunk = 1u << 22;
    uint8_t* d_p = nullptr;
    unsigned long long* d_o = nullptr;
    bool ok = SST_CUDA_OK(cudaMalloc(&s->d_kmer, (cells + 1) * 4)) && SST_CUDA_OK(cudaMalloc(&d_p, (size_t)chunk * k + 64)) &&
              SST_CUDA_OK(cudaMalloc(&d_o, ((size_t)chunk + 1) * 8));
    for (unsigned long long x0 = 0; ok && x0 < cells; x0 += chunk) {
        const unsigned cnt = (unsigned)std::min<unsigned long long>(chunk, cells - x0);
        kmer_patterns_kernel<<<sm_count(s->device) * 8, 256, 0, st>>>(x0, cnt, k, d_p, d_o);
        ok = sa_search_launch(s, d_p, (const uint64_t*)d_o, (unsigned long long)cnt * k, cnt, SST_SA_BINARY, s->d_kmer + x0, nullptr, nullptr, st) == SST_OK;
    }
    const uint32_t n32 = (uint32_t)s->n;
    ok = ok && SST_CUDA_OK(cudaMemcpyAsync(s->d_kmer + cells, &n32, 4, cudaMemcpyHostToDevice, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(d_p);
    cudaFree(d_o);
    if (!ok) { cudaFree(s->d_kmer); s->d_kmer = nullptr; return false; }
    s->kmer_k = k;  // set last: the searches above ran without the table
    return true;
}


// ---- replica of a finished index on another device (sst_multi_sa_*): device-to-device copies, no rebuild ----
namespace sst {
sst_sa* clone_sa(const sst_sa* src, int device) {
    clear_error();
    if (!src) { set_error(SST_ERR_ARG, "null argument"); return nullptr; }
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard g(device);
    if (!g.ok) return nullptr;
    if (device != src->device) {  // direct NVLink path when the devices can reach each other (otherwise the copy is staged)
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            (void)cudaGetLastError();
        }
    }
    auto* s = new sst_sa();
    s->device = device;
    s->n = src->n;
    s->pivot_levels = src->pivot_levels;
    s->kmer_k = src->kmer_k;
    cudaStream_t st = thread_stream(device);
    bool ok = true;
    auto copy = [&](auto*& dst, const auto* from, size_t bytes) {
        if (!ok || !from || !bytes) return;
        ok = SST_CUDA_OK(cudaMalloc(&dst, bytes)) && SST_CUDA_OK(cudaMemcpyPeerAsync(dst, device, from, src->device, bytes, st));
    };
    copy(s->d_text, src->d_text, s->n + 64);
    copy(s->d_sa, src->d_sa, s->n * 4);
    if (src->pivot_levels) copy(s->d_pivots, src->d_pivots, pivot_table_entries(src->pivot_levels) * sizeof(uint4));
    if (src->kmer_k) copy(s->d_kmer, src->d_kmer, ((1ull << (2 * src->kmer_k)) + 1) * 4);
    copy(s->d_sax, src->d_sax, s->n * sizeof(uint2));
    copy(s->d_saw, src->d_saw, s->n * sizeof(uint4));
    if (src->d_text2) copy(s->d_text2, src->d_text2, (s->n / 16 + 16) * 4);
    if (src->d_cells) copy(s->d_cells, src->d_cells, (1ull << (2 * src->kmer_k)) * 128ull);
    ok = ok && SST_CUDA_OK(cudaStreamSynchronize(st));
    if (!ok) { cudaFree(s->d_text); cudaFree(s->d_sa); cudaFree(s->d_pivots); cudaFree(s->d_kmer); cudaFree(s->d_sax); cudaFree(s->d_saw); cudaFree(s->d_text2); cudaFree(s->d_cells); delete s; return nullptr; }
    return s;
}
}  // namespace sst

extern "C" {

int sst_sa_search_device(const sst_sa_t* s, const uint8_t* d_pats, const uint64_t* d_pat_off, size_t npat, int mode,
                         uint32_t* d_out_lo, uint32_t* d_out_hi, uint32_t* d_out_pos, void* stream) {
    clear_error();
    if (!s || (npat && (!d_pat_off || !d_out_lo))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (mode != SST_SA_BINARY && mode != SST_SA_MLR) { set_error(SST_ERR_ARG, "unknown SA search mode"); return SST_ERR_ARG; }
    if (npat == 0) return SST_OK;
    DeviceGuard g(s->device);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;  // NULL == the CUDA legacy default stream
    // total pattern bytes (bounds the aligned loads); one small read-back on the caller's stream
    unsigned long long total = 0;
    if (!SST_CUDA_OK(cudaMemcpyAsync(&total, d_pat_off + npat, 8, cudaMemcpyDeviceToHost, st)) || !SST_CUDA_OK(cudaStreamSynchronize(st)))
        return SST_ERR_CUDA;
    return sa_search_launch(s, d_pats, d_pat_off, total, npat, mode, d_out_lo, d_out_hi, d_out_pos, st);
}

// The reference threads a probe counter through every search function (`cnt: &mut usize`, sa_search.rs:98-112,423-436: one
// increment per loop iteration).  This entry point runs exactly that loop -- the plain binary search over [0, n), no table,
// on the sub-warp kernel -- and returns sa[l] and the number of iterations per pattern.  A tracing aid, not the fast path.
int sst_sa_search_probes(const sst_sa_t* s, const uint8_t* pats, const uint64_t* pat_off, size_t npat, uint32_t* out_pos,
                         uint32_t* out_probes) {
    clear_error();
    if (!s || (npat && (!pat_off || !out_probes))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (npat == 0) return SST_OK;
    DeviceGuard g(s->device);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t st = thread_stream(s->device);
    const size_t byte0 = pat_off[0], bytes = pat_off[npat] - byte0;
    uint8_t* d_p = nullptr;
    uint64_t* d_o = nullptr;
    uint32_t *d_lo = nullptr, *d_pos = nullptr, *d_pr = nullptr;
    bool ok = SST_CUDA_OK(cudaMalloc(&d_p, bytes + 64)) && SST_CUDA_OK(cudaMalloc(&d_o, (npat + 1) * 8)) && SST_CUDA_OK(cudaMalloc(&d_lo, npat * 4)) &&
              SST_CUDA_OK(cudaMalloc(&d_pos, npat * 4)) && SST_CUDA_OK(cudaMalloc(&d_pr, npat * 4)) &&
              (!bytes || SST_CUDA_OK(cudaMemcpyAsync(d_p, pats + byte0, bytes, cudaMemcpyHostToDevice, st))) &&
              SST_CUDA_OK(cudaMemcpyAsync(d_o, pat_off, (npat + 1) * 8, cudaMemcpyHostToDevice, st));
    if (ok) {
        SaParams p{};
        p.text = s->d_text; p.sa = s->d_sa; p.n = s->n;
        p.pats = d_p - byte0; p.pat_off = (const unsigned long long*)d_o; p.npat = npat; p.pats_bytes = pat_off[npat];
        p.out_lo = d_lo; p.out_pos = d_pos; p.out_probes = d_pr;
        launch_search<8>(p, SST_SA_BINARY, st, s->device);
        ok = SST_CUDA_OK(cudaGetLastError()) && (!out_pos || SST_CUDA_OK(cudaMemcpyAsync(out_pos, d_pos, npat * 4, cudaMemcpyDeviceToHost, st))) &&
             SST_CUDA_OK(cudaMemcpyAsync(out_probes, d_pr, npat * 4, cudaMemcpyDeviceToHost, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(d_p); cudaFree(d_o); cudaFree(d_lo); cudaFree(d_pos); cudaFree(d_pr);
    return ok ? SST_OK : SST_ERR_CUDA;
}

// Per (host thread, device) staging ring for the host-buffer SA path (see Staging in stree_search.cu).
namespace {
struct SaStaging {
    size_t cap_pat = 0, cap_bytes = 0;
    uint8_t* p[3] = {};
    uint64_t* o[3] = {};
    uint32_t *lo[3] = {}, *hi[3] = {}, *pos[3] = {};
    cudaEvent_t e_in[3] = {}, e_k[3] = {}, e_out[3] = {};
    bool events = false;
    int device = -1;
    ~SaStaging() {
        if (device < 0) return;
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return; }
        for (int b = 0; b < 3; b++) {
            cudaFree(p[b]); cudaFree(o[b]); cudaFree(lo[b]); cudaFree(hi[b]); cudaFree(pos[b]);
            if (events) { cudaEventDestroy(e_in[b]); cudaEventDestroy(e_k[b]); cudaEventDestroy(e_out[b]); }
        }
        (void)cudaGetLastError();
        if (prev >= 0) cudaSetDevice(prev);
    }
    bool ensure(size_t npat, size_t bytes) {
        if (!events) {
            for (int b = 0; b < 3; b++)
                if (!SST_CUDA_OK(cudaEventCreateWithFlags(&e_in[b], cudaEventDisableTiming)) ||
                    !SST_CUDA_OK(cudaEventCreateWithFlags(&e_k[b], cudaEventDisableTiming)) ||
                    !SST_CUDA_OK(cudaEventCreateWithFlags(&e_out[b], cudaEventDisableTiming)))
                    return false;
            events = true;
        }
        if (npat > cap_pat) {
            for (int b = 0; b < 3; b++) {
                cudaFree(o[b]); cudaFree(lo[b]); cudaFree(hi[b]); cudaFree(pos[b]);
                o[b] = nullptr; lo[b] = hi[b] = pos[b] = nullptr;
                if (!SST_CUDA_OK(cudaMalloc(&o[b], (npat + 1) * 8)) || !SST_CUDA_OK(cudaMalloc(&lo[b], npat * 4)) ||
                    !SST_CUDA_OK(cudaMalloc(&hi[b], npat * 4)) || !SST_CUDA_OK(cudaMalloc(&pos[b], npat * 4))) { cap_pat = 0; return false; }
            }
            cap_pat = npat;
        }
        if (bytes > cap_bytes) {
            const size_t want = bytes + bytes / 4;
            for (int b = 0; b < 3; b++) {
                cudaFree(p[b]);
                p[b] = nullptr;
                if (!SST_CUDA_OK(cudaMalloc(&p[b], want + 64))) { cap_bytes = 0; return false; }
            }
            cap_bytes = want;
        }
        return true;
    }
};
thread_local SaStaging g_sa_staging[64];
}  // namespace

// Host buffers: patterns go down in chunks through a three-buffer ring (H2D | kernel | D2H on three
// streams).  Offsets stay absolute: the kernel is given a pattern base shifted by the chunk's first offset.
int sst_sa_search(const sst_sa_t* s, const uint8_t* pats, const uint64_t* pat_off, size_t npat, int mode, uint32_t* out_lo,
                  uint32_t* out_hi, uint32_t* out_pos) {
    clear_error();
    if (!s || (npat && (!pat_off || !out_lo))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (mode != SST_SA_BINARY && mode != SST_SA_MLR) { set_error(SST_ERR_ARG, "unknown SA search mode"); return SST_ERR_ARG; }
    if (npat == 0) return SST_OK;
    const int dev = s->device;
    if (dev < 0 || dev >= 64) { set_error(SST_ERR_ARG, "device index out of range"); return SST_ERR_ARG; }
    DeviceGuard g(dev);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t s_in = thread_copy_stream(dev, 0), s_k = thread_stream(dev), s_out = thread_copy_stream(dev, 1);
    if (!s_in || !s_k || !s_out) return SST_ERR_CUDA;
    const size_t chunk = (size_t)opt(OPT_SA_CHUNK);
    const size_t nchunks = div_ceil(npat, chunk);
    size_t max_bytes = 0;
    for (size_t c = 0; c < nchunks; c++) {
        const size_t a = c * chunk, b = std::min(npat, a + chunk);
        max_bytes = std::max<size_t>(max_bytes, pat_off[b] - pat_off[a]);
    }
    SaStaging& sg = g_sa_staging[dev];
    sg.device = dev;
    if (!sg.ensure(std::min(chunk, npat), max_bytes)) return SST_ERR_CUDA;
    int rc = SST_OK;
    for (size_t c = 0; c < nchunks && rc == SST_OK; c++) {
        const int b = (int)(c % 3);
        const size_t a = c * chunk, e = std::min(npat, a + chunk), cnt = e - a;
        const size_t byte0 = pat_off[a], bytes = pat_off[e] - byte0;
        if (c >= 3 && !SST_CUDA_OK(cudaStreamWaitEvent(s_in, sg.e_out[b], 0))) { rc = SST_ERR_CUDA; break; }
        if ((bytes && !SST_CUDA_OK(cudaMemcpyAsync(sg.p[b], pats + byte0, bytes, cudaMemcpyHostToDevice, s_in))) ||
            !SST_CUDA_OK(cudaMemcpyAsync(sg.o[b], pat_off + a, (cnt + 1) * 8, cudaMemcpyHostToDevice, s_in)) ||
            !SST_CUDA_OK(cudaEventRecord(sg.e_in[b], s_in)) || !SST_CUDA_OK(cudaStreamWaitEvent(s_k, sg.e_in[b], 0))) { rc = SST_ERR_CUDA; break; }
        rc = sa_search_launch(s, sg.p[b] - byte0, sg.o[b], pat_off[e], cnt, mode, sg.lo[b], out_hi ? sg.hi[b] : nullptr,
                              out_pos ? sg.pos[b] : nullptr, s_k);
        if (rc != SST_OK) break;
        if (!SST_CUDA_OK(cudaEventRecord(sg.e_k[b], s_k)) || !SST_CUDA_OK(cudaStreamWaitEvent(s_out, sg.e_k[b], 0)) ||
            !SST_CUDA_OK(cudaMemcpyAsync(out_lo + a, sg.lo[b], cnt * 4, cudaMemcpyDeviceToHost, s_out)) ||
            (out_hi && !SST_CUDA_OK(cudaMemcpyAsync(out_hi + a, sg.hi[b], cnt * 4, cudaMemcpyDeviceToHost, s_out))) ||
            (out_pos && !SST_CUDA_OK(cudaMemcpyAsync(out_pos + a, sg.pos[b], cnt * 4, cudaMemcpyDeviceToHost, s_out))) ||
            !SST_CUDA_OK(cudaEventRecord(sg.e_out[b], s_out))) { rc = SST_ERR_CUDA; break; }
    }
    if (!SST_CUDA_OK(cudaStreamSynchronize(s_in)) || !SST_CUDA_OK(cudaStreamSynchronize(s_k)) || !SST_CUDA_OK(cudaStreamSynchronize(s_out)))
        rc = rc == SST_OK ? SST_ERR_CUDA : rc;
    return rc;
}

}  // extern "C"
