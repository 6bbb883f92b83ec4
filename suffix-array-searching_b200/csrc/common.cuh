// common.cuh -- internal declarations shared by the translation units of libsst_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "sst_b200.h"

namespace sst {

constexpr int kMaxLevels = 12;
constexpr uint32_t kMax = SST_MAX;   // static-search-tree/src/node.rs:5
constexpr int kNodeSlots = 16;       // BTreeNode<16>, 64 bytes (node.rs:7-11)

void set_error(int status, const std::string& msg);
void clear_error();
bool cuda_ok(cudaError_t e, const char* what, const char* file, int line);

#define SST_CUDA_OK(call) ::sst::cuda_ok((call), #call, __FILE__, __LINE__)

// Selects `device` for the current host thread for the lifetime of the guard.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device);
    ~DeviceGuard();
};

// One non-blocking stream per (host thread, device); created lazily, never destroyed.
cudaStream_t thread_stream(int device);
// Two extra per-thread streams used by the chunked host-buffer pipeline.
cudaStream_t thread_copy_stream(int device, int which);

void configure_l2_fetch(int device);  // must run with `device` current
int sm_count(int device);             // cached per device
unsigned cur_sms();                   // sm_count of the calling thread's current device (launch sites run under a DeviceGuard)

// ---- library options -------------------------------------------------------------------------
// Every tuning / A-B switch of the library lives in ONE table: a default, overridden ONCE when the library is loaded by
// the environment variable SST_<NAME> (so a tool can still be started as `SST_BK_R=256 python tool.py`), and afterwards
// only through sst_set_option() (tests, bench).  Nothing on a launch path calls getenv: the table is an array of
// atomics, so sst_query* stay re-entrant while another thread changes an option.  -1 means "auto" where noted.
//   X(name, default, min, max)
#define SST_OPTION_LIST(X)                                                                                                   \
    X(DEBUG, 0, 0, 9)                                                                                                        \
    X(L2_FETCH, 64, 0, 128)              /* cudaLimitMaxL2FetchGranularity set at first use; 0 = leave */                    \
    X(NO_BIND, 0, 0, 1)                  /* sst_multi_* workers do not pin themselves to the GPU's CPUs */                   \
    /* direct kernels (stree_search.cu) */                                                                                    \
    X(SCHEME, -1, -1, 8)                 /* what SST_SCHEME_AUTO resolves to for plain trees; -1 = the measured rule */       \
    X(HINTS, 3, 0, 3)                                                                                                        \
    X(L1_LEVEL_KB, 256, 0, 1 << 20)                                                                                          \
    X(USE_C5, 1, 0, 1)                                                                                                       \
    X(C5, 0, 0, 1)                       /* build the 16-bit copy of the last inner level (measured slower: off) */          \
    X(THREADS, 1024, 32, 1024)           /* CTA size of stree_search_fast; rounded down to a multiple of 32 */                \
    X(WAVES, 1, 1, 64)                                                                                                       \
    X(GRID_CAP, 0, 0, 1 << 20)                                                                                               \
    X(PERSIST, 0, 0, 300)                                                                                                    \
    X(PERSIST_MB, -1, -1, 1 << 20)                                                                                           \
    X(PGROUP, 1, 0, 1)                                                                                                       \
    X(PT, 1, 1, 2)                                                                                                           \
    X(T, 2, 1, 2)                                                                                                            \
    X(TABLE_G, 2, 2, 4)                                                                                                      \
    X(TABLE_MIN_NQ, 1 << 17, 0, 1ll << 40)                                                                                   \
    X(CHUNK, 1 << 22, 1024, 1ll << 32)   /* queries per chunk of the host-buffer path */                                     \
    X(HOST_GRID_CAP, 0, -1, 1 << 20)                                                                                         \
    /* reordered-batch pipeline (bucketed.cu) */                                                                              \
    X(BK_AUTO_MIN_N, 1 << 25, 0, 1ll << 40)                                                                                  \
    X(BK_AUTO_MIN_NQ, -1, -1, 1ll << 40) /* -1 = measured crossover by tree size */                                           \
    X(BK_MIN_N, 1 << 22, 0, 1ll << 40)   /* smaller trees get no pipeline arrays */                                          \
    X(BK_G, -1, -1, 16)                  /* keys per separator: 8 / 16; -1 = by size */                                       \
    X(BK_R, -1, -1, 32768)               /* separators per bucket (power of two >= 64); -1 = by size */                       \
    X(BK_CHUNK, 32768, 16384, 1 << 24)   /* queries per search work item */                                                   \
    X(BK_TIMING, 0, 0, 2)                /* per-stage CUDA-event times (synchronises; bench/tools only) */                    \
    X(BK_CHUNK2_LOG2, 15, 14, 24)        /* log2 of the queries per search work item of the V2 pipeline */                    \
    X(BK_COMPACT, 1, 0, 1)               /* Compact layout: keep a dense copy of the keys so that large batches take the pipeline */ \
    X(BK_SEP16, -1, -1, 1)               /* 16-bit separators (65536 per bucket, 8 keys each): -1 = above 2^28 slots, 0 = never, 1 = always */ \
    X(BK_SUB_LOG2, 30, 20, 30)           /* log2 of the queries per pipeline run (sub-batch); halved when the scratch does not fit */ \
    X(BK_HYBRID, 1, 0, 4)                                                                                                    \
    X(BK_VEC, 1, 0, 1)                                                                                                       \
    X(BK_MOVE_THREADS, 1024, 512, 1024)                                                                                      \
    X(BK_PREFETCH, 1, 0, 1)                                                                                                  \
    X(BK_MOVE_CTAS, -1, -1, 8)                                                                                               \
    X(BK_PROBE, 1, 0, 1)                                                                                                     \
    /* suffix arrays (sa.cu) */                                                                                               \
    X(SA_PIVOT_LEVELS, -1, -1, 33)       /* -1 = all but the last ~3 levels */                                                \
    X(SA_TABLE_GB, -1, -1, 1 << 10)      /* pivot-table budget; -1 = min(half of free memory, 64 GB) */                       \
    X(SA_USE_LEVELS, 64, 0, 64)                                                                                              \
    X(SA_USE_KMER, 1, 0, 1)                                                                                                  \
    X(SA_USE_INLINE, 1, 0, 1)                                                                                                \
    X(SA_CELLS, 1, 0, 1)                 /* build the packed 64-byte k-mer cells (range + first five entries in one line) */ \
    X(SA_USE_CELLS, 1, 0, 1)             /* use them */ \
    X(SA_LANES, 1, 1, 32)                                                                                                    \
    X(SA_SORT_LEVELS, 12, 0, 30)                                                                                             \
    X(SA_SORT_MIN, -1, -1, 1ll << 62)    /* batches of at least this many patterns search in sorted order; -1 = never */      \
    X(SA_MINB, 0, 0, 5)                                                                                                      \
    X(SA_INLINE, 1, 0, 32)               /* 0 = no inlined bases, 15 / 32 = 8- / 16-byte entries, 1 = widest that fits */     \
    X(SA_GRID, 0, 0, 1 << 20)            /* cap on the blocks per SM of the search kernel's grid; 0 = one block per 256 patterns */ \
    X(SA_PACKED_TEXT, 1, 0, 1)           /* build the 2-bit packed text next to the inlined bases */                          \
    X(SA_USE_PACKED_TEXT, 1, 0, 1)       /* compare behind the inlined bases through it */                                    \
    X(SA_INLINE_DIV, 2, 1, 64)                                                                                               \
    X(SA_KMER, 1, 0, 1)                                                                                                      \
    X(SA_KMER_K, 16, 1, 16)                                                                                                  \
    X(SA_KMER_FORCE, 0, 0, 16)                                                                                               \
    X(SA_CHUNK, 1 << 20, 16, 1ll << 32)  /* patterns per chunk of the host-buffer path */                                     \
    X(SA_L2_64B, 1, 0, 2)                /* random loads of the SA search ask L2 for 64-byte fills (0 = plain __ldg) */        \
    X(SA_VALIDATE, 1, 0, 1)              /* sst_sa_from_parts checks the caller's suffix array (sa_search.rs:36-38) */

enum Opt : int {
#define SST_OPT_ENUM(name, dflt, lo, hi) OPT_##name,
    SST_OPTION_LIST(SST_OPT_ENUM)
#undef SST_OPT_ENUM
    OPT_COUNT
};
long long opt(Opt o);
size_t max_smem_optin(int device);
bool device_usable(int device);

inline size_t div_ceil(size_t a, size_t b) { return (a + b - 1) / b; }

// Eytzinger layout helpers (shared by the builder and the search kernel): in-order rank of BFS slot k.
__host__ __device__ inline unsigned long long eytz_subtree_size(unsigned long long k, int depth, unsigned long long n, int H) {
    // nodes of the subtree rooted at BFS slot k (depth `depth`, root = 0) in a tree of n slots whose last level H is filled from the left
    if (k > n) return 0;
    const int below = H - depth;                       // levels below k down to level H
    const unsigned long long full = (1ull << below) - 1ull;   // k's level .. H-1
    const unsigned long long first = k << below;       // first slot of k's span on level H
    unsigned long long last_level = 0;
    if (first <= n) { last_level = n + 1 - first; const unsigned long long cap = 1ull << below; if (last_level > cap) last_level = cap; }
    return full + last_level;
}
__host__ __device__ inline unsigned long long eytz_rank(unsigned long long k, unsigned long long n, int H) {
#ifdef __CUDA_ARCH__
    const int depth = 63 - __clzll((long long)k);
#else
    const int depth = 63 - __builtin_clzll(k);
#endif
    unsigned long long base = 0, c = 1;
    for (int b = depth - 1; b >= 0; b--) {
        const unsigned bit = (unsigned)((k >> b) & 1ull);
        const int d = depth - 1 - b;                   // depth of c
        if (bit) base += eytz_subtree_size(2 * c, d + 1, n, H) + 1;
        c = 2 * c + bit;
    }
    return base + eytz_subtree_size(2 * k, depth + 1, n, H);
}



}  // namespace sst

// What the search kernels need to know about an index.  Passed by value (__grid_constant__).
// All positions are in u32 "slots" (4 bytes); a node is 16 slots.
struct SstTreeView {
    const uint32_t* tree;                          // device base of the node array
    unsigned long long level_slot[sst::kMaxLevels]; // first slot of each level (within a part for COMPACT)
    uint32_t mult[sst::kMaxLevels];                // s' = s * mult[h] + 16 * count after level h
    int levels;                                    // offsets.len()
    int variant;                                   // SST_PLAIN .. SST_MAP
    uint32_t node_b;                               // B: 16 or 15
    uint32_t shift;                                // part = q >> shift
    unsigned long long parts;
    uint32_t start_mul;                            // first slot = part * start_mul (SIMPLE/L1: 16, OVERLAPPING: 16 - overlap)
    unsigned long long part_stride;                // COMPACT: slots per part (bpp * 16)
    const uint32_t* prefix_map;                    // MAP
    unsigned long long leaf_slots;                 // slots of the leaf level (per part for COMPACT)
    unsigned long long n;                          // number of keys
    // sorted-array index of record for the partitioned layouts
    const uint32_t* part_start;                    // [parts + 1] index of each part's first key
    const unsigned long long* part_pos;            // [parts] leaf slot of each part's first key
    // COMPACT only: dense copy of every part's non-leaf nodes (part p at upper + p * upper_stride slots).
    // The image interleaves them with the leaves over the whole allocation, which costs a TLB miss per
    // level (measured: 13.9 vs 28 Gq/s for the same DRAM traffic); the copy keeps them in a few pages.
    const uint32_t* upper;
    unsigned long long upper_stride;
};

// Auxiliary arrays of the reordered-batch pipeline (bucketed.cu); nb == 0: not built for this index.
struct BkAux {
    uint32_t* d_dense = nullptr;  // COMPACT only: dense copy of the sorted keys, the "leaf level" the pipeline reads (the image interleaves the parts' levels)
    uint32_t* d_sep = nullptr;    // [nb * r] last key of every g-key block of the leaf level, 0xffffffff beyond the keys
    uint32_t* d_split = nullptr;  // [nb + 1] split[0] = 0, split[b] = last key before bucket b, split[nb] = MAX
    uint16_t* d_bt = nullptr;     // bucket table over the top 12 key bits
    uint16_t* d_jump = nullptr;   // [nb][8200] per-bucket jump table into its separators
    uint2* d_meta = nullptr;      // [nb] {lo, shift} of the jump table
    uint16_t* d_sep16 = nullptr;  // 16-bit mode (above 2^28 slots): [nb * r] low `shift` bits of sep - lo, i.e. the separator's offset inside its jump cell
    unsigned cells = 0;           // jump cells per bucket: r (32-bit separators) or r / 2 (16-bit mode)
    unsigned nb = 0, nbp = 0, r = 0, bits = 0;
    unsigned g = 8;               // keys per separator: 8 (half node) or 16 (node)
    unsigned long long m8 = 0;    // blocks of g keys that hold keys
    unsigned long long n_flat = 0;  // slots of the sorted array the pipeline searches: n, or the flat leaf level of Simple / L1 / Overlapping
};

struct sst_index {
    int device = 0;
    int variant = SST_PLAIN;
    uint32_t node_b = 16;
    uint32_t flags = 0;
    size_t n = 0;
    size_t n_blocks = 0;          // nodes in the image; one MAX guard node follows on the device
    uint32_t* d_tree = nullptr;
    int levels = 0;
    size_t offsets[sst::kMaxLevels] = {};      // node offsets, as the reference's `offsets`
    size_t layer_sizes[sst::kMaxLevels] = {};  // nodes per level (per part where the reference says so)
    size_t layer_blocks[sst::kMaxLevels] = {}; // nodes actually allocated per level
    // partition parameters (partitioned_s_tree.rs:19-32)
    size_t shift = 0, parts = 1, bpp = 0, l1 = 0, overlap = 0, max_bucket = 0;
    bool has_overlap = false;
    uint32_t* d_prefix_map = nullptr;
    size_t prefix_map_len = 0;
    uint32_t* d_part_start = nullptr;             // [parts + 1]
    unsigned long long* d_part_pos = nullptr;     // [parts]
    uint32_t* d_upper = nullptr;                  // COMPACT: dense copy of the non-leaf nodes of every part
    size_t l1_field = 0;                          // the `l1` struct field (max(l1,16) for OL)
    size_t eytz_words = 0;                        // SST_EYTZINGER: n + 1 array entries (entry 0 = u32::MAX)
    int eytz_h = 0;                               // SST_EYTZINGER: depth of the last level
    // shared-memory rank table replacing levels [0, top_level) of the plain B=16 tree
    int top_level = 0;
    size_t top_nbound = 0;
    uint16_t* d_top_table = nullptr;              // [2^15 + 1] separators-before-bucket counts
    uint16_t* d_top_low = nullptr;                // [top_nbound] low 16 bits of each separator
    // 16-bit compressed copy of the last internal level (plain B=16 trees whose level is HBM/L2 sized)
    uint16_t* d_c5 = nullptr;                     // [nodes * 16] separators minus the node's base
    uint32_t* d_h5 = nullptr;                     // [nodes] base (first separator), 0xffffffff = use the exact node
    BkAux bk;                                     // reordered-batch pipeline (plain B=16 trees, 2^22..2^28 keys)
    size_t auto_min_nq = 0;                       // SCHEME_AUTO takes the pipeline from this batch size on; 0 = the rule of resolve_scheme, set by sst_query_calibrate
    bool persist_ok = false;                      // persisting-L2 carve-out configured on this device
    size_t persist_window_max = 0;
    SstTreeView view{};
};

namespace sst {
// builders (stree_build.cu)
sst_index* build_plain(const uint32_t* d_sorted, bool sorted_is_owned_leaf, size_t n, uint32_t node_b, uint32_t flags, int device);
sst_index* build_partitioned(const uint32_t* d_sorted, size_t n, uint32_t b, int variant, int device);
sst_index* build_eytzinger(const uint32_t* d_sorted, size_t n, int device);
void finalize_view(sst_index* idx);
bool build_top_table(sst_index* idx, const uint32_t* d_sorted);
bool build_compressed_level(sst_index* idx);
// search (stree_search.cu)
int launch_query(const sst_index* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_vals, unsigned long long* d_idx,
                 int scheme, cudaStream_t stream);
int query_launch_count(const sst_index* idx, int scheme, size_t nq, bool want_idx);
// reordered-batch pipeline (bucketed.cu)
bool build_bucket_aux(sst_index* idx, const uint32_t* d_sorted = nullptr);  // d_sorted: the builder's input (COMPACT keeps a copy)
void free_bucket_aux(sst_index* idx);
bool bucketed_eligible(const sst_index* idx);
int last_stage_ms(double* out, int n);  // stage times of this thread's last pipeline run under SST_BK_TIMING=1
int launch_bucketed(const sst_index* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_vals, unsigned long long* d_idx,
                    cudaStream_t stream);
int reserve_bucketed(const sst_index* idx, size_t nq, bool want_idx);  // pre-size the calling thread's scratch
void release_bucketed_scratch();
size_t bucketed_sub_batch();                                           // queries per pipeline run (larger batches run in several)                                       // free the calling thread's scratch on every device
// suffix arrays (sa.cu): replica of a finished index on another device, copied device to device
struct sst_sa* clone_sa(const struct sst_sa* src, int device);
}  // namespace sst
