// stree_build.cu -- GPU layout builders: sorted u32 keys in HBM -> node-packed S+-tree image.
//
// Replaces (reference paths relative to static-search-tree/src):
//   STree::new_params                         s_tree.rs:72-176
//   PartitionedSTree::get_part_size/max_overlap   partitioned_s_tree.rs:111-227
//   PartitionedSTree<Compact>::try_new        partitioned_s_tree.rs:241-351
//   PartitionedSTree<Simple|L1|Overlapping|Map>::try_new   partitioned_s_tree.rs:364-649
//
// The reference fills the image with sequential loops (a running write index over the keys,
// then bottom-up copies from the leaf layer).  Here every slot of the image is computed
// independently from a closed form, one thread per slot:
//   * leaf slots gather from the sorted array through a per-part position table (a max-plus
//     scan of the bucket histogram done on the host over <= 2^27 parts);
//   * internal slots gather the separator straight from the sorted array (plain, Map) or from
//     the already-written leaf layer (partitioned layouts, whose separators include fill keys).
// The image is bit-identical to the reference's (tests compare it against the oracle).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace sst {
namespace {

constexpr int kBuildThreads = 256;

void free_index(sst_index* idx);  // defined below

inline unsigned grid_for(size_t work, int threads = kBuildThreads) {
    size_t b = div_ceil(work, (size_t)threads);
    return (unsigned)std::min<size_t>(b, (size_t)1 << 30);
}

// ---- shape math: s_tree.rs:22-45 TreeBase<B> ---------------------------------------------------
size_t tb_prev_keys(size_t n, size_t B) { return div_ceil(div_ceil(n, B), B + 1) * B; }
size_t tb_height(size_t n, size_t B) {
    size_t h = 1;
    while (n > B) { n = tb_prev_keys(n, B); h++; }
    return h;
}
size_t tb_layer_size(size_t n, size_t h, size_t height, size_t B) {
    for (size_t i = h; i + 1 < height; i++) n = tb_prev_keys(n, B);
    return n;
}
size_t ipow(size_t b, size_t e) { size_t r = 1; while (e--) r *= b; return r; }

// ---- input validation: s_tree.rs:87-89 (v <= MAX), partitioned_s_tree.rs:112 (is_sorted) -------
__global__ void check_keys_kernel(const uint32_t* __restrict__ v, size_t n, unsigned* __restrict__ flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned bad = 0;
    for (; i < n; i += stride) {
        uint32_t x = v[i];
        if (x > kMax) bad |= 1u;
        if (i + 1 < n && v[i + 1] < x) bad |= 2u;
    }
    if (bad) atomicOr(flags, bad);
}

struct PlainShape {
    size_t n;
    unsigned B;
    int levels;
    int left_max;
    unsigned long long level_slot[kMaxLevels];  // first slot of each level in the image
    unsigned long long level_slots[kMaxLevels]; // slots in each level (16 * nodes)
    unsigned long long stride_pow[kMaxLevels];  // (B+1)^(H-2-h) for internal levels
};

// Leaf layer: s_tree.rs:132-145.  Slot s of leaf node j.
__global__ void plain_leaf_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ tree, PlainShape sh,
                                  int zero_tail) {
    const unsigned long long slots = sh.level_slots[sh.levels - 1];
    const unsigned long long base = sh.level_slot[sh.levels - 1];
    const size_t n = sh.n;
    const unsigned B = sh.B;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < slots;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long j = t >> 4;
        const unsigned s = (unsigned)(t & 15);
        const unsigned long long i = s < B ? j * B + s : (j + 1) * B;  // slot B of a B<16 node: next leaf's first key (:137-139)
        uint32_t v;
        if (i < n) v = vals[i];
        else if (j == n / B && s >= n % B) v = kMax;  // :142-145
        else if (s >= B && n % B == 0 && j + 1 == n / B) v = kMax;  // deviation: reference leaves a stray 0 here
        else v = zero_tail ? 0u : kMax;  // other nodes stay zero (hugepages)
        tree[base + t] = v;
    }
}

// Internal layers: s_tree.rs:149-173, closed form per slot.
__global__ void plain_inner_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ tree, PlainShape sh,
                                   unsigned long long total_inner_slots) {
    const unsigned B = sh.B;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total_inner_slots;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        // find the level this slot belongs to (levels 0 .. H-2 are enumerated root first)
        unsigned long long r = t;
        int h = 0;
        while (r >= sh.level_slots[h]) { r -= sh.level_slots[h]; h++; }
        const unsigned long long node = r >> 4;
        const unsigned j = (unsigned)(r & 15);
        uint32_t v = kMax;
        if (j < B) {
            const unsigned long long k = (node * (B + 1) + j + 1) * sh.stride_pow[h];
            if (k * B < sh.n) v = sh.left_max ? vals[k * B - 1] : vals[k * B];
        }
        tree[sh.level_slot[h] + r] = v;
    }
}

__global__ void fill_kernel(uint32_t* __restrict__ p, unsigned long long count, uint32_t value) {
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < count;
         t += (unsigned long long)gridDim.x * blockDim.x)
        p[t] = value;
}


// ---- shared-memory rank table for the top of the plain tree (see stree_search.cu) --------------
// Separator before node c (c >= 1) of level t: the key the builder formula of s_tree.rs:156-172
// places between the subtrees of nodes c-1 and c, i.e. vals[f*16 - 1] (left_max) or vals[f*16]
// with f = c * 17^(H-1-t) the first leaf below node c.
__global__ void top_bounds_kernel(const uint32_t* __restrict__ vals, unsigned nbound, unsigned long long leaf_stride, int left_max,
                                  uint32_t* __restrict__ bkeys, uint16_t* __restrict__ low) {
    for (unsigned c = blockIdx.x * blockDim.x + threadIdx.x; c < nbound; c += gridDim.x * blockDim.x) {
        const unsigned long long i = (unsigned long long)(c + 1) * leaf_stride * 16ull;
        const uint32_t key = left_max ? vals[i - 1] : vals[i];
        bkeys[c] = key;
        low[c] = (uint16_t)(key & 0xffffu);
    }
}
// table[b] = number of separators whose key >> 16 is < b, b in [0, 2^15]
__global__ void top_table_kernel(const uint32_t* __restrict__ bkeys, unsigned nbound, uint16_t* __restrict__ table) {
    for (unsigned b = blockIdx.x * blockDim.x + threadIdx.x; b <= (1u << 15); b += gridDim.x * blockDim.x) {
        unsigned lo = 0, hi = nbound;
        while (lo < hi) {
            const unsigned m = (lo + hi) >> 1;
            if ((bkeys[m] >> 16) < b) lo = m + 1; else hi = m;
        }
        table[b] = (uint16_t)lo;
    }
}

bool validate_keys(const uint32_t* d_sorted, size_t n, int device) {
    unsigned* d_flags = nullptr;
    if (!SST_CUDA_OK(cudaMalloc(&d_flags, sizeof(unsigned)))) return false;
    cudaStream_t st = thread_stream(device);
    bool ok = SST_CUDA_OK(cudaMemsetAsync(d_flags, 0, sizeof(unsigned), st));
    unsigned flags = 0;
    if (ok) {
        check_keys_kernel<<<std::min<unsigned>(grid_for(n), cur_sms() * 16), kBuildThreads, 0, st>>>(d_sorted, n, d_flags);
        ok = SST_CUDA_OK(cudaGetLastError()) &&
             SST_CUDA_OK(cudaMemcpyAsync(&flags, d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, st)) &&
             SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(d_flags);
    if (!ok) return false;
    if (flags & 1u) { set_error(SST_ERR_ARG, "key larger than i32::MAX (reference: assert!(v <= MAX), s_tree.rs:87-89)"); return false; }
    if (flags & 2u) { set_error(SST_ERR_ARG, "keys are not sorted"); return false; }
    return true;
}

// Optional persisting-L2 carve-out (SST_PERSIST != 0): lets a per-launch access-policy window pin
// the last internal level.  Device-wide limit, so only touched when asked for.
void configure_persisting_l2(sst_index* idx) {
    if (opt(OPT_PERSIST) == 0) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, idx->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, idx->device);
    if (max_persist <= 0 || max_window <= 0) return;
    size_t want = (size_t)max_persist;
    if (const long long mb = opt(OPT_PERSIST_MB); mb >= 0) want = std::min<size_t>((size_t)mb << 20, (size_t)max_persist);
    const cudaError_t rc = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    if (opt(OPT_DEBUG)) {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
        fprintf(stderr, "[sst] persisting L2: max=%d window_max=%d set rc=%d limit now=%zu\n", max_persist, max_window, (int)rc, got);
    }
    if (rc != cudaSuccess) { (void)cudaGetLastError(); return; }
    idx->persist_ok = true;
    idx->persist_window_max = (size_t)max_window;
}

}  // namespace

// =================================================================================================
// Plain S+-tree
// =================================================================================================
sst_index* build_plain(const uint32_t* d_sorted, bool, size_t n, uint32_t node_b, uint32_t flags, int device) {
    clear_error();
    if (n == 0) { set_error(SST_ERR_ARG, "empty input (reference: assert!(n > 0), s_tree.rs:93)"); return nullptr; }
    if (node_b != 16 && node_b != 15) { set_error(SST_ERR_UNSUPPORTED, "node_b must be 16 (STree16) or 15 (STree15)"); return nullptr; }
    if (flags & ~(uint32_t)(SST_LEFT_MAX | SST_REVERSE_STORAGE | SST_FULL_ARRAY)) { set_error(SST_ERR_ARG, "unknown flags"); return nullptr; }
    if ((flags & SST_FULL_ARRAY) && (flags & SST_REVERSE_STORAGE)) {
        set_error(SST_ERR_ARG, "Full array only makes sense in forward layout (s_tree.rs:77-82)");
        return nullptr;
    }
    if (n >= ((size_t)1 << 32)) { set_error(SST_ERR_UNSUPPORTED, "n must be < 2^32"); return nullptr; }
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard guard(device);
    if (!guard.ok) return nullptr;
    configure_l2_fetch(device);
    if (!validate_keys(d_sorted, n, device)) return nullptr;

    const size_t B = node_b;
    const bool full = flags & SST_FULL_ARRAY, reverse = flags & SST_REVERSE_STORAGE;
    const size_t H = tb_height(n, B);
    if (H > (size_t)kMaxLevels) { set_error(SST_ERR_UNSUPPORTED, "tree too high"); return nullptr; }
    auto* idx = new sst_index();
    idx->device = device; idx->variant = SST_PLAIN; idx->node_b = node_b; idx->flags = flags; idx->n = n;
    idx->levels = (int)H;
    size_t n_blocks = 0;
    for (size_t h = 0; h < H; h++) {  // s_tree.rs:96-104
        idx->layer_sizes[h] = full ? ipow(B + 1, h) : div_ceil(tb_layer_size(n, h, H, B), B);
        idx->layer_blocks[h] = idx->layer_sizes[h];
        n_blocks += idx->layer_sizes[h];
    }
    if (n_blocks * 64 > ((size_t)64 << 30)) {  // vec_on_hugepages cap, util.rs:135
        set_error(SST_ERR_CAPACITY, "tree larger than 64 GiB (util.rs:135)");
        delete idx;
        return nullptr;
    }
    size_t sum = 0;
    for (size_t h = 0; h < H; h++) {  // s_tree.rs:106-123
        if (!reverse) { idx->offsets[h] = sum; sum += idx->layer_sizes[h]; }
        else { sum += idx->layer_sizes[h]; idx->offsets[h] = n_blocks - sum; }
    }
    idx->n_blocks = n_blocks;
    if (!SST_CUDA_OK(cudaMalloc(&idx->d_tree, (n_blocks + 1) * 64))) { delete idx; return nullptr; }

    PlainShape sh{};
    sh.n = n; sh.B = (unsigned)B; sh.levels = (int)H; sh.left_max = (flags & SST_LEFT_MAX) ? 1 : 0;
    unsigned long long inner_slots = 0;
    for (size_t h = 0; h < H; h++) {
        sh.level_slot[h] = (unsigned long long)idx->offsets[h] * 16;
        sh.level_slots[h] = (unsigned long long)idx->layer_sizes[h] * 16;
        sh.stride_pow[h] = h + 2 <= H ? ipow(B + 1, H - 2 - h) : 0;
        if (h + 1 < H) inner_slots += sh.level_slots[h];
    }
    cudaStream_t st = thread_stream(device);
    plain_leaf_kernel<<<std::min<unsigned>(grid_for(sh.level_slots[H - 1]), cur_sms() * 32), kBuildThreads, 0, st>>>(
        d_sorted, idx->d_tree, sh, 1);
    if (inner_slots)
        plain_inner_kernel<<<std::min<unsigned>(grid_for(inner_slots), cur_sms() * 32), kBuildThreads, 0, st>>>(
            d_sorted, idx->d_tree, sh, inner_slots);
    fill_kernel<<<1, 16, 0, st>>>(idx->d_tree + n_blocks * 16, 16, kMax);  // guard node
    if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaStreamSynchronize(st))) {
        cudaFree(idx->d_tree);
        delete idx;
        return nullptr;
    }
    if (!build_top_table(idx, d_sorted) || !build_compressed_level(idx) || !build_bucket_aux(idx)) {
        free_index(idx);
        return nullptr;
    }
    configure_persisting_l2(idx);
    finalize_view(idx);
    return idx;
}

// =================================================================================================
// Prefix-partitioned S+-trees
// =================================================================================================
namespace {

// part_start[p] = index of the first key with (key >> shift) >= p, for p in [0, parts]; the
// bucket histogram of partitioned_s_tree.rs:123-127 is its adjacent difference.
__global__ void part_start_kernel(const uint32_t* __restrict__ vals, size_t n, unsigned shift, unsigned long long parts,
                                  uint32_t* __restrict__ part_start) {
    for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p <= parts;
         p += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long key = p << shift;  // may exceed 32 bits for p == parts
        size_t l = 0, r = n;
        while (l < r) {
            size_t m = (l + r) >> 1;
            if ((unsigned long long)vals[m] < key) l = m + 1; else r = m;
        }
        part_start[p] = (uint32_t)l;
    }
}

struct PartShape {
    size_t n;
    int levels;
    int variant;
    int has_overlap;
    unsigned long long parts, extra_parts, max_bucket, part_size, subtree, bpp;
    unsigned long long l1, overlap;
    unsigned long long level_slot[kMaxLevels];   // first slot of each level (within a part for COMPACT)
    unsigned long long ls[kMaxLevels];           // nodes per level per part (global for MAP)
    unsigned long long blocks[kMaxLevels];       // nodes allocated per level (non-compact)
    unsigned long long stride_pow[kMaxLevels];   // 17^(H-2-h)
    unsigned long long n_ne;                     // number of non-empty parts
    unsigned long long last_ne_part;
};

// Leaf layer of Simple / L1 / Overlapping (partitioned_s_tree.rs:501-527): key i of part p goes to
// slot part_pos[p] + (i - part_start[p]); gaps take the first key of the next non-empty part.
__global__ void ps_leaf_full_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ tree, PartShape sh,
                                    const unsigned long long* __restrict__ ne_pos, const uint32_t* __restrict__ ne_start,
                                    const uint32_t* __restrict__ ne_cnt) {
    const unsigned long long slots = sh.blocks[sh.levels - 1] * 16, base = sh.level_slot[sh.levels - 1];
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < slots;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        // last non-empty part a with ne_pos[a] <= s
        unsigned long long lo = 0, hi = sh.n_ne;  // first a with ne_pos[a] > s
        while (lo < hi) {
            unsigned long long m = (lo + hi) >> 1;
            if (ne_pos[m] <= s) lo = m + 1; else hi = m;
        }
        uint32_t v;
        if (lo == 0) v = vals[0];  // before the first key: filled with the first key (:506-514)
        else {
            const unsigned long long a = lo - 1, off = s - ne_pos[a];
            if (off < ne_cnt[a]) v = vals[ne_start[a] + off];
            else if (lo < sh.n_ne) v = vals[ne_start[lo]];
            else v = kMax;
        }
        tree[base + s] = v;
    }
}

// Leaf layer of Map: the sorted array verbatim (partitioned_s_tree.rs:501-527 with MAP).
__global__ void ps_leaf_map_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ tree, PartShape sh) {
    const unsigned long long slots = sh.ls[sh.levels - 1] * 16, base = sh.level_slot[sh.levels - 1];
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < slots;
         s += (unsigned long long)gridDim.x * blockDim.x)
        tree[base + s] = s < sh.n ? vals[s] : kMax;
}

// Leaf layer of Compact (partitioned_s_tree.rs:283-307).
__global__ void ps_leaf_compact_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ tree, PartShape sh,
                                       const uint32_t* __restrict__ part_start) {
    const unsigned long long per_part = sh.ls[sh.levels - 1] * 16, total = per_part * sh.parts;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = t / per_part, j = t % per_part;
        const unsigned long long st = part_start[p], cnt = part_start[p + 1] - st;
        uint32_t v = kMax;
        if (j < cnt) v = vals[st + j];
        else if ((j >> 4) == (cnt >> 4) && p < sh.last_ne_part) v = vals[st + cnt];  // first key of the next non-empty part
        tree[p * sh.bpp * 16 + sh.level_slot[sh.levels - 1] + j] = v;
    }
}

// Inner layers of the partitioned layouts: partitioned_s_tree.rs:310-329 (Compact), :569-585.
// Reads the leaf layer written before.
__global__ void ps_inner_kernel(uint32_t* __restrict__ tree, PartShape sh, int h_first, unsigned long long slots_per_part) {
    const unsigned long long np = sh.variant == SST_COMPACT ? sh.parts : sh.parts + sh.extra_parts;
    const unsigned long long total = slots_per_part * np;
    const int H = sh.levels;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = t / slots_per_part;
        unsigned long long r = t % slots_per_part;
        int h = h_first;
        while (r >= sh.ls[h] * 16) { r -= sh.ls[h] * 16; h++; }
        const unsigned long long node = r >> 4, j = r & 15;
        const unsigned long long k = (node * 17 + j + 1) * sh.stride_pow[h];
        uint32_t v = kMax;
        if (sh.variant == SST_COMPACT) {
            const unsigned long long pb = p * sh.bpp * 16;
            if (k * 16 < sh.max_bucket) v = tree[pb + sh.level_slot[H - 1] + (k - 1) * 16 + 15];
            tree[pb + sh.level_slot[h] + r] = v;
        } else {
            if (k * 16 < sh.max_bucket) v = tree[sh.level_slot[H - 1] + (sh.ls[H - 1] * p + k - 1) * 16 + 15];
            tree[sh.level_slot[h] + sh.ls[h] * p * 16 + r] = v;
        }
    }
}

// Inner layers h >= 1 of Map: the ordinary left-max S+-tree over all n keys (:556-568).
__global__ void ps_inner_map_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ tree, PartShape sh,
                                    unsigned long long total) {
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long r = t;
        int h = 1;
        while (r >= sh.ls[h] * 16) { r -= sh.ls[h] * 16; h++; }
        const unsigned long long node = r >> 4, j = r & 15;
        const unsigned long long k = (node * 17 + j + 1) * sh.stride_pow[h];
        tree[sh.level_slot[h] + r] = (k * 16 < sh.n) ? vals[k * 16 - 1] : kMax;
    }
}

// Level 0 with overlap (Overlapping with Some(o), Map): flat array of subtree maxima (:534-552).
__global__ void ps_level0_kernel(uint32_t* __restrict__ tree, PartShape sh, unsigned long long slots, unsigned long long range) {
    const unsigned long long leaf = sh.level_slot[sh.levels - 1];
    const unsigned long long leaf_slots = (sh.variant == SST_MAP ? sh.ls[sh.levels - 1] : sh.blocks[sh.levels - 1]) * 16;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < slots;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t v = kMax;
        if (i < range) {
            const unsigned long long j = (i + 1) * sh.subtree - 1;
            if (j < leaf_slots) v = tree[leaf + j];
        }
        tree[sh.level_slot[0] + i] = v;
    }
}

// prefix_map of Map (:599-617): first level-0 slot whose key has prefix >= p, clamped.
__global__ void ps_prefix_map_kernel(const uint32_t* __restrict__ tree, PartShape sh, unsigned shift, uint32_t* __restrict__ pm) {
    const unsigned long long slots = sh.ls[0] * 16, max_idx = slots - 16;
    const uint32_t* l0 = tree + sh.level_slot[0];
    for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < sh.parts;
         p += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long res = 0;
        if (p > 0) {
            unsigned long long lo = 0, hi = slots;
            while (lo < hi) {
                unsigned long long m = (lo + hi) >> 1;
                if (((unsigned long long)l0[m] >> shift) < p) lo = m + 1; else hi = m;
            }
            res = lo < max_idx ? lo : max_idx;
        }
        pm[p] = (uint32_t)res;
    }
}

// COMPACT: gather the first `upper_nodes` nodes (all non-leaf levels) of every part into a dense array.
__global__ void compact_upper_kernel(const uint32_t* __restrict__ tree, unsigned long long parts, unsigned long long bpp,
                                     unsigned long long upper_nodes, uint32_t* __restrict__ upper) {
    const unsigned long long total = parts * upper_nodes * 4;  // in 16-byte chunks
    const uint4* src = reinterpret_cast<const uint4*>(tree);
    uint4* dst = reinterpret_cast<uint4*>(upper);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long part = t / (upper_nodes * 4), r = t % (upper_nodes * 4);
        dst[t] = src[part * bpp * 4 + r];
    }
}

// partitioned_s_tree.rs:200-227
bool max_overlap(const std::vector<size_t>& buckets, size_t subtree_size, size_t* out) {
    if (buckets.size() == 1) {
        if (buckets[0] <= subtree_size) { *out = 0; return true; }
        return false;
    }
    const size_t capacity = 16 * subtree_size;
    for (int overlap = 15; overlap >= 0; overlap--) {
        size_t x = 0;
        bool ok = true;
        for (size_t b : buckets) {
            x += b;
            if (x > capacity) { ok = false; break; }
            const size_t sub = (16 - (size_t)overlap) * subtree_size;
            x = x > sub ? x - sub : 0;
        }
        if (ok) { *out = (size_t)overlap; return true; }
    }
    return false;
}

template <class T>
bool upload(T** d, const std::vector<T>& h, cudaStream_t st) {
    *d = nullptr;
    if (!SST_CUDA_OK(cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(T)))) return false;
    if (h.empty()) return true;
    return SST_CUDA_OK(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

void free_index(sst_index* idx) {
    if (!idx) return;
    cudaFree(idx->d_tree);
    cudaFree(idx->d_prefix_map);
    cudaFree(idx->d_part_start);
    cudaFree(idx->d_part_pos);
    cudaFree(idx->d_upper);
    cudaFree(idx->d_top_table);
    cudaFree(idx->d_top_low);
    cudaFree(idx->d_c5);
    cudaFree(idx->d_h5);
    free_bucket_aux(idx);
    delete idx;
}

}  // namespace

// 16-bit copy of the last internal level: node -> (base = first separator, 16 deltas).  Halves the L2
// footprint of the one level that competes with leaf traffic for the cache (63 -> 33+4 MB at 2^28 keys).
// A node whose separators span 2^16 or more (sparse keys, MAX padding) is flagged and read exactly.
__global__ void compress_level_kernel(const uint32_t* __restrict__ level, unsigned long long nodes, uint16_t* __restrict__ low,
                                      uint32_t* __restrict__ high) {
    for (unsigned long long nd = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; nd < nodes;
         nd += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(level + nd * 16);
        uint32_t k[16];
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint4 v = src[i]; k[4 * i] = v.x; k[4 * i + 1] = v.y; k[4 * i + 2] = v.z; k[4 * i + 3] = v.w; }
        const uint32_t h = k[0];
        bool uniform = h != 0xffffffffu;
#pragma unroll
        for (int i = 1; i < 16; i++) uniform = uniform && k[i] >= h && (k[i] - h) < 65536u;
        high[nd] = uniform ? h : 0xffffffffu;
        uint4* dst = reinterpret_cast<uint4*>(low + nd * 16);
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = uniform ? (((k[2 * i] - h) & 0xffffu) | ((k[2 * i + 1] - h) << 16)) : 0u;
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

bool build_compressed_level(sst_index* idx) {
    if (idx->variant != SST_PLAIN || idx->node_b != 16 || idx->levels < 3) return true;
    // Opt-in experiment (SST_C5=1): measured at 2^28 keys it cuts DRAM reads by 15 % (7.99 -> 6.77 GB per
    // 10^8 queries, L2 hit 47 -> 54 %) but the extra load and instructions cost 8 % of run time.
    if (!opt(OPT_C5)) return true;
    const int h = idx->levels - 2;
    const size_t nodes = idx->layer_sizes[h];
    if (idx->top_level > h) return true;
    cudaStream_t st = thread_stream(idx->device);
    bool ok = SST_CUDA_OK(cudaMalloc(&idx->d_c5, nodes * 32)) && SST_CUDA_OK(cudaMalloc(&idx->d_h5, nodes * 4));
    if (ok) {
        compress_level_kernel<<<std::min<unsigned>(grid_for(nodes), cur_sms() * 16), kBuildThreads, 0, st>>>(
            idx->d_tree + idx->offsets[h] * 16, nodes, idx->d_c5, idx->d_h5);
        ok = SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (!ok) { cudaFree(idx->d_c5); cudaFree(idx->d_h5); idx->d_c5 = nullptr; idx->d_h5 = nullptr; }
    return ok;
}

// Chooses the deepest level t whose node count fits 16-bit ranks and shared memory, and builds
// the rank table for it.  Plain B=16 trees only; a tree of height 1 has no table.
bool build_top_table(sst_index* idx, const uint32_t* d_sorted) {
    if (idx->variant != SST_PLAIN || idx->node_b != 16 || idx->levels < 2) return true;
    const size_t H = idx->levels, n = idx->n;
    const size_t smem_cap = max_smem_optin(idx->device);
    const size_t table_bytes = (((size_t)(1u << 15) + 1) * 2 + 15) & ~(size_t)15;
    int t = 0;
    size_t nb = 0;
    for (size_t lvl = 1; lvl < H; lvl++) {
        const size_t real_nodes = div_ceil(tb_layer_size(n, lvl, H, 16), 16);  // full_array has more slots, not more nodes
        if (real_nodes < 2) continue;
        const size_t bounds = real_nodes - 1;
        if (bounds > 65535) break;
        if (table_bytes + ((bounds * 2 + 15) & ~(size_t)15) + 4096 > smem_cap) break;
        t = (int)lvl;
        nb = bounds;
    }
    if (t == 0) return true;
    cudaStream_t st = thread_stream(idx->device);
    uint32_t* d_bkeys = nullptr;
    const size_t low_bytes = (nb * 2 + 15) & ~(size_t)15;
    bool ok = SST_CUDA_OK(cudaMalloc(&d_bkeys, nb * 4)) && SST_CUDA_OK(cudaMalloc(&idx->d_top_table, table_bytes)) &&
              SST_CUDA_OK(cudaMalloc(&idx->d_top_low, low_bytes)) && SST_CUDA_OK(cudaMemsetAsync(idx->d_top_table, 0, table_bytes, st)) &&
              SST_CUDA_OK(cudaMemsetAsync(idx->d_top_low, 0, low_bytes, st));
    if (ok) {
        top_bounds_kernel<<<std::min<unsigned>(grid_for(nb), cur_sms() * 8), kBuildThreads, 0, st>>>(
            d_sorted, (unsigned)nb, ipow(17, H - 1 - t), (idx->flags & SST_LEFT_MAX) ? 1 : 0, d_bkeys, idx->d_top_low);
        top_table_kernel<<<std::min<unsigned>(grid_for((1u << 15) + 1), cur_sms() * 8), kBuildThreads, 0, st>>>(d_bkeys, (unsigned)nb, idx->d_top_table);
        ok = SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(d_bkeys);
    if (!ok) {
        cudaFree(idx->d_top_table); cudaFree(idx->d_top_low);
        idx->d_top_table = nullptr; idx->d_top_low = nullptr;
        return false;
    }
    idx->top_level = t;
    idx->top_nbound = nb;
    return true;
}

sst_index* build_partitioned(const uint32_t* d_sorted, size_t n, uint32_t b, int variant, int device) {
    clear_error();
    if (n == 0) { set_error(SST_ERR_ARG, "empty input"); return nullptr; }
    if (variant < SST_SIMPLE || variant > SST_MAP) { set_error(SST_ERR_ARG, "unknown partitioned layout"); return nullptr; }
    if (n >= ((size_t)1 << 32)) { set_error(SST_ERR_UNSUPPORTED, "n must be < 2^32"); return nullptr; }
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard guard(device);
    if (!guard.ok) return nullptr;
    configure_l2_fetch(device);
    if (!validate_keys(d_sorted, n, device)) return nullptr;
    cudaStream_t st = thread_stream(device);
    const size_t B = 16;
    const bool COMPACT = variant == SST_COMPACT, MAPV = variant == SST_MAP;
    const bool OL = variant == SST_OVERLAPPING || MAPV, L1V = variant == SST_L1 || OL;

    // ---- get_part_size (partitioned_s_tree.rs:111-190) ----
    uint32_t last = 0;
    if (!SST_CUDA_OK(cudaMemcpy(&last, d_sorted + (n - 1), 4, cudaMemcpyDeviceToHost))) return nullptr;
    if (last == 0) { set_error(SST_ERR_ARG, "largest key is 0 (reference: ilog2(0) panics, partitioned_s_tree.rs:116)"); return nullptr; }
    const size_t bits = 1 + (63 - __builtin_clzll((unsigned long long)last));
    size_t shift = bits > b ? bits - b : 0;
    size_t parts = (size_t)1 << (bits - shift);
    if (parts > ((size_t)1 << 27)) { set_error(SST_ERR_UNSUPPORTED, "more than 2^27 parts"); return nullptr; }
    std::vector<uint32_t> fine(parts + 1);  // part_start at the requested granularity
    {
        uint32_t* d_ps = nullptr;
        if (!SST_CUDA_OK(cudaMalloc(&d_ps, (parts + 1) * 4))) return nullptr;
        part_start_kernel<<<std::min<unsigned>(grid_for(parts + 1), cur_sms() * 32), kBuildThreads, 0, st>>>(d_sorted, n, (unsigned)shift, parts, d_ps);
        bool ok = SST_CUDA_OK(cudaGetLastError()) &&
                  SST_CUDA_OK(cudaMemcpyAsync(fine.data(), d_ps, (parts + 1) * 4, cudaMemcpyDeviceToHost, st)) &&
                  SST_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(d_ps);
        if (!ok) return nullptr;
    }
    const size_t shift_fine = shift;
    auto buckets_at = [&](size_t sh2, size_t parts2) {
        std::vector<size_t> bs(parts2);
        const size_t step = (size_t)1 << (sh2 - shift_fine);
        for (size_t p = 0; p < parts2; p++) bs[p] = (size_t)fine[(p + 1) * step] - fine[p * step] + (COMPACT ? 1 : 0);
        return bs;
    };
    auto get_height = [&](size_t x) { return tb_height(MAPV ? div_ceil(x * 17, 16) : x, B); };
    std::vector<size_t> bucket_sizes = buckets_at(shift, parts);
    size_t max_bucket = *std::max_element(bucket_sizes.begin(), bucket_sizes.end());
    size_t height = get_height(max_bucket);
    for (size_t b2 = b;;) {  // :139-168
        if (b2 == 0) break;
        b2 -= 1;
        if (b2 > bits) break;
        const size_t shift2 = bits > b2 ? bits - b2 : 0, parts2 = (size_t)1 << (bits - shift2);
        std::vector<size_t> bs2 = buckets_at(shift2, parts2);
        const size_t mb2 = *std::max_element(bs2.begin(), bs2.end());
        const size_t h2 = get_height(mb2);
        if (h2 > height) break;
        shift = shift2; parts = parts2; max_bucket = mb2; bucket_sizes.swap(bs2); height = h2;
    }
    if (height > (size_t)kMaxLevels) { set_error(SST_ERR_UNSUPPORTED, "tree too high"); return nullptr; }
    bool has_overlap = false;
    size_t overlap = 0;
    const size_t subtree = height == 1 ? 1 : B * ipow(B + 1, height - 2);
    if (MAPV) { has_overlap = true; overlap = 0; }
    else if (OL) has_overlap = max_overlap(bucket_sizes, subtree, &overlap);

    // part_start at the final granularity
    std::vector<uint32_t> ps(parts + 1);
    {
        const size_t step = (size_t)1 << (shift - shift_fine);
        for (size_t p = 0; p <= parts; p++) ps[p] = fine[p * step];
    }
    fine.clear();
    fine.shrink_to_fit();

    auto* idx = new sst_index();
    idx->device = device; idx->variant = variant; idx->node_b = 16; idx->n = n; idx->levels = (int)height;
    idx->shift = shift; idx->parts = parts; idx->max_bucket = max_bucket; idx->has_overlap = has_overlap;
    idx->overlap = has_overlap ? overlap : 0;

    PartShape sh{};
    sh.n = n; sh.levels = (int)height; sh.variant = variant; sh.has_overlap = has_overlap;
    sh.parts = parts; sh.max_bucket = max_bucket; sh.subtree = subtree; sh.overlap = overlap;
    size_t l1 = 0, n_blocks = 0, extra_parts = 0;
    std::vector<size_t> ls(height);
    if (COMPACT) {  // :254-269
        for (size_t h = 0; h < height; h++) ls[h] = div_ceil(tb_layer_size(max_bucket, h, height, B), B);
        size_t bpp = 0;
        for (size_t h = 0; h < height; h++) { idx->offsets[h] = bpp; bpp += ls[h]; }
        idx->bpp = bpp; sh.bpp = bpp;
        n_blocks = parts * bpp;
        for (size_t h = 0; h < height; h++) idx->layer_blocks[h] = ls[h] * parts;
    } else {
        if (MAPV) {  // :376-386
            for (size_t h = 0; h < height; h++) ls[h] = div_ceil(tb_layer_size(n, h, height, B), B);
            if (height > 1) ls[0] = div_ceil(div_ceil(tb_layer_size(n, 1, height, B), B), B);
        } else if (!L1V) {  // :387-388
            for (size_t h = 0; h < height; h++) ls[h] = ipow(B + 1, h);
        } else {  // :399-409
            if (OL) l1 = has_overlap ? 16 - overlap : 17;
            else l1 = div_ceil(tb_layer_size(max_bucket, 1, height, B), B);
            for (size_t h = 0; h < height; h++) ls[h] = div_ceil(ipow(B + 1, h) * l1, B + 1);
        }
        if (!MAPV) {  // :414-448
            if (ls[0] != 1) { set_error(SST_ERR_ARG, "unexpected root size (partitioned_s_tree.rs:415-419)"); delete idx; return nullptr; }
            extra_parts = l1 == 0 ? 0 : div_ceil(has_overlap ? overlap : 0, l1);
            for (size_t h = 0; h < height; h++) idx->layer_blocks[h] = ls[h] * (parts + extra_parts);
            if (has_overlap) idx->layer_blocks[0] = div_ceil(parts * (16 - overlap) + overlap, 16);
        } else {
            for (size_t h = 0; h < height; h++) idx->layer_blocks[h] = ls[h];
        }
        for (size_t h = 0; h < height; h++) { idx->offsets[h] = n_blocks; n_blocks += idx->layer_blocks[h]; }
    }
    for (size_t h = 0; h < height; h++) idx->layer_sizes[h] = ls[h];
    if (n_blocks * 64 > ((size_t)32 << 30)) {  // :271-274, :463-466
        set_error(SST_ERR_CAPACITY, "partitioned tree larger than 32 GiB (reference returns None)");
        delete idx;
        return nullptr;
    }
    if (MAPV && 24 + parts * 4 > 4 * n * 4) {  // :594-597 (size_of_val(&tree) is the 24-byte Vec header)
        set_error(SST_ERR_CAPACITY, "prefix map larger than 4x the input (reference returns None)");
        delete idx;
        return nullptr;
    }
    idx->n_blocks = n_blocks;
    idx->l1 = l1;
    idx->l1_field = OL ? std::max<size_t>(l1, 16) : l1;  // :636-643
    sh.l1 = l1; sh.extra_parts = extra_parts;
    size_t part_size = l1 * subtree;                      // :487
    if (!OL) part_size = B * ls[height - 1];              // :490-493
    sh.part_size = part_size;
    for (size_t h = 0; h < height; h++) {
        sh.level_slot[h] = (unsigned long long)idx->offsets[h] * 16;
        sh.ls[h] = ls[h];
        sh.blocks[h] = idx->layer_blocks[h];
        sh.stride_pow[h] = h + 2 <= height ? ipow(B + 1, height - 2 - h) : 0;
    }

    // ---- position tables (host, O(parts)) ----
    std::vector<unsigned long long> part_pos(parts, 0);
    std::vector<unsigned long long> ne_pos;
    std::vector<uint32_t> ne_start, ne_cnt;
    size_t last_ne = 0;
    {
        unsigned long long run = 0;
        for (size_t p = 0; p < parts; p++) {
            const size_t cnt = ps[p + 1] - ps[p];
            if (COMPACT || MAPV) {
                part_pos[p] = MAPV ? ps[p] : 0;
                if (cnt) last_ne = p;
                continue;
            }
            if (cnt == 0) { part_pos[p] = run; continue; }
            run = std::max<unsigned long long>(run, (unsigned long long)p * part_size);  // :506-514
            part_pos[p] = run;
            ne_pos.push_back(run); ne_start.push_back(ps[p]); ne_cnt.push_back((uint32_t)cnt);
            run += cnt;
            last_ne = p;
        }
        if (!COMPACT && !MAPV && run > (unsigned long long)idx->layer_blocks[height - 1] * 16) {
            set_error(SST_ERR_ARG, "internal: leaf layer overflow");
            delete idx;
            return nullptr;
        }
    }
    sh.n_ne = ne_pos.size();
    sh.last_ne_part = last_ne;

    bool ok = SST_CUDA_OK(cudaMalloc(&idx->d_tree, (n_blocks + 1) * 64));
    unsigned long long* d_ne_pos = nullptr;
    uint32_t *d_ne_start = nullptr, *d_ne_cnt = nullptr;
    ok = ok && upload(&idx->d_part_start, ps, st) && upload(&idx->d_part_pos, part_pos, st);
    const unsigned maxgrid = cur_sms() * 32;
    const int H = (int)height;
    if (ok) {
        if (COMPACT) {
            ps_leaf_compact_kernel<<<std::min(grid_for(ls[H - 1] * 16 * parts), maxgrid), kBuildThreads, 0, st>>>(d_sorted, idx->d_tree, sh, idx->d_part_start);
            size_t inner = 0;
            for (int h = 0; h + 1 < H; h++) inner += ls[h] * 16;
            if (inner) ps_inner_kernel<<<std::min(grid_for(inner * parts), maxgrid), kBuildThreads, 0, st>>>(idx->d_tree, sh, 0, inner);
        } else if (MAPV) {
            ps_leaf_map_kernel<<<std::min(grid_for(ls[H - 1] * 16), maxgrid), kBuildThreads, 0, st>>>(d_sorted, idx->d_tree, sh);
            size_t inner = 0;
            for (int h = 1; h + 1 < H; h++) inner += ls[h] * 16;
            if (inner) ps_inner_map_kernel<<<std::min(grid_for(inner), maxgrid), kBuildThreads, 0, st>>>(d_sorted, idx->d_tree, sh, inner);
            if (H > 1) ps_level0_kernel<<<std::min(grid_for(ls[0] * 16), maxgrid), kBuildThreads, 0, st>>>(idx->d_tree, sh, ls[0] * 16, ls[1] - 1);
            idx->prefix_map_len = parts;
            ok = SST_CUDA_OK(cudaMalloc(&idx->d_prefix_map, parts * 4));
            if (ok) ps_prefix_map_kernel<<<std::min(grid_for(parts), maxgrid), kBuildThreads, 0, st>>>(idx->d_tree, sh, (unsigned)shift, idx->d_prefix_map);
        } else {
            ok = upload(&d_ne_pos, ne_pos, st) && upload(&d_ne_start, ne_start, st) && upload(&d_ne_cnt, ne_cnt, st);
            if (ok) {
                ps_leaf_full_kernel<<<std::min(grid_for(idx->layer_blocks[H - 1] * 16), maxgrid), kBuildThreads, 0, st>>>(d_sorted, idx->d_tree, sh, d_ne_pos, d_ne_start, d_ne_cnt);
                const int h_first = has_overlap ? 1 : 0;  // with overlap, level 0 is the flat array (:534-552)
                size_t inner = 0;
                for (int h = h_first; h + 1 < H; h++) inner += ls[h] * 16;
                if (inner) ps_inner_kernel<<<std::min(grid_for(inner * (parts + extra_parts)), maxgrid), kBuildThreads, 0, st>>>(idx->d_tree, sh, h_first, inner);
                if (has_overlap && H > 1)
                    ps_level0_kernel<<<std::min(grid_for(idx->layer_blocks[0] * 16), maxgrid), kBuildThreads, 0, st>>>(idx->d_tree, sh, idx->layer_blocks[0] * 16, parts * l1 + overlap);
            }
        }
        if (ok) fill_kernel<<<1, 16, 0, st>>>(idx->d_tree + n_blocks * 16, 16, kMax);
        if (ok && COMPACT && H > 1) {
            const size_t upper_nodes = idx->offsets[H - 1];
            ok = SST_CUDA_OK(cudaMalloc(&idx->d_upper, parts * upper_nodes * 64));
            if (ok) compact_upper_kernel<<<std::min(grid_for(parts * upper_nodes * 4), maxgrid), kBuildThreads, 0, st>>>(
                        idx->d_tree, parts, idx->bpp, upper_nodes, idx->d_upper);
        }
        ok = ok && SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(d_ne_pos); cudaFree(d_ne_start); cudaFree(d_ne_cnt);
    if (ok) ok = build_bucket_aux(idx, d_sorted);  // large batches take the reordered-batch pipeline over the (flat, sorted) leaf level; Compact: over a dense copy of the keys
    if (!ok) { free_index(idx); return nullptr; }
    finalize_view(idx);
    return idx;
}

// =================================================================================================
// Eytzinger baseline (eytzinger.rs:37-63).  The reference fills the array by an in-order recursion;
// here slot k (BFS index) computes its in-order rank in closed form: walk from the root to k and add,
// at every right turn, the size of the left subtree + 1; finally add the size of k's own left subtree.
// =================================================================================================
namespace {
__global__ void eytzinger_build_kernel(const uint32_t* __restrict__ vals, unsigned long long n, int H, uint32_t* __restrict__ out) {
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k <= n; k += (unsigned long long)gridDim.x * blockDim.x)
        out[k] = k == 0 ? 0xffffffffu : vals[eytz_rank(k, n, H)];
}
__global__ void check_sorted_kernel(const uint32_t* __restrict__ v, size_t n, unsigned* __restrict__ flags) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += (size_t)gridDim.x * blockDim.x)
        if (v[i + 1] < v[i]) atomicOr(flags, 2u);
}
}  // namespace

sst_index* build_eytzinger(const uint32_t* d_sorted, size_t n, int device) {
    clear_error();
    if (n == 0) { set_error(SST_ERR_ARG, "empty input"); return nullptr; }
    if (n >= ((size_t)1 << 32) - 1) { set_error(SST_ERR_UNSUPPORTED, "n must be < 2^32 - 1"); return nullptr; }
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard guard(device);
    if (!guard.ok) return nullptr;
    cudaStream_t st = thread_stream(device);
    unsigned* d_flags = nullptr;
    unsigned flags = 0;
    bool ok = SST_CUDA_OK(cudaMalloc(&d_flags, 4)) && SST_CUDA_OK(cudaMemsetAsync(d_flags, 0, 4, st));
    if (ok) {
        check_sorted_kernel<<<std::min<unsigned>(grid_for(n), cur_sms() * 16), kBuildThreads, 0, st>>>(d_sorted, n, d_flags);
        ok = SST_CUDA_OK(cudaMemcpyAsync(&flags, d_flags, 4, cudaMemcpyDeviceToHost, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    }
    cudaFree(d_flags);
    if (!ok) return nullptr;
    if (flags) { set_error(SST_ERR_ARG, "keys are not sorted"); return nullptr; }
    auto* idx = new sst_index();
    idx->device = device; idx->variant = SST_EYTZINGER; idx->node_b = 16; idx->n = n;
    int H = 0;
    while ((2ull << H) <= n) H++;                       // depth of the last level: floor(log2 n)
    int lg = 0;
    while ((2ull << lg) <= n + 1) lg++;
    idx->levels = lg + 1;                               // layers() = (n + 1).ilog2() + 1 (eytzinger.rs:72-74)
    idx->n_blocks = div_ceil(n + 1, 16);
    idx->eytz_words = n + 1;
    if (!SST_CUDA_OK(cudaMalloc(&idx->d_tree, (idx->n_blocks + 1) * 64))) { delete idx; return nullptr; }
    eytzinger_build_kernel<<<std::min<unsigned>(grid_for(n + 1), cur_sms() * 32), kBuildThreads, 0, st>>>(d_sorted, n, H, idx->d_tree);
    if (!SST_CUDA_OK(cudaGetLastError()) || !SST_CUDA_OK(cudaStreamSynchronize(st))) { cudaFree(idx->d_tree); delete idx; return nullptr; }
    idx->eytz_h = H;
    finalize_view(idx);
    return idx;
}

void finalize_view(sst_index* idx) {
    SstTreeView& v = idx->view;
    v = SstTreeView{};
    v.tree = idx->d_tree;
    v.levels = idx->levels;
    v.variant = idx->variant;
    v.node_b = idx->node_b;
    v.shift = (uint32_t)idx->shift;
    v.parts = idx->parts;
    v.n = idx->n;
    v.prefix_map = idx->d_prefix_map;
    v.part_start = idx->d_part_start;
    v.part_pos = idx->d_part_pos;
    const int H = idx->levels;
    if (idx->variant != SST_EYTZINGER) {
        for (int h = 0; h < H; h++) {
            v.level_slot[h] = (unsigned long long)idx->offsets[h] * 16;
            v.mult[h] = idx->node_b + 1;
        }
        v.leaf_slots = (unsigned long long)idx->layer_blocks[H - 1] * 16;
    }
    if (idx->variant == SST_EYTZINGER) {  // not a node tree: the view only carries the array
        v.levels = 1;
        v.leaf_slots = idx->eytz_words;
        return;
    }
    switch (idx->variant) {
        case SST_PLAIN: v.start_mul = 0; break;
        case SST_SIMPLE: v.start_mul = 16; break;                                       // partitioned_s_tree.rs:664-665
        case SST_COMPACT:                                                                // :703-721
            v.start_mul = 0;
            v.part_stride = (unsigned long long)idx->bpp * 16;
            v.leaf_slots = (unsigned long long)idx->layer_sizes[H - 1] * 16;
            v.upper = idx->d_upper;
            v.upper_stride = (unsigned long long)idx->offsets[H - 1] * 16;
            break;
        case SST_L1: v.start_mul = 16; v.mult[0] = (uint32_t)idx->l1_field; break;       // :745-759
        case SST_OVERLAPPING: v.start_mul = (uint32_t)(16 - idx->overlap); v.mult[0] = (uint32_t)idx->l1_field; break;  // :795-810
        case SST_MAP: v.start_mul = 0; v.mult[0] = 16; break;                            // :844-860
    }
}

}  // namespace sst

// =================================================================================================
// C ABI
// =================================================================================================
using namespace sst;

extern "C" {

// The keys may have been produced on any stream of the caller (e.g. a sort on torch's stream):
// builders are not on the hot path, so they simply wait for the whole device first.
static bool sync_device(int device) {
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return false; }
    DeviceGuard g(device);
    return g.ok && SST_CUDA_OK(cudaDeviceSynchronize());
}

sst_index_t* sst_stree_build_device(const uint32_t* d_sorted, size_t n, uint32_t node_b, uint32_t flags, int device) {
    clear_error();
    if (!sync_device(device)) return nullptr;
    return build_plain(d_sorted, false, n, node_b, flags, device);
}

static uint32_t* upload_keys(const uint32_t* sorted, size_t n, int device) {
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return nullptr; }
    DeviceGuard guard(device);
    if (!guard.ok) return nullptr;
    uint32_t* d = nullptr;
    if (!SST_CUDA_OK(cudaMalloc(&d, std::max<size_t>(n, 1) * 4))) return nullptr;
    // Stream-ordered with the builder's kernels (same per-thread stream).  A plain cudaMemcpy from
    // pageable memory may return before the DMA has landed and would not order against that stream.
    cudaStream_t st = thread_stream(device);
    if (n && (!SST_CUDA_OK(cudaMemcpyAsync(d, sorted, n * 4, cudaMemcpyHostToDevice, st)) || !SST_CUDA_OK(cudaStreamSynchronize(st)))) {
        cudaFree(d);
        return nullptr;
    }
    return d;
}

sst_index_t* sst_stree_build(const uint32_t* sorted, size_t n, uint32_t node_b, uint32_t flags, int device) {
    clear_error();
    if (n == 0 || !sorted) { set_error(SST_ERR_ARG, "empty input (reference: assert!(n > 0), s_tree.rs:93)"); return nullptr; }
    uint32_t* d = upload_keys(sorted, n, device);
    if (!d) return nullptr;
    sst_index* idx = build_plain(d, false, n, node_b, flags, device);
    { DeviceGuard g(device); cudaFree(d); }
    return idx;
}

sst_index_t* sst_pstree_build_device(const uint32_t* d_sorted, size_t n, uint32_t b, int variant, int device) {
    clear_error();
    if (!sync_device(device)) return nullptr;
    return build_partitioned(d_sorted, n, b, variant, device);
}

sst_index_t* sst_pstree_build(const uint32_t* sorted, size_t n, uint32_t b, int variant, int device) {
    clear_error();
    if (n == 0 || !sorted) { set_error(SST_ERR_ARG, "empty input"); return nullptr; }
    uint32_t* d = upload_keys(sorted, n, device);
    if (!d) return nullptr;
    sst_index* idx = build_partitioned(d, n, b, variant, device);
    { DeviceGuard g(device); cudaFree(d); }
    return idx;
}

sst_index_t* sst_eytzinger_build_device(const uint32_t* d_sorted, size_t n, int device) {
    clear_error();
    if (!sync_device(device)) return nullptr;
    return build_eytzinger(d_sorted, n, device);
}

sst_index_t* sst_eytzinger_build(const uint32_t* sorted, size_t n, int device) {
    clear_error();
    if (n == 0 || !sorted) { set_error(SST_ERR_ARG, "empty input"); return nullptr; }
    uint32_t* d = upload_keys(sorted, n, device);
    if (!d) return nullptr;
    sst_index* idx = build_eytzinger(d, n, device);
    { DeviceGuard g(device); cudaFree(d); }
    return idx;
}

void sst_index_free(sst_index_t* idx) {
    if (!idx) return;
    DeviceGuard g(idx->device);
    free_index(idx);
}

// SearchIndex::size: s_tree.rs:56-58, partitioned_s_tree.rs:101-103
size_t sst_index_size_bytes(const sst_index_t* idx) {
    if (!idx) return 0;
    if (idx->variant == SST_EYTZINGER) return idx->eytz_words * 4;  // eytzinger.rs:76-78
    return idx->n_blocks * 64 + idx->prefix_map_len * 4;
}
size_t sst_index_image_words(const sst_index_t* idx) {
    if (!idx) return 0;
    return idx->variant == SST_EYTZINGER ? idx->eytz_words : idx->n_blocks * 16;
}
// SearchIndex::layers: s_tree.rs:52-54, partitioned_s_tree.rs:105-107
size_t sst_index_layers(const sst_index_t* idx) { return idx ? (size_t)idx->levels + (idx->variant == SST_MAP ? 1 : 0) : 0; }
size_t sst_index_len(const sst_index_t* idx) { return idx ? idx->n : 0; }
int sst_index_device(const sst_index_t* idx) { return idx ? idx->device : -1; }
int sst_index_variant(const sst_index_t* idx) { return idx ? idx->variant : -1; }
size_t sst_index_nodes(const sst_index_t* idx) { return idx ? idx->n_blocks : 0; }
size_t sst_index_levels(const sst_index_t* idx) { return idx ? (size_t)idx->levels : 0; }

int sst_index_offsets(const sst_index_t* idx, uint64_t* out) {
    if (!idx || !out) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    for (int h = 0; h < idx->levels; h++) out[h] = idx->offsets[h];
    return SST_OK;
}

int sst_index_image(const sst_index_t* idx, uint32_t* out) {
    if (!idx || !out) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    DeviceGuard g(idx->device);
    if (!g.ok) return SST_ERR_CUDA;
    return SST_CUDA_OK(cudaMemcpy(out, idx->d_tree, sst_index_image_words(idx) * 4, cudaMemcpyDeviceToHost)) ? SST_OK : SST_ERR_CUDA;
}

int sst_index_params(const sst_index_t* idx, uint64_t* out) {
    if (!idx || !out) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    out[0] = idx->shift; out[1] = idx->parts; out[2] = idx->bpp; out[3] = idx->l1_field; out[4] = idx->overlap;
    out[5] = idx->has_overlap; out[6] = idx->max_bucket; out[7] = idx->prefix_map_len;
    return SST_OK;
}

int sst_index_prefix_map(const sst_index_t* idx, uint32_t* out) {
    if (!idx || !out) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (!idx->prefix_map_len) return SST_OK;
    DeviceGuard g(idx->device);
    if (!g.ok) return SST_ERR_CUDA;
    return SST_CUDA_OK(cudaMemcpy(out, idx->d_prefix_map, idx->prefix_map_len * 4, cudaMemcpyDeviceToHost)) ? SST_OK : SST_ERR_CUDA;
}

}  // extern "C"
