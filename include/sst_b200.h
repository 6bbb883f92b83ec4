/* sst_b200.h -- C ABI of the B200 (sm_100a) static-search-tree / suffix-array search library.
 *
 * This is the drop-in boundary for the ONE data-parallel hot path of
 * RagnarGrootKoerkamp/suffix-array-searching: batched lower_bound over a static sorted u32
 * array in the S+-tree layout (plain and prefix-partitioned), and suffix-array pattern search.
 * The reference has no FFI of its own (no extern "C" anywhere); the entry points below are what
 * a Rust `extern "C"` block would bind so that the reference's `SearchIndex` / `SearchScheme`
 * traits (static-search-tree/src/lib.rs:30-61) and the free SA search functions
 * (suffix-array-searching/src/sa_search.rs:98-112) can be served by the GPU.  INTEGRATION.md
 * shows that binding.  Paths below: sst = static-search-tree/src, sas = suffix-array-searching/src.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA or torch types in any signature (streams are
 *     passed as `void*` holding a cudaStream_t; NULL is the CUDA legacy default stream, which is
 *     also what torch.cuda.current_stream().cuda_stream returns for torch's default stream).
 *   - the caller owns every input/output buffer; a handle owns its device memory.
 *   - functions returning int return 0 on success, non-zero on failure; pointer-returning
 *     builders return NULL on failure.  sst_last_error() gives the message (thread-local).
 *     The reference's convention is panic (assert!/unwrap, Cargo.toml:12 panic=abort); the Rust
 *     shim turns a non-zero status into panic!, and a NULL from sst_pstree_build whose
 *     sst_last_status() is SST_ERR_CAPACITY into `None` (partitioned_s_tree.rs:271-274,463-466).
 *   - every query entry point is re-entrant: an index is immutable after build and may be
 *     queried from many host threads at once (`SearchIndex: Sync`, lib.rs:30).
 *   - there is no CPU fallback: without a usable sm_100 device every call fails loudly.
 *   - keys and queries must be <= SST_MAX (0x7fffffff): node comparison is SIGNED like the
 *     reference's AVX2 compare (sst/node.rs:91-108); builders reject larger keys
 *     (sst/s_tree.rs:87-89).  A query above every key returns value SST_MAX and index n (the
 *     reference reads zeroed allocation slack there).
 */
#ifndef SST_B200_H
#define SST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SST_MAX 0x7fffffffu /* sst/node.rs:5 */

/* status codes */
enum {
    SST_OK = 0,
    SST_ERR_CUDA = 1,      /* CUDA runtime/driver error, or no sm_100 device */
    SST_ERR_ARG = 2,       /* reference would panic: empty input, key > MAX, unsorted, bad flags */
    SST_ERR_CAPACITY = 3,  /* reference returns None: index would exceed its memory cap */
    SST_ERR_UNSUPPORTED = 4
};

/* STree::new_params flags (sst/s_tree.rs:72-77) */
enum { SST_LEFT_MAX = 1, SST_REVERSE_STORAGE = 2, SST_FULL_ARRAY = 4 };

/* PartitionedSTree layouts (sst/partitioned_s_tree.rs:34-98) */
enum { SST_PLAIN = 0, SST_SIMPLE = 1, SST_COMPACT = 2, SST_L1 = 3, SST_OVERLAPPING = 4, SST_MAP = 5,
       SST_EYTZINGER = 6 /* baseline layout of sst/eytzinger.rs, built by sst_eytzinger_build */ };

/* search schemes (kernel variants) selectable per query call, like the reference's many
 * SearchScheme closures over one index (sst/bin/bench.rs:93-96).  All return identical results. */
enum {
    SST_SCHEME_AUTO = 0,     /* best measured kernel for the index */
    SST_SCHEME_GROUP4 = 1,   /* 4 lanes x 16 B per node (LDG.128), every level from L1/L2 */
    SST_SCHEME_GROUP16 = 2,  /* 16 lanes x 4 B per node + ballot/popc (north-star baseline) */
    SST_SCHEME_GROUP2 = 3,   /* 2 lanes x 32 B per node (LDG.256) */
    SST_SCHEME_GENERIC = 4,  /* one thread per query, any layout */
    SST_SCHEME_TABLE = 5,    /* top levels answered by a shared-memory rank table (TMA-staged), rest as GROUP2 */
    SST_SCHEME_BINSEARCH = 6, /* baseline: SortedVec::binary_search (sst/binary_search.rs:36-49) over the leaf level */
    SST_SCHEME_BUCKETED = 7   /* reordered batch: every 16384-query tile sorted by key range, each range answered from shared memory
                                 + one leaf sector per query, in place, answers un-permuted (B=16 trees of 2^22..2^30 keys, large batches) */
};

/* SA search modes */
enum { SST_SA_BINARY = 0, SST_SA_MLR = 1 };

typedef struct sst_index sst_index_t; /* S+-tree / partitioned S+-tree on one device */
typedef struct sst_sa sst_sa_t;       /* text + suffix array on one device */
typedef struct sst_multi sst_multi_t; /* replicas of one index on several devices */
typedef struct sst_multi_sa sst_multi_sa_t; /* replicas of one text + suffix array on several devices */

/* ---- library ------------------------------------------------------------------------------ */
const char* sst_last_error(void);
int sst_last_status(void);
int sst_device_count(void);          /* number of usable sm_100 devices, 0 if none */
/* Page-locked host memory for query/result buffers: sst_query overlaps H2D, kernel and D2H only when the
 * host buffers are pinned (pageable memory is staged by the driver).  NULL on failure. */
void* sst_host_alloc(size_t bytes);
void sst_host_free(void* p);
const char* sst_version(void);
/* Tuning / A-B options (the table in csrc/common.cuh: BK_R, SA_CHUNK, SCHEME, ...).  `name` with or without the SST_
 * prefix.  Each option starts at its default, or at the value of the environment variable SST_<NAME> read ONCE when the
 * library is loaded; afterwards only these calls change it (nothing on a query path reads the environment).  A value
 * outside the option's range is rejected with SST_ERR_ARG.  A change applies to calls and index builds that start after it. */
int sst_set_option(const char* name, long long value);
int sst_get_option(const char* name, long long* out_value);
void sst_reset_options(void);          /* back to the load-time values */
int sst_option_count(void);
const char* sst_option_name(int i);    /* NULL when i is out of range */
/* Binds the calling host thread to the CPUs local to `device` (PCIe/NUMA topology from sysfs), so that the page-locked
 * buffers it allocates afterwards and its copies stay on the GPU's socket.  The reference pins nothing (rayon workers,
 * sst/bin/bench.rs:558-573) because its data never leaves host memory; here every query crosses PCIe once each way.
 * Returns the size of the CPU set, 0 when the topology is not visible (nothing changed), < 0 on a CUDA error.
 * sst_multi_* workers call it for their device. */
int sst_bind_thread_to_device(int device);

/* ---- S+-tree: replaces STree::<B,16>::new_params (sst/s_tree.rs:72-176) ------------------- */
/* `sorted` is a HOST pointer to n ascending keys, all <= SST_MAX.  node_b is B (16, or 15 for
 * STree15, sst/s_tree.rs:19-20).  flags = OR of SST_LEFT_MAX | SST_REVERSE_STORAGE | SST_FULL_ARRAY;
 * SearchIndex::new is flags = 0 (sst/s_tree.rs:47-50). */
sst_index_t* sst_stree_build(const uint32_t* sorted, size_t n, uint32_t node_b, uint32_t flags, int device);
/* Same, keys already resident on `device` (the GPU layout builder proper). */
sst_index_t* sst_stree_build_device(const uint32_t* d_sorted, size_t n, uint32_t node_b, uint32_t flags, int device);

/* ---- prefix-partitioned S+-tree: replaces PartitionedSTree::<16,16,Tp>::try_new
 *      (sst/partitioned_s_tree.rs:241-351 Compact, :364-649 Simple/L1/Overlapping/Map).
 *      NULL with sst_last_status()==SST_ERR_CAPACITY is the reference's `None`. */
sst_index_t* sst_pstree_build(const uint32_t* sorted, size_t n, uint32_t b, int variant, int device);
sst_index_t* sst_pstree_build_device(const uint32_t* d_sorted, size_t n, uint32_t b, int variant, int device);

/* ---- Eytzinger baseline: replaces Eytzinger::new + search (sst/eytzinger.rs:37-89).  One-based BFS layout of
 *      the sorted keys with element 0 = u32::MAX; UNSIGNED compares; a query above every key returns
 *      0xffffffff (eytzinger.rs:222-229) and index n.  Image = n + 1 words (sst_index_image_words). */
sst_index_t* sst_eytzinger_build(const uint32_t* sorted, size_t n, int device);
sst_index_t* sst_eytzinger_build_device(const uint32_t* d_sorted, size_t n, int device);

void sst_index_free(sst_index_t* idx);

/* SearchIndex::size (bytes) and ::layers (sst/lib.rs:35-40; s_tree.rs:52-58; partitioned_s_tree.rs:100-107) */
size_t sst_index_size_bytes(const sst_index_t* idx);
size_t sst_index_layers(const sst_index_t* idx);
size_t sst_index_len(const sst_index_t* idx);   /* n */
int sst_index_device(const sst_index_t* idx);
int sst_index_variant(const sst_index_t* idx);

/* Introspection for layout parity tests (the `tree`, `offsets`, `shift`, `bpp`, `l1`, `overlap`,
 * `prefix_map` fields of sst/s_tree.rs:14-17 and sst/partitioned_s_tree.rs:19-32). */
size_t sst_index_nodes(const sst_index_t* idx);                  /* tree.len() */
size_t sst_index_levels(const sst_index_t* idx);                 /* offsets.len() */
int sst_index_offsets(const sst_index_t* idx, uint64_t* out);     /* out[levels], node units */
size_t sst_index_image_words(const sst_index_t* idx);             /* u32 words in the image: nodes*16, or n+1 for Eytzinger */
int sst_index_image(const sst_index_t* idx, uint32_t* out);       /* out[image_words], device -> host */
/* out[8] = shift, parts, bpp, l1, overlap, has_overlap, max_bucket, prefix_map_len */
int sst_index_params(const sst_index_t* idx, uint64_t* out);
int sst_index_prefix_map(const sst_index_t* idx, uint32_t* out);

/* ---- queries: replaces SearchScheme::query (sst/lib.rs:51-61) and the batch_* / search
 *      functions behind it (sst/s_tree.rs:196-385, partitioned_s_tree.rs:654-880).
 * out_vals[i] = first key >= qs[i] (the reference's return value);
 * out_idx[i]  = its index in the sorted input (nullable; the `l` of sst/binary_search.rs:36-49).
 * Any nq is accepted (the reference's `batched` asserts nq % P == 0, lib.rs:85-92). */
int sst_query(const sst_index_t* idx, const uint32_t* qs, size_t nq, uint32_t* out_vals, uint64_t* out_idx,
              int scheme); /* HOST buffers: H2D, kernel, D2H, synchronous */
int sst_query_device(const sst_index_t* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_out_vals,
                     uint64_t* d_out_idx, int scheme, void* stream); /* DEVICE buffers, asynchronous on `stream` */
/* Pre-sizes the calling thread's scratch buffers of the reordered-batch pipeline (6-10 bytes per query, at most 2^30 queries'
 * worth) for batches of up to nq queries on this index, so that later sst_query_device calls from this thread allocate
 * nothing: they are then asynchronous on `stream` and can be captured into a CUDA graph.  Without it the first large
 * batch allocates (and synchronises the device) once.  sst_query_release frees the calling thread's scratch on every device
 * (it is also freed when the thread exits). */
int sst_query_reserve(const sst_index_t* idx, size_t nq, int want_index);
void sst_query_release(void);
/* Measures on this index, on its own device, from which batch size on the reordered-batch pipeline beats the direct kernel
 * (synthetic queries that follow the key distribution, batches of 2^20 .. max_nq), and makes SST_SCHEME_AUTO use that
 * crossover for this index instead of the default rule.  A few tens of milliseconds; 12 bytes of device memory per query of
 * the largest batch while it runs.  *out_min_nq (nullable): the crossover, SIZE_MAX if the pipeline never won, 0 for an
 * index the pipeline does not serve (nothing to calibrate). */
int sst_query_calibrate(sst_index_t* idx, size_t max_nq, size_t* out_min_nq);
/* Number of kernel launches sst_query_device issues for this index/scheme (for launch accounting). */
int sst_query_launches(const sst_index_t* idx, int scheme);
/* The kernel SST_SCHEME_AUTO resolves to for a batch of nq queries on this index (*out_scheme; partitioned layouts
 * report SST_SCHEME_AUTO = their lane-group kernel) and the number of kernel launches the call issues (*out_launches). */
int sst_query_plan(const sst_index_t* idx, size_t nq, int scheme, int want_index, int* out_scheme, int* out_launches);

/* ---- suffix arrays: replaces SaNaive::build / SA::build (sas/sa_search.rs:30-57,
 *      sas/experiments.rs:19-38; the libsais call at sa_search.rs:33) and binary_search
 *      (sas/sa_search.rs:98-112, sas/experiments.rs:51-64).  A handle also carries GPU-only accelerators that do not
 *      change any result: a pivot-prefix table and, for texts over {0,1,2,3}, a k-mer table (SaNaive's prefix `table`,
 *      sa_search.rs:59-85), the next 15 or 48 bases of every suffix inlined next to its entry ("Inlining values",
 *      todo.org:18-19) and the text at 2 bits per base for the compares behind them; each is skipped when device memory
 *      is short. */
sst_sa_t* sst_sa_build(const uint8_t* text, size_t n, int device);               /* GPU SA construction */
sst_sa_t* sst_sa_build_device(const uint8_t* d_text, size_t n, int device);
sst_sa_t* sst_sa_from_parts(const uint8_t* text, size_t n, const uint32_t* sa, int device); /* upload a ready SA */
void sst_sa_free(sst_sa_t* sa);
size_t sst_sa_len(const sst_sa_t* sa);
int sst_sa_get(const sst_sa_t* sa, uint32_t* out_sa);                             /* device -> host, n entries */
/* out_sa[i] = sa[positions[i]] (0xffffffff for a position >= n): e.g. the occurrences sa[lo .. hi) of a pattern. */
int sst_sa_gather(const sst_sa_t* sa, const uint64_t* positions, size_t count, uint32_t* out_sa);
/* Number of adjacent suffix pairs violating strict order (sas/sa_search.rs:36-38); 0 == valid. */
int sst_sa_check(const sst_sa_t* sa, uint64_t* out_violations);
/* patterns are packed back to back in `pats`; pattern i is pats[pat_off[i] .. pat_off[i+1]).
 * out_lo[i]  = l of binary_search (first suffix >= pattern);
 * out_hi[i]  = first index >= lo whose suffix does not start with the pattern (nullable; not in the reference);
 * out_pos[i] = sa[lo] -- the reference's return value (nullable; 0xffffffff when lo == n). */
int sst_sa_search(const sst_sa_t* sa, const uint8_t* pats, const uint64_t* pat_off, size_t npat, int mode,
                  uint32_t* out_lo, uint32_t* out_hi, uint32_t* out_pos);
int sst_sa_search_device(const sst_sa_t* sa, const uint8_t* d_pats, const uint64_t* d_pat_off, size_t npat, int mode,
                         uint32_t* d_out_lo, uint32_t* d_out_hi, uint32_t* d_out_pos, void* stream);
/* The reference's probe counter (`cnt: &mut usize`, one increment per loop iteration of binary_search, sas/sa_search.rs:98-112,
 * printed per query by bench, :423-436): runs exactly that loop on the device (plain binary search over [0, n), no table) and
 * returns out_pos[i] = sa[l] (nullable) and out_probes[i] = its number of iterations.  Host buffers; a tracing aid. */
int sst_sa_search_probes(const sst_sa_t* sa, const uint8_t* pats, const uint64_t* pat_off, size_t npat, uint32_t* out_pos,
                         uint32_t* out_probes);

/* ---- input data formats (the callers' side of the path) --------------------------------------
 * sst_fasta_encode: FASTA text -> one byte per base in 0..3, replaces read_fasta_file
 *   (sas/util.rs:144-169: headers dropped, line ends stripped, A/C/G/T in either case -> 0..3, every
 *   other byte -> 0).  out_codes needs room for `len` bytes; *out_len receives the number of bases.
 * sst_kmer_keys: the `--human` key mode of the bench harness (sst/bin/bench.rs:60-76): key i is the
 *   2-bit pack of codes[i..i+k) (first base most significant) masked to 31 bits, key 0 is SST_MAX;
 *   count = min(n - k + 1, max_keys).  sort != 0 also sorts them (bench.rs:89) so that they can go
 *   straight into sst_stree_build_device. */
int sst_fasta_encode(const char* fasta, size_t len, uint8_t* out_codes, size_t* out_len, int device);
int sst_fasta_encode_device(const char* d_fasta, size_t len, uint8_t* d_out_codes, size_t* out_len, int device);
int sst_kmer_keys(const uint8_t* codes, size_t n, uint32_t k, size_t max_keys, uint32_t* out_keys, size_t* out_count, int sort,
                  int device);
int sst_kmer_keys_device(const uint8_t* d_codes, size_t n, uint32_t k, size_t max_keys, uint32_t* d_out_keys, size_t* out_count,
                         int sort, int device);

/* ---- multi-GPU: index replicated per device, query batch sharded contiguously
 *      (chunk = ceil(nq / G), the rule of sst/bin/bench.rs:558-573), one host thread and one
 *      stream per device, no collective.  The host keys are uploaded once; the other replicas receive them
 *      device to device and run the GPU builder in parallel. */
sst_multi_t* sst_multi_stree_build(const uint32_t* sorted, size_t n, uint32_t node_b, uint32_t flags,
                                   const int* devices, int n_devices);
sst_multi_t* sst_multi_pstree_build(const uint32_t* sorted, size_t n, uint32_t b, int variant,
                                    const int* devices, int n_devices);
int sst_multi_query(const sst_multi_t* m, const uint32_t* qs, size_t nq, uint32_t* out_vals, uint64_t* out_idx,
                    int scheme);
/* The same with the shards already resident on the replicas' devices: shard i = d_qs[i] (nq[i] queries) with outputs
 * d_out_vals[i] / d_out_idx[i] (d_out_idx may be NULL) on the device of replica i.  Returns when every shard is done. */
int sst_multi_query_device(const sst_multi_t* m, const uint32_t* const* d_qs, const size_t* nq, uint32_t* const* d_out_vals,
                           uint64_t* const* d_out_idx, int scheme);
int sst_multi_devices(const sst_multi_t* m);
void sst_multi_free(sst_multi_t* m);
/* Suffix arrays: text + SA (+ the GPU-only accelerators) replicated per device, the pattern batch sharded contiguously
 * (chunk = ceil(npat / G)); replaces the serial callers of sas/sa_search.rs:423-451 in the harness shape of
 * sst/bin/bench.rs:558-573.  The index is built ONCE on devices[0] and copied to the other devices peer to peer. */
sst_multi_sa_t* sst_multi_sa_build(const uint8_t* text, size_t n, const int* devices, int n_devices);
sst_multi_sa_t* sst_multi_sa_from_parts(const uint8_t* text, size_t n, const uint32_t* sa, const int* devices, int n_devices);
int sst_multi_sa_search(const sst_multi_sa_t* m, const uint8_t* pats, const uint64_t* pat_off, size_t npat, int mode,
                        uint32_t* out_lo, uint32_t* out_hi, uint32_t* out_pos);
int sst_multi_sa_devices(const sst_multi_sa_t* m);
void sst_multi_sa_free(sst_multi_sa_t* m);

/* ---- measurement helpers (used by bench.py; not part of the reference surface) ------------ */
/* Runs sst_query_device `iters` times on an internal stream and returns the mean kernel time in
 * milliseconds measured with CUDA events on that stream (<0 on error). */
double sst_time_query_device(const sst_index_t* idx, const uint32_t* d_qs, size_t nq, uint32_t* d_out_vals,
                             uint64_t* d_out_idx, int scheme, int warmup, int iters);
/* Random 64-byte gather probe over `bytes` of device memory: the practical ceiling of the
 * access pattern of one tree level.  Returns GB/s (<0 on error). lanes_per_node in {1,2,4,8,16}. */
double sst_probe_gather64(int device, size_t bytes, size_t n_gathers, int lanes_per_node, int iters);
/* With the option BK_TIMING = 1 the reordered-batch pipeline times its stages with CUDA events (and synchronises the
 * stream); this returns the calling thread's last {partition, plan, -, search, un-permute} times in ms (0 = nothing
 * recorded; with BK_V1 = 1: {rank, plan, scatter, search, gather}). */
int sst_last_stage_ms(double* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* SST_B200_H */
