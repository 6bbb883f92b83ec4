// formats.cu -- the data formats on the input side of the hot path, on the GPU.
//
// Replaces (reference paths):
//   read_fasta_file                      suffix-array-searching/src/util.rs:144-169
//     (needletail::parse_fastx_file picks the format from the first byte.  FASTA ('>'): header lines start with '>',
//      sequence lines are concatenated with line ends stripped.  FASTQ ('@'): four-line records -- header, sequence,
//      '+' line, qualities -- of which only the second line is sequence.  A/C/G/T in either case map to 0..3, every
//      other byte to 0)
//   k-mer key generation of `--human`    static-search-tree/src/bin/bench.rs:60-76
//     (rolling 2-bit pack of k = 16 bases, masked to 31 bits; vals[0] = MAX)
//   the key sort before every build      static-search-tree/src/bin/bench.rs:89 (rdst radix sort)
//
// Parsing is data-parallel: a max-scan gives every byte the start of its line, bytes of header lines
// and line ends are dropped by a flagged compaction (cub::DeviceScan / cub::DeviceSelect are the scan
// and compaction primitives; the classification and packing kernels are hand-written).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>

#include "common.cuh"

namespace sst {
namespace {

constexpr int kThreads = 256;
inline unsigned grid_for(size_t work) { return (unsigned)std::min<size_t>(div_ceil(work, (size_t)kThreads), (size_t)cur_sms() * 32); }

// start-of-line marker: position i if byte i starts a line, else 0 (a running max gives the line start)
__global__ void fasta_line_starts(const char* __restrict__ s, size_t n, unsigned long long* __restrict__ start) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        start[i] = (i == 0 || s[i - 1] == '\n') ? (unsigned long long)i : 0ull;
}

struct MaxOp {
    __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; }
};

__device__ __forceinline__ uint8_t base_code(char c) {  // util.rs:145-155: map[] is zero except for ACGT/acgt
    switch (c) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0;
    }
}

__global__ void fasta_classify(const char* __restrict__ s, size_t n, const unsigned long long* __restrict__ line_start,
                               uint8_t* __restrict__ code, uint8_t* __restrict__ keep) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const char c = s[i];
        const bool header = s[line_start[i]] == '>';
        keep[i] = (!header && c != '\n' && c != '\r') ? 1 : 0;
        code[i] = base_code(c);
    }
}

// FASTQ: number of line ends before byte i (an exclusive sum gives every byte its line number; line % 4 == 1 is sequence)
__global__ void fastq_line_ends(const char* __restrict__ s, size_t n, unsigned long long* __restrict__ flag) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) flag[i] = s[i] == '\n' ? 1ull : 0ull;
}
__global__ void fastq_classify(const char* __restrict__ s, size_t n, const unsigned long long* __restrict__ line_no, uint8_t* __restrict__ code,
                               uint8_t* __restrict__ keep) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const char c = s[i];
        keep[i] = ((line_no[i] & 3ull) == 1ull && c != '\n' && c != '\r') ? 1 : 0;
        code[i] = base_code(c);
    }
}

// bench.rs:64-73: key_i = (2-bit pack of codes[i .. i+k), first base most significant) & (2^(2k) - 1) & MAX
__global__ void kmer_keys_kernel(const uint8_t* __restrict__ codes, size_t count, unsigned k, uint32_t* __restrict__ keys) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long key = 0;
        for (unsigned j = 0; j < k; j++) key = (key << 2) | (unsigned long long)(codes[i + j] & 3u);
        if (k < 32) key &= (1ull << (2 * k)) - 1ull;
        uint32_t v = (uint32_t)key & kMax;
        if (i == 0) v = kMax;  // bench.rs:74
        keys[i] = v;
    }
}

}  // namespace
}  // namespace sst

using namespace sst;

extern "C" {

int sst_fasta_encode_device(const char* d_fasta, size_t len, uint8_t* d_out_codes, size_t* out_len, int device) {
    clear_error();
    if (!out_len || (len && (!d_fasta || !d_out_codes))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    *out_len = 0;
    if (len == 0) return SST_OK;
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return SST_ERR_CUDA; }
    DeviceGuard g(device);
    if (!g.ok || !SST_CUDA_OK(cudaDeviceSynchronize())) return SST_ERR_CUDA;
    cudaStream_t st = thread_stream(device);
    unsigned long long* d_start = nullptr;
    uint8_t *d_code = nullptr, *d_keep = nullptr;
    unsigned long long* d_count = nullptr;
    void* tmp = nullptr;
    size_t tb1 = 0, tb2 = 0;
    bool ok = SST_CUDA_OK(cudaMalloc(&d_start, len * 8)) && SST_CUDA_OK(cudaMalloc(&d_code, len)) && SST_CUDA_OK(cudaMalloc(&d_keep, len)) &&
              SST_CUDA_OK(cudaMalloc(&d_count, 8));
    ok = ok && SST_CUDA_OK(cub::DeviceScan::InclusiveScan(nullptr, tb1, d_start, d_start, MaxOp(), (unsigned long long)len, st)) &&
         SST_CUDA_OK(cub::DeviceSelect::Flagged(nullptr, tb2, d_code, d_keep, d_out_codes, d_count, (unsigned long long)len, st));
    ok = ok && SST_CUDA_OK(cudaMalloc(&tmp, std::max(tb1, tb2)));
    unsigned long long count = 0;
    char first = 0;
    ok = ok && SST_CUDA_OK(cudaMemcpyAsync(&first, d_fasta, 1, cudaMemcpyDeviceToHost, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    if (ok && first == '@') {  // FASTQ (needletail decides by the first byte): line number of every byte, line % 4 == 1 is sequence
        size_t tb3 = 0;
        ok = SST_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_start, d_start, (unsigned long long)len, st));
        if (ok && tb3 > std::max(tb1, tb2)) { cudaFree(tmp); tmp = nullptr; ok = SST_CUDA_OK(cudaMalloc(&tmp, tb3)); }
        if (ok) {
            fastq_line_ends<<<grid_for(len), kThreads, 0, st>>>(d_fasta, len, d_start);
            ok = SST_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb3, d_start, d_start, (unsigned long long)len, st));
        }
        if (ok) {
            fastq_classify<<<grid_for(len), kThreads, 0, st>>>(d_fasta, len, d_start, d_code, d_keep);
            size_t tb = std::max(tb2, tb3);
            ok = SST_CUDA_OK(cub::DeviceSelect::Flagged(tmp, tb, d_code, d_keep, d_out_codes, d_count, (unsigned long long)len, st)) &&
                 SST_CUDA_OK(cudaMemcpyAsync(&count, d_count, 8, cudaMemcpyDeviceToHost, st)) && SST_CUDA_OK(cudaStreamSynchronize(st)) &&
                 SST_CUDA_OK(cudaGetLastError());
        }
    } else if (ok) {
        fasta_line_starts<<<grid_for(len), kThreads, 0, st>>>(d_fasta, len, d_start);
        size_t tb = tb1;
        ok = SST_CUDA_OK(cub::DeviceScan::InclusiveScan(tmp, tb, d_start, d_start, MaxOp(), (unsigned long long)len, st));
        if (ok) {
            fasta_classify<<<grid_for(len), kThreads, 0, st>>>(d_fasta, len, d_start, d_code, d_keep);
            tb = tb2;
            ok = SST_CUDA_OK(cub::DeviceSelect::Flagged(tmp, tb, d_code, d_keep, d_out_codes, d_count, (unsigned long long)len, st)) &&
                 SST_CUDA_OK(cudaMemcpyAsync(&count, d_count, 8, cudaMemcpyDeviceToHost, st)) && SST_CUDA_OK(cudaStreamSynchronize(st)) &&
                 SST_CUDA_OK(cudaGetLastError());
        }
    }
    cudaFree(d_start); cudaFree(d_code); cudaFree(d_keep); cudaFree(d_count); cudaFree(tmp);
    if (!ok) return SST_ERR_CUDA;
    *out_len = (size_t)count;
    return SST_OK;
}

int sst_fasta_encode(const char* fasta, size_t len, uint8_t* out_codes, size_t* out_len, int device) {
    clear_error();
    if (!out_len || (len && (!fasta || !out_codes))) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    *out_len = 0;
    if (len == 0) return SST_OK;
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return SST_ERR_CUDA; }
    DeviceGuard g(device);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t st = thread_stream(device);
    char* d_in = nullptr;
    uint8_t* d_out = nullptr;
    bool ok = SST_CUDA_OK(cudaMalloc(&d_in, len)) && SST_CUDA_OK(cudaMalloc(&d_out, len)) &&
              SST_CUDA_OK(cudaMemcpyAsync(d_in, fasta, len, cudaMemcpyHostToDevice, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    int rc = ok ? sst_fasta_encode_device(d_in, len, d_out, out_len, device) : SST_ERR_CUDA;
    if (rc == SST_OK && *out_len)
        rc = SST_CUDA_OK(cudaMemcpy(out_codes, d_out, *out_len, cudaMemcpyDeviceToHost)) ? SST_OK : SST_ERR_CUDA;
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}

int sst_kmer_keys_device(const uint8_t* d_codes, size_t n, uint32_t k, size_t max_keys, uint32_t* d_out_keys, size_t* out_count,
                         int sort, int device) {
    clear_error();
    if (!out_count || !d_codes || !d_out_keys) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    if (k < 1 || k > 32) { set_error(SST_ERR_ARG, "k must be in 1..32"); return SST_ERR_ARG; }
    *out_count = 0;
    if (n < k) return SST_OK;
    const size_t count = std::min(n - k + 1, max_keys);  // bench.rs:69
    if (count == 0) return SST_OK;
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return SST_ERR_CUDA; }
    DeviceGuard g(device);
    if (!g.ok || !SST_CUDA_OK(cudaDeviceSynchronize())) return SST_ERR_CUDA;
    cudaStream_t st = thread_stream(device);
    bool ok = true;
    if (!sort) {
        kmer_keys_kernel<<<grid_for(count), kThreads, 0, st>>>(d_codes, count, k, d_out_keys);
    } else {  // bench.rs:89: vals.radix_sort_unstable()
        uint32_t* d_tmp = nullptr;
        void* tmp = nullptr;
        size_t tb = 0;
        ok = SST_CUDA_OK(cudaMalloc(&d_tmp, count * 4)) &&
             SST_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_tmp, d_out_keys, (unsigned long long)count, 0, 31, st)) &&
             SST_CUDA_OK(cudaMalloc(&tmp, tb));
        if (ok) {
            kmer_keys_kernel<<<grid_for(count), kThreads, 0, st>>>(d_codes, count, k, d_tmp);
            ok = SST_CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, tb, d_tmp, d_out_keys, (unsigned long long)count, 0, 31, st));
        }
        ok = ok && SST_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(d_tmp); cudaFree(tmp);
    }
    ok = ok && SST_CUDA_OK(cudaGetLastError()) && SST_CUDA_OK(cudaStreamSynchronize(st));
    if (!ok) return SST_ERR_CUDA;
    *out_count = count;
    return SST_OK;
}

int sst_kmer_keys(const uint8_t* codes, size_t n, uint32_t k, size_t max_keys, uint32_t* out_keys, size_t* out_count, int sort,
                  int device) {
    clear_error();
    if (!out_count || !codes || !out_keys) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    *out_count = 0;
    if (k < 1 || k > 32) { set_error(SST_ERR_ARG, "k must be in 1..32"); return SST_ERR_ARG; }
    if (n < k || max_keys == 0) return SST_OK;
    if (!device_usable(device)) { set_error(SST_ERR_CUDA, "device is not an sm_100 GPU (no CPU fallback)"); return SST_ERR_CUDA; }
    DeviceGuard g(device);
    if (!g.ok) return SST_ERR_CUDA;
    cudaStream_t st = thread_stream(device);
    const size_t count = std::min(n - k + 1, max_keys), need = count + k - 1;
    uint8_t* d_c = nullptr;
    uint32_t* d_k = nullptr;
    bool ok = SST_CUDA_OK(cudaMalloc(&d_c, need)) && SST_CUDA_OK(cudaMalloc(&d_k, count * 4)) &&
              SST_CUDA_OK(cudaMemcpyAsync(d_c, codes, need, cudaMemcpyHostToDevice, st)) && SST_CUDA_OK(cudaStreamSynchronize(st));
    int rc = ok ? sst_kmer_keys_device(d_c, need, k, max_keys, d_k, out_count, sort, device) : SST_ERR_CUDA;
    if (rc == SST_OK && *out_count)
        rc = SST_CUDA_OK(cudaMemcpy(out_keys, d_k, *out_count * 4, cudaMemcpyDeviceToHost)) ? SST_OK : SST_ERR_CUDA;
    cudaFree(d_c); cudaFree(d_k);
    return rc;
}

}  // extern "C"
