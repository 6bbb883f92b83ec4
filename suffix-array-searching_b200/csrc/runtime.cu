// runtime.cu -- error state, device selection and per-thread streams.
#include <atomic>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <sched.h>

#include "common.cuh"

namespace sst {

namespace {
thread_local std::string g_err;
thread_local int g_status = SST_OK;
constexpr int kMaxDevices = 64;
thread_local cudaStream_t g_streams[kMaxDevices][3] = {};

// ---- options table (see SST_OPTION_LIST in common.cuh) ----
struct OptDesc { const char* name; long long dflt, lo, hi; };
const OptDesc kOptDesc[OPT_COUNT] = {
#define SST_OPT_DESC(name, dflt, lo, hi) {#name, (long long)(dflt), (long long)(lo), (long long)(hi)},
    SST_OPTION_LIST(SST_OPT_DESC)
#undef SST_OPT_DESC
};
std::atomic<long long> g_opt[OPT_COUNT];
long long g_opt_initial[OPT_COUNT];  // default or the environment's value at load time (what sst_reset_options restores)
long long clamp_opt(int o, long long v) { return v < kOptDesc[o].lo ? kOptDesc[o].lo : v > kOptDesc[o].hi ? kOptDesc[o].hi : v; }
// Runs once when the library is loaded (single-threaded): the only getenv calls of the library.
struct OptInit {
    OptInit() {
        for (int o = 0; o < OPT_COUNT; o++) {
            long long v = kOptDesc[o].dflt;
            char env[64];
            snprintf(env, sizeof env, "SST_%s", kOptDesc[o].name);
            if (const char* e = getenv(env); e && *e) v = clamp_opt(o, strtoll(e, nullptr, 10));
            g_opt_initial[o] = v;
            g_opt[o].store(v, std::memory_order_relaxed);
        }
    }
} g_opt_init;
int find_opt(const char* name) {
    if (!name) return -1;
    if (!strncmp(name, "SST_", 4)) name += 4;
    for (int o = 0; o < OPT_COUNT; o++)
        if (!strcmp(name, kOptDesc[o].name)) return o;
    return -1;
}
}  // namespace

long long opt(Opt o) { return g_opt[o].load(std::memory_order_relaxed); }

void set_error(int status, const std::string& msg) {
    g_status = status;
    g_err = msg;
}
void clear_error() {
    g_status = SST_OK;
    g_err.clear();
}

bool cuda_ok(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return true;
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    set_error(SST_ERR_CUDA, buf);
    (void)cudaGetLastError();
    return false;
}

DeviceGuard::DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = SST_CUDA_OK(cudaSetDevice(device));
}
DeviceGuard::~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
}

static cudaStream_t get_stream(int device, int which) {
    if (device < 0 || device >= kMaxDevices) return nullptr;
    cudaStream_t& s = g_streams[device][which];
    if (!s) {
        DeviceGuard g(device);
        if (!g.ok) return nullptr;
        if (!SST_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking))) s = nullptr;
    }
    return s;
}
cudaStream_t thread_stream(int device) { return get_stream(device, 0); }
cudaStream_t thread_copy_stream(int device, int which) { return get_stream(device, 1 + (which & 1)); }

// L2 fetch granularity (cudaLimitMaxL2FetchGranularity).  The hot path reads isolated 64-byte
// nodes; with the default 128-byte granularity every leaf miss moves twice the bytes it needs
// (measured: dram__bytes_read = 2x algorithmic, profiles/).  SST_L2_FETCH overrides (0 = leave).
void configure_l2_fetch(int device) {
    static thread_local bool done[kMaxDevices] = {};
    if (device < 0 || device >= kMaxDevices || done[device]) return;
    done[device] = true;
    const int want = (int)opt(OPT_L2_FETCH);
    if (want <= 0) return;
    size_t before = 0, after = 0;
    cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
    const cudaError_t rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)want);
    cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
    if (rc != cudaSuccess) (void)cudaGetLastError();
    if (opt(OPT_DEBUG)) fprintf(stderr, "[sst] L2 fetch granularity: before=%zu want=%d rc=%d after=%zu\n", before, want, (int)rc, after);
}

int sm_count(int device) {
    static std::atomic<int> cache[kMaxDevices];
    if (device >= 0 && device < kMaxDevices)
        if (const int c = cache[device].load(std::memory_order_relaxed)) return c;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    if (device >= 0 && device < kMaxDevices) cache[device].store(v, std::memory_order_relaxed);
    return v;
}
unsigned cur_sms() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
    const int v = sm_count(d);
    return v > 0 ? (unsigned)v : 1u;
}
size_t max_smem_optin(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
    return (size_t)v;
}
bool device_usable(int device) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return major == 10;  // the library is built for sm_100a only
}

}  // namespace sst

extern "C" {

const char* sst_last_error(void) { return sst::g_err.c_str(); }
int sst_last_status(void) { return sst::g_status; }
const char* sst_version(void) { return "sst_b200 0.2 (sm_100a)"; }

// Options: `name` with or without the SST_ prefix.  Values outside the option's range are rejected (SST_ERR_ARG), so a
// typo cannot turn every launch into a failure.  Takes effect for calls (and index builds) that start afterwards.
int sst_set_option(const char* name, long long value) {
    sst::clear_error();
    const int o = sst::find_opt(name);
    if (o < 0) { sst::set_error(SST_ERR_ARG, std::string("unknown option ") + (name ? name : "(null)")); return SST_ERR_ARG; }
    if (value < sst::kOptDesc[o].lo || value > sst::kOptDesc[o].hi) {
        sst::set_error(SST_ERR_ARG, std::string("option ") + sst::kOptDesc[o].name + " out of range [" + std::to_string(sst::kOptDesc[o].lo) + ", " +
                                        std::to_string(sst::kOptDesc[o].hi) + "]");
        return SST_ERR_ARG;
    }
    sst::g_opt[o].store(value, std::memory_order_relaxed);
    return SST_OK;
}
int sst_get_option(const char* name, long long* out_value) {
    sst::clear_error();
    const int o = sst::find_opt(name);
    if (o < 0 || !out_value) { sst::set_error(SST_ERR_ARG, "unknown option or null output"); return SST_ERR_ARG; }
    *out_value = sst::g_opt[o].load(std::memory_order_relaxed);
    return SST_OK;
}
void sst_reset_options(void) {
    for (int o = 0; o < sst::OPT_COUNT; o++) sst::g_opt[o].store(sst::g_opt_initial[o], std::memory_order_relaxed);
}
int sst_option_count(void) { return sst::OPT_COUNT; }
const char* sst_option_name(int i) { return i >= 0 && i < sst::OPT_COUNT ? sst::kOptDesc[i].name : nullptr; }

// Page-locked host buffers for full-speed sst_query / sst_sa_search (PCIe DMA without a staging copy).
void* sst_host_alloc(size_t bytes) {
    sst::clear_error();
    void* p = nullptr;
    if (!SST_CUDA_OK(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable))) return nullptr;
    return p;
}
void sst_host_free(void* p) {
    if (p) (void)cudaFreeHost(p);
}

// Host-side locality: on a multi-socket box a page-locked buffer that lives on the other socket's memory makes every
// H2D/D2H cross the socket interconnect (measured: 8 ranks, 0.8 GB per rank and step: 16.8 Gq/s in total, 2.1 per GPU,
// against 10.5 for one GPU alone).  Binds the calling thread to the CPUs that are local to `device`
// (/sys/bus/pci/devices/<bus id>/local_cpulist); memory the thread allocates afterwards (first touch, cudaHostAlloc) is
// then local too.  Returns the number of CPUs in the set, 0 if the topology is not visible (nothing changed), -1 on error.
int sst_bind_thread_to_device(int device) {
    sst::clear_error();
    char bus[32] = {0};
    if (!SST_CUDA_OK(cudaDeviceGetPCIBusId(bus, sizeof(bus), device))) return -1;
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return 0;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    int count = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {  // "0-31,64-95"
        int a = -1, b = -1;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) {}
        else if (sscanf(tok, "%d", &a) == 1) b = a;
        for (int c = a; a >= 0 && c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); count++; }
    }
    if (count == 0) return 0;
    cpu_set_t cur;
    if (sched_getaffinity(0, sizeof(cur), &cur) == 0) {  // stay inside the set the process was given (cgroups, taskset)
        cpu_set_t both;
        CPU_AND(&both, &set, &cur);
        if (CPU_COUNT(&both) == 0) return 0;
        set = both;
        count = CPU_COUNT(&both);
    }
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return 0;
    return count;
}

int sst_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < n; d++) usable += sst::device_usable(d) ? 1 : 0;
    return usable;
}

}  // extern "C"
