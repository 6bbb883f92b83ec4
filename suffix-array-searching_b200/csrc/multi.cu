// multi.cu -- one index replica per GPU, query batch sharded contiguously.
//
// Mirrors the reference's only parallel harness: `threads` tasks over a shared index, task i
// takes qs[i*chunk .. (i+1)*chunk) with chunk = ceil(nq / threads)
// (static-search-tree/src/bin/bench.rs:558-573, src/util.rs:88-113).  Here a task is one
// long-lived host thread driving one device through its own streams and staging buffers; the
// shards need no exchange, so there is no collective and no NCCL.
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

typedef struct sst_multi_sa sst_multi_sa_t;

namespace {

// One worker per replica: owns the host thread whose thread-local streams / staging ring serve
// that device, so repeated sst_multi_query calls reuse them.
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = false, stop = false;

    Worker() {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(m);
            while (true) {
                cv.wait(lk, [this] { return has_job || stop; });
                if (stop) return;
                auto j = std::move(job);
                has_job = false;
                lk.unlock();
                j();
                lk.lock();
                done = true;
                cv.notify_all();
            }
        });
    }
    void post(std::function<void()> j) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(j);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return done; });
    }
    ~Worker() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
            cv.notify_all();
        }
        if (th.joinable()) th.join();
    }
};

}  // namespace

struct sst_multi {
    std::vector<sst_index_t*> replicas;
    std::vector<std::unique_ptr<Worker>> workers;
    std::mutex call_mutex;  // one sharded query at a time per multi handle
};

using namespace sst;

namespace {

template <class F>
sst_multi* build_replicas(const int* devices, int n_devices, F build_one) {
    clear_error();
    if (!devices || n_devices < 1) { set_error(SST_ERR_ARG, "need at least one device"); return nullptr; }
    auto* m = new sst_multi();
    m->replicas.assign(n_devices, nullptr);
    for (int i = 0; i < n_devices; i++) m->workers.emplace_back(new Worker());
    std::vector<std::string> errs(n_devices);
    std::vector<int> stats(n_devices, SST_OK);
    for (int i = 0; i < n_devices; i++)
        m->workers[i]->post([&, i] {
            if (!opt(OPT_NO_BIND)) (void)sst_bind_thread_to_device(devices[i]);  // worker stays on the GPU's socket
            m->replicas[i] = build_one(devices[i]);
            if (!m->replicas[i]) { errs[i] = sst_last_error(); stats[i] = sst_last_status(); }
        });
    for (int i = 0; i < n_devices; i++) m->workers[i]->wait();
    for (int i = 0; i < n_devices; i++)
        if (!m->replicas[i]) {
            const int st = stats[i];
            const std::string msg = "device " + std::to_string(devices[i]) + ": " + errs[i];
            sst_multi_free(m);
            set_error(st, msg);
            return nullptr;
        }
    return m;
}

}  // namespace

namespace {
// The sorted keys cross PCIe ONCE (to the first device); every other replica receives them device to device
// (cudaMemcpyPeerAsync: NVLink when the devices are peers) and runs the GPU layout builder on its own copy, all in parallel.
template <class F>
sst_multi* build_tree_replicas(const uint32_t* sorted, size_t n, const int* devices, int n_devices, F build_from_device) {
    clear_error();
    if (!sorted || n == 0) { set_error(SST_ERR_ARG, "empty input"); return nullptr; }
    if (!devices || n_devices < 1) { set_error(SST_ERR_ARG, "need at least one device"); return nullptr; }
    uint32_t* d0 = nullptr;
    {
        DeviceGuard g(devices[0]);
        if (!g.ok || !SST_CUDA_OK(cudaMalloc(&d0, n * 4)) || !SST_CUDA_OK(cudaMemcpy(d0, sorted, n * 4, cudaMemcpyHostToDevice))) { cudaFree(d0); return nullptr; }
    }
    const int dev0 = devices[0];
    sst_multi* m = build_replicas(devices, n_devices, [&](int dev) -> sst_index_t* {
        if (dev == dev0) return build_from_device(d0, dev);
        DeviceGuard g(dev);
        if (!g.ok) return nullptr;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dev, dev0) == cudaSuccess && can) (void)cudaDeviceEnablePeerAccess(dev0, 0);
        (void)cudaGetLastError();  // (already enabled / not supported: the copy below is then staged by the driver)
        uint32_t* d = nullptr;
        cudaStream_t st = thread_stream(dev);
        sst_index_t* r = nullptr;
        if (SST_CUDA_OK(cudaMalloc(&d, n * 4)) && SST_CUDA_OK(cudaMemcpyPeerAsync(d, dev, d0, dev0, n * 4, st)) && SST_CUDA_OK(cudaStreamSynchronize(st)))
            r = build_from_device(d, dev);
        cudaFree(d);
        return r;
    });
    {
        DeviceGuard g(dev0);
        cudaFree(d0);
    }
    return m;
}

}  // namespace

// ---- suffix-array replicas -----------------------------------------------------------------------
extern "C" void sst_multi_sa_free(struct sst_multi_sa* m);
struct sst_multi_sa {
    std::vector<sst_sa_t*> replicas;
    std::vector<std::unique_ptr<Worker>> workers;
    std::mutex call_mutex;
};

namespace {
// The index is BUILT once, on the first device (suffix array, pivot table, k-mer table, inlined bases), and copied to the
// others device to device in doubling rounds (round r: replicas [0, 2^r) feed replicas [2^r, 2^(r+1))), so that G replicas
// cost one build plus log2(G) rounds of copies at NVLink speed instead of G builds.
sst_multi_sa* build_sa_replicas(const int* devices, int n_devices, const std::function<sst_sa_t*(int)>& build_first) {
    clear_error();
    if (!devices || n_devices < 1) { set_error(SST_ERR_ARG, "need at least one device"); return nullptr; }
    auto* m = new sst_multi_sa();
    m->replicas.assign(n_devices, nullptr);
    for (int i = 0; i < n_devices; i++) m->workers.emplace_back(new Worker());
    std::vector<std::string> errs(n_devices);
    std::vector<int> stats(n_devices, SST_OK);
    auto run = [&](int i, std::function<sst_sa_t*()> f) {  // (f is copied into the worker's job)
        m->workers[i]->post([&, i, f] {
            if (!opt(OPT_NO_BIND)) (void)sst_bind_thread_to_device(devices[i]);
            m->replicas[i] = f();
            if (!m->replicas[i]) { errs[i] = sst_last_error(); stats[i] = sst_last_status(); }
        });
    };
    run(0, std::function<sst_sa_t*()>([&]() -> sst_sa_t* { return build_first(devices[0]); }));
    m->workers[0]->wait();
    bool ok = m->replicas[0] != nullptr;
    for (int have = 1; ok && have < n_devices; have *= 2) {
        const int end = std::min(n_devices, 2 * have);
        for (int i = have; i < end; i++) run(i, std::function<sst_sa_t*()>([&, i, have]() -> sst_sa_t* { return clone_sa(m->replicas[i - have], devices[i]); }));
        for (int i = have; i < end; i++) m->workers[i]->wait();
        for (int i = have; i < end; i++) ok = ok && m->replicas[i] != nullptr;
    }
    if (!ok) {
        int bad = 0;
        while (bad < n_devices - 1 && (m->replicas[bad] || stats[bad] == SST_OK)) bad++;
        const int st = stats[bad];
        const std::string msg = "device " + std::to_string(devices[bad]) + ": " + errs[bad];
        sst_multi_sa_free(m);
        set_error(st, msg);
        return nullptr;
    }
    return m;
}
}  // namespace

extern "C" {

sst_multi_t* sst_multi_stree_build(const uint32_t* sorted, size_t n, uint32_t node_b, uint32_t flags, const int* devices,
                                   int n_devices) {
    return build_tree_replicas(sorted, n, devices, n_devices, [&](const uint32_t* d, int dev) { return sst_stree_build_device(d, n, node_b, flags, dev); });
}

sst_multi_t* sst_multi_pstree_build(const uint32_t* sorted, size_t n, uint32_t b, int variant, const int* devices,
                                    int n_devices) {
    return build_tree_replicas(sorted, n, devices, n_devices, [&](const uint32_t* d, int dev) { return sst_pstree_build_device(d, n, b, variant, dev); });
}

sst_multi_sa_t* sst_multi_sa_build(const uint8_t* text, size_t n, const int* devices, int n_devices) {
    return build_sa_replicas(devices, n_devices, [&](int dev) { return sst_sa_build(text, n, dev); });
}

sst_multi_sa_t* sst_multi_sa_from_parts(const uint8_t* text, size_t n, const uint32_t* sa, const int* devices, int n_devices) {
    return build_sa_replicas(devices, n_devices, [&](int dev) { return sst_sa_from_parts(text, n, sa, dev); });
}

int sst_multi_sa_devices(const sst_multi_sa_t* m) { return m ? (int)m->replicas.size() : 0; }

// Patterns sharded contiguously: replica i takes patterns [i * chunk, (i + 1) * chunk) with chunk = ceil(npat / G), the
// rule of the reference's thread harness (static-search-tree/src/bin/bench.rs:558-573); the SA callers of the reference are
// serial loops (sa_search.rs:423-451).  Offsets stay absolute into `pats`.
int sst_multi_sa_search(const sst_multi_sa_t* cm, const uint8_t* pats, const uint64_t* pat_off, size_t npat, int mode, uint32_t* out_lo,
                        uint32_t* out_hi, uint32_t* out_pos) {
    clear_error();
    auto* m = const_cast<sst_multi_sa_t*>(cm);
    if (!m || m->replicas.empty()) { set_error(SST_ERR_ARG, "null multi index"); return SST_ERR_ARG; }
    if (npat && (!pat_off || !out_lo)) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    std::lock_guard<std::mutex> call(m->call_mutex);
    const size_t G = m->replicas.size();
    const size_t chunk = div_ceil(npat, G);
    std::vector<int> rc(G, SST_OK);
    std::vector<std::string> errs(G);
    for (size_t i = 0; i < G; i++)
        m->workers[i]->post([&, i] {
            const size_t s = std::min(npat, i * chunk), e = std::min(npat, (i + 1) * chunk);
            if (e > s) {
                rc[i] = sst_sa_search(m->replicas[i], pats, pat_off + s, e - s, mode, out_lo + s, out_hi ? out_hi + s : nullptr, out_pos ? out_pos + s : nullptr);
                if (rc[i] != SST_OK) errs[i] = sst_last_error();
            }
        });
    for (size_t i = 0; i < G; i++) m->workers[i]->wait();
    for (size_t i = 0; i < G; i++)
        if (rc[i] != SST_OK) { set_error(rc[i], errs[i]); return rc[i]; }
    return SST_OK;
}

void sst_multi_sa_free(sst_multi_sa_t* m) {
    if (!m) return;
    m->workers.clear();
    for (auto* r : m->replicas) sst_sa_free(r);
    delete m;
}

int sst_multi_devices(const sst_multi_t* m) { return m ? (int)m->replicas.size() : 0; }

int sst_multi_query(const sst_multi_t* cm, const uint32_t* qs, size_t nq, uint32_t* out_vals, uint64_t* out_idx, int scheme) {
    clear_error();
    auto* m = const_cast<sst_multi_t*>(cm);
    if (!m || m->replicas.empty()) { set_error(SST_ERR_ARG, "null multi index"); return SST_ERR_ARG; }
    std::lock_guard<std::mutex> call(m->call_mutex);
    const size_t G = m->replicas.size();
    const size_t chunk = div_ceil(nq, G);  // bench.rs:558
    std::vector<int> rc(G, SST_OK);
    std::vector<std::string> errs(G);
    for (size_t i = 0; i < G; i++)
        m->workers[i]->post([&, i] {
            const size_t s = std::min(nq, i * chunk), e = std::min(nq, (i + 1) * chunk);  // bench.rs:567-569
            if (e > s) {
                rc[i] = sst_query(m->replicas[i], qs + s, e - s, out_vals + s, out_idx ? out_idx + s : nullptr, scheme);
                if (rc[i] != SST_OK) errs[i] = sst_last_error();
            }
        });
    for (size_t i = 0; i < G; i++) m->workers[i]->wait();
    for (size_t i = 0; i < G; i++)
        if (rc[i] != SST_OK) { set_error(rc[i], errs[i]); return rc[i]; }
    return SST_OK;
}

// Device-resident shards: shard i (d_qs[i], nq[i] queries, outputs d_out_vals[i] / d_out_idx[i]) already lives on the device
// of replica i, so an in-process multi-GPU caller is not forced through PCIe.  Every replica's worker launches on its own
// stream and the call returns when all shards are done.
int sst_multi_query_device(const sst_multi_t* cm, const uint32_t* const* d_qs, const size_t* nq, uint32_t* const* d_out_vals,
                           uint64_t* const* d_out_idx, int scheme) {
    clear_error();
    auto* m = const_cast<sst_multi_t*>(cm);
    if (!m || m->replicas.empty() || !d_qs || !nq || !d_out_vals) { set_error(SST_ERR_ARG, "null argument"); return SST_ERR_ARG; }
    std::lock_guard<std::mutex> call(m->call_mutex);
    const size_t G = m->replicas.size();
    std::vector<int> rc(G, SST_OK);
    std::vector<std::string> errs(G);
    for (size_t i = 0; i < G; i++)
        m->workers[i]->post([&, i] {
            if (!nq[i]) return;
            const int dev = sst_index_device(m->replicas[i]);
            DeviceGuard g(dev);
            cudaStream_t st = thread_stream(dev);
            rc[i] = g.ok && st ? sst_query_device(m->replicas[i], d_qs[i], nq[i], d_out_vals[i], d_out_idx ? d_out_idx[i] : nullptr, scheme, st) : SST_ERR_CUDA;
            if (rc[i] == SST_OK && !SST_CUDA_OK(cudaStreamSynchronize(st))) rc[i] = SST_ERR_CUDA;
            if (rc[i] != SST_OK) errs[i] = sst_last_error();
        });
    for (size_t i = 0; i < G; i++) m->workers[i]->wait();
    for (size_t i = 0; i < G; i++)
        if (rc[i] != SST_OK) { set_error(rc[i], errs[i]); return rc[i]; }
    return SST_OK;
}

void sst_multi_free(sst_multi_t* m) {
    if (!m) return;
    m->workers.clear();  // joins the worker threads (their thread-local staging is released with them)
    for (auto* r : m->replicas) sst_index_free(r);
    delete m;
}

}  // extern "C"
