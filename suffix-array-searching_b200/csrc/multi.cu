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

#include <cstdlib>

#include "common.cuh"

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
            if (!getenv("SST_NO_BIND")) (void)sst_bind_thread_to_device(devices[i]);  // worker stays on the GPU's socket
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

extern "C" {

sst_multi_t* sst_multi_stree_build(const uint32_t* sorted, size_t n, uint32_t node_b, uint32_t flags, const int* devices,
                                   int n_devices) {
    return build_replicas(devices, n_devices, [&](int dev) { return sst_stree_build(sorted, n, node_b, flags, dev); });
}

sst_multi_t* sst_multi_pstree_build(const uint32_t* sorted, size_t n, uint32_t b, int variant, const int* devices,
                                    int n_devices) {
    return build_replicas(devices, n_devices, [&](int dev) { return sst_pstree_build(sorted, n, b, variant, dev); });
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

void sst_multi_free(sst_multi_t* m) {
    if (!m) return;
    m->workers.clear();  // joins the worker threads (their thread-local staging is released with them)
    for (auto* r : m->replicas) sst_index_free(r);
    delete m;
}

}  // extern "C"
