// Device group: single-process multi-GPU behind the C ABI ("bartint_init(n_gpus)").
//
// The reference's callers are ONE R process (src/BARTInt.R:259-262, 312-314; src/survey_design/bartMean.R:74-82), so
// the GPUs of a box have to be driven from inside the library: bartint_init(n) opens a group over n devices
// (peer access, one worker thread + stream per member, an NCCL communicator per member from ncclCommInitAll), and
// the entry points that take a matrix staged on the group (bartint_stage_matrix with device = BARTINT_DEVICE_GROUP)
// shard its rows over the members.  The path shards without a data-path collective except one exchange per call:
// the members' per-draw partial sums (S doubles each) are all-gathered with ncclAllGather on the compute streams and
// added in member order, so the k-GPU result does not depend on NCCL's reduction order.
//
// Member 0 is driven by the calling thread exactly like the single-GPU path.  Every other member has a worker
// thread that owns its device context and runs posted closures in order: the CUDA runtime calls of different members
// (tens of microseconds each) then run on different host threads instead of queueing behind one another.
// NCCL is loaded with dlopen: a box without it can still use one GPU.
#include <dlfcn.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace bartint {
const char* get_error();

namespace {

// the handful of NCCL declarations used (ABI-stable since NCCL 2.x; nccl.h need not be present to build)
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclDouble = 8;
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString;
    }
};

struct Worker {
    int device = 0;
    cudaStream_t st = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv, cv_idle;
    std::deque<std::function<int()>> q;
    bool busy = false, stop = false;
    int rc = BARTINT_OK;
    std::string err;

    void loop() {
        cudaSetDevice(device);
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        {
            std::lock_guard<std::mutex> lk(mu);
            busy = false;
        }
        cv_idle.notify_all();
        for (;;) {
            std::function<int()> fn;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) break;          // stop requested and nothing left
                fn = std::move(q.front());
                q.pop_front();
                busy = true;
            }
            const int r = fn();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (r != BARTINT_OK && rc == BARTINT_OK) { rc = r; err = get_error(); }
                busy = false;
            }
            cv_idle.notify_all();
        }
        if (st) cudaStreamDestroy(st);
    }
};

struct Group {
    std::vector<int> devices;                       // member -> CUDA device
    std::vector<std::unique_ptr<Worker>> workers;   // [0] unused (member 0 = the calling thread)
    std::vector<ncclComm_t> comms;
    bool nccl = false;
    bool peer_to_root = true;                       // every member can store into member 0's memory
};

std::mutex g_mu;
Group* g_group = nullptr;
int g_generation = 0;
NcclApi g_nccl;

void destroy_group(Group* g) {
    if (!g) return;
    for (auto& w : g->workers) {
        if (!w) continue;
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->stop = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    if (g->nccl)
        for (ncclComm_t c : g->comms)
            if (c) g_nccl.CommDestroy(c);
    delete g;
}

}  // namespace

int group_size() { return g_group ? (int)g_group->devices.size() : 1; }
int group_generation() { return g_generation; }
int group_device(int g) { return g_group ? g_group->devices[(size_t)g] : 0; }
bool group_has(int device) { return g_group != nullptr && g_group->devices[0] == device; }
cudaStream_t group_stream(int g) { return g_group && g >= 1 ? g_group->workers[(size_t)g]->st : nullptr; }
// The members can store into member 0's memory (NVLink peer access, or the same device in test mode) and the NCCL
// exchange has not been asked for: the design step's gather-to-root is then done with peer writes.
bool group_peer_reads() { return g_group != nullptr && g_group->peer_to_root; }
bool group_peer_writes() {
    const char* e = getenv("BARTINT_GROUP_EXCHANGE");          // read per call: the tests switch it
    const bool want_nccl = e && strcmp(e, "nccl") == 0;
    return g_group != nullptr && g_group->peer_to_root && !(want_nccl && g_group->nccl);
}
const char* group_exchange_name() {
    return !g_group || g_group->devices.size() < 2 ? "none (one device)"
           : group_peer_writes() ? "peer writes of the per-draw partial sums into member 0's buffer (NVLink; ncclAllGather with "
                                   "BARTINT_GROUP_EXCHANGE=nccl and in bartint_eval) + member-ordered sum"
           : g_group->nccl ? "ncclAllGather of the per-draw partial sums + member-ordered sum"
                           : "peer copies of the per-draw partial sums + member-ordered sum (same-device test mode)";
}

void group_post(int g, std::function<int()> fn) {
    Worker* w = g_group->workers[(size_t)g].get();
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->q.push_back(std::move(fn));
    }
    w->cv.notify_one();
}

int group_wait() {
    if (!g_group) return BARTINT_OK;
    int rc = BARTINT_OK;
    for (auto& w : g_group->workers) {
        if (!w) continue;
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv_idle.wait(lk, [&] { return w->q.empty() && !w->busy; });
        if (w->rc != BARTINT_OK && rc == BARTINT_OK) {
            rc = w->rc;
            set_error("group member on device %d: %s", w->device, w->err.c_str());
        }
        w->rc = BARTINT_OK;
        w->err.clear();
    }
    return rc;
}

int group_allgather(double* const* send, double* const* recv, size_t count, cudaStream_t const* streams) {
    Group* g = g_group;
    const int G = group_size();
    if (g && g->nccl) {
        ncclResult_t r = g_nccl.GroupStart();
        for (int k = 0; k < G && r == 0; ++k)
            r = g_nccl.AllGather(send[k], recv[k], count, kNcclDouble, g->comms[(size_t)k], streams[k]);
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == 0) r = r2;
        BI_REQUIRE(r == 0, BARTINT_ERR_CUDA, "ncclAllGather: %s", g_nccl.GetErrorString(r));
        return BARTINT_OK;
    }
    // same-device test mode (BARTINT_GROUP_DEVICES lists one GPU several times; NCCL refuses duplicate devices):
    // every member's row is copied to member 0's buffer on the member's own stream, member 0's stream waits for them
    int cur = 0;
    BI_CUDA(cudaGetDevice(&cur));
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < G && e == cudaSuccess; ++k) {
        const int dev = group_device(k);
        e = cudaSetDevice(dev);
        if (e == cudaSuccess)
            e = cudaMemcpyPeerAsync(recv[0] + (size_t)k * count, group_device(0), send[k], dev, sizeof(double) * count, streams[k]);
        if (e == cudaSuccess && k > 0) {
            cudaEvent_t ev = nullptr;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(ev, streams[k]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(streams[0], ev, 0);
            if (ev) cudaEventDestroy(ev);
        }
    }
    cudaSetDevice(cur);
    BI_CUDA(e);
    return BARTINT_OK;
}

}  // namespace bartint

using namespace bartint;

extern "C" {

int bartint_init(int n_gpus) {
    std::lock_guard<std::mutex> lk(g_mu);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available: libbartint_cuda has no CPU path");
        return BARTINT_ERR_NODEVICE;
    }
    std::vector<int> devices;
    if (const char* env = getenv("BARTINT_GROUP_DEVICES")) {       // e.g. "0,0": two members on one GPU (tests)
        for (const char* p = env; *p;) {
            char* end = nullptr;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            devices.push_back((int)v);
            p = *end ? end + 1 : end;
        }
        if (n_gpus > 0 && (int)devices.size() > n_gpus) devices.resize((size_t)n_gpus);
    } else {
        const int n = n_gpus <= 0 ? n_dev : n_gpus;
        for (int k = 0; k < n; ++k) devices.push_back(k);
    }
    BI_REQUIRE(!devices.empty() && devices.size() <= 64, BARTINT_ERR_ARG, "bad device group size %zu", devices.size());
    bool distinct = true;
    for (size_t a = 0; a < devices.size(); ++a) {
        BI_REQUIRE(devices[a] >= 0 && devices[a] < n_dev, BARTINT_ERR_ARG, "device %d of the group does not exist (%d visible)",
                   devices[a], n_dev);
        for (size_t b = 0; b < a; ++b) distinct = distinct && devices[a] != devices[b];
    }
    if (g_group) {
        group_wait();
        destroy_group(g_group);
        g_group = nullptr;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    std::unique_ptr<Group> g(new Group());
    g->devices = devices;
    const int G = (int)devices.size();
    // peer access both ways (direct NVLink copies of the forest images), also for the stream-ordered allocator's pools
    for (int a = 0; a < G; ++a) {
        cudaSetDevice(devices[(size_t)a]);
        cudaMemPool_t pool = nullptr;
        cudaDeviceGetDefaultMemPool(&pool, devices[(size_t)a]);
        uint64_t keep = UINT64_MAX;
        if (pool) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        for (int b = 0; b < G; ++b) {
            if (devices[(size_t)a] == devices[(size_t)b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[(size_t)a], devices[(size_t)b]);
            if (!can) { if (b == 0) g->peer_to_root = false; continue; }
            cudaError_t pe = cudaDeviceEnablePeerAccess(devices[(size_t)b], 0);
            if (pe != cudaSuccess) (void)cudaGetLastError();                   // already enabled
            if (pool) {
                cudaMemAccessDesc desc{};
                desc.location.type = cudaMemLocationTypeDevice;
                desc.location.id = devices[(size_t)b];
                desc.flags = cudaMemAccessFlagsProtReadWrite;
                if (cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) (void)cudaGetLastError();
            }
        }
    }
    cudaSetDevice(cur);
    g->workers.resize((size_t)G);
    for (int k = 1; k < G; ++k) {
        g->workers[(size_t)k].reset(new Worker());
        Worker* w = g->workers[(size_t)k].get();
        w->device = devices[(size_t)k];
        w->busy = true;                      // until its stream exists
        w->th = std::thread([w] { w->loop(); });
    }
    for (int k = 1; k < G; ++k) {
        Worker* w = g->workers[(size_t)k].get();
        std::unique_lock<std::mutex> wl(w->mu);
        w->cv_idle.wait(wl, [&] { return !w->busy; });
    }
    if (G > 1 && distinct) {
        if (!g_nccl.load()) {
            destroy_group(g.release());
            set_error("bartint_init(%d): NCCL (libnccl.so.2) could not be loaded: %s", G, dlerror());
            return BARTINT_ERR_UNSUPPORTED;
        }
        g->comms.assign((size_t)G, nullptr);
        const ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), G, devices.data());
        if (r != 0) {
            destroy_group(g.release());
            set_error("ncclCommInitAll over %d devices: %s", G, g_nccl.GetErrorString(r));
            return BARTINT_ERR_CUDA;
        }
        g->nccl = true;
    }
    cudaSetDevice(devices[0]);
    g_group = g.release();
    ++g_generation;
#ifdef _OPENMP
    // the members' worker threads need cores of their own: the exporter's OpenMP team spin-waits between its parallel
    // regions, and a team as large as the machine leaves the workers waiting for a scheduler quantum (milliseconds)
    if (G > 1) omp_set_num_threads(std::max(1, std::min(omp_get_max_threads(), omp_get_num_procs() - (G - 1))));
#endif
    return BARTINT_OK;
}

int bartint_group_size(void) { return group_size(); }

const char* bartint_group_exchange(void) { return group_exchange_name(); }

void bartint_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_group) return;
    group_wait();
    destroy_group(g_group);
    g_group = nullptr;
    ++g_generation;
}

}  // extern "C"
