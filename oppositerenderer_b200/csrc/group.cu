// group.cu — multi-GPU behind the C ABI: iteration partition over the GPUs of one box and ONE NCCL reduce of the running sums.
//
// Replaces the reference's distribution layer: the client's hand-out loop (Client/DistributedApplication.cpp:96-122 — the next
// global iteration number and its PPM radius go to whichever render server is free), the servers' render loop and result
// packets (Server/server/RenderServerRenderer.cpp:154-175) and the client-side sum of the returned W*H float3 buffers
// (Gui/gui/RenderWidget.cpp:120,197).  Here global iteration i belongs to rank i % G as its localIterationNumber i / G; its
// radius is the closed form of the driver's recurrence (opr_ppm_radius_for_iteration), so nothing but the 128-byte NCCL id is
// exchanged before the reduce.  Two shapes of the same thing:
//   (a) one process per GPU (torchrun / MPI):  opr_comm_get_unique_id -> (launcher transports 128 bytes) -> opr_comm_init_rank
//       -> opr_reduce_output;
//   (b) one process, G GPUs:  opr_group_* — one renderer and one host thread per device, ncclCommInitAll.
// The reduce never targets a running sum: it lands in a scratch buffer on the root, so it can be repeated during a progressive
// render without counting earlier sums twice.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy already mapped by the host application — e.g. PyTorch's — or
// the system one), so the renderer itself has no link-time dependency on it.
#include "renderer.h"
#include "../../include/opr.h"
#include <dlfcn.h>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

using namespace opr;

namespace {

// ---- the few NCCL entry points used, with the types of nccl.h (2.x ABI) ---------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // ncclSuccess = 0
enum { ncclFloat32 = 7, ncclSum = 0 };    // ncclDataType_t / ncclRedOp_t values of nccl.h

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

Nccl& nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.lib) break; }
        if (!n.lib) { n.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return; }
        auto sym = [&](const char* s) { void* p = dlsym(n.lib, s); if (!p && n.error.empty()) n.error = std::string("NCCL symbol missing: ") + s; return p; };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.Reduce = (decltype(n.Reduce))sym("ncclReduce");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    });
    if (!n.error.empty()) throw std::runtime_error(n.error);
    return n;
}

void ncclCheck(ncclResult_t r, const char* what)
{
    if (r != 0) throw std::runtime_error(std::string(what) + ": " + nccl().GetErrorString(r));
}

constexpr unsigned SEED_STRIDE_LOG2 = 28;    // 2^28 RNG states per rank, 16 ranks: disjoint `seed + i` ranges modulo 2^32
constexpr int MAX_RANKS = 1 << (32 - SEED_STRIDE_LOG2);

// scratch buffer the reduce lands in (root) — sized for the current image
float* reducedBuffer(Renderer& R)
{
    const size_t n = (size_t)R.width * R.height * 3;
    if (R.reducedCount != n) {
        if (R.reduced) devFree(R.reduced);
        R.reduced = nullptr; R.reducedCount = 0;
        devAlloc((void**)&R.reduced, n * sizeof(float));
        R.reducedCount = n;
    }
    return R.reduced;
}

// queues ncclReduce(sum) of R.output on R.stream, behind everything rendered so far
void queueReduce(Renderer& R, int root)
{
    if (!R.buffersAllocated) throw std::runtime_error("opr_reduce_output before the first iteration");
    if (!R.ncclComm) throw std::runtime_error("opr_reduce_output: no communicator (opr_comm_init_rank / opr_group_create first)");
    if ((size_t)R.rsW * R.rsH > ((size_t)1 << SEED_STRIDE_LOG2))
        throw std::runtime_error("RNG state buffer exceeds the per-rank seed stride: ranks would share random streams");
    OPR_CUDA_CHECK(cudaSetDevice(R.device));
    R.joinStreams();
    const size_t n = (size_t)R.width * R.height * 3;
    float* dst = R.commRank == root ? reducedBuffer(R) : nullptr;
    ncclCheck(nccl().Reduce(R.output, dst, n, ncclFloat32, ncclSum, root, (ncclComm_t)R.ncclComm, R.stream), "ncclReduce");
}

}  // namespace

// One worker thread per renderer: CUDA contexts, streams and the renderer's host state stay on one thread each, as in the
// reference (one RenderServerRenderer thread per server, Server/server/RenderServerRenderer.cpp:56-80).
struct opr_group {
    struct Worker {
        std::unique_ptr<opr_renderer> r;
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        std::function<void()> job;
        bool busy = false, quit = false;
        std::string error;
    };
    std::vector<std::unique_ptr<Worker>> w;

    void start(Worker& k)
    {
        k.th = std::thread([&k] {
            for (;;) {
                std::function<void()> job;
                {
                    std::unique_lock<std::mutex> l(k.m);
                    k.cv.wait(l, [&] { return k.quit || (k.busy && k.job); });
                    if (k.quit) return;
                    job = std::move(k.job); k.job = nullptr;
                }
                try { job(); } catch (const std::exception& e) { k.error = e.what(); } catch (...) { k.error = "unknown error"; }
                { std::lock_guard<std::mutex> l(k.m); k.busy = false; }
                k.cv.notify_all();
            }
        });
    }
    // runs f(rank, renderer) on every worker in parallel; throws the first failure
    void all(const std::function<void(int, Renderer&)>& f)
    {
        for (size_t i = 0; i < w.size(); i++) {
            Worker& k = *w[i];
            std::lock_guard<std::mutex> l(k.m);
            k.error.clear(); k.busy = true;
            k.job = [&f, &k, i] { f((int)i, k.r->r); };
            k.cv.notify_all();
        }
        std::string err;
        for (size_t i = 0; i < w.size(); i++) {
            Worker& k = *w[i];
            std::unique_lock<std::mutex> l(k.m);
            k.cv.wait(l, [&] { return !k.busy; });
            if (err.empty() && !k.error.empty()) err = "rank " + std::to_string(i) + ": " + k.error;
        }
        if (!err.empty()) throw std::runtime_error(err);
    }
    ~opr_group()
    {
        for (auto& k : w) {
            { std::lock_guard<std::mutex> l(k->m); k->quit = true; }
            k->cv.notify_all();
            if (k->th.joinable()) k->th.join();
        }
        for (auto& k : w) if (k->r && k->r->r.ncclComm) {
            try { cudaSetDevice(k->r->r.device); nccl().CommDestroy((ncclComm_t)k->r->r.ncclComm); } catch (...) {}
            k->r->r.ncclComm = nullptr;
        }
    }
};

extern "C" {

void opr_set_last_error(const char* msg);   // capi.cu
#define OPR_GTRY(...) try { __VA_ARGS__; return 0; } catch (const std::exception& e) { opr_set_last_error(e.what()); return 1; } catch (...) { opr_set_last_error("unknown error"); return 1; }

uint32_t opr_group_rank_seed(uint32_t baseSeed, int rank)
{
    return baseSeed + ((uint32_t)(rank & (MAX_RANKS - 1)) << SEED_STRIDE_LOG2);
}

int opr_comm_get_unique_id(void* id128)
{
    OPR_GTRY({
        ncclUniqueId id;
        ncclCheck(nccl().GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(id128, &id, sizeof id);
    })
}

int opr_comm_init_rank(opr_renderer* r, const void* id128, int rank, int world)
{
    OPR_GTRY({
        Renderer& R = r->r;
        if (!R.initialized) throw std::runtime_error("opr_comm_init_rank before opr_initialize");
        if (R.ncclComm) throw std::runtime_error("opr_comm_init_rank: this renderer already has a communicator");
        if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world) throw std::runtime_error("opr_comm_init_rank: rank / world out of range (at most 16 ranks)");
        OPR_CUDA_CHECK(cudaSetDevice(R.device));
        ncclUniqueId id; memcpy(&id, id128, sizeof id);
        ncclComm_t c = nullptr;
        ncclCheck(nccl().CommInitRank(&c, world, id, rank), "ncclCommInitRank");
        R.ncclComm = c; R.commRank = rank; R.commWorld = world;
    })
}

int opr_comm_destroy(opr_renderer* r)
{
    OPR_GTRY({
        Renderer& R = r->r;
        if (!R.ncclComm) return 0;
        OPR_CUDA_CHECK(cudaSetDevice(R.device));
        OPR_CUDA_CHECK(cudaStreamSynchronize(R.stream));
        ncclCheck(nccl().CommDestroy((ncclComm_t)R.ncclComm), "ncclCommDestroy");
        R.ncclComm = nullptr; R.commRank = 0; R.commWorld = 1;
    })
}

int opr_reduce_output(opr_renderer* r, int root, void* dstHost)
{
    OPR_GTRY({
        Renderer& R = r->r;
        queueReduce(R, root);
        if (dstHost && R.commRank == root) {
            OPR_CUDA_CHECK(cudaMemcpyAsync(dstHost, R.reduced, R.reducedCount * sizeof(float), cudaMemcpyDeviceToHost, R.stream));
            OPR_CUDA_CHECK(cudaStreamSynchronize(R.stream));
        }
    })
}

void* opr_get_reduced_device_ptr(opr_renderer* r) { return r->r.reduced; }

int opr_group_create(opr_group** out, int numDevices, const int* ordinals, const opr_config* cfg)
{
    OPR_GTRY({
        if (numDevices < 1 || numDevices > MAX_RANKS) throw std::runtime_error("opr_group_create: 1..16 devices");
        std::unique_ptr<opr_group> g(new opr_group());
        const uint32_t baseSeed = (cfg && cfg->seedIsSet) ? cfg->randomSeed : Renderer().seed;
        std::vector<int> devs(numDevices);
        for (int i = 0; i < numDevices; i++) {
            devs[i] = ordinals ? ordinals[i] : i;
            std::unique_ptr<opr_group::Worker> k(new opr_group::Worker());
            opr_config c{};
            if (cfg) c = *cfg;
            c.seedIsSet = 1; c.randomSeed = opr_group_rank_seed(baseSeed, i);
            opr_renderer* r = nullptr;
            if (opr_create(&r, &c)) throw std::runtime_error(opr_last_error());
            k->r.reset(r);
            g->w.push_back(std::move(k));
        }
        for (auto& k : g->w) g->start(*k);
        g->all([&](int i, Renderer& R) { R.initialize(devs[i]); });
        if (numDevices > 1) {
            std::vector<ncclComm_t> comms(numDevices);
            ncclCheck(nccl().CommInitAll(comms.data(), numDevices, devs.data()), "ncclCommInitAll");
            for (int i = 0; i < numDevices; i++) { Renderer& R = g->w[i]->r->r; R.ncclComm = comms[i]; R.commRank = i; R.commWorld = numDevices; }
        }
        *out = g.release();
    })
}

void opr_group_destroy(opr_group* g) { delete g; }
int opr_group_size(opr_group* g) { return (int)g->w.size(); }
opr_renderer* opr_group_renderer(opr_group* g, int rank) { return (rank >= 0 && rank < (int)g->w.size()) ? g->w[rank]->r.get() : nullptr; }

int opr_group_init_scene(opr_group* g, const opr_scene_desc* scene)
{
    OPR_GTRY(g->all([&](int, Renderer& R) { R.initScene(scene); }))
}

int opr_group_render_next_iterations(opr_group* g, uint64_t first, uint64_t count, double initialRadius, double alpha, const opr_request_details* details)
{
    OPR_GTRY({
        const uint64_t G = g->w.size();
        g->all([&](int rank, Renderer& R) {
            for (uint64_t i = first; i < first + count; i++) {
                if (i % G != (uint64_t)rank) continue;
                R.renderNextIteration(i, i / G, (float)opr_ppm_radius_for_iteration(initialRadius, i, alpha), true, *details);
            }
        });
    })
}

int opr_group_reduce_output(opr_group* g, void* dstHost)
{
    OPR_GTRY({
        const size_t G = g->w.size();
        Renderer& R0 = g->w[0]->r->r;
        if (G == 1) { R0.getOutputBuffer(dstHost); return 0; }
        // single-process multi-GPU: the per-device calls of one collective are bracketed by ncclGroupStart / End on one thread
        ncclCheck(nccl().GroupStart(), "ncclGroupStart");
        std::string err;
        for (size_t i = 0; i < G; i++) {
            try { queueReduce(g->w[i]->r->r, 0); } catch (const std::exception& e) { if (err.empty()) err = e.what(); }
        }
        ncclCheck(nccl().GroupEnd(), "ncclGroupEnd");
        if (!err.empty()) throw std::runtime_error(err);
        OPR_CUDA_CHECK(cudaSetDevice(R0.device));
        OPR_CUDA_CHECK(cudaMemcpyAsync(dstHost, R0.reduced, R0.reducedCount * sizeof(float), cudaMemcpyDeviceToHost, R0.stream));
        OPR_CUDA_CHECK(cudaStreamSynchronize(R0.stream));
        for (size_t i = 1; i < G; i++) { Renderer& R = g->w[i]->r->r; OPR_CUDA_CHECK(cudaSetDevice(R.device)); OPR_CUDA_CHECK(cudaStreamSynchronize(R.stream)); }
    })
}

}  // extern "C"
