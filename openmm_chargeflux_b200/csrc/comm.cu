// comm.cu -- multi-GPU plumbing of the C ABI: an evaluation is split over the ranks of an NCCL
// communicator (k-vector rows + direct-space i-clusters per rank, kspace.cu / direct.cu) and completed by ONE sum
// all-reduce of the int64 fixed-point reduction buffer [3*Npad + 8] over NVLink. Integer sums make the result
// independent of the reduction order. The reference has no multi-GPU path (its CUDA platform binds contexts[0] only,
// platforms/cuda/src/CudaCoulKernelFactory.cpp:40; CudaCoulKernels.cpp:477-481 splits nothing but exclusion tiles).
//
// Two ways in:
//   one process (or thread) per GPU   cfx_create(shard_rank, shard_count) + cfx_comm_init(id) on every rank; then
//                                     cfx_execute (host buffers) / cfx_execute_sharded (device buffers) on every rank
//   one process, several GPUs         cfx_multi_create(devices) -- what a plugin inside one OpenMM process can use
//
// NCCL is loaded with dlopen (libnccl.so.2: the copy already in the process, e.g. PyTorch's, else the system one), so
// libcfx_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include "cfx_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace cfx {

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    const char* names[] = {getenv("CFX_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) throw std::runtime_error(std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
    auto sym = [&](const char* name) {
        void* p = dlsym(api.lib, name);
        if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + name);
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    return api;
}

void ncclCheck(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(r));
}

// the first collective of a communicator allocates and connects: do it eagerly, outside any stream capture
void warmUp(State& st) {
    CFX_CUDA(cudaSetDevice(st.device));
    CFX_CUDA(cudaMemsetAsync(st.reduceBuf, 0, sizeof(long long)*(3*(size_t) st.Npad + 8), st.stream));
}

} // namespace

void commAllReduce(State& st, long long* buf, size_t count, cudaStream_t s) {
    if (!st.comm) throw std::runtime_error("sharded handle without a communicator: call cfx_comm_init first");
    ncclCheck(nccl().AllReduce(buf, buf, count, ncclInt64, ncclSum, static_cast<ncclComm_t>(st.comm), s), "ncclAllReduce");
}

void commDestroy(State& st) {
    if (st.comm) { nccl().CommDestroy(static_cast<ncclComm_t>(st.comm)); st.comm = nullptr; }
    if (st.reduceBuf) { cudaFree(st.reduceBuf); st.reduceBuf = nullptr; }
}

} // namespace cfx

using namespace cfx;

struct cfx_multi {
    std::vector<cfx_handle*> handles;
    std::vector<int> devices;
    double* hPos = nullptr;            // pinned staging shared by all devices
    int N = 0;
};

extern "C" {

int cfx_comm_get_unique_id(void* id) {
    try {
        if (!id) { setLastError("null argument"); return CFX_ERR_ARGUMENT; }
        static_assert(sizeof(ncclUniqueId) == CFX_COMM_ID_BYTES, "CFX_COMM_ID_BYTES must match ncclUniqueId");
        ncclUniqueId u;
        ncclCheck(nccl().GetUniqueId(&u), "ncclGetUniqueId");
        memcpy(id, &u, sizeof(u));
        return CFX_OK;
    } catch (const std::exception& e) { setLastError(e.what()); return CFX_ERR_CUDA; }
}

int cfx_comm_init(cfx_handle* h, const void* id) {
    try {
        if (!h || !id) { setLastError("null argument"); return CFX_ERR_ARGUMENT; }
        State& st = h->st;
        if (st.comm) { setLastError("the handle already has a communicator"); return CFX_ERR_STATE; }
        CFX_CUDA(cudaSetDevice(st.device));
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclComm_t comm = nullptr;
        ncclCheck(nccl().CommInitRank(&comm, st.shardCount, u, st.shardRank), "ncclCommInitRank");
        st.comm = comm;
        CFX_CUDA(cudaMalloc(&st.reduceBuf, sizeof(long long)*(3*(size_t) st.Npad + 8)));
        warmUp(st);
        commAllReduce(st, st.reduceBuf, 3*(size_t) st.Npad + 8, st.stream);
        CFX_CUDA(cudaStreamSynchronize(st.stream));
        return CFX_OK;
    } catch (const std::exception& e) { setLastError(e.what()); return CFX_ERR_CUDA; }
}

int cfx_comm_size(const cfx_handle* h) { return h ? (h->st.comm ? h->st.shardCount : 0) : 0; }

/* ---- one process, several GPUs ---- */
void cfx_multi_destroy(cfx_multi* m) {
    if (!m) return;
    for (cfx_handle* h : m->handles) if (h) cfx_destroy(h);
    if (m->hPos) cudaFreeHost(m->hPos);
    delete m;
}

int cfx_multi_create(const cfx_system_desc* desc, const int32_t* devices, int32_t numDevices, cfx_multi** out) {
    struct Guard { cfx_multi* m = nullptr; ~Guard() { if (m) cfx_multi_destroy(m); } } guard;
    try {
        if (!desc || !devices || !out || numDevices < 1) { setLastError("bad argument"); return CFX_ERR_ARGUMENT; }
        *out = nullptr;
        cfx_multi* m = guard.m = new cfx_multi();
        m->N = desc->num_particles;
        m->devices.assign(devices, devices + numDevices);
        for (int r = 0; r < numDevices; r++) {
            cfx_options o;
            memset(&o, 0, sizeof(o));
            o.device = devices[r]; o.shard_rank = r; o.shard_count = numDevices; o.use_graph = 1;
            cfx_handle* h = nullptr;
            const int rc = cfx_create(desc, &o, &h);
            if (rc != CFX_OK) return rc;                        // (last error already set)
            m->handles.push_back(h);
        }
        std::vector<ncclComm_t> comms(numDevices);
        ncclCheck(nccl().CommInitAll(comms.data(), numDevices, m->devices.data()), "ncclCommInitAll");
        for (int r = 0; r < numDevices; r++) {
            State& st = m->handles[r]->st;
            st.comm = comms[r];
            CFX_CUDA(cudaSetDevice(st.device));
            CFX_CUDA(cudaMalloc(&st.reduceBuf, sizeof(long long)*(3*(size_t) st.Npad + 8)));
            warmUp(st);
        }
        ncclCheck(nccl().GroupStart(), "ncclGroupStart");
        for (int r = 0; r < numDevices; r++) {
            State& st = m->handles[r]->st;
            commAllReduce(st, st.reduceBuf, 3*(size_t) st.Npad + 8, st.stream);
        }
        ncclCheck(nccl().GroupEnd(), "ncclGroupEnd");
        for (int r = 0; r < numDevices; r++) {
            CFX_CUDA(cudaSetDevice(m->handles[r]->st.device));
            CFX_CUDA(cudaStreamSynchronize(m->handles[r]->st.stream));
        }
        CFX_CUDA(cudaMallocHost(&m->hPos, sizeof(double)*3*std::max(m->N, 1)));
        guard.m = nullptr;
        *out = m;
        return CFX_OK;
    } catch (const std::exception& e) { setLastError(e.what()); return CFX_ERR_CUDA; }
}

int cfx_multi_num_devices(const cfx_multi* m) { return m ? (int) m->handles.size() : 0; }
cfx_handle* cfx_multi_handle(cfx_multi* m, int32_t index) {
    return (m && index >= 0 && index < (int) m->handles.size()) ? m->handles[index] : nullptr;
}

int cfx_multi_execute(cfx_multi* m, const double* positions, const double* box, int includeForces, int includeEnergy,
                      double* energy, double* forces) {
    try {
        if (!m || !positions) { setLastError("null argument"); return CFX_ERR_ARGUMENT; }
        const size_t vecBytes = sizeof(double)*3*(size_t) m->N;
        memcpy(m->hPos, positions, vecBytes);
        const int n = (int) m->handles.size();
        // every device: its own copy of the positions, then its shard of the evaluation (one CUDA graph per device)
        for (int r = 0; r < n; r++) {
            State& st = m->handles[r]->st;
            CFX_CUDA(cudaSetDevice(st.device));
            CFX_CUDA(cudaMemcpyAsync(st.pos, m->hPos, vecBytes, cudaMemcpyHostToDevice, st.stream));
            const int rc = cfx_execute_shard(m->handles[r], st.pos, box, includeForces, includeEnergy, st.reduceBuf, st.stream);
            if (rc != CFX_OK) return rc;
        }
        ncclCheck(nccl().GroupStart(), "ncclGroupStart");
        for (int r = 0; r < n; r++) {
            State& st = m->handles[r]->st;
            commAllReduce(st, st.reduceBuf, 3*(size_t) st.Npad + 8, st.stream);
        }
        ncclCheck(nccl().GroupEnd(), "ncclGroupEnd");
        // device 0 converts and returns the (identical everywhere) result
        State& s0 = m->handles[0]->st;
        CFX_CUDA(cudaSetDevice(s0.device));
        launchFinalize(s0, s0.reduceBuf, s0.reduceBuf + 3*(size_t) s0.Npad, s0.stream);
        CFX_CUDA(cudaMemcpyAsync(s0.hForce, s0.forceOut, vecBytes, cudaMemcpyDeviceToHost, s0.stream));
        CFX_CUDA(cudaMemcpyAsync(s0.hEnergy, s0.energyOut, sizeof(double)*CFX_E_COUNT, cudaMemcpyDeviceToHost, s0.stream));
        for (int r = n - 1; r >= 0; r--) {
            CFX_CUDA(cudaSetDevice(m->handles[r]->st.device));
            CFX_CUDA(cudaStreamSynchronize(m->handles[r]->st.stream));
        }
        if (energy) memcpy(energy, s0.hEnergy, sizeof(double)*CFX_E_COUNT);
        if (forces) for (size_t k = 0; k < 3*(size_t) m->N; k++) forces[k] += s0.hForce[k];
        return CFX_OK;
    } catch (const std::exception& e) { setLastError(e.what()); return CFX_ERR_CUDA; }
}

} // extern "C"
