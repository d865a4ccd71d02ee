// msv_multi.cu -- several GPUs of one box behind the C ABI, from ONE process (msv_cuda_multi_* in include/msv_cuda.h).
//
// The reference has no multi-device path at all (one context, queue on its first device: reference
// algorithms/MSV_HMM.cpp:230,317).  Sequences are independent, so the database is cut into contiguous slices of equal
// DP-cell count (msv_host_partition_by_cells); slice g is uploaded to and scanned on GPU g by its own host thread through
// the same pipelined end-to-end path as msv_cuda_score_batch.  The only exchange is the gather of the fp32 scores:
//   MSV_GATHER_HOST : every GPU downloads its slice straight into the caller's host array (no device-side gather);
//   MSV_GATHER_PEER : the scan kernels store every score straight into ONE array on the first GPU over NVLink peer
//                     access (the fused gather of msv_cuda_db_score_gather); one download of the whole array follows;
//   MSV_GATHER_NCCL : every GPU scans into its own buffer, then one grouped ncclSend / ncclRecv round collects the
//                     slices on the first GPU (communicators from ncclCommInitAll); one download follows.
// NCCL is bound at run time (dlopen of libnccl.so.2), so the library itself has no link-time dependency on it.
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "msv_internal.hpp"

namespace {

// ---- the few NCCL entry points this file needs, bound lazily ----------------------------------------------------------
// (types as in nccl.h: ncclComm_t is an opaque pointer, ncclResult_t / ncclDataType_t are ints, ncclFloat32 == 7)
typedef void* nccl_comm;
struct Nccl_api {
    int (*comm_init_all)(nccl_comm*, int, const int*) = nullptr;
    int (*comm_destroy)(nccl_comm) = nullptr;
    int (*group_start)() = nullptr;
    int (*group_end)() = nullptr;
    int (*send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char* (*error_string)(int) = nullptr;
    bool ok = false;
    std::string why;
};
constexpr int kNcclFloat32 = 7;

const Nccl_api& nccl_api() {
    static const Nccl_api api = [] {
        Nccl_api a;
        void* handle = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            a.why = std::string("libnccl.so.2 not loadable: ") + (dlerror() ? dlerror() : "?");
            return a;
        }
        const auto bind = [&](auto& fn, const char* symbol) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(handle, symbol));
            if (!fn && a.why.empty()) a.why = std::string("symbol missing in libnccl: ") + symbol;
        };
        bind(a.comm_init_all, "ncclCommInitAll");
        bind(a.comm_destroy, "ncclCommDestroy");
        bind(a.group_start, "ncclGroupStart");
        bind(a.group_end, "ncclGroupEnd");
        bind(a.send, "ncclSend");
        bind(a.recv, "ncclRecv");
        bind(a.error_string, "ncclGetErrorString");
        a.ok = a.why.empty();
        return a;
    }();
    return api;
}

} // namespace

struct msv_multi {
    std::vector<msv_model*> models; // not owned
    std::vector<int> devices;
    // per GPU: a score buffer (grow-only) and a stream for the NCCL round; GPU 0's buffer holds the whole job's scores
    std::vector<float*> d_scores;
    std::vector<size_t> capacity;
    std::vector<cudaStream_t> streams;
    std::vector<char> peer_ok; // GPU g can store into GPU 0's memory
    std::vector<nccl_comm> comms;
    size_t last_n = 0;
};

namespace {

int reserve_scores(msv_multi* multi, int g, size_t floats) {
    if (floats <= multi->capacity[g]) return MSV_OK;
    Device_guard guard(multi->devices[g]);
    MSV_CUDA_TRY(guard.status);
    cudaFree(multi->d_scores[g]);
    multi->d_scores[g] = nullptr;
    multi->capacity[g] = 0;
    const size_t want = floats + floats / 8 + 256;
    MSV_CUDA_TRY(cudaMalloc(&multi->d_scores[g], want * sizeof(float)));
    multi->capacity[g] = want;
    return MSV_OK;
}

int nccl_fail(const Nccl_api& api, int result, const char* what) {
    return fail(MSV_ERR_CUDA, "%s failed: %s", what, api.error_string ? api.error_string(result) : "NCCL error");
}

} // namespace

extern "C" {

int msv_cuda_multi_create(msv_model* const* models, int ngpu, msv_multi** out) {
    if (!out) return fail(MSV_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!models || ngpu < 1 || ngpu > 8) return fail(MSV_ERR_INVALID_ARGUMENT, "between 1 and 8 models (one per GPU) are supported");
    auto* multi = new (std::nothrow) msv_multi();
    if (!multi) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    for (int g = 0; g < ngpu; ++g) {
        int device = -1;
        if (!models[g] || msv_cuda_model_device(models[g], &device) != MSV_OK) {
            delete multi;
            return fail(MSV_ERR_INVALID_ARGUMENT, "models[%d] is NULL", g);
        }
        multi->models.push_back(models[g]);
        multi->devices.push_back(device);
    }
    multi->d_scores.assign(ngpu, nullptr);
    multi->capacity.assign(ngpu, 0);
    multi->streams.assign(ngpu, nullptr);
    multi->peer_ok.assign(ngpu, 0);
    // peer access towards the first GPU (where the gathered array lives); "already enabled" is fine
    for (int g = 0; g < ngpu; ++g) {
        Device_guard guard(multi->devices[g]);
        if (guard.status != cudaSuccess) continue;
        (void)cudaStreamCreateWithFlags(&multi->streams[g], cudaStreamNonBlocking);
        if (multi->devices[g] == multi->devices[0]) {
            multi->peer_ok[g] = 1;
            continue;
        }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, multi->devices[g], multi->devices[0]) == cudaSuccess && can) {
            const cudaError_t err = cudaDeviceEnablePeerAccess(multi->devices[0], 0);
            multi->peer_ok[g] = (err == cudaSuccess || err == cudaErrorPeerAccessAlreadyEnabled) ? 1 : 0;
        }
        (void)cudaGetLastError();
    }
    *out = multi;
    return MSV_OK;
}

int msv_cuda_multi_destroy(msv_multi* multi) {
    if (!multi) return MSV_OK;
    if (!multi->comms.empty() && nccl_api().ok)
        for (nccl_comm c : multi->comms)
            if (c) nccl_api().comm_destroy(c);
    for (size_t g = 0; g < multi->devices.size(); ++g) {
        Device_guard guard(multi->devices[g]);
        cudaFree(multi->d_scores[g]);
        if (multi->streams[g]) cudaStreamDestroy(multi->streams[g]);
    }
    delete multi;
    return MSV_OK;
}

int msv_cuda_multi_score_batch(msv_multi* multi, const uint8_t* residues, const uint64_t* offsets, size_t n, float* scores_host,
                               int gather) {
    if (!multi) return fail(MSV_ERR_INVALID_ARGUMENT, "multi is NULL");
    if (n && (!offsets || !scores_host)) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    if (gather != MSV_GATHER_HOST && gather != MSV_GATHER_PEER && gather != MSV_GATHER_NCCL)
        return fail(MSV_ERR_INVALID_ARGUMENT, "unknown gather mode %d", gather);
    const int ngpu = static_cast<int>(multi->devices.size());
    multi->last_n = 0;
    if (n == 0) return MSV_OK;
    std::vector<size_t> bounds(static_cast<size_t>(ngpu) + 1);
    if (int rc = msv_host_partition_by_cells(offsets, n, ngpu, bounds.data())) return rc;

    if (gather == MSV_GATHER_PEER)
        for (int g = 0; g < ngpu; ++g)
            if (!multi->peer_ok[g])
                return fail(MSV_ERR_CUDA, "GPU %d cannot store into GPU %d's memory (no peer access); use MSV_GATHER_HOST or MSV_GATHER_NCCL",
                            multi->devices[g], multi->devices[0]);
    const Nccl_api* nccl = nullptr;
    if (gather == MSV_GATHER_NCCL) {
        nccl = &nccl_api();
        if (!nccl->ok) return fail(MSV_ERR_CUDA, "NCCL unavailable: %s", nccl->why.c_str());
        if (multi->comms.empty()) {
            multi->comms.assign(ngpu, nullptr);
            const int result = nccl->comm_init_all(multi->comms.data(), ngpu, multi->devices.data());
            if (result != 0) {
                multi->comms.clear();
                return nccl_fail(*nccl, result, "ncclCommInitAll");
            }
        }
    }
    // device buffers: GPU 0 holds the whole job's scores (device-side gathers), the others their own slice (NCCL only)
    if (gather != MSV_GATHER_HOST) {
        if (int rc = reserve_scores(multi, 0, n)) return rc;
        if (gather == MSV_GATHER_NCCL)
            for (int g = 1; g < ngpu; ++g)
                if (int rc = reserve_scores(multi, g, bounds[g + 1] - bounds[g])) return rc;
    }

    // ---- one host thread per GPU: upload + bucket + scan of its slice (pipelined inside the call) ----
    std::vector<int> status(ngpu, MSV_OK);
    std::vector<std::string> message(ngpu);
    const bool trace = std::getenv("MSV_MULTI_TRACE") != nullptr; // tuning aid: per-GPU wall time of the slice, on stderr
    const auto call_begin = std::chrono::steady_clock::now();
    const auto work = [&](int g) {
        const size_t first = bounds[g], last = bounds[g + 1];
        if (first == last) return;
        const auto begin = std::chrono::steady_clock::now();
        std::vector<uint64_t> local(offsets + first, offsets + last + 1); // the slice's offsets, rebased to its first residue
        const uint64_t base = local.front();
        for (auto& o : local) o -= base;
        const uint8_t* slice = residues ? residues + base : nullptr;
        int rc;
        if (gather == MSV_GATHER_HOST) {
            rc = msv_cuda_score_batch(multi->models[g], slice, local.data(), last - first, scores_host + first);
        } else {
            // PEER: the kernel on GPU g stores into GPU 0's array over NVLink; NCCL: into GPU g's own buffer
            float* target = (gather == MSV_GATHER_PEER || g == 0) ? multi->d_scores[0] + first : multi->d_scores[g];
            rc = msv_cuda_score_batch_gather(multi->models[g], slice, local.data(), last - first, &target, 1, 0);
        }
        if (rc != MSV_OK) {
            status[g] = rc;
            message[g] = msv_cuda_last_error();
        }
        if (trace) {
            const auto end = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[msv_multi] GPU %d: %zu sequences, started %.3f ms into the call, took %.3f ms\n", multi->devices[g], last - first,
                         std::chrono::duration<double, std::milli>(begin - call_begin).count(),
                         std::chrono::duration<double, std::milli>(end - begin).count());
        }
    };
    {
        std::vector<std::thread> pool;
        for (int g = 1; g < ngpu; ++g) pool.emplace_back(work, g);
        work(0);
        for (auto& t : pool) t.join();
    }
    for (int g = 0; g < ngpu; ++g)
        if (status[g] != MSV_OK) return fail(status[g], "GPU %d: %s", multi->devices[g], message[g].c_str());
    if (gather == MSV_GATHER_HOST) return MSV_OK;

    if (gather == MSV_GATHER_NCCL && ngpu > 1) {
        int result = nccl->group_start();
        for (int g = 1; g < ngpu && result == 0; ++g) {
            const size_t count = bounds[g + 1] - bounds[g];
            if (count == 0) continue;
            result = nccl->send(multi->d_scores[g], count, kNcclFloat32, 0, multi->comms[g], multi->streams[g]);
            if (result == 0) result = nccl->recv(multi->d_scores[0] + bounds[g], count, kNcclFloat32, g, multi->comms[0], multi->streams[0]);
        }
        const int end_result = nccl->group_end();
        if (result != 0 || end_result != 0) return nccl_fail(*nccl, result != 0 ? result : end_result, "ncclSend/ncclRecv");
        for (int g = 1; g < ngpu; ++g) {
            Device_guard guard(multi->devices[g]);
            MSV_CUDA_TRY(cudaStreamSynchronize(multi->streams[g]));
        }
    }
    Device_guard guard(multi->devices[0]);
    MSV_CUDA_TRY(guard.status);
    MSV_CUDA_TRY(cudaMemcpyAsync(scores_host, multi->d_scores[0], n * sizeof(float), cudaMemcpyDeviceToHost, multi->streams[0]));
    MSV_CUDA_TRY(cudaStreamSynchronize(multi->streams[0]));
    multi->last_n = n;
    return MSV_OK;
}

int msv_cuda_multi_gathered(const msv_multi* multi, const float** scores_device, size_t* n, int* device) {
    if (!multi) return fail(MSV_ERR_INVALID_ARGUMENT, "multi is NULL");
    if (scores_device) *scores_device = multi->last_n ? multi->d_scores[0] : nullptr;
    if (n) *n = multi->last_n;
    if (device) *device = multi->devices[0];
    return MSV_OK;
}

} // extern "C"
