// msv_internal.hpp -- what the translation units of libmsv_cuda.so share (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "msv_cuda.h"

namespace msv_detail {
int fail(int code, const char* fmt, ...); // records the message for msv_cuda_last_error() and returns `code`
void count_launch();                      // msv_cuda_launch_count bookkeeping
} // namespace msv_detail
using msv_detail::fail;

#define MSV_CUDA_TRY(expr)                                                                                             \
    do {                                                                                                               \
        cudaError_t err__ = (expr);                                                                                    \
        if (err__ != cudaSuccess) {                                                                                    \
            const int code__ = (err__ == cudaErrorNoDevice || err__ == cudaErrorInsufficientDriver) ? MSV_ERR_NO_DEVICE \
                               : (err__ == cudaErrorMemoryAllocation)                               ? MSV_ERR_OUT_OF_MEMORY \
                                                                                                    : MSV_ERR_CUDA;    \
            (void)cudaGetLastError();                                                                                  \
            return fail(code__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__);        \
        }                                                                                                              \
    } while (0)

struct Device_guard {
    int previous = -1;
    cudaError_t status;
    explicit Device_guard(int device) {
        status = cudaGetDevice(&previous);
        if (status == cudaSuccess && previous != device) status = cudaSetDevice(device);
    }
    ~Device_guard() {
        if (previous >= 0) (void)cudaSetDevice(previous);
    }
};

// ---- device-resident database (opaque handle of the C ABI) --------------------------------------------------------------
constexpr int kMaxChunks = 32; // pipelined upload: at most this many upload/scan stages per batch
constexpr uint32_t kProfileStep = 32, kProfileBuckets = 2048; // host length profile: 32-row buckets up to 65 504 rows
struct msv_db {
    int device = 0;
    size_t n = 0;
    uint64_t total = 0;
    uint64_t longest = 0;
    // device buffers (grow-only capacities so a workspace database can be refilled without reallocating)
    uint8_t* d_residues = nullptr;
    size_t cap_residues = 0;
    uint64_t* d_offsets = nullptr;
    uint32_t* d_order = nullptr;
    float* d_scores = nullptr;
    float* d_stats = nullptr; // bit scores | P-values, 2 * cap_n, allocated on first use
    size_t cap_stats = 0;
    size_t cap_n = 0;
    float2* d_length_tr = nullptr;
    size_t cap_tr = 0;
    uint32_t* d_hist = nullptr; // hist | cursor, 2 * kBuckets; twice (the pipelined upload buckets alternate stages on two streams)
    unsigned int* d_queue = nullptr;
    unsigned long long* d_first_bad = nullptr;
    std::vector<float2> h_length_tr; // host copy, extended lazily
    std::vector<uint32_t> h_lengths; // sequence lengths, kept on the host only for small databases (launch planning)
    // length profile of the whole database for the launch planner (lane-group plans put the longest sequences on fast
    // CTAs): sequences and rows per bucket of kProfileStep lengths, the last bucket open-ended; empty when the offsets
    // never were on the host (databases parsed on the device)
    std::vector<uint32_t> h_profile_count;
    std::vector<uint64_t> h_profile_rows;
    // pipelined upload (msv_cuda_score_batch): copy engine and scan overlap
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr, compute_stream2 = nullptr;
    cudaEvent_t stage_copied[kMaxChunks] = {};
    cudaEvent_t reserved = nullptr, other_done = nullptr;
    // filter stages kept on the device (filter_cuda.cu): scratch + the survivors of the last MSV filter stage
    void* filter_scratch = nullptr;
    void (*filter_scratch_free)(void*) = nullptr;
    size_t n_survivors = 0;
};

namespace msv_detail {
// (re)fill a database handle from host buffers on the default stream: upload, validate, bucket longest-first; buffers only
// grow, so a workspace handle can be refilled call after call without reallocating (msv_cuda.cu)
int db_refill(msv_db* db, const uint8_t* residues, const uint64_t* offsets, size_t n);
int db_free(msv_db* db);
// for producers that fill a database ON the device (fasta_cuda.cu): make room for `total` residues / `n` sequences and the
// per-length transition table up to `longest` (queued on `stream`); then, once d_residues / d_offsets hold valid codes and
// offsets and db->n / total / longest are set, bucket the sequences longest-first
int db_reserve_for(msv_db* db, uint64_t total, size_t n, uint64_t longest, cudaStream_t stream);
int db_bucket(msv_db* db, cudaStream_t stream);
// FASTA text (host memory) -> `db`, parsed on the device (fasta_cuda.cu); *rejected = records dropped for a foreign character
int db_fill_from_fasta(msv_db* db, const char* text, size_t bytes, size_t* rejected);
} // namespace msv_detail
