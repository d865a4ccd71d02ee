// filter_cuda.cu -- the filter stages that FOLLOW the scans, kept on the device: selection of the sequences whose P-value
// passes a threshold (stream compaction, database order preserved), the Viterbi scan of exactly those survivors by index
// list, and the download of the hits only (msv_cuda_db_filter_pipeline in include/msv_cuda.h).
//
// The reference stops at the raw score: it parses the models' Gumbel parameters (data_readers/Profile_HMM.hpp:34-35,
// Profile_HMM.cpp:82-92) and never uses them.  HMMER3's acceleration pipeline is what those parameters are for: MSV filter
// (P <= F1 = 0.02) -> Viterbi filter on the ~2 % survivors (P <= F2 = 1e-3) -> the expensive stages.  Round 1 downloaded three
// n-float arrays and selected on the host, and re-packed the survivors on the host for the second scan; here only the
// hits cross PCIe and the second scan reads the resident database through the survivors' index list.
#include <algorithm>
#include <new>

#include "msv_internal.hpp"

namespace {

constexpr int kThreads = 256;

// candidate t of m: sequence q = candidates ? candidates[t] : t; kept when pvalues[q] <= threshold
__global__ void __launch_bounds__(kThreads) select_count_kernel(const float* __restrict__ pvalues, const uint32_t* __restrict__ candidates, uint32_t m,
                                                               float threshold, uint32_t* __restrict__ tile_counts) {
    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    const bool keep = t < m && pvalues[candidates ? candidates[t] : t] <= threshold;
    const uint32_t count = __syncthreads_count(keep);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = count;
}

// exclusive scan of the tile counts (one CTA), total -> *total
__global__ void __launch_bounds__(1024) select_scan_kernel(const uint32_t* __restrict__ tile_counts, uint32_t tiles, uint32_t* __restrict__ tile_base,
                                                          uint32_t* __restrict__ total) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (tiles + blockDim.x - 1) / blockDim.x;
    const uint32_t first = min(tiles, threadIdx.x * per), last = min(tiles, first + per);
    uint32_t sum = 0;
    for (uint32_t i = first; i < last; ++i) sum += tile_counts[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t k = 0; k < blockDim.x; ++k) {
            const uint32_t v = part[k];
            part[k] = run;
            run += v;
        }
        *total = run;
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t i = first; i < last; ++i) {
        tile_base[i] = run;
        run += tile_counts[i];
    }
}

// the kept sequences, in candidate order, with their statistics gathered next to them
__global__ void __launch_bounds__(kThreads) select_scatter_kernel(const float* __restrict__ pvalues, const uint32_t* __restrict__ candidates, uint32_t m,
                                                                 float threshold, const uint32_t* __restrict__ tile_base,
                                                                 const float* __restrict__ scores, const float* __restrict__ bits,
                                                                 uint32_t* __restrict__ selected, float* __restrict__ hit_scores,
                                                                 float* __restrict__ hit_bits, float* __restrict__ hit_pvalues) {
    __shared__ uint32_t warp_counts[kThreads / 32];
    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    const uint32_t q = t < m ? (candidates ? candidates[t] : t) : 0u;
    const float p = t < m ? pvalues[q] : 2.0f;
    const bool keep = t < m && p <= threshold;
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_counts[warp] = __popc(ballot);
    __syncthreads();
    uint32_t at = tile_base[blockIdx.x] + __popc(ballot & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) at += warp_counts[w];
    if (keep) {
        selected[at] = q;
        hit_scores[at] = scores[q];
        hit_bits[at] = bits[q];
        hit_pvalues[at] = p;
    }
}

// bit score and Gumbel P-value of the listed sequences only (same formulas as msv_filter_statistics_kernel)
__global__ void __launch_bounds__(kThreads) subset_statistics_kernel(const float* __restrict__ scores, const uint64_t* __restrict__ offsets,
                                                                    const uint32_t* __restrict__ indices, uint32_t m, double mu, double lambda,
                                                                    float* __restrict__ bits_out, float* __restrict__ p_out) {
    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= m) return;
    const uint32_t q = indices[t];
    const double L = static_cast<double>(offsets[q + 1] - offsets[q]);
    const double null1 = L * log(L / (L + 1.0)) + log(1.0 / (L + 1.0));
    const double bits = (static_cast<double>(scores[q]) - null1) / 0.69314718055994530942;
    const double ey = -exp(-lambda * (bits - mu));
    const double p = fabs(ey) < 5e-9 ? -ey : 1.0 - exp(ey);
    bits_out[q] = static_cast<float>(bits);
    p_out[q] = static_cast<float>(p);
}

struct Filter_scratch { // hangs off the database (grow-only)
    uint32_t* d_selected[2] = {nullptr, nullptr}; // survivors of stage 1 / stage 2
    uint32_t* d_tiles = nullptr;                  // counts | bases | total
    float* d_hits = nullptr;                      // scores | bits | pvalues of the selected, compact
    float* d_stats = nullptr;                     // bits | pvalues per sequence (stage 1), bits | pvalues (stage 2)
    float* d_second = nullptr;                    // stage-2 raw scores per sequence
    size_t capacity = 0;
};

int reserve(msv_db* db, Filter_scratch*& scratch_out) {
    auto* scratch = static_cast<Filter_scratch*>(db->filter_scratch);
    if (!scratch) {
        scratch = new (std::nothrow) Filter_scratch();
        if (!scratch) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
        db->filter_scratch = scratch;
        db->filter_scratch_free = [](void* raw) {
            auto* s = static_cast<Filter_scratch*>(raw);
            cudaFree(s->d_selected[0]);
            cudaFree(s->d_selected[1]);
            cudaFree(s->d_tiles);
            cudaFree(s->d_hits);
            cudaFree(s->d_stats);
            cudaFree(s->d_second);
            delete s;
        };
    }
    if (scratch->capacity < db->n) {
        for (void* ptr : {static_cast<void*>(scratch->d_selected[0]), static_cast<void*>(scratch->d_selected[1]), static_cast<void*>(scratch->d_tiles),
                          static_cast<void*>(scratch->d_hits), static_cast<void*>(scratch->d_stats), static_cast<void*>(scratch->d_second)})
            cudaFree(ptr);
        *scratch = Filter_scratch();
        const size_t cap = db->n + db->n / 8 + 256;
        const size_t tiles = (cap + kThreads - 1) / kThreads;
        MSV_CUDA_TRY(cudaMalloc(&scratch->d_selected[0], cap * sizeof(uint32_t)));
        MSV_CUDA_TRY(cudaMalloc(&scratch->d_selected[1], cap * sizeof(uint32_t)));
        MSV_CUDA_TRY(cudaMalloc(&scratch->d_tiles, (2 * tiles + 8) * sizeof(uint32_t)));
        MSV_CUDA_TRY(cudaMalloc(&scratch->d_hits, 3 * cap * sizeof(float)));
        MSV_CUDA_TRY(cudaMalloc(&scratch->d_stats, 4 * cap * sizeof(float)));
        MSV_CUDA_TRY(cudaMalloc(&scratch->d_second, cap * sizeof(float)));
        scratch->capacity = cap;
    }
    scratch_out = scratch;
    return MSV_OK;
}

// select among m candidates; *count on the host (one small synchronisation); the compact statistics land in scratch->d_hits
int select(msv_db* db, Filter_scratch* s, const float* scores, const float* bits, const float* pvalues, const uint32_t* candidates, size_t m,
           float threshold, uint32_t* selected, size_t* count) {
    *count = 0;
    if (m == 0) return MSV_OK;
    const uint32_t m32 = static_cast<uint32_t>(m), tiles = (m32 + kThreads - 1) / kThreads;
    uint32_t* counts = s->d_tiles;
    uint32_t* bases = counts + (s->capacity + kThreads - 1) / kThreads;
    uint32_t* total = bases + (s->capacity + kThreads - 1) / kThreads;
    select_count_kernel<<<tiles, kThreads>>>(pvalues, candidates, m32, threshold, counts);
    select_scan_kernel<<<1, 1024>>>(counts, tiles, bases, total);
    select_scatter_kernel<<<tiles, kThreads>>>(pvalues, candidates, m32, threshold, bases, scores, bits, selected, s->d_hits, s->d_hits + s->capacity,
                                               s->d_hits + 2 * s->capacity);
    for (int k = 0; k < 3; ++k) msv_detail::count_launch();
    MSV_CUDA_TRY(cudaGetLastError());
    uint32_t found = 0;
    MSV_CUDA_TRY(cudaMemcpy(&found, total, sizeof found, cudaMemcpyDeviceToHost));
    *count = found;
    (void)db;
    return MSV_OK;
}

int download_hits(Filter_scratch* s, const uint32_t* selected, size_t count, size_t capacity, uint32_t* hit_index, float* hit_score, float* hit_bits,
                  float* hit_pvalue) {
    const size_t take = std::min(count, capacity);
    if (take == 0) return MSV_OK;
    if (hit_index) MSV_CUDA_TRY(cudaMemcpy(hit_index, selected, take * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (hit_score) MSV_CUDA_TRY(cudaMemcpy(hit_score, s->d_hits, take * sizeof(float), cudaMemcpyDeviceToHost));
    if (hit_bits) MSV_CUDA_TRY(cudaMemcpy(hit_bits, s->d_hits + s->capacity, take * sizeof(float), cudaMemcpyDeviceToHost));
    if (hit_pvalue) MSV_CUDA_TRY(cudaMemcpy(hit_pvalue, s->d_hits + 2 * s->capacity, take * sizeof(float), cudaMemcpyDeviceToHost));
    return MSV_OK;
}

} // namespace

extern "C" {

int msv_cuda_db_msv_filter(msv_model* model, msv_db* db, float mu, float lambda, float threshold, uint32_t* hit_index, float* hit_score,
                           float* hit_bits, float* hit_pvalue, size_t capacity, size_t* n_hits) {
    if (!model || !db || !n_hits) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    *n_hits = 0;
    db->n_survivors = 0;
    if (db->n == 0) return MSV_OK;
    Device_guard guard(db->device);
    MSV_CUDA_TRY(guard.status);
    Filter_scratch* s = nullptr;
    if (int rc = reserve(db, s)) return rc;
    float* bits = s->d_stats;
    float* pvalues = s->d_stats + s->capacity;
    if (int rc = msv_cuda_db_score_device(model, db, db->d_scores, nullptr)) return rc;
    if (int rc = msv_cuda_db_filter_device(db, db->d_scores, mu, lambda, bits, pvalues, nullptr)) return rc;
    size_t found = 0;
    if (int rc = select(db, s, db->d_scores, bits, pvalues, nullptr, db->n, threshold, s->d_selected[0], &found)) return rc;
    db->n_survivors = found; // the index list stays on the device for msv_cuda_db_viterbi_filter_survivors
    *n_hits = found;
    return download_hits(s, s->d_selected[0], found, capacity, hit_index, hit_score, hit_bits, hit_pvalue);
}

int msv_cuda_db_viterbi_filter_survivors(msv_viterbi_model* model, msv_db* db, float mu, float lambda, float threshold, uint32_t* hit_index,
                                         float* hit_score, float* hit_bits, float* hit_pvalue, size_t capacity, size_t* n_hits) {
    if (!model || !db || !n_hits) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    *n_hits = 0;
    const size_t survivors = db->n_survivors;
    if (survivors == 0) return MSV_OK;
    auto* s = static_cast<Filter_scratch*>(db->filter_scratch);
    if (!s || survivors > db->n) return fail(MSV_ERR_INVALID_ARGUMENT, "no survivors of an MSV filter stage on this database");
    Device_guard guard(db->device);
    MSV_CUDA_TRY(guard.status);
    float* bits = s->d_stats + 2 * s->capacity;
    float* pvalues = s->d_stats + 3 * s->capacity;
    // the second scan reads the resident database through the survivors' index list; nothing is re-packed or re-uploaded
    if (int rc = msv_cuda_db_viterbi_subset_device(model, db, s->d_selected[0], survivors, s->d_second, nullptr)) return rc;
    const uint32_t m32 = static_cast<uint32_t>(survivors);
    subset_statistics_kernel<<<(m32 + kThreads - 1) / kThreads, kThreads>>>(s->d_second, db->d_offsets, s->d_selected[0], m32, static_cast<double>(mu),
                                                                            static_cast<double>(lambda), bits, pvalues);
    msv_detail::count_launch();
    MSV_CUDA_TRY(cudaGetLastError());
    size_t found = 0;
    if (int rc = select(db, s, s->d_second, bits, pvalues, s->d_selected[0], survivors, threshold, s->d_selected[1], &found)) return rc;
    *n_hits = found;
    return download_hits(s, s->d_selected[1], found, capacity, hit_index, hit_score, hit_bits, hit_pvalue);
}

} // extern "C"
