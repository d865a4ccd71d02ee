// msv_cuda.cu -- implementation of the C ABI declared in include/msv_cuda.h.
//
// Host side of the boundary: model upload (table re-layout for the kernel), database upload + validation +
// longest-first bucketing, kernel dispatch by geometry, single-sequence latency path.  The kernels themselves are in
// msv_kernels.cuh.  Reference behaviour replaced: MSV_HMM::parallel_run_on_sequence and its OpenCL helpers
// (reference algorithms/MSV_HMM.cpp:118-430).
#include "msv_cuda.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "msv_internal.hpp"
#include "msv_kernels.cuh"
#include "msv_registry.hpp"
#include "msv_wave_kernels.cuh"

static thread_local std::string g_last_error;
static thread_local uint64_t g_launches = 0;

namespace msv_detail {
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
void count_launch() { ++g_launches; }
} // namespace msv_detail

namespace {

// ---- kernel registry: msv_registry.hpp; the table is compiled in four parts and joined here ---------------------------
using msv_registry::Geometry;
using msv_registry::Scan_kernel;
using msv_registry::Rows;
using msv_registry::kExact;
using msv_registry::kWhole;
using msv_registry::kBlocks;
using msv_registry::quiet_rows;

const std::vector<Geometry>& all_geometries() {
    static const std::vector<Geometry> table = [] {
        std::vector<Geometry> t;
        for (auto part : {msv_registry::part0, msv_registry::part1, msv_registry::part2, msv_registry::part3}) {
            size_t count = 0;
            const Geometry* entries = part(&count);
            t.insert(t.end(), entries, entries + count);
        }
        return t;
    }();
    return table;
}

const Geometry* find_geometry(int G, int K, int KT, int threads = 0, int variant = 0) {
    for (const auto& g : all_geometries())
        if (g.G == G && g.K == K && g.KT == KT && (threads == 0 || g.threads == threads) && g.variant == variant) return &g;
    return nullptr;
}

int round_up4(size_t v) { return static_cast<int>((v + 3) / 4 * 4); }

// Default: one warp per sequence with the shared-memory + tensor-memory split (that kernel needs the last column of
// lane 31 to be padding, hence 32*K > columns).  The generic family (8/16/32 lanes per sequence, shared memory only) is
// kept for comparison and for devices/contexts where TMEM is unavailable.
// MSV_CUDA_GEOMETRY="G,K[,KT[,threads[,variant]]]" overrides the choice (tuning aid; without KT the generic kernel is selected).
const Geometry* choose_geometry(size_t columns) {
    if (const char* env = std::getenv("MSV_CUDA_GEOMETRY")) {
        int G = 0, K = 0, KT = -1, T = 0, V = 0;
        const int got = std::sscanf(env, "%d,%d,%d,%d,%d", &G, &K, &KT, &T, &V);
        if (got == 2 && static_cast<size_t>(G) * K > columns)
            if (const Geometry* g = find_geometry(G, K, -1)) return g;
        if (got >= 3 && static_cast<size_t>(G) * K > columns)
            if (const Geometry* g = find_geometry(G, K, KT, T, V)) return g;
    }
    // Measured on B200 over the 24 fixture models (profiles/r01/sweep_models_v2.jsonl, sweep_kt_v2.jsonl,
    // sweep_tmem_ahead_models.jsonl): the warp-per-sequence kernel with the shared-memory/tensor-memory split wins at every
    // model length; how many columns per lane come from tensor memory, and whether they are loaded a row ahead
    // (variant 1), is picked per K from those sweeps.
    // 32*K > columns (lane 31 ends in a padding column).  K moves in steps of two up to 58 (less padding: +2..7 % on the
    // fixture models, profiles/r01/sweep_models_v5_even_k.jsonl) and in steps of four beyond (there the 18-column tensor
    // part of the odd steps spills registers and loses to the next multiple of four).
    int K = std::max(4, static_cast<int>((columns + 1 + 63) / 64 * 2));
    if (K > 58) K = round_up4(K);
    if (K > msv::kMaxColumnsPerLane) return nullptr;
    struct Choice {
        int KT, variant;
    };
    const auto pick = [](int k) -> Choice {
        if (k % 4 == 2) return {k <= 18 ? k : 18, 1}; // columns per lane in steps of two: 18 = 16 + 2 tensor-memory columns
        switch (k) {
        case 4: return {0, 0};
        case 8: case 12: return {8, 1};
        case 16: return {16, 1};
        case 20: return {16, 1}; // (round 2, profiles/r02/kt_sweep_v1.txt: 8.65 vs 8.40 TCUPS for the row-ahead variant)
        case 24: case 28: return {16, 1};
        case 32: case 36: return {24, 1};
        case 40: return {16, 1};
        case 44: return {24, 1};
        case 52: return {16, 0}; // (round 2, kt_sweep_v1.txt: 9.52 vs 9.35 TCUPS)
        case 48: case 56: return {16, 1};
        default: return {24, 1}; // 60 .. 88
        }
    };
    const Choice choice = pick(K);
    return find_geometry(32, K, choice.KT, 0, choice.variant);
}

// Eight lanes per sequence (four sequences per warp) pays off for short models: measured on B200
// (profiles/r01/sweep_generic_v3.jsonl) it beats the warp plan for LENG 200-447 when there are enough sequences to
// balance its four times more slots (plan_launch checks that per launch).
const Geometry* choose_octet_geometry(size_t columns) {
    const int K = std::max(4, static_cast<int>((columns + 1 + 7) / 8 + 1) / 2 * 2); // 8*K > columns, K even
    return (columns >= 64 && K <= 56) ? find_geometry(8, K, -1) : nullptr;
}
// Four lanes per sequence (eight sequences per warp) for the shortest models: less padding (LENG 100: 112 instead of 128
// slots) and the per-row bookkeeping of a warp is shared by eight sequences.
const Geometry* choose_narrow_geometry(size_t columns) {
    const int K = std::max(4, static_cast<int>((columns + 1 + 3) / 4 + 1) / 2 * 2); // 4*K > columns, K even
    return (columns >= 32 && K <= 56) ? find_geometry(4, K, -1) : nullptr;
}

// Four warps per sequence: the plan for few/long sequences and for models beyond one warp's registers.
const Geometry* choose_quad_geometry(size_t columns) {
    const int K = std::max(4, round_up4((columns + 127) / 128));
    for (const auto& g : all_geometries())
        if (g.G == 128 && g.K == K) return &g;
    return nullptr;
}

// ---- single-sequence latency kernels (msv_wave_kernels.cuh), one pair per columns-per-lane K --------------------------
constexpr size_t kWaveDefaultMaxWarps = 16; // default chain length limit (the smallest K that stays within it is chosen)
struct Wave_kernels {
    int K;
    void (*with_inline_residues)(const msv::Wave_params, const msv::Wave_inline_residues);
    void (*with_device_residues)(const msv::Wave_params, const msv::Wave_no_residues);
};
template <int K> constexpr Wave_kernels wave_entry() { return {K, msv::msv_wave_kernel<K, true>, msv::msv_wave_kernel<K, false>}; }
const Wave_kernels g_wave_kernels[] = {wave_entry<2>(), wave_entry<4>(), wave_entry<6>(), wave_entry<8>(), wave_entry<12>(), wave_entry<16>()};

// Columns per lane of the chain: more warps shorten a row (each warp has fewer cells) until per-row bookkeeping and the
// pipeline fill dominate.  MSV_CUDA_WAVE_K overrides (tuning aid).
const Wave_kernels* choose_wave_kernels(size_t columns) {
    const size_t max_warps = static_cast<size_t>(msv::kWaveWarpsPerCta) * msv::kWaveMaxCtas;
    int want = 0;
    if (const char* env = std::getenv("MSV_CUDA_WAVE_K")) want = std::atoi(env);
    const Wave_kernels* fallback = nullptr;
    for (const auto& k : g_wave_kernels) {
        const size_t warps = (columns + 32 * k.K - 1) / (32 * static_cast<size_t>(k.K));
        if (warps > max_warps) continue;
        if (k.K == want) return &k;
        if (!fallback && (warps <= kWaveDefaultMaxWarps || k.K == 16)) fallback = &k;
    }
    return want ? nullptr : fallback;
}

} // namespace

// ---- opaque handles ---------------------------------------------------------------------------------------------
struct msv_model {
    int device = 0;
    size_t model_length = 0;
    // A plan = one kernel geometry + the emission table laid out for it (shared-memory part, then tensor-memory part).
    struct Plan {
        const Geometry* geo = nullptr;
        size_t table_bytes = 0;  // whole table in HBM
        size_t shared_bytes = 0; // dynamic shared memory of the scan kernel
        float4* d_table = nullptr;
    };
    Plan bulk;  // many sequences: one warp per sequence (or whatever MSV_CUDA_GEOMETRY forces)
    Plan octet; // short models and enough sequences to balance 4x more slots: eight lanes per sequence (may be absent)
    Plan narrow; // the shortest models, eight times more slots: four lanes per sequence (may be absent)
    Plan quad;  // few or very long sequences, single-sequence calls: four warps per sequence (may be absent)
    bool forced = false; // MSV_CUDA_GEOMETRY was given: always use `bulk`
    float tr_B_Mk = 0, tr_E_C = 0, tr_E_J = 0;
    int sm_count = 0;
    msv_db* workspace = nullptr; // reused by msv_cuda_score_batch / msv_cuda_score_sequence
    // Feedback for the choice between the two speculating warp kernels (launch_scan): running totals on the device
    // ([0] failed speculations, [1] sequences offered), copied after every such launch into pinned host memory, which the
    // next launch reads without waiting -- a heuristic input, so a stale value is fine.
    unsigned int* d_speculation_totals = nullptr;
    volatile unsigned int* h_speculation_totals = nullptr;
    unsigned int seen_failures = 0, seen_offered = 0; // totals at the last decision that had enough new sequences behind it
    float hit_share = 0.0f; // failed / scanned over the sequences between the last two decisions
    // single-sequence latency path (msv_wave_kernels.cuh): a chain of warps over a thread-block cluster; built when
    // tr_E_C == tr_E_J (the kernel speculates B = N + move) and the chain fits a cluster
    struct Wave {
        int K = 0;               // columns per lane
        uint32_t warps = 0, ctas = 0;
        size_t shared_bytes = 0, shared_limit = 0;
        float* d_table = nullptr;
        msv::Wave_accumulator* d_accumulator = nullptr;
        msv::Wave_result* h_result = nullptr; // pinned + mapped: the kernel writes the result, the host polls it
        msv::Wave_result* d_result = nullptr; // the same memory as the device sees it
        uint8_t* d_residues = nullptr;        // sequences too long for the kernel parameters
        uint8_t* h_staging = nullptr;         // pinned
        size_t capacity = 0;
        cudaStream_t stream = nullptr;
        uint32_t tag = 0;
        const Wave_kernels* kernels = nullptr;
        // the communication-free variant (one thread per diagonal phase, msv_diag_kernel): the reference-layout table with
        // every row extended by its first four columns; used when table + 4 bytes per row fit one SM's shared memory
        float* d_diag_table = nullptr;
        uint32_t diag_period = 0, diag_ctas = 0;
        size_t diag_table_bytes = 0;
    } wave;
};

namespace {

constexpr uint32_t kBuckets = 1u << 16;

int db_release(msv_db* db) {
    if (!db) return MSV_OK;
    Device_guard guard(db->device);
    cudaFree(db->d_residues);
    cudaFree(db->d_offsets);
    cudaFree(db->d_order);
    cudaFree(db->d_scores);
    cudaFree(db->d_stats);
    cudaFree(db->d_length_tr);
    cudaFree(db->d_hist);
    cudaFree(db->d_queue);
    cudaFree(db->d_first_bad);
    if (db->filter_scratch && db->filter_scratch_free) db->filter_scratch_free(db->filter_scratch);
    if (db->copy_stream) {
        cudaStreamDestroy(db->copy_stream);
        cudaStreamDestroy(db->compute_stream);
        cudaStreamDestroy(db->compute_stream2);
        cudaEventDestroy(db->other_done);
        for (auto ev : db->stage_copied) cudaEventDestroy(ev);
        cudaEventDestroy(db->reserved);
    }
    delete db;
    return MSV_OK;
}

// Host pass over the offsets: monotonic, total and longest sequence.
int db_check_offsets(const uint8_t* residues, const uint64_t* offsets, size_t n, uint64_t* total_out, uint64_t* longest_out) {
    if (n > 0 && (!offsets || offsets[0] != 0)) return fail(MSV_ERR_INVALID_ARGUMENT, "offsets[0] must be 0");
    if (n >= (1ull << 32) - 1) return fail(MSV_ERR_INVALID_ARGUMENT, "more than 2^32-2 sequences in one database");
    // one sequential pass over the offsets (~0.4 ms per million sequences; measured on the B200 host, splitting it over
    // threads costs more in thread start-up than it saves)
    uint64_t longest = 0;
    for (size_t q = 0; q < n; ++q) {
        if (offsets[q + 1] < offsets[q]) return fail(MSV_ERR_INVALID_ARGUMENT, "offsets not monotonic at %zu", q);
        longest = std::max<uint64_t>(longest, offsets[q + 1] - offsets[q]);
    }
    const uint64_t total = n ? offsets[n] : 0;
    if (total > 0 && !residues) return fail(MSV_ERR_INVALID_ARGUMENT, "residues is NULL");
    if (longest >= (1ull << 27)) // the per-length transition table would exceed 1 GB; no protein comes close
        return fail(MSV_ERR_INVALID_ARGUMENT, "sequence of %llu residues exceeds the supported 2^27-1", static_cast<unsigned long long>(longest));
    *total_out = total;
    *longest_out = longest;
    return MSV_OK;
}

// Grow-only device buffers + the per-length (tr_loop, tr_move) table (host libm, reference MSV_HMM.cpp:59-64).
int db_reserve(msv_db* db, uint64_t total, size_t n, uint64_t longest, cudaStream_t stream) {
    const size_t need_res = static_cast<size_t>(total) + msv::kResiduePadBytes + 16;
    if (need_res > db->cap_residues) {
        cudaFree(db->d_residues);
        db->d_residues = nullptr;
        db->cap_residues = 0;
        const size_t cap = need_res + need_res / 8;
        MSV_CUDA_TRY(cudaMalloc(&db->d_residues, cap));
        db->cap_residues = cap;
    }
    if (n + 1 > db->cap_n) {
        cudaFree(db->d_offsets);
        cudaFree(db->d_order);
        cudaFree(db->d_scores);
        db->d_offsets = nullptr;
        db->d_order = nullptr;
        db->d_scores = nullptr;
        db->cap_n = 0;
        const size_t cap = n + 1 + n / 8;
        MSV_CUDA_TRY(cudaMalloc(&db->d_offsets, cap * sizeof(uint64_t)));
        MSV_CUDA_TRY(cudaMalloc(&db->d_order, cap * sizeof(uint32_t)));
        MSV_CUDA_TRY(cudaMalloc(&db->d_scores, cap * sizeof(float)));
        db->cap_n = cap;
    }
    if (!db->d_hist) {
        MSV_CUDA_TRY(cudaMalloc(&db->d_hist, 4 * kBuckets * sizeof(uint32_t)));
        MSV_CUDA_TRY(cudaMalloc(&db->d_queue, 2 * kMaxChunks * sizeof(unsigned int))); // per upload stage: queue head, head of its second part (next_ticket)
        MSV_CUDA_TRY(cudaMalloc(&db->d_first_bad, sizeof(unsigned long long)));
    }
    if (db->h_length_tr.size() < longest + 1) {
        const size_t from = db->h_length_tr.size();
        db->h_length_tr.resize(longest + 1);
        for (size_t L = from; L <= longest; ++L) {
            float lo, mv;
            msv_host_length_transitions(L, &lo, &mv);
            db->h_length_tr[L] = make_float2(lo, mv);
        }
    }
    if (db->h_length_tr.size() > db->cap_tr) {
        cudaFree(db->d_length_tr);
        db->d_length_tr = nullptr;
        db->cap_tr = 0;
        MSV_CUDA_TRY(cudaMalloc(&db->d_length_tr, db->h_length_tr.size() * sizeof(float2)));
        MSV_CUDA_TRY(cudaMemcpyAsync(db->d_length_tr, db->h_length_tr.data(), db->h_length_tr.size() * sizeof(float2),
                                     cudaMemcpyHostToDevice, stream));
        db->cap_tr = db->h_length_tr.size();
    }
    const unsigned long long none = ~0ull;
    MSV_CUDA_TRY(cudaMemcpyAsync(db->d_first_bad, &none, sizeof none, cudaMemcpyHostToDevice, stream));
    MSV_CUDA_TRY(cudaMemsetAsync(db->d_queue, 0, 2 * kMaxChunks * sizeof(unsigned int), stream));
    return MSV_OK;
}

// Validate the residue codes of sequences [first, first+count) and bucket them longest-first into
// d_order[first .. first+count) (indices relative to `first`).  Everything is queued on `stream`.
int db_prepare_range(msv_db* db, size_t first, size_t count, uint64_t residue_begin, uint64_t residue_end, uint64_t longest,
                     cudaStream_t stream, int scratch = 0) {
    uint32_t* const d_hist = db->d_hist + static_cast<size_t>(scratch) * 2 * kBuckets; // hist | cursor of this stream
    if (count == 0) return MSV_OK;
    if (residue_end > residue_begin) {
        // 16-byte words that cover the range; neighbouring ranges may be checked twice, which is harmless
        const uint64_t word_begin = residue_begin / 16, word_end = (residue_end + 15) / 16;
        const uint64_t words = word_end - word_begin;
        const int blocks = static_cast<int>(std::min<uint64_t>((words + 255) / 256, 148 * 16));
        msv::db_validate_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(db->d_residues) + word_begin, words,
                                                             residue_end - word_begin * 16, word_begin * 16, db->d_first_bad);
        ++g_launches;
    }
    if (count == 1) { // single-sequence calls: the order is trivially {0}
        MSV_CUDA_TRY(cudaMemsetAsync(db->d_order + first, 0, sizeof(uint32_t), stream));
        return MSV_OK;
    }
    uint32_t shift = 0;
    while ((longest >> shift) >= kBuckets) ++shift;
    const uint32_t used_buckets = static_cast<uint32_t>(std::min<uint64_t>((longest >> shift) + 1, kBuckets));
    MSV_CUDA_TRY(cudaMemsetAsync(d_hist, 0, used_buckets * sizeof(uint32_t), stream));
    const uint32_t n32 = static_cast<uint32_t>(count);
    const int blocks_n = static_cast<int>((count + 255) / 256);
    msv::db_histogram_kernel<<<blocks_n, 256, 0, stream>>>(db->d_offsets + first, n32, shift, used_buckets, d_hist);
    msv::db_scan_kernel<<<1, 1024, 0, stream>>>(d_hist, used_buckets, d_hist + kBuckets);
    msv::db_scatter_kernel<<<blocks_n, 256, 0, stream>>>(db->d_offsets + first, n32, shift, used_buckets, d_hist + kBuckets,
                                                         db->d_order + first);
    g_launches += 3;
    MSV_CUDA_TRY(cudaGetLastError());
    return MSV_OK;
}

int db_read_validation(msv_db* db, const uint8_t* residues, cudaStream_t stream) {
    unsigned long long first_bad = ~0ull;
    MSV_CUDA_TRY(cudaMemcpyAsync(&first_bad, db->d_first_bad, sizeof first_bad, cudaMemcpyDeviceToHost, stream));
    MSV_CUDA_TRY(cudaStreamSynchronize(stream));
    if (first_bad != ~0ull) {
        db->n = 0;
        return fail(MSV_ERR_BAD_RESIDUE, "residue code %u at position %llu is outside 0..19",
                    static_cast<unsigned>(residues[first_bad]), first_bad);
    }
    return MSV_OK;
}

// Sequences and rows per bucket of kProfileStep lengths for sequences [first, first + count) (see msv_db::h_profile_*).
struct Length_profile {
    const uint32_t* count = nullptr;
    const uint64_t* rows = nullptr;
};
void profile_lengths(const uint64_t* offsets, size_t first, size_t count, std::vector<uint32_t>& counts, std::vector<uint64_t>& rows) {
    counts.assign(kProfileBuckets, 0);
    rows.assign(kProfileBuckets, 0);
    for (size_t q = first; q < first + count; ++q) {
        const uint64_t len = offsets[q + 1] - offsets[q];
        const size_t bucket = static_cast<size_t>(std::min<uint64_t>(len / kProfileStep, kProfileBuckets - 1));
        ++counts[bucket];
        rows[bucket] += len;
    }
}

// Small databases keep their lengths on the host: the launch planner needs them to balance few long sequences.  Every
// database whose offsets pass through the host can keep its length profile.
void db_keep_lengths(msv_db* db, const uint64_t* offsets, size_t n) {
    db->h_lengths.clear();
    db->h_profile_count.clear();
    db->h_profile_rows.clear();
    if (n == 0) return;
    if (std::getenv("MSV_CUDA_FAST_CTAS")) profile_lengths(offsets, 0, n, db->h_profile_count, db->h_profile_rows); // (experiment, see plan_launch)
    if (n > 65536) return;
    db->h_lengths.resize(n);
    for (size_t q = 0; q < n; ++q) db->h_lengths[q] = static_cast<uint32_t>(offsets[q + 1] - offsets[q]);
}

// (Re)fill `db` from host buffers in one piece: upload, validate, per-length transitions, longest-first order.
// Returns after the validation result has been read back (one small synchronisation).
int db_fill(msv_db* db, const uint8_t* residues, const uint64_t* offsets, size_t n, cudaStream_t stream) {
    uint64_t total = 0, longest = 0;
    if (int rc = db_check_offsets(residues, offsets, n, &total, &longest)) return rc;
    if (int rc = db_reserve(db, total, n, longest, stream)) return rc;
    if (total) MSV_CUDA_TRY(cudaMemcpyAsync(db->d_residues, residues, total, cudaMemcpyHostToDevice, stream));
    MSV_CUDA_TRY(cudaMemsetAsync(db->d_residues + total, 0, msv::kResiduePadBytes + 16, stream));
    if (n) {
        MSV_CUDA_TRY(cudaMemcpyAsync(db->d_offsets, offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    } else {
        MSV_CUDA_TRY(cudaMemsetAsync(db->d_offsets, 0, sizeof(uint64_t), stream));
    }
    db->n = n;
    db->total = total;
    db->longest = longest;
    db_keep_lengths(db, offsets, n);
    if (n == 0) return MSV_OK;
    if (int rc = db_prepare_range(db, 0, n, 0, total, longest, stream)) return rc;
    return db_read_validation(db, residues, stream);
}

} // namespace

namespace msv_detail {
int db_refill(msv_db* db, const uint8_t* residues, const uint64_t* offsets, size_t n) { return db_fill(db, residues, offsets, n, nullptr); }
int db_free(msv_db* db) { return db_release(db); }
int db_reserve_for(msv_db* db, uint64_t total, size_t n, uint64_t longest, cudaStream_t stream) { return db_reserve(db, total, n, longest, stream); }
int db_bucket(msv_db* db, cudaStream_t stream) { return db_prepare_range(db, 0, db->n, 0, 0, db->longest, stream); }
} // namespace msv_detail

namespace {

// ---- launch planning ------------------------------------------------------------------------------------------------
// With millions of sequences every plan is balanced and the choice is static.  With few (long) sequences the scan is a
// scheduling problem: sequences are serial, a slot (warp / four warps) scans one at a time, so the makespan is set by
// the longest-processing-time-first assignment of whole sequences to slots -- and fewer, faster slots can win.
// The planner simulates that assignment for each candidate (plan, slots per SM) and weighs it with the measured
// throughput of that kernel at that occupancy (B200, profiles/r01/occupancy_curve_*.jsonl, relative to the warp kernel
// at full occupancy).
struct Launch_plan {
    const msv_model::Plan* plan;
    size_t slots_per_cta;
    // lane-group plans: the n_long longest sequences go to fast_ctas CTAs that run with fast_slots slots (next_ticket)
    uint32_t n_long = 0, fast_ctas = 0, fast_slots = 0;
};

// throughput of the warp kernel at `fraction` of its maximum warps per SM (measured: 25 % -> 0.51, 50 % -> 0.885, 75 % -> 0.98)
double warp_occupancy_factor(double fraction) {
    static const double x[] = {0.0, 0.25, 0.5, 0.75, 1.0}, y[] = {0.0, 0.51, 0.885, 0.98, 1.0};
    for (int i = 1; i < 5; ++i)
        if (fraction <= x[i]) return y[i - 1] + (y[i] - y[i - 1]) * (fraction - x[i - 1]) / (x[i] - x[i - 1]);
    return 1.0;
}

// throughput of the four-warp kernel with `groups` of `max_groups` per SM (measured at K = 20: 3.4 ... 7.0 TCUPS vs 9.2)
double quad_occupancy_factor(size_t groups, size_t max_groups) {
    static const double x[] = {0.0, 1.0 / 6, 2.0 / 6, 3.0 / 6, 4.0 / 6, 5.0 / 6, 1.0}, y[] = {0.0, 0.37, 0.61, 0.685, 0.724, 0.745, 0.758};
    const double fraction = static_cast<double>(groups) / static_cast<double>(max_groups);
    for (int i = 1; i < 7; ++i)
        if (fraction <= x[i] + 1e-9) return y[i - 1] + (y[i] - y[i - 1]) * (fraction - x[i - 1]) / (x[i] - x[i - 1]);
    return y[6];
}

// rows of the busiest slot when `lengths` (sorted, longest first) are handed to `slots` workers greedily
uint64_t lpt_makespan(const std::vector<uint32_t>& lengths, size_t slots) {
    if (lengths.empty()) return 0;
    if (slots >= lengths.size()) return lengths.front();
    std::vector<uint64_t> heap(slots, 0); // min-heap of slot loads
    for (const uint32_t len : lengths) {
        std::pop_heap(heap.begin(), heap.end(), std::greater<uint64_t>());
        heap.back() += len;
        std::push_heap(heap.begin(), heap.end(), std::greater<uint64_t>());
    }
    return *std::max_element(heap.begin(), heap.end());
}

// Lane-group plans: which sequences are too long for a slot of a full CTA, and how many fast CTAs they need.
// A slot of a full CTA scans `average` rows in the time of the whole launch; a slot of a fast CTA (warps_fast of
// warps_full warps, measured throughput `relative`) scans relative * warps_full / warps_fast times as many.
bool plan_fast_ctas(const Length_profile& profile, uint64_t residues, uint64_t longest, size_t sms, size_t slots_full, size_t per_warp,
                    Launch_plan& plan) {
    const uint64_t average = residues / (sms * slots_full); // rows per slot of a balanced launch
    const size_t warps_full = slots_full / per_warp;
    int want_ctas = 0, want_warps = 0, want_cut = 0; // MSV_CUDA_FAST_CTAS="ctas,warps,rows", or "auto" for the model below
    if (const char* env = std::getenv("MSV_CUDA_FAST_CTAS")) std::sscanf(env, "%d,%d,%d", &want_ctas, &want_warps, &want_cut);
    if (!want_ctas && 10 * average >= 12 * std::max<uint64_t>(longest, 1)) return true; // nothing is too long: plain queue, all slots
    for (const size_t warps_fast : {size_t(8), size_t(4)}) {
        if (want_warps ? warps_fast != static_cast<size_t>(want_warps) : warps_fast * 2 > warps_full) continue;
        // throughput of a CTA at that occupancy relative to a full one (B200, profiles/r02/short_model_sweep_v1.jsonl:
        // 8 warps 0.69..0.90, 4 warps about half)
        const double relative = warps_fast == 8 ? 0.75 : 0.5;
        const double speedup = relative * static_cast<double>(warps_full) / static_cast<double>(warps_fast);
        if (!want_warps && static_cast<double>(longest) > 0.95 * speedup * static_cast<double>(average)) continue;
        // long = at least 0.8 x the rows of an average slot, at a bucket boundary of the profile
        const uint64_t cut = want_cut ? static_cast<uint64_t>(want_cut) / kProfileStep * kProfileStep
                                      : std::max<uint64_t>(2 * kProfileStep, average * 8 / 10 / kProfileStep * kProfileStep);
        if (cut / kProfileStep >= kProfileBuckets) continue;
        uint64_t n_long = 0, rows_long = 0;
        for (size_t b = cut / kProfileStep; b < kProfileBuckets; ++b) n_long += profile.count[b], rows_long += profile.rows[b];
        if (n_long == 0) return true;
        const double share = static_cast<double>(rows_long) / static_cast<double>(residues);
        const size_t ctas = want_ctas ? static_cast<size_t>(want_ctas)
                                      : static_cast<size_t>(std::ceil(1.15 * share / relative * static_cast<double>(sms)));
        if (ctas < 1 || 2 * ctas > sms) continue;
        plan.n_long = static_cast<uint32_t>(n_long);
        plan.fast_ctas = static_cast<uint32_t>(ctas);
        plan.fast_slots = static_cast<uint32_t>(warps_fast * per_warp);
        return true;
    }
    return false;
}

Launch_plan plan_launch(const msv_model* model, const msv_db* db, size_t first, size_t count, uint64_t residues,
                        const Length_profile* profile = nullptr) {
    const auto max_slots = [](const msv_model::Plan& plan) { return static_cast<size_t>(plan.geo->threads / plan.geo->G); };
    if (!model->bulk.geo) return {&model->quad, max_slots(model->quad)};
    if (model->forced) return {&model->bulk, max_slots(model->bulk)};
    const size_t sms = static_cast<size_t>(model->sm_count);
    const size_t bulk_slots = sms * max_slots(model->bulk);

    // plenty of sequences: static choice (the eight-lane plan needs about as many rows per slot as the longest sequence
    // has, because it has four times more slots -- measured break-even, profiles/r01/sweep_generic_v3.jsonl)
    // (planning costs ~1 ms of host time at most; it pays only when the sequences are long enough for balance to matter)
    const bool few = count < 4 * bulk_slots && residues >= 1000 * static_cast<uint64_t>(count) && db->h_lengths.size() == db->n &&
                     model->bulk.geo->G == 32;
    if (!few) {
        // Lane-group plans (short models) have 4x / 8x more slots than the warp plan, and a slot scans a sequence at a
        // quarter / an eighth of a warp's speed: with few sequences per slot the scan ends when the longest sequence does.
        // So the slot count is CUT (fewer CTA threads: fewer, faster slots) until every slot gets 1.2x the rows of the longest
        // sequence, down to a third of the maximum; the SM keeps most of its throughput on the way.  Measured on B200
        // (profiles/r02/short_model_sweep_v1.jsonl, 100.hmm x 100 k sequences, four lanes per sequence): 192 slots per CTA
        // 5.0 TCUPS, 96: 6.3, 64: 6.6, 48: 5.9; with 1 M sequences every slot count is balanced and the maximum wins (7.5).
        // Four lanes per sequence beat eight wherever both exist (100.hmm 7.5 vs 6.2, 200.hmm 8.4 vs 7.6 TCUPS at 1 M).
        // Round 2 experiment, NOT the default: keep every slot on most CTAs and hand the longest sequences to a few fast
        // CTAs instead (plan_fast_ctas / next_ticket).  Measured on B200 over 60 settings (profiles/r02/fast_cta_sweep_v1..3.jsonl,
        // 100..400.hmm x 100 k sequences): within -9 .. +3 % of the slot cut and never clearly ahead -- what the full CTAs
        // lose at this database size is not only the longest sequences.  MSV_CUDA_FAST_CTAS="ctas,warps,rows" (or "auto")
        // switches it on for further tuning; results are the same bits either way (test_long_sequences_on_fast_ctas).
        for (const msv_model::Plan* plan : {&model->narrow, &model->octet}) {
            if (!plan->geo) continue;
            const size_t per_warp = 32 / static_cast<size_t>(plan->geo->G);
            const size_t most = max_slots(*plan), least = std::max(per_warp, most / 3 / per_warp * per_warp);
            if (profile && profile->count && count >= 2 * sms * most && std::getenv("MSV_CUDA_FAST_CTAS")) {
                Launch_plan chosen{plan, most};
                if (plan_fast_ctas(*profile, residues, db->longest, sms, most, per_warp, chosen)) return chosen;
            }
            for (size_t slots = most; slots >= least; slots -= per_warp)
                if (10 * (residues / (sms * slots)) >= 12 * std::max<uint64_t>(db->longest, 1)) return {plan, slots};
        }
        if (model->quad.geo && count < 2 * bulk_slots) return {&model->quad, max_slots(model->quad)};
        return {&model->bulk, max_slots(model->bulk)};
    }

    // few sequences: pick (plan, slots per SM) by simulated makespan x measured per-slot speed
    std::vector<uint32_t> lengths(db->h_lengths.begin() + static_cast<std::ptrdiff_t>(first),
                                  db->h_lengths.begin() + static_cast<std::ptrdiff_t>(first + count));
    std::sort(lengths.begin(), lengths.end(), std::greater<uint32_t>());
    Launch_plan best{&model->bulk, max_slots(model->bulk)};
    double best_cost = std::numeric_limits<double>::infinity();
    const auto consider = [&](const msv_model::Plan& plan, size_t slots_per_sm, double relative_throughput) {
        const size_t ctas = std::min(sms, (count + slots_per_sm - 1) / slots_per_sm);
        const double cost = static_cast<double>(lpt_makespan(lengths, ctas * slots_per_sm)) * static_cast<double>(slots_per_sm) /
                            relative_throughput;
        if (cost < best_cost) {
            best_cost = cost;
            best = {&plan, slots_per_sm};
        }
    };
    const size_t max_warps = max_slots(model->bulk);
    for (size_t warps = 4; warps <= max_warps; warps += 4) // whole warps per scheduler: 4 SM sub-partitions
        consider(model->bulk, warps, warp_occupancy_factor(static_cast<double>(warps) / static_cast<double>(max_warps)));
    if (model->quad.geo)
        for (size_t groups = 1; groups <= max_slots(model->quad); ++groups)
            consider(model->quad, groups, quad_occupancy_factor(groups, max_slots(model->quad)));
    return best;
}

constexpr bool kGatherPushByDefault = false; // see launch_scan

// Row variant of the warp kernel for a launch (only where the speculating kernels exist, i.e. quiet_rows(K) != kExact).
Rows rows_for(int K, float hit_share, bool long_sequences) {
    const Rows quiet = quiet_rows(K);
    const bool blocks_are_cheap = quiet == kBlocks || K < 24; // measured: block-wise speculation is at least as fast as exact rows there
    if (long_sequences) return blocks_are_cheap ? kBlocks : kExact;
    if (hit_share <= 0.03f) return quiet;
    return (blocks_are_cheap && hit_share < 0.2f) ? kBlocks : kExact;
}

// One launch of the scan over sequences [first, first+count) of `db`; scores go to d_scores[first ..).
Length_profile whole_profile(const msv_db* db) { // of the whole database, when its offsets passed through the host
    if (db->h_profile_count.size() != kProfileBuckets) return {};
    return {db->h_profile_count.data(), db->h_profile_rows.data()};
}

int launch_scan(msv_model* model, msv_db* db, size_t first, size_t count, uint64_t residues, int queue_slot, float* d_scores,
                cudaStream_t stream, float* const* mirrors = nullptr, int n_mirrors = 0, const Length_profile* stage_profile = nullptr) {
    if (count == 0) return MSV_OK;
    const Length_profile whole = (first == 0 && count == db->n) ? whole_profile(db) : Length_profile{};
    const Launch_plan chosen = plan_launch(model, db, first, count, residues, stage_profile ? stage_profile : &whole);
    const msv_model::Plan& plan = *chosen.plan;
    const Geometry* geo = plan.geo;
    msv::Scan_params p{};
    p.table = plan.d_table;
    p.residues = db->d_residues;
    p.offsets = db->d_offsets + first;
    p.order = db->d_order + first;
    p.length_tr = db->d_length_tr;
    p.scores = d_scores + first;
    p.queue_head = db->d_queue + 2 * queue_slot;
    p.first_bad = db->d_first_bad;
    p.n = static_cast<uint32_t>(count);
    p.table_bytes = static_cast<uint32_t>(plan.shared_bytes);
    p.tr_B_Mk = model->tr_B_Mk;
    p.tr_E_C = model->tr_E_C;
    p.tr_E_J = model->tr_E_J;
    // Gather of a sharded run (mirrors = the peers' copies of the job's score array).  Two forms, same result:
    //   stores : every score is stored into all copies by the lane that computed it (store_score) -- no second kernel, the
    //            transfer rides along with the scan; n_mirrors remote 4-byte stores per sequence;
    //   push   : the scan writes this GPU's copy only, and a small kernel behind it copies the slice into the peers' copies
    //            with coalesced stores (score_push_kernel).
    // MSV_CUDA_GATHER=stores|push overrides the default.
    const char* gather_env = std::getenv("MSV_CUDA_GATHER");
    const bool push = n_mirrors > 0 && (gather_env ? std::strcmp(gather_env, "push") == 0 : kGatherPushByDefault);
    p.n_mirrors = push ? 0u : static_cast<uint32_t>(n_mirrors);
    for (int r = 0; r < n_mirrors; ++r) p.mirrors[r] = mirrors[r] + first;
    const auto push_slice = [&]() -> int { // this launch's slice is contiguous: sequences [first, first + count)
        if (!push) return MSV_OK;
        msv::Score_mirrors peers{};
        for (int r = 0; r < n_mirrors; ++r) peers.copy[r] = mirrors[r] + first;
        const int blocks = static_cast<int>(std::min<size_t>(static_cast<size_t>(model->sm_count), (count + 1023) / 1024));
        msv::score_push_kernel<<<blocks, 256, 0, stream>>>(d_scores + first, peers, static_cast<uint32_t>(n_mirrors), count);
        ++g_launches;
        MSV_CUDA_TRY(cudaGetLastError());
        return MSV_OK;
    };
    MSV_CUDA_TRY(cudaMemsetAsync(db->d_queue + 2 * queue_slot, 0, 2 * sizeof(unsigned int), stream));
    // persistent CTAs, at most one per SM; a "slot" scans one sequence at a time (lane group, warp or four warps)
    const size_t threads_per_slot = static_cast<size_t>(geo->G);
    size_t slots = std::min<size_t>(chosen.slots_per_cta, geo->threads / threads_per_slot);
    if (const char* env = std::getenv(geo->G == 128 ? "MSV_CUDA_QUAD_GROUPS" : "MSV_CUDA_BULK_SLOTS")) // tuning aid
        slots = std::min<size_t>(geo->threads / threads_per_slot, std::max(1, std::atoi(env)));
    const size_t ctas = std::max<size_t>(1, std::min<size_t>(model->sm_count, (count + slots - 1) / slots));
    if (ctas == 1) slots = std::min(slots, count); // do not launch slots that would find the queue empty
    const int threads = static_cast<int>(std::max<size_t>(32, (slots * threads_per_slot + 31) / 32 * 32));
    const bool cj_same = std::memcmp(&model->tr_E_C, &model->tr_E_J, sizeof(float)) == 0;
    const bool group_speculation = geo->fn_group_spec && cj_same && residues / count <= 2048 && count >= 4096 &&
                                   !std::getenv("MSV_CUDA_NO_SPECULATION");
    if (chosen.fast_ctas && chosen.fast_ctas < ctas && chosen.fast_slots < slots) {
        p.n_long = chosen.n_long;
        p.fast_ctas = chosen.fast_ctas;
        p.fast_threads = static_cast<uint32_t>(chosen.fast_slots * threads_per_slot);
    }
    if (group_speculation) { // lane groups with speculative rows; a sequence whose speculation fails is repeated exactly in the kernel
        geo->fn_group_spec<<<static_cast<int>(ctas), threads, plan.shared_bytes, stream>>>(p);
        ++g_launches;
        MSV_CUDA_TRY(cudaGetLastError());
        return push_slice();
    }
    Scan_kernel kernel = cj_same ? geo->fn_cj_same : geo->fn;
    // Three row variants of the warp kernel (msv_kernels.cuh), same bits: exact rows; speculation on whole sequences (vote at
    // the end, a failed sequence is scanned again: one extra pass per hit, T = W / (1 + 1.09 f) for a share f of failing
    // sequences); speculation in checkpointed blocks of 64 rows (a hit costs one block: T = B / (1 + 0.25 f)).  Which is
    // fastest depends on K (quiet_rows) and on f (profiles/r02/hit_rate_modes_v1.jsonl, 1400.hmm at f = 0 / 0.02 / 0.1 / 0.5:
    // whole 10.07 / 9.85 / 9.08 / 6.53, blocks 9.80 / 9.75 / 9.56 / 8.73, exact 9.78 throughout).  Every variant counts the
    // sequences that fail (or would fail) the vote; the totals travel to pinned host memory after each launch and the next
    // launch reads them without waiting: a database that turns out hit-rich gets exact rows from the second scan on (blocks
    // where blocks are the fastest variant anyway and hits are not the majority), long sequences never speculate as a whole.
    // MSV_CUDA_SPECULATION=whole|blocks|none overrides (tuning aid; MSV_CUDA_NO_SPECULATION is the older spelling of none).
    bool feedback = false;
    if (cj_same && geo->fn_cj_same_exact) {
        const char* env = std::getenv("MSV_CUDA_SPECULATION");
        const std::string forced_mode = std::getenv("MSV_CUDA_NO_SPECULATION") ? "none" : env ? env : "";
        if (model->h_speculation_totals && &plan == &model->bulk) {
            feedback = true;
            const unsigned int failures = model->h_speculation_totals[0], offered = model->h_speculation_totals[1];
            if (offered - model->seen_offered >= 4096u) { // unsigned differences: the totals may wrap
                model->hit_share = static_cast<float>(failures - model->seen_failures) / static_cast<float>(offered - model->seen_offered);
                model->seen_failures = failures;
                model->seen_offered = offered;
            }
        }
        Rows rows = rows_for(geo->K, model->hit_share, residues / count > 1024);
        if (forced_mode == "none") rows = kExact;
        else if (forced_mode == "blocks") rows = kBlocks;
        else if (forced_mode == "whole") rows = kWhole;
        kernel = rows == kExact ? geo->fn_cj_same_exact : rows == kBlocks ? geo->fn_cj_same_blocks : geo->fn_cj_same;
    }
    if (feedback) p.speculation_failures = model->d_speculation_totals;
    kernel<<<static_cast<int>(ctas), threads, plan.shared_bytes, stream>>>(p);
    ++g_launches;
    MSV_CUDA_TRY(cudaGetLastError());
    if (int rc = push_slice()) return rc;
    if (feedback)
        MSV_CUDA_TRY(cudaMemcpyAsync(const_cast<unsigned int*>(model->h_speculation_totals), model->d_speculation_totals,
                                     2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    return MSV_OK;
}

// End-to-end batch with the upload hidden behind the scan: the database is cut into contiguous stages of sequences;
// stage s+1 is copied to the device (copy stream) while stage s is validated, bucketed and scanned (compute stream).
// Scores go to `scores_host` (when not NULL) and stay in `d_scores_out` (NULL: the workspace's own buffer); `mirrors` are the
// peers' copies of a fused gather (launch_scan).  Synchronous: returns after the validation verdict has been read back.
int score_batch_pipelined(msv_model* model, msv_db* db, const uint8_t* residues, const uint64_t* offsets, size_t n,
                          float* scores_host, float* d_scores_out = nullptr, float* const* mirrors = nullptr, int n_mirrors = 0) {
    uint64_t total = 0, longest = 0;
    if (int rc = db_check_offsets(residues, offsets, n, &total, &longest)) return rc;
    if (!db->copy_stream) {
        MSV_CUDA_TRY(cudaStreamCreateWithFlags(&db->copy_stream, cudaStreamNonBlocking));
        MSV_CUDA_TRY(cudaStreamCreateWithFlags(&db->compute_stream, cudaStreamNonBlocking));
        MSV_CUDA_TRY(cudaStreamCreateWithFlags(&db->compute_stream2, cudaStreamNonBlocking));
        for (auto& ev : db->stage_copied) MSV_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        MSV_CUDA_TRY(cudaEventCreateWithFlags(&db->reserved, cudaEventDisableTiming));
        MSV_CUDA_TRY(cudaEventCreateWithFlags(&db->other_done, cudaEventDisableTiming));
    }
    cudaStream_t copy = db->copy_stream, compute = db->compute_stream;
    // Stages alternate between two compute streams: the persistent scan kernel of a stage ends with a tail in which most SMs
    // are already idle (every CTA waits for its last warp), and with the next stage on another stream its bucketing kernels
    // and scan CTAs move onto those SMs as they become free instead of waiting for the last CTA of the stage before.
    // MSV_CUDA_ONE_COMPUTE_STREAM=1 keeps everything on one stream (tuning aid).
    cudaStream_t const lanes[2] = {compute, std::getenv("MSV_CUDA_ONE_COMPUTE_STREAM") ? compute : db->compute_stream2};
    if (int rc = db_reserve(db, total, n, longest, compute)) return rc;
    MSV_CUDA_TRY(cudaEventRecord(db->reserved, compute));
    MSV_CUDA_TRY(cudaStreamWaitEvent(copy, db->reserved, 0)); // buffers may have been reallocated
    if (lanes[1] != compute) MSV_CUDA_TRY(cudaStreamWaitEvent(lanes[1], db->reserved, 0));
    db->n = n;
    db->total = total;
    db->longest = longest;
    db_keep_lengths(db, offsets, n);
    if (n == 0) return MSV_OK;
    float* const d_out = d_scores_out ? d_scores_out : db->d_scores;

    // Stages grow geometrically: the first one is small so that the scan starts almost immediately, and every later
    // upload (tens of GB/s over PCIe/C2C) finishes long before the scan of the stage before it (a few GB/s of residues).
    size_t bounds[kMaxChunks + 1];
    int stages = 0;
    bounds[0] = 0;
    {
        // few sequences: one stage, so that the launch planner can balance them all at once (the upload is short anyway)
        const size_t warp_slots = static_cast<size_t>(model->sm_count) * 16;
        // first stage: 8 MB (~23 k sequences).  2 MB (round 1) started the scan 0.1 ms earlier, but a 6 k-sequence launch on 2368
        // warp slots ends when its longest sequence does -- ~0.8 ms for 0.28 ms of work.  MSV_CUDA_FIRST_STAGE_MB overrides.
        uint64_t first_mb = 8;
        if (const char* env = std::getenv("MSV_CUDA_FIRST_STAGE_MB")) first_mb = std::max(1, std::atoi(env));
        uint64_t stage_bytes = first_mb << 20, cut = 0;
        // every stage must be uploaded before the scan of the stage before it ends: the scan consumes ~8e12 / LENG bytes per
        // second, the link delivers ~45e9, so a stage may be LENG / 220 times the size of its predecessor (at most 4x, at least
        // 1.5x: short models are upload-bound whatever the stages are).  MSV_CUDA_STAGE_GROWTH overrides.
        double growth = std::min(4.0, std::max(1.5, static_cast<double>(model->model_length - 1) / 220.0));
        if (const char* env = std::getenv("MSV_CUDA_STAGE_GROWTH")) growth = std::max(1.1, std::atof(env));
        while (n >= 4 * warp_slots && stages < kMaxChunks - 1 && cut + stage_bytes + (stage_bytes >> 1) < total) {
            cut += stage_bytes;
            size_t q = static_cast<size_t>(std::lower_bound(offsets, offsets + n + 1, cut) - offsets);
            q = std::min(q, n);
            if (q > bounds[stages]) bounds[++stages] = q;
            stage_bytes = static_cast<uint64_t>(static_cast<double>(stage_bytes) * growth);
        }
        if (bounds[stages] < n || stages == 0) bounds[++stages] = n;
    }

    std::vector<uint32_t> stage_counts;
    std::vector<uint64_t> stage_rows;
    const auto enqueue = [&]() -> int {
        MSV_CUDA_TRY(cudaMemsetAsync(db->d_residues + total, 0, msv::kResiduePadBytes + 16, copy));
        for (int s = 0; s < stages; ++s) {
            const size_t first = bounds[s], last = bounds[s + 1];
            const uint64_t begin = offsets[first], end = offsets[last];
            // the stage's own offsets travel with it (8 MB for a million sequences would otherwise delay the first scan by 0.15 ms)
            MSV_CUDA_TRY(cudaMemcpyAsync(db->d_offsets + first, offsets + first, (last - first + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, copy));
            if (end > begin)
                MSV_CUDA_TRY(cudaMemcpyAsync(db->d_residues + begin, residues + begin, end - begin, cudaMemcpyHostToDevice, copy));
            cudaStream_t const lane = lanes[s & 1];
            MSV_CUDA_TRY(cudaEventRecord(db->stage_copied[s], copy));
            MSV_CUDA_TRY(cudaStreamWaitEvent(lane, db->stage_copied[s], 0));
            if (int rc = db_prepare_range(db, first, last - first, begin, end, longest, lane, s & 1)) return rc;
            Length_profile stage_profile;
            if (stages > 1 && (model->narrow.geo || model->octet.geo) && std::getenv("MSV_CUDA_FAST_CTAS")) { // (experiment, see plan_launch)
                profile_lengths(offsets, first, last - first, stage_counts, stage_rows);
                stage_profile = {stage_counts.data(), stage_rows.data()};
            }
            if (int rc = launch_scan(model, db, first, last - first, end - begin, s, d_out, lane, mirrors, n_mirrors,
                                     stage_profile.count ? &stage_profile : nullptr))
                return rc;
        }
        if (lanes[1] != compute) { // everything the second stream did happens before the download and the verdict
            MSV_CUDA_TRY(cudaEventRecord(db->other_done, lanes[1]));
            MSV_CUDA_TRY(cudaStreamWaitEvent(compute, db->other_done, 0));
        }
        if (scores_host) MSV_CUDA_TRY(cudaMemcpyAsync(scores_host, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, compute));
        return db_read_validation(db, residues, compute);
    };
    const int rc = enqueue();
    if (rc != MSV_OK) {
        // copies that still read the caller's buffers (and kernels that use the workspace) may be queued: drain both streams
        // before handing the buffers back, and leave the workspace empty
        const std::string message = g_last_error;
        (void)cudaStreamSynchronize(copy);
        (void)cudaStreamSynchronize(compute);
        (void)cudaStreamSynchronize(db->compute_stream2);
        (void)cudaGetLastError();
        db->n = 0;
        g_last_error = message;
    }
    return rc;
}

} // namespace

// ---- single-sequence latency path -----------------------------------------------------------------------------------
namespace {

void wave_release(msv_model* model) {
    auto& w = model->wave;
    cudaFree(w.d_table);
    cudaFree(w.d_diag_table);
    cudaFree(w.d_accumulator);
    cudaFree(w.d_residues);
    if (w.h_result) cudaFreeHost(w.h_result);
    if (w.h_staging) cudaFreeHost(w.h_staging);
    if (w.stream) cudaStreamDestroy(w.stream);
    w = msv_model::Wave();
}

cudaError_t wave_build(msv_model* model, const float* emission_scores, size_t columns) {
    auto& w = model->wave;
    const Wave_kernels* kernels = choose_wave_kernels(columns);
    if (!kernels || columns == 0) return cudaErrorInvalidConfiguration;
    const int K = kernels->K;
    const size_t warps = (columns + 32 * static_cast<size_t>(K) - 1) / (32 * static_cast<size_t>(K));
    size_t ctas = 1;
    while (ctas * msv::kWaveWarpsPerCta < warps) ctas *= 2; // cluster sizes 1, 2, 4, 8
    // table: [warp][residue][pair][lane][2]; lane l of warp g owns model columns (g*32 + l)*K + 1 ... + K, -inf beyond the model
    const size_t floats = ctas * msv::kWaveWarpsPerCta * MSV_ALPHABET * static_cast<size_t>(K) * 32;
    std::vector<float> laid(floats, -std::numeric_limits<float>::infinity());
    const size_t model_length = columns + 1;
    for (size_t g = 0; g < warps; ++g)
        for (int res = 0; res < MSV_ALPHABET; ++res)
            for (int lane = 0; lane < 32; ++lane)
                for (int j = 0; j < K; ++j) {
                    const size_t col = (g * 32 + lane) * K + j + 1;
                    if (col > columns) continue;
                    laid[(((g * MSV_ALPHABET + res) * (K / 2) + j / 2) * 32 + lane) * 2 + j % 2] = emission_scores[res * model_length + col];
                }
    w.K = K;
    w.warps = static_cast<uint32_t>(warps);
    w.ctas = static_cast<uint32_t>(ctas);
    w.shared_bytes = static_cast<size_t>(msv::kWaveWarpsPerCta) * MSV_ALPHABET * K * 128; // the table slice; + 4 bytes per row of the sequence
    int optin = 0;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, model->device) != cudaSuccess || optin < 65536) return cudaErrorInvalidConfiguration;
    w.shared_limit = static_cast<size_t>(optin) - 4096; // static shared memory of the kernel (mailboxes) + slack
    if (w.shared_limit < w.shared_bytes + 4096) return cudaErrorInvalidConfiguration;
    w.kernels = kernels;
    cudaError_t err = cudaMalloc(&w.d_table, floats * sizeof(float));
    if (err == cudaSuccess) err = cudaMemcpy(w.d_table, laid.data(), floats * sizeof(float), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMalloc(&w.d_accumulator, sizeof(msv::Wave_accumulator));
    if (err == cudaSuccess) err = cudaMemset(w.d_accumulator, 0, sizeof(msv::Wave_accumulator));
    if (err == cudaSuccess) err = cudaHostAlloc(&w.h_result, sizeof(msv::Wave_result), cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess) {
        std::memset(w.h_result, 0, sizeof(msv::Wave_result));
        err = cudaHostGetDevicePointer(&w.d_result, w.h_result, 0);
    }
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking);
    // the diagonal-worker kernel: [residue][P + 4] floats, P = model_length; it needs 80 (P + 4) bytes of shared memory
    const size_t diag_bytes = static_cast<size_t>(MSV_ALPHABET) * (model_length + 4) * sizeof(float);
    if (err == cudaSuccess && model_length >= 8 && diag_bytes + 4096 <= w.shared_limit && !std::getenv("MSV_CUDA_NO_DIAGONAL")) {
        std::vector<float> extended(static_cast<size_t>(MSV_ALPHABET) * (model_length + 4));
        for (int res = 0; res < MSV_ALPHABET; ++res)
            for (size_t col = 0; col < model_length + 4; ++col)
                extended[res * (model_length + 4) + col] = emission_scores[res * model_length + col % model_length];
        err = cudaMalloc(&w.d_diag_table, diag_bytes);
        if (err == cudaSuccess) err = cudaMemcpy(w.d_diag_table, extended.data(), diag_bytes, cudaMemcpyHostToDevice);
        for (const void* fn : {reinterpret_cast<const void*>(msv::msv_diag_kernel<true>), reinterpret_cast<const void*>(msv::msv_diag_kernel<false>)})
            if (err == cudaSuccess) err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(w.shared_limit));
        w.diag_period = static_cast<uint32_t>(model_length);
        w.diag_table_bytes = diag_bytes;
        const size_t diag_warps = (model_length + 31) / 32;
        w.diag_ctas = static_cast<uint32_t>((diag_warps + msv::kWaveWarpsPerCta - 1) / msv::kWaveWarpsPerCta);
    }
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernels->with_inline_residues), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(w.shared_limit));
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernels->with_device_residues), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(w.shared_limit));
    return err;
}

// Scores one sequence with the wavefront kernel.  *rescore is set when the kernel's speculation did not hold (the sequence
// contains a real hit): the caller then runs the exact kernel.
int wave_score(msv_model* model, const uint8_t* residues, size_t length, float* score, bool* rescore) {
    auto& w = model->wave;
    *rescore = false;
    const size_t row_bytes = (length + 3) / 4 * 16 + 32; // one word per row of the sequence, next to the table
    // first choice: one thread per diagonal phase, no communication (needs the whole table in one SM's shared memory)
    const bool diagonal = w.d_diag_table && w.diag_table_bytes + row_bytes <= w.shared_limit;
    // every CTA of the chain kernel keeps the same words next to its table slice; a sequence too long for that (tens of
    // thousands of residues) goes to the four-warp kernel
    const size_t shared_bytes = diagonal ? w.diag_table_bytes + row_bytes : w.shared_bytes + row_bytes;
    if (shared_bytes > w.shared_limit) {
        *rescore = true;
        return MSV_OK;
    }
    msv::Wave_params p{};
    p.table = w.d_table;
    p.accumulator = w.d_accumulator;
    p.result = w.d_result;
    p.length = static_cast<uint32_t>(length);
    p.warps = w.warps;
    p.tag = ++w.tag ? w.tag : ++w.tag; // never 0: that is what an untouched result slot holds
    p.tr_B_Mk = model->tr_B_Mk;
    p.tr_E_J = model->tr_E_J;
    msv_host_length_transitions(length, &p.loop, &p.move); // host libm, reference MSV_HMM.cpp:59-64

    cudaLaunchConfig_t config{};
    config.gridDim = dim3(w.ctas);
    config.blockDim = dim3(msv::kWaveWarpsPerCta * 32);
    config.dynamicSmemBytes = shared_bytes;
    config.stream = w.stream;
    cudaLaunchAttribute cluster{};
    cluster.id = cudaLaunchAttributeClusterDimension;
    cluster.val.clusterDim.x = w.ctas;
    cluster.val.clusterDim.y = 1;
    cluster.val.clusterDim.z = 1;
    config.attrs = &cluster;
    config.numAttrs = 1;

    msv::Diag_params dp{};
    if (diagonal) {
        dp.table = w.d_diag_table;
        dp.accumulator = w.d_accumulator;
        dp.result = w.d_result;
        dp.length = p.length;
        dp.period = w.diag_period;
        dp.tag = p.tag;
        dp.tr_B_Mk = p.tr_B_Mk;
        dp.tr_E_J = p.tr_E_J;
        dp.loop = p.loop;
        dp.move = p.move;
        config.gridDim = dim3(w.diag_ctas);
        config.numAttrs = 0; // independent CTAs: no cluster
    }
    if (length <= msv::kWaveInlineBytes) {
        msv::Wave_inline_residues block; // the sequence rides in the kernel parameters: one launch, no copy
        block.words[(length ? length - 1 : 0) / 4] = 0;
        if (length) std::memcpy(block.words, residues, length);
        if (diagonal) MSV_CUDA_TRY(cudaLaunchKernelEx(&config, msv::msv_diag_kernel<true>, dp, block));
        else MSV_CUDA_TRY(cudaLaunchKernelEx(&config, w.kernels->with_inline_residues, p, block));
    } else {
        const size_t padded = (length + 3) / 4 * 4;
        if (padded > w.capacity) {
            cudaFree(w.d_residues);
            if (w.h_staging) cudaFreeHost(w.h_staging);
            w.d_residues = nullptr;
            w.h_staging = nullptr;
            w.capacity = 0;
            const size_t want = padded + padded / 4;
            MSV_CUDA_TRY(cudaMalloc(&w.d_residues, want));
            MSV_CUDA_TRY(cudaHostAlloc(&w.h_staging, want, cudaHostAllocPortable));
            w.capacity = want;
        }
        std::memcpy(w.h_staging, residues, length);
        std::memset(w.h_staging + length, 0, padded - length);
        MSV_CUDA_TRY(cudaMemcpyAsync(w.d_residues, w.h_staging, padded, cudaMemcpyHostToDevice, w.stream));
        p.residues = dp.residues = w.d_residues;
        if (diagonal) MSV_CUDA_TRY(cudaLaunchKernelEx(&config, msv::msv_diag_kernel<false>, dp, msv::Wave_no_residues{}));
        else MSV_CUDA_TRY(cudaLaunchKernelEx(&config, w.kernels->with_device_residues, p, msv::Wave_no_residues{}));
    }
    ++g_launches;
    // the kernel writes the result into pinned host memory; poll it (a stream query every so often notices a failed launch)
    volatile msv::Wave_result* result = w.h_result;
    for (uint32_t spins = 1; result->tag != p.tag; ++spins) {
        if ((spins & 0x3ffu) == 0) {
            const cudaError_t state = cudaStreamQuery(w.stream);
            if (state != cudaErrorNotReady) {
                MSV_CUDA_TRY(state);
                if (result->tag != p.tag) return fail(MSV_ERR_CUDA, "the single-sequence kernel finished without a result");
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    const uint32_t status = result->status;
    if (status & 2u) {
        size_t at = 0;
        while (at < length && residues[at] < MSV_ALPHABET) ++at;
        return fail(MSV_ERR_BAD_RESIDUE, "residue code %u at position %zu is outside 0..19", at < length ? static_cast<unsigned>(residues[at]) : 255u, at);
    }
    if (status & 1u) *rescore = true;
    else *score = result->score;
    return MSV_OK;
}

} // namespace

// =====================================================================================================================
extern "C" {

int msv_cuda_abi_version(void) { return MSV_CUDA_ABI_VERSION; }

const char* msv_cuda_last_error(void) { return g_last_error.c_str(); }

uint64_t msv_cuda_launch_count(int reset) {
    const uint64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int msv_cuda_device_count(int* count) {
    if (!count) return fail(MSV_ERR_INVALID_ARGUMENT, "count is NULL");
    *count = 0;
    cudaError_t err = cudaGetDeviceCount(count);
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        *count = 0;
        return fail(MSV_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(err));
    }
    return MSV_OK;
}

// ---- host-side model arithmetic (libm, fp32) --------------------------------------------------------------------------
// Background residue frequencies used for the log-odds ratio; the values HMMER's p7_AminoFrequencies returns and the
// reference hard-codes (MSV_HMM.cpp:21-27).
static const float k_background[MSV_ALPHABET] = {0.0787945f, 0.0151600f, 0.0535222f, 0.0668298f, 0.0397062f,
                                                 0.0695071f, 0.0229198f, 0.0590092f, 0.0594422f, 0.0963728f,
                                                 0.0237718f, 0.0414386f, 0.0482904f, 0.0395639f, 0.0540978f,
                                                 0.0683364f, 0.0540687f, 0.0673417f, 0.0114135f, 0.0304133f};

int msv_host_emission_table(const float* match_emissions, size_t model_length, float* table) {
    if (!match_emissions || !table) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    for (size_t col = 0; col < model_length; ++col) {
        const float* node = match_emissions + col * MSV_ALPHABET;
        for (int res = 0; res < MSV_ALPHABET; ++res) table[res * model_length + col] = logf(node[res] / k_background[res]);
    }
    return MSV_OK;
}

int msv_host_model_transitions(size_t model_length, float* tr_B_Mk, float* tr_E_C, float* tr_E_J) {
    if (!tr_B_Mk || !tr_E_C || !tr_E_J) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    const volatile float expected_hits = 2.0f; // "nu"; volatile keeps logf a run-time libm call
    *tr_B_Mk = logf(2.0f / static_cast<float>(model_length * (model_length + 1)));
    *tr_E_C = logf((expected_hits - 1.0f) / expected_hits);
    *tr_E_J = logf(1.0f / expected_hits);
    return MSV_OK;
}

int msv_host_length_transitions(size_t residues, float* tr_loop, float* tr_move) {
    if (!tr_loop || !tr_move) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    const float denom = static_cast<float>(residues + 3);
    *tr_loop = logf(static_cast<float>(residues) / denom);
    *tr_move = logf(3.0f / denom);
    return MSV_OK;
}

int msv_host_encode(const char* letters, size_t n, uint8_t* codes, size_t* bad_at) {
    static const struct Lut {
        uint8_t v[256];
        Lut() {
            std::memset(v, 0xff, sizeof v);
            const char* order = "ACDEFGHIKLMNPQRSTVWY";
            for (int i = 0; i < MSV_ALPHABET; ++i) v[static_cast<unsigned char>(order[i])] = static_cast<uint8_t>(i);
        }
    } lut;
    if ((!letters || !codes) && n) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    uint8_t worst = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t c = lut.v[static_cast<unsigned char>(letters[i])];
        codes[i] = c;
        worst |= c;
    }
    if (worst & 0x80) {
        size_t at = 0;
        while (lut.v[static_cast<unsigned char>(letters[at])] != 0xff) ++at;
        if (bad_at) *bad_at = at;
        return fail(MSV_ERR_BAD_RESIDUE, "letter '%c' at position %zu is not one of ACDEFGHIKLMNPQRSTVWY", letters[at], at);
    }
    return MSV_OK;
}

int msv_host_partition_by_cells(const uint64_t* offsets, size_t n, int parts, size_t* bounds) {
    if (parts < 1 || !bounds || (n && !offsets)) return fail(MSV_ERR_INVALID_ARGUMENT, "bad partition request");
    bounds[0] = 0;
    const uint64_t base = n ? offsets[0] : 0, total = n ? offsets[n] - base : 0;
    for (int part = 1; part < parts; ++part) {
        // first sequence boundary at or after the ideal cut (binary search); the nearer of it and its predecessor wins
        const uint64_t want = base + total / parts * part + (total % parts) * part / parts;
        size_t q = static_cast<size_t>(std::lower_bound(offsets, offsets + n + 1, want) - offsets);
        if (q > 0 && q <= n && offsets[q] - want > want - offsets[q - 1]) --q;
        bounds[part] = std::min(q, n);
    }
    bounds[parts] = n;
    for (int part = 1; part <= parts; ++part) bounds[part] = std::max(bounds[part], bounds[part - 1]);
    return MSV_OK;
}

// ---- model ----------------------------------------------------------------------------------------------------------
int msv_cuda_model_create(const float* emission_scores, size_t model_length, float tr_B_Mk, float tr_E_C, float tr_E_J,
                          int device, msv_model** out) {
    if (!out) return fail(MSV_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!emission_scores || model_length < 1) return fail(MSV_ERR_INVALID_ARGUMENT, "empty model");
    int count = 0;
    if (int rc = msv_cuda_device_count(&count)) return rc;
    if (count == 0) return fail(MSV_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= count) return fail(MSV_ERR_INVALID_ARGUMENT, "device %d out of range (have %d)", device, count);

    const size_t columns = model_length - 1; // without the dummy M0
    const bool forced = std::getenv("MSV_CUDA_GEOMETRY") != nullptr;
    const Geometry* bulk_geo = choose_geometry(columns);
    const Geometry* quad_geo = choose_quad_geometry(columns);
    if (!bulk_geo && !quad_geo)
        return fail(MSV_ERR_MODEL_TOO_LONG, "model of %zu columns exceeds the on-chip table capacity (128 lanes x 44 columns)", columns);

    Device_guard guard(device);
    MSV_CUDA_TRY(guard.status);
    cudaDeviceProp prop{};
    MSV_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MSV_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                    prop.minor);

    auto* model = new (std::nothrow) msv_model();
    if (!model) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    model->device = device;
    model->model_length = model_length;
    model->tr_B_Mk = tr_B_Mk;
    model->tr_E_C = tr_E_C;
    model->tr_E_J = tr_E_J;
    model->sm_count = prop.multiProcessorCount;
    model->forced = forced;

    // kernel layout (see msv_kernels.cuh).  Lane g of a sequence's G lanes owns model columns g*K + j + 1, j = 0..K-1;
    // -inf beyond the model.
    //   shared-memory part : [residue][quad q][lane g][4]   j = KT + 4q + c      (generic family: KT = 0)
    //   tensor-memory part : [residue][lane g][KT]          j = 0 .. KT-1
    const auto build = [&](const Geometry* geo, msv_model::Plan& plan) -> cudaError_t {
        const int G = geo->G, K = geo->K, KT = std::max(geo->KT, 0), KS = K - KT;
        const int copies = (geo->KT < 0 && G < 8) ? 8 / G : 1; // lane groups narrower than a quarter-warp: interleaved copies
        const int R = G * copies;                               // lane slots per quad row of the shared-memory table
        const bool pair = geo->KT < 0 && KS % 4 == 2; // lane groups: the two highest columns, 16 lane slots x 2 floats behind the quads
        const size_t row_floats = static_cast<size_t>(KS / 4) * R * 4 + (pair ? 32 : 0);
        const size_t shared_floats = static_cast<size_t>(MSV_ALPHABET) * row_floats;
        const size_t floats = shared_floats + static_cast<size_t>(MSV_ALPHABET) * KT * G;
        std::vector<float> laid(std::max<size_t>(floats, 4), -std::numeric_limits<float>::infinity());
        const auto emission = [&](int res, int g, int j) {
            const size_t col = static_cast<size_t>(g) * K + j + 1;
            return col <= columns ? emission_scores[res * model_length + col] : -std::numeric_limits<float>::infinity();
        };
        const bool tensor_high = geo->variant == 1; // TMEM_AHEAD kernels keep the tensor-memory columns at the top of a lane
        const int shared_first = tensor_high ? 0 : KT, tensor_first = tensor_high ? KS : 0;
        for (int res = 0; res < MSV_ALPHABET; ++res)
            for (int g = 0; g < G; ++g) {
                for (int js = 0; js < KS / 4 * 4; ++js)
                    for (int copy = 0; copy < copies; ++copy)
                        laid[res * row_floats + (static_cast<size_t>(js / 4) * R + copy * G + g) * 4 + js % 4] = emission(res, g, shared_first + js);
                if (pair)
                    for (int copy = 0; copy < 16 / G; ++copy)
                        for (int jp = 0; jp < 2; ++jp)
                            laid[res * row_floats + static_cast<size_t>(KS / 4) * R * 4 + (copy * G + g) * 2 + jp] = emission(res, g, KS / 4 * 4 + jp);
                for (int jt = 0; jt < KT; ++jt)
                    laid[shared_floats + (static_cast<size_t>(res) * G + g) * KT + jt] = emission(res, g, tensor_first + jt);
            }
        plan.table_bytes = laid.size() * sizeof(float);
        plan.shared_bytes = geo->shared_bytes();
        if (static_cast<size_t>(prop.sharedMemPerBlockOptin) < plan.shared_bytes + 1024) return cudaErrorInvalidConfiguration;
        cudaError_t err = cudaMalloc(&plan.d_table, plan.table_bytes);
        if (err == cudaSuccess) err = cudaMemcpy(plan.d_table, laid.data(), plan.table_bytes, cudaMemcpyHostToDevice);
        for (Scan_kernel fn : {geo->fn, geo->fn_cj_same, geo->fn_cj_same_exact, geo->fn_cj_same_blocks, geo->fn_group_spec})
            if (err == cudaSuccess && fn)
                err = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(plan.shared_bytes));
        if (err == cudaSuccess) plan.geo = geo;
        return err;
    };
    cudaError_t err = cudaSuccess;
    if (bulk_geo) err = build(bulk_geo, model->bulk);
    if (err == cudaSuccess && quad_geo && !forced) {
        const cudaError_t quad_err = build(quad_geo, model->quad);
        if (!bulk_geo) err = quad_err; // the quad plan is optional unless it is the only one
    }
    if (err == cudaSuccess && bulk_geo && !forced) {
        if (const Geometry* octet_geo = choose_octet_geometry(columns)) (void)build(octet_geo, model->octet);    // optional
        if (const Geometry* narrow_geo = choose_narrow_geometry(columns)) (void)build(narrow_geo, model->narrow); // optional
        (void)cudaGetLastError();
    }
    if (err != cudaSuccess || (!model->bulk.geo && !model->quad.geo)) {
        cudaFree(model->bulk.d_table);
        cudaFree(model->octet.d_table);
        cudaFree(model->narrow.d_table);
        cudaFree(model->quad.d_table);
        delete model;
        (void)cudaGetLastError();
        if (err == cudaErrorInvalidConfiguration)
            return fail(MSV_ERR_MODEL_TOO_LONG, "emission table of a %zu-column model exceeds shared + tensor memory", columns);
        return fail(MSV_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(err));
    }
    if (model->bulk.geo && model->bulk.geo->fn_cj_same_blocks) { // optional: without it the whole-sequence mode stays
        void* pinned = nullptr;
        if (cudaMalloc(&model->d_speculation_totals, 2 * sizeof(unsigned int)) == cudaSuccess &&
            cudaMemset(model->d_speculation_totals, 0, 2 * sizeof(unsigned int)) == cudaSuccess &&
            cudaHostAlloc(&pinned, 2 * sizeof(unsigned int), cudaHostAllocPortable) == cudaSuccess) {
            model->h_speculation_totals = static_cast<volatile unsigned int*>(pinned);
            model->h_speculation_totals[0] = model->h_speculation_totals[1] = 0;
        } else {
            cudaFree(model->d_speculation_totals);
            model->d_speculation_totals = nullptr;
            (void)cudaGetLastError();
        }
    }
    // the single-sequence latency plan is optional: any failure here just leaves the four-warp kernel in charge
    if (std::memcmp(&tr_E_C, &tr_E_J, sizeof(float)) == 0 && !forced && !std::getenv("MSV_CUDA_NO_WAVE")) {
        if (wave_build(model, emission_scores, columns) != cudaSuccess) {
            wave_release(model);
            (void)cudaGetLastError();
        }
    }
    *out = model;
    return MSV_OK;
}

int msv_cuda_model_destroy(msv_model* model) {
    if (!model) return MSV_OK;
    db_release(model->workspace);
    {
        Device_guard guard(model->device);
        wave_release(model);
        cudaFree(model->d_speculation_totals);
        if (model->h_speculation_totals) cudaFreeHost(const_cast<unsigned int*>(model->h_speculation_totals));
        cudaFree(model->bulk.d_table);
        cudaFree(model->octet.d_table);
        cudaFree(model->narrow.d_table);
        cudaFree(model->quad.d_table);
    }
    delete model;
    return MSV_OK;
}

int msv_cuda_model_geometry(const msv_model* model, int* lanes_per_sequence, int* columns_per_lane,
                            int* tensor_columns_per_lane, int* threads_per_cta, size_t* shared_bytes) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    const msv_model::Plan& plan = model->bulk.geo ? model->bulk : model->quad; // the throughput plan
    if (lanes_per_sequence) *lanes_per_sequence = plan.geo->G;
    if (columns_per_lane) *columns_per_lane = plan.geo->K;
    if (tensor_columns_per_lane) *tensor_columns_per_lane = plan.geo->KT;
    if (threads_per_cta) *threads_per_cta = plan.geo->threads;
    if (shared_bytes) *shared_bytes = plan.shared_bytes;
    return MSV_OK;
}

int msv_cuda_model_wave_geometry(const msv_model* model, int* columns_per_lane, int* warps, int* ctas, int* diagonal_ctas) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    if (columns_per_lane) *columns_per_lane = model->wave.kernels ? model->wave.K : 0;
    if (warps) *warps = static_cast<int>(model->wave.warps);
    if (ctas) *ctas = static_cast<int>(model->wave.ctas);
    if (diagonal_ctas) *diagonal_ctas = model->wave.d_diag_table ? static_cast<int>(model->wave.diag_ctas) : 0;
    return MSV_OK;
}

int msv_cuda_model_speculation(const msv_model* model, unsigned int* failed, unsigned int* scanned, int* rows_next) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    const unsigned int f = model->h_speculation_totals ? model->h_speculation_totals[0] : 0u;
    const unsigned int o = model->h_speculation_totals ? model->h_speculation_totals[1] : 0u;
    if (failed) *failed = f;
    if (scanned) *scanned = o;
    if (rows_next) { // the decision launch_scan would take now (it also folds the totals in once 4096 new sequences are behind them)
        const float share = (o - model->seen_offered >= 4096u) ? static_cast<float>(f - model->seen_failures) / static_cast<float>(o - model->seen_offered)
                                                                : model->hit_share;
        const Geometry* geo = model->bulk.geo;
        *rows_next = (geo && geo->fn_cj_same_exact) ? static_cast<int>(rows_for(geo->K, share, false)) : MSV_ROWS_EXACT;
    }
    return MSV_OK;
}

int msv_cuda_model_plan(const msv_model* model, const msv_db* db, int* lanes_per_sequence, int* sequences_per_cta) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    const Length_profile whole = whole_profile(db);
    const Launch_plan chosen = plan_launch(model, db, 0, db->n, db->total, &whole);
    if (lanes_per_sequence) *lanes_per_sequence = chosen.plan->geo->G;
    if (sequences_per_cta) *sequences_per_cta = static_cast<int>(chosen.slots_per_cta);
    return MSV_OK;
}

int msv_cuda_model_plan_long_sequences(const msv_model* model, const msv_db* db, unsigned int* n_long, int* fast_ctas, int* fast_sequences_per_cta) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    const Length_profile whole = whole_profile(db);
    const Launch_plan chosen = plan_launch(model, db, 0, db->n, db->total, &whole);
    if (n_long) *n_long = chosen.n_long;
    if (fast_ctas) *fast_ctas = static_cast<int>(chosen.fast_ctas);
    if (fast_sequences_per_cta) *fast_sequences_per_cta = static_cast<int>(chosen.fast_slots);
    return MSV_OK;
}

// ---- database -------------------------------------------------------------------------------------------------------
int msv_cuda_db_create(int device, const uint8_t* residues, const uint64_t* offsets, size_t n, msv_db** out) {
    if (!out) return fail(MSV_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    int count = 0;
    if (int rc = msv_cuda_device_count(&count)) return rc;
    if (device < 0 || device >= count) return fail(MSV_ERR_INVALID_ARGUMENT, "device %d out of range (have %d)", device, count);
    Device_guard guard(device);
    MSV_CUDA_TRY(guard.status);
    auto* db = new (std::nothrow) msv_db();
    if (!db) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    db->device = device;
    const int rc = db_fill(db, residues, offsets, n, nullptr);
    if (rc != MSV_OK) {
        db_release(db);
        return rc;
    }
    *out = db;
    return MSV_OK;
}

int msv_cuda_db_destroy(msv_db* db) { return db_release(db); }

int msv_cuda_db_info(const msv_db* db, size_t* n, uint64_t* total_residues, uint64_t* longest) {
    if (!db) return fail(MSV_ERR_INVALID_ARGUMENT, "db is NULL");
    if (n) *n = db->n;
    if (total_residues) *total_residues = db->total;
    if (longest) *longest = db->longest;
    return MSV_OK;
}

// ---- scoring --------------------------------------------------------------------------------------------------------
int msv_cuda_db_score_device(msv_model* model, msv_db* db, float* scores_device, void* cuda_stream) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    if (model->device != db->device) return fail(MSV_ERR_INVALID_ARGUMENT, "model and database live on different devices");
    if (db->n && !scores_device) return fail(MSV_ERR_INVALID_ARGUMENT, "scores_device is NULL");
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    return launch_scan(model, db, 0, db->n, db->total, 0, scores_device, static_cast<cudaStream_t>(cuda_stream));
}

int msv_cuda_db_score_gather(msv_model* model, msv_db* db, float* const* gathered, int n_gathered, size_t first_index,
                             void* cuda_stream) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    if (model->device != db->device) return fail(MSV_ERR_INVALID_ARGUMENT, "model and database live on different devices");
    if (n_gathered < 1 || n_gathered > msv::kMaxScoreMirrors + 1 || !gathered)
        return fail(MSV_ERR_INVALID_ARGUMENT, "between 1 and %d gathered arrays are supported", msv::kMaxScoreMirrors + 1);
    for (int r = 0; r < n_gathered; ++r)
        if (db->n && !gathered[r]) return fail(MSV_ERR_INVALID_ARGUMENT, "gathered[%d] is NULL", r);
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    float* shifted[msv::kMaxScoreMirrors + 1];
    for (int r = 0; r < n_gathered; ++r) shifted[r] = gathered[r] + first_index;
    return launch_scan(model, db, 0, db->n, db->total, 0, shifted[0], static_cast<cudaStream_t>(cuda_stream), shifted + 1,
                       n_gathered - 1);
}

int msv_cuda_db_score(msv_model* model, msv_db* db, float* scores_host) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    if (model->device != db->device) return fail(MSV_ERR_INVALID_ARGUMENT, "model and database live on different devices");
    if (db->n && !scores_host) return fail(MSV_ERR_INVALID_ARGUMENT, "scores_host is NULL");
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    if (int rc = launch_scan(model, db, 0, db->n, db->total, 0, db->d_scores, nullptr)) return rc;
    if (db->n) MSV_CUDA_TRY(cudaMemcpy(scores_host, db->d_scores, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    return MSV_OK;
}

int msv_cuda_score_batch(msv_model* model, const uint8_t* residues, const uint64_t* offsets, size_t n, float* scores_host) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    if (n && (!offsets || !scores_host)) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    if (!model->workspace) {
        model->workspace = new (std::nothrow) msv_db();
        if (!model->workspace) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
        model->workspace->device = model->device;
    }
    return score_batch_pipelined(model, model->workspace, residues, offsets, n, scores_host);
}

int msv_cuda_score_batch_gather(msv_model* model, const uint8_t* residues, const uint64_t* offsets, size_t n, float* const* gathered,
                                int n_gathered, size_t first_index) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    if (n && !offsets) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_gathered < 1 || n_gathered > msv::kMaxScoreMirrors + 1 || !gathered)
        return fail(MSV_ERR_INVALID_ARGUMENT, "between 1 and %d gathered arrays are supported", msv::kMaxScoreMirrors + 1);
    for (int r = 0; r < n_gathered; ++r)
        if (n && !gathered[r]) return fail(MSV_ERR_INVALID_ARGUMENT, "gathered[%d] is NULL", r);
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    if (!model->workspace) {
        model->workspace = new (std::nothrow) msv_db();
        if (!model->workspace) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
        model->workspace->device = model->device;
    }
    float* shifted[msv::kMaxScoreMirrors + 1];
    for (int r = 0; r < n_gathered; ++r) shifted[r] = gathered[r] + first_index;
    return score_batch_pipelined(model, model->workspace, residues, offsets, n, nullptr, shifted[0], shifted + 1, n_gathered - 1);
}

int msv_cuda_score_fasta(msv_model* model, const char* text, size_t bytes, float* scores_host, size_t capacity, size_t* n_sequences,
                         size_t* rejected) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    if (n_sequences) *n_sequences = 0;
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    if (!model->workspace) {
        model->workspace = new (std::nothrow) msv_db();
        if (!model->workspace) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
        model->workspace->device = model->device;
    }
    msv_db* db = model->workspace;
    if (int rc = msv_detail::db_fill_from_fasta(db, text, bytes, rejected)) {
        db->n = 0;
        return rc;
    }
    if (n_sequences) *n_sequences = db->n;
    if (db->n > capacity) return fail(MSV_ERR_INVALID_ARGUMENT, "the text holds %zu sequences but scores_host has room for %zu", db->n, capacity);
    if (db->n == 0) return MSV_OK;
    if (!scores_host) return fail(MSV_ERR_INVALID_ARGUMENT, "scores_host is NULL");
    if (int rc = launch_scan(model, db, 0, db->n, db->total, 0, db->d_scores, nullptr)) return rc;
    MSV_CUDA_TRY(cudaMemcpy(scores_host, db->d_scores, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    return MSV_OK;
}

int msv_cuda_model_device(const msv_model* model, int* device) {
    if (!model || !device) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    *device = model->device;
    return MSV_OK;
}

int msv_cuda_db_filter_device(msv_db* db, const float* scores_device, float mu, float lambda, float* bits_device,
                              float* pvalues_device, void* cuda_stream) {
    if (!db) return fail(MSV_ERR_INVALID_ARGUMENT, "db is NULL");
    if (db->n == 0) return MSV_OK;
    if (!scores_device) return fail(MSV_ERR_INVALID_ARGUMENT, "scores_device is NULL");
    if (!(lambda > 0.0f)) return fail(MSV_ERR_INVALID_ARGUMENT, "lambda must be positive");
    Device_guard guard(db->device);
    MSV_CUDA_TRY(guard.status);
    const uint32_t n32 = static_cast<uint32_t>(db->n);
    msv::msv_filter_statistics_kernel<<<(n32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        scores_device, db->d_offsets, n32, static_cast<double>(mu), static_cast<double>(lambda), bits_device, pvalues_device);
    ++g_launches;
    MSV_CUDA_TRY(cudaGetLastError());
    return MSV_OK;
}

int msv_cuda_db_score_filter(msv_model* model, msv_db* db, float mu, float lambda, float* scores_host, float* bits_host,
                             float* pvalues_host) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    if (model->device != db->device) return fail(MSV_ERR_INVALID_ARGUMENT, "model and database live on different devices");
    if (db->n == 0) return MSV_OK;
    if (!scores_host) return fail(MSV_ERR_INVALID_ARGUMENT, "scores_host is NULL");
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    if (db->cap_stats < db->n) {
        cudaFree(db->d_stats);
        db->d_stats = nullptr;
        db->cap_stats = 0;
        MSV_CUDA_TRY(cudaMalloc(&db->d_stats, 2 * db->cap_n * sizeof(float)));
        db->cap_stats = db->cap_n;
    }
    float* d_bits = db->d_stats;
    float* d_p = db->d_stats + db->cap_stats;
    if (int rc = launch_scan(model, db, 0, db->n, db->total, 0, db->d_scores, nullptr)) return rc;
    if (int rc = msv_cuda_db_filter_device(db, db->d_scores, mu, lambda, d_bits, d_p, nullptr)) return rc;
    MSV_CUDA_TRY(cudaMemcpy(scores_host, db->d_scores, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    if (bits_host) MSV_CUDA_TRY(cudaMemcpy(bits_host, d_bits, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    if (pvalues_host) MSV_CUDA_TRY(cudaMemcpy(pvalues_host, d_p, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    return MSV_OK;
}

int msv_cuda_host_register(const void* buffer, size_t bytes) {
    if (!buffer || bytes == 0) return MSV_OK;
    int count = 0;
    if (int rc = msv_cuda_device_count(&count)) return rc;
    MSV_CUDA_TRY(cudaHostRegister(const_cast<void*>(buffer), bytes, cudaHostRegisterPortable));
    return MSV_OK;
}

int msv_cuda_host_unregister(const void* buffer) {
    if (!buffer) return MSV_OK;
    MSV_CUDA_TRY(cudaHostUnregister(const_cast<void*>(buffer)));
    return MSV_OK;
}

int msv_cuda_score_sequence(msv_model* model, const uint8_t* residues, size_t length, float* score) {
    if (!score) return fail(MSV_ERR_INVALID_ARGUMENT, "score is NULL");
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    if (length && !residues) return fail(MSV_ERR_INVALID_ARGUMENT, "residues is NULL");
    if (model->wave.kernels && length < (1ull << 27)) {
        Device_guard guard(model->device);
        MSV_CUDA_TRY(guard.status);
        bool rescore = false;
        if (int rc = wave_score(model, residues, length, score, &rescore)) return rc;
        if (!rescore) return MSV_OK;
        // a real hit: J overtook N somewhere, the speculative rows do not apply -- the exact kernel below scores it
    }
    const uint64_t offsets[2] = {0, length};
    return msv_cuda_score_batch(model, residues, offsets, 1, score);
}

} // extern "C"
