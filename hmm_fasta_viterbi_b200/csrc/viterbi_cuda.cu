// viterbi_cuda.cu -- the msv_cuda_viterbi_* part of the C ABI (include/msv_cuda.h): model upload for the Plan-7 local
// Viterbi scan and its dispatch over a device-resident database.  Kernels: viterbi_kernels.cuh.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "msv_internal.hpp"
#include "viterbi_kernels.cuh"

namespace {

using Scan_kernel = void (*)(const msv::Scan_params);
struct Viterbi_geometry {
    int K, threads;
    Scan_kernel fn, fn_cj_same, fn_cj_same_spec; // general, tr_E_C == tr_E_J, the latter with speculative rows
    int G = 32;                                  // lanes per sequence: 32 (one warp) or 8 (four sequences per warp, short models)
    size_t shared_bytes() const { return static_cast<size_t>(MSV_ALPHABET + 2) * K * G * sizeof(float); }
    size_t table_floats() const { return static_cast<size_t>(MSV_ALPHABET + 2) * K * G + 32 * 5 * static_cast<size_t>(K) + 32 * 8; }
};

// three state registers per column: the register file, not shared memory, bounds the warps per SM.  Registers are handed
// out to a CTA in units of four warps, so only multiples of 128 threads are worth considering: 768 / 512 / 384 / 256 threads
// leave 80 / 128 / 168 / 255 registers each; the kernel wants 76, 108, 118, 126, 141, 154, 164, 176, 187, 200 ... at K = 4, 8, ... 40
// (and is content with 158 / 168 at K = 32 / 36 when that is the budget).
constexpr int viterbi_threads_for(int K) { return K <= 4 ? 768 : K <= 16 ? 512 : K <= 36 ? 384 : 256; }
template <int K> constexpr Viterbi_geometry viterbi_entry() {
    return Viterbi_geometry{K, viterbi_threads_for(K), msv::viterbi_scan_warp_kernel<K, viterbi_threads_for(K), false>,
                            msv::viterbi_scan_warp_kernel<K, viterbi_threads_for(K), true>,
                            msv::viterbi_scan_warp_kernel<K, viterbi_threads_for(K), true, true>};
}
const Viterbi_geometry k_viterbi_geometries[] = {
    viterbi_entry<4>(),  viterbi_entry<8>(),  viterbi_entry<12>(), viterbi_entry<16>(), viterbi_entry<20>(),
    viterbi_entry<24>(), viterbi_entry<28>(), viterbi_entry<32>(), viterbi_entry<36>(), viterbi_entry<40>(),
    viterbi_entry<44>(), viterbi_entry<48>(), viterbi_entry<52>(), viterbi_entry<56>(), viterbi_entry<60>(),
    viterbi_entry<64>(), viterbi_entry<68>(), viterbi_entry<72>(), viterbi_entry<76>(), viterbi_entry<80>(),
};

// eight lanes per sequence (viterbi_scan_group_kernel), exact rows only
constexpr int viterbi_group_threads_for(int K) { return K <= 12 ? 512 : K <= 28 ? 384 : 256; } // 118 ... 167, 207 ... 254 registers
template <int K> constexpr Viterbi_geometry viterbi_group_entry() {
    return Viterbi_geometry{K, viterbi_group_threads_for(K), msv::viterbi_scan_group_kernel<8, K, viterbi_group_threads_for(K), false>,
                            msv::viterbi_scan_group_kernel<8, K, viterbi_group_threads_for(K), true>,
                            msv::viterbi_scan_group_kernel<8, K, viterbi_group_threads_for(K), true>, 8};
}
const Viterbi_geometry k_viterbi_group_geometries[] = {
    viterbi_group_entry<8>(),  viterbi_group_entry<12>(), viterbi_group_entry<16>(), viterbi_group_entry<20>(), viterbi_group_entry<24>(),
    viterbi_group_entry<28>(), viterbi_group_entry<32>(), viterbi_group_entry<36>(), viterbi_group_entry<40>(), viterbi_group_entry<44>(),
    viterbi_group_entry<48>(), viterbi_group_entry<52>(), viterbi_group_entry<56>(),
};
const Viterbi_geometry* choose_viterbi_group_geometry(size_t model_length) {
    for (const Viterbi_geometry& g : k_viterbi_group_geometries)
        if (static_cast<size_t>(g.K) * 8 >= model_length) return &g;
    return nullptr;
}

const Viterbi_geometry* choose_viterbi_geometry(size_t model_length) {
    // 32 * K slots hold columns 0 .. LENG (column 0 is the dummy the recurrence needs at its left end)
    for (const Viterbi_geometry& g : k_viterbi_geometries)
        if (static_cast<size_t>(g.K) * 32 >= model_length) return &g;
    return nullptr;
}

} // namespace

struct msv_viterbi_model {
    int device = 0;
    size_t model_length = 0;
    const Viterbi_geometry* geo = nullptr; // one warp per sequence
    float4* d_table = nullptr;
    const Viterbi_geometry* group_geo = nullptr; // eight lanes per sequence: short models, when there are sequences enough
    float4* d_group_table = nullptr;
    float tr_B_Mk = 0, tr_E_C = 0, tr_E_J = 0;
    int sm_count = 0;
    msv_db* workspace = nullptr; // reused by msv_cuda_viterbi_batch (grow-only buffers)
};

extern "C" {

int msv_host_viterbi_transitions(const float* transitions, size_t model_length, float* log_transitions) {
    if (!transitions || !log_transitions) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    for (size_t i = 0; i < model_length * MSV_TRANSITIONS; ++i) log_transitions[i] = logf(transitions[i]);
    return MSV_OK;
}

int msv_cuda_viterbi_model_create(const float* emission_scores, const float* log_transitions, size_t model_length, float tr_B_Mk,
                                  float tr_E_C, float tr_E_J, int device, msv_viterbi_model** out) {
    if (!out) return fail(MSV_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!emission_scores || !log_transitions || model_length < 1) return fail(MSV_ERR_INVALID_ARGUMENT, "empty model");
    for (size_t i = 0; i < model_length * MSV_TRANSITIONS; ++i)
        if (!(log_transitions[i] <= 0.0f)) // also rejects NaN
            return fail(MSV_ERR_INVALID_ARGUMENT, "log_transitions[%zu] = %g is not the logarithm of a probability", i,
                        static_cast<double>(log_transitions[i]));
    int count = 0;
    if (int rc = msv_cuda_device_count(&count)) return rc;
    if (count == 0) return fail(MSV_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= count) return fail(MSV_ERR_INVALID_ARGUMENT, "device %d out of range (have %d)", device, count);
    const Viterbi_geometry* geo = choose_viterbi_geometry(model_length);
    if (!geo)
        return fail(MSV_ERR_MODEL_TOO_LONG, "Viterbi: model of %zu columns exceeds 32 lanes x %d columns", model_length - 1,
                    msv::kViterbiMaxColumnsPerLane);

    Device_guard guard(device);
    MSV_CUDA_TRY(guard.status);
    cudaDeviceProp prop{};
    MSV_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MSV_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    if (static_cast<size_t>(prop.sharedMemPerBlockOptin) < geo->shared_bytes() + 1024)
        return fail(MSV_ERR_MODEL_TOO_LONG, "Viterbi: tables of a %zu-column model exceed shared memory", model_length - 1);

    // kernel layout, see viterbi_kernels.cuh.  A sequence occupies G lanes (32, or 8 in the lane-group kernel); slot
    // s = lane * K + j holds model column s - pad (right aligned).
    const long columns = static_cast<long>(model_length) - 1; // LENG
    const float ninf = -std::numeric_limits<float>::infinity();
    enum { MM = 0, MI = 1, MD = 2, IM = 3, II = 4, DM = 5, DD = 6 }; // order of Profile_HMM::transitions (Profile_HMM.hpp:29)
    const auto tr = [&](long node, int which) { return log_transitions[node * MSV_TRANSITIONS + which]; };
    const auto upload = [&](const Viterbi_geometry* g, float4** d_table) -> cudaError_t {
        const int K = g->K, G = g->G;
        const long pad = static_cast<long>(G) * K - 1 - columns;
        // every transition is stored with the column it leaves; columns 1 .. LENG-1 have successors, nothing else does
        const auto own = [&](long slot, int which) {
            const long c = slot - pad;
            return (c >= 1 && c <= columns - 1) ? tr(c, which) : ninf;
        };
        std::vector<float> laid(g->table_floats(), ninf);
        const size_t row = static_cast<size_t>(K) * G;
        float* md = laid.data() + MSV_ALPHABET * row;
        float* dd = md + row;
        float* tensor = dd + row;
        float* edge = tensor + 32 * 5 * static_cast<size_t>(K);
        for (int lane = 0; lane < G; ++lane) {
            for (int j = 0; j < K; ++j) {
                const long slot = static_cast<long>(lane) * K + j, c = slot - pad;
                const int q = j / 4, w = j % 4;
                const size_t at = (static_cast<size_t>(q) * G + lane) * 4 + w;
                for (int res = 0; res < MSV_ALPHABET; ++res)
                    laid[res * row + at] = (c >= 1 && c <= columns) ? emission_scores[res * model_length + c] : ninf;
                md[at] = own(slot, MD);
                dd[at] = own(slot, DD);
                for (int copy = lane; copy < 32; copy += G) { // the tensor-memory part covers all 32 lanes of a warp
                    float* t = tensor + (static_cast<size_t>(copy) * (K / 4) + q) * 20;
                    t[w] = own(slot, MM);
                    t[4 + w] = own(slot, IM);
                    t[8 + w] = own(slot, DM);
                    t[12 + w] = own(slot, MI);
                    t[16 + w] = own(slot, II);
                }
            }
            const long last = static_cast<long>(lane) * K + K - 1; // the last lane: column LENG, which has no successor
            const int which[5] = {MM, IM, DM, MD, DD};
            for (int i = 0; i < 5; ++i) edge[lane * 8 + i] = own(last, which[i]);
        }
        cudaError_t err = cudaMalloc(d_table, laid.size() * sizeof(float));
        if (err == cudaSuccess) err = cudaMemcpy(*d_table, laid.data(), laid.size() * sizeof(float), cudaMemcpyHostToDevice);
        for (Scan_kernel fn : {g->fn, g->fn_cj_same, g->fn_cj_same_spec})
            if (err == cudaSuccess)
                err = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(g->shared_bytes()));
        return err;
    };

    auto* model = new (std::nothrow) msv_viterbi_model();
    if (!model) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    model->device = device;
    model->model_length = model_length;
    model->geo = geo;
    model->tr_B_Mk = tr_B_Mk;
    model->tr_E_C = tr_E_C;
    model->tr_E_J = tr_E_J;
    model->sm_count = prop.multiProcessorCount;
    const cudaError_t err = upload(geo, &model->d_table);
    if (err != cudaSuccess) {
        cudaFree(model->d_table);
        delete model;
        (void)cudaGetLastError();
        return fail(err == cudaErrorMemoryAllocation ? MSV_ERR_OUT_OF_MEMORY : MSV_ERR_CUDA, "Viterbi model upload failed: %s",
                    cudaGetErrorString(err));
    }
    // the lane-group plan is optional: short models only, and a failure to set it up just leaves the warp plan
    if (const Viterbi_geometry* group = choose_viterbi_group_geometry(model_length)) {
        if (static_cast<size_t>(prop.sharedMemPerBlockOptin) >= group->shared_bytes() + 1024 &&
            upload(group, &model->d_group_table) == cudaSuccess) {
            model->group_geo = group;
        } else {
            cudaFree(model->d_group_table);
            model->d_group_table = nullptr;
            (void)cudaGetLastError();
        }
    }
    *out = model;
    return MSV_OK;
}

int msv_cuda_viterbi_model_destroy(msv_viterbi_model* model) {
    if (!model) return MSV_OK;
    msv_detail::db_free(model->workspace);
    {
        Device_guard guard(model->device);
        cudaFree(model->d_table);
        cudaFree(model->d_group_table);
    }
    delete model;
    return MSV_OK;
}

int msv_cuda_viterbi_model_geometry(const msv_viterbi_model* model, int* columns_per_lane, int* threads_per_cta, size_t* shared_bytes) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    if (columns_per_lane) *columns_per_lane = model->geo->K; // of the one-warp-per-sequence plan
    if (threads_per_cta) *threads_per_cta = model->geo->threads;
    if (shared_bytes) *shared_bytes = model->geo->shared_bytes();
    return MSV_OK;
}

} // extern "C"

// One launch over the sequences listed in `order` (n of them; `residues` = their total length, for the plan); scores land at
// scores_device[original sequence index].
static int viterbi_launch(msv_viterbi_model* model, msv_db* db, const uint32_t* order, size_t n, uint64_t residues, float* scores_device,
                          cudaStream_t stream, bool likely_hits = false) {
    {
    // Eight lanes per sequence when the model is short and every slot gets enough work to balance (it has four times more
    // slots than the warp plan; same criterion as the MSV planner: rows per slot vs the longest sequence).
    // MSV_CUDA_VITERBI_GROUPS=0 / 1 forces the choice (tuning and test aid).
    const char* forced = std::getenv("MSV_CUDA_VITERBI_GROUPS");
    const bool balanced = model->group_geo &&
                          4 * (residues / (static_cast<uint64_t>(model->sm_count) * (model->group_geo->threads / 8))) >=
                              3 * std::max<uint64_t>(db->longest, 1);
    const bool grouped = model->group_geo && (forced ? forced[0] == '1' : balanced);
    const Viterbi_geometry* geo = grouped ? model->group_geo : model->geo;
    msv::Scan_params p{};
    p.table = grouped ? model->d_group_table : model->d_table;
    p.residues = db->d_residues;
    p.offsets = db->d_offsets;
    p.order = order;
    p.length_tr = db->d_length_tr;
    p.scores = scores_device;
    p.queue_head = db->d_queue;
    p.first_bad = db->d_first_bad;
    p.n = static_cast<uint32_t>(n);
    p.table_bytes = static_cast<uint32_t>(geo->shared_bytes());
    p.tr_B_Mk = model->tr_B_Mk;
    p.tr_E_C = model->tr_E_C;
    p.tr_E_J = model->tr_E_J;
    p.n_mirrors = 0;
    MSV_CUDA_TRY(cudaMemsetAsync(db->d_queue, 0, sizeof(unsigned int), stream));
    // persistent CTAs, one per SM, one warp per sequence (or per four) in flight; fewer warps when there are fewer sequences
    const size_t warps_per_cta = static_cast<size_t>(geo->threads) / 32;
    const size_t per_warp = 32 / static_cast<size_t>(geo->G);
    const size_t warps_wanted = (n + per_warp - 1) / per_warp;
    const size_t ctas = std::max<size_t>(1, std::min<size_t>(model->sm_count, (warps_wanted + warps_per_cta - 1) / warps_per_cta));
    const size_t warps = ctas == 1 ? std::min(warps_per_cta, warps_wanted) : std::min(warps_per_cta, (warps_wanted + ctas - 1) / ctas);
    const bool cj_same = std::memcmp(&model->tr_E_C, &model->tr_E_J, sizeof(float)) == 0;
    // speculative rows unless the database is one of long sequences, which would mostly be scanned twice
    // ... or one of filter survivors, which are hits more often than not (their speculation would fail)
    const bool speculate = !likely_hits && residues / n <= msv::kViterbiSpeculationMaxLength / 2 && !std::getenv("MSV_CUDA_NO_SPECULATION");
    const Scan_kernel kernel = !cj_same ? geo->fn : speculate ? geo->fn_cj_same_spec : geo->fn_cj_same;
    kernel<<<static_cast<int>(ctas), static_cast<int>(warps * 32), geo->shared_bytes(), stream>>>(p);
    msv_detail::count_launch();
    MSV_CUDA_TRY(cudaGetLastError());
    return MSV_OK;
    }
}

extern "C" {

int msv_cuda_db_viterbi_device(msv_viterbi_model* model, msv_db* db, float* scores_device, void* cuda_stream) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    if (model->device != db->device) return fail(MSV_ERR_INVALID_ARGUMENT, "model and database live on different devices");
    if (db->n == 0) return MSV_OK;
    if (!scores_device) return fail(MSV_ERR_INVALID_ARGUMENT, "scores_device is NULL");
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    return viterbi_launch(model, db, db->d_order, db->n, db->total, scores_device, static_cast<cudaStream_t>(cuda_stream));
}

int msv_cuda_db_viterbi_subset_device(msv_viterbi_model* model, msv_db* db, const uint32_t* indices_device, size_t count, float* scores_device,
                                      void* cuda_stream) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    if (model->device != db->device) return fail(MSV_ERR_INVALID_ARGUMENT, "model and database live on different devices");
    if (count == 0) return MSV_OK;
    if (count > db->n || !indices_device || !scores_device) return fail(MSV_ERR_INVALID_ARGUMENT, "bad subset");
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    // (the subset's own residue count is not known on the host: the database's mean length stands in for the plan)
    const uint64_t residues = std::max<uint64_t>(1, db->total / db->n) * count;
    return viterbi_launch(model, db, indices_device, count, residues, scores_device, static_cast<cudaStream_t>(cuda_stream), true);
}

int msv_cuda_db_viterbi(msv_viterbi_model* model, msv_db* db, float* scores_host) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
    if (db->n && !scores_host) return fail(MSV_ERR_INVALID_ARGUMENT, "scores_host is NULL");
    if (int rc = msv_cuda_db_viterbi_device(model, db, db->d_scores, nullptr)) return rc;
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    if (db->n) MSV_CUDA_TRY(cudaMemcpy(scores_host, db->d_scores, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    return MSV_OK;
}

int msv_cuda_db_viterbi_filter(msv_viterbi_model* model, msv_db* db, float mu, float lambda, float* scores_host, float* bits_host,
                               float* pvalues_host) {
    if (!model || !db) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL handle");
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
    if (int rc = msv_cuda_db_viterbi_device(model, db, db->d_scores, nullptr)) return rc;
    if (int rc = msv_cuda_db_filter_device(db, db->d_scores, mu, lambda, d_bits, d_p, nullptr)) return rc;
    MSV_CUDA_TRY(cudaMemcpy(scores_host, db->d_scores, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    if (bits_host) MSV_CUDA_TRY(cudaMemcpy(bits_host, d_bits, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    if (pvalues_host) MSV_CUDA_TRY(cudaMemcpy(pvalues_host, d_p, db->n * sizeof(float), cudaMemcpyDeviceToHost));
    return MSV_OK;
}

int msv_cuda_viterbi_batch(msv_viterbi_model* model, const uint8_t* residues, const uint64_t* offsets, size_t n, float* scores_host) {
    if (!model) return fail(MSV_ERR_INVALID_ARGUMENT, "model is NULL");
    if (n && (!offsets || !scores_host)) return fail(MSV_ERR_INVALID_ARGUMENT, "NULL argument");
    Device_guard guard(model->device);
    MSV_CUDA_TRY(guard.status);
    if (!model->workspace) {
        model->workspace = new (std::nothrow) msv_db();
        if (!model->workspace) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
        model->workspace->device = model->device;
    }
    if (int rc = msv_detail::db_refill(model->workspace, residues, offsets, n)) return rc;
    return msv_cuda_db_viterbi(model, model->workspace, scores_host);
}

} // extern "C"
