#include "MSV_HMM.hpp"

#include <algorithm>
#include <limits>
#include <stdexcept>

#include "msv_cuda.h"

namespace {

[[noreturn]] void throw_last_error(const char* what, int status) {
    const auto message = std::string(what) + ": " + msv_cuda_last_error();
    if (status == MSV_ERR_BAD_RESIDUE) throw std::out_of_range(message);
    throw std::runtime_error(message);
}

} // namespace

Device_database::Device_database(const Packed_sequences& database, int device)
    : sequences(database.size()), residues(database.total_residues()), device_index(device) {
    msv_db* raw = nullptr;
    const auto status = msv_cuda_db_create(device, database.residues.data(), database.offsets.data(), database.size(), &raw);
    if (status != MSV_OK) throw_last_error("Device_database: cannot upload the database", status);
    resident = std::shared_ptr<msv_db>(raw, [](msv_db* db) { msv_cuda_db_destroy(db); });
}

// Model preparation.  All fp32 expressions (log-odds table, B->M_k, E->C, E->J) are evaluated by the shared host
// helpers of the C ABI so that the C++ class and every other binding produce the same bits as the reference
// constructor (reference MSV_HMM.cpp:35-53).
MSV_HMM::MSV_HMM(const Profile_HMM& base_hmm)
    : model_length(base_hmm.model_length), msv_mu(base_hmm.stats_local_msv_mu), msv_lambda(base_hmm.stats_local_msv_lambda) {
    // a truncated .hmm leaves fewer nodes than LENG announces (the reader only prints a warning, like the reference's):
    // never read past what was parsed
    if (base_hmm.match_emissions.size() != model_length)
        throw std::invalid_argument("MSV_HMM: profile HMM is incomplete (" + std::to_string(base_hmm.match_emissions.size()) +
                                    " of " + std::to_string(model_length) + " nodes)");
    emission_scores.resize(NUM_OF_AMINO_ACIDS * model_length);
    if (model_length > 0) {
        msv_host_emission_table(base_hmm.match_emissions.front().data(), model_length, emission_scores.data());
    }
    msv_host_model_transitions(model_length, &tr_B_Mk, &tr_E_C, &tr_E_J);
}

void MSV_HMM::set_device(int device) {
    device_index = device;
    device_model.reset();
}

msv_model* MSV_HMM::on_device() {
    if (!device_model) {
        msv_model* raw = nullptr;
        const auto status =
            msv_cuda_model_create(emission_scores.data(), model_length, tr_B_Mk, tr_E_C, tr_E_J, device_index, &raw);
        if (status != MSV_OK) throw_last_error("MSV_HMM: cannot create the device model", status);
        device_model = std::shared_ptr<msv_model>(raw, [](msv_model* m) { msv_cuda_model_destroy(m); });
    }
    return device_model.get();
}

// Host recurrence (the API's sequential entry point; semantics of reference MSV_HMM.cpp:74-113).  One row of M
// values is updated in place from the last column to the first, so each cell still reads the previous row's left
// neighbour; E, J, C, N, B are scalars.  Same operands, same fp32 adds, exact max => same bits as the reference.
Log_score MSV_HMM::run_on_sequence(const Protein_sequence& seq) {
    constexpr auto minus_infinity = -std::numeric_limits<Log_score>::infinity();
    const auto residues = seq.empty() ? size_t(0) : seq.size() - 1; // seq[0] is the '#' sentinel
    auto codes = std::vector<uint8_t>(residues);
    if (msv_host_encode(seq.data() + (seq.empty() ? 0 : 1), residues, codes.data(), nullptr) != MSV_OK)
        throw_last_error("MSV_HMM::run_on_sequence", MSV_ERR_BAD_RESIDUE);

    Log_score tr_loop, tr_move;
    msv_host_length_transitions(residues, &tr_loop, &tr_move);

    auto row = std::vector<Log_score>(model_length ? model_length : 1, minus_infinity);
    auto J = minus_infinity, C = minus_infinity;
    auto N = Log_score(0), B = tr_move;
    for (const auto code : codes) {
        const auto* emission = emission_scores.data() + static_cast<size_t>(code) * model_length;
        const auto entry = B + tr_B_Mk;
        auto E = minus_infinity;
        for (auto k = model_length; k-- > 1;) {
            const auto best_in = row[k - 1] < entry ? entry : row[k - 1];
            const auto cell = emission[k] + best_in;
            row[k] = cell;
            if (E < cell) E = cell;
        }
        const auto J_stay = J + tr_loop, J_from_E = E + tr_E_J;
        J = J_stay < J_from_E ? J_from_E : J_stay;
        const auto C_stay = C + tr_loop, C_from_E = E + tr_E_C;
        C = C_stay < C_from_E ? C_from_E : C_stay;
        N = N + tr_loop;
        const auto B_from_N = N + tr_move, B_from_J = J + tr_move;
        B = B_from_N < B_from_J ? B_from_J : B_from_N;
    }
    return C + tr_move;
}

Log_score MSV_HMM::parallel_run_on_sequence(const Protein_sequence& seq, bool /*should_specialize*/) {
    const auto residues = seq.empty() ? size_t(0) : seq.size() - 1;
    auto codes = std::vector<uint8_t>(residues);
    if (msv_host_encode(seq.data() + (seq.empty() ? 0 : 1), residues, codes.data(), nullptr) != MSV_OK)
        throw_last_error("MSV_HMM::parallel_run_on_sequence", MSV_ERR_BAD_RESIDUE);
    auto score = Log_score(0);
    const auto status = msv_cuda_score_sequence(on_device(), codes.data(), residues, &score);
    if (status != MSV_OK) throw_last_error("MSV_HMM::parallel_run_on_sequence", status);
    return score;
}

std::vector<Log_score> MSV_HMM::parallel_run_on_sequences(const Packed_sequences& database) {
    auto scores = std::vector<Log_score>(database.size());
    const auto status = msv_cuda_score_batch(on_device(), database.residues.data(), database.offsets.data(), database.size(),
                                             scores.data());
    if (status != MSV_OK) throw_last_error("MSV_HMM::parallel_run_on_sequences", status);
    return scores;
}

std::vector<Log_score> MSV_HMM::parallel_run_on_sequences(const Device_database& database) {
    if (database.device() != device_index) set_device(database.device());
    auto scores = std::vector<Log_score>(database.size());
    const auto status = msv_cuda_db_score(on_device(), database.handle(), scores.data());
    if (status != MSV_OK) throw_last_error("MSV_HMM::parallel_run_on_sequences", status);
    return scores;
}

// Scan, statistics AND selection happen on the device (msv_cuda_db_msv_filter): only the hits come back.  The survivors'
// index list stays on the GPU next to the database, where Viterbi_HMM::viterbi_filter_survivors picks it up.
std::vector<MSV_hit> MSV_HMM::msv_filter(const Device_database& database, float threshold) {
    if (database.device() != device_index) set_device(database.device());
    auto capacity = std::max<size_t>(1024, database.size() / 16); // ~2 % pass at F1 = 0.02; grown when a database is richer
    for (;;) {
        auto index = std::vector<uint32_t>(capacity);
        auto scores = std::vector<float>(capacity), bits = std::vector<float>(capacity), p_values = std::vector<float>(capacity);
        auto found = size_t(0);
        const auto status = msv_cuda_db_msv_filter(on_device(), database.handle(), msv_mu, msv_lambda, threshold, index.data(), scores.data(),
                                                   bits.data(), p_values.data(), capacity, &found);
        if (status != MSV_OK) throw_last_error("MSV_HMM::msv_filter", status);
        if (found > capacity) {
            capacity = found;
            continue;
        }
        auto hits = std::vector<MSV_hit>(found);
        for (size_t i = 0; i < found; ++i) hits[i] = MSV_hit{index[i], scores[i], bits[i], p_values[i]};
        return hits;
    }
}

struct MSV_HMM::Replicas {
    std::vector<int> devices;
    std::vector<msv_model*> models;
    msv_multi* multi = nullptr;
    ~Replicas() {
        msv_cuda_multi_destroy(multi);
        for (auto* m : models) msv_cuda_model_destroy(m);
    }
};

std::vector<Log_score> MSV_HMM::parallel_run_on_sequences(const Packed_sequences& database, const std::vector<int>& devices,
                                                          Score_gather gather) {
    if (devices.empty()) throw std::invalid_argument("MSV_HMM::parallel_run_on_sequences: no devices given");
    if (!replicas || replicas->devices != devices) {
        auto fresh = std::make_shared<Replicas>();
        fresh->devices = devices;
        for (const auto device : devices) {
            msv_model* raw = nullptr;
            const auto status = msv_cuda_model_create(emission_scores.data(), model_length, tr_B_Mk, tr_E_C, tr_E_J, device, &raw);
            if (status != MSV_OK) throw_last_error("MSV_HMM: cannot create the device model", status);
            fresh->models.push_back(raw);
        }
        const auto status = msv_cuda_multi_create(fresh->models.data(), static_cast<int>(fresh->models.size()), &fresh->multi);
        if (status != MSV_OK) throw_last_error("MSV_HMM: cannot set up the multi-GPU scan", status);
        replicas = fresh;
    }
    auto scores = std::vector<Log_score>(database.size());
    const auto status = msv_cuda_multi_score_batch(replicas->multi, database.residues.data(), database.offsets.data(), database.size(),
                                                   scores.data(), static_cast<int>(gather));
    if (status != MSV_OK) throw_last_error("MSV_HMM::parallel_run_on_sequences", status);
    return scores;
}

std::vector<Log_score> MSV_HMM::parallel_run_on_sequences(const Protein_sequences& sequences) {
    return parallel_run_on_sequences(Packed_sequences::from_sequences(sequences));
}
