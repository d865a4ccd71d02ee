#include "Viterbi_HMM.hpp"

#include <algorithm>
#include <stdexcept>
#include <string>

#include "msv_cuda.h"

namespace {

[[noreturn]] void throw_viterbi_error(const char* what, int status) {
    const auto message = std::string(what) + ": " + msv_cuda_last_error();
    if (status == MSV_ERR_BAD_RESIDUE) throw std::out_of_range(message);
    throw std::runtime_error(message);
}

} // namespace

// Every logf of the model is evaluated here, on the host, by the helpers all bindings share.
Viterbi_HMM::Viterbi_HMM(const Profile_HMM& base_hmm)
    : model_length(base_hmm.model_length), mu(base_hmm.stats_local_viterbi_mu), lambda(base_hmm.stats_local_viterbi_lambda) {
    if (base_hmm.match_emissions.size() != model_length || base_hmm.transitions.size() != model_length)
        throw std::invalid_argument("Viterbi_HMM: profile HMM is incomplete");
    emission_scores.resize(NUM_OF_AMINO_ACIDS * model_length);
    log_transitions.resize(NUM_OF_TRANSITIONS * model_length);
    if (model_length > 0) {
        msv_host_emission_table(base_hmm.match_emissions.front().data(), model_length, emission_scores.data());
        msv_host_viterbi_transitions(base_hmm.transitions.front().data(), model_length, log_transitions.data());
    }
    msv_host_model_transitions(model_length, &tr_B_Mk, &tr_E_C, &tr_E_J);
}

void Viterbi_HMM::set_device(int device) {
    device_index = device;
    device_model.reset();
}

msv_viterbi_model* Viterbi_HMM::on_device() {
    if (!device_model) {
        msv_viterbi_model* raw = nullptr;
        const auto status = msv_cuda_viterbi_model_create(emission_scores.data(), log_transitions.data(), model_length, tr_B_Mk,
                                                          tr_E_C, tr_E_J, device_index, &raw);
        if (status != MSV_OK) throw_viterbi_error("Viterbi_HMM: cannot create the device model", status);
        device_model = std::shared_ptr<msv_viterbi_model>(raw, [](msv_viterbi_model* m) { msv_cuda_viterbi_model_destroy(m); });
    }
    return device_model.get();
}

Log_score Viterbi_HMM::parallel_run_on_sequence(const Protein_sequence& seq) {
    const auto residues = seq.empty() ? size_t(0) : seq.size() - 1; // seq[0] is the '#' sentinel
    auto codes = std::vector<uint8_t>(residues);
    if (msv_host_encode(seq.data() + (seq.empty() ? 0 : 1), residues, codes.data(), nullptr) != MSV_OK)
        throw_viterbi_error("Viterbi_HMM::parallel_run_on_sequence", MSV_ERR_BAD_RESIDUE);
    const uint64_t offsets[2] = {0, residues};
    auto score = Log_score(0);
    const auto status = msv_cuda_viterbi_batch(on_device(), codes.data(), offsets, 1, &score);
    if (status != MSV_OK) throw_viterbi_error("Viterbi_HMM::parallel_run_on_sequence", status);
    return score;
}

std::vector<Log_score> Viterbi_HMM::parallel_run_on_sequences(const Packed_sequences& database) {
    auto scores = std::vector<Log_score>(database.size());
    const auto status =
        msv_cuda_viterbi_batch(on_device(), database.residues.data(), database.offsets.data(), database.size(), scores.data());
    if (status != MSV_OK) throw_viterbi_error("Viterbi_HMM::parallel_run_on_sequences", status);
    return scores;
}

std::vector<Log_score> Viterbi_HMM::parallel_run_on_sequences(const Protein_sequences& sequences) {
    return parallel_run_on_sequences(Packed_sequences::from_sequences(sequences));
}

std::vector<Log_score> Viterbi_HMM::parallel_run_on_sequences(const Device_database& database) {
    if (database.device() != device_index) set_device(database.device());
    auto scores = std::vector<Log_score>(database.size());
    const auto status = msv_cuda_db_viterbi(on_device(), database.handle(), scores.data());
    if (status != MSV_OK) throw_viterbi_error("Viterbi_HMM::parallel_run_on_sequences", status);
    return scores;
}

std::vector<MSV_hit> Viterbi_HMM::viterbi_filter(const Device_database& database, float threshold) {
    if (database.device() != device_index) set_device(database.device());
    const auto n = database.size();
    auto scores = std::vector<float>(n), bits = std::vector<float>(n), p_values = std::vector<float>(n);
    const auto status =
        msv_cuda_db_viterbi_filter(on_device(), database.handle(), mu, lambda, scores.data(), bits.data(), p_values.data());
    if (status != MSV_OK) throw_viterbi_error("Viterbi_HMM::viterbi_filter", status);
    auto hits = std::vector<MSV_hit>();
    for (size_t q = 0; q < n; ++q)
        if (p_values[q] <= threshold) hits.push_back(MSV_hit{q, scores[q], bits[q], p_values[q]});
    return hits;
}

std::vector<MSV_hit> Viterbi_HMM::viterbi_filter_survivors(const Device_database& database, float threshold) {
    if (database.device() != device_index) set_device(database.device());
    auto capacity = std::max<size_t>(1024, database.size() / 64);
    for (;;) {
        auto index = std::vector<uint32_t>(capacity);
        auto scores = std::vector<float>(capacity), bits = std::vector<float>(capacity), p_values = std::vector<float>(capacity);
        auto found = size_t(0);
        const auto status = msv_cuda_db_viterbi_filter_survivors(on_device(), database.handle(), mu, lambda, threshold, index.data(), scores.data(),
                                                                 bits.data(), p_values.data(), capacity, &found);
        if (status != MSV_OK) throw_viterbi_error("Viterbi_HMM::viterbi_filter_survivors", status);
        if (found > capacity) { // (the survivors list is untouched by this stage, so the call can simply be repeated)
            capacity = found;
            continue;
        }
        auto hits = std::vector<MSV_hit>(found);
        for (size_t i = 0; i < found; ++i) hits[i] = MSV_hit{index[i], scores[i], bits[i], p_values[i]};
        return hits;
    }
}
