// bindings.cpp -- C handle API over the C++ host classes (Profile_HMM, FASTA_protein_sequences, Packed_sequences,
// MSV_HMM) so that Python (hmm_fasta_viterbi_b200/host.py, via ctypes) drives exactly the code a C++ caller uses.
// Exceptions never cross the boundary: every function returns 0 on success or -1 and keeps the message for
// msvh_last_error().
#include <cstring>
#include <exception>
#include <string>

#include "MSV_HMM.hpp"
#include "Viterbi_HMM.hpp"
#include "Synthetic_database.hpp"
#include "msv_cuda.h"

namespace {
thread_local std::string last_error;

template <class F> int guarded(F&& body) {
    try {
        body();
        return 0;
    } catch (const std::out_of_range& e) {
        last_error = std::string("out_of_range: ") + e.what();
        return -2;
    } catch (const std::exception& e) {
        last_error = e.what();
        return -1;
    }
}
} // namespace

extern "C" {

const char* msvh_last_error(void) { return last_error.c_str(); }

// ---- Profile_HMM ----
void* msvh_profile_load(const char* path) {
    Profile_HMM* p = nullptr;
    guarded([&] { p = new Profile_HMM(path); });
    return p;
}
void msvh_profile_free(void* p) { delete static_cast<Profile_HMM*>(p); }
size_t msvh_profile_model_length(void* p) { return static_cast<Profile_HMM*>(p)->model_length; }
const char* msvh_profile_name(void* p) { return static_cast<Profile_HMM*>(p)->name.c_str(); }
size_t msvh_profile_rows(void* p, int which) {
    auto* h = static_cast<Profile_HMM*>(p);
    return which == 0 ? h->match_emissions.size() : which == 1 ? h->insert_emissions.size() : h->transitions.size();
}
void msvh_profile_copy(void* p, int which, float* out) {
    auto* h = static_cast<Profile_HMM*>(p);
    if (which == 0 && !h->match_emissions.empty())
        std::memcpy(out, h->match_emissions.data(), h->match_emissions.size() * sizeof(h->match_emissions[0]));
    if (which == 1 && !h->insert_emissions.empty())
        std::memcpy(out, h->insert_emissions.data(), h->insert_emissions.size() * sizeof(h->insert_emissions[0]));
    if (which == 2 && !h->transitions.empty())
        std::memcpy(out, h->transitions.data(), h->transitions.size() * sizeof(h->transitions[0]));
}
void msvh_profile_stats(void* p, float* out6) {
    auto* h = static_cast<Profile_HMM*>(p);
    const float v[6] = {h->stats_local_msv_mu,         h->stats_local_msv_lambda,    h->stats_local_viterbi_mu,
                        h->stats_local_viterbi_lambda, h->stats_local_forward_theta, h->stats_local_forward_lambda};
    std::memcpy(out6, v, sizeof v);
}

// ---- FASTA_protein_sequences ----
void* msvh_fasta_load(const char* path) {
    FASTA_protein_sequences* f = nullptr;
    guarded([&] { f = new FASTA_protein_sequences(path); });
    return f;
}
void msvh_fasta_free(void* f) { delete static_cast<FASTA_protein_sequences*>(f); }
size_t msvh_fasta_count(void* f) { return static_cast<FASTA_protein_sequences*>(f)->sequences.size(); }
const char* msvh_fasta_record(void* f, size_t i) { return static_cast<FASTA_protein_sequences*>(f)->sequences[i].c_str(); }

// ---- Packed_sequences ----
void* msvh_packed_from_fasta_file(const char* path, size_t* rejected) {
    Packed_sequences* p = nullptr;
    guarded([&] { p = new Packed_sequences(Packed_sequences::from_fasta_file(path, rejected)); });
    return p;
}
void* msvh_packed_from_fasta(void* fasta) {
    Packed_sequences* p = nullptr;
    guarded([&] { p = new Packed_sequences(Packed_sequences::from_sequences(static_cast<FASTA_protein_sequences*>(fasta)->sequences)); });
    return p;
}
void* msvh_packed_synthetic_swissprot_like(size_t count, uint64_t seed) {
    Packed_sequences* p = nullptr;
    guarded([&] { p = new Packed_sequences(synthetic_swissprot_like(count, seed)); });
    return p;
}
void* msvh_packed_synthetic_long_uniform(size_t count, uint64_t seed, size_t shortest, size_t longest) {
    Packed_sequences* p = nullptr;
    guarded([&] { p = new Packed_sequences(synthetic_long_uniform(count, seed, shortest, longest)); });
    return p;
}
void* msvh_packed_from_arrays(const uint8_t* residues, const uint64_t* offsets, size_t n) {
    auto* p = new Packed_sequences();
    p->offsets.assign(offsets, offsets + n + 1);
    p->residues.assign(residues, residues + offsets[n]);
    return p;
}
void msvh_packed_free(void* p) { delete static_cast<Packed_sequences*>(p); }
size_t msvh_packed_count(void* p) { return static_cast<Packed_sequences*>(p)->size(); }
uint64_t msvh_packed_total(void* p) { return static_cast<Packed_sequences*>(p)->total_residues(); }
const uint8_t* msvh_packed_residues(void* p) { return static_cast<Packed_sequences*>(p)->residues.data(); }
const uint64_t* msvh_packed_offsets(void* p) { return static_cast<Packed_sequences*>(p)->offsets.data(); }

// ---- MSV_HMM ----
void* msvh_msv_create(void* profile, int device) {
    MSV_HMM* m = nullptr;
    guarded([&] {
        m = new MSV_HMM(*static_cast<Profile_HMM*>(profile));
        m->set_device(device);
    });
    return m;
}
void* msvh_msv_clone(void* m) { return new MSV_HMM(*static_cast<MSV_HMM*>(m)); }
void msvh_msv_free(void* m) { delete static_cast<MSV_HMM*>(m); }
size_t msvh_msv_length(void* m) { return static_cast<MSV_HMM*>(m)->length(); }
int msvh_msv_run_on_sequence(void* m, const char* seq, float* score) {
    return guarded([&] { *score = static_cast<MSV_HMM*>(m)->run_on_sequence(seq); });
}
int msvh_msv_parallel_run_on_sequence(void* m, const char* seq, int should_specialize, float* score) {
    return guarded([&] { *score = static_cast<MSV_HMM*>(m)->parallel_run_on_sequence(seq, should_specialize != 0); });
}
int msvh_msv_parallel_run_on_packed(void* m, void* packed, float* scores) {
    return guarded([&] {
        const auto got = static_cast<MSV_HMM*>(m)->parallel_run_on_sequences(*static_cast<Packed_sequences*>(packed));
        if (!got.empty()) std::memcpy(scores, got.data(), got.size() * sizeof(float));
    });
}

void* msvh_device_database_create(void* packed, int device) {
    Device_database* d = nullptr;
    guarded([&] { d = new Device_database(*static_cast<Packed_sequences*>(packed), device); });
    return d;
}
void msvh_device_database_free(void* d) { delete static_cast<Device_database*>(d); }
int msvh_msv_parallel_run_on_device_database(void* m, void* d, float* scores) {
    return guarded([&] {
        const auto got = static_cast<MSV_HMM*>(m)->parallel_run_on_sequences(*static_cast<Device_database*>(d));
        if (!got.empty()) std::memcpy(scores, got.data(), got.size() * sizeof(float));
    });
}

// hits are returned as four parallel arrays of capacity `capacity`; the return value is the number of hits (or < 0)
long msvh_msv_filter(void* m, void* d, float threshold, size_t capacity, uint64_t* index, float* score, float* bits, float* p) {
    long found = -1;
    const int status = guarded([&] {
        const auto hits = static_cast<MSV_HMM*>(m)->msv_filter(*static_cast<Device_database*>(d), threshold);
        found = static_cast<long>(hits.size());
        for (size_t i = 0; i < hits.size() && i < capacity; ++i) {
            index[i] = hits[i].sequence;
            score[i] = hits[i].score;
            bits[i] = hits[i].bits;
            p[i] = hits[i].p_value;
        }
    });
    return status == 0 ? found : status;
}

int msvh_msv_parallel_run_on_packed_devices(void* m, void* packed, const int* devices, int n_devices, int gather, float* scores) {
    return guarded([&] {
        const auto got = static_cast<MSV_HMM*>(m)->parallel_run_on_sequences(*static_cast<Packed_sequences*>(packed),
                                                                             std::vector<int>(devices, devices + n_devices),
                                                                             static_cast<MSV_HMM::Score_gather>(gather));
        if (!got.empty()) std::memcpy(scores, got.data(), got.size() * sizeof(float));
    });
}

// ---- Viterbi_HMM ----
void* msvh_viterbi_create(void* profile, int device) {
    Viterbi_HMM* m = nullptr;
    guarded([&] {
        m = new Viterbi_HMM(*static_cast<Profile_HMM*>(profile));
        m->set_device(device);
    });
    return m;
}
void msvh_viterbi_free(void* m) { delete static_cast<Viterbi_HMM*>(m); }
int msvh_viterbi_parallel_run_on_sequence(void* m, const char* seq, float* score) {
    return guarded([&] { *score = static_cast<Viterbi_HMM*>(m)->parallel_run_on_sequence(seq); });
}
int msvh_viterbi_parallel_run_on_packed(void* m, void* packed, float* scores) {
    return guarded([&] {
        const auto got = static_cast<Viterbi_HMM*>(m)->parallel_run_on_sequences(*static_cast<Packed_sequences*>(packed));
        if (!got.empty()) std::memcpy(scores, got.data(), got.size() * sizeof(float));
    });
}
long msvh_viterbi_filter(void* m, void* d, float threshold, size_t capacity, uint64_t* index, float* score, float* bits, float* p) {
    long found = -1;
    const int status = guarded([&] {
        const auto hits = static_cast<Viterbi_HMM*>(m)->viterbi_filter(*static_cast<Device_database*>(d), threshold);
        found = static_cast<long>(hits.size());
        for (size_t i = 0; i < hits.size() && i < capacity; ++i) {
            index[i] = hits[i].sequence;
            score[i] = hits[i].score;
            bits[i] = hits[i].bits;
            p[i] = hits[i].p_value;
        }
    });
    return status == 0 ? found : status;
}
long msvh_viterbi_filter_survivors(void* m, void* d, float threshold, size_t capacity, uint64_t* index, float* score, float* bits, float* p) {
    long found = -1;
    const int status = guarded([&] {
        const auto hits = static_cast<Viterbi_HMM*>(m)->viterbi_filter_survivors(*static_cast<Device_database*>(d), threshold);
        found = static_cast<long>(hits.size());
        for (size_t i = 0; i < hits.size() && i < capacity; ++i) {
            index[i] = hits[i].sequence;
            score[i] = hits[i].score;
            bits[i] = hits[i].bits;
            p[i] = hits[i].p_value;
        }
    });
    return status == 0 ? found : status;
}
int msvh_viterbi_parallel_run_on_device_database(void* m, void* d, float* scores) {
    return guarded([&] {
        const auto got = static_cast<Viterbi_HMM*>(m)->parallel_run_on_sequences(*static_cast<Device_database*>(d));
        if (!got.empty()) std::memcpy(scores, got.data(), got.size() * sizeof(float));
    });
}

} // extern "C"
