// TEST INFRASTRUCTURE ONLY -- a C handle API over the *reference's own, unmodified* classes.
//
// oracle/Makefile compiles this file together with /root/reference/{algorithms/MSV_HMM.cpp,
// data_readers/Profile_HMM.cpp, data_readers/FASTA_protein_sequences.cpp} (read where they lie, never copied) and
// oracle/cl_stub/CL/cl2.hpp into oracle/_ref/libmsv_ref.so.  It is used to (a) pin oracle/msv_oracle.c,
// (b) generate tests/golden/*.json, (c) serve as bench.py's `cpu_baseline.kind = "reference"` leg.
// Nothing in the product path links or loads it.
//
// The reference keeps the emission table and transition scores private (algorithms/MSV_HMM.hpp:25-40).  To read
// them for the table-parity test, this file re-declares access for the one include below; the class layout is
// unchanged, so the object code of MSV_HMM.cpp is unaffected.
#define private public
#include "MSV_HMM.hpp"
#undef private

#include <cstring>
#include <thread>
#include <vector>

extern "C" {

// ---- Profile_HMM (data_readers/Profile_HMM.hpp:21-49) ----
void* ref_profile_load(const char* path) { return new Profile_HMM(path); }
void ref_profile_free(void* p) { delete static_cast<Profile_HMM*>(p); }
size_t ref_profile_model_length(void* p) { return static_cast<Profile_HMM*>(p)->model_length; }
const char* ref_profile_name(void* p) { return static_cast<Profile_HMM*>(p)->name.c_str(); }
size_t ref_profile_rows(void* p, int which) {
    auto* h = static_cast<Profile_HMM*>(p);
    return which == 0 ? h->match_emissions.size() : which == 1 ? h->insert_emissions.size() : h->transitions.size();
}
void ref_profile_copy(void* p, int which, float* out) {
    auto* h = static_cast<Profile_HMM*>(p);
    if (which == 0)
        for (size_t i = 0; i < h->match_emissions.size(); ++i)
            std::memcpy(out + i * NUM_OF_AMINO_ACIDS, h->match_emissions[i].data(), sizeof(float) * NUM_OF_AMINO_ACIDS);
    else if (which == 1)
        for (size_t i = 0; i < h->insert_emissions.size(); ++i)
            std::memcpy(out + i * NUM_OF_AMINO_ACIDS, h->insert_emissions[i].data(), sizeof(float) * NUM_OF_AMINO_ACIDS);
    else
        for (size_t i = 0; i < h->transitions.size(); ++i)
            std::memcpy(out + i * NUM_OF_TRANSITIONS, h->transitions[i].data(), sizeof(float) * NUM_OF_TRANSITIONS);
}
void ref_profile_stats(void* p, float* out6) {
    auto* h = static_cast<Profile_HMM*>(p);
    out6[0] = h->stats_local_msv_mu;
    out6[1] = h->stats_local_msv_lambda;
    out6[2] = h->stats_local_viterbi_mu;
    out6[3] = h->stats_local_viterbi_lambda;
    out6[4] = h->stats_local_forward_theta;
    out6[5] = h->stats_local_forward_lambda;
}

// ---- FASTA_protein_sequences (data_readers/FASTA_protein_sequences.hpp:9-14) ----
void* ref_fasta_load(const char* path) { return new FASTA_protein_sequences(path); }
void ref_fasta_free(void* p) { delete static_cast<FASTA_protein_sequences*>(p); }
size_t ref_fasta_count(void* p) { return static_cast<FASTA_protein_sequences*>(p)->sequences.size(); }
const char* ref_fasta_record(void* p, size_t i) { return static_cast<FASTA_protein_sequences*>(p)->sequences[i].c_str(); }

// ---- MSV_HMM (algorithms/MSV_HMM.hpp:17-44) ----
void* ref_msv_create(void* profile) { return new MSV_HMM(*static_cast<Profile_HMM*>(profile)); }
void ref_msv_free(void* m) { delete static_cast<MSV_HMM*>(m); }
size_t ref_msv_model_length(void* m) { return static_cast<MSV_HMM*>(m)->model_length; }
void ref_msv_copy_table(void* m, float* out /* [20][model_length] */) {
    auto* msv = static_cast<MSV_HMM*>(m);
    std::memcpy(out, msv->emission_scores.data(), sizeof(float) * msv->emission_scores.size());
}
void ref_msv_transitions(void* m, float* out3) {
    auto* msv = static_cast<MSV_HMM*>(m);
    out3[0] = msv->tr_B_Mk;
    out3[1] = msv->tr_E_C;
    out3[2] = msv->tr_E_J;
}
// MSV_HMM::run_on_sequence, MSV_HMM.cpp:74-113.  `seq` carries the leading '#'.
float ref_msv_run_on_sequence(void* m, const char* seq) { return static_cast<MSV_HMM*>(m)->run_on_sequence(seq); }

// Batch driver for the CPU baseline: codes/offsets -> the reference's string form -> run_on_sequence, one MSV_HMM
// copy per thread because the class mutates tr_loop/tr_move per call (MSV_HMM.cpp:59-64).
void ref_msv_run_batch(void* m, const unsigned char* codes, const unsigned long long* offsets, size_t n, float* scores,
                       int threads) {
    static const char letters[] = "ACDEFGHIKLMNPQRSTVWY";
    if (threads < 1) threads = 1;
    // contiguous slices balanced by residue count: thread t owns sequences [bounds[t], bounds[t+1])
    auto bounds = std::vector<size_t>(static_cast<size_t>(threads) + 1, n);
    bounds[0] = 0;
    const auto total = n ? offsets[n] - offsets[0] : 0ull;
    for (int t = 1; t < threads; ++t) {
        const auto want = offsets[0] + total * static_cast<unsigned long long>(t) / threads;
        auto q = bounds[t - 1];
        while (q < n && offsets[q + 1] <= want) ++q;
        bounds[t] = q;
    }
    auto work = [&](int t) {
        auto local = MSV_HMM(*static_cast<MSV_HMM*>(m));
        auto text = std::string();
        for (auto q = bounds[t]; q < bounds[t + 1]; ++q) {
            text.assign(1, '#');
            for (auto r = offsets[q]; r < offsets[q + 1]; ++r) text.push_back(letters[codes[r]]);
            scores[q] = local.run_on_sequence(text);
        }
    };
    auto pool = std::vector<std::thread>();
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
}

} // extern "C"
