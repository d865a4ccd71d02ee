#pragma once
// Viterbi_HMM -- Plan-7 local multihit Viterbi scoring (match, insert and delete states) on the B200.
//
// The reference announces "algorithms based on the Viterbi algorithm" (README.md:2-3) and parses everything such a scan
// needs (Profile_HMM::transitions, data_readers/Profile_HMM.hpp:27-29) but implements only MSV.  This class is the next
// algorithm on the same plumbing, shaped like MSV_HMM (algorithms/MSV_HMM.hpp:17-23): constructed from a Profile_HMM,
// scored through parallel_run_on_sequence(s).  Conventions are those of MSV_HMM so that the two scores are comparable:
// same log-odds emission table, uniform local entry, E -> C / E -> J = log(1/2), length-dependent loop / move scores;
// node transitions are logf() of the parsed probabilities; insert emissions score 0 (HMMER3's profile configuration).
// The recurrence is spelled out in include/msv_cuda.h.  GPU only: there is no host implementation.

#include <cstddef>
#include <memory>
#include <vector>

#include "MSV_HMM.hpp"

struct msv_viterbi_model; // opaque device model of the C ABI

class Viterbi_HMM {
  public:
    explicit Viterbi_HMM(const Profile_HMM& base_hmm);

    // one sequence ("#" + letters, as FASTA_protein_sequences produces), synchronous
    Log_score parallel_run_on_sequence(const Protein_sequence& seq);

    // whole database in one launch; scores come back in input order
    std::vector<Log_score> parallel_run_on_sequences(const Protein_sequences& sequences);
    std::vector<Log_score> parallel_run_on_sequences(const Packed_sequences& database);
    // database already resident in HBM (shared with MSV_HMM: upload once, run both scans)
    std::vector<Log_score> parallel_run_on_sequences(const Device_database& database);

    // The Viterbi *filter*, second stage of HMMER3's pipeline: scan, bit scores and Gumbel P-values (STATS LOCAL VITERBI) on
    // the device; keeps sequences with P <= threshold (HMMER3's default F2 is 1e-3).  Hits reuse MSV_hit.
    std::vector<MSV_hit> viterbi_filter(const Device_database& database, float threshold = 1e-3f);
    // The same restricted to the SURVIVORS of the last MSV_HMM::msv_filter on this database -- HMMER3's pipeline order.  The
    // survivors' index list never left the GPU: the scan reads the resident database through it, nothing is re-packed.
    std::vector<MSV_hit> viterbi_filter_survivors(const Device_database& database, float threshold = 1e-3f);

    size_t length() const { return model_length; } // LENG + 1
    int device() const { return device_index; }
    void set_device(int device);

    // Gumbel parameters of Viterbi bit scores (STATS LOCAL VITERBI), for callers that go on to P-values
    float viterbi_mu() const { return mu; }
    float viterbi_lambda() const { return lambda; }

  private:
    size_t model_length;
    std::vector<Log_score> emission_scores; // [NUM_OF_AMINO_ACIDS][model_length]
    std::vector<Log_score> log_transitions; // [model_length][NUM_OF_TRANSITIONS]
    Log_score tr_B_Mk, tr_E_C, tr_E_J;
    float mu = 0.0f, lambda = 0.0f;

    int device_index = 0;
    std::shared_ptr<msv_viterbi_model> device_model; // created on first use, shared by copies

    msv_viterbi_model* on_device();
};
