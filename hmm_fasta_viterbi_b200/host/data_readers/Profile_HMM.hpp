#pragma once
// Profile_HMM -- reader for HMMER3 ASCII profile files (the subset the MSV path needs).
//
// Drop-in for the reference's data_readers/Profile_HMM.hpp:21-49: same type names, same public data members with
// the same meaning, same constructor; callers written against the reference (test_hmm_parsing.cpp, test_MSV.cpp,
// benchmark_MSV*.cpp) compile unchanged.  The parsing rules, including the reference's quirks, are documented in
// Profile_HMM.cpp next to the code that implements them.

#include <array>
#include <cstddef>
#include <string>
#include <string_view>
#include <vector>

enum : int { NUM_OF_AMINO_ACIDS = 20, NUM_OF_TRANSITIONS = 7 }; // columns of an emission row / a transition row

typedef float Probability;
typedef std::string Profile_name;

template <int N> using Probabilities_array = std::array<Probability, N>;
template <int N> using Probabilities_arrays_vector = std::vector<Probabilities_array<N>>;

struct Profile_HMM {
    // Parse `file_path`.  On an unreadable file a message goes to stdout and the object stays empty
    // (model_length == 0), which is what the reference does (Profile_HMM.cpp:49-53).
    explicit Profile_HMM(const std::string& file_path);

    Profile_name name;
    size_t model_length = 0; // LENG + 1 (counts the begin node)

    // Row i describes node i of the model; row 0 is the begin node: match_emissions[0] is all zero,
    // insert_emissions[0] / transitions[0] come from the two lines that follow COMPO.
    // Values are probabilities, exp(-x) of the file's negative natural logs.
    Probabilities_arrays_vector<NUM_OF_AMINO_ACIDS> match_emissions, insert_emissions;
    Probabilities_arrays_vector<NUM_OF_TRANSITIONS> transitions; // m->m m->i m->d i->m i->i d->m d->d

    // STATS LOCAL lines: Gumbel location/slope for MSV and Viterbi scores, exponential tail for Forward scores.
    float stats_local_msv_mu = 0.0f, stats_local_msv_lambda = 0.0f;
    float stats_local_viterbi_mu = 0.0f, stats_local_viterbi_lambda = 0.0f;
    float stats_local_forward_theta = 0.0f, stats_local_forward_lambda = 0.0f;

  private:
    // Consumes the text of one .hmm file; returns false when a mandatory section is missing.
    bool parse(std::string_view text);
};
