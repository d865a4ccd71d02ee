#pragma once
// MSV_HMM -- Multiple Segment Viterbi scoring of protein sequences against a profile HMM, B200 edition.
//
// Drop-in for the reference's algorithms/MSV_HMM.hpp:9-24: same aliases, same class name, same three public entry
// points with the same argument meaning, so test_MSV.cpp, benchmark_MSV.cpp and benchmark_MSV_1400.cpp compile and
// run unchanged.  What is underneath is new:
//   * parallel_run_on_sequence() no longer builds an OpenCL context, JIT-compiles kernels and issues 13 launches per
//     residue (reference MSV_HMM.cpp:287-423); it calls the sm_100a CUDA library through the C ABI in
//     include/msv_cuda.h (one launch per call).  There is no CPU fallback: without a B200 it throws.
//   * parallel_run_on_sequences() (new) scores a whole database in one launch -- the throughput entry point.
//   * `should_specialize` is kept for source compatibility.  The reference used it to pick JIT-specialised kernels,
//     which silently changed the arithmetic (constants rounded to 6 decimals, MSV_HMM.cpp:325-336).  Here every model
//     always runs a kernel specialised (at compile time) for its geometry, and both values of the flag return the
//     same bit-exact fp32 score.
//   * the object is cheap to copy and move (benchmark_MSV.cpp:35-36 stores it in a growing vector): the device-side
//     model is shared between copies and released with the last one.

#include <cstddef>
#include <memory>
#include <string>
#include <vector>

#include "FASTA_protein_sequences.hpp"
#include "Packed_sequences.hpp"
#include "Profile_HMM.hpp"

typedef float Log_score; // natural-log odds, fp32 like the reference

// aliases of the reference's header that no caller uses; kept so that code naming them still compiles
template <int N> using Log_scores_array = std::array<Log_score, N>;
template <int N> using Log_scores_arrays_vector = std::vector<Log_scores_array<N>>;
typedef std::string Kernels_source_code; // kernels are compiled in; nothing is read at run time

struct msv_model; // opaque device model of the C ABI
struct msv_db;    // opaque device-resident database of the C ABI
struct msv_multi; // opaque multi-GPU handle of the C ABI

// A sequence database uploaded once and kept in HBM (validated, bucketed longest-first), to be scanned by any number of
// models without touching the host again -- the natural shape of "all models against one database"
// (reference benchmark_MSV.cpp:26-41 loops 24 models over the same sequences).  Cheap to copy (shared handle).
class Device_database {
  public:
    explicit Device_database(const Packed_sequences& database, int device = 0);

    size_t size() const { return sequences; }
    uint64_t total_residues() const { return residues; }
    int device() const { return device_index; }
    msv_db* handle() const { return resident.get(); }

  private:
    std::shared_ptr<msv_db> resident;
    size_t sequences = 0;
    uint64_t residues = 0;
    int device_index = 0;
};

// One sequence that passed the MSV filter (HMMER3 conventions: bit score against the null length model, Gumbel
// P-value with the model's STATS LOCAL MSV parameters).
struct MSV_hit {
    size_t sequence;   // index in the database
    Log_score score;   // raw MSV score, nats (what parallel_run_on_sequence returns)
    float bits;
    float p_value;
};

class MSV_HMM {
  public:
    explicit MSV_HMM(const Profile_HMM& base_hmm);

    // CPU entry point of the API: scalar fp32 recurrence on the host (one rolling row).
    Log_score run_on_sequence(const Protein_sequence& seq);

    // GPU entry point of the API: one sequence, synchronous.
    Log_score parallel_run_on_sequence(const Protein_sequence& seq, bool should_specialize = false);

    // GPU, whole database in one launch; scores come back in input order.
    std::vector<Log_score> parallel_run_on_sequences(const Protein_sequences& sequences);
    std::vector<Log_score> parallel_run_on_sequences(const Packed_sequences& database);

    // GPU, database already resident in HBM (Device_database): only the scores cross PCIe.
    std::vector<Log_score> parallel_run_on_sequences(const Device_database& database);

    // The MSV *filter*: scan, convert to bit scores and P-values on the device, keep sequences with P <= threshold
    // (HMMER3's default first-stage threshold F1 is 0.02).  The reference parses the model's Gumbel parameters
    // (Profile_HMM.hpp:34-35) but stops at the raw score.
    std::vector<MSV_hit> msv_filter(const Device_database& database, float threshold = 0.02f);

    // The same over several GPUs of one box from ONE process (msv_cuda_multi_score_batch): the database is cut into
    // contiguous slices of equal cell count, one host thread and one device model per GPU.  `gather` says how the scores
    // come together: every GPU downloads its slice straight into the result (host), the scan kernels store them into one
    // array on the first GPU over NVLink peer access (peer), or one grouped NCCL send/receive round does (nccl).  Same bits.
    // (The one-process-per-GPU variant is hmm_fasta_viterbi_b200/sharded.py + bench.py.)
    enum class Score_gather { host = 0, peer = 1, nccl = 2 };
    std::vector<Log_score> parallel_run_on_sequences(const Packed_sequences& database, const std::vector<int>& devices,
                                                     Score_gather gather = Score_gather::host);

    size_t length() const { return model_length; } // LENG + 1
    int device() const { return device_index; }
    void set_device(int device); // drops the device model; it is re-created lazily on the chosen GPU

  private:
    size_t model_length;
    std::vector<Log_score> emission_scores; // [NUM_OF_AMINO_ACIDS][model_length], column 0 is the -inf begin column

    Log_score tr_B_Mk; // B -> M_k, uniform entry
    Log_score tr_E_C;  // E -> C
    Log_score tr_E_J;  // E -> J
    float msv_mu = 0.0f, msv_lambda = 0.0f; // Gumbel location / slope of MSV bit scores (STATS LOCAL MSV)

    int device_index = 0;
    std::shared_ptr<msv_model> device_model; // created on first GPU call, shared by copies

    // several GPUs from one process: one device model per GPU + the C ABI's multi-GPU handle, built on first use
    struct Replicas;
    std::shared_ptr<Replicas> replicas;

    msv_model* on_device();
};
