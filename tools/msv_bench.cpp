// msv_bench -- microsecond-resolution benchmark of the MSV scan through the C++ host interface.
//
// The reference's benchmark programs (algorithms/benchmark_MSV.cpp, benchmark_MSV_1400.cpp) add up per-call times that
// are truncated to whole milliseconds (benchmark_helper.hpp:36-38); with this implementation a call takes well under a
// millisecond, so they print 0.  This harness reports, per model:
//   * per-call latency of MSV_HMM::parallel_run_on_sequence (the reference's device entry point), best of N, in us;
//   * the same sequences through MSV_HMM::run_on_sequence (host);
//   * throughput of MSV_HMM::parallel_run_on_sequences on a synthetic database, in GCUPS: host buffers in and out
//     ("batch"), and with the database uploaded once as a Device_database ("resident").
//
//   build/msv_bench [--models DIR] [--fasta FILE] [--sequences N] [--repeat R] [--only NAME.hmm]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

#include "MSV_HMM.hpp"
#include "Viterbi_HMM.hpp"
#include "Synthetic_database.hpp"

namespace {
using Clock = std::chrono::steady_clock;
double micros_since(Clock::time_point t0) { return std::chrono::duration<double, std::micro>(Clock::now() - t0).count(); }
} // namespace

int main(int argc, char** argv) {
    auto models_dir = std::string("fixtures/profile_HMMs");
    auto fasta_file = std::string("fixtures/FASTA_files/random_FASTA.fsa");
    auto only = std::string();
    auto sequences = size_t(100000);
    auto repeat = 5;
    for (int i = 1; i + 1 < argc; i += 2) {
        const auto key = std::string(argv[i]);
        if (key == "--models") models_dir = argv[i + 1];
        else if (key == "--fasta") fasta_file = argv[i + 1];
        else if (key == "--sequences") sequences = std::strtoull(argv[i + 1], nullptr, 10);
        else if (key == "--repeat") repeat = std::atoi(argv[i + 1]);
        else if (key == "--only") only = argv[i + 1];
    }

    auto fasta = FASTA_protein_sequences(fasta_file);
    auto residues_in_fasta = size_t(0);
    for (const auto& seq : fasta.sequences) residues_in_fasta += seq.size() - 1;
    const auto database = synthetic_swissprot_like(sequences, 1400);

    auto files = std::vector<std::filesystem::path>();
    for (const auto& entry : std::filesystem::directory_iterator(models_dir))
        if (entry.path().extension() == ".hmm" && (only.empty() || entry.path().filename() == only)) files.push_back(entry.path());
    std::sort(files.begin(), files.end(), [](const auto& a, const auto& b) { return std::stoi(a.stem()) < std::stoi(b.stem()); });

    const auto pinned = Pinned_sequences(database); // "batch" uploads from page-locked memory
    const auto resident = Device_database(database);
    std::printf("%-10s %6s | %12s %12s %10s | %14s %10s | %12s %10s | %12s %10s\n", "model", "LENG", "par us/call", "seq us/call",
                "par GCUPS", "batch ms", "GCUPS", "resident ms", "GCUPS", "Viterbi ms", "GCUPS");
    for (const auto& file : files) {
        auto msv = MSV_HMM(Profile_HMM(file.string()));
        const auto leng = msv.length() - 1;

        auto best_par = 1e300, best_seq = 1e300;
        auto checksum = 0.0f;
        msv.parallel_run_on_sequence(fasta.sequences.front()); // model upload, workspace
        for (int r = 0; r < repeat; ++r) {
            auto t0 = Clock::now();
            for (const auto& seq : fasta.sequences) checksum += msv.parallel_run_on_sequence(seq);
            best_par = std::min(best_par, micros_since(t0) / fasta.sequences.size());
        }
        for (int r = 0; r < std::min(repeat, 2); ++r) {
            auto t0 = Clock::now();
            for (const auto& seq : fasta.sequences) checksum -= msv.run_on_sequence(seq);
            best_seq = std::min(best_seq, micros_since(t0) / fasta.sequences.size());
        }

        auto best_batch = 1e300;
        msv.parallel_run_on_sequences(database);
        for (int r = 0; r < repeat; ++r) {
            auto t0 = Clock::now();
            const auto scores = msv.parallel_run_on_sequences(database);
            best_batch = std::min(best_batch, micros_since(t0));
            checksum += scores.front();
        }
        auto best_resident = 1e300;
        msv.parallel_run_on_sequences(resident);
        for (int r = 0; r < repeat; ++r) {
            auto t0 = Clock::now();
            const auto scores = msv.parallel_run_on_sequences(resident);
            best_resident = std::min(best_resident, micros_since(t0));
            checksum -= scores.front();
        }
        // the Plan-7 local Viterbi scan (Viterbi_HMM) over the same resident database
        auto viterbi = Viterbi_HMM(Profile_HMM(file.string()));
        auto best_viterbi = 1e300;
        viterbi.parallel_run_on_sequences(resident);
        for (int r = 0; r < repeat; ++r) {
            auto t0 = Clock::now();
            const auto scores = viterbi.parallel_run_on_sequences(resident);
            best_viterbi = std::min(best_viterbi, micros_since(t0));
            checksum += scores.front();
        }
        const auto cells_per_call = static_cast<double>(leng) * residues_in_fasta / fasta.sequences.size();
        const auto cells_batch = static_cast<double>(leng) * database.total_residues();
        std::printf("%-10s %6zu | %12.1f %12.1f %10.2f | %14.3f %10.1f | %12.3f %10.1f | %12.3f %10.1f   (checksum %g)\n",
                    file.filename().c_str(), leng, best_par, best_seq, cells_per_call / best_par / 1e3, best_batch / 1e3,
                    cells_batch / best_batch / 1e3, best_resident / 1e3, cells_batch / best_resident / 1e3, best_viterbi / 1e3,
                    cells_batch / best_viterbi / 1e3, checksum);
    }
    return 0;
}
