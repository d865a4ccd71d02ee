// msv_scan -- command-line front end: scan a FASTA database with one or more profile HMMs on the GPU and print either
// every raw MSV score or the sequences that pass the MSV filter.  (The reference's main.cpp only prints "Work in
// progress"; this is the program a user of the library would start from.)
//
//   build/msv_scan [--all] [--F1 0.02] [--viterbi [--F2 0.001]] [--device 0] model.hmm [more.hmm ...] database.fasta
//
// Output (tab separated): model, sequence index (0-based, among the records that survive the reader's filter), length,
// raw score (nats), bit score, P-value.  With --viterbi the sequences that pass the MSV filter are rescored by the Plan-7
// local Viterbi scan (Viterbi_HMM) and only those with a Viterbi P-value <= F2 are printed, with three more columns
// (Viterbi score, bits, P-value): the first two stages of HMMER3's acceleration pipeline.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "MSV_HMM.hpp"
#include "Viterbi_HMM.hpp"

int main(int argc, char** argv) {
    auto all = false;
    auto threshold = 0.02f, threshold2 = 1e-3f;
    auto second_stage = false;
    auto device = 0;
    auto files = std::vector<std::string>();
    for (int i = 1; i < argc; ++i) {
        const auto arg = std::string(argv[i]);
        if (arg == "--all") all = true;
        else if (arg == "--F1" && i + 1 < argc) threshold = std::strtof(argv[++i], nullptr);
        else if (arg == "--F2" && i + 1 < argc) threshold2 = std::strtof(argv[++i], nullptr);
        else if (arg == "--viterbi") second_stage = true;
        else if (arg == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else files.push_back(arg);
    }
    if (files.size() < 2) {
        std::fprintf(stderr, "usage: %s [--all] [--F1 P] [--viterbi [--F2 P]] [--device N] model.hmm [more.hmm ...] database.fasta\n", argv[0]);
        return 2;
    }
    try {
        const auto t0 = std::chrono::steady_clock::now();
        auto rejected = size_t(0);
        const auto database = Packed_sequences::from_fasta_file(files.back(), &rejected);
        const auto resident = Device_database(database, device);
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "# %zu sequences, %llu residues (%zu records rejected) read and uploaded in %.1f ms\n", database.size(),
                     static_cast<unsigned long long>(database.total_residues()), rejected,
                     std::chrono::duration<double, std::milli>(t1 - t0).count());
        std::printf(second_stage ? "#model\tsequence\tlength\tscore_nats\tbits\tp_value\tviterbi_nats\tviterbi_bits\tviterbi_p_value\n"
                                 : "#model\tsequence\tlength\tscore_nats\tbits\tp_value\n");
        for (size_t f = 0; f + 1 < files.size(); ++f) {
            const auto profile = Profile_HMM(files[f]);
            if (profile.model_length == 0) {
                std::fprintf(stderr, "cannot read model %s\n", files[f].c_str());
                return 1;
            }
            auto msv = MSV_HMM(profile);
            msv.set_device(device);
            const auto s0 = std::chrono::steady_clock::now();
            const auto hits = msv.msv_filter(resident, all ? 2.0f : threshold); // P <= 1 always: --all keeps everything
            const auto ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s0).count();
            auto reported = hits.size();
            if (second_stage) {
                auto survivors = std::vector<size_t>();
                for (const auto& hit : hits) survivors.push_back(hit.sequence);
                auto viterbi = Viterbi_HMM(profile);
                viterbi.set_device(device);
                const auto confirmed = viterbi.viterbi_filter(Device_database(database.subset(survivors), device), all ? 2.0f : threshold2);
                for (const auto& second : confirmed) {
                    const auto& first = hits[second.sequence]; // index among the survivors
                    std::printf("%s\t%zu\t%zu\t%.6g\t%.4f\t%.4g\t%.6g\t%.4f\t%.4g\n", profile.name.c_str(), first.sequence,
                                database.length(first.sequence), first.score, first.bits, first.p_value, second.score, second.bits,
                                second.p_value);
                }
                reported = confirmed.size();
            } else {
                for (const auto& hit : hits)
                    std::printf("%s\t%zu\t%zu\t%.6g\t%.4f\t%.4g\n", profile.name.c_str(), hit.sequence, database.length(hit.sequence),
                                hit.score, hit.bits, hit.p_value);
            }
            const auto cells = static_cast<double>(profile.model_length - 1) * static_cast<double>(database.total_residues());
            std::fprintf(stderr, "# %s (LENG %zu): %zu of %zu sequences reported, %.2f ms, %.0f GCUPS\n", profile.name.c_str(),
                         profile.model_length - 1, reported, database.size(), ms, cells / ms / 1e6);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "msv_scan: %s\n", e.what());
        return 1;
    }
    return 0;
}
