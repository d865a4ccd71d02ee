// host_selftest.cpp -- exercises the C++ host layer (readers, packed layout, MSV_HMM::run_on_sequence) on the fixture
// files and on malformed inputs.  Built with -fsanitize=address,undefined by tests/test_host_cpu.py, so that the text
// parsers are checked for memory errors and undefined behaviour (the reference's readers have several: back() on an
// empty vector for a FASTA file without a header, remove_prefix(npos) on blank lines).
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <stdexcept>
#include <string>

#include "MSV_HMM.hpp"
#include "Synthetic_database.hpp"

namespace {
int failures = 0;
#define CHECK(cond)                                                                                                    \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                              \
            ++failures;                                                                                                \
        }                                                                                                              \
    } while (0)

uint32_t bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, sizeof u);
    return u;
}

std::string write_temp(const std::string& dir, const std::string& name, const std::string& text) {
    const auto path = dir + "/" + name;
    std::ofstream(path, std::ios::binary) << text;
    return path;
}
} // namespace

int main(int argc, char** argv) {
    const auto fixtures = std::string(argc > 1 ? argv[1] : "fixtures");
    const auto scratch = std::string(argc > 2 ? argv[2] : "/tmp");

    // ---- fixture models and sequences; known-answer bits from the reference ----
    auto fasta = FASTA_protein_sequences(fixtures + "/FASTA_files/fasta_like_example.fsa");
    CHECK(fasta.sequences.size() == 4);
    CHECK(fasta.sequences[0] == "#ACDEFGHIKLMNPQTVWY");
    auto model_count = 0;
    for (const auto& entry : std::filesystem::directory_iterator(fixtures + "/profile_HMMs")) {
        if (entry.path().extension() != ".hmm") continue;
        const auto profile = Profile_HMM(entry.path().string());
        CHECK(profile.model_length == static_cast<size_t>(std::stoi(entry.path().stem())) + 1);
        CHECK(profile.match_emissions.size() == profile.model_length);
        CHECK(profile.insert_emissions.size() == profile.model_length);
        CHECK(profile.transitions.size() == profile.model_length);
        auto msv = MSV_HMM(profile);
        auto copy = msv; // copyable, like the reference's class (benchmark_MSV.cpp:35-36)
        const auto score = copy.run_on_sequence(fasta.sequences[0]);
        if (entry.path().filename() == "100.hmm") CHECK(bits(score) == 0xc114d20bu);
        if (entry.path().filename() == "1400.hmm") CHECK(bits(score) == 0xc147de90u);
        if (entry.path().filename() == "2405.hmm") CHECK(bits(score) == 0xc16cb0feu);
        ++model_count;
    }
    CHECK(model_count == 24);

    // ---- packed layout ----
    const auto packed = Packed_sequences::from_sequences(fasta.sequences);
    CHECK(packed.size() == 4 && packed.total_residues() == 181);
    for (size_t q = 0; q < packed.size(); ++q) CHECK(packed.to_sequence(q) == fasta.sequences[q]);
    auto rejected = size_t(99);
    const auto direct = Packed_sequences::from_fasta_file(fixtures + "/FASTA_files/fasta_like_example.fsa", &rejected);
    CHECK(rejected == 0 && direct.residues == packed.residues && direct.offsets == packed.offsets);
    const auto bounds = packed.cell_balanced_bounds(3);
    CHECK(bounds.size() == 4 && bounds.front() == 0 && bounds.back() == 4);
    const auto part = packed.slice(1, 3);
    CHECK(part.size() == 2 && part.to_sequence(0) == fasta.sequences[1] && part.to_sequence(1) == fasta.sequences[2]);
    const auto picked = packed.subset({3, 0, 3});
    CHECK(picked.size() == 3 && picked.to_sequence(0) == fasta.sequences[3] && picked.to_sequence(1) == fasta.sequences[0] &&
          picked.to_sequence(2) == fasta.sequences[3] && picked.offsets.back() == picked.residues.size());
    CHECK(packed.subset({}).size() == 0);
    auto threw = false;
    try {
        packed.subset({4});
    } catch (const std::out_of_range&) {
        threw = true;
    }
    CHECK(threw);
    threw = false;
    try {
        Packed_sequences::from_sequences({"#ACDX"});
    } catch (const std::out_of_range&) {
        threw = true;
    }
    CHECK(threw);
    const auto synthetic = synthetic_swissprot_like(2000, 7);
    CHECK(synthetic.size() == 2000 && synthetic.offsets.back() == synthetic.residues.size());

    // ---- malformed inputs must not crash ----
    CHECK(FASTA_protein_sequences(write_temp(scratch, "empty.fsa", "")).sequences.empty());
    CHECK(FASTA_protein_sequences(write_temp(scratch, "noheader.fsa", "ACDEF\nGHIK\n")).sequences.empty());
    CHECK(FASTA_protein_sequences(write_temp(scratch, "onlyheader.fsa", ">x")).sequences == Protein_sequences{"#"});
    CHECK(Packed_sequences::from_fasta_file(write_temp(scratch, "empty2.fsa", "")).size() == 0);
    CHECK(Packed_sequences::from_fasta_file(write_temp(scratch, "noheader2.fsa", "ACDEF\n\n")).size() == 0);
    CHECK(Packed_sequences::from_fasta_file(write_temp(scratch, "crlf.fsa", ">a\r\nACD\r\n"), &rejected).size() == 0 && rejected == 1);
    {
        auto whole = std::ifstream(fixtures + "/profile_HMMs/100.hmm", std::ios::binary);
        const auto text = std::string(std::istreambuf_iterator<char>(whole), std::istreambuf_iterator<char>());
        for (const auto cut : {size_t(0), size_t(10), size_t(200), size_t(700), size_t(1500), text.size() / 2, text.size() - 5}) {
            const auto truncated = Profile_HMM(write_temp(scratch, "cut.hmm", text.substr(0, cut)));
            CHECK(truncated.match_emissions.size() <= truncated.model_length || truncated.model_length == 0);
        }
        auto blank_lines = text;
        blank_lines.insert(blank_lines.find("STATS"), "\n   \n\n");
        CHECK(Profile_HMM(write_temp(scratch, "blank.hmm", blank_lines)).model_length == 101);
    }
    CHECK(Profile_HMM(scratch + "/does_not_exist.hmm").model_length == 0);

    std::printf(failures ? "host selftest: %d failure(s)\n" : "host selftest ok\n", failures);
    return failures ? 1 : 0;
}
