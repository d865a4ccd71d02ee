#include "FASTA_protein_sequences.hpp"

#include <array>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <string_view>

// Record rules (behaviour of the reference's data_readers/FASTA_protein_sequences.cpp:9-44):
//   * a line whose first character is '>' opens a new record; the header text is discarded;
//   * every other line is appended verbatim to the open record (no trimming, so a '\r' or a blank inside a line is
//     a foreign character);
//   * after reading, a record that holds any character other than the 20 amino-acid letters ACDEFGHIKLMNPQRSTVWY is
//     dropped as a whole with a warning on stdout; the survivors keep their relative order.
// Differences by design: the file is read once into memory and classified with a 256-entry table instead of a hash
// lookup per character (the reference's filter is its slowest part on large databases), and text before the first
// '>' is ignored instead of being undefined behaviour.

namespace {

constexpr auto make_residue_table() {
    auto table = std::array<bool, 256>{};
    for (const char c : std::string_view("ACDEFGHIKLMNPQRSTVWY")) table[static_cast<unsigned char>(c)] = true;
    return table;
}
constexpr auto is_residue = make_residue_table();

} // namespace

FASTA_protein_sequences::FASTA_protein_sequences(const std::string& file_path) {
    auto file = std::ifstream(file_path, std::ios::binary);
    if (file.fail()) {
        std::cout << "Failed to open " << file_path << '\n';
        return;
    }
    file.seekg(0, std::ios::end);
    auto text = std::string(static_cast<size_t>(std::max<std::streamoff>(file.tellg(), 0)), '\0');
    file.seekg(0, std::ios::beg);
    file.read(text.data(), static_cast<std::streamsize>(text.size()));
    auto rest = std::string_view(text);

    auto record_open = false;
    auto foreign = '\0'; // first foreign character of the open record, if any
    const auto close_record = [&] {
        if (record_open && foreign != '\0') {
            std::cout << "Warning: sequence " << sequences.back() << " was rejected.\nReason: prohibited symbol " << foreign
                      << " in " << file_path << " FASTA file\n\n";
            sequences.pop_back();
        }
        foreign = '\0';
    };

    while (!rest.empty()) {
        const auto eol = rest.find('\n');
        const auto line = rest.substr(0, eol);
        rest.remove_prefix(eol == std::string_view::npos ? rest.size() : eol + 1);
        if (!line.empty() && line.front() == '>') {
            close_record();
            sequences.emplace_back("#");
            record_open = true;
        } else if (record_open) {
            for (const char c : line)
                if (!is_residue[static_cast<unsigned char>(c)] && c != '#' && foreign == '\0') foreign = c;
            sequences.back().append(line);
        }
    }
    close_record();
}
