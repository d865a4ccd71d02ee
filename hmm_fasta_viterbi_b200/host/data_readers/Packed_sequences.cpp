#include "Packed_sequences.hpp"

#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>

#include "msv_cuda.h"

namespace {
constexpr char letters[] = "ACDEFGHIKLMNPQRSTVWY";
}

void Packed_sequences::append(const Protein_sequence& seq) {
    const auto skip = static_cast<size_t>(!seq.empty() && seq.front() == '#');
    const auto count = seq.size() - skip;
    const auto at = residues.size();
    residues.resize(at + count);
    auto bad_at = size_t(0);
    if (msv_host_encode(seq.data() + skip, count, residues.data() + at, &bad_at) != MSV_OK) {
        residues.resize(at);
        throw std::out_of_range(std::string("Packed_sequences: ") + msv_cuda_last_error());
    }
    offsets.push_back(residues.size());
}

Packed_sequences Packed_sequences::from_sequences(const Protein_sequences& sequences) {
    auto packed = Packed_sequences();
    auto total = size_t(0);
    for (const auto& seq : sequences) total += seq.size();
    packed.residues.reserve(total);
    packed.offsets.reserve(sequences.size() + 1);
    for (const auto& seq : sequences) packed.append(seq);
    return packed;
}

Packed_sequences Packed_sequences::from_fasta_file(const std::string& file_path, size_t* rejected) {
    auto packed = Packed_sequences();
    if (rejected) *rejected = 0;
    auto file = std::unique_ptr<std::FILE, int (*)(std::FILE*)>(std::fopen(file_path.c_str(), "rb"), &std::fclose);
    if (!file) throw std::runtime_error("Failed to open " + file_path);

    // letter -> code, 0xff for everything else ('\n' is handled before the lookup)
    uint8_t code_of[256];
    std::memset(code_of, 0xff, sizeof code_of);
    for (int i = 0; i < MSV_ALPHABET; ++i) code_of[static_cast<unsigned char>(letters[i])] = static_cast<uint8_t>(i);

    auto chunk = std::vector<char>(1 << 20);
    auto in_header = false, line_start = true, open = false, poisoned = false;
    auto record_begin = size_t(0);
    const auto close_record = [&] {
        if (!open) return;
        if (poisoned) {
            packed.residues.resize(record_begin);
            if (rejected) ++*rejected;
        } else {
            packed.offsets.push_back(packed.residues.size());
        }
        open = poisoned = false;
    };
    for (;;) {
        const auto got = std::fread(chunk.data(), 1, chunk.size(), file.get());
        if (got == 0) break;
        packed.residues.reserve(packed.residues.size() + got);
        for (size_t i = 0; i < got; ++i) {
            const auto c = chunk[i];
            if (c == '\n') {
                in_header = false;
                line_start = true;
                continue;
            }
            if (line_start && c == '>') {
                close_record();
                open = true;
                record_begin = packed.residues.size();
                in_header = true;
            } else if (!in_header && open) {
                // Any non-letter poisons the record.  (The reference's filter lets a '#' inside a record through,
                // FASTA_protein_sequences.cpp:30, and its scorer then throws on it, MSV_HMM.cpp:101; here such a
                // record is rejected when read.)
                const auto code = code_of[static_cast<unsigned char>(c)];
                if (code == 0xff)
                    poisoned = true;
                else
                    packed.residues.push_back(code);
            }
            line_start = false;
        }
    }
    close_record();
    return packed;
}

Protein_sequence Packed_sequences::to_sequence(size_t q) const {
    auto seq = Protein_sequence(1, '#');
    seq.reserve(length(q) + 1);
    for (auto r = offsets[q]; r < offsets[q + 1]; ++r) seq.push_back(letters[residues[r]]);
    return seq;
}

std::vector<size_t> Packed_sequences::cell_balanced_bounds(int parts) const {
    auto bounds = std::vector<size_t>(static_cast<size_t>(parts > 0 ? parts : 1) + 1);
    if (msv_host_partition_by_cells(offsets.data(), size(), parts, bounds.data()) != MSV_OK)
        throw std::invalid_argument(msv_cuda_last_error());
    return bounds;
}

Packed_sequences Packed_sequences::slice(size_t first, size_t last) const {
    auto part = Packed_sequences();
    part.residues.assign(residues.begin() + static_cast<std::ptrdiff_t>(offsets[first]),
                         residues.begin() + static_cast<std::ptrdiff_t>(offsets[last]));
    part.offsets.resize(last - first + 1);
    for (auto q = first; q <= last; ++q) part.offsets[q - first] = offsets[q] - offsets[first];
    return part;
}
