#include "Packed_sequences.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "msv_cuda.h"

namespace {
constexpr char letters[] = "ACDEFGHIKLMNPQRSTVWY";
}

Pinned_sequences::Pinned_sequences(const Packed_sequences& database) {
    if (msv_cuda_host_register(database.residues.data(), database.residues.size()) == MSV_OK) residues = database.residues.data();
    if (msv_cuda_host_register(database.offsets.data(), database.offsets.size() * sizeof(uint64_t)) == MSV_OK)
        offsets = database.offsets.data();
    // pinning is an optimisation: without a device (or if the driver refuses) the database stays pageable
}

Pinned_sequences::~Pinned_sequences() {
    if (residues) msv_cuda_host_unregister(residues);
    if (offsets) msv_cuda_host_unregister(offsets);
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

namespace {

// Result of parsing one slice of a FASTA file that starts at a record header.
struct Parsed_slice {
    std::vector<uint8_t> residues;
    std::vector<uint64_t> lengths;
    size_t rejected = 0;
};

// letter -> code, 0xff for everything else
struct Code_table {
    uint8_t of[256];
    Code_table() {
        std::memset(of, 0xff, sizeof of);
        for (int i = 0; i < MSV_ALPHABET; ++i) of[static_cast<unsigned char>(letters[i])] = static_cast<uint8_t>(i);
    }
};

// Parse [begin, end): `begin` points at a '>' that starts a line (or begin == end).  Record rules as in
// FASTA_protein_sequences.cpp: header text dropped, other lines appended verbatim, a record with any non-letter is
// dropped whole.  (The reference's filter lets a '#' inside a record through, FASTA_protein_sequences.cpp:30, and its
// scorer then throws on it, MSV_HMM.cpp:101; here such a record is rejected when read.)
void parse_slice(const char* begin, const char* end, Parsed_slice& out) {
    static const Code_table table;
    out.residues.resize(static_cast<size_t>(end - begin)); // upper bound; shrunk at the end
    auto* dst = out.residues.data();
    auto* record_start = dst;
    auto open = false;
    uint8_t poison = 0;
    const auto close_record = [&] {
        if (!open) return;
        if (poison & 0x80) {
            dst = record_start;
            ++out.rejected;
        } else {
            out.lengths.push_back(static_cast<uint64_t>(dst - record_start));
        }
    };
    for (const char* line = begin; line < end;) {
        const auto* eol = static_cast<const char*>(std::memchr(line, '\n', static_cast<size_t>(end - line)));
        const char* stop = eol ? eol : end;
        if (line < stop && *line == '>') {
            close_record();
            open = true;
            poison = 0;
            record_start = dst;
        } else if (open) {
            for (const char* c = line; c < stop; ++c) {
                const auto code = table.of[static_cast<unsigned char>(*c)];
                *dst++ = code;
                poison |= code;
            }
        }
        line = stop + 1;
    }
    close_record();
    out.residues.resize(static_cast<size_t>(dst - out.residues.data()));
}

} // namespace

// The file is mapped, cut at record boundaries into one slice per hardware thread, parsed in parallel and stitched
// together with a prefix sum -- the reference's reader is a single-threaded getline loop plus a hash lookup per
// character (FASTA_protein_sequences.cpp:18-41), which would dominate the end-to-end time of a TCUPS scan.
Packed_sequences Packed_sequences::from_fasta_file(const std::string& file_path, size_t* rejected) {
    auto packed = Packed_sequences();
    if (rejected) *rejected = 0;
    const int fd = ::open(file_path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("Failed to open " + file_path);
    struct stat info {};
    if (::fstat(fd, &info) != 0 || info.st_size == 0) {
        ::close(fd);
        if (info.st_size == 0) return packed;
        throw std::runtime_error("Failed to stat " + file_path);
    }
    const auto size = static_cast<size_t>(info.st_size);
    void* mapping = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (mapping == MAP_FAILED) throw std::runtime_error("Failed to map " + file_path);
    const auto* text = static_cast<const char*>(mapping);
    const char* const text_end = text + size;

    // next record header at or after `from`: a '>' at the start of a line
    const auto next_header = [&](const char* from) -> const char* {
        if (from == text && *from == '>') return from;
        for (const char* p = from; p < text_end;) {
            const auto* nl = static_cast<const char*>(std::memchr(p, '\n', static_cast<size_t>(text_end - p)));
            if (!nl || nl + 1 >= text_end) return text_end;
            if (nl[1] == '>') return nl + 1;
            p = nl + 1;
        }
        return text_end;
    };

    const auto workers = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), size / (4u << 20) + 1));
    auto cuts = std::vector<const char*>(workers + 1, text_end);
    cuts[0] = next_header(text);
    for (size_t w = 1; w < workers; ++w) cuts[w] = std::max(cuts[w - 1], next_header(text + size / workers * w));
    auto slices = std::vector<Parsed_slice>(workers);
    {
        auto pool = std::vector<std::thread>();
        for (size_t w = 1; w < workers; ++w) pool.emplace_back([&, w] { parse_slice(cuts[w], cuts[w + 1], slices[w]); });
        parse_slice(cuts[0], cuts[1], slices[0]);
        for (auto& th : pool) th.join();
    }
    ::munmap(mapping, size);

    auto total_residues = size_t(0), total_records = size_t(0);
    for (const auto& sl : slices) {
        total_residues += sl.residues.size();
        total_records += sl.lengths.size();
        if (rejected) *rejected += sl.rejected;
    }
    packed.residues.resize(total_residues);
    packed.offsets.resize(total_records + 1);
    auto residue_at = size_t(0), record_at = size_t(0);
    auto copies = std::vector<std::thread>();
    for (auto& sl : slices) {
        auto running = static_cast<uint64_t>(residue_at);
        for (const auto len : sl.lengths) {
            running += len;
            packed.offsets[++record_at] = running;
        }
        if (!sl.residues.empty())
            copies.emplace_back([&packed, &sl, residue_at] { std::memcpy(packed.residues.data() + residue_at, sl.residues.data(), sl.residues.size()); });
        residue_at += sl.residues.size();
    }
    for (auto& th : copies) th.join();
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

Packed_sequences Packed_sequences::subset(const std::vector<size_t>& indices) const {
    auto part = Packed_sequences();
    auto total = uint64_t(0);
    for (const auto q : indices) total += offsets.at(q + 1) - offsets[q];
    part.residues.reserve(total);
    part.offsets.reserve(indices.size() + 1);
    for (const auto q : indices) {
        part.residues.insert(part.residues.end(), residues.begin() + static_cast<std::ptrdiff_t>(offsets[q]),
                             residues.begin() + static_cast<std::ptrdiff_t>(offsets[q + 1]));
        part.offsets.push_back(part.residues.size());
    }
    return part;
}
