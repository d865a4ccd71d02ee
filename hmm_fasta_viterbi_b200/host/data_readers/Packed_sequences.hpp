#pragma once
// Packed_sequences -- the device-facing layout of a sequence database (new in this implementation; the reference
// hands one std::string at a time to its device path, MSV_HMM.cpp:382-383).
//
//   residues : all sequences back to back, one byte per residue, codes 0..19 in the order
//              A C D E F G H I K L M N P Q R S T V W Y (the order of the .hmm columns, reference MSV_HMM.cpp:29-31);
//              no '#' sentinel
//   offsets  : n + 1 entries; sequence q is residues[offsets[q] .. offsets[q + 1])
//
// This is exactly what the C ABI (include/msv_cuda.h) takes.  Length bucketing (longest first) happens on the device
// when the database is uploaded; cell-balanced sharding for several GPUs is `cell_balanced_bounds`.

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "FASTA_protein_sequences.hpp"

struct Packed_sequences;

// Page-locks the buffers of a Packed_sequences for as long as the guard lives, so that repeated uploads of the same
// database (one scan per model) run at full link speed.  The database must not be modified or moved meanwhile.
class Pinned_sequences {
  public:
    explicit Pinned_sequences(const Packed_sequences& database);
    ~Pinned_sequences();
    Pinned_sequences(const Pinned_sequences&) = delete;
    Pinned_sequences& operator=(const Pinned_sequences&) = delete;

  private:
    const void* residues = nullptr;
    const void* offsets = nullptr;
};

struct Packed_sequences {
    std::vector<uint8_t> residues;
    std::vector<uint64_t> offsets = {0};

    size_t size() const { return offsets.size() - 1; }
    size_t length(size_t q) const { return static_cast<size_t>(offsets[q + 1] - offsets[q]); }
    uint64_t total_residues() const { return offsets.back(); }

    // Append one sequence given as letters (with or without the leading '#').  Throws std::out_of_range on a letter
    // outside the alphabet -- the error the reference raises from unordered_map::at (MSV_HMM.cpp:101,383).
    void append(const Protein_sequence& seq);

    static Packed_sequences from_sequences(const Protein_sequences& sequences);

    // Read a FASTA file straight into packed form with the record rules of FASTA_protein_sequences (records with a
    // foreign character are dropped whole).  `rejected`, when given, receives the number of dropped records.
    static Packed_sequences from_fasta_file(const std::string& file_path, size_t* rejected = nullptr);

    // Sequence q back as the reference's string form "#" + letters.
    Protein_sequence to_sequence(size_t q) const;

    // parts + 1 boundaries of contiguous slices with (nearly) equal residue counts.
    std::vector<size_t> cell_balanced_bounds(int parts) const;
    Packed_sequences slice(size_t first, size_t last) const;
    // the listed sequences, in the listed order (e.g. the survivors of a filter stage)
    Packed_sequences subset(const std::vector<size_t>& indices) const;
};
