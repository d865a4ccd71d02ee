#pragma once
// FASTA_protein_sequences -- protein FASTA reader.
//
// Drop-in for the reference's data_readers/FASTA_protein_sequences.hpp:6-14 (same aliases, same class, same public
// member).  Every record is stored as "#" + residues: index 0 is a sentinel so that residue i of the biological
// sequence sits at index i, which is what MSV_HMM expects (reference MSV_HMM.cpp:61,100).
//
// For the GPU path the same file can be read straight into the packed device layout with
// Packed_sequences::from_fasta_file (Packed_sequences.hpp), which applies identical record rules without building
// one std::string per record.

#include <string>
#include <vector>

using Protein_sequence = std::string;
using Protein_sequences = std::vector<Protein_sequence>;

class FASTA_protein_sequences {
  public:
    explicit FASTA_protein_sequences(const std::string& file_path);

    Protein_sequences sequences;
};
