#ifndef MSV_B200_FASTA_PROTEIN_SEQUENCES_HPP
#define MSV_B200_FASTA_PROTEIN_SEQUENCES_HPP
// FASTA_protein_sequences -- protein FASTA reader.
//
// Source-compatible with the reference's data_readers/FASTA_protein_sequences.hpp:6-14: the two type names, the
// reader's name, its one-argument constructor and its `sequences` member are what callers written against the
// reference use (test_fasta_parsing.cpp:6, test_MSV.cpp:17, benchmark_helper.hpp:10).  Every record is stored as
// "#" + residues: index 0 is a sentinel so that residue i of the biological sequence sits at index i, which is what
// MSV_HMM expects (reference MSV_HMM.cpp:61,100).
//
// For the GPU path the same file can be read straight into the packed device layout with
// Packed_sequences::from_fasta_file (Packed_sequences.hpp), which applies identical record rules without building
// one std::string per record.

#include <string>
#include <vector>

typedef std::string Protein_sequence;                    // "#ACDEF..."
typedef std::vector<Protein_sequence> Protein_sequences; // one entry per surviving record, file order

struct FASTA_protein_sequences {
    Protein_sequences sequences;

    // Reads `file_path`; an unreadable file leaves `sequences` empty and prints a message (as the reference does).
    explicit FASTA_protein_sequences(const std::string& file_path);
};

#endif
