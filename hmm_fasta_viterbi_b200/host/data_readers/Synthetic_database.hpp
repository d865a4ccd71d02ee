#pragma once
// Seeded synthetic protein databases for the benchmark configurations of BASELINE.md section 5 (the reference ships
// only an unseeded 3-sequence generator, FASTA_files/random_FASTA_generator.py).  Everything is generated directly
// in packed form; nothing goes through FASTA text.

#include <cstdint>

#include "Packed_sequences.hpp"

// Swiss-Prot-like: L = clip(round(exp(N(5.70, 0.55^2))), 30, 3000), residues i.i.d. from the background amino-acid
// frequencies (16-bit quantised), engine std::mt19937_64(seed).   Configs 3 and 4.
Packed_sequences synthetic_swissprot_like(size_t count, uint64_t seed);

// Titin-like: L ~ U[shortest, longest], residues uniform over the 20 letters.   Config 5.
Packed_sequences synthetic_long_uniform(size_t count, uint64_t seed, size_t shortest, size_t longest);
