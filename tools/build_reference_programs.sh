#!/bin/bash
# Compile the reference's OWN, UNCHANGED test and benchmark programs against this implementation.
#
# Sources are read where they lie under $REF (nothing is copied into the repository):
#   algorithms/test_MSV.cpp, algorithms/benchmark_MSV.cpp, algorithms/benchmark_MSV_1400.cpp,
#   data_readers/test_hmm_parsing.cpp, data_readers/test_fasta_parsing.cpp
# They include "MSV_HMM.hpp" / "Profile_HMM.hpp" / "FASTA_protein_sequences.hpp" / "benchmark_helper.hpp"; the first
# three resolve to hmm_fasta_viterbi_b200/host (this implementation), benchmark_helper.hpp is the reference's own
# header-only timing loop.  The reference's directory contract is reproduced (compile_clang_in_build_dir.sh:3-7):
#   build/profile_HMMs, build/FASTA_files, programs in build/algorithms and build/data_readers (they open ../...).
# build/ is git-ignored but travels to the GPU box, where /root/reference does not exist.
set -euo pipefail
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/hmm_fasta_viterbi_b200
HOST=$PKG/host
OUT=$ROOT/build
CXX=${CXX:-g++}

# this repository's own microsecond-resolution harness (no reference sources involved)
mkdir -p "$OUT"
$CXX -std=c++20 -O2 -Wall -Wextra -I"$HOST/algorithms" -I"$HOST/data_readers" -I"$ROOT/include" "$ROOT/tools/msv_bench.cpp" \
    -o "$OUT/msv_bench" -L"$PKG" -lmsv_host -lmsv_cuda -Wl,-rpath,"$PKG" -Wl,-rpath,'$ORIGIN/../hmm_fasta_viterbi_b200'
$CXX -std=c++20 -O2 -Wall -Wextra -I"$HOST/algorithms" -I"$HOST/data_readers" -I"$ROOT/include" "$ROOT/tools/msv_scan.cpp" \
    -o "$OUT/msv_scan" -L"$PKG" -lmsv_host -lmsv_cuda -Wl,-rpath,"$PKG" -Wl,-rpath,'$ORIGIN/../hmm_fasta_viterbi_b200'

[ -d "$REF/algorithms" ] || { echo "reference tree $REF not present; keeping prebuilt build/ (if any)"; exit 0; }
mkdir -p "$OUT/algorithms" "$OUT/data_readers"
cp -ru "$ROOT/fixtures/profile_HMMs" "$OUT/"
cp -ru "$ROOT/fixtures/FASTA_files" "$OUT/"

# the reference's flags (CMakeLists.txt:4-6) with a portable -march; asserts stay enabled (no -DNDEBUG) so that the
# two reader tests are real tests
FLAGS="-std=c++20 -Wall -Wextra -pedantic -Werror -Wno-unused-function -Wno-unused-variable -march=x86-64-v3 -O3 -pipe"
INC="-I$HOST/algorithms -I$HOST/data_readers -I$ROOT/include"
LIBS="-L$PKG -lmsv_host -lmsv_cuda -Wl,-rpath,$PKG -Wl,-rpath,\$ORIGIN/../../hmm_fasta_viterbi_b200 -lstdc++fs"

# The programs include their headers with quotes, which the compiler resolves next to the including file first.  A
# directory of symbolic links puts the reference's sources (links into $REF, not copies) next to THIS implementation's
# headers, so "MSV_HMM.hpp" etc. resolve to hmm_fasta_viterbi_b200/host while benchmark_helper.hpp stays the
# reference's own.
SRC=$OUT/src
rm -rf "$SRC" && mkdir -p "$SRC/algorithms" "$SRC/data_readers"
for f in test_MSV.cpp benchmark_MSV.cpp benchmark_MSV_1400.cpp benchmark_helper.hpp; do ln -s "$REF/algorithms/$f" "$SRC/algorithms/$f"; done
for f in test_hmm_parsing.cpp test_fasta_parsing.cpp; do ln -s "$REF/data_readers/$f" "$SRC/data_readers/$f"; done
for f in MSV_HMM.hpp; do ln -s "$HOST/algorithms/$f" "$SRC/algorithms/$f"; done
for f in Profile_HMM.hpp FASTA_protein_sequences.hpp Packed_sequences.hpp; do ln -s "$HOST/data_readers/$f" "$SRC/data_readers/$f"; done

for prog in test_MSV benchmark_MSV benchmark_MSV_1400; do
    $CXX $FLAGS $INC "$SRC/algorithms/$prog.cpp" -o "$OUT/algorithms/$prog" $LIBS
done
for prog in test_hmm_parsing test_fasta_parsing; do
    $CXX $FLAGS $INC "$SRC/data_readers/$prog.cpp" -o "$OUT/data_readers/$prog" $LIBS
done
echo "reference programs built in $OUT"
