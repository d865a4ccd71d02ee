// fasta_cuda.cu -- FASTA text -> packed database on the device (placeholder translation unit; filled in below).
#include "msv_internal.hpp"
