"""hmm_fasta_viterbi_b200 -- B200-native MSV (Multiple Segment Viterbi) profile-HMM scan.

Layers (bottom up):
  libmsv_cuda.so   csrc/            hand-written sm_100a kernels behind the C ABI of include/msv_cuda.h
  libmsv_host.so   host/            C++ host layer with the reference's interface: Profile_HMM,
                                    FASTA_protein_sequences, MSV_HMM (+ Packed_sequences, synthetic databases,
                                    Viterbi_HMM)
  _cabi.py, host.py                 ctypes marshalling for tests, bench.py and multi-GPU drivers

Importing this package requires both shared libraries to be built (``__graft_entry__.build()``); there is no CPU
fallback for the GPU path.
"""
from . import _cabi
from ._cabi import Database, Model, MsvCudaError, ViterbiModel
from .host import Device_database, FASTA_protein_sequences, MSV_HMM, Packed_sequences, Profile_HMM, Viterbi_HMM

__all__ = ["Database", "Device_database", "Model", "MsvCudaError", "FASTA_protein_sequences", "MSV_HMM", "Packed_sequences", "Profile_HMM",
           "ViterbiModel", "Viterbi_HMM", "_cabi"]
