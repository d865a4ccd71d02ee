// msv_registry.hpp -- the kernel registry of libmsv_cuda.so: which instantiations of the scan kernels (msv_kernels.cuh) exist, as
// a table of Geometry entries.  The table is compiled in four parts (msv_registry_part.cu with MSV_REGISTRY_PART = 0..3) so that
// `make -j` builds the ~700 kernel instantiations side by side; msv_cuda.cu joins the parts.
#pragma once
#include <algorithm>
#include <cstddef>

#include "msv_cuda.h"
#include "msv_kernels.cuh"

namespace msv_registry {

// ---- kernel registry ------------------------------------------------------------------------------------------------
// Two kernel families (msv_kernels.cuh):
//   generic : G = 8/16/32 lanes per sequence, whole table in shared memory           (KT = -1 below)
//   warp    : G = 32, table split between shared memory and KT tensor-memory columns per lane (KT = 0, 8, 16, 24)
constexpr int threads_for(int K) { return K <= 20 ? 1024 : K <= 40 ? 768 : K <= 56 ? 640 : 512; }
constexpr int warp_threads_for(int K, int KT) { return K + (KT > 0 ? 8 : 0) <= 24 ? 1024 : K <= 30 ? 768 : K <= 36 ? 640 : 512; }

using Scan_kernel = void (*)(const msv::Scan_params);
struct Geometry {
    int G, K, KT, threads;
    Scan_kernel fn;         // general transitions
    Scan_kernel fn_cj_same; // tr_E_C == tr_E_J bitwise (C is J); same as fn for the generic family
    int variant = 0;        // 0: tensor-memory columns are each lane's lowest; 1: TMEM_AHEAD (they are the highest, loaded a row ahead)
    Scan_kernel fn_cj_same_exact = nullptr; // warp family: when fn_cj_same speculates B = N + move (and verifies), the kernel that never does
    Scan_kernel fn_cj_same_blocks = nullptr; // warp family: speculation in checkpointed blocks (fn_cj_same speculates on whole sequences)
    Scan_kernel fn_group_spec = nullptr;    // lane-group family (G = 4, 8): speculative scan; a failed sequence is repeated exactly inside the kernel
    // (four lanes per sequence: two interleaved copies of the table, see msv_scan_kernel)
    // lane-group family with K % 4 == 2: the two highest columns of a lane are a pair behind the quads, 128 bytes per residue
    size_t shared_bytes() const {
        if (KT < 0) return static_cast<size_t>(MSV_ALPHABET) * (static_cast<size_t>(K / 4) * std::max(G, 8) * 16 + (K % 4 ? 128 : 0));
        return static_cast<size_t>(MSV_ALPHABET) * (K - KT) * G * sizeof(float);
    }
};

// speculative lane-group scan.  Round 1 (failed speculations went to a second launch): ahead of the exact one by 15 % at
// K = 16 and behind it from K = 40 up.  With the exact redo inside the kernel (round 2) it is ahead everywhere the lane-group
// plans are used: 300.hmm (8 x 38) 6.53 -> 8.13 TCUPS, 400.hmm (8 x 52) 7.00 -> 8.16 on 100 k sequences
// (profiles/r02/group_spec_sweep_v1.txt); beyond K = 56 the warp-per-sequence kernel wins anyway (group_spec_sweep_v2.txt)
template <int G, int K> constexpr Scan_kernel group_spec_kernel() {
    if constexpr ((G == 8 && K <= 56) || G == 4) return msv::msv_scan_group_spec_kernel<G, K, threads_for(K)>;
    else return nullptr;
}
template <int G, int K> constexpr Geometry generic_entry() {
    return Geometry{G, K, -1, threads_for(K), msv::msv_scan_kernel<G, K, threads_for(K), false>,
                    msv::msv_scan_kernel<G, K, threads_for(K), true>, 0, nullptr, nullptr, group_spec_kernel<G, K>()};
}
// Speculative rows (B = N + move, verified per sequence; msv_kernels.cuh) are instantiated where B200 sweeps showed them
// ahead of the exact rows (round 1, profiles/r01/sweep_speculation*.txt: +13 % at K = 4, +2..6 % at K = 16..22 and 32..38;
// behind by 5..8 % from K = 48 up).
// The row variant of the warp kernel that is fastest on a database WITHOUT hits, per columns-per-lane K: B200 sweep over the
// fixture models, profiles/r02/speculation_sweep_v2.txt (100 k sequences; exact / whole-sequence / block-wise speculation, e.g.
// K = 44: 9.76 / 10.05 / 9.77 TCUPS, K = 38: 9.16 / 9.32 / 9.57, K = 42: 9.41 / 9.29 / 9.08, K = 48: 9.87 / 9.63 / 8.91).
// It is not monotone in K -- the compiler's schedule of three different loop nests at the register limit -- and beyond
// K = 44 the two row bodies no longer share the instruction cache.  Unmeasured K keep round 1's rule.
enum Rows { kExact = 0, kWhole = 1, kBlocks = 2 };
constexpr Rows quiet_rows(int K) {
    if (K > 44 || K == 26 || K == 28 || K == 42) return kExact;
    if (K == 16 || K == 38) return kBlocks;
    return kWhole;
}
constexpr bool speculation_pays(int K) { return quiet_rows(K) != kExact; }
template <int K, int KT, int T, bool AHEAD> constexpr Scan_kernel cj_same_kernel() {
    if constexpr (speculation_pays(K)) return msv::msv_scan_warp_kernel<K, KT, T, true, AHEAD, 1>;
    else return msv::msv_scan_warp_kernel<K, KT, T, true, AHEAD>;
}
template <int K, int KT, int T, bool AHEAD> constexpr Scan_kernel cj_same_blocks_kernel() {
    if constexpr (speculation_pays(K)) return msv::msv_scan_warp_kernel<K, KT, T, true, AHEAD, 2>;
    else return nullptr;
}
template <int K, int KT, int T, bool AHEAD> constexpr Scan_kernel cj_same_exact_kernel() {
    if constexpr (speculation_pays(K)) return msv::msv_scan_warp_kernel<K, KT, T, true, AHEAD, 0, true>;
    else return nullptr;
}
template <int K, int KT> constexpr Geometry warp_entry() {
    return Geometry{32, K, KT, warp_threads_for(K, KT), msv::msv_scan_warp_kernel<K, KT, warp_threads_for(K, KT), false>,
                    cj_same_kernel<K, KT, warp_threads_for(K, KT), false>(), 0, cj_same_exact_kernel<K, KT, warp_threads_for(K, KT), false>(),
                    cj_same_blocks_kernel<K, KT, warp_threads_for(K, KT), false>()};
}

constexpr int quad_threads_for(int K) { return K <= 12 ? 1024 : K <= 28 ? 768 : 512; }
template <int K, int KT> constexpr Geometry quad_entry() { // four warps (128 lanes) per sequence
    return Geometry{128, K, KT, quad_threads_for(K), msv::msv_scan_quad_kernel<K, KT, quad_threads_for(K), false>,
                    msv::msv_scan_quad_kernel<K, KT, quad_threads_for(K), true>};
}
template <int K, int KT, int T> constexpr Geometry warp_entry_threads() {
    return Geometry{32, K, KT, T, msv::msv_scan_warp_kernel<K, KT, T, false>, cj_same_kernel<K, KT, T, false>(), 0,
                    cj_same_exact_kernel<K, KT, T, false>(), cj_same_blocks_kernel<K, KT, T, false>()};
}

template <int K, int KT, int T> constexpr Geometry warp_entry_ahead() {
    return Geometry{32, K, KT, T, msv::msv_scan_warp_kernel<K, KT, T, false, true>, cj_same_kernel<K, KT, T, true>(), 1,
                    cj_same_exact_kernel<K, KT, T, true>(), cj_same_blocks_kernel<K, KT, T, true>()};
}


// one part of the table (msv_registry_part.cu)
const Geometry* part0(size_t* count);
const Geometry* part1(size_t* count);
const Geometry* part2(size_t* count);
const Geometry* part3(size_t* count);

} // namespace msv_registry
