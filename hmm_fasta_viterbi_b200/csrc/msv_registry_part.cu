// msv_registry_part.cu -- one quarter of the kernel registry (msv_registry.hpp), compiled four times with
// -DMSV_REGISTRY_PART=0..3: the instantiations of the scan kernels are what takes the compile time of this library, and in
// four translation units `make -j` builds them side by side.  -DMSV_QUICK_BUILD (development aid) keeps only what a
// 1400-column model needs, all in part 0, so that a kernel experiment compiles in seconds.
#define MSV_KERNELS_TEMPLATES_ONLY // the non-template kernels of msv_kernels.cuh belong to msv_cuda.cu alone
#include "msv_registry.hpp"

#ifndef MSV_REGISTRY_PART
#error "compile with -DMSV_REGISTRY_PART=0..3"
#endif

namespace msv_registry {

#define MSV_FOR_EACH_K(X, A)                                                                                           \
    X(A, 4) X(A, 8) X(A, 12) X(A, 16) X(A, 20) X(A, 24) X(A, 28) X(A, 32) X(A, 36) X(A, 40) X(A, 44) X(A, 48) X(A, 52)  \
    X(A, 56) X(A, 60) X(A, 64) X(A, 68) X(A, 72) X(A, 76) X(A, 80) X(A, 84) X(A, 88)
#define MSV_FOR_EACH_K_TO_56(X, A)                                                                                     \
    X(A, 4) X(A, 8) X(A, 12) X(A, 16) X(A, 20) X(A, 24) X(A, 28) X(A, 32) X(A, 36) X(A, 40) X(A, 44) X(A, 48) X(A, 52) X(A, 56)
#define MSV_FOR_EACH_K_FROM_24(X, A)                                                                                   \
    X(A, 24) X(A, 28) X(A, 32) X(A, 36) X(A, 40) X(A, 44) X(A, 48) X(A, 52) X(A, 56) X(A, 60) X(A, 64) X(A, 68) X(A, 72) \
    X(A, 76) X(A, 80) X(A, 84) X(A, 88)
#define MSV_FOR_EACH_K_FROM_28(X, A)                                                                                   \
    X(A, 28) X(A, 32) X(A, 36) X(A, 40) X(A, 44) X(A, 48) X(A, 52) X(A, 56) X(A, 60) X(A, 64) X(A, 68) X(A, 72) X(A, 76)  \
    X(A, 80) X(A, 84) X(A, 88)
// lane groups with K % 4 == 2 (less padding for the short models; G = 4 and 8 only)
#define MSV_FOR_EACH_K_PAIR(X, A)                                                                                      \
    X(A, 6) X(A, 10) X(A, 14) X(A, 18) X(A, 22) X(A, 26) X(A, 30) X(A, 34) X(A, 38) X(A, 42) X(A, 46) X(A, 50) X(A, 54)
#define MSV_GENERIC(G, K) generic_entry<G, K>(),
#define MSV_WARP(KT, K) warp_entry<K, KT>(),
#define MSV_WARP_AHEAD(KT, K) warp_entry_ahead<K, KT, warp_threads_for(K, KT)>(),

#ifdef MSV_QUICK_BUILD
#if MSV_REGISTRY_PART == 0
static const Geometry entries[] = {generic_entry<32, 44>(), warp_entry_ahead<44, 24, 512>(), warp_entry<44, 16>(), quad_entry<12, 8>(),
                                   generic_entry<4, 28>(), generic_entry<8, 16>()
#ifdef MSV_QUICK_EXTRA
                                   , MSV_QUICK_EXTRA
#endif
};
#define MSV_PART_HAS_ENTRIES
#endif
#elif MSV_REGISTRY_PART == 0 // lane-group family: G = 4, 8, 16, 32 lanes per sequence, whole table in shared memory
static const Geometry entries[] = {MSV_FOR_EACH_K_TO_56(MSV_GENERIC, 4) MSV_FOR_EACH_K(MSV_GENERIC, 8) MSV_FOR_EACH_K(MSV_GENERIC, 16)
                                       MSV_FOR_EACH_K(MSV_GENERIC, 32) MSV_FOR_EACH_K_PAIR(MSV_GENERIC, 4) MSV_FOR_EACH_K_PAIR(MSV_GENERIC, 8)};
#define MSV_PART_HAS_ENTRIES
#elif MSV_REGISTRY_PART == 1 // warp per sequence, tensor-memory columns at the bottom of a lane (variant 0)
static const Geometry entries[] = {MSV_WARP(0, 4) MSV_WARP(0, 8) MSV_WARP(8, 8) MSV_WARP(8, 12) MSV_WARP(8, 16) MSV_WARP(8, 20) MSV_WARP(16, 16)
                                       MSV_WARP(16, 20) MSV_WARP(24, 24) MSV_FOR_EACH_K_FROM_24(MSV_WARP, 16) MSV_FOR_EACH_K_FROM_28(MSV_WARP, 24)
                                           MSV_WARP(0, 44) MSV_WARP(8, 44) warp_entry_threads<44, 16, 640>(),
                                   warp_entry_threads<44, 16, 448>(), warp_entry_threads<44, 16, 384>()};
#define MSV_PART_HAS_ENTRIES
#elif MSV_REGISTRY_PART == 2 // warp per sequence, tensor-memory columns loaded a row ahead (variant 1), K in steps of four
static const Geometry entries[] = {MSV_FOR_EACH_K_FROM_24(MSV_WARP_AHEAD, 16) MSV_FOR_EACH_K_FROM_28(MSV_WARP_AHEAD, 24) MSV_WARP_AHEAD(16, 16)
                                       MSV_WARP_AHEAD(16, 20) MSV_WARP_AHEAD(8, 8) MSV_WARP_AHEAD(8, 12)};
#define MSV_PART_HAS_ENTRIES
#elif MSV_REGISTRY_PART == 3 // variant 1 with K in steps of two (tensor-memory part 18 = 16 + 2 columns), and four warps per sequence
static const Geometry entries[] = {MSV_WARP_AHEAD(6, 6) MSV_WARP_AHEAD(10, 10) MSV_WARP_AHEAD(14, 14) MSV_WARP_AHEAD(18, 18) MSV_WARP_AHEAD(18, 22)
                                       MSV_WARP_AHEAD(18, 26) MSV_WARP_AHEAD(18, 30) MSV_WARP_AHEAD(18, 34) MSV_WARP_AHEAD(18, 38)
                                           MSV_WARP_AHEAD(18, 42) MSV_WARP_AHEAD(18, 46) MSV_WARP_AHEAD(18, 50) MSV_WARP_AHEAD(18, 54)
                                               MSV_WARP_AHEAD(18, 58) quad_entry<4, 0>(),
                                   quad_entry<8, 8>(), quad_entry<12, 8>(), quad_entry<16, 16>(), quad_entry<20, 16>(), quad_entry<24, 16>(),
                                   quad_entry<28, 16>(), quad_entry<32, 16>(), quad_entry<36, 16>(), quad_entry<40, 24>(), quad_entry<44, 24>()};
#define MSV_PART_HAS_ENTRIES
#endif

#define MSV_PART_NAME_(n) part##n
#define MSV_PART_NAME(n) MSV_PART_NAME_(n)
const Geometry* MSV_PART_NAME(MSV_REGISTRY_PART)(size_t* count) {
#ifdef MSV_PART_HAS_ENTRIES
    *count = sizeof entries / sizeof entries[0];
    return entries;
#else
    *count = 0;
    return nullptr;
#endif
}

} // namespace msv_registry
