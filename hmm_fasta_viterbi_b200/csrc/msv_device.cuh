// msv_device.cuh -- device-side building blocks shared by the scan kernels of libmsv_cuda.so (msv_kernels.cuh,
// viterbi_kernels.cuh): launch parameters, mbarrier / bulk-async (TMA) copies, shared- and tensor-memory accessors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace msv {

constexpr int kAlphabet = 20;
constexpr int kMaxColumnsPerLane = 88;
constexpr int kMaxScoreMirrors = 7; // a fused gather reaches the other GPUs of an 8-GPU NVSwitch domain
constexpr uint32_t kResiduePadBytes = 64; // bytes readable past the last residue of the database

struct Scan_params {
    const float4* table;      // emission table in kernel layout (global memory), table_bytes long
    const uint8_t* residues;  // concatenated residue codes, padded by kResiduePadBytes
    const uint64_t* offsets;  // n + 1
    const uint32_t* order;    // n sequence indices, longest first
    const float2* length_tr;  // (tr_loop, tr_move) indexed by sequence length (host libm, reference MSV_HMM.cpp:59-64)
    float* scores;            // n, original order
    unsigned int* queue_head; // work queue cursor, zero before launch
    const unsigned long long* first_bad; // position of the first invalid residue code found by db_validate_kernel, or ~0
    uint32_t n;
    uint32_t table_bytes;
    float tr_B_Mk, tr_E_C, tr_E_J;
    // fused gather: every score is also stored into these buffers (peer GPUs' copies of the gathered score array, mapped
    // into this GPU's address space over NVLink; already offset to this shard's first sequence).  0 = plain scan.
    uint32_t n_mirrors;
    float* mirrors[kMaxScoreMirrors];
    // lane-group kernels, long sequences on fast CTAs (see next_ticket): the first n_long entries of `order` (the longest
    // sequences) are handed to CTAs [0, fast_ctas), which run with fast_threads threads only -- fewer, faster slots --
    // while the other CTAs start behind them; queue_head[1] is the cursor of that second part.  n_long = 0: one plain queue.
    uint32_t n_long, fast_ctas, fast_threads;
    // running totals for the host's choice of the speculation mode (may be NULL): [0] sequences whose speculation failed,
    // [1] sequences offered to a speculating warp kernel
    unsigned int* speculation_failures;
};

// Work queue of the lane-group kernels.  `order` lists the sequences longest first.  A lane group scans a sequence at 1/slots
// of its SM's speed, so on a CTA with hundreds of slots the few longest sequences of a database would finish long after
// everything else (3000 rows at ~530 clocks per row against ~0.5 ms for the whole scan of 100 k sequences).  Cutting the slot
// count of EVERY CTA (what the planner did before) trades throughput for it everywhere; instead a few "fast" CTAs run with
// a fraction of the threads, take the long sequences first (tickets [0, n_long)) and carry on with the rest, while the full
// CTAs work through tickets [n_long, n) only.
__device__ __forceinline__ uint32_t next_ticket(const Scan_params& p, const bool fast_cta) {
    if (p.n_long == 0) return atomicAdd(p.queue_head, 1u);
    if (fast_cta) {
        const uint32_t t = atomicAdd(p.queue_head, 1u);
        if (t < p.n_long) return t;
    }
    // (a full CTA never takes a long sequence, not even when nothing else is left: started late on a slow slot it would
    // finish long after the fast CTAs have dealt with the rest)
    return p.n_long + atomicAdd(p.queue_head + 1, 1u);
}

// the one store per sequence: local result plus, for the fused gather, the same 4 bytes into every peer's array
__device__ __forceinline__ void store_score(const Scan_params& p, uint32_t idx, float score) {
    p.scores[idx] = score;
    for (uint32_t r = 0; r < p.n_mirrors; ++r) p.mirrors[r][idx] = score;
}

// ---- mbarrier / bulk-async (TMA) helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbarrier_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbarrier_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbarrier_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// global -> shared bulk copy executed by the TMA unit (SASS: UBLKCP); completion is counted in bytes on `bar`.
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- group max --------------------------------------------------------------------------------------------------
template <int G> __device__ __forceinline__ float group_max(float v) {
    if constexpr (G == 32) {
        float r;
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
        return r;
    } else {
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
        return v;
    }
}

// ---- shared-memory loads with explicit 32-bit addresses -----------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t shared_address) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(shared_address));
    return v;
}

__device__ __forceinline__ float2 lds64(uint32_t shared_address) {
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(shared_address));
    return v;
}

// same, but never hoisted out of a loop or merged: for loop-invariant tables that must NOT be promoted to registers
__device__ __forceinline__ float4 lds128_volatile(uint32_t shared_address) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(shared_address));
    return v;
}

// ---- tensor memory (TMEM) accessors: tcgen05.ld/st.32x32b.xN gives thread t of a warp N consecutive 32-bit columns of
// TMEM lane 32*(warp%4)+t ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_load8(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr));
}

__device__ __forceinline__ void tmem_load16(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr));
}

// tcgen05.wait::ld; the loaded registers are in/out operands so that no use of them can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait8(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])::"memory");
}
__device__ __forceinline__ void tmem_wait16(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])::"memory");
}

__device__ __forceinline__ void tmem_store2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tmem_load2(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_load4(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait2(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1])::"memory");
}
__device__ __forceinline__ void tmem_wait4(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3])::"memory");
}

// KT (even, <= 24) columns per lane as a sum of power-of-two pieces: 24 = 16+8, 18 = 16+2, 14 = 8+4+2, ...
template <int KT> __device__ __forceinline__ void tmem_load(uint32_t taddr, float* v) {
    static_assert(KT >= 2 && KT <= 24 && KT % 2 == 0, "TMEM columns per lane");
    constexpr int P = KT >= 16 ? 16 : KT >= 8 ? 8 : KT >= 4 ? 4 : 2;
    if constexpr (P == 16) tmem_load16(taddr, v);
    if constexpr (P == 8) tmem_load8(taddr, v);
    if constexpr (P == 4) tmem_load4(taddr, v);
    if constexpr (P == 2) tmem_load2(taddr, v);
    if constexpr (KT > P) tmem_load<KT - P>(taddr + P, v + P);
}
// (the waits after the first are free: tcgen05.wait::ld covers every earlier load)
template <int KT> __device__ __forceinline__ void tmem_wait(float* v) {
    constexpr int P = KT >= 16 ? 16 : KT >= 8 ? 8 : KT >= 4 ? 4 : 2;
    if constexpr (P == 16) tmem_wait16(v);
    if constexpr (P == 8) tmem_wait8(v);
    if constexpr (P == 4) tmem_wait4(v);
    if constexpr (P == 2) tmem_wait2(v);
    if constexpr (KT > P) tmem_wait<KT - P>(v + P);
}

} // namespace msv
