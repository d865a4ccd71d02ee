// msv_wave_kernels.cuh -- ONE sequence per launch, as fast as the row dependency allows: the latency kernel behind
// msv_cuda_score_sequence, i.e. MSV_HMM::parallel_run_on_sequence (reference algorithms/MSV_HMM.cpp:269-430, where one
// residue costs 13 kernel launches) and the benchmark_MSV_1400 workload (reference benchmark_MSV_1400.cpp:5-16).
//
// A sequence is serial in the row direction, but with the speculative row B = N + move (see msv_scan_warp_kernel: B does
// not depend on E while J <= N, verified at the end) the ONLY dependency between rows is M[i][k] <- M[i-1][k-1].  So the
// model columns are cut into blocks of 32*K columns, one WARP per block, and the warps form a systolic chain: warp g
// runs row i as soon as warp g-1 has published the last column of its row i-1.  No barrier, no E exchange, nobody waits
// for a whole row: every warp streams through all rows at its own pace, one chunk of four rows (one residue word) behind
// its left neighbour.
//   * the chain spans a thread-block CLUSTER (four warps per CTA, one per SM sub-partition, up to 8 CTAs), so that a
//     1400-column model runs on 8..22 warps of 2..6 SMs instead of sharing one SM's issue slots;
//   * hand-over between neighbours goes through a small ring in the CONSUMER's shared memory (distributed shared memory
//     across CTAs): once per four rows the producer's lane 31 stores its four boundary values with one 16-byte remote
//     store into a slot that holds NaNs until then -- the data is its own "ready" flag (a score is never NaN), so there is
//     no flag, no fence and no mbarrier on the row-to-row path (measured: a release store at cluster scope costs a
//     MEMBAR.ALL.GPU per chunk on sm_100a, an mbarrier try_wait ~100 clocks).  The consumer reads slot c+1 while it
//     computes chunk c, re-reads it only if it was still empty, empties it again and reports its progress back with a plain
//     remote store that the producer looks at only when the ring might be full;
//   * the residues are turned into emission-row offsets once, into shared memory, by the validation pass that every CTA
//     runs while its table slice is still in flight -- the row loop never touches global or constant memory;
//   * every lane carries its own share of J (j = max(j + loop, E_lane + tEJ)); N and B are running sums every warp keeps
//     redundantly; at the end the shares are combined with one atomic max per warp, the last warp to finish writes
//     score, verdict (speculation held / residue code invalid) and the call's tag straight into pinned host memory,
//     which the host polls -- no memcpy, no stream synchronisation on the way back;
//   * short sequences travel INSIDE the kernel parameters (constant bank), so a call is a single launch.
// When the verification fails (J overtook N: the sequence contains a real hit) the host re-scores the sequence with the
// exact four-warp kernel.  Same bits in both cases: every cell does the reference's add on the reference's operands.
#pragma once

#include <type_traits>

#include "msv_device.cuh"

namespace msv {

constexpr int kWaveWarpsPerCta = 4;
constexpr int kWaveMaxCtas = 8;
constexpr uint32_t kWaveSlots = 16;           // ring of four-row chunks buffered between two neighbouring warps
constexpr uint32_t kWaveInlineBytes = 3968;   // longest sequence that travels inside the kernel parameters

struct Wave_result { // pinned, mapped host memory
    float score;
    uint32_t status;  // bit 0: speculation failed (rescore exactly), bit 1: a residue code >= 20
    uint32_t tag;     // written last: the call this result belongs to
    uint32_t pad;
};

struct Wave_accumulator { // device memory, all zero between calls
    unsigned int best;    // max over warps of J, as an order-preserving unsigned
    unsigned int status;
    unsigned int finished; // warps done
    unsigned int pad;
};

struct Wave_params {
    const float* table;         // [warp][residue][pair][lane][2], -inf beyond the model
    const uint8_t* residues;    // device memory, 4-byte aligned, readable to a multiple of 4 (when not inline)
    Wave_accumulator* accumulator;
    Wave_result* result;        // device-visible address of the pinned result slot
    uint32_t length;
    uint32_t warps;             // warps in the chain (32*K*warps >= model columns)
    uint32_t tag;
    float tr_B_Mk, tr_E_J, loop, move;
};

struct Wave_inline_residues {
    uint32_t words[kWaveInlineBytes / 4];
};
struct Wave_no_residues { // stands in for the inline block in the kernel that reads the sequence from device memory
    uint32_t words[1];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t shared_address, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(shared_address), "r"(cta));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of `local` (an object in my shared memory) as it appears in CTA `cta` of the cluster
template <typename T> __device__ __forceinline__ T* map_generic_to_cta(T* local, uint32_t cta) {
    uint64_t in = reinterpret_cast<uint64_t>(local), out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(in), "r"(cta));
    return reinterpret_cast<T*>(out);
}
// plain (weak) stores through a generic address into a neighbour's shared memory: fire and forget
__device__ __forceinline__ void store_remote_v4(float4* generic, float a, float b, float c, float d) {
    asm volatile("st.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(generic), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void store_remote_u32(uint32_t* generic, uint32_t v) {
    asm volatile("st.u32 [%0], %1;" ::"l"(generic), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t shared_address) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_address) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128_volatile_u32(uint32_t shared_address) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(shared_address) : "memory");
    return v;
}
__device__ __forceinline__ void sts128_u32(uint32_t shared_address, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(shared_address), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// float -> unsigned that orders like the float (for atomicMax); 0 is below every float, including -inf
__device__ __forceinline__ unsigned int ordered_bits(float v) {
    const unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

constexpr uint32_t kWaveEmpty = 0x7fffffffu; // a NaN: what a ring slot holds until my neighbour's values arrive (a score is never NaN)

template <int K, bool INLINE>
__global__ void __launch_bounds__(kWaveWarpsPerCta * 32, 1)
msv_wave_kernel(const __grid_constant__ Wave_params p,
                const __grid_constant__ std::conditional_t<INLINE, Wave_inline_residues, Wave_no_residues> inl) {
    static_assert(K % 2 == 0 && K >= 2 && K <= 16, "columns per lane");
    constexpr uint32_t ROW_BYTES = K * 128;              // one residue's emissions for one warp: K/2 pairs x 32 lanes x 8 B
    constexpr uint32_t WARP_TABLE_BYTES = kAlphabet * ROW_BYTES;
    constexpr uint32_t TABLE_BYTES = kWaveWarpsPerCta * WARP_TABLE_BYTES;

    // dynamic shared memory: my four warps' slices of the emission table, then one word per row of the sequence: the
    // offset of that residue's emission row (x * ROW_BYTES), padded to whole chunks of four rows
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // what my left neighbour sends me: slot c % kWaveSlots holds the last column of its rows 4c+1 .. 4c+4 (kWaveEmpty until then)
    __shared__ __align__(16) uint32_t ring[kWaveWarpsPerCta][kWaveSlots][4];
    __shared__ uint32_t taken[kWaveWarpsPerCta]; // chunks my RIGHT neighbour has consumed (and emptied again)
    __shared__ __align__(8) uint64_t table_ready;
    __shared__ uint32_t bad_code;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t cta = cluster_ctarank();
    const uint32_t g = cta * kWaveWarpsPerCta + warp; // position in the chain
    const bool in_chain = g < p.warps;
    const float NEG_INF = __int_as_float(0xff800000);

    // ---- set-up: my slice of the table by TMA; meanwhile validate the residues and turn them into row offsets ----
    if (threadIdx.x == 0) {
        mbarrier_init(&table_ready, 1);
        bad_code = 0;
    }
    for (uint32_t i = lane; i < kWaveSlots * 4; i += 32) (&ring[warp][0][0])[i] = kWaveEmpty;
    if (lane == 0) taken[warp] = 0;
    __syncthreads();
    const uint32_t my_warps = min(static_cast<uint32_t>(kWaveWarpsPerCta), p.warps > cta * kWaveWarpsPerCta ? p.warps - cta * kWaveWarpsPerCta : 0u);
    if (threadIdx.x == 0 && my_warps > 0) {
        const uint32_t bytes = my_warps * WARP_TABLE_BYTES;
        mbarrier_expect_tx(&table_ready, bytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.table) + static_cast<size_t>(cta) * TABLE_BYTES;
        for (uint32_t at = 0; at < bytes; at += 32768) tma_bulk_load(smem_raw + at, src + at, min(32768u, bytes - at), &table_ready);
    }
    const uint32_t words = (p.length + 3) / 4;
    uint4* row_offsets = reinterpret_cast<uint4*>(smem_raw + TABLE_BYTES);
    {
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(p.residues);
        uint32_t bad = 0;
        for (uint32_t c = threadIdx.x; c < words; c += blockDim.x) {
            uint32_t w;
            if constexpr (INLINE) w = inl.words[c];
            else w = __ldg(gw + c);
            if (c == words - 1 && (p.length & 3u)) w &= (1u << (8u * (p.length & 3u))) - 1u; // bytes past the end do not count
            bad |= (((w & 0x7f7f7f7fu) + 0x6c6c6c6cu) | w) & 0x80808080u;                   // some byte >= 20
            row_offsets[c] = make_uint4((w & 0xffu) * ROW_BYTES, ((w >> 8) & 0xffu) * ROW_BYTES, ((w >> 16) & 0xffu) * ROW_BYTES,
                                        (w >> 24) * ROW_BYTES);
        }
        if (bad) bad_code = 1;
    }
    // every CTA's mailboxes must exist before a neighbour writes into them (and every CTA reaches the same verdict on the
    // residue codes, so either all of them scan or none does); also makes row_offsets visible to all my warps
    cluster_sync_all();
    const bool invalid = bad_code != 0;

    float J = NEG_INF, N = 0.0f;
    if (my_warps > 0) mbarrier_wait(&table_ready, 0); // (unconditionally: nobody leaves while a bulk copy into its shared memory is in flight)
    if (in_chain && !invalid) {
        const uint32_t tab = smem_u32(smem_raw) + warp * WARP_TABLE_BYTES + lane * 8;
        const float tBMk = p.tr_B_Mk, tEJ = p.tr_E_J, loop = p.loop, move = p.move;
        const bool has_left = g > 0, has_right = g + 1 < p.warps;
        // my mailboxes (local) and my neighbours' (possibly in another CTA of the cluster)
        const uint32_t my_ring = smem_u32(&ring[warp][0][0]), my_taken = smem_u32(&taken[warp]);
        const uint32_t right_cta = (g + 1) / kWaveWarpsPerCta, right_warp = (g + 1) % kWaveWarpsPerCta;
        const uint32_t left_cta = has_left ? (g - 1) / kWaveWarpsPerCta : 0u, left_warp = has_left ? (g - 1) % kWaveWarpsPerCta : 0u;
        float4* const right_ring = reinterpret_cast<float4*>(map_generic_to_cta(&ring[right_warp][0][0], right_cta));
        uint32_t* const left_taken = map_generic_to_cta(&taken[left_warp], left_cta);

        float m[K];
#pragma unroll
        for (int j = 0; j < K; ++j) m[j] = NEG_INF; // MSV_HMM.cpp:86
        float B = move;                             // MSV_HMM.cpp:96-97 (N = 0 above)

        // Emissions are fetched one chunk (four rows) AHEAD of their use: with one warp per scheduler nothing else hides the
        // shared-memory latency, and the in-order issue would stall on every row otherwise.
        struct Chunk_emissions {
            float2 e[4][K / 2];
        };
        const auto fetch = [&](Chunk_emissions& into, const uint4 o) {
            const uint32_t at[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < K / 2; ++q)
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(into.e[r][q].x), "=f"(into.e[r][q].y) : "r"(tab + at[r] + q * 256));
        };
        // one row: `from_left` = last column of my left neighbour's previous row (-inf for the first block: the dummy column M0)
        const auto row = [&](const float2 (&ev)[K / 2], const float from_left) {
            const float bt = B + tBMk; // MSV_HMM.cpp:103, B -> M_k entry
            const float up = __shfl_up_sync(0xffffffffu, m[K - 1], 1);
            const float left = lane == 0 ? from_left : up;
            float e = NEG_INF;
#pragma unroll
            for (int q = K / 2 - 1; q >= 0; --q) { // highest column first: every cell reads its not-yet-overwritten left neighbour
                const int j = 2 * q;
                m[j + 1] = ev[q].y + fmaxf(m[j], bt);
                m[j] = ev[q].x + fmaxf(j ? m[j > 0 ? j - 1 : 0] : left, bt);
                e = fmaxf(fmaxf(e, m[j + 1]), m[j]); // MSV_HMM.cpp:104
            }
            J = fmaxf(J + loop, e + tEJ); // this lane's share of J (MSV_HMM.cpp:107); combined once, after the last row
            N = N + loop;                 // MSV_HMM.cpp:109
            B = N + move;                 // MSV_HMM.cpp:110 while J <= N -- verified after the last row
        };

        // Consumer side.  Chunk c (rows 4c+1 .. 4c+4) needs my neighbour's rows 4c .. 4c+3: the last value of its chunk c-1 and
        // the first three of its chunk c.  Slot c+1 is read while chunk c is computed; if it is still (partly) empty then, it
        // is read again when it is needed.
        const uint32_t chunks = p.length / 4, rest = p.length & 3u;
        const uint32_t incoming_chunks = has_left ? chunks + (rest ? 1u : 0u) : 0u; // chunks my neighbour sends
        const auto slot_of = [&](uint32_t c) { return my_ring + (c % kWaveSlots) * 16u; };
        const auto is_empty = [](const uint4 v) { return v.x == kWaveEmpty || v.y == kWaveEmpty || v.z == kWaveEmpty || v.w == kWaveEmpty; };
        uint4 incoming = make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u);
        if (incoming_chunks > 0) incoming = lds128_volatile_u32(slot_of(0));
        float previous_last = NEG_INF; // "row 0": nothing enters from the left before the first row
        uint32_t taken_seen = 0;
        const uint32_t offsets_at = smem_u32(row_offsets);
        const auto offsets_of = [&](uint32_t c) { return lds128_volatile_u32(offsets_at + c * 16u); };
        Chunk_emissions ahead;
        if (words) fetch(ahead, offsets_of(0));
        uint4 offsets_next = words > 1 ? offsets_of(1) : make_uint4(0, 0, 0, 0);

#pragma unroll 2
        for (uint32_t c = 0; c < chunks; ++c) {
            const Chunk_emissions now = ahead;
            if (c + 1 < words) fetch(ahead, offsets_next);
            if (c + 2 < words) offsets_next = offsets_of(c + 2);
            uint4 cur = incoming;
            if (has_left) {
                while (is_empty(cur)) cur = lds128_volatile_u32(slot_of(c));
                if ((c & 3u) == 3u && lane < 4) { // chunks c-3 .. c are in registers: free their slots, then tell my neighbour
                    sts128_u32(slot_of(c - lane), kWaveEmpty, kWaveEmpty, kWaveEmpty, kWaveEmpty);
                    __syncwarp(0xfu);
                    if (lane == 0) store_remote_u32(left_taken, c + 1);
                }
                if (c + 1 < incoming_chunks) incoming = lds128_volatile_u32(slot_of(c + 1));
            }
            // back-pressure: my consumer must have emptied the slot I am about to fill
            if (has_right) {
                while (c >= taken_seen + kWaveSlots) taken_seen = lds_volatile_u32(my_taken);
            }
            row(now.e[0], previous_last);
            const float b0 = m[K - 1];
            row(now.e[1], __uint_as_float(cur.x));
            const float b1 = m[K - 1];
            row(now.e[2], __uint_as_float(cur.y));
            const float b2 = m[K - 1];
            row(now.e[3], __uint_as_float(cur.z));
            previous_last = __uint_as_float(cur.w);
            if (has_right && lane == 31) store_remote_v4(right_ring + (c % kWaveSlots), b0, b1, b2, m[K - 1]);
        }
        // ---- the last one to three rows ----
        if (rest) {
            const uint32_t c = chunks;
            uint4 cur = incoming;
            if (has_left) {
                while (is_empty(cur)) cur = lds128_volatile_u32(slot_of(c));
            }
            if (has_right) {
                while (c >= taken_seen + kWaveSlots) taken_seen = lds_volatile_u32(my_taken);
            }
            const float from_left[3] = {previous_last, __uint_as_float(cur.x), __uint_as_float(cur.y)};
            float b[3] = {NEG_INF, NEG_INF, NEG_INF};
#pragma unroll
            for (uint32_t r = 0; r < 3; ++r) {
                if (r < rest) {
                    row(ahead.e[r], from_left[r]);
                    b[r] = m[K - 1];
                }
            }
            if (has_right && lane == 31) store_remote_v4(right_ring + (c % kWaveSlots), b[0], b[1], b[2], NEG_INF);
        }
    }

    // ---- combine: max of the lanes' shares of J, did the speculation hold, and hand the result to the host ----
    if (in_chain) {
        const bool suspect = __any_sync(0xffffffffu, J >= N); // J overtook N at some row: B was not N + move there
        float Jw;
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(Jw) : "f"(J));
        if (lane == 0) {
            atomicMax(&p.accumulator->best, ordered_bits(Jw));
            const unsigned int status = (suspect && !invalid ? 1u : 0u) | (invalid ? 2u : 0u);
            if (status) atomicOr(&p.accumulator->status, status);
            __threadfence();
            if (atomicAdd(&p.accumulator->finished, 1u) == p.warps - 1) { // every warp's contribution is in
                __threadfence();
                const float best = from_ordered_bits(atomicExch(&p.accumulator->best, 0u));
                const unsigned int verdict = atomicExch(&p.accumulator->status, 0u);
                atomicExch(&p.accumulator->finished, 0u); // all zero again for the next call
                volatile Wave_result* out = p.result;
                out->score = best + p.move; // MSV_HMM.cpp:112 (C == J: tr_E_C == tr_E_J)
                out->status = verdict;
                __threadfence_system();
                out->tag = p.tag;
            }
        }
    }
    // nobody leaves while a neighbour may still store into its shared memory
    cluster_sync_all();
}

// =====================================================================================================================
// The same single sequence with NO communication at all: one THREAD per diagonal phase.
//
// With the speculative row the only dependency is M[i][k] <- M[i-1][k-1]: cells on different diagonals never meet.  Worker
// w (one thread) walks down the rows 1 .. L and, at row i, owns column (i - 1 - w) mod P of the P = LENG + 1 columns
// 0 .. LENG -- i.e. it follows one diagonal to the last column, wraps to column 0 and follows the next one.  Column 0 is the
// reference's dummy column M0, whose emission is -inf (MSV_HMM.cpp:38-45 with the zero-filled row 0 of match_emissions,
// Profile_HMM.cpp:110-111): passing through it resets the worker's value to -inf exactly as M[i][0] = -inf does in the
// reference, so there is no wrap logic in the arithmetic.  P workers cover every column of every row.  Each worker keeps
// its own share of J and its own copy of N / B (running sums, the same in every thread); at the end the shares are combined
// like in the chain kernel above.  Per row and thread: one LDS (all lanes of a warp read consecutive columns of the same
// residue's row: conflict-free, and the table is the reference's own [residue][column] layout), two maxima, six adds.
// The critical path is the row count times one max + one add: ~20 us for 3500 rows, whatever the model length, on
// ceil(P / 128) SMs with one warp per scheduler.  Needs the whole table (80 bytes per column) plus 4 bytes per row in one
// SM's shared memory: models up to ~2650 columns; longer ones use the chain kernel.
// =====================================================================================================================
struct Diag_params {
    const float* table;         // [residue][P + 4]: the reference's table, every row extended by its own first four columns
    const uint8_t* residues;    // device memory (when not inline)
    Wave_accumulator* accumulator;
    Wave_result* result;
    uint32_t length;
    uint32_t period;            // P = model_length = LENG + 1
    uint32_t tag;
    float tr_B_Mk, tr_E_J, loop, move;
};

template <bool INLINE>
__global__ void __launch_bounds__(kWaveWarpsPerCta * 32, 1)
msv_diag_kernel(const __grid_constant__ Diag_params p,
                const __grid_constant__ std::conditional_t<INLINE, Wave_inline_residues, Wave_no_residues> inl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t table_ready;
    __shared__ uint32_t bad_code;

    const int lane = threadIdx.x & 31;
    const uint32_t row_bytes = (p.period + 4u) * 4u;    // one residue's emissions, with the four wrapped columns
    const uint32_t table_bytes = kAlphabet * row_bytes; // a multiple of 16
    const float NEG_INF = __int_as_float(0xff800000);

    if (threadIdx.x == 0) {
        mbarrier_init(&table_ready, 1);
        bad_code = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbarrier_expect_tx(&table_ready, table_bytes);
        for (uint32_t at = 0; at < table_bytes; at += 32768)
            tma_bulk_load(smem_raw + at, reinterpret_cast<const unsigned char*>(p.table) + at, min(32768u, table_bytes - at), &table_ready);
    }
    // while the table is in flight: validate the residues and turn them into row offsets (one word per row, chunks of four)
    const uint32_t words = (p.length + 3) / 4;
    uint4* row_offsets = reinterpret_cast<uint4*>(smem_raw + (table_bytes + 15u) / 16u * 16u);
    {
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(p.residues);
        uint32_t bad = 0;
        for (uint32_t c = threadIdx.x; c < words; c += blockDim.x) {
            uint32_t w;
            if constexpr (INLINE) w = inl.words[c];
            else w = __ldg(gw + c);
            if (c == words - 1 && (p.length & 3u)) w &= (1u << (8u * (p.length & 3u))) - 1u;
            bad |= (((w & 0x7f7f7f7fu) + 0x6c6c6c6cu) | w) & 0x80808080u; // some byte >= 20
            row_offsets[c] = make_uint4((w & 0xffu) * row_bytes, ((w >> 8) & 0xffu) * row_bytes, ((w >> 16) & 0xffu) * row_bytes, (w >> 24) * row_bytes);
        }
        if (bad) bad_code = 1;
    }
    __syncthreads();
    const bool invalid = bad_code != 0;

    float J = NEG_INF, N = 0.0f;
    mbarrier_wait(&table_ready, 0); // (also when the residues are invalid: nobody leaves while a bulk copy into its shared memory is in flight)
    if (!invalid) {
        const float tBMk = p.tr_B_Mk, tEJ = p.tr_E_J, loop = p.loop, move = p.move;
        const uint32_t period_bytes = p.period * 4u;
        // worker w is on column (i - 1 - w) mod P at row i: at the first row on column (P - w mod P) mod P
        const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) % p.period; // (surplus lanes of the last warp repeat early workers)
        uint32_t column_bytes = w ? (p.period - w) * 4u : 0u;
        const uint32_t tab = smem_u32(smem_raw);
        const uint32_t offsets_at = smem_u32(row_offsets);
        float y = NEG_INF, B = move; // MSV_HMM.cpp:86, 96-97

        const auto fetch = [&](float (&e)[4], const uint4 o, const uint32_t at) { // the four emissions of a chunk: columns at .. at + 3
            const uint32_t base = tab + at;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(e[0]) : "r"(base + o.x));
            asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(e[1]) : "r"(base + o.y));
            asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(e[2]) : "r"(base + o.z));
            asm volatile("ld.shared.f32 %0, [%1+12];" : "=f"(e[3]) : "r"(base + o.w));
        };
        const auto row = [&](const float e) {
            const float bt = B + tBMk;    // MSV_HMM.cpp:103, B -> M_k entry
            y = e + fmaxf(y, bt);         // MSV_HMM.cpp:103: M[i][k] from M[i-1][k-1]; e = -inf on the dummy column resets the diagonal
            J = fmaxf(J + loop, y + tEJ); // this thread's share of J (MSV_HMM.cpp:104,107)
            N = N + loop;                 // MSV_HMM.cpp:109
            B = N + move;                 // MSV_HMM.cpp:110 while J <= N -- verified after the last row
        };
        const auto advance = [&](uint32_t at) { // four columns on, modulo P
            at += 16u;
            return at >= period_bytes ? at - period_bytes : at;
        };
        const auto offsets_of = [&](uint32_t c) { return lds128_volatile_u32(offsets_at + c * 16u); };

        const uint32_t chunks = p.length / 4, rest = p.length & 3u;
        float ahead[4] = {NEG_INF, NEG_INF, NEG_INF, NEG_INF};
        if (words) fetch(ahead, offsets_of(0), column_bytes);
        column_bytes = advance(column_bytes);
        uint4 offsets_next = words > 1 ? offsets_of(1) : make_uint4(0, 0, 0, 0);
#pragma unroll 2
        for (uint32_t c = 0; c < chunks; ++c) {
            const float e0 = ahead[0], e1 = ahead[1], e2 = ahead[2], e3 = ahead[3];
            if (c + 1 < words) fetch(ahead, offsets_next, column_bytes); // emissions one chunk ahead of their use
            column_bytes = advance(column_bytes);
            if (c + 2 < words) offsets_next = offsets_of(c + 2);
            row(e0);
            row(e1);
            row(e2);
            row(e3);
        }
#pragma unroll
        for (uint32_t r = 0; r < 3; ++r)
            if (r < rest) row(ahead[r]);
    }

    // ---- combine: max of the threads' shares of J, did the speculation hold, and hand the result to the host ----
    const bool suspect = __any_sync(0xffffffffu, J >= N);
    float Jw;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(Jw) : "f"(J));
    if (lane == 0) {
        atomicMax(&p.accumulator->best, ordered_bits(Jw));
        const unsigned int status = (suspect && !invalid ? 1u : 0u) | (invalid ? 2u : 0u);
        if (status) atomicOr(&p.accumulator->status, status);
        __threadfence();
        if (atomicAdd(&p.accumulator->finished, 1u) == gridDim.x * (blockDim.x / 32) - 1) { // every warp's contribution is in
            __threadfence();
            const float best = from_ordered_bits(atomicExch(&p.accumulator->best, 0u));
            const unsigned int verdict = atomicExch(&p.accumulator->status, 0u);
            atomicExch(&p.accumulator->finished, 0u);
            volatile Wave_result* out = p.result;
            out->score = best + p.move; // MSV_HMM.cpp:112 (C == J: tr_E_C == tr_E_J)
            out->status = verdict;
            __threadfence_system();
            out->tag = p.tag;
        }
    }
}

} // namespace msv
