// msv_kernels.cuh -- hand-written sm_100a kernels of the MSV scan.
//
// What the kernels compute (per sequence, fp32, only `+` and `max`): the recurrence of the reference's
// MSV_HMM::run_on_sequence (reference algorithms/MSV_HMM.cpp:96-112), i.e. the fusion of its six OpenCL kernels
// (algorithms/MSV_kernels.cl:1-65: init_dp, init_N_B, M_states_handler, copy_M, reduction_step, E_J_C_N_B_handler)
// and of the host loop that launches them 13 times per residue (MSV_HMM.cpp:382-423) into ONE launch per database.
//
// Three kernel families share one design (B200-first, not a translation):
//   msv_scan_warp_kernel        one warp per sequence; emissions from shared memory AND tensor memory   -- the hot kernel
//   msv_scan_kernel             G = 8/16/32 lanes per sequence, 32/G sequences per warp                  -- short models
//   msv_scan_group_spec_kernel  the same with speculative rows, failed sequences repeated exactly in place -- short models (default)
//   msv_scan_quad_kernel        four warps per sequence; table distributed over shared + tensor memory   -- few/long sequences,
//                                                                                                          models > 2815 columns
// Common to all:
//   * the lanes that own a sequence keep its DP row in REGISTERS (m[K]: K consecutive model columns per lane); the row
//     never touches memory;
//   * the k-1 dependency is satisfied inside a lane by updating m[] from the highest column downwards (every cell
//     reads the not-yet-overwritten left neighbour), and across lanes by ONE shuffle per row; the shuffle ROTATES, and
//     because the last column of the last lane is -inf padding, lane 0 receives the -inf of the dummy column M0 for free;
//   * E = max_k M is a per-lane FMNMX3 chain followed by a cross-lane max (CREDUX.MAX.F32 for whole warps, xor-shuffles
//     for lane groups); the special states N/J/B are carried redundantly by every lane; C is J when tr_E_C == tr_E_J;
//   * the emission table ([residue][column]) is staged ONCE per CTA: the shared-memory part by bulk-async (TMA) copies
//     completing on an mbarrier, laid out [residue][quad][lane][4] so that every access is a conflict-free LDS.128; the
//     tensor-memory part by tcgen05.st, read back with tcgen05.ld; columns beyond the model are -inf;
//   * CTAs are persistent (one per SM); slots pull sequences longest-first from a global atomic queue;
//   * residues stream from HBM as aligned 32-bit words (4 residues), prefetched one word ahead.
//   * speculative rows (warp and shortest-model kernels): B = max(N, J) + move is N + move while J <= N, so the row needs no
//     cross-lane E; one vote per sequence verifies it, failures are scanned again with the exact row (argument at
//     msv_scan_warp_kernel).
// Exactness: every cell performs the same fp32 add on the same operands as the reference; max is exact and
// order-independent for the non-NaN values that can occur (-inf and finite numbers only), so scores are
// bit-identical to the reference for any order of the E reduction.
#pragma once

#include <type_traits>

#include "msv_device.cuh"

namespace msv {

// ---- the scan, lane-group family: G = 8, 16 or 32 lanes per sequence, whole table in shared memory ------------------
// A warp scans 32/G sequences at once; every group pulls its own sequences from the queue.  Rows are executed in
// warp-uniform chunks (the minimum over the groups of their remaining rows), so the row loop itself never diverges;
// retiring a sequence and fetching the next one happens between chunks.  Per-row bookkeeping is shared by the 32/G
// sequences of the warp, which is what makes this family the faster one for short models.
// The host guarantees G*K > model columns: the last column of a group's last lane is -inf padding, so the rotating
// shuffle hands lane 0 of each group the -inf of the dummy column M0 (same trick as in the warp kernel below).
// G = 4: a quarter-warp (the unit in which an LDS.128 is served) holds TWO groups, whose residues differ; the table is
// therefore stored as two interleaved copies ([residue][quad][copy][lane][4]: copy 0 in banks 0-15, copy 1 in banks 16-31),
// even groups read copy 0 and odd groups copy 1, and every access stays conflict-free whatever the residues are.
template <int G, int K, int THREADS, bool CJ_SAME>
__global__ void __launch_bounds__(THREADS, 1) msv_scan_kernel(const Scan_params p) {
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "lanes per sequence");
    static_assert(K % 2 == 0 && (K % 4 == 0 || G <= 8) && K >= 4 && K <= kMaxColumnsPerLane, "columns per lane");
    constexpr int LANES_PER_QUAD_ROW = G < 8 ? 8 : G;
    constexpr uint32_t QUAD_BYTES = LANES_PER_QUAD_ROW * 16;
    // K % 4 == 2 (G = 4, 8: less padding, e.g. 4 x 26 = 104 slots for LENG 100 instead of 112): the two highest columns of
    // every lane are a PAIR stored behind the quads of the residue row, 128 bytes = 16 lane slots x 8 bytes, read with one
    // LDS.64.  An LDS.64 is served per half-warp, which holds 16 / G groups with different residues, so the pair is stored
    // 16 / G times side by side and lane l reads slot l % 16: every half-warp touches each bank once, whatever the residues.
    constexpr bool PAIR = K % 4 == 2;
    constexpr uint32_t PAIR_OFFSET = (K / 4) * QUAD_BYTES;
    constexpr uint32_t ROW_BYTES = PAIR_OFFSET + (PAIR ? 128u : 0u); // bytes per residue row of the table
    constexpr uint32_t COPY_CHUNK = 32768;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t table_ready;

    // never index the table with an unvalidated residue code: the validation kernel precedes this launch on the stream
    if (p.first_bad != nullptr && *p.first_bad != ~0ull) return;

    // ---- stage the emission table: one thread programs the TMA unit, everybody waits on the mbarrier ----
    if (threadIdx.x == 0) mbarrier_init(&table_ready, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbarrier_expect_tx(&table_ready, p.table_bytes);
        for (uint32_t at = 0; at < p.table_bytes; at += COPY_CHUNK) {
            const uint32_t bytes = min(COPY_CHUNK, p.table_bytes - at);
            tma_bulk_load(smem_raw + at, reinterpret_cast<const unsigned char*>(p.table) + at, bytes, &table_ready);
        }
    }
    mbarrier_wait(&table_ready, 0);
    const bool fast_cta = blockIdx.x < p.fast_ctas;
    if (fast_cta && threadIdx.x >= p.fast_threads) return; // whole warps; nobody meets at a barrier after this point

    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
    const int left_lane = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1)); // rotate inside the group
    const uint32_t tab_lane = smem_u32(smem_raw) + (lane & (LANES_PER_QUAD_ROW - 1)) * 16;
    [[maybe_unused]] const uint32_t pair_lane = smem_u32(smem_raw) + PAIR_OFFSET + (lane & 15) * 8;
    const float NEG_INF = __int_as_float(0xff800000);
    const float tBMk = p.tr_B_Mk, tEC = p.tr_E_C, tEJ = p.tr_E_J;

    float m[K];
#pragma unroll
    for (int j = 0; j < K; ++j) m[j] = NEG_INF;
    float J = NEG_INF, C = NEG_INF, N = 0.0f, B = NEG_INF, loop = 0.0f, move = 0.0f;

    uint32_t remaining = 0, idx = 0;
    bool active = false, done = false;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.residues);
    uint32_t buf = 0, nextw = 0, have = 0x7fffffffu;

    for (;;) {
        // ---- retire finished sequences, pull new ones (group-uniform control flow) ----
        while (remaining == 0 && !done) {
            if (active) {
                if (gl == 0) store_score(p, idx, (CJ_SAME ? J : C) + move); // MSV_HMM.cpp:112
                active = false;
            }
            uint32_t ticket = 0;
            if (gl == 0) ticket = next_ticket(p, fast_cta);
            ticket = __shfl_sync(gmask, ticket, 0, G);
            if (ticket >= p.n) {
                done = true;
                buf = 0;
                nextw = 0;
                have = 0x7fffffffu;
                break;
            }
            idx = __ldg(p.order + ticket);
            const uint64_t begin = __ldg(p.offsets + idx);
            const uint32_t len = static_cast<uint32_t>(__ldg(p.offsets + idx + 1) - begin);
            const float2 tr = __ldg(p.length_tr + len);
            loop = tr.x;
            move = tr.y;
#pragma unroll
            for (int j = 0; j < K; ++j) m[j] = NEG_INF; // MSV_HMM.cpp:86
            J = NEG_INF;
            C = NEG_INF;
            N = 0.0f;  // MSV_HMM.cpp:96
            B = move;  // MSV_HMM.cpp:97
            const uint32_t mis = static_cast<uint32_t>(begin) & 3u;
            wp = reinterpret_cast<const uint32_t*>(p.residues + (begin - mis));
            buf = __ldg(wp) >> (8u * mis);
            nextw = __ldg(wp + 1);
            wp += 2;
            have = 4u - mis;
            remaining = len;
            active = true;
        }

        // ---- rows that every group of this warp can run without anyone finishing: warp-uniform trip count ----
        const uint32_t mine = done ? 0xffffffffu : remaining;
        const uint32_t steps = (G == 32) ? mine : __reduce_min_sync(0xffffffffu, mine);
        if (steps == 0xffffffffu) break;
        if (!done) remaining -= steps;

#pragma unroll 1
        for (uint32_t t = 0; t < steps; ++t) {
            const uint32_t erow = tab_lane + (buf & 0xffu) * ROW_BYTES;
            [[maybe_unused]] const uint32_t prow = pair_lane + (buf & 0xffu) * ROW_BYTES;
            buf >>= 8;
            if (--have == 0) {
                buf = nextw;
                have = 4;
                if (!done) nextw = __ldg(wp);
                ++wp;
            }
            const float bt = B + tBMk; // MSV_HMM.cpp:103, the B -> M_k entry
            const float left = __shfl_sync(0xffffffffu, m[K - 1], left_lane);

            float e = NEG_INF;
            if constexpr (PAIR) { // the two highest columns first (every cell reads its not-yet-overwritten left neighbour)
                const float2 ev = lds64(prow);
                m[K - 1] = ev.y + fmaxf(m[K - 2], bt);
                m[K - 2] = ev.x + fmaxf(m[K - 3], bt);
                e = fmaxf(m[K - 1], m[K - 2]);
            }
#pragma unroll
            for (int q = K / 4 - 1; q >= 0; --q) {
                const float4 ev = lds128(erow + q * QUAD_BYTES);
                const int j = 4 * q;
                m[j + 3] = ev.w + fmaxf(m[j + 2], bt);
                m[j + 2] = ev.z + fmaxf(m[j + 1], bt);
                m[j + 1] = ev.y + fmaxf(m[j], bt);
                m[j] = ev.x + fmaxf(q ? m[j > 0 ? j - 1 : 0] : left, bt);
                e = fmaxf(fmaxf(e, m[j + 3]), m[j + 2]); // MSV_HMM.cpp:104
                e = fmaxf(fmaxf(e, m[j + 1]), m[j]);
            }
            const float E = group_max<G>(e);

            J = fmaxf(J + loop, E + tEJ);                         // MSV_HMM.cpp:107
            if constexpr (!CJ_SAME) C = fmaxf(C + loop, E + tEC); // MSV_HMM.cpp:108
            N = N + loop;                                         // MSV_HMM.cpp:109
            B = fmaxf(N, J) + move; // MSV_HMM.cpp:110: max(N+move, J+move) == max(N, J)+move exactly (rounding is monotone)
        }
    }
}

// ---- the lane-group scan with SPECULATIVE rows (see msv_scan_warp_kernel for the argument) ----------------------------
// For short models the per-row bookkeeping of the exact row -- a group-wide max by shuffles, the J and B maxima -- costs as
// much as the cells.  While J <= N, B is N + move and needs none of it: the row is cells + one FMNMX3 chain + this lane's
// share of J.  32/G sequences advance per warp instruction, four rows per residue word.  Verification is one group vote
// when a sequence retires.  A sequence that fails it (it contains a real hit: ~0.2 % of random sequences) is scanned again
// AT ONCE by the same group with the exact row -- sequences are handed out longest first, so a long sequence that has to be
// repeated is repeated early, not as a lonely tail after everybody else has finished (an earlier version collected the
// failures in a list for a second launch, whose run time was the longest failed sequence at single-warp speed).  While any
// group of a warp is in its exact pass the whole warp executes exact rows; they are valid for speculating groups too
// (a lane's share of J only grows towards the true J, and B = max(N, J) + move is N + move for as long as their speculation
// holds).  tr_E_C == tr_E_J only.  G = 4 uses two interleaved copies of the table (see msv_scan_kernel).
template <int G, int K, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) msv_scan_group_spec_kernel(const Scan_params p) {
    static_assert(G == 4 || G == 8 || G == 16, "lanes per sequence");
    static_assert(K % 2 == 0 && (K % 4 == 0 || G <= 8) && K >= 4 && K <= kMaxColumnsPerLane, "columns per lane");
    constexpr int LANES_PER_QUAD_ROW = G < 8 ? 8 : G; // G = 4: two copies of the table side by side, one per neighbouring group
    constexpr uint32_t QUAD_BYTES = LANES_PER_QUAD_ROW * 16;
    constexpr bool PAIR = K % 4 == 2; // the two highest columns of a lane: one LDS.64 from 16 lane slots behind the quads (msv_scan_kernel)
    constexpr uint32_t PAIR_OFFSET = (K / 4) * QUAD_BYTES;
    constexpr uint32_t ROW_BYTES = PAIR_OFFSET + (PAIR ? 128u : 0u);
    constexpr uint32_t COPY_CHUNK = 32768;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t table_ready;

    if (p.first_bad != nullptr && *p.first_bad != ~0ull) return;

    if (threadIdx.x == 0) mbarrier_init(&table_ready, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbarrier_expect_tx(&table_ready, p.table_bytes);
        for (uint32_t at = 0; at < p.table_bytes; at += COPY_CHUNK) {
            const uint32_t bytes = min(COPY_CHUNK, p.table_bytes - at);
            tma_bulk_load(smem_raw + at, reinterpret_cast<const unsigned char*>(p.table) + at, bytes, &table_ready);
        }
    }
    mbarrier_wait(&table_ready, 0);
    const bool fast_cta = blockIdx.x < p.fast_ctas;
    if (fast_cta && threadIdx.x >= p.fast_threads) return; // whole warps; nobody meets at a barrier after this point

    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = ((1u << G) - 1u) << (lane & ~(G - 1));
    const int left_lane = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1)); // rotate inside the group
    const uint32_t tab_lane = smem_u32(smem_raw) + (lane & (LANES_PER_QUAD_ROW - 1)) * 16;
    [[maybe_unused]] const uint32_t pair_lane = smem_u32(smem_raw) + PAIR_OFFSET + (lane & 15) * 8;
    const float NEG_INF = __int_as_float(0xff800000);
    const float tBMk = p.tr_B_Mk, tEJ = p.tr_E_J;

    float m[K];
#pragma unroll
    for (int j = 0; j < K; ++j) m[j] = NEG_INF;
    float J = NEG_INF, N = 0.0f, B = NEG_INF, loop = 0.0f, move = 0.0f; // J: this LANE's share of J (the whole J in an exact pass)

    uint32_t remaining = 0, idx = 0, len = 0;
    uint64_t begin = 0;
    bool active = false, done = false, exact = false;
    // residue window of the group's sequence: the next residue is byte `phase/8` of the 64-bit value (whi:wlo)
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.residues);
    uint32_t wlo = 0, whi = 0, phase = 0;

    const auto cells = [&](const uint32_t x, float& e) {
        const uint32_t erow = tab_lane + x * ROW_BYTES;
        const float bt = B + tBMk;
        const float left = __shfl_sync(0xffffffffu, m[K - 1], left_lane);
        e = NEG_INF;
        if constexpr (PAIR) {
            const float2 ev = lds64(pair_lane + x * ROW_BYTES);
            m[K - 1] = ev.y + fmaxf(m[K - 2], bt);
            m[K - 2] = ev.x + fmaxf(m[K - 3], bt);
            e = fmaxf(m[K - 1], m[K - 2]);
        }
#pragma unroll
        for (int q = K / 4 - 1; q >= 0; --q) {
            const float4 ev = lds128(erow + q * QUAD_BYTES);
            const int j = 4 * q;
            m[j + 3] = ev.w + fmaxf(m[j + 2], bt);
            m[j + 2] = ev.z + fmaxf(m[j + 1], bt);
            m[j + 1] = ev.y + fmaxf(m[j], bt);
            m[j] = ev.x + fmaxf(q ? m[j > 0 ? j - 1 : 0] : left, bt);
            e = fmaxf(fmaxf(e, m[j + 3]), m[j + 2]);
            e = fmaxf(fmaxf(e, m[j + 1]), m[j]);
        }
    };
    const auto row = [&](const uint32_t x) {
        float e;
        cells(x, e);
        J = fmaxf(J + loop, e + tEJ);
        N = N + loop;
        B = N + move; // = max(N, J) + move while J <= N -- verified when the sequence retires
    };
    const auto exact_row = [&](const uint32_t x) {
        float e;
        cells(x, e);
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) e = fmaxf(e, __shfl_xor_sync(0xffffffffu, e, d)); // E of the group (MSV_HMM.cpp:104)
        J = fmaxf(J + loop, e + tEJ);                                                              // MSV_HMM.cpp:107
        N = N + loop;
        B = fmaxf(N, J) + move; // MSV_HMM.cpp:110
    };
    const auto start_sequence = [&] { // (re)start the scan of sequence `idx` at its first row
#pragma unroll
        for (int j = 0; j < K; ++j) m[j] = NEG_INF;
        J = NEG_INF;
        N = 0.0f;
        B = move;
        const uint32_t mis = static_cast<uint32_t>(begin) & 3u;
        wp = reinterpret_cast<const uint32_t*>(p.residues + (begin - mis));
        wlo = __ldg(wp);
        whi = __ldg(wp + 1);
        wp += 2;
        phase = 8u * mis;
        remaining = len;
    };

    for (;;) {
        // ---- retire finished sequences, pull new ones (group-uniform control flow) ----
        while (remaining == 0 && !done) {
            if (active) {
                if (exact) {
                    if (gl == 0) store_score(p, idx, J + move); // MSV_HMM.cpp:112 (every lane holds the whole J)
                    exact = false;
                } else {
                    // if J overtook N at any row, this lane's or a neighbour's share still is >= N now (both decay by + loop)
                    const bool suspect = __any_sync(gmask, J >= N);
                    float best = J;
#pragma unroll
                    for (int d = G / 2; d > 0; d >>= 1) best = fmaxf(best, __shfl_xor_sync(gmask, best, d));
                    if (suspect && len > 0) { // the speculation did not hold: the same sequence again, exactly, right now
                        exact = true;
                        start_sequence();
                        continue;
                    }
                    if (gl == 0) store_score(p, idx, best + move);
                }
                active = false;
            }
            uint32_t ticket = 0;
            if (gl == 0) ticket = next_ticket(p, fast_cta);
            ticket = __shfl_sync(gmask, ticket, 0, G);
            if (ticket >= p.n) {
                done = true;
                wp = reinterpret_cast<const uint32_t*>(p.residues); // finished groups keep executing rows on harmless input
                wlo = whi = phase = 0;
                break;
            }
            idx = __ldg(p.order + ticket);
            begin = __ldg(p.offsets + idx);
            len = static_cast<uint32_t>(__ldg(p.offsets + idx + 1) - begin);
            const float2 tr = __ldg(p.length_tr + len);
            loop = tr.x;
            move = tr.y;
            start_sequence();
            active = true;
        }

        // ---- rows that every group of this warp can run without anyone finishing: warp-uniform trip count ----
        const uint32_t mine = done ? 0xffffffffu : remaining;
        const uint32_t steps = __reduce_min_sync(0xffffffffu, mine);
        if (steps == 0xffffffffu) break;
        if (!done) remaining -= steps;
        const uint32_t advance = done ? 0u : 1u;

        if (__any_sync(0xffffffffu, exact && !done)) { // some group of this warp is in its exact pass: exact rows for everybody
#pragma unroll 1
            for (uint32_t t = steps; t > 0; --t) {
                const uint32_t x = __funnelshift_r(wlo, whi, phase) & 0xffu;
                phase += 8u;
                if (phase == 32u) {
                    phase = 0;
                    wlo = whi;
                    whi = __ldg(wp);
                    wp += advance;
                }
                exact_row(x);
            }
            continue;
        }
#pragma unroll 1
        for (uint32_t t = steps >> 2; t > 0; --t) { // one residue word = four rows
            const uint32_t word = __funnelshift_r(wlo, whi, phase);
            wlo = whi;
            whi = __ldg(wp);
            wp += advance;
            row(word & 0xffu);
            row((word >> 8) & 0xffu);
            row((word >> 16) & 0xffu);
            row(word >> 24);
        }
#pragma unroll 1
        for (uint32_t t = steps & 3u; t > 0; --t) {
            const uint32_t x = __funnelshift_r(wlo, whi, phase) & 0xffu;
            phase += 8u;
            if (phase == 32u) {
                phase = 0;
                wlo = whi;
                whi = __ldg(wp);
                wp += advance;
            }
            row(x);
        }
    }
}

// =====================================================================================================================
// Warp-per-sequence scan with the emission table split between shared memory and TENSOR MEMORY (TMEM).
//
// Why: with one LDS.128 per four cells the generic kernel above saturates the shared-memory pipe (ncu: 88 % of the
// LSU wavefront peak at M = 1400) while the fp32 pipes still have head-room.  Blackwell has a second on-chip memory
// with its own read path: 256 KB of TMEM per SM (512 columns x 128 lanes x 32 bit), read with tcgen05.ld at
// ~300 B/clk/SM (tools/microbench.cu), i.e. more than twice the 128 B/clk of shared memory.  tcgen05.ld.32x32b.xN gives
// thread t of a warp N consecutive columns of TMEM lane 32*(warp%4)+t -- exactly "N emissions of my own model columns".
// So each lane's first KT model columns come from TMEM (column = residue*KT + j, one copy per lane quarter), the other
// KS = K-KT from shared memory.  The TMEM load for a row is issued first, the shared-memory columns are processed while
// it is in flight, tcgen05.wait::ld, then the TMEM columns.  The address must be warp-uniform, which is why this
// variant exists for G == 32 only (all lanes of the warp scan the same residue).
// =====================================================================================================================
// Table in global memory for this kernel:
//   [0, 20*KS*128)            shared-memory part  [residue][quad q][lane][4]   column j = KT + 4q + c
//   [20*KS*128, +20*32*KT*4)  TMEM part           [residue][lane][KT]          column j = 0 .. KT-1
// (model column of lane l, index j:  l*K + j + 1;  -inf beyond the model)
//
// Per-row bookkeeping is trimmed to what the fp32 "ALU" pipe (FMNMX, half rate) can least afford:
//   * the host guarantees that the last column of lane 31 is padding (32*K > model columns), so that column is -inf
//     for ever and a ROTATING shuffle hands lane 0 the -inf of column 0 for free (no select);
//   * when tr_E_C and tr_E_J are the same bits (always, for the reference's nu = 2, MSV_HMM.cpp:49-53) the C and J
//     recurrences are the same function of the same inputs, so C == J and only J is carried (CJ_SAME);
//   * one FMNMX3 chain accumulates E; max(N+move, J+move) is computed as max(N, J)+move (rounding is monotone).
//   * TMEM_AHEAD: the tensor-memory columns are the TOP KT columns of each lane and their emissions for row i+1 are
//     requested while row i is still being computed, so every row starts on operands that are already in registers
//     (instead of waiting for the first LDS) and the shared-memory loads of the row land behind that work.
//   * SPECULATE (with CJ_SAME): B[i] = max(N[i], J[i]) + move equals N[i] + move for as long as J <= N, which is the
//     whole sequence unless it contains a hit worth more than the ~19 nats of entry cost (HMMER 3.1's SSV observation;
//     ~0.2 % of random sequences).  The speculative row therefore takes B = N + move, which needs no E: the warp-wide
//     reduction, the uniform->vector move and the J/B maxima leave the row-to-row critical path.  Each lane carries
//     j = max(j + loop, E_lane + tEJ); rounding is monotone, so max over lanes of j is exactly J as long as the
//     speculation held.  Verification is one vote per SEQUENCE: if J[i] > N[i] at some row then j >= N from that row on
//     in the lane that saw it (j and N decay by the same + loop), so "some lane ends with j >= N" catches every
//     sequence whose B ever differed; those are scanned again with the exact row.  Same bits, always.
#ifndef MSV_SPEC_BLOCK_ROWS
#define MSV_SPEC_BLOCK_ROWS 64
#endif

constexpr uint32_t kSpeculationBlockRows = MSV_SPEC_BLOCK_ROWS; // rows between two checkpoints of the speculative rows (a multiple of 16)

// SPECULATE: 0 = exact rows only; 1 = speculate on the whole sequence, verify once at its end, scan it again exactly when the
// vote fails; 2 = speculate in blocks with checkpoints (below).  Mode 1 is the faster one when hits are rare (the compiler
// schedules its plain loop 2-3 % better: 10.08 vs 9.85 TCUPS at K = 44, profiles/r02/variant_sweep_v3.txt) and loses a whole
// extra pass per hit; mode 2 loses one block per hit.  The host picks per launch from the share of failed speculations it
// observed in the previous scan of the same database (Scan_params::speculation_failures) and from the sequence lengths.
// REPORT: exact rows that also count the sequences which would have failed the speculation (only instantiated for the K that
// have speculating variants to choose from: two more instructions per sequence cost up to 6 % at some K -- 2207.hmm 9.91 ->
// 9.31 TCUPS -- through nothing but a different register allocation).
template <int K, int KT, int THREADS, bool CJ_SAME, bool TMEM_AHEAD = false, int SPECULATE = 0, bool REPORT = false>
__global__ void __launch_bounds__(THREADS, 1) msv_scan_warp_kernel(const Scan_params p) {
    constexpr bool SPEC = SPECULATE != 0 && CJ_SAME;
    constexpr bool CHECKPOINTS = SPECULATE == 2 && CJ_SAME;
    static_assert(KT >= 0 && KT <= 24 && KT % 2 == 0, "TMEM columns per lane");
    constexpr int KS = K - KT;
    static_assert(K % 2 == 0 && KS >= 0 && KS % 4 == 0 && K <= kMaxColumnsPerLane, "columns per lane");
    constexpr uint32_t ROW_BYTES = KS * 32 * 4;
    constexpr uint32_t SMEM_TABLE_BYTES = kAlphabet * ROW_BYTES;
    constexpr uint32_t COPY_CHUNK = 32768;
    constexpr uint32_t TMEM_COLUMNS = 512;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t table_ready;
    __shared__ uint32_t tmem_base_slot;

    int lane = threadIdx.x & 31;
    // opaque to the compiler from here on: at this kernel's register limit it otherwise RE-READS SR_TID (a ~25-clock S2R) inside
    // the row loop to rebuild the lane's table address and shuffle source instead of keeping them in registers
    asm volatile("" : "+r"(lane));
    const int warp = threadIdx.x >> 5;

    // never index the table with an unvalidated residue code: the validation kernel precedes this launch on the stream
    if (p.first_bad != nullptr && *p.first_bad != ~0ull) return;
    // [1] counts the sequences scanned, [0] those whose speculation failed -- or, with exact rows, would have failed: once J
    // exceeds N it stays >= N to the end of the sequence (both grow by the same + loop, and rounding is monotone), which is
    // exactly what the speculating rows vote on.  The host picks the row variant of the next launch from these (launch_scan).
    if constexpr (CJ_SAME && (SPECULATE != 0 || REPORT))
        if (blockIdx.x == 0 && threadIdx.x == 0 && p.speculation_failures) atomicAdd(p.speculation_failures + 1, p.n);

    // ---- stage the shared-memory part with the TMA unit ----
    if (threadIdx.x == 0) mbarrier_init(&table_ready, 1);
    if constexpr (KT > 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                         "n"(TMEM_COLUMNS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if constexpr (KT > 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (KS > 0) {
        if (threadIdx.x == 0) {
            mbarrier_expect_tx(&table_ready, SMEM_TABLE_BYTES);
#pragma unroll 1
            for (uint32_t at = 0; at + 1 <= SMEM_TABLE_BYTES; at += COPY_CHUNK) {
                const uint32_t bytes = min(COPY_CHUNK, SMEM_TABLE_BYTES - at);
                tma_bulk_load(smem_raw + at, reinterpret_cast<const unsigned char*>(p.table) + at, bytes, &table_ready);
            }
        }
    }
    // ---- fill the TMEM part: warps 0..3 each write the copy of their own lane quarter ----
    uint32_t tmem_lane_base = 0;
    if constexpr (KT > 0) {
        tmem_lane_base = tmem_base_slot + ((static_cast<uint32_t>(warp & 3) * 32u) << 16);
        if (warp < 4) {
            const float4* src =
                reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(p.table) + SMEM_TABLE_BYTES);
            for (int x = 0; x < kAlphabet; ++x) {
#pragma unroll
                for (int c = 0; c < KT / 2; ++c) {
                    const float2 a = __ldg(reinterpret_cast<const float2*>(src) + ((x * 32 + lane) * KT + 2 * c) / 2);
                    tmem_store2(tmem_lane_base + x * KT + 2 * c, a.x, a.y);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if constexpr (KS > 0) mbarrier_wait(&table_ready, 0);

    const uint32_t tab_lane = smem_u32(smem_raw) + lane * 16;
    const int left_lane = (lane + 31) & 31;
    const float NEG_INF = __int_as_float(0xff800000);
    const float tBMk = p.tr_B_Mk, tEC = p.tr_E_C, tEJ = p.tr_E_J;

    for (;;) {
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(p.queue_head, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= p.n) break;
        const uint32_t idx = __ldg(p.order + ticket);
        const uint64_t begin = __ldg(p.offsets + idx);
        const uint32_t len = static_cast<uint32_t>(__ldg(p.offsets + idx + 1) - begin);
        const float2 tr = __ldg(p.length_tr + len);
        const float loop = tr.x, move = tr.y;

        float m[K];
#pragma unroll
        for (int j = 0; j < K; ++j) m[j] = NEG_INF; // MSV_HMM.cpp:86
        float J = NEG_INF, C = NEG_INF, N = 0.0f, B = move; // MSV_HMM.cpp:96-97

        float te[KT > 0 ? KT : 1]; // emissions of the tensor-memory columns for the row at hand
        auto any_row = [&](auto exact_tag, const uint32_t x, const uint32_t x_next) {
            constexpr bool EXACT = decltype(exact_tag)::value;
            constexpr bool AHEAD = TMEM_AHEAD && KT > 0;
            if constexpr (KT > 0 && !AHEAD) tmem_load<KT>(tmem_lane_base + x * KT, te);
            const uint32_t erow = tab_lane + x * ROW_BYTES;
            const float bt = B + tBMk; // MSV_HMM.cpp:103, B -> M_k entry
            // lane 0 receives lane 31's last column, which is padding and therefore -inf: the dummy column M0
            const float left = __shfl_sync(0xffffffffu, m[K - 1], left_lane);
            float e = NEG_INF;
            // columns are updated from the highest down; TB / SB = first column served by tensor / shared memory
            [[maybe_unused]] constexpr int TB = AHEAD ? KS : 0, SB = AHEAD ? 0 : KT;
            const auto tensor_columns = [&] {
                if constexpr (KT > 0) {
                    tmem_wait<KT>(te);
#pragma unroll
                    for (int jj = KT - 1; jj >= 1; jj -= 2) {
                        const int j = TB + jj;
                        m[j] = te[jj] + fmaxf(m[j - 1], bt);
                        m[j - 1] = te[jj - 1] + fmaxf(j > 1 ? m[j > 1 ? j - 2 : 0] : left, bt);
                        e = fmaxf(fmaxf(e, m[j]), m[j - 1]); // MSV_HMM.cpp:104
                    }
                }
            };
            const auto shared_columns = [&] {
#pragma unroll
                for (int q = KS / 4 - 1; q >= 0; --q) {
                    const float4 ev = lds128(erow + q * 512);
                    const int j = SB + 4 * q;
                    m[j + 3] = ev.w + fmaxf(m[j + 2], bt);
                    m[j + 2] = ev.z + fmaxf(m[j + 1], bt);
                    m[j + 1] = ev.y + fmaxf(m[j], bt);
                    m[j] = ev.x + fmaxf(j ? m[j > 0 ? j - 1 : 0] : left, bt);
                    e = fmaxf(fmaxf(e, m[j + 3]), m[j + 2]);
                    e = fmaxf(fmaxf(e, m[j + 1]), m[j]);
                }
            };
            if constexpr (AHEAD) {
                tensor_columns();
                tmem_load<KT>(tmem_lane_base + x_next * KT, te); // for the next row; lands while this row finishes
                shared_columns();
            } else {
                shared_columns();
                tensor_columns();
            }
            if constexpr (EXACT) {
                float E;
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(E) : "f"(e));
                J = fmaxf(J + loop, E + tEJ);                           // MSV_HMM.cpp:107
                if constexpr (!CJ_SAME) C = fmaxf(C + loop, E + tEC);   // MSV_HMM.cpp:108
                N = N + loop;                                           // MSV_HMM.cpp:109
                B = fmaxf(N, J) + move; // MSV_HMM.cpp:110: max(N+move, J+move) == max(N, J)+move exactly (rounding is monotone)
            } else {
                J = fmaxf(J + loop, e + tEJ); // this lane's share of J; the lanes are combined once, after the last row
                N = N + loop;
                B = N + move;                 // = max(N, J) + move while J <= N -- verified below
            }
        };
        using Exact_row = std::bool_constant<true>;
        using Main_row = std::bool_constant<!SPEC>;
        auto row = [&](const uint32_t x, const uint32_t x_next) { any_row(Main_row{}, x, x_next); };

        // residues arrive as aligned 32-bit words (4 per load, prefetched two words ahead); a funnel shift undoes the
        // byte misalignment of the sequence start.  `word` holds the next four residues, `ahead` the four after them.
        const uint32_t shift = (static_cast<uint32_t>(begin) & 3u) * 8u;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.residues + (begin & ~static_cast<uint64_t>(3)));
        uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
        wp += 3;
        uint32_t word = __funnelshift_r(w0, w1, shift);
        // (an empty sequence's "first residue" is a foreign byte: clamp, as for every prefetch index below)
        if constexpr (TMEM_AHEAD && KT > 0) tmem_load<KT>(tmem_lane_base + min(word & 0xffu, static_cast<uint32_t>(kAlphabet - 1)) * KT, te);
        if constexpr (!CHECKPOINTS) {
        // long sequences are likely enough to contain a hit that speculating on them would mostly mean scanning them twice
        const bool speculate = SPEC && len <= 4096u;
        const uint32_t quads = (!SPEC || speculate) ? len >> 2 : 0u;
        // 4 rows (one residue word) per loop iteration, 8 or 16 where B200 sweeps showed a gain (+1..4 %, +10 % at K = 4;
        // profiles/r01/sweep_models_v6_row_unroll.jsonl, sweep_force_unroll2.txt).  It is not monotone in K -- the
        // instruction scheduler's luck -- and the longest bodies stop fitting the instruction cache (-10 % at K = 76).
        constexpr bool EIGHT_ROWS = K == 20 || K == 30 || K == 32 || K == 36 || K == 42 || K == 44 || K == 52 || K == 58 ||
                                    K == 60; // (K = 68 was in this list in round 1; re-measured: 9.25 TCUPS with 8 rows, 10.24 with 4 -- profiles/r02/unroll_sweep_v1/v2.txt; K = 54 likewise, by 0.6 %)
#ifdef MSV_FORCE_UNROLL // development aid (with MSV_QUICK_BUILD)
        constexpr int WORD_UNROLL = MSV_FORCE_UNROLL;
#else
        constexpr int WORD_UNROLL = K == 4 ? 4 : EIGHT_ROWS ? 2 : 1;
#endif
#pragma unroll WORD_UNROLL
        for (uint32_t i = 0; i < quads; ++i) {
            const uint32_t ahead = __funnelshift_r(w1, w2, shift);
            w1 = w2;
            w2 = __ldg(wp);
            ++wp;
            const uint32_t x0 = __byte_perm(word, 0, 0x4440), x1 = __byte_perm(word, 0, 0x4441);
            const uint32_t x2 = __byte_perm(word, 0, 0x4442), x3 = __byte_perm(word, 0, 0x4443);
            row(x0, x1);
            row(x1, x2);
            row(x2, x3);
            // the row after the last one of a sequence is somebody else's byte (the next sequence's, or memory that a later
            // upload stage has not filled or validated yet): it only feeds the tensor-memory PREFETCH, whose result is then
            // dropped, but it must never become a TMEM address outside the table -- hence the clamp to a real residue code
            row(x3, min(__byte_perm(ahead, 0, 0x4440), static_cast<uint32_t>(kAlphabet - 1)));
            word = ahead;
        }
#pragma unroll 1
        for (uint32_t r = (!SPEC || speculate) ? len & 3u : 0u; r > 0; --r) {
            row(word & 0xffu, min((word >> 8) & 0xffu, static_cast<uint32_t>(kAlphabet - 1)));
            word >>= 8;
        }
        if constexpr (SPEC) {
            if (!speculate || __any_sync(0xffffffffu, J >= N)) {
                if (speculate && lane == 0 && p.speculation_failures) atomicAdd(p.speculation_failures, 1u);
                // J may have overtaken N at some row, where B was then not N + move: scan this sequence again, exactly
                if constexpr (TMEM_AHEAD && KT > 0) tmem_wait<KT>(te); // retire the request made by the last speculative row
#pragma unroll
                for (int j = 0; j < K; ++j) m[j] = NEG_INF;
                J = NEG_INF, N = 0.0f, B = move;
                const uint8_t* rp = p.residues + begin;
                uint32_t x = __ldg(rp);
                if constexpr (TMEM_AHEAD && KT > 0) tmem_load<KT>(tmem_lane_base + x * KT, te);
#pragma unroll 1
                for (uint32_t i = 0; i < len; ++i) {
                    // the byte after the last residue is another sequence's or padding: prefetch index only, clamped (see above)
                    const uint32_t x_next = min(static_cast<uint32_t>(__ldg(rp + i + 1)), static_cast<uint32_t>(kAlphabet - 1));
                    any_row(Exact_row{}, x, x_next);
                    x = x_next;
                }
            } else {
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(J) : "f"(J));
            }
        }
        } else {
        const uint32_t quads = len >> 2;
        // 4 rows (one residue word) per loop iteration, 8 or 16 where B200 sweeps showed a gain (+1..4 %, +10 % at K = 4;
        // profiles/r01/sweep_models_v6_row_unroll.jsonl, sweep_force_unroll2.txt).  It is not monotone in K -- the
        // instruction scheduler's luck -- and the longest bodies stop fitting the instruction cache (-10 % at K = 76).
        constexpr bool EIGHT_ROWS = K == 20 || K == 30 || K == 32 || K == 36 || K == 42 || K == 44 || K == 52 || K == 58 ||
                                    K == 60; // (K = 68 was in this list in round 1; re-measured: 9.25 TCUPS with 8 rows, 10.24 with 4 -- profiles/r02/unroll_sweep_v1/v2.txt; K = 54 likewise, by 0.6 %)
#ifdef MSV_FORCE_UNROLL // development aid (with MSV_QUICK_BUILD)
        constexpr int WORD_UNROLL = MSV_FORCE_UNROLL;
#else
        constexpr int WORD_UNROLL = K == 4 ? 4 : EIGHT_ROWS ? 2 : 1;
#endif
        // SPECULATION WITH CHECKPOINTS.  The speculative rows run in blocks of kSpeculationBlockRows rows; before a block the
        // row state (m[], this lane's share of J, N) is parked in local memory, after it ONE vote asks whether J has overtaken
        // N.  If it has, only that block is lost: the state of its first row is restored -- it is exact, the vote before it
        // passed -- the lanes' shares are combined into J, and the rest of the sequence runs with the exact row.  A hit
        // therefore costs at most one block of rows plus the slower exact rows behind it, whatever the sequence length
        // (before: the whole sequence was scanned twice, and sequences beyond 4096 rows did not speculate at all).
#ifdef MSV_PARKED_PLAIN // development aid: let the compiler decide where the checkpoint lives
        [[maybe_unused]] float parked[SPEC ? K + 2 : 1];
#else
        [[maybe_unused]] volatile float parked[SPEC ? K + 2 : 1]; // volatile: really in local memory, not 46 more registers
#endif
        // (register diet: this kernel sits at its 128-register limit, and two more live values made the compiler recompute
        // lane constants from SR_TID inside the hot loop, -5 %.  Hence: the block's first word is derived from its end --
        // blocks start at multiples of the block size -- and the sequence index waits in shared memory.)
        constexpr uint32_t BLOCK_WORDS = kSpeculationBlockRows / 4;
        static_assert((BLOCK_WORDS & (BLOCK_WORDS - 1)) == 0, "block size must be a power of two");
        const auto four_rows = [&] { // one residue word
            const uint32_t ahead = __funnelshift_r(w1, w2, shift);
            w1 = w2;
            w2 = __ldg(wp);
            ++wp;
            const uint32_t x0 = __byte_perm(word, 0, 0x4440), x1 = __byte_perm(word, 0, 0x4441);
            const uint32_t x2 = __byte_perm(word, 0, 0x4442), x3 = __byte_perm(word, 0, 0x4443);
            row(x0, x1);
            row(x1, x2);
            row(x2, x3);
            // the row after the last one of a sequence is somebody else's byte (the next sequence's, or memory that a later
            // upload stage has not filled or validated yet): it only feeds the tensor-memory PREFETCH, whose result is then
            // dropped, but it must never become a TMEM address outside the table -- hence the clamp to a real residue code
            row(x3, min(__byte_perm(ahead, 0, 0x4440), static_cast<uint32_t>(kAlphabet - 1)));
            word = ahead;
        };
        const auto park = [&] {
            if constexpr (SPEC) {
#pragma unroll
                for (int j = 0; j < K; ++j) parked[j] = m[j];
                parked[K] = J;
                parked[K + 1] = N;
            }
        };
        bool overtaken = false;
        uint32_t first_row = 0; // first row of the block in which J overtook N
        uint32_t block_end = 0; // one past the last residue word of the block at hand
        for (;;) {
            const uint32_t block_first = block_end;
            block_end = SPEC ? min(quads, block_end + BLOCK_WORDS) : quads;
            park();
            // (written as a plain loop with `#pragma unroll`: measured 5-6 % faster than WORD_UNROLL hand-placed copies of the
            // body with exit tests between them -- the compiler's schedule of the row body differs, variant_sweep_v2/v3.txt)
#pragma unroll WORD_UNROLL
            for (uint32_t i = block_first; i < block_end; ++i) four_rows();
            const bool last_block = block_end == quads;
            if (last_block) {
#pragma unroll 1
                for (uint32_t r = len & 3u; r > 0; --r) {
                    row(word & 0xffu, min((word >> 8) & 0xffu, static_cast<uint32_t>(kAlphabet - 1)));
                    word >>= 8;
                }
            }
            if constexpr (SPEC) {
                if (__any_sync(0xffffffffu, J >= N)) {
                    overtaken = true;
                    break;
                }
            }
            if (last_block) break;
        }
        first_row = block_end ? 4u * ((block_end - 1u) & ~(BLOCK_WORDS - 1u)) : 0u; // blocks start at multiples of their size
        if constexpr (SPEC) {
            if (overtaken) {
                if (lane == 0 && p.speculation_failures) atomicAdd(p.speculation_failures, 1u);
                // J overtook N inside this block, where B was then not N + move: back to the block's first row, exactly from there
                if constexpr (TMEM_AHEAD && KT > 0) tmem_wait<KT>(te); // retire the request made by the last speculative row
#pragma unroll
                for (int j = 0; j < K; ++j) m[j] = parked[j];
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(J) : "f"(parked[K])); // the lanes' shares -> J
                N = parked[K + 1];
                B = N + move; // J < N here: the vote before this block passed
                // (the sequence's start is re-read here rather than kept in two registers through the hot loop: at the
                // 128-register limit of this kernel that made the compiler recompute lane constants from SR_TID inside the loop)
                const uint8_t* rp = p.residues + __ldg(p.offsets + idx) + first_row;
                uint32_t x = first_row < len ? __ldg(rp) : 0u;
                if constexpr (TMEM_AHEAD && KT > 0) tmem_load<KT>(tmem_lane_base + x * KT, te);
#pragma unroll 1
                for (uint32_t i = first_row; i < len; ++i) {
                    // the byte after the last residue is another sequence's or padding: prefetch index only, clamped (see above)
                    const uint32_t x_next = min(static_cast<uint32_t>(__ldg(rp + (i - first_row) + 1)), static_cast<uint32_t>(kAlphabet - 1));
                    any_row(Exact_row{}, x, x_next);
                    x = x_next;
                }
                if constexpr (TMEM_AHEAD && KT > 0) tmem_wait<KT>(te);
            } else {
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(J) : "f"(J));
            }
        }
        }
        if constexpr (CJ_SAME && !SPEC && REPORT) // exact rows: the same statistic the speculating rows produce (see the top of the kernel)
            if (lane == 0 && p.speculation_failures && J >= N) atomicAdd(p.speculation_failures, 1u);
        if (lane == 0) store_score(p, idx, (CJ_SAME ? J : C) + move); // MSV_HMM.cpp:112
    }

    if constexpr (KT > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "n"(TMEM_COLUMNS) : "memory");
    }
}

// =====================================================================================================================
// Four warps per sequence ("quad" kernel): for FEW and LONG sequences (config 5: 2 048 sequences of 10-35 k residues,
// the single-sequence API) and for models too long for one warp (LENG > 2 815).
//
// The row is serial in the sequence direction, so when there are fewer sequences than warp slots the makespan of the
// warp-per-sequence kernel is the longest sequence at one sixteenth of an SM.  Here a group of four warps shares one
// sequence: global lane gl = 32*(warp%4) + lane owns model columns gl*K+1 .. gl*K+K, so each warp holds a quarter of
// the row.  Per row the four warps exchange two words through shared memory and meet at one named barrier:
//   * the last column of each warp (boundary for the next warp's lane 0), double-buffered by row parity;
//   * each warp's partial E (after CREDUX); every warp then reduces the four partial maxima and carries N/J/B itself.
// The emission table is DISTRIBUTED rather than replicated: warp position q reads TMEM lane quarter q, which holds
// only that quarter's slice, so shared memory + tensor memory together hold up to 20 x 5 631 emissions (483 KB) --
// this is the "chunked" staging for models that do not fit shared memory alone.
// =====================================================================================================================
template <int K, int KT, int THREADS, bool CJ_SAME>
__global__ void __launch_bounds__(THREADS, 1) msv_scan_quad_kernel(const Scan_params p) {
    static_assert(KT == 0 || KT == 8 || KT == 16 || KT == 24, "TMEM columns per lane");
    static_assert(THREADS % 128 == 0, "whole groups of four warps");
    constexpr int KS = K - KT;
    static_assert(K % 4 == 0 && KS >= 0 && KS % 4 == 0, "columns per lane");
    constexpr uint32_t QUAD_BYTES = 128 * 16;         // one float4 per lane of the group
    constexpr uint32_t ROW_BYTES = (KS / 4) * QUAD_BYTES;
    constexpr uint32_t SMEM_TABLE_BYTES = kAlphabet * ROW_BYTES;
    constexpr uint32_t COPY_CHUNK = 32768;
    constexpr uint32_t TMEM_COLUMNS = 512;
    constexpr int MAX_GROUPS = THREADS / 128;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t table_ready;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float exchange[MAX_GROUPS][2][8]; // [group][row parity][4 partial E | 4 boundary columns]
    __shared__ uint32_t ticket_slot[MAX_GROUPS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wq = warp & 3;     // position inside the group == TMEM lane quarter
    const int group = warp >> 2;

    if (p.first_bad != nullptr && *p.first_bad != ~0ull) return;

    if (threadIdx.x == 0) mbarrier_init(&table_ready, 1);
    if constexpr (KT > 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                         "n"(TMEM_COLUMNS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if constexpr (KT > 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (KS > 0) {
        if (threadIdx.x == 0) {
            mbarrier_expect_tx(&table_ready, SMEM_TABLE_BYTES);
#pragma unroll 1
            for (uint32_t at = 0; at + 1 <= SMEM_TABLE_BYTES; at += COPY_CHUNK) {
                const uint32_t bytes = min(COPY_CHUNK, SMEM_TABLE_BYTES - at);
                tma_bulk_load(smem_raw + at, reinterpret_cast<const unsigned char*>(p.table) + at, bytes, &table_ready);
            }
        }
    }
    uint32_t tmem_lane_base = 0;
    if constexpr (KT > 0) {
        tmem_lane_base = tmem_base_slot + ((static_cast<uint32_t>(wq) * 32u) << 16);
        if (warp < 4) { // the first group fills all four quarters, each warp its own slice of the model
            const float4* src =
                reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(p.table) + SMEM_TABLE_BYTES);
            for (int x = 0; x < kAlphabet; ++x) {
#pragma unroll
                for (int c = 0; c < KT / 2; ++c) {
                    const float2 a = __ldg(reinterpret_cast<const float2*>(src) + ((x * 128 + wq * 32 + lane) * KT + 2 * c) / 2);
                    tmem_store2(tmem_lane_base + x * KT + 2 * c, a.x, a.y);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if constexpr (KS > 0) mbarrier_wait(&table_ready, 0);

    const uint32_t tab_lane = smem_u32(smem_raw) + (wq * 32 + lane) * 16;
    const uint32_t xchg = smem_u32(&exchange[group][0][0]);
    const int barrier_id = 1 + group;
    const float NEG_INF = __int_as_float(0xff800000);
    const float tBMk = p.tr_B_Mk, tEC = p.tr_E_C, tEJ = p.tr_E_J;
    auto group_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(barrier_id) : "memory"); };

    for (;;) {
        if (wq == 0 && lane == 0) ticket_slot[group] = atomicAdd(p.queue_head, 1u);
        group_sync();
        const uint32_t ticket = ticket_slot[group];
        if (ticket >= p.n) break;
        const uint32_t idx = __ldg(p.order + ticket);
        const uint64_t begin = __ldg(p.offsets + idx);
        const uint32_t len = static_cast<uint32_t>(__ldg(p.offsets + idx + 1) - begin);
        const float2 tr = __ldg(p.length_tr + len);
        const float loop = tr.x, move = tr.y;

        float m[K];
#pragma unroll
        for (int j = 0; j < K; ++j) m[j] = NEG_INF;
        float J = NEG_INF, C = NEG_INF, N = 0.0f, B = move;
        // boundary columns of "row 0" are -inf: written for parity 1 (row 1 reads parity (1-1)&1 ^ ... see below)
        if (lane < 8) {
            exchange[group][0][lane] = NEG_INF;
            exchange[group][1][lane] = NEG_INF;
        }
        group_sync();

        uint32_t parity = 0; // buffer written by the current row; the previous row wrote parity ^ 1
        auto row = [&](const uint32_t x) {
            float te[KT > 0 ? KT : 1];
            if constexpr (KT > 0) tmem_load<KT>(tmem_lane_base + x * KT, te);
            const uint32_t erow = tab_lane + x * ROW_BYTES;
            const float bt = B + tBMk;
            // left neighbour of this lane's first column: previous lane, or the previous warp's last column of the
            // previous row (double-buffered in shared memory), or -inf for the very first column of the model
            float left = __shfl_up_sync(0xffffffffu, m[K - 1], 1);
            if (lane == 0) {
                float carried;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(carried) : "r"(xchg + ((parity ^ 1u) * 8u + 4u + (wq > 0 ? wq - 1 : 0)) * 4u));
                left = wq == 0 ? NEG_INF : carried;
            }
            float e = NEG_INF;
#pragma unroll
            for (int q = KS / 4 - 1; q >= 0; --q) {
                const float4 ev = lds128(erow + q * QUAD_BYTES);
                const int j = KT + 4 * q;
                m[j + 3] = ev.w + fmaxf(m[j + 2], bt);
                m[j + 2] = ev.z + fmaxf(m[j + 1], bt);
                m[j + 1] = ev.y + fmaxf(m[j], bt);
                m[j] = ev.x + fmaxf(j ? m[j > 0 ? j - 1 : 0] : left, bt);
                e = fmaxf(fmaxf(e, m[j + 3]), m[j + 2]);
                e = fmaxf(fmaxf(e, m[j + 1]), m[j]);
            }
            if constexpr (KT > 0) {
                tmem_wait<KT>(te);
#pragma unroll
                for (int j = KT - 1; j >= 1; j -= 2) {
                    m[j] = te[j] + fmaxf(m[j - 1], bt);
                    m[j - 1] = te[j - 1] + fmaxf(j > 1 ? m[j > 1 ? j - 2 : 0] : left, bt);
                    e = fmaxf(fmaxf(e, m[j]), m[j - 1]);
                }
            }
            float warp_max;
            asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(warp_max) : "f"(e));
            if (lane == 31) {
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(xchg + (parity * 8u + 4u + wq) * 4u), "f"(m[K - 1]) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(xchg + (parity * 8u + wq) * 4u), "f"(warp_max) : "memory");
            }
            group_sync();
            float4 parts;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(parts.x), "=f"(parts.y), "=f"(parts.z), "=f"(parts.w)
                         : "r"(xchg + parity * 32u));
            const float E = fmaxf(fmaxf(fmaxf(parts.x, parts.y), parts.z), parts.w);
            J = fmaxf(J + loop, E + tEJ);
            if constexpr (!CJ_SAME) C = fmaxf(C + loop, E + tEC);
            N = N + loop;
            B = fmaxf(N, J) + move;
            parity ^= 1u;
        };

        const uint32_t shift = (static_cast<uint32_t>(begin) & 3u) * 8u;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.residues + (begin & ~static_cast<uint64_t>(3)));
        uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1);
        wp += 2;
        const uint32_t quads = len >> 2;
#pragma unroll 1
        for (uint32_t i = 0; i < quads; ++i) {
            const uint32_t word = __funnelshift_r(w0, w1, shift);
            w0 = w1;
            w1 = __ldg(wp);
            ++wp;
            row(__byte_perm(word, 0, 0x4440));
            row(__byte_perm(word, 0, 0x4441));
            row(__byte_perm(word, 0, 0x4442));
            row(__byte_perm(word, 0, 0x4443));
        }
        uint32_t word = __funnelshift_r(w0, w1, shift);
#pragma unroll 1
        for (uint32_t r = len & 3u; r > 0; --r) {
            row(word & 0xffu);
            word >>= 8;
        }
        if (wq == 0 && lane == 0) store_score(p, idx, (CJ_SAME ? J : C) + move);
    }

    if constexpr (KT > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "n"(TMEM_COLUMNS) : "memory");
    }
}

#ifndef MSV_KERNELS_TEMPLATES_ONLY // (the registry parts include this file for the kernel templates only)
// ---- gather of a sharded run, "push" form: after the scan, this GPU's slice of the job's score array (its own copy) is
// written into every peer's copy with coalesced stores over NVLink (128 bytes per warp instruction and peer).  The other
// form stores each score into all copies from the scan kernel itself (store_score); see launch_scan for which is used when.
struct Score_mirrors {
    float* copy[kMaxScoreMirrors];
};
__global__ void score_push_kernel(const float* __restrict__ own, const Score_mirrors peers, uint32_t n_peers, uint64_t n) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = own[i];
        for (uint32_t r = 0; r < n_peers; ++r) peers.copy[r][i] = v;
    }
}

// ---- database preparation kernels -------------------------------------------------------------------------------
// Validate residue codes (reference: unordered_map::at throws on a foreign letter, MSV_HMM.cpp:101) -- 16 B per thread.
__global__ void db_validate_kernel(const uint4* __restrict__ words, uint64_t n_words16, uint64_t n_bytes, uint64_t base_position,
                                   unsigned long long* __restrict__ first_bad) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_words16; i += stride) {
        const uint4 w = words[i];
        const uint32_t v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // a byte is bad iff it is >= 20: (b + 108) sets bit 7 for b >= 20 (b < 128), or b itself has bit 7
            const uint32_t t = ((v[k] & 0x7f7f7f7fu) + 0x6c6c6c6cu) | v[k];
            if (t & 0x80808080u) {
                for (int b = 0; b < 4; ++b) {
                    const uint64_t pos = i * 16 + k * 4 + b;
                    if (pos < n_bytes && ((v[k] >> (8 * b)) & 0xffu) >= kAlphabet)
                        atomicMin(first_bad, static_cast<unsigned long long>(base_position + pos));
                }
            }
        }
    }
}

// Longest-first bucketing (counting sort on length >> shift), three tiny kernels.
__global__ void db_histogram_kernel(const uint64_t* __restrict__ offsets, uint32_t n, uint32_t shift, uint32_t buckets,
                                    uint32_t* __restrict__ hist) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint64_t len = offsets[q + 1] - offsets[q];
    const uint32_t b = static_cast<uint32_t>(min(len >> shift, static_cast<uint64_t>(buckets - 1)));
    atomicAdd(hist + b, 1u);
}

// cursor[b] = number of sequences in buckets longer than b  (exclusive scan from the top bucket down); one CTA.
__global__ void db_scan_kernel(const uint32_t* __restrict__ hist, uint32_t buckets, uint32_t* __restrict__ cursor) {
    __shared__ uint32_t partial[1024];
    const uint32_t per = (buckets + blockDim.x - 1) / blockDim.x;
    // thread t owns the descending range of buckets [hi - per*t, ...)
    const int64_t top = static_cast<int64_t>(buckets) - 1 - static_cast<int64_t>(per) * threadIdx.x;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < per; ++k) {
        const int64_t b = top - k;
        if (b >= 0) sum += hist[b];
    }
    partial[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t t = 0; t < blockDim.x; ++t) {
            const uint32_t v = partial[t];
            partial[t] = run;
            run += v;
        }
    }
    __syncthreads();
    uint32_t run = partial[threadIdx.x];
    for (uint32_t k = 0; k < per; ++k) {
        const int64_t b = top - k;
        if (b >= 0) {
            cursor[b] = run;
            run += hist[b];
        }
    }
}

__global__ void db_scatter_kernel(const uint64_t* __restrict__ offsets, uint32_t n, uint32_t shift, uint32_t buckets,
                                  uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint64_t len = offsets[q + 1] - offsets[q];
    const uint32_t b = static_cast<uint32_t>(min(len >> shift, static_cast<uint64_t>(buckets - 1)));
    order[atomicAdd(cursor + b, 1u)] = q;
}

// ---- MSV filter statistics: raw score -> bit score -> Gumbel P-value (HMMER3 conventions) ----------------------------
__global__ void msv_filter_statistics_kernel(const float* __restrict__ scores, const uint64_t* __restrict__ offsets, uint32_t n,
                                             double mu, double lambda, float* __restrict__ bits_out, float* __restrict__ p_out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const double L = static_cast<double>(offsets[q + 1] - offsets[q]);
    const double null1 = L * log(L / (L + 1.0)) + log(1.0 / (L + 1.0));
    const double bits = (static_cast<double>(scores[q]) - null1) / 0.69314718055994530942;
    const double ey = -exp(-lambda * (bits - mu));
    const double p = fabs(ey) < 5e-9 ? -ey : 1.0 - exp(ey);
    if (bits_out) bits_out[q] = static_cast<float>(bits);
    if (p_out) p_out[q] = static_cast<float>(p);
}

#endif // MSV_KERNELS_TEMPLATES_ONLY

} // namespace msv
