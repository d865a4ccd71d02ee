// viterbi_kernels.cuh -- Plan-7 local multihit Viterbi scan for B200 (sm_100a): match, insert and delete states.
//
// This is SURVEY.md section 8(f) rank 4, the step the reference names as its direction (README.md:2-3) and for which it
// parses `transitions` (data_readers/Profile_HMM.hpp:27-29) without using them.  Same database, same special states,
// same emission table and same uniform local entry as the MSV path (reference algorithms/MSV_HMM.cpp:38-53,59-64,
// 107-112); per cell (i = residue, k = model column):
//     M[i][k] = e[x_i][k] + max(M[i-1][k-1] + tMM[k-1], I[i-1][k-1] + tIM[k-1], D[i-1][k-1] + tDM[k-1], B[i-1] + tBMk)
//     I[i][k] = max(M[i-1][k] + tMI[k], I[i-1][k] + tII[k])                  (insert emissions score 0, as in HMMER3)
//     D[i][k] = max(M[i][k-1] + tMD[k-1], D[i][k-1] + tDD[k-1])
//     E[i]    = max(max_k M[i][k], D[i][LENG])
// Every transcendental is evaluated on the host; the device only adds and takes maxima in fp32, each add on the same
// two operands as the scalar evaluation (oracle/viterbi_oracle.c), so scores are bit-identical to it:
//   * max is exact and order-free for the values that occur (finite and -inf), and rounding is monotone, so
//     fl(max(a, b) + t) == max(fl(a + t), fl(b + t)): the delete chain may be evaluated as a maximum over paths as long
//     as every path is summed left to right -- which is what the cross-lane propagation below does.
//
// Mapping.  One warp per sequence; lane l keeps K consecutive model columns of all three states in registers
// (m[K], in[K], d[K]).  The model is RIGHT-aligned in the 32*K slots: slot s = l*K + j holds column s - (32K - 1 - LENG),
// so the last real column is always the last slot of lane 31 (D[LENG] enters E with one predicated max) and slot 0 is
// always a dummy (-inf emissions and transitions), which makes the rotating shuffles below need no select.
//
// One ASCENDING pass per row does all three states of a column together.  Every transition is stored with the column it
// LEAVES ("outgoing"), so column j first turns its previous-row states into what column j+1 will receive,
//     a[j] = max(m[j] + tMM[j], in[j] + tIM[j], d[j] + tDM[j]),
// and only then overwrites them with the new row: m[j] = e + max(a[j-1], B + tBMk), in[j] from the old m[j], in[j], and
// d[j] = max(m_new[j-1] + tMD[j-1], d_new[j-1] + tDD[j-1]).  The serial delete chain therefore runs one column behind the
// match/insert arithmetic of the same pass, whose independent instructions hide its latency.
//
// Across lanes: the lane to the left hands over a[K-1] (previous row; one shuffle at the top of the row) and, after the
// pass, c = max(m_new[K-1] + tMD[K-1], d_new[K-1] + tDD[K-1]), the value of D in this lane's first slot.  Each lane has run
// its own chain with d[0] = -inf; the carried-in path c, c + tDD[0], c + tDD[0] + tDD[1], ... is then pushed through the
// lane four columns at a time for as long as it still improves some column in some lane (it loses ~1 nat per column
// against the local alternatives, so it normally dies within the first group), and the hand-over is repeated only if a
// carried-in path crossed an entire lane.  Every path is summed left to right, as the scalar evaluation does.
//
// Where the operands come from (per cell: 7 transitions + 1 emission = 32 B, more than shared memory alone can feed):
//   * tensor memory, 5 words per column (tMM, tIM, tDM, tMI, tII), laid out per lane in groups of 4 columns (20 words:
//     one tcgen05.ld.x16 + one .x4), fetched one group ahead;
//   * shared memory: the emission row of the residue (one LDS.128 per 4 columns) and tMD, tDD (two LDS.128 per 4
//     columns), staged once per CTA by the TMA unit.
#pragma once

#include <type_traits>

#include "msv_device.cuh"

namespace msv {

constexpr uint32_t kViterbiSpeculationMaxLength = 4096;
constexpr int kViterbiMaxColumnsPerLane = 80; // 22 * K * 128 B of shared memory

// Table in global memory (floats), K columns per lane, Q = K / 4; every transition belongs to the column it leaves:
//   [0, 20*K*32)                  emissions              [residue][q][lane][4]
//   [.., + K*32)                  tMD                    [q][lane][4]
//   [.., + K*32)                  tDD                    [q][lane][4]          <- end of the shared-memory part
//   [.., + 32*5*K)                tensor-memory part     [lane][q][tMM x4 | tIM x4 | tDM x4 | tMI x4 | tII x4]
//   [.., + 32*8)                  per lane: tMM, tIM, tDM, tMD, tDD of the lane's LAST slot (what it hands to the right; -inf
//                                 for lane 31), 3 unused words
// SPECULATE (tr_E_C == tr_E_J only): as in the MSV warp kernel, B[i] = max(N[i], J[i]) + move is N[i] + move while J <= N, so
// the row needs no warp-wide E: each lane carries its share j = max(j + loop, E_lane + tEJ), one vote per sequence checks
// "no lane ends with j >= N", and a sequence that fails it is scanned again with the exact row.  Same bits.
template <int K, int THREADS, bool CJ_SAME, bool SPECULATE = false>
__global__ void __launch_bounds__(THREADS, 1) viterbi_scan_warp_kernel(const Scan_params p) {
    constexpr bool SPEC = SPECULATE && CJ_SAME;
    static_assert(K % 4 == 0 && K >= 4 && K <= kViterbiMaxColumnsPerLane, "columns per lane");
    constexpr int Q = K / 4;
    constexpr uint32_t ROW_BYTES = K * 32 * 4; // one residue's emissions; also the size of the tMD and of the tDD block
    constexpr uint32_t EMISSION_BYTES = kAlphabet * ROW_BYTES;
    constexpr uint32_t SMEM_TABLE_BYTES = EMISSION_BYTES + 2 * ROW_BYTES;
    constexpr uint32_t TENSOR_WORDS = 5 * K; // per lane
    constexpr uint32_t COPY_CHUNK = 32768;
    constexpr uint32_t TMEM_COLUMNS = 512;
    static_assert(TENSOR_WORDS <= TMEM_COLUMNS, "transitions of one lane must fit its tensor-memory lane");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t table_ready;
    __shared__ uint32_t tmem_base_slot;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    if (p.first_bad != nullptr && *p.first_bad != ~0ull) return; // never index the table with an unvalidated residue code

    if (threadIdx.x == 0) mbarrier_init(&table_ready, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "n"(TMEM_COLUMNS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        mbarrier_expect_tx(&table_ready, SMEM_TABLE_BYTES);
#pragma unroll 1
        for (uint32_t at = 0; at < SMEM_TABLE_BYTES; at += COPY_CHUNK) {
            const uint32_t bytes = min(COPY_CHUNK, SMEM_TABLE_BYTES - at);
            tma_bulk_load(smem_raw + at, reinterpret_cast<const unsigned char*>(p.table) + at, bytes, &table_ready);
        }
    }
    const float* tensor_src = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p.table) + SMEM_TABLE_BYTES);
    const uint32_t tmem_lane_base = tmem_base_slot + ((static_cast<uint32_t>(warp & 3) * 32u) << 16);
    if (warp < 4) { // each of the first four warps writes the copy of its own lane quarter
        const float2* src = reinterpret_cast<const float2*>(tensor_src + static_cast<size_t>(lane) * TENSOR_WORDS);
#pragma unroll 4
        for (uint32_t w = 0; w < TENSOR_WORDS / 2; ++w) {
            const float2 a = __ldg(src + w);
            tmem_store2(tmem_lane_base + 2 * w, a.x, a.y);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbarrier_wait(&table_ready, 0);

    const float4 edge = __ldg(reinterpret_cast<const float4*>(tensor_src + 32 * TENSOR_WORDS) + 2 * lane);
    const float edge_dd = __ldg(tensor_src + 32 * TENSOR_WORDS + 8 * lane + 4);
    const uint32_t tab_lane = smem_u32(smem_raw) + lane * 16;
    const uint32_t md_lane = tab_lane + EMISSION_BYTES, dd_lane = md_lane + ROW_BYTES;
    const float4 dd_first = lds128(dd_lane); // tDD of this lane's first four slots, kept in registers
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

        float m[K] = {}, in[K] = {}, d[K] = {}; // (set by scan() below before any row reads them)
        float J = 0.0f, C = 0.0f, N = 0.0f, B = 0.0f;

        // transitions of the column group at hand / of the next one (tensor memory, double buffered across the unrolled loop)
        float tq[2][20];

        auto row = [&](auto exact_tag, const uint32_t x) {
            constexpr bool EXACT = decltype(exact_tag)::value;
            const uint32_t erow = tab_lane + x * ROW_BYTES;
            const float bt = B + tBMk;
            // what the first slot of the lane to the right receives from this lane's last slot (previous row); lane 31's
            // edge transitions are -inf, so lane 0 receives -inf: nothing enters the model's left end
            const float a_last = fmaxf(fmaxf(m[K - 1] + edge.x, in[K - 1] + edge.y), d[K - 1] + edge.z);
            float a_prev = __shfl_sync(0xffffffffu, a_last, left_lane);
            float e = NEG_INF;
            float md_prev = NEG_INF, dd_prev = NEG_INF; // tMD, tDD of the column one to the left (none for the lane's first slot)
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float* t = tq[q & 1];
                tmem_wait<20>(t);
                if (q + 1 < Q) tmem_load<20>(tmem_lane_base + (q + 1) * 20, tq[(q + 1) & 1]);
                const float4 ev = lds128(erow + q * 512);
                const float4 md4 = lds128_volatile(md_lane + q * 512), dd4 = lds128_volatile(dd_lane + q * 512);
                const float em[4] = {ev.x, ev.y, ev.z, ev.w}, md[4] = {md4.x, md4.y, md4.z, md4.w}, dd[4] = {dd4.x, dd4.y, dd4.z, dd4.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = 4 * q + c;
                    const float a = j == K - 1 ? a_last : fmaxf(fmaxf(m[j] + t[c], in[j] + t[4 + c]), d[j] + t[8 + c]);
                    const float ins = fmaxf(m[j] + t[12 + c], in[j] + t[16 + c]);
                    const float m_new = em[c] + fmaxf(a_prev, bt);
                    d[j] = j == 0 ? NEG_INF : fmaxf(m[j > 0 ? j - 1 : 0] + md_prev, d[j > 0 ? j - 1 : 0] + dd_prev); // m[j-1], d[j-1]: new row
                    m[j] = m_new;
                    in[j] = ins;
                    e = fmaxf(e, m_new);
                    a_prev = a;
                    md_prev = md[c];
                    dd_prev = dd[c];
                }
                // after the highest group: request the lowest group for the next row; issued after the group's arithmetic
                // because for odd Q both use the same registers
                if (q == Q - 1) tmem_load<20>(tmem_lane_base, tq[0]);
            }
            // ---- delete paths that arrive from the lane to the left ----
            // The first four columns are done unconditionally with tDD from registers (the carried-in path nearly always
            // improves them): no vote, no shared-memory load on this serial stretch.  Further groups only while a vote says
            // the path is still alive somewhere; the hand-over repeats only if some lane's last column changed.
            for (;;) {
                const float last_before = d[K - 1];
                const float c_out = fmaxf(m[K - 1] + edge.w, d[K - 1] + edge_dd); // D of the right neighbour's first slot
                float carried = __shfl_sync(0xffffffffu, c_out, left_lane);       // lane 0 receives lane 31's -inf
                d[0] = fmaxf(d[0], carried);
                carried = carried + dd_first.x;
                d[1] = fmaxf(d[1], carried);
                carried = carried + dd_first.y;
                d[2] = fmaxf(d[2], carried);
                carried = carried + dd_first.z;
                d[3] = fmaxf(d[3], carried);
                carried = carried + dd_first.w;
                bool crossed = true; // a carried path is still alive after the last column of some lane
#pragma unroll
                for (int q = 1; q < Q; ++q) {
                    const int j = 4 * q;
                    if (!__any_sync(0xffffffffu, carried > d[j])) { // dead everywhere: nothing further can change
                        crossed = false;
                        break;
                    }
                    const float4 dd = lds128_volatile(dd_lane + q * 512);
                    d[j] = fmaxf(d[j], carried);
                    carried = carried + dd.x;
                    d[j + 1] = fmaxf(d[j + 1], carried);
                    carried = carried + dd.y;
                    d[j + 2] = fmaxf(d[j + 2], carried);
                    carried = carried + dd.z;
                    d[j + 3] = fmaxf(d[j + 3], carried);
                    carried = carried + dd.w;
                }
                if (!crossed || !__any_sync(0xffffffffu, d[K - 1] > last_before)) break;
            }
            if (lane == 31) e = fmaxf(e, d[K - 1]); // D[LENG] -> E
            if constexpr (EXACT) {
                float E;
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(E) : "f"(e));
                J = fmaxf(J + loop, E + tEJ);
                if constexpr (!CJ_SAME) C = fmaxf(C + loop, E + tEC);
                N = N + loop;
                B = fmaxf(N, J) + move;
            } else {
                J = fmaxf(J + loop, e + tEJ); // this lane's share of J
                N = N + loop;
                B = N + move;                 // = max(N, J) + move while J <= N; verified after the last row
            }
        };

        // one pass over the sequence with the given kind of row.  Residues arrive as aligned 32-bit words; a funnel shift
        // undoes the byte misalignment of the sequence start.
        auto scan = [&](auto exact_tag) {
#pragma unroll
            for (int j = 0; j < K; ++j) m[j] = in[j] = d[j] = NEG_INF;
            J = NEG_INF, C = NEG_INF, N = 0.0f, B = move;
            tmem_load<20>(tmem_lane_base, tq[0]);
            const uint32_t shift = (static_cast<uint32_t>(begin) & 3u) * 8u;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.residues + (begin & ~static_cast<uint64_t>(3)));
            uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1);
            wp += 2;
            uint32_t word = __funnelshift_r(w0, w1, shift);
#pragma unroll 1
            for (uint32_t i = 0; i < len; ++i) {
                row(exact_tag, word & 0xffu);
                word >>= 8;
                if ((i & 3u) == 3u) {
                    w0 = w1;
                    w1 = __ldg(wp);
                    ++wp;
                    word = __funnelshift_r(w0, w1, shift);
                }
            }
            tmem_wait<20>(tq[0]); // retire the group that was requested for a row that does not exist
        };
        if (SPEC && len <= kViterbiSpeculationMaxLength) {
            scan(std::bool_constant<!SPEC>{});
            if (__any_sync(0xffffffffu, J >= N)) scan(std::bool_constant<true>{}); // J may have overtaken N: exact rows
            else asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(J) : "f"(J));
        } else {
            scan(std::bool_constant<true>{});
        }
        if (lane == 0) store_score(p, idx, (CJ_SAME ? J : C) + move);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "n"(TMEM_COLUMNS) : "memory");
}

// =====================================================================================================================
// Lane-group variant for SHORT models: G = 8 lanes per sequence, four sequences per warp.
//
// With one warp per sequence a 100- or 200-column model leaves each lane 4 or 8 columns: the per-row bookkeeping outweighs
// the cells and a carried-in delete path often outlives a whole lane, so the hand-over repeats.  Here a lane keeps K = 16 ..
// 56 columns of one of FOUR sequences; bookkeeping, votes and the tensor-memory transition loads (their addresses do not
// depend on the residue, so they stay warp-uniform; every group reads the same values, replicated over the 32 TMEM lanes)
// are shared by the four.  Emissions differ per group and come from shared memory, [residue][q][lane of group][4].
// Groups retire and fetch sequences independently; rows run in warp-uniform chunks (the minimum over the groups of their
// remaining rows), as in msv_scan_kernel.  Exact rows only (the group-wide E is three shuffles).
// Table: as for the warp kernel with 32 replaced by G in the shared-memory part and in the edge block; the tensor-memory
// part is [32 lanes][...] with lane l holding the transitions of group lane l % G.
// =====================================================================================================================
template <int G, int K, int THREADS, bool CJ_SAME>
__global__ void __launch_bounds__(THREADS, 1) viterbi_scan_group_kernel(const Scan_params p) {
    static_assert(G == 8 || G == 16, "lanes per sequence");
    static_assert(K % 4 == 0 && K >= 4 && K <= kViterbiMaxColumnsPerLane, "columns per lane");
    constexpr int Q = K / 4;
    constexpr uint32_t QUAD_BYTES = G * 16;
    constexpr uint32_t ROW_BYTES = K * G * 4;
    constexpr uint32_t EMISSION_BYTES = kAlphabet * ROW_BYTES;
    constexpr uint32_t SMEM_TABLE_BYTES = EMISSION_BYTES + 2 * ROW_BYTES;
    constexpr uint32_t TENSOR_WORDS = 5 * K;
    constexpr uint32_t COPY_CHUNK = 32768;
    constexpr uint32_t TMEM_COLUMNS = 512;
    static_assert(TENSOR_WORDS <= TMEM_COLUMNS, "transitions of one lane must fit its tensor-memory lane");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t table_ready;
    __shared__ uint32_t tmem_base_slot;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1);

    if (p.first_bad != nullptr && *p.first_bad != ~0ull) return;

    if (threadIdx.x == 0) mbarrier_init(&table_ready, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "n"(TMEM_COLUMNS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        mbarrier_expect_tx(&table_ready, SMEM_TABLE_BYTES);
#pragma unroll 1
        for (uint32_t at = 0; at < SMEM_TABLE_BYTES; at += COPY_CHUNK) {
            const uint32_t bytes = min(COPY_CHUNK, SMEM_TABLE_BYTES - at);
            tma_bulk_load(smem_raw + at, reinterpret_cast<const unsigned char*>(p.table) + at, bytes, &table_ready);
        }
    }
    const float* tensor_src = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p.table) + SMEM_TABLE_BYTES);
    const uint32_t tmem_lane_base = tmem_base_slot + ((static_cast<uint32_t>(warp & 3) * 32u) << 16);
    if (warp < 4) {
        const float2* src = reinterpret_cast<const float2*>(tensor_src + static_cast<size_t>(lane) * TENSOR_WORDS);
#pragma unroll 4
        for (uint32_t w = 0; w < TENSOR_WORDS / 2; ++w) {
            const float2 a = __ldg(src + w);
            tmem_store2(tmem_lane_base + 2 * w, a.x, a.y);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbarrier_wait(&table_ready, 0);

    const float4 edge = __ldg(reinterpret_cast<const float4*>(tensor_src + 32 * TENSOR_WORDS) + 2 * gl);
    const float edge_dd = __ldg(tensor_src + 32 * TENSOR_WORDS + 8 * gl + 4);
    const uint32_t tab_lane = smem_u32(smem_raw) + gl * 16;
    const uint32_t md_lane = tab_lane + EMISSION_BYTES, dd_lane = md_lane + ROW_BYTES;
    const float4 dd_first = lds128(dd_lane);
    const unsigned gmask = ((1u << G) - 1u) << (lane & ~(G - 1));
    const int left_lane = (lane & ~(G - 1)) | ((gl + G - 1) & (G - 1)); // rotate inside the group
    const float NEG_INF = __int_as_float(0xff800000);
    const float tBMk = p.tr_B_Mk, tEC = p.tr_E_C, tEJ = p.tr_E_J;

    float m[K], in[K], d[K];
#pragma unroll
    for (int j = 0; j < K; ++j) m[j] = in[j] = d[j] = NEG_INF;
    float J = NEG_INF, C = NEG_INF, N = 0.0f, B = NEG_INF, loop = 0.0f, move = 0.0f;
    uint32_t remaining = 0, idx = 0;
    bool active = false, done = false;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.residues);
    uint32_t wlo = 0, whi = 0, phase = 0; // residue window of the group's sequence: next residue = byte phase/8 of (whi:wlo)

    float tq[2][20];
    tmem_load<20>(tmem_lane_base, tq[0]);

    auto row = [&](const uint32_t x) {
        const uint32_t erow = tab_lane + x * ROW_BYTES;
        const float bt = B + tBMk;
        const float a_last = fmaxf(fmaxf(m[K - 1] + edge.x, in[K - 1] + edge.y), d[K - 1] + edge.z);
        float a_prev = __shfl_sync(0xffffffffu, a_last, left_lane); // the group's last lane hands -inf to its first
        float e = NEG_INF;
        float md_prev = NEG_INF, dd_prev = NEG_INF;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float* t = tq[q & 1];
            tmem_wait<20>(t);
            if (q + 1 < Q) tmem_load<20>(tmem_lane_base + (q + 1) * 20, tq[(q + 1) & 1]);
            const float4 ev = lds128(erow + q * QUAD_BYTES);
            const float4 md4 = lds128_volatile(md_lane + q * QUAD_BYTES), dd4 = lds128_volatile(dd_lane + q * QUAD_BYTES);
            const float em[4] = {ev.x, ev.y, ev.z, ev.w}, md[4] = {md4.x, md4.y, md4.z, md4.w}, dd[4] = {dd4.x, dd4.y, dd4.z, dd4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 4 * q + c;
                const float a = j == K - 1 ? a_last : fmaxf(fmaxf(m[j] + t[c], in[j] + t[4 + c]), d[j] + t[8 + c]);
                const float ins = fmaxf(m[j] + t[12 + c], in[j] + t[16 + c]);
                const float m_new = em[c] + fmaxf(a_prev, bt);
                d[j] = j == 0 ? NEG_INF : fmaxf(m[j > 0 ? j - 1 : 0] + md_prev, d[j > 0 ? j - 1 : 0] + dd_prev);
                m[j] = m_new;
                in[j] = ins;
                e = fmaxf(e, m_new);
                a_prev = a;
                md_prev = md[c];
                dd_prev = dd[c];
            }
            if (q == Q - 1) tmem_load<20>(tmem_lane_base, tq[0]);
        }
        for (;;) { // delete paths that arrive from the lane to the left (see the warp kernel)
            const float last_before = d[K - 1];
            const float c_out = fmaxf(m[K - 1] + edge.w, d[K - 1] + edge_dd);
            float carried = __shfl_sync(0xffffffffu, c_out, left_lane);
            d[0] = fmaxf(d[0], carried);
            carried = carried + dd_first.x;
            d[1] = fmaxf(d[1], carried);
            carried = carried + dd_first.y;
            d[2] = fmaxf(d[2], carried);
            carried = carried + dd_first.z;
            d[3] = fmaxf(d[3], carried);
            carried = carried + dd_first.w;
            bool crossed = true;
#pragma unroll
            for (int q = 1; q < Q; ++q) {
                const int j = 4 * q;
                if (!__any_sync(0xffffffffu, carried > d[j])) {
                    crossed = false;
                    break;
                }
                const float4 dd = lds128_volatile(dd_lane + q * QUAD_BYTES);
                d[j] = fmaxf(d[j], carried);
                carried = carried + dd.x;
                d[j + 1] = fmaxf(d[j + 1], carried);
                carried = carried + dd.y;
                d[j + 2] = fmaxf(d[j + 2], carried);
                carried = carried + dd.z;
                d[j + 3] = fmaxf(d[j + 3], carried);
                carried = carried + dd.w;
            }
            if (!crossed || !__any_sync(0xffffffffu, d[K - 1] > last_before)) break;
        }
        if (gl == G - 1) e = fmaxf(e, d[K - 1]); // D[LENG] -> E
#pragma unroll
        for (int s = G / 2; s > 0; s >>= 1) e = fmaxf(e, __shfl_xor_sync(0xffffffffu, e, s));
        J = fmaxf(J + loop, e + tEJ);
        if constexpr (!CJ_SAME) C = fmaxf(C + loop, e + tEC);
        N = N + loop;
        B = fmaxf(N, J) + move;
    };

    for (;;) {
        while (remaining == 0 && !done) { // retire / fetch, group-uniform
            if (active) {
                if (gl == 0) store_score(p, idx, (CJ_SAME ? J : C) + move);
                active = false;
            }
            uint32_t ticket = 0;
            if (gl == 0) ticket = atomicAdd(p.queue_head, 1u);
            ticket = __shfl_sync(gmask, ticket, 0, G);
#pragma unroll
            for (int j = 0; j < K; ++j) m[j] = in[j] = d[j] = NEG_INF;
            J = NEG_INF;
            C = NEG_INF;
            N = 0.0f;
            if (ticket >= p.n) { // finished groups keep executing rows on a dead state: everything stays -inf, no vote fires
                done = true;
                B = NEG_INF;
                loop = 0.0f;
                move = NEG_INF;
                wp = reinterpret_cast<const uint32_t*>(p.residues);
                wlo = whi = phase = 0;
                break;
            }
            idx = __ldg(p.order + ticket);
            const uint64_t begin = __ldg(p.offsets + idx);
            const uint32_t len = static_cast<uint32_t>(__ldg(p.offsets + idx + 1) - begin);
            const float2 tr = __ldg(p.length_tr + len);
            loop = tr.x;
            move = tr.y;
            B = move;
            const uint32_t mis = static_cast<uint32_t>(begin) & 3u;
            wp = reinterpret_cast<const uint32_t*>(p.residues + (begin - mis));
            wlo = __ldg(wp);
            whi = __ldg(wp + 1);
            wp += 2;
            phase = 8u * mis;
            remaining = len;
            active = true;
        }
        const uint32_t mine = done ? 0xffffffffu : remaining;
        const uint32_t steps = __reduce_min_sync(0xffffffffu, mine);
        if (steps == 0xffffffffu) break;
        if (!done) remaining -= steps;
        const uint32_t advance = done ? 0u : 1u;
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
            row(x);
        }
    }
    tmem_wait<20>(tq[0]);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "n"(TMEM_COLUMNS) : "memory");
}

} // namespace msv
