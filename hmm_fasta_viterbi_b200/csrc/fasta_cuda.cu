// fasta_cuda.cu -- FASTA text -> packed, device-resident database, parsed ON THE GPU (msv_cuda_db_create_from_fasta,
// msv_cuda_score_fasta in include/msv_cuda.h).
//
// What it replaces: FASTA_protein_sequences::FASTA_protein_sequences (reference data_readers/FASTA_protein_sequences.cpp:9-44),
// a getline loop plus a hash-set lookup per character, and the re-encoding every scoring call does on top of it
// (reference algorithms/MSV_HMM.cpp:101).  With a scan that runs at ~10 TCUPS that reader is what an end-to-end run waits
// for, so here the raw text goes over the link as it is and the GPU classifies, encodes and cuts it:
//   * record rules are the reference's: a line that starts with '>' opens a record and its text is dropped, every byte of
//     every other line belongs to the open record, a record with any byte outside ACDEFGHIKLMNPQRSTVWY is dropped WHOLE
//     (FASTA_protein_sequences.cpp:26-41; a '\r' is such a byte), text before the first header is ignored;
//   * byte work, HBM-bound: the text is read three times and the codes written twice, a few hundred microseconds per
//     100 MB against milliseconds for the upload -- which is therefore the part that is engineered: pageable text (an
//     mmap'ed file) is staged through a ring of pinned buffers by several host threads, each with its own stream.
//
// Passes (TILE = 4096 bytes per CTA, 16 per thread):
//   A  per tile: does it contain a line start, and what kind (header / body) is its last line          -> tile_flags
//   B  one CTA: kind of the line each tile starts in (the last line start before it)                    -> tile_entry
//   C  per tile: headers and residue bytes in it                                                        -> tile_counts
//   D  one CTA: exclusive scan of both counts over the tiles, totals                                    -> tile_base, totals
//   E  per tile: for every byte its record and its rank among the residue bytes; writes the code to rank, the rank at
//      every header to record_start, and marks records that contain a foreign byte                      -> codes, record_*
//   F  records: keep = not marked (record 0 = text before the first header is never kept); exclusive scan of keep and
//      keep * length                                                                                    -> offsets, totals, longest
//   G  one warp per kept record copies its codes to their final place (skipped when nothing was rejected: the codes of
//      pass E are then already the database)
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "msv_internal.hpp"

namespace {

constexpr int kThreads = 256;
constexpr int kBytesPerThread = 16;
constexpr uint32_t kTile = kThreads * kBytesPerThread;

// letter -> code (order ACDEFGHIKLMNPQRSTVWY, reference MSV_HMM.cpp:29-31), 0xff for anything else
__device__ __forceinline__ uint32_t residue_code(uint32_t c) {
    // 'A'..'Y' -> 0..24, then a 25-entry table packed five bits per entry (31 = foreign) into 64-bit words of 12 entries
    const uint32_t i = c - 'A';
    if (i > 24u) return 0xffu;
    //   A  B  C  D  E  F  G  H  I  J  K  L  |  M   N   O  P   Q   R   S   T   U  V   W   X  |  Y
    //   0  x  1  2  3  4  5  6  7  x  8  9  |  10  11  x  12  13  14  15  16  x  17  18  x  |  19
    constexpr unsigned long long a_to_l = 0ull | (31ull << 5) | (1ull << 10) | (2ull << 15) | (3ull << 20) | (4ull << 25) | (5ull << 30) |
                                          (6ull << 35) | (7ull << 40) | (31ull << 45) | (8ull << 50) | (9ull << 55);
    constexpr unsigned long long m_to_x = 10ull | (11ull << 5) | (31ull << 10) | (12ull << 15) | (13ull << 20) | (14ull << 25) | (15ull << 30) |
                                          (16ull << 35) | (31ull << 40) | (17ull << 45) | (18ull << 50) | (31ull << 55);
    const uint32_t v = i < 12u ? static_cast<uint32_t>(a_to_l >> (5u * i)) & 31u : i < 24u ? static_cast<uint32_t>(m_to_x >> (5u * (i - 12u))) & 31u : 19u;
    return v == 31u ? 0xffu : v;
}

struct Span { // what a thread learns from its 16 bytes without knowing the line it starts in
    uint32_t has_line_start; // a line starts inside the span
    uint32_t last_is_header; // kind of the last line that starts inside it
};

// the 16 bytes of thread `t` of tile `tile` (zero beyond the text), and the byte just before them ('\n' before the text)
__device__ __forceinline__ void load_span(const uint8_t* __restrict__ text, uint64_t bytes, uint64_t at, uint8_t (&b)[kBytesPerThread],
                                          uint8_t& before, uint32_t& valid) {
    valid = at >= bytes ? 0u : static_cast<uint32_t>(min(static_cast<uint64_t>(kBytesPerThread), bytes - at));
    if (valid == kBytesPerThread) {
        const uint4 w = *reinterpret_cast<const uint4*>(text + at); // the text buffer is 16-byte aligned
        const uint32_t v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < kBytesPerThread; ++i) b[i] = static_cast<uint8_t>(v[i / 4] >> (8 * (i % 4)));
    } else {
#pragma unroll
        for (int i = 0; i < kBytesPerThread; ++i) b[i] = static_cast<uint32_t>(i) < valid ? text[at + i] : 0;
    }
    before = at == 0 ? '\n' : (at <= bytes ? text[at - 1] : 0);
}

__device__ __forceinline__ Span classify_span(const uint8_t (&b)[kBytesPerThread], uint8_t before, uint32_t valid) {
    Span s{0, 0};
    uint8_t prev = before;
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i) {
        if (static_cast<uint32_t>(i) < valid && prev == '\n') {
            s.has_line_start = 1;
            s.last_is_header = b[i] == '>';
        }
        prev = b[i];
    }
    return s;
}

// kind (1 = header line) of the line in which each thread's span starts: the last line start in an earlier span of the
// tile, else the tile's entry kind.  One ballot per warp, the warps chained through shared memory.
__device__ __forceinline__ uint32_t entry_kind_in_tile(const Span s, uint32_t tile_entry, uint32_t* warp_exit /* [kThreads/32][2] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t starts = __ballot_sync(0xffffffffu, s.has_line_start);
    const uint32_t headers = __ballot_sync(0xffffffffu, s.last_is_header);
    if (lane == 0) {
        warp_exit[2 * warp] = starts != 0;
        warp_exit[2 * warp + 1] = starts ? (headers >> (31 - __clz(starts))) & 1u : 0u;
    }
    __syncthreads();
    uint32_t kind = tile_entry;
    for (int w = 0; w < warp; ++w)
        if (warp_exit[2 * w]) kind = warp_exit[2 * w + 1];
    const uint32_t before_me = starts & ((1u << lane) - 1u);
    if (before_me) kind = (headers >> (31 - __clz(before_me))) & 1u;
    __syncthreads();
    return kind;
}

// ---- pass A -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) fasta_tile_flags_kernel(const uint8_t* __restrict__ text, uint64_t bytes, uint8_t* __restrict__ tile_flags) {
    __shared__ uint32_t warp_exit[kThreads / 32][2];
    uint8_t b[kBytesPerThread], before;
    uint32_t valid;
    load_span(text, bytes, static_cast<uint64_t>(blockIdx.x) * kTile + threadIdx.x * kBytesPerThread, b, before, valid);
    const Span s = classify_span(b, before, valid);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t starts = __ballot_sync(0xffffffffu, s.has_line_start);
    const uint32_t headers = __ballot_sync(0xffffffffu, s.last_is_header);
    if (lane == 0) {
        warp_exit[warp][0] = starts != 0;
        warp_exit[warp][1] = starts ? (headers >> (31 - __clz(starts))) & 1u : 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t any = 0, kind = 0;
        for (int w = 0; w < kThreads / 32; ++w)
            if (warp_exit[w][0]) any = 1, kind = warp_exit[w][1];
        tile_flags[blockIdx.x] = static_cast<uint8_t>(any | (kind << 1));
    }
}

// ---- pass B: entry kind of every tile = exit kind of the nearest earlier tile that has a line start (one CTA) -----------------
__global__ void __launch_bounds__(1024) fasta_tile_entry_kernel(const uint8_t* __restrict__ tile_flags, uint32_t tiles, uint8_t* __restrict__ tile_entry) {
    __shared__ uint32_t chunk_exit[1024];
    const uint32_t per = (tiles + blockDim.x - 1) / blockDim.x;
    const uint32_t first = threadIdx.x * per, last = min(tiles, first + per);
    uint32_t mine = 2; // 2 = no line start in my tiles
    for (uint32_t t = first; t < last; ++t)
        if (tile_flags[t] & 1) mine = (tile_flags[t] >> 1) & 1;
    chunk_exit[threadIdx.x] = mine;
    __syncthreads();
    uint32_t kind = 0; // (the first tile starts at a line start, so its entry kind is never used)
    for (int c = static_cast<int>(threadIdx.x) - 1; c >= 0; --c)
        if (chunk_exit[c] != 2) {
            kind = chunk_exit[c];
            break;
        }
    for (uint32_t t = first; t < last; ++t) {
        tile_entry[t] = static_cast<uint8_t>(kind);
        if (tile_flags[t] & 1) kind = (tile_flags[t] >> 1) & 1;
    }
}

// what one thread's 16 bytes hold, once the kind of the line it starts in is known
struct Counts {
    uint32_t headers, residues;
};
__device__ __forceinline__ Counts count_span(const uint8_t (&b)[kBytesPerThread], uint8_t before, uint32_t valid, uint32_t kind) {
    Counts c{0, 0};
    uint8_t prev = before;
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i) {
        if (static_cast<uint32_t>(i) < valid) {
            if (prev == '\n') {
                kind = b[i] == '>';
                c.headers += kind;
            }
            c.residues += (kind == 0 && b[i] != '\n');
        }
        prev = b[i];
    }
    return c;
}

// exclusive scan of (headers, residues) over the threads of a CTA; returns this thread's prefix, `total` the CTA's sum
__device__ __forceinline__ Counts block_exclusive_scan(Counts mine, Counts& total, uint32_t (*warp_sums)[2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t h = mine.headers, r = mine.residues;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t oh = __shfl_up_sync(0xffffffffu, h, d), orr = __shfl_up_sync(0xffffffffu, r, d);
        if (lane >= d) h += oh, r += orr;
    }
    if (lane == 31) warp_sums[warp][0] = h, warp_sums[warp][1] = r;
    __syncthreads();
    Counts prefix{h - mine.headers, r - mine.residues};
    total = Counts{0, 0};
    for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) prefix.headers += warp_sums[w][0], prefix.residues += warp_sums[w][1];
        total.headers += warp_sums[w][0];
        total.residues += warp_sums[w][1];
    }
    __syncthreads();
    return prefix;
}

// ---- pass C ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) fasta_tile_counts_kernel(const uint8_t* __restrict__ text, uint64_t bytes,
                                                                    const uint8_t* __restrict__ tile_entry, uint32_t* __restrict__ tile_headers,
                                                                    uint32_t* __restrict__ tile_residues) {
    __shared__ uint32_t warp_exit[kThreads / 32][2];
    __shared__ uint32_t warp_sums[kThreads / 32][2];
    uint8_t b[kBytesPerThread], before;
    uint32_t valid;
    load_span(text, bytes, static_cast<uint64_t>(blockIdx.x) * kTile + threadIdx.x * kBytesPerThread, b, before, valid);
    const uint32_t kind = entry_kind_in_tile(classify_span(b, before, valid), tile_entry[blockIdx.x], &warp_exit[0][0]);
    Counts total;
    block_exclusive_scan(count_span(b, before, valid, kind), total, warp_sums);
    if (threadIdx.x == 0) {
        tile_headers[blockIdx.x] = total.headers;
        tile_residues[blockIdx.x] = total.residues;
    }
}

// ---- pass D / F-b: exclusive scan of one or two uint32 arrays into uint64 bases, one CTA; totals[k] = sum of array k ------------
__global__ void __launch_bounds__(1024) scan_pair_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t n,
                                                         uint64_t* __restrict__ base_a, uint64_t* __restrict__ base_b, uint64_t* __restrict__ totals) {
    __shared__ uint64_t part[1024][2];
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t first = min(n, threadIdx.x * per), last = min(n, first + per);
    uint64_t sa = 0, sb = 0;
    for (uint32_t i = first; i < last; ++i) {
        sa += a[i];
        if (b) sb += b[i];
    }
    part[threadIdx.x][0] = sa;
    part[threadIdx.x][1] = sb;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t ra = 0, rb = 0;
        for (uint32_t t = 0; t < blockDim.x; ++t) {
            const uint64_t va = part[t][0], vb = part[t][1];
            part[t][0] = ra;
            part[t][1] = rb;
            ra += va;
            rb += vb;
        }
        totals[0] = ra;
        totals[1] = rb;
    }
    __syncthreads();
    uint64_t ra = part[threadIdx.x][0], rb = part[threadIdx.x][1];
    for (uint32_t i = first; i < last; ++i) {
        base_a[i] = ra;
        ra += a[i];
        if (b) {
            base_b[i] = rb;
            rb += b[i];
        }
    }
}

// ---- pass E ---------------------------------------------------------------------------------------------------------------
// record r >= 1 is the r-th header; record 0 is whatever precedes the first header.  record_start[r] = rank (among all
// residue bytes) of the record's first residue; record_bad[r] != 0 when it contains a foreign byte.
__global__ void __launch_bounds__(kThreads) fasta_encode_kernel(const uint8_t* __restrict__ text, uint64_t bytes,
                                                               const uint8_t* __restrict__ tile_entry, const uint64_t* __restrict__ tile_header_base,
                                                               const uint64_t* __restrict__ tile_residue_base, uint8_t* __restrict__ codes,
                                                               uint64_t* __restrict__ record_start, uint8_t* __restrict__ record_bad) {
    __shared__ uint32_t warp_exit[kThreads / 32][2];
    __shared__ uint32_t warp_sums[kThreads / 32][2];
    uint8_t b[kBytesPerThread], before;
    uint32_t valid;
    load_span(text, bytes, static_cast<uint64_t>(blockIdx.x) * kTile + threadIdx.x * kBytesPerThread, b, before, valid);
    uint32_t kind = entry_kind_in_tile(classify_span(b, before, valid), tile_entry[blockIdx.x], &warp_exit[0][0]);
    Counts total;
    const Counts prefix = block_exclusive_scan(count_span(b, before, valid, kind), total, warp_sums);
    uint64_t record = tile_header_base[blockIdx.x] + prefix.headers;
    uint64_t rank = tile_residue_base[blockIdx.x] + prefix.residues;
    uint8_t prev = before;
    bool bad = false; // a foreign byte in the record this thread is currently in
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i) {
        if (static_cast<uint32_t>(i) < valid) {
            if (prev == '\n') {
                kind = b[i] == '>';
                if (kind) {
                    if (bad) record_bad[record] = 1;
                    bad = false;
                    ++record;
                    record_start[record] = rank;
                }
            }
            if (kind == 0 && b[i] != '\n') {
                const uint32_t code = residue_code(b[i]);
                codes[rank++] = static_cast<uint8_t>(code);
                bad |= code == 0xffu;
            }
        }
        prev = b[i];
    }
    if (bad) record_bad[record] = 1;
}

// ---- pass F: per record length and keep flag; tile sums for the scan over records ------------------------------------------
__global__ void __launch_bounds__(kThreads) fasta_record_keep_kernel(const uint64_t* __restrict__ record_start, const uint8_t* __restrict__ record_bad,
                                                                    uint64_t records /* incl. record 0 */, uint64_t total_residues,
                                                                    uint32_t* __restrict__ tile_kept, uint32_t* __restrict__ tile_kept_residues,
                                                                    unsigned long long* __restrict__ longest) {
    __shared__ uint32_t warp_sums[kThreads / 32][2];
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
    Counts mine{0, 0};
    if (r >= 1 && r < records && !record_bad[r]) {
        const uint64_t len = (r + 1 < records ? record_start[r + 1] : total_residues) - record_start[r];
        mine = Counts{1, static_cast<uint32_t>(min(len, static_cast<uint64_t>(0xffffffffull)))};
        atomicMax(longest, static_cast<unsigned long long>(len));
    }
    Counts total;
    block_exclusive_scan(mine, total, warp_sums);
    if (threadIdx.x == 0) {
        tile_kept[blockIdx.x] = total.headers;
        tile_kept_residues[blockIdx.x] = total.residues;
    }
}

// offsets[k] = first residue of the k-th kept record in the final array; source[k] = where its codes sit after pass E
__global__ void __launch_bounds__(kThreads) fasta_record_place_kernel(const uint64_t* __restrict__ record_start, const uint8_t* __restrict__ record_bad,
                                                                     uint64_t records, uint64_t total_residues, const uint64_t* __restrict__ tile_kept_base,
                                                                     const uint64_t* __restrict__ tile_residue_base, uint64_t* __restrict__ offsets,
                                                                     uint64_t* __restrict__ source) {
    __shared__ uint32_t warp_sums[kThreads / 32][2];
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
    Counts mine{0, 0};
    uint64_t start = 0;
    if (r >= 1 && r < records && !record_bad[r]) {
        start = record_start[r];
        const uint64_t len = (r + 1 < records ? record_start[r + 1] : total_residues) - start;
        mine = Counts{1, static_cast<uint32_t>(len)};
    }
    Counts total;
    const Counts prefix = block_exclusive_scan(mine, total, warp_sums);
    if (mine.headers) {
        const uint64_t k = tile_kept_base[blockIdx.x] + prefix.headers;
        offsets[k] = tile_residue_base[blockIdx.x] + prefix.residues;
        source[k] = start;
    }
}

// ---- pass G: one warp per kept record moves its codes to their final place --------------------------------------------------
__global__ void __launch_bounds__(kThreads) fasta_compact_kernel(const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets,
                                                                const uint64_t* __restrict__ source, uint64_t kept, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = static_cast<uint64_t>(gridDim.x) * (kThreads / 32);
    for (uint64_t k = static_cast<uint64_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); k < kept; k += warps) {
        const uint64_t to = offsets[k], len = offsets[k + 1] - to, from = source[k];
        for (uint64_t i = lane; i < len; i += 32) out[to + i] = codes[from + i];
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
// Grow-only scratch of the parser, one per device, shared by every call in the process (calls are serialised by a mutex: the
// parse itself takes well under a millisecond per 100 MB, the scan that follows runs outside the lock).
struct Fasta_scratch {
    int device = -1;
    uint8_t* d_text = nullptr;
    size_t cap_text = 0;
    uint8_t* d_codes = nullptr;
    size_t cap_codes = 0;
    void* d_tiles = nullptr; // flags | entry | headers | residues | header_base | residue_base
    size_t cap_tiles = 0;
    void* d_records = nullptr; // start | source | bad | kept tile arrays
    size_t cap_records = 0;
    uint64_t* d_totals = nullptr; // [0..1] tiles, [2..3] records, [4] longest
    uint64_t* h_totals = nullptr; // pinned
    // staging ring for pageable text: `workers` host threads, two pinned buffers and one stream each
    static constexpr int kWorkers = 8;
    static constexpr size_t kStage = 8u << 20;
    uint8_t* h_stage[kWorkers][2] = {};
    cudaStream_t streams[kWorkers] = {};
    cudaEvent_t events[kWorkers][2] = {};
    cudaStream_t stream = nullptr;
    std::mutex lock;
};

Fasta_scratch* scratch_for(int device) {
    static std::mutex registry_lock;
    static std::vector<Fasta_scratch*> registry;
    std::lock_guard<std::mutex> guard(registry_lock);
    for (auto* s : registry)
        if (s->device == device) return s;
    auto* s = new (std::nothrow) Fasta_scratch();
    if (!s) return nullptr;
    s->device = device;
    registry.push_back(s);
    return s;
}

template <typename T> int grow(T*& pointer, size_t& capacity, size_t bytes) {
    if (bytes <= capacity) return MSV_OK;
    cudaFree(pointer);
    pointer = nullptr;
    capacity = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    MSV_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&pointer), want));
    capacity = want;
    return MSV_OK;
}

// text (host) -> s->d_text.  Pinned or device-accessible text goes in one copy; pageable text (an mmap'ed file) is staged by
// kWorkers threads through pinned buffers so that the page-cache reads, the staging copies and the DMA all overlap.
int upload_text(Fasta_scratch* s, const char* text, size_t bytes) {
    cudaPointerAttributes attributes{};
    const bool pinned = cudaPointerGetAttributes(&attributes, text) == cudaSuccess &&
                        (attributes.type == cudaMemoryTypeHost || attributes.type == cudaMemoryTypeManaged || attributes.type == cudaMemoryTypeDevice);
    (void)cudaGetLastError();
    if (pinned || bytes < (1u << 20)) {
        MSV_CUDA_TRY(cudaMemcpyAsync(s->d_text, text, bytes, cudaMemcpyDefault, s->stream));
        return MSV_OK;
    }
    if (!s->streams[0]) {
        for (int w = 0; w < Fasta_scratch::kWorkers; ++w) {
            MSV_CUDA_TRY(cudaStreamCreateWithFlags(&s->streams[w], cudaStreamNonBlocking));
            for (int k = 0; k < 2; ++k) {
                MSV_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_stage[w][k]), Fasta_scratch::kStage, cudaHostAllocPortable));
                MSV_CUDA_TRY(cudaEventCreateWithFlags(&s->events[w][k], cudaEventDisableTiming));
            }
        }
    }
    const size_t chunks = (bytes + Fasta_scratch::kStage - 1) / Fasta_scratch::kStage;
    const int workers = static_cast<int>(std::min<size_t>(Fasta_scratch::kWorkers, chunks));
    std::vector<cudaError_t> status(workers, cudaSuccess);
    const int device = s->device;
    const auto work = [&](int w) {
        cudaError_t err = cudaSetDevice(device);
        int turn = 0;
        for (size_t c = w; c < chunks && err == cudaSuccess; c += workers, turn ^= 1) {
            const size_t at = c * Fasta_scratch::kStage, len = std::min(Fasta_scratch::kStage, bytes - at);
            if (c >= static_cast<size_t>(2 * workers)) err = cudaEventSynchronize(s->events[w][turn]); // the buffer's previous copy is out
            if (err != cudaSuccess) break;
            std::memcpy(s->h_stage[w][turn], text + at, len);
            err = cudaMemcpyAsync(s->d_text + at, s->h_stage[w][turn], len, cudaMemcpyHostToDevice, s->streams[w]);
            if (err == cudaSuccess) err = cudaEventRecord(s->events[w][turn], s->streams[w]);
        }
        if (err == cudaSuccess) err = cudaStreamSynchronize(s->streams[w]);
        status[w] = err;
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < workers; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
    for (const cudaError_t err : status) MSV_CUDA_TRY(err);
    return MSV_OK;
}

} // namespace

namespace msv_detail {

// Fill `db` from FASTA text: upload, parse on the device, bucket longest-first.  Returns after the database is ready.
int db_fill_from_fasta(msv_db* db, const char* text, size_t bytes, size_t* rejected_out) {
    if (rejected_out) *rejected_out = 0;
    if (bytes && !text) return fail(MSV_ERR_INVALID_ARGUMENT, "text is NULL");
    if (bytes >= (1ull << 40)) return fail(MSV_ERR_INVALID_ARGUMENT, "FASTA text of %zu bytes is beyond the supported 1 TB", bytes);
    Fasta_scratch* s = scratch_for(db->device);
    if (!s) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    std::lock_guard<std::mutex> guard(s->lock);
    if (!s->stream) {
        MSV_CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        MSV_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_totals), 8 * sizeof(uint64_t)));
        MSV_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_totals), 8 * sizeof(uint64_t), cudaHostAllocPortable));
    }
    cudaStream_t stream = s->stream;
    const uint64_t tiles64 = (bytes + kTile - 1) / kTile;
    if (tiles64 >= (1ull << 31)) return fail(MSV_ERR_INVALID_ARGUMENT, "FASTA text too large");
    const uint32_t tiles = static_cast<uint32_t>(tiles64);
    uint64_t records = 1, residue_bytes = 0;
    if (bytes) {
        if (int rc = grow(s->d_text, s->cap_text, bytes + 64)) return rc;
        if (int rc = grow(s->d_codes, s->cap_codes, bytes + 64)) return rc;
        // per tile: flags u8 | entry u8 | headers u32 | residues u32 | header_base u64 | residue_base u64
        const size_t tile_bytes = static_cast<size_t>(tiles) * (1 + 1 + 4 + 4 + 8 + 8) + 64;
        uint8_t* raw = static_cast<uint8_t*>(s->d_tiles);
        if (int rc = grow(raw, s->cap_tiles, tile_bytes)) return rc;
        s->d_tiles = raw;
        const size_t t8 = (static_cast<size_t>(tiles) + 7) / 8 * 8;
        uint64_t* header_base = reinterpret_cast<uint64_t*>(raw);
        uint64_t* residue_base = header_base + tiles;
        uint32_t* tile_headers = reinterpret_cast<uint32_t*>(residue_base + tiles);
        uint32_t* tile_residues = tile_headers + tiles;
        uint8_t* tile_flags = reinterpret_cast<uint8_t*>(tile_residues + tiles);
        uint8_t* tile_entry = tile_flags + t8;

        if (int rc = upload_text(s, text, bytes)) return rc;
        fasta_tile_flags_kernel<<<tiles, kThreads, 0, stream>>>(s->d_text, bytes, tile_flags);
        fasta_tile_entry_kernel<<<1, 1024, 0, stream>>>(tile_flags, tiles, tile_entry);
        fasta_tile_counts_kernel<<<tiles, kThreads, 0, stream>>>(s->d_text, bytes, tile_entry, tile_headers, tile_residues);
        scan_pair_kernel<<<1, 1024, 0, stream>>>(tile_headers, tile_residues, tiles, header_base, residue_base, s->d_totals);
        for (int k = 0; k < 4; ++k) count_launch();
        MSV_CUDA_TRY(cudaGetLastError());
        MSV_CUDA_TRY(cudaMemcpyAsync(s->h_totals, s->d_totals, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        MSV_CUDA_TRY(cudaStreamSynchronize(stream));
        records = s->h_totals[0] + 1; // + record 0, the text before the first header
        residue_bytes = s->h_totals[1];

        // per record: start u64 | source u64 | bad u8; per record tile: kept u32 | kept_residues u32 | kept_base u64 | residue_base u64
        const uint32_t record_tiles = static_cast<uint32_t>((records + kThreads - 1) / kThreads);
        const size_t record_bytes = static_cast<size_t>(records + 1) * (8 + 8 + 1) + static_cast<size_t>(record_tiles) * (4 + 4 + 8 + 8) + 256;
        uint8_t* rraw = static_cast<uint8_t*>(s->d_records);
        if (int rc = grow(rraw, s->cap_records, record_bytes)) return rc;
        s->d_records = rraw;
        uint64_t* record_start = reinterpret_cast<uint64_t*>(rraw);
        uint64_t* source = record_start + (records + 1);
        uint64_t* kept_base = source + (records + 1);
        uint64_t* kept_residue_base = kept_base + record_tiles;
        uint32_t* tile_kept = reinterpret_cast<uint32_t*>(kept_residue_base + record_tiles);
        uint32_t* tile_kept_residues = tile_kept + record_tiles;
        uint8_t* record_bad = reinterpret_cast<uint8_t*>(tile_kept_residues + record_tiles);
        MSV_CUDA_TRY(cudaMemsetAsync(record_bad, 0, records + 1, stream));
        MSV_CUDA_TRY(cudaMemsetAsync(record_start, 0, sizeof(uint64_t), stream));
        MSV_CUDA_TRY(cudaMemsetAsync(s->d_totals + 4, 0, sizeof(uint64_t), stream));
        fasta_encode_kernel<<<tiles, kThreads, 0, stream>>>(s->d_text, bytes, tile_entry, header_base, residue_base, s->d_codes, record_start, record_bad);
        fasta_record_keep_kernel<<<record_tiles, kThreads, 0, stream>>>(record_start, record_bad, records, residue_bytes, tile_kept, tile_kept_residues,
                                                                        reinterpret_cast<unsigned long long*>(s->d_totals + 4));
        scan_pair_kernel<<<1, 1024, 0, stream>>>(tile_kept, tile_kept_residues, record_tiles, kept_base, kept_residue_base, s->d_totals + 2);
        for (int k = 0; k < 3; ++k) count_launch();
        MSV_CUDA_TRY(cudaGetLastError());
        MSV_CUDA_TRY(cudaMemcpyAsync(s->h_totals + 2, s->d_totals + 2, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        MSV_CUDA_TRY(cudaStreamSynchronize(stream));
        const uint64_t kept = s->h_totals[2], kept_residues = s->h_totals[3], longest = s->h_totals[4];
        if (kept >= (1ull << 32) - 1) return fail(MSV_ERR_INVALID_ARGUMENT, "more than 2^32-2 sequences in one database");
        if (longest >= (1ull << 27)) return fail(MSV_ERR_INVALID_ARGUMENT, "sequence of %llu residues exceeds the supported 2^27-1", static_cast<unsigned long long>(longest));
        if (int rc = db_reserve_for(db, kept_residues, kept, longest, stream)) return rc;
        fasta_record_place_kernel<<<record_tiles, kThreads, 0, stream>>>(record_start, record_bad, records, residue_bytes, kept_base, kept_residue_base,
                                                                         db->d_offsets, source);
        MSV_CUDA_TRY(cudaMemcpyAsync(db->d_offsets + kept, s->d_totals + 3, sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
        count_launch();
        if (kept_residues == residue_bytes) { // nothing dropped: the codes of the encode pass are the database as they are
            if (kept_residues) MSV_CUDA_TRY(cudaMemcpyAsync(db->d_residues, s->d_codes, kept_residues, cudaMemcpyDeviceToDevice, stream));
        } else if (kept) {
            const int blocks = static_cast<int>(std::min<uint64_t>((kept + kThreads / 32 - 1) / (kThreads / 32), 148 * 32));
            fasta_compact_kernel<<<blocks, kThreads, 0, stream>>>(s->d_codes, db->d_offsets, source, kept, db->d_residues);
            count_launch();
        }
        MSV_CUDA_TRY(cudaMemsetAsync(db->d_residues + kept_residues, 0, 64 + 16, stream));
        MSV_CUDA_TRY(cudaGetLastError());
        db->n = kept;
        db->total = kept_residues;
        db->longest = longest;
        db->h_lengths.clear();
        if (rejected_out) *rejected_out = static_cast<size_t>(records - 1 - kept);
        if (int rc = db_bucket(db, stream)) return rc;
        MSV_CUDA_TRY(cudaStreamSynchronize(stream));
        return MSV_OK;
    }
    // empty text: an empty database
    if (int rc = db_reserve_for(db, 0, 0, 0, stream)) return rc;
    MSV_CUDA_TRY(cudaMemsetAsync(db->d_offsets, 0, sizeof(uint64_t), stream));
    MSV_CUDA_TRY(cudaMemsetAsync(db->d_residues, 0, 64 + 16, stream));
    MSV_CUDA_TRY(cudaStreamSynchronize(stream));
    db->n = 0;
    db->total = 0;
    db->longest = 0;
    db->h_lengths.clear();
    return MSV_OK;
}

} // namespace msv_detail

extern "C" {

int msv_cuda_db_create_from_fasta(int device, const char* text, size_t bytes, msv_db** out, size_t* rejected) {
    if (!out) return fail(MSV_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    int count = 0;
    if (int rc = msv_cuda_device_count(&count)) return rc;
    if (device < 0 || device >= count) return fail(MSV_ERR_INVALID_ARGUMENT, "device %d out of range (have %d)", device, count);
    Device_guard guard(device);
    MSV_CUDA_TRY(guard.status);
    auto* db = new (std::nothrow) msv_db();
    if (!db) return fail(MSV_ERR_OUT_OF_MEMORY, "host allocation failed");
    db->device = device;
    const int rc = msv_detail::db_fill_from_fasta(db, text, bytes, rejected);
    if (rc != MSV_OK) {
        msv_detail::db_free(db);
        return rc;
    }
    *out = db;
    return MSV_OK;
}

int msv_cuda_db_refill_from_fasta(msv_db* db, const char* text, size_t bytes, size_t* rejected) {
    if (!db) return fail(MSV_ERR_INVALID_ARGUMENT, "db is NULL");
    Device_guard guard(db->device);
    MSV_CUDA_TRY(guard.status);
    const int rc = msv_detail::db_fill_from_fasta(db, text, bytes, rejected);
    if (rc != MSV_OK) db->n = 0;
    return rc;
}

int msv_cuda_db_download(const msv_db* db, uint8_t* residues, uint64_t* offsets) {
    if (!db) return fail(MSV_ERR_INVALID_ARGUMENT, "db is NULL");
    Device_guard guard(db->device);
    MSV_CUDA_TRY(guard.status);
    if (residues && db->total) MSV_CUDA_TRY(cudaMemcpy(residues, db->d_residues, db->total, cudaMemcpyDeviceToHost));
    if (offsets) {
        if (db->n) MSV_CUDA_TRY(cudaMemcpy(offsets, db->d_offsets, (db->n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        else offsets[0] = 0;
    }
    return MSV_OK;
}

} // extern "C"
