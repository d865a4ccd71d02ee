// TEST INFRASTRUCTURE ONLY -- the synthetic databases of BASELINE.json configs 3-5, generated WITHOUT the product
// libraries, so that bench.py's `--impl reference` arm (the reference's own CPU code on the host cores) and the CPU
// tests can build the same workload without loading libmsv_host.so / libmsv_cuda.so.
//
// Same recipe as SURVEY.md section 8(d): lengths L = clip(round(exp(N(5.70, 0.55^2))), 30, 3000), residues i.i.d. from
// the background frequencies the reference hard-codes (algorithms/MSV_HMM.cpp:21-27), std::mt19937_64(seed).  The
// draws are made in the same order as the product's generator (host/data_readers/Synthetic_database.cpp), and
// tests/test_host_cpu.py checks that both produce the same bytes.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

namespace {

constexpr std::array<double, 20> k_background = {0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198,
                                                 0.0590092, 0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639,
                                                 0.0540978, 0.0683364, 0.0540687, 0.0673417, 0.0114135, 0.0304133};

// 16-bit uniform number -> residue code (inverse CDF sampled at 65536 points)
std::vector<uint8_t> residue_lookup(const std::array<double, 20>& weights) {
    double sum = 0.0;
    for (double w : weights) sum += w;
    std::vector<uint8_t> lut(65536);
    double running = 0.0;
    size_t filled = 0;
    for (size_t code = 0; code < 20; ++code) {
        running += weights[code] / sum;
        const size_t end = code == 19 ? lut.size() : static_cast<size_t>(std::llround(running * 65536.0));
        while (filled < end && filled < lut.size()) lut[filled++] = static_cast<uint8_t>(code);
    }
    return lut;
}

struct Synthetic {
    std::vector<uint8_t> residues;
    std::vector<uint64_t> offsets;
};

void draw_residues(Synthetic& db, const std::vector<uint64_t>& lengths, const std::vector<uint8_t>& lut, std::mt19937_64& rng) {
    db.offsets.assign(1, 0);
    for (uint64_t len : lengths) db.offsets.push_back(db.offsets.back() + len);
    const size_t total = db.offsets.back();
    db.residues.resize(total);
    size_t at = 0;
    while (at + 4 <= total) { // four residues per 64-bit draw, low 16 bits first
        uint64_t bits = rng();
        for (int k = 0; k < 4; ++k, bits >>= 16) db.residues[at++] = lut[bits & 0xffff];
    }
    if (at < total) {
        uint64_t bits = rng();
        for (; at < total; bits >>= 16) db.residues[at++] = lut[bits & 0xffff];
    }
}

} // namespace

extern "C" {

void* oracle_synthetic_swissprot_like(size_t count, uint64_t seed) {
    auto* db = new Synthetic();
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> log_length(5.70, 0.55);
    std::vector<uint64_t> lengths(count);
    for (auto& len : lengths) len = static_cast<uint64_t>(std::clamp(std::llround(std::exp(log_length(rng))), 30LL, 3000LL));
    draw_residues(*db, lengths, residue_lookup(k_background), rng);
    return db;
}

void* oracle_synthetic_long_uniform(size_t count, uint64_t seed, size_t shortest, size_t longest) {
    auto* db = new Synthetic();
    std::mt19937_64 rng(seed);
    std::uniform_int_distribution<size_t> length(shortest, longest);
    std::vector<uint64_t> lengths(count);
    for (auto& len : lengths) len = length(rng);
    std::array<double, 20> flat;
    flat.fill(1.0);
    draw_residues(*db, lengths, residue_lookup(flat), rng);
    return db;
}

size_t oracle_synthetic_count(void* h) { return static_cast<Synthetic*>(h)->offsets.size() - 1; }
const uint8_t* oracle_synthetic_residues(void* h) { return static_cast<Synthetic*>(h)->residues.data(); }
const uint64_t* oracle_synthetic_offsets(void* h) { return static_cast<Synthetic*>(h)->offsets.data(); }
void oracle_synthetic_free(void* h) { delete static_cast<Synthetic*>(h); }

} // extern "C"
