#include "Synthetic_database.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <random>

namespace {

// HMMER's default amino-acid background (order A C D E F G H I K L M N P Q R S T V W Y).
constexpr std::array<double, 20> background = {0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198,
                                               0.0590092, 0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639,
                                               0.0540978, 0.0683364, 0.0540687, 0.0673417, 0.0114135, 0.0304133};

// 65536-entry inverse CDF: a 16-bit uniform number -> residue code.
std::vector<uint8_t> inverse_cdf(const std::array<double, 20>& weights) {
    auto total = 0.0;
    for (const auto w : weights) total += w;
    auto lut = std::vector<uint8_t>(65536);
    auto cumulative = 0.0;
    auto at = size_t(0);
    for (size_t code = 0; code < weights.size(); ++code) {
        cumulative += weights[code] / total;
        const auto upto = code + 1 == weights.size() ? lut.size() : static_cast<size_t>(std::llround(cumulative * 65536.0));
        for (; at < upto && at < lut.size(); ++at) lut[at] = static_cast<uint8_t>(code);
    }
    return lut;
}

void fill_residues(Packed_sequences& db, const std::vector<size_t>& lengths, const std::vector<uint8_t>& lut,
                   std::mt19937_64& engine) {
    auto total = size_t(0);
    db.offsets.assign(1, 0);
    db.offsets.reserve(lengths.size() + 1);
    for (const auto len : lengths) {
        total += len;
        db.offsets.push_back(total);
    }
    db.residues.resize(total);
    auto* out = db.residues.data();
    auto i = size_t(0);
    for (; i + 4 <= total; i += 4) { // four residues per 64-bit draw
        const auto bits = engine();
        out[i] = lut[bits & 0xffff];
        out[i + 1] = lut[(bits >> 16) & 0xffff];
        out[i + 2] = lut[(bits >> 32) & 0xffff];
        out[i + 3] = lut[bits >> 48];
    }
    if (i < total) {
        auto bits = engine();
        for (; i < total; ++i, bits >>= 16) out[i] = lut[bits & 0xffff];
    }
}

} // namespace

Packed_sequences synthetic_swissprot_like(size_t count, uint64_t seed) {
    auto engine = std::mt19937_64(seed);
    auto log_length = std::normal_distribution<double>(5.70, 0.55);
    auto lengths = std::vector<size_t>(count);
    for (auto& len : lengths) len = static_cast<size_t>(std::clamp(std::llround(std::exp(log_length(engine))), 30LL, 3000LL));
    auto db = Packed_sequences();
    fill_residues(db, lengths, inverse_cdf(background), engine);
    return db;
}

Packed_sequences synthetic_long_uniform(size_t count, uint64_t seed, size_t shortest, size_t longest) {
    auto engine = std::mt19937_64(seed);
    auto length = std::uniform_int_distribution<size_t>(shortest, longest);
    auto lengths = std::vector<size_t>(count);
    for (auto& len : lengths) len = length(engine);
    auto uniform = std::array<double, 20>();
    uniform.fill(1.0);
    auto db = Packed_sequences();
    fill_residues(db, lengths, inverse_cdf(uniform), engine);
    return db;
}
