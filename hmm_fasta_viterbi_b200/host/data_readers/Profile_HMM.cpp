#include "Profile_HMM.hpp"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <algorithm>

// Parsing contract (behaviour of the reference's data_readers/Profile_HMM.cpp:8-122, restated):
//   * the file is scanned strictly forward: NAME, LENG, three STATS lines, COMPO, then nodes 1..LENG;
//   * a "tag" matches a line when the line's first non-blank characters START WITH the tag (prefix match), so
//     node 1 is found by the first line beginning with "1" after COMPO -- the order of the file does the rest;
//   * the value of a tagged line is everything after its first word;
//   * a probability field is stored as expf(-strtof(field)); "*" does not parse as a number, strtof yields 0 and
//     the stored probability is therefore 1.0 (pinned by the reference's own test, test_hmm_parsing.cpp:35-36);
//   * model_length = LENG + 1.
// Unlike the reference this reader loads the file once and walks it with string_views (no per-line allocation), and
// blank lines are skipped instead of being undefined behaviour.

namespace {

constexpr std::string_view blanks = " \t\r";

// Forward-only cursor over the lines of a text buffer.
class Line_cursor {
  public:
    explicit Line_cursor(std::string_view text) : rest(text) {}

    bool next(std::string_view& line) {
        if (rest.empty()) return false;
        const auto eol = rest.find('\n');
        line = rest.substr(0, eol);
        rest.remove_prefix(eol == std::string_view::npos ? rest.size() : eol + 1);
        return true;
    }

    // Advance to the first line that starts (after blanks) with `tag`; yield what follows that line's first word.
    bool seek(std::string_view tag, std::string_view& value) {
        auto line = std::string_view();
        while (next(line)) {
            const auto start = line.find_first_not_of(' ');
            if (start == std::string_view::npos) continue;
            line.remove_prefix(start);
            if (line.substr(0, tag.size()) == tag) {
                value = after_word(line);
                return true;
            }
        }
        return false;
    }

    static std::string_view after_word(std::string_view s) {
        const auto word_end = s.find(' ');
        if (word_end == std::string_view::npos) return {};
        s.remove_prefix(word_end);
        const auto next_word = s.find_first_not_of(' ');
        return next_word == std::string_view::npos ? std::string_view() : s.substr(next_word);
    }

  private:
    std::string_view rest;
};

// strtof needs a terminated buffer; fields are short, so copy the field.
float field_to_float(std::string_view& fields) {
    const auto start = fields.find_first_not_of(blanks);
    if (start == std::string_view::npos) {
        fields = {};
        return 0.0f;
    }
    fields.remove_prefix(start);
    const auto len = std::min(fields.find_first_of(blanks), fields.size());
    char buffer[64];
    const auto n = std::min(len, sizeof buffer - 1);
    fields.copy(buffer, n);
    buffer[n] = '\0';
    fields.remove_prefix(len);
    return std::strtof(buffer, nullptr);
}

template <int N> Probabilities_array<N> negative_logs_to_probabilities(std::string_view fields) {
    auto row = Probabilities_array<N>();
    for (auto& p : row) p = std::exp(-1 * field_to_float(fields));
    return row;
}

} // namespace

Profile_HMM::Profile_HMM(const std::string& file_path) {
    auto file = std::ifstream(file_path, std::ios::binary);
    if (file.fail()) {
        std::cout << "Failed to open " << file_path << '\n';
        return;
    }
    file.seekg(0, std::ios::end);
    auto text = std::string(static_cast<size_t>(std::max<std::streamoff>(file.tellg(), 0)), '\0');
    file.seekg(0, std::ios::beg);
    file.read(text.data(), static_cast<std::streamsize>(text.size()));
    if (!parse(text)) std::cout << "Incomplete profile HMM in " << file_path << '\n';
}

bool Profile_HMM::parse(std::string_view text) {
    auto lines = Line_cursor(text);
    auto value = std::string_view();

    if (!lines.seek("NAME", value)) return false;
    name = Profile_name(value);

    if (!lines.seek("LENG", value)) return false;
    model_length = static_cast<size_t>(std::atoi(std::string(value).c_str())) + 1;

    for (int i = 0; i < 3; ++i) {
        if (!lines.seek("STATS", value)) return false;
        value = Line_cursor::after_word(value); // drop LOCAL
        const auto kind = value.empty() ? '\0' : value.front();
        auto numbers = Line_cursor::after_word(value);
        const auto first = field_to_float(numbers);
        const auto second = field_to_float(numbers);
        if (kind == 'M') {
            stats_local_msv_mu = first;
            stats_local_msv_lambda = second;
        } else if (kind == 'V') {
            stats_local_viterbi_mu = first;
            stats_local_viterbi_lambda = second;
        } else if (kind == 'F') {
            stats_local_forward_theta = first;
            stats_local_forward_lambda = second;
        }
    }

    if (!lines.seek("COMPO", value)) return false;
    match_emissions.reserve(model_length);
    insert_emissions.reserve(model_length);
    transitions.reserve(model_length);

    // begin node: no match emissions; its insert emissions and transitions follow COMPO
    auto line = std::string_view();
    match_emissions.emplace_back();
    if (!lines.next(line)) return false;
    insert_emissions.push_back(negative_logs_to_probabilities<NUM_OF_AMINO_ACIDS>(line));
    if (!lines.next(line)) return false;
    transitions.push_back(negative_logs_to_probabilities<NUM_OF_TRANSITIONS>(line));

    for (size_t node = 1; node < model_length; ++node) {
        if (!lines.seek(std::to_string(node), value)) return false;
        match_emissions.push_back(negative_logs_to_probabilities<NUM_OF_AMINO_ACIDS>(value));
        if (!lines.next(line)) return false;
        insert_emissions.push_back(negative_logs_to_probabilities<NUM_OF_AMINO_ACIDS>(line));
        if (!lines.next(line)) return false;
        transitions.push_back(negative_logs_to_probabilities<NUM_OF_TRANSITIONS>(line));
    }
    return true;
}
