/*
 * msv_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product path.
 *
 * A plain-C, single-file CPU restatement of the reference's MSV path, used as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs (and nowhere else).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit against
 *   (1) tests/golden/msv_scores.json -- 24 models x 7 sequences of IEEE-754 bit patterns produced by the
 *       reference's own unmodified sources compiled here (oracle/Makefile -> oracle/_ref/libmsv_ref.so), and
 *   (2) oracle/_ref itself whenever that library is present, on random inputs,
 * and against the reader KATs the reference's own tests hold (data_readers/test_hmm_parsing.cpp:23-36,
 * data_readers/test_fasta_parsing.cpp:8-14).
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.  The code is written from
 * the behaviour of those lines, not copied from them: C instead of C++, two rolling rows instead of the
 * reference's full (L+1) x (model_length+5) matrix, explicit error returns instead of exceptions.
 *
 * Floating point: all arithmetic is IEEE binary32; every transcendental is glibc logf/expf evaluated at run time
 * on the host, exactly as in the reference (Profile_HMM.cpp:40, MSV_HMM.cpp:42,51-53,62-63).  Build with
 * -fno-fast-math (default) so that no contraction or reassociation happens; there are no multiplies in the
 * recurrence, so FMA contraction cannot occur either.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_ALPHABET 20
#define ORACLE_TRANSITIONS 7

/* ------------------------------------------------------------------------------------------------------------ */
/* Constants: MSV_HMM.cpp:21-27 (background frequencies, HMMER p7_AminoFrequencies order A C D E ... Y) and      */
/* MSV_HMM.cpp:29-31 (residue letter -> column index).                                                          */
/* ------------------------------------------------------------------------------------------------------------ */
static const float oracle_background[ORACLE_ALPHABET] = {
    0.0787945f, 0.0151600f, 0.0535222f, 0.0668298f, 0.0397062f, 0.0695071f, 0.0229198f, 0.0590092f, 0.0594422f, 0.0963728f,
    0.0237718f, 0.0414386f, 0.0482904f, 0.0395639f, 0.0540978f, 0.0683364f, 0.0540687f, 0.0673417f, 0.0114135f, 0.0304133f};

static const char oracle_letters[ORACLE_ALPHABET + 1] = "ACDEFGHIKLMNPQRSTVWY";

int oracle_residue_code(char c) {
    const char* p = (c != '\0') ? strchr(oracle_letters, c) : NULL;
    return p ? (int)(p - oracle_letters) : -1;
}

const float* oracle_background_frequencies(void) { return oracle_background; }

/* ------------------------------------------------------------------------------------------------------------ */
/* .hmm reader: data_readers/Profile_HMM.cpp:8-122                                                               */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct oracle_hmm {
    size_t model_length;    /* LENG + 1 (dummy node 0), Profile_HMM.cpp:70 */
    char name[512];         /* NAME value, Profile_HMM.cpp:63 */
    float* match_emissions; /* [model_length][20]; row 0 all zero, Profile_HMM.cpp:110-111 */
    float* insert_emissions; /* [model_length][20] */
    float* transitions;     /* [model_length][7] */
    float stats[6];         /* msv mu, msv lambda, viterbi mu, viterbi lambda, forward theta, forward lambda */
} oracle_hmm;

/* Profile_HMM.cpp:8-13 -- drop the current word and the blanks after it, `times` times. */
static const char* skip_words(const char* s, int times) {
    for (int t = 0; t < times; ++t) {
        while (*s != '\0' && *s != ' ') ++s;
        while (*s == ' ') ++s;
    }
    return s;
}

/* getline that strips the '\n' (std::getline semantics); returns 0 at EOF with nothing read. */
static int read_line(FILE* f, char** buf, size_t* cap) {
    ssize_t got = getline(buf, cap, f);
    if (got < 0) return 0;
    if (got > 0 && (*buf)[got - 1] == '\n') (*buf)[got - 1] = '\0';
    return 1;
}

/* Profile_HMM.cpp:15-26 -- scan forward for the first line whose first non-blank text starts with `tag`
 * (a PREFIX match, not a token match); return what follows the first word of that line. */
static const char* value_after_tag(FILE* f, const char* tag, char** buf, size_t* cap) {
    const size_t tag_len = strlen(tag);
    while (read_line(f, buf, cap)) {
        const char* s = *buf;
        while (*s == ' ') ++s;
        if (strncmp(s, tag, tag_len) == 0) return skip_words(s, 1);
    }
    return NULL;
}

/* Profile_HMM.cpp:35-45 -- N fields, each stored as expf(-strtof(field)); "*" parses as 0 and so becomes 1.0. */
static void parse_probabilities(const char* s, int n, float* out) {
    while (*s == ' ') ++s;
    for (int i = 0; i < n; ++i) {
        out[i] = expf(-1 * strtof(s, NULL));
        s = skip_words(s, 1);
    }
}

void oracle_hmm_free(oracle_hmm* h) {
    if (!h) return;
    free(h->match_emissions);
    free(h->insert_emissions);
    free(h->transitions);
    free(h);
}

/* Profile_HMM.cpp:48-60 (ctor), :62-71 (NAME, LENG), :73-93 (three STATS LOCAL lines), :95-122 (COMPO + nodes). */
oracle_hmm* oracle_hmm_load(const char* path) {
    FILE* f = fopen(path, "r");
    if (!f) return NULL;
    oracle_hmm* h = (oracle_hmm*)calloc(1, sizeof(oracle_hmm));
    char* buf = NULL;
    size_t cap = 0;
    const char* v;
    int ok = 0;

    if (!(v = value_after_tag(f, "NAME", &buf, &cap))) goto done;
    snprintf(h->name, sizeof h->name, "%s", v);
    if (!(v = value_after_tag(f, "LENG", &buf, &cap))) goto done;
    h->model_length = (size_t)atoi(v) + 1;

    for (int i = 0; i < 3; ++i) {
        if (!(v = value_after_tag(f, "STATS", &buf, &cap))) goto done;
        v = skip_words(v, 1); /* past LOCAL */
        int slot = (v[0] == 'M') ? 0 : (v[0] == 'V') ? 2 : (v[0] == 'F') ? 4 : -1;
        if (slot >= 0) {
            char* rest = NULL;
            const char* nums = skip_words(v, 1);
            h->stats[slot] = strtof(nums, &rest);
            h->stats[slot + 1] = strtof(rest, NULL);
        }
    }

    if (!value_after_tag(f, "COMPO", &buf, &cap)) goto done;
    h->match_emissions = (float*)calloc(h->model_length * ORACLE_ALPHABET, sizeof(float));
    h->insert_emissions = (float*)calloc(h->model_length * ORACLE_ALPHABET, sizeof(float));
    h->transitions = (float*)calloc(h->model_length * ORACLE_TRANSITIONS, sizeof(float));
    if (!read_line(f, &buf, &cap)) goto done;
    parse_probabilities(buf, ORACLE_ALPHABET, h->insert_emissions);
    if (!read_line(f, &buf, &cap)) goto done;
    parse_probabilities(buf, ORACLE_TRANSITIONS, h->transitions);

    for (size_t node = 1; node < h->model_length; ++node) {
        char tag[32];
        snprintf(tag, sizeof tag, "%zu", node);
        if (!(v = value_after_tag(f, tag, &buf, &cap))) goto done;
        parse_probabilities(v, ORACLE_ALPHABET, h->match_emissions + node * ORACLE_ALPHABET);
        if (!read_line(f, &buf, &cap)) goto done;
        parse_probabilities(buf, ORACLE_ALPHABET, h->insert_emissions + node * ORACLE_ALPHABET);
        if (!read_line(f, &buf, &cap)) goto done;
        parse_probabilities(buf, ORACLE_TRANSITIONS, h->transitions + node * ORACLE_TRANSITIONS);
    }
    ok = 1;
done:
    free(buf);
    fclose(f);
    if (!ok) {
        oracle_hmm_free(h);
        return NULL;
    }
    return h;
}

size_t oracle_hmm_model_length(const oracle_hmm* h) { return h->model_length; }
const char* oracle_hmm_name(const oracle_hmm* h) { return h->name; }
const float* oracle_hmm_match(const oracle_hmm* h) { return h->match_emissions; }
const float* oracle_hmm_insert(const oracle_hmm* h) { return h->insert_emissions; }
const float* oracle_hmm_transitions(const oracle_hmm* h) { return h->transitions; }
const float* oracle_hmm_stats(const oracle_hmm* h) { return h->stats; }

/* ------------------------------------------------------------------------------------------------------------ */
/* FASTA reader: data_readers/FASTA_protein_sequences.cpp:9-44                                                   */
/* Each '>' line opens a record "#"; other lines are appended verbatim; afterwards every record holding a        */
/* character outside "#ACDEFGHIKLMNPQRSTVWY" is dropped whole.                                                   */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct oracle_fasta {
    size_t count;
    char** records; /* each NUL-terminated, with the leading '#' */
} oracle_fasta;

void oracle_fasta_free(oracle_fasta* fa) {
    if (!fa) return;
    for (size_t i = 0; i < fa->count; ++i) free(fa->records[i]);
    free(fa->records);
    free(fa);
}

oracle_fasta* oracle_fasta_load(const char* path) {
    FILE* f = fopen(path, "r");
    if (!f) return NULL;
    oracle_fasta* fa = (oracle_fasta*)calloc(1, sizeof(oracle_fasta));
    size_t cap_records = 0, cur_len = 0, cur_cap = 0;
    char* buf = NULL;
    size_t cap = 0;
    while (read_line(f, &buf, &cap)) {
        if (buf[0] == '>') {
            if (fa->count == cap_records) {
                cap_records = cap_records ? cap_records * 2 : 16;
                fa->records = (char**)realloc(fa->records, cap_records * sizeof(char*));
            }
            cur_cap = 64;
            cur_len = 1;
            fa->records[fa->count] = (char*)malloc(cur_cap);
            memcpy(fa->records[fa->count], "#", 2);
            ++fa->count;
        } else if (fa->count > 0) { /* the reference has undefined behaviour when the file does not start with '>' */
            size_t add = strlen(buf);
            char** rec = &fa->records[fa->count - 1];
            if (cur_len + add + 1 > cur_cap) {
                while (cur_len + add + 1 > cur_cap) cur_cap *= 2;
                *rec = (char*)realloc(*rec, cur_cap);
            }
            memcpy(*rec + cur_len, buf, add + 1);
            cur_len += add;
        }
    }
    free(buf);
    fclose(f);
    /* FASTA_protein_sequences.cpp:26-41 -- reject whole records with a foreign character, keep order of the rest */
    size_t kept = 0;
    for (size_t i = 0; i < fa->count; ++i) {
        int good = 1;
        for (const char* p = fa->records[i]; *p; ++p)
            if (*p != '#' && oracle_residue_code(*p) < 0) {
                good = 0;
                break;
            }
        if (good)
            fa->records[kept++] = fa->records[i];
        else
            free(fa->records[i]);
    }
    fa->count = kept;
    return fa;
}

size_t oracle_fasta_count(const oracle_fasta* fa) { return fa->count; }
const char* oracle_fasta_record(const oracle_fasta* fa, size_t i) { return fa->records[i]; }

/* ------------------------------------------------------------------------------------------------------------ */
/* Model preparation: MSV_HMM.cpp:35-57                                                                          */
/*   emission[j][i] = logf(match[i][j] / background[j]),  j residue, i column, row stride model_length (:38-45)  */
/*   tr_B_Mk = logf(2 / float(model_length * (model_length + 1)))  with model_length = LENG + 1 (:51)            */
/*   tr_E_C = logf((nu - 1) / nu), tr_E_J = logf(1 / nu), nu = 2 (:49,52-53)                                     */
/* ------------------------------------------------------------------------------------------------------------ */
void oracle_msv_prepare(const float* match_emissions, size_t model_length, float* emission_scores, float* transitions3) {
    for (size_t i = 0; i < model_length; ++i)
        for (size_t j = 0; j < ORACLE_ALPHABET; ++j)
            emission_scores[j * model_length + i] = logf(match_emissions[i * ORACLE_ALPHABET + j] / oracle_background[j]);
    volatile float nu = 2.0f; /* volatile: evaluate with the run-time libm like everything else */
    transitions3[0] = logf(2.0f / (float)(model_length * (model_length + 1)));
    transitions3[1] = logf((nu - 1.0f) / nu);
    transitions3[2] = logf(1.0f / nu);
}

/* MSV_HMM.cpp:59-64 -- length-dependent N/C/J loop and move scores; `residues` excludes the '#' sentinel. */
void oracle_msv_length_transitions(size_t residues, float* tr_loop, float* tr_move) {
    *tr_loop = logf((float)residues / (float)(residues + 3));
    *tr_move = logf(3 / (float)(residues + 3));
}

/* ------------------------------------------------------------------------------------------------------------ */
/* The recurrence: MSV_HMM.cpp:74-113 (two rolling rows, as the reference's own device path keeps, :291-292,422) */
/* codes[0..L) are residue indices 0..19.  Returns C[L] + tr_move.                                               */
/* ------------------------------------------------------------------------------------------------------------ */
static inline float max2(float a, float b) { return (a < b) ? b : a; } /* std::max semantics */

float oracle_msv_score_codes(const float* emission_scores, size_t model_length, const float* transitions3,
                             const uint8_t* codes, size_t L, float* scratch /* 2*model_length floats or NULL */) {
    const float tr_B_Mk = transitions3[0], tr_E_C = transitions3[1], tr_E_J = transitions3[2];
    float tr_loop, tr_move;
    oracle_msv_length_transitions(L, &tr_loop, &tr_move);

    float* own = NULL;
    if (!scratch) scratch = own = (float*)malloc(2 * model_length * sizeof(float));
    float* prev = scratch;
    float* cur = scratch + model_length;
    for (size_t k = 0; k < model_length; ++k) prev[k] = cur[k] = -INFINITY; /* :86 */

    float J = -INFINITY, C = -INFINITY, N = 0.0f, B = tr_move; /* :96-97 */
    for (size_t i = 0; i < L; ++i) {
        const float* e = emission_scores + (size_t)codes[i] * model_length; /* :101 */
        const float entry = B + tr_B_Mk;
        float E = -INFINITY;
        for (size_t k = 1; k < model_length; ++k) { /* :102-105; column 0 stays -inf */
            const float m = e[k] + max2(prev[k - 1], entry);
            cur[k] = m;
            E = max2(E, m);
        }
        J = max2(J + tr_loop, E + tr_E_J); /* :107 */
        C = max2(C + tr_loop, E + tr_E_C); /* :108 */
        N = N + tr_loop;                   /* :109 */
        B = max2(N + tr_move, J + tr_move); /* :110 */
        float* t = prev;
        prev = cur;
        cur = t;
    }
    free(own);
    return C + tr_move; /* :112 */
}

/* Same, from the reference's string form ("#" + letters).  Returns 0 and sets *score, or -1 on a foreign letter
 * (the reference throws std::out_of_range from unordered_map::at, MSV_HMM.cpp:101). */
int oracle_msv_score_string(const float* emission_scores, size_t model_length, const float* transitions3,
                            const char* seq_with_sentinel, float* score) {
    const size_t n = strlen(seq_with_sentinel);
    const size_t L = n ? n - 1 : 0;
    uint8_t* codes = (uint8_t*)malloc(L ? L : 1);
    for (size_t i = 0; i < L; ++i) {
        int c = oracle_residue_code(seq_with_sentinel[i + 1]);
        if (c < 0) {
            free(codes);
            return -1;
        }
        codes[i] = (uint8_t)c;
    }
    *score = oracle_msv_score_codes(emission_scores, model_length, transitions3, codes, L, NULL);
    free(codes);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Batch driver over a packed database (codes + offsets), used by the parity tests and by the "port" CPU         */
/* baseline in bench.py.  Threads take contiguous slices of sequences, as BASELINE.md section 4 prescribes.      */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const float* emis;
    size_t model_length;
    const float* tr3;
    const uint8_t* codes;
    const uint64_t* offsets;
    size_t begin, end;
    float* out;
} oracle_slice;

static void* oracle_slice_main(void* arg) {
    oracle_slice* s = (oracle_slice*)arg;
    float* scratch = (float*)malloc(2 * s->model_length * sizeof(float));
    for (size_t q = s->begin; q < s->end; ++q)
        s->out[q] = oracle_msv_score_codes(s->emis, s->model_length, s->tr3, s->codes + s->offsets[q],
                                           (size_t)(s->offsets[q + 1] - s->offsets[q]), scratch);
    free(scratch);
    return NULL;
}

void oracle_msv_score_batch(const float* emission_scores, size_t model_length, const float* transitions3,
                            const uint8_t* codes, const uint64_t* offsets, size_t n, float* scores, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n && n > 0) threads = (int)n;
    oracle_slice* sl = (oracle_slice*)calloc((size_t)threads, sizeof(oracle_slice));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    /* slices balanced by residue count */
    const uint64_t total = n ? offsets[n] - offsets[0] : 0;
    size_t at = 0;
    for (int t = 0; t < threads; ++t) {
        size_t end = at;
        const uint64_t want = offsets[0] + (total * (uint64_t)(t + 1)) / (uint64_t)threads;
        while (end < n && offsets[end + 1] <= want) ++end;
        if (t == threads - 1) end = n;
        sl[t] = (oracle_slice){emission_scores, model_length, transitions3, codes, offsets, at, end, scores};
        at = end;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, oracle_slice_main, &sl[t]);
    oracle_slice_main(&sl[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    free(sl);
    free(th);
}
