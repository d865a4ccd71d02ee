/*
 * viterbi_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product path.
 *
 * CPU checker for the Plan-7 local multihit Viterbi scan ("Viterbi over the
 * already-parsed transitions", the reference README.md:2-3 names it as the project's direction; the reference has no
 * implementation of it).
 *
 * Parity status: PARITY UNPINNED.  There is no reference code, test or golden vector for this recurrence in
 * /root/reference.  The algorithm restated here is the published one (Durbin, Eddy, Krogh, Mitchison 1998, ch. 5;
 * Eddy 2011, "Accelerated profile HMM searches", the generic Viterbi of HMMER3) in the conventions the reference
 * already fixes for its MSV path, so that both scans see the same model:
 *   * match emission log-odds          logf(match[k][x] / background[x])            (MSV_HMM.cpp:38-45)
 *   * insert emission log-odds         0, as in HMMER3's profile configuration (inserts emit at background)
 *   * node transitions                 logf(p) of the seven probabilities Profile_HMM parses per node
 *                                      (Profile_HMM.hpp:27-29: m->m m->i m->d i->m i->i d->m d->d)
 *   * local entry B -> M_k             uniform, logf(2 / (model_length (model_length + 1)))   (MSV_HMM.cpp:51)
 *   * local exit M_k -> E, D_M -> E    0
 *   * E -> C, E -> J                   logf(1/2)                                      (MSV_HMM.cpp:52-53)
 *   * N/C/J loop and move              logf(L / (L+3)), logf(3 / (L+3))               (MSV_HMM.cpp:59-64)
 * Pinned by construction only: tests check (1) hand-computable cases, (2) that with delete/insert transitions
 * disabled and m->m = 1 the recurrence collapses to the MSV recurrence bit for bit (which IS pinned against the
 * reference), (3) an independent full-matrix evaluation of the same equations written in numpy.
 *
 * Recurrence, i = 1..L, k = 1..M (M = model_length - 1), every operation an IEEE binary32 add or max:
 *   Mx[i][k] = e[x_i][k] + max( Mx[i-1][k-1] + tMM[k-1], Ix[i-1][k-1] + tIM[k-1], Dx[i-1][k-1] + tDM[k-1], B[i-1] + tBMk )
 *   Ix[i][k] =             max( Mx[i-1][k] + tMI[k],     Ix[i-1][k] + tII[k] )                         k < M
 *   Dx[i][k] =             max( Mx[i][k-1] + tMD[k-1],   Dx[i][k-1] + tDD[k-1] )                        k >= 2
 *   E[i]     = max( max_k Mx[i][k], Dx[i][M] )
 *   J, C, N, B as in the MSV path (MSV_HMM.cpp:107-110); score = C[L] + tr_move.
 * Column 0 and row 0 are -inf; Dx[i][1] = -inf; there is no I_M.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VIT_TRANSITIONS 7
enum { T_MM = 0, T_MI = 1, T_MD = 2, T_IM = 3, T_II = 4, T_DM = 5, T_DD = 6 };

void oracle_msv_length_transitions(size_t residues, float* tr_loop, float* tr_move); /* msv_oracle.c */

/* log_transitions[k * 7 + t] = logf(transitions[k * 7 + t]) for every node k = 0 .. model_length - 1 */
void oracle_viterbi_prepare(const float* transitions, size_t model_length, float* log_transitions) {
    for (size_t i = 0; i < model_length * VIT_TRANSITIONS; ++i) log_transitions[i] = logf(transitions[i]);
}

static inline float vmax(float a, float b) { return (a < b) ? b : a; }

float oracle_viterbi_score_codes(const float* emission_scores, const float* log_transitions, size_t model_length,
                                 const float* transitions3, const uint8_t* codes, size_t L, float* scratch /* 6*model_length or NULL */) {
    const float tBMk = transitions3[0], tEC = transitions3[1], tEJ = transitions3[2];
    const size_t M = model_length - 1;
    float tr_loop, tr_move;
    oracle_msv_length_transitions(L, &tr_loop, &tr_move);

    float* own = NULL;
    if (!scratch) scratch = own = (float*)malloc(6 * model_length * sizeof(float));
    float *pm = scratch, *pi = pm + model_length, *pd = pi + model_length;
    float *cm = pd + model_length, *ci = cm + model_length, *cd = ci + model_length;
    for (size_t k = 0; k < model_length; ++k) pm[k] = pi[k] = pd[k] = cm[k] = ci[k] = cd[k] = -INFINITY;
#define TR(k, t) log_transitions[(k) * VIT_TRANSITIONS + (t)]

    float J = -INFINITY, C = -INFINITY, N = 0.0f, B = tr_move;
    for (size_t i = 0; i < L; ++i) {
        const float* e = emission_scores + (size_t)codes[i] * model_length;
        const float entry = B + tBMk;
        float E = -INFINITY;
        cm[0] = ci[0] = cd[0] = -INFINITY;
        for (size_t k = 1; k <= M; ++k) {
            float best = entry;
            if (k >= 2) {
                best = vmax(best, pm[k - 1] + TR(k - 1, T_MM));
                best = vmax(best, pi[k - 1] + TR(k - 1, T_IM));
                best = vmax(best, pd[k - 1] + TR(k - 1, T_DM));
            }
            cm[k] = e[k] + best;
            E = vmax(E, cm[k]);
            ci[k] = (k < M) ? vmax(pm[k] + TR(k, T_MI), pi[k] + TR(k, T_II)) : -INFINITY;
            cd[k] = (k >= 2) ? vmax(cm[k - 1] + TR(k - 1, T_MD), cd[k - 1] + TR(k - 1, T_DD)) : -INFINITY;
        }
        if (M >= 1) E = vmax(E, cd[M]);
        J = vmax(J + tr_loop, E + tEJ);
        C = vmax(C + tr_loop, E + tEC);
        N = N + tr_loop;
        B = vmax(N + tr_move, J + tr_move);
        float* t;
        t = pm, pm = cm, cm = t;
        t = pi, pi = ci, ci = t;
        t = pd, pd = cd, cd = t;
    }
#undef TR
    free(own);
    return C + tr_move;
}

typedef struct {
    const float *emis, *logtr;
    size_t model_length;
    const float* tr3;
    const uint8_t* codes;
    const uint64_t* offsets;
    size_t begin, end;
    float* out;
} vit_slice;

static void* vit_slice_main(void* arg) {
    vit_slice* s = (vit_slice*)arg;
    float* scratch = (float*)malloc(6 * s->model_length * sizeof(float));
    for (size_t q = s->begin; q < s->end; ++q)
        s->out[q] = oracle_viterbi_score_codes(s->emis, s->logtr, s->model_length, s->tr3, s->codes + s->offsets[q],
                                               (size_t)(s->offsets[q + 1] - s->offsets[q]), scratch);
    free(scratch);
    return NULL;
}

void oracle_viterbi_score_batch(const float* emission_scores, const float* log_transitions, size_t model_length,
                                const float* transitions3, const uint8_t* codes, const uint64_t* offsets, size_t n, float* scores,
                                int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n && n > 0) threads = (int)n;
    vit_slice* sl = (vit_slice*)calloc((size_t)threads, sizeof(vit_slice));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    const uint64_t total = n ? offsets[n] - offsets[0] : 0;
    size_t at = 0;
    for (int t = 0; t < threads; ++t) {
        size_t end = at;
        const uint64_t want = offsets[0] + (total * (uint64_t)(t + 1)) / (uint64_t)threads;
        while (end < n && offsets[end + 1] <= want) ++end;
        if (t == threads - 1) end = n;
        sl[t] = (vit_slice){emission_scores, log_transitions, model_length, transitions3, codes, offsets, at, end, scores};
        at = end;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, vit_slice_main, &sl[t]);
    vit_slice_main(&sl[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    free(sl);
    free(th);
}
