/*
 * msv_cuda.h -- C ABI of the B200 (sm_100a) MSV profile-HMM scan.
 *
 * This is the drop-in boundary for the one hot path of IvanTyulyandin/HMM_FASTA_Viterbi: everything that
 * MSV_HMM::parallel_run_on_sequence (reference algorithms/MSV_HMM.cpp:269-430) and its six OpenCL kernels
 * (algorithms/MSV_kernels.cl:1-65) compute, re-designed as one batched CUDA launch.  Plain pointers and sizes
 * only; no C++/torch types.  Host bindings (the C++ MSV_HMM class in hmm_fasta_viterbi_b200/host, the ctypes
 * wrapper in hmm_fasta_viterbi_b200/_cabi.py, the stubs in INTEGRATION.md) sit on top of exactly these symbols.
 *
 * Conventions
 *   - every function returns an int status: MSV_OK (0) or a negative MSV_ERR_* code; the message for the last
 *     failure on the calling thread is available from msv_cuda_last_error().  Nothing "prints and continues"
 *     (contrast: check_errors, reference MSV_HMM.cpp:198-203).
 *   - residues are byte codes 0..19 in the order A C D E F G H I K L M N P Q R S T V W Y
 *     (reference MSV_HMM.cpp:29-31); a sequence has NO '#' sentinel here.
 *   - a database is the concatenation of all sequences plus a uint64 offsets array of n+1 entries
 *     (sequence q = residues[offsets[q] .. offsets[q+1]) ).
 *   - scores are fp32 and bit-identical to MSV_HMM::run_on_sequence (reference MSV_HMM.cpp:74-113).
 *   - there is no CPU fallback: without a CUDA device every msv_cuda_* call fails with MSV_ERR_NO_DEVICE.
 *   - thread safety: calls that share a msv_model or a msv_db must be serialised by the caller (the reference's
 *     MSV_HMM is not re-entrant either, MSV_HMM.cpp:59-64); distinct handles may be used from distinct threads.
 *     The asynchronous entry points (msv_cuda_db_score_device, msv_cuda_db_score_gather, msv_cuda_db_viterbi_device,
 *     msv_cuda_db_viterbi_subset_device) must ALSO be ordered on the device: scans of one msv_db share its work queue and
 *     scratch lists, so two scans of the same database (with whatever models) must not overlap on the GPU -- launch them
 *     on one stream, or order the streams with events.  Scans of different databases may overlap freely.
 */
#ifndef MSV_CUDA_H
#define MSV_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSV_CUDA_ABI_VERSION 3
#define MSV_ALPHABET 20
#define MSV_TRANSITIONS 7 /* m->m m->i m->d i->m i->i d->m d->d, the order of Profile_HMM::transitions (Profile_HMM.hpp:29) */

enum {
    MSV_OK = 0,
    MSV_ERR_INVALID_ARGUMENT = -1,
    MSV_ERR_NO_DEVICE = -2,
    MSV_ERR_CUDA = -3,
    MSV_ERR_BAD_RESIDUE = -4,      /* a residue code >= 20 (reference: std::out_of_range from .at(), MSV_HMM.cpp:101,383) */
    MSV_ERR_MODEL_TOO_LONG = -5,
    MSV_ERR_OUT_OF_MEMORY = -6
};

typedef struct msv_model msv_model; /* device-resident model: emission table in kernel layout + transition scores */
typedef struct msv_db msv_db;       /* device-resident packed sequence database, bucketed longest-first */

int msv_cuda_abi_version(void);
const char* msv_cuda_last_error(void);
int msv_cuda_device_count(int* count);

/* ---------------------------------------------------------------------------------------------------------------
 * Host-side model arithmetic.  Every transcendental of the path is evaluated on the HOST with the same fp32
 * expressions and the same libm as the reference, and shipped to the device as data; the device only adds and
 * takes maxima.  These helpers are the single implementation all host bindings share.
 * ------------------------------------------------------------------------------------------------------------- */

/* table[j * model_length + i] = logf(match[i * 20 + j] / background[j])     (reference MSV_HMM.cpp:38-45) */
int msv_host_emission_table(const float* match_emissions, size_t model_length, float* table);
/* tr_B_Mk = logf(2 / float(model_length * (model_length + 1))), tr_E_C = tr_E_J = logf(1/2)   (MSV_HMM.cpp:47-53) */
int msv_host_model_transitions(size_t model_length, float* tr_B_Mk, float* tr_E_C, float* tr_E_J);
/* tr_loop = logf(L / float(L + 3)), tr_move = logf(3 / float(L + 3))                          (MSV_HMM.cpp:59-64) */
int msv_host_length_transitions(size_t residues, float* tr_loop, float* tr_move);
/* letters -> codes; returns MSV_OK or MSV_ERR_BAD_RESIDUE (and the offending position in *bad_at if not NULL) */
int msv_host_encode(const char* letters, size_t n, uint8_t* codes, size_t* bad_at);
/* Cut n sequences into `parts` contiguous slices of (nearly) equal residue count, i.e. equal DP cell count for a
 * fixed model: slice p = sequences [bounds[p], bounds[p+1]).  `bounds` has parts+1 entries.  Used to shard a
 * database over GPUs / ranks. */
int msv_host_partition_by_cells(const uint64_t* offsets, size_t n, int parts, size_t* bounds);

/* ---------------------------------------------------------------------------------------------------------------
 * Model.  Replaces the per-call context + JIT + 20 emission buffers of the reference (MSV_HMM.cpp:287-340).
 *   emission_scores : [20][model_length] fp32, row stride model_length, column 0 = dummy M0 (= -inf),
 *                     exactly the layout of MSV_HMM::emission_scores (MSV_HMM.hpp:27-28, MSV_HMM.cpp:43)
 *   model_length    : LENG + 1 (reference Profile_HMM.cpp:70)
 * The table is re-laid out for the kernel ([residue][column-quad][lane][4], -inf padded) and kept in HBM; each
 * CTA stages it into shared memory with bulk-async (TMA) copies and, for whole-warp geometries, partly into tensor memory.
 * ------------------------------------------------------------------------------------------------------------- */
int msv_cuda_model_create(const float* emission_scores, size_t model_length, float tr_B_Mk, float tr_E_C, float tr_E_J,
                          int device, msv_model** out);
int msv_cuda_model_destroy(msv_model* model);
/* geometry of the model's throughput plan: lanes per sequence (32 = one warp, 128 = four warps when the model is too
 * long for one warp; a model may also carry an 8-lane plan for short models and a 128-lane plan for few/long sequences,
 * chosen per launch), model columns per lane, how many of those columns are served from tensor memory (TMEM; -1 = kernel
 * family without TMEM), threads per CTA, dynamic shared memory bytes.  Any pointer may be NULL. */
int msv_cuda_model_geometry(const msv_model* model, int* lanes_per_sequence, int* columns_per_lane,
                            int* tensor_columns_per_lane, int* threads_per_cta, size_t* shared_bytes);

/* geometry of the single-sequence latency kernels behind msv_cuda_score_sequence (msv_wave_kernels.cuh).  First choice: one
 * thread per diagonal phase, no communication, the whole table in every CTA's shared memory -- *diagonal_ctas CTAs of 128
 * threads (0 = the model is too long for it).  Otherwise a chain of warps over a thread-block cluster: columns per lane
 * (0 = the model has no such plan either), warps in the chain, CTAs in the cluster.  Any pointer may be NULL. */
int msv_cuda_model_wave_geometry(const msv_model* model, int* columns_per_lane, int* warps, int* ctas, int* diagonal_ctas);

/* Row variants of the warp-per-sequence kernel -- the same bits always.  Exact rows compute B = max(N, J) + move from the
 * row's E.  Speculative rows take B = N + move, which holds while no hit has lifted J above N, verify it, and repeat
 * exactly what failed: on whole sequences (a failed sequence is scanned again: fastest when hits are rare) or in checkpointed
 * blocks of 64 rows (a hit costs one block).  The library picks per launch from the model length, the mean sequence length
 * and the share of sequences that failed (or, with exact rows, would have failed) the check in the previous scans with this
 * model.  This reads the running totals behind that choice (as of the last finished scan) and the variant the next scan of
 * ordinary-length sequences would use.  Any pointer may be NULL. */
#define MSV_ROWS_EXACT 0
#define MSV_ROWS_SPECULATE_WHOLE 1
#define MSV_ROWS_SPECULATE_BLOCKS 2
int msv_cuda_model_speculation(const msv_model* model, unsigned int* failed, unsigned int* scanned, int* rows_next);

/* which launch plan a scan of the whole of `db` with `model` would use (introspection for tests and tuning): lanes per
 * sequence of the chosen kernel family (8, 32 or 128) and sequences in flight per CTA.  Does not launch anything. */
int msv_cuda_model_plan(const msv_model* model, const msv_db* db, int* lanes_per_sequence, int* sequences_per_cta);
/* ... and, for the lane-group plans of short models, how that scan would treat the longest sequences: a lane group scans a
 * sequence at 1/sequences_per_cta of its SM's speed, so the *n_long longest sequences (0 = none need it) are handed to
 * *fast_ctas CTAs that run only *fast_sequences_per_cta sequences at a time, while all other CTAs keep every slot busy with
 * the rest.  Known only for databases whose offsets passed through the host (not for msv_cuda_db_create_from_fasta). */
int msv_cuda_model_plan_long_sequences(const msv_model* model, const msv_db* db, unsigned int* n_long, int* fast_ctas,
                                       int* fast_sequences_per_cta);

/* ---------------------------------------------------------------------------------------------------------------
 * Database (device resident).  Uploads residues + offsets, validates the codes, computes the per-length
 * tr_loop/tr_move table on the host, and buckets the sequences longest-first on the device.
 * `residues`/`offsets` are HOST pointers and may be freed after the call.
 * ------------------------------------------------------------------------------------------------------------- */
int msv_cuda_db_create(int device, const uint8_t* residues, const uint64_t* offsets, size_t n, msv_db** out);
int msv_cuda_db_destroy(msv_db* db);
int msv_cuda_db_info(const msv_db* db, size_t* n, uint64_t* total_residues, uint64_t* longest);

/* FASTA text in HOST memory (e.g. an mmap'ed file) -> device-resident database, parsed ON the GPU: the raw text is
 * uploaded (pageable memory is staged through pinned buffers by several host threads) and classified, encoded and cut
 * into records by a handful of HBM-bound kernels (csrc/fasta_cuda.cu).  Record rules are those of the reference's
 * FASTA_protein_sequences (data_readers/FASTA_protein_sequences.cpp:9-44): a line starting with '>' opens a record and is
 * dropped, all other lines belong to the open record, a record with any character outside ACDEFGHIKLMNPQRSTVWY (a '\r'
 * is one) is dropped WHOLE, text before the first header is ignored.  *rejected (may be NULL) = dropped records. */
int msv_cuda_db_create_from_fasta(int device, const char* text, size_t bytes, msv_db** out, size_t* rejected);
/* the same into an EXISTING handle (its device buffers only grow, so a handle can be refilled file after file without
 * reallocating); on failure the handle is left empty */
int msv_cuda_db_refill_from_fasta(msv_db* db, const char* text, size_t bytes, size_t* rejected);
/* the packed form of a resident database back on the host (residues: total_residues bytes, offsets: n + 1 entries; either
 * may be NULL) -- what the device-side FASTA parser is tested with */
int msv_cuda_db_download(const msv_db* db, uint8_t* residues, uint64_t* offsets);

/* ---------------------------------------------------------------------------------------------------------------
 * Scoring.  One kernel launch scores a resident database (msv_cuda_score_batch issues one launch per upload stage so
 * that the copy engine and the scan overlap).  Limits: model_length - 1 <= 5631 columns, sequences < 2^27 residues,
 * fewer than 2^32 - 1 sequences per database.
 * ------------------------------------------------------------------------------------------------------------- */
/* resident path: scores_device is a DEVICE pointer to n floats (original sequence order); asynchronous on
 * `cuda_stream` (a cudaStream_t passed as void*, NULL = default stream). */
int msv_cuda_db_score_device(msv_model* model, msv_db* db, float* scores_device, void* cuda_stream);
/* Scan fused with the gather of a sharded run (one process per GPU, NVLink / NVSwitch): `gathered[r]`, r < n_gathered <= 8,
 * are DEVICE-ADDRESSABLE pointers to every participant's copy of the whole gathered score array -- this GPU's own copy
 * and the peers' copies mapped into this process (CUDA IPC / symmetric memory; torch.distributed._symmetric_memory in
 * bench.py).  The kernel stores the score of local sequence q into gathered[r][first_index + q] for every r, straight
 * from the lane that computed it, so no collective follows the scan: after a barrier every GPU holds all scores.
 * (The reference has no multi-device path; the NCCL all-gather of msv_cuda_db_score_device's output is the plain
 * alternative.)  Asynchronous on `cuda_stream`. */
int msv_cuda_db_score_gather(msv_model* model, msv_db* db, float* const* gathered, int n_gathered, size_t first_index,
                             void* cuda_stream);
/* resident database, host result: synchronous, copies n floats into scores_host. */
int msv_cuda_db_score(msv_model* model, msv_db* db, float* scores_host);
/* end-to-end path with HOST buffers in and out: upload + bucket + scan + download in one synchronous call.
 * This is what MSV_HMM::parallel_run_on_sequences (the batch entry point the C++ class adds) calls. */
int msv_cuda_score_batch(msv_model* model, const uint8_t* residues, const uint64_t* offsets, size_t n, float* scores_host);
/* The end-to-end call of a SHARDED run (one process per GPU): HOST buffers of this rank's slice in, and the scores leave
 * through the fused gather of msv_cuda_db_score_gather -- local sequence q goes to gathered[r][first_index + q] for every
 * r < n_gathered (this GPU's copy first, then the peers' copies mapped over NVLink) -- instead of a host array.  Upload,
 * validation, bucketing and scan are pipelined as in msv_cuda_score_batch.  Synchronous for THIS GPU's work; the caller
 * then runs its cross-rank barrier (bench.py: the symmetric-memory barrier) before anybody reads the gathered array. */
int msv_cuda_score_batch_gather(msv_model* model, const uint8_t* residues, const uint64_t* offsets, size_t n, float* const* gathered,
                                int n_gathered, size_t first_index);
/* FASTA text -> scores in one synchronous call: msv_cuda_db_create_from_fasta into the model's workspace, one scan, D2H.
 * scores_host has room for `capacity` floats; *n_sequences (may be NULL) receives the number of sequences found -- when it
 * exceeds `capacity` the call fails with MSV_ERR_INVALID_ARGUMENT and nothing is written. */
int msv_cuda_score_fasta(msv_model* model, const char* text, size_t bytes, float* scores_host, size_t capacity, size_t* n_sequences,
                         size_t* rejected);
/* the CUDA device a model lives on */
int msv_cuda_model_device(const msv_model* model, int* device);
/* one sequence, synchronous: the body of MSV_HMM::parallel_run_on_sequence (reference MSV_HMM.cpp:269-430).  One launch
 * of a latency kernel (the sequence travels in the kernel parameters when it is at most 3968 residues long, the result
 * comes back through pinned host memory); a sequence that contains a real hit is re-scored by the exact kernel. */
int msv_cuda_score_sequence(msv_model* model, const uint8_t* residues, size_t length, float* score);

/* ---------------------------------------------------------------------------------------------------------------
 * Several GPUs of one box from ONE process.  The reference has no multi-device path (single context, queue on its first
 * device: MSV_HMM.cpp:230,317).  The host database is cut into `ngpu` contiguous slices of equal DP-cell count
 * (msv_host_partition_by_cells); slice g is uploaded to and scanned on models[g]'s GPU by its own host thread, pipelined
 * like msv_cuda_score_batch.  The scan has no exchange step; `gather` selects how the fp32 scores come together:
 *   MSV_GATHER_HOST  every GPU downloads its slice straight into scores_host (no device-side gather);
 *   MSV_GATHER_PEER  the scan kernels store every score into ONE array on models[0]'s GPU over NVLink peer access (the
 *                    fused gather of msv_cuda_db_score_gather), then one download;
 *   MSV_GATHER_NCCL  every GPU scans into its own buffer, one grouped ncclSend / ncclRecv round (communicators from
 *                    ncclCommInitAll; libnccl.so.2 is bound at run time) collects them on models[0]'s GPU, then one download.
 * After a PEER or NCCL call the whole job's scores also stay resident on models[0]'s GPU (msv_cuda_multi_gathered) for
 * device-side follow-up stages.  Scores are the same bits in every mode.  `models` (one per GPU, same model) are not owned.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct msv_multi msv_multi;
enum { MSV_GATHER_HOST = 0, MSV_GATHER_PEER = 1, MSV_GATHER_NCCL = 2 };
int msv_cuda_multi_create(msv_model* const* models, int ngpu, msv_multi** out);
int msv_cuda_multi_destroy(msv_multi* multi);
int msv_cuda_multi_score_batch(msv_multi* multi, const uint8_t* residues, const uint64_t* offsets, size_t n, float* scores_host,
                               int gather);
int msv_cuda_multi_gathered(const msv_multi* multi, const float** scores_device, size_t* n, int* device);

/* ---------------------------------------------------------------------------------------------------------------
 * MSV filter statistics (the step that follows the scan in HMMER3's pipeline; the reference parses the model's
 * "STATS LOCAL MSV mu lambda" line, Profile_HMM.cpp:82-85, but never uses it).  For every sequence of `db`, from the
 * raw scores (nats) left on the device by msv_cuda_db_score_device:
 *     null1(L) = L * log(L / (L + 1)) + log(1 / (L + 1))            length model of the null hypothesis
 *     bits     = (score - null1(L)) / ln 2
 *     P        = 1 - exp(-exp(-lambda * (bits - mu)))               Gumbel survival; -expm1 form for tiny values
 * evaluated in fp64 on the device and stored as fp32.  Either output pointer may be NULL.  Asynchronous on
 * `cuda_stream`.  Unlike the scan this is ordinary floating point: results agree with an fp64 host evaluation to
 * 1e-6 relative (tests/test_gpu_parity.py), not bit for bit.
 * ------------------------------------------------------------------------------------------------------------- */
int msv_cuda_db_filter_device(msv_db* db, const float* scores_device, float mu, float lambda, float* bits_device,
                              float* pvalues_device, void* cuda_stream);

/* Scan + statistics in one synchronous call on a resident database: raw scores, bit scores and P-values are copied to
 * the three host arrays (n floats each; bits_host / pvalues_host may be NULL).  This is the body of
 * MSV_HMM::msv_filter in the C++ layer. */
int msv_cuda_db_score_filter(msv_model* model, msv_db* db, float mu, float lambda, float* scores_host, float* bits_host,
                             float* pvalues_host);

/* ---------------------------------------------------------------------------------------------------------------
 * The filter stages kept ON THE DEVICE (HMMER3's acceleration pipeline: MSV filter, P <= F1 = 0.02, then the Viterbi filter
 * on the survivors, P <= F2 = 1e-3).  One synchronous call per stage; only the HITS cross PCIe:
 *   msv_cuda_db_msv_filter                 scan + statistics + selection (stream compaction in database order) of the
 *                                          sequences with P <= threshold; their indices stay on the device as "survivors";
 *   msv_cuda_db_viterbi_filter_survivors   Viterbi scan of exactly those survivors (the resident database is read through
 *                                          the index list, nothing is re-packed), statistics, second selection.
 * The hit arrays have room for `capacity` entries each (any may be NULL); *n_hits is the number selected (it may exceed
 * `capacity`: then the first `capacity` are returned).  hit_index is the sequence's index in the database.
 * ------------------------------------------------------------------------------------------------------------- */
int msv_cuda_db_msv_filter(msv_model* model, msv_db* db, float mu, float lambda, float threshold, uint32_t* hit_index, float* hit_score,
                           float* hit_bits, float* hit_pvalue, size_t capacity, size_t* n_hits);

/* Page-lock (pin) a caller-owned host buffer so that uploads from it run at full PCIe / C2C speed and overlap with the
 * scan (msv_cuda_score_batch, msv_cuda_db_create take any host memory; pageable memory is staged by the driver at a
 * fraction of the link speed).  Registration is expensive (of the order of 1 ms per 4 MB): do it once per database. */
int msv_cuda_host_register(const void* buffer, size_t bytes);
int msv_cuda_host_unregister(const void* buffer);

/* ---------------------------------------------------------------------------------------------------------------
 * Plan-7 local multihit Viterbi scan: match, insert AND delete states over the node transitions that the reference
 * parses (Profile_HMM::transitions, data_readers/Profile_HMM.hpp:27-29, Profile_HMM.cpp:105-121) and never uses -- the
 * direction its README.md:2-3 names.  The reference has no implementation to be bit-compatible with; the recurrence is
 * the published one (HMMER3's generic Viterbi) in the conventions of the reference's MSV path, so both scans see one
 * model: same emission table, uniform local entry tr_B_Mk, free local exit M_k -> E and D_LENG -> E, tr_E_C / tr_E_J,
 * length-dependent loop / move scores; insert emissions score 0 as in HMMER3's profile configuration.
 *     M[i][k] = e[x_i][k] + max(M[i-1][k-1] + tMM[k-1], I[i-1][k-1] + tIM[k-1], D[i-1][k-1] + tDM[k-1], B[i-1] + tr_B_Mk)
 *     I[i][k] = max(M[i-1][k] + tMI[k], I[i-1][k] + tII[k])            k < LENG
 *     D[i][k] = max(M[i][k-1] + tMD[k-1], D[i][k-1] + tDD[k-1])        k >= 2
 *     E[i]    = max(max_k M[i][k], D[i][LENG]);   J, C, N, B and the score as in the MSV path (MSV_HMM.cpp:107-112)
 * fp32 throughout, adds and maxima only on the device; scores are bit-identical to the scalar evaluation of these
 * equations (oracle/viterbi_oracle.c).  Models up to 32 x 80 - 1 = 2559 columns.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct msv_viterbi_model msv_viterbi_model;

/* log_transitions[node * 7 + t] = logf(transitions[node * 7 + t]) for node = 0 .. model_length - 1 (probabilities as
 * Profile_HMM stores them, including its '*' -> 1.0 quirk, Profile_HMM.cpp:40; rows 0 and LENG are never read). */
int msv_host_viterbi_transitions(const float* transitions, size_t model_length, float* log_transitions);
/* emission_scores as for msv_cuda_model_create; log_transitions [model_length][7], every value <= 0 (-inf allowed). */
int msv_cuda_viterbi_model_create(const float* emission_scores, const float* log_transitions, size_t model_length, float tr_B_Mk,
                                  float tr_E_C, float tr_E_J, int device, msv_viterbi_model** out);
int msv_cuda_viterbi_model_destroy(msv_viterbi_model* model);
/* one warp per sequence: model columns per lane, threads per CTA, dynamic shared memory bytes.  Any pointer may be NULL. */
int msv_cuda_viterbi_model_geometry(const msv_viterbi_model* model, int* columns_per_lane, int* threads_per_cta, size_t* shared_bytes);
/* resident database (the same msv_db the MSV scan uses), one launch; scores_device: DEVICE pointer to n floats in
 * original sequence order; asynchronous on `cuda_stream`. */
int msv_cuda_db_viterbi_device(msv_viterbi_model* model, msv_db* db, float* scores_device, void* cuda_stream);
/* resident database, host result (synchronous). */
int msv_cuda_db_viterbi(msv_viterbi_model* model, msv_db* db, float* scores_host);
/* Viterbi scan + filter statistics on a resident database (the second stage of HMMER3's pipeline: bit score against the
 * null length model and Gumbel P-value with the model's "STATS LOCAL VITERBI" mu / lambda, same formulas as
 * msv_cuda_db_filter_device); n floats per host array, bits_host / pvalues_host may be NULL. */
int msv_cuda_db_viterbi_filter(msv_viterbi_model* model, msv_db* db, float mu, float lambda, float* scores_host, float* bits_host,
                               float* pvalues_host);
/* a SUBSET of a resident database, given as `count` sequence indices in device memory (e.g. the survivors of a filter
 * stage); scores_device is indexed by the ORIGINAL sequence index (n floats); asynchronous on `cuda_stream`. */
int msv_cuda_db_viterbi_subset_device(msv_viterbi_model* model, msv_db* db, const uint32_t* indices_device, size_t count, float* scores_device,
                                      void* cuda_stream);
/* second filter stage: see msv_cuda_db_msv_filter */
int msv_cuda_db_viterbi_filter_survivors(msv_viterbi_model* model, msv_db* db, float mu, float lambda, float threshold, uint32_t* hit_index,
                                         float* hit_score, float* hit_bits, float* hit_pvalue, size_t capacity, size_t* n_hits);
/* HOST buffers in and out: upload + bucket + scan + download in one synchronous call. */
int msv_cuda_viterbi_batch(msv_viterbi_model* model, const uint8_t* residues, const uint64_t* offsets, size_t n, float* scores_host);

/* kernel launches issued by this library on the calling thread since the last reset (for bench.py's gpu_launches) */
uint64_t msv_cuda_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* MSV_CUDA_H */
