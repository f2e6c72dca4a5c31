/*
 * n2v_b200.h -- C ABI of libn2v_b200.so: the B200 (sm_100a) implementation of the
 * node2vec-by-ecc embedding hot path
 *
 *     alias-table build  ->  second-order biased random walks  ->  skip-gram negative sampling
 *     src/node2vec.py:176-204   src/node2vec.py:55-111            src/main.py:82-90 (gensim 3.2.0)
 *
 * (paths relative to the reference checkout). The reference is pure Python and has no FFI of
 * its own; these entry points are what a binding for that path binds (see INTEGRATION.md for
 * the ctypes stub that replaces src/node2vec.py and gensim.models.Word2Vec in place).
 *
 * Conventions
 *  - Every function returns 0 on success or a negative N2V_E* code; n2v_last_error() gives a
 *    thread-local message. There is NO CPU fallback: without a CUDA device every compute entry
 *    point fails with N2V_ECUDA.
 *  - No allocation inside. Every buffer is a DEVICE pointer owned by the caller (sizes passed
 *    explicitly); scratch space is passed in after asking the matching *_workspace_bytes().
 *  - Asynchronous on `stream` (a cudaStream_t passed as void*), no hidden synchronisation,
 *    current device, re-entrant across streams.
 *  - Graphs are CSR over compact node ids 0..n-1: row_ptr int64[n+1], col int32[nnz] ascending
 *    within each row (== the reference's sorted(G.neighbors(v)) under an order-preserving
 *    relabel), w float64[nnz] or NULL for an unweighted graph (every weight 1).
 *  - An alias slot is 8 bytes: { int32 alias, uint32 threshold }. With u2 = r2 * 2^-32 the
 *    reference's test `u2 < q[kk]` (node2vec.py:278) is `r2 < ceil(q * 2^32)`; entries whose
 *    threshold would be 2^32 store alias = own index so that either branch returns kk.
 *  - Walk randomness: Philox4x32-10, counter (walk_id lo, walk_id hi, step, trial),
 *    key = seed; alias mode uses words 0,1 of trial 0 as the reference's two
 *    np.random.rand() calls of that step (u = r * 2^-32).
 */
#ifndef N2V_B200_H
#define N2V_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define N2V_OK 0
#define N2V_EINVAL (-1)   /* bad argument */
#define N2V_ECUDA (-2)    /* CUDA runtime error (incl. no device) */
#define N2V_ENOMEM (-3)   /* workspace too small */

typedef struct { int32_t alias; uint32_t thr; } n2v_slot_t;

const char *n2v_last_error(void);
int n2v_version(void);
/* number of SMs of the current device (grid sizing on the host side); <0 on error */
int n2v_sm_count(void);

/* ---- (1) CSR builder: networkx ingest replacement ------------------------------------
 * replaces: nx.read_edgelist/to_undirected adjacency + sorted(G.neighbors(.)) --
 * src/main.py:66-80, src/main_link.py:20-34, src/node2vec.py:67,184.
 * COO arcs (src[i] -> dst[i], optional w[i]); undirected != 0 adds the reverse arcs.
 * Duplicate arcs collapse, the LAST one in input order wins (networkx add_edge semantics).
 * Output: row_ptr[n+1], col/w_out with capacity (undirected ? 2m : m), *nnz_out (device). */
size_t n2v_csr_workspace_bytes(int64_t m, int32_t n_nodes, int undirected);
int n2v_csr_from_coo(const int32_t *src, const int32_t *dst, const double *w, int64_t m,
                     int32_t n_nodes, int undirected, void *workspace, size_t workspace_bytes,
                     int64_t *row_ptr, int32_t *col, double *w_out, int64_t *nnz_out,
                     void *stream);

/* ---- (2) alias tables ------------------------------------------------------------------
 * etab_ptr[e] = offset of the edge table of arc e (size deg(col[e])), etab_ptr[nnz] = total
 * = Sigma_e deg(col[e]) (Sigma deg^2 when undirected): decides alias vs rejection mode. */
size_t n2v_etab_workspace_bytes(int64_t nnz);
int n2v_etab_offsets(const int64_t *row_ptr, const int32_t *col, int32_t n_nodes, int64_t nnz,
                     int64_t *etab_ptr, void *workspace, size_t workspace_bytes, void *stream);

/* replaces: per-node loop of preprocess_transition_probs (node2vec.py:184-188) and of
 * preprocess_transition_probs_popularity (:213-218, popwalk bit 0; is_item[v] != 0 marks the
 * '9999999'-prefixed item nodes). popwalk bit 1: the weights already are probabilities, skip
 * the normalisation -- then each row is exactly alias_setup(probs) (node2vec.py:240-269).
 * slots[nnz] is laid out like col. work_J int32[nnz] / work_q float64[nnz] are scratch that
 * holds, on return, the reference's raw (J, q) of every node table. */
int n2v_alias_build_nodes(const int64_t *row_ptr, const int32_t *col, const double *w,
                          int32_t n_nodes, const uint8_t *is_item, int popwalk,
                          n2v_slot_t *slots, int32_t *work_J, double *work_q, void *stream);

/* replaces: get_alias_edge (node2vec.py:133-152) for every arc, i.e. the edge loop of
 * preprocess_transition_probs (:194-199). Builds the tables of arcs [arc_begin, arc_end) into
 * slots[etab_ptr[e] ...]; work_J/work_q hold etab_ptr[arc_end]-etab_ptr[arc_begin] entries
 * (pass full-size arrays and keep them to read the reference's raw (J, q) back).
 * symmetric != 0 promises an undirected (symmetric) CSR. popwalk != 0 builds get_alias_edge_pop
 * instead (node2vec.py:154-174: weight / len(G[nbr]), return edge additionally / p, no q). */
int n2v_alias_build_edges(const int64_t *row_ptr, const int32_t *col, const double *w,
                          int32_t n_nodes, double p, double q, int symmetric, int popwalk,
                          const int64_t *etab_ptr, int64_t arc_begin, int64_t arc_end,
                          n2v_slot_t *slots, int32_t *work_J, double *work_q, void *stream);

/* ---- (3) walks -------------------------------------------------------------------------
 * replaces: simulate_walks / node2vec_walk (node2vec.py:55-95) and, same output, the
 * on-the-fly pair (:34-53,:97-111). Walk i starts at starts[i], has global id
 * walk_id_base + i; walks is int32[n_walks, L] padded with -1 after lens[i] (dead ends,
 * node2vec.py:76-77). Bit-exact to the reference under injected Philox uniforms. */
int n2v_walk_alias(const int64_t *row_ptr, const int32_t *col, const n2v_slot_t *node_slots,
                   const int64_t *etab_ptr, const n2v_slot_t *edge_slots, const int32_t *starts,
                   int64_t n_walks, int32_t L, uint64_t seed, uint64_t walk_id_base,
                   int32_t *walks, int32_t *lens, void *stream);

/* Second form of the alias walker, bit-identical output: every arc gets one 32-byte,
 * sector-aligned record {edge-table offset, row start << 24 | degree of its head, head node} so
 * that a step is two dependent sector reads (alias slot, then the chosen arc's record) instead of
 * four to five. recs: nnz * n2v_arc_record_bytes() bytes, 32-byte aligned (n2v_pack_arcs;
 * *overflow_flag set when a degree >= 2^24 / row start >= 2^40 does not fit); packed_rows as for
 * n2v_walk_reject_indexed (first step). */
size_t n2v_arc_record_bytes(void);
int n2v_pack_arcs(const int64_t *row_ptr, const int32_t *col, const int64_t *etab_ptr, int64_t nnz,
                  void *recs, int *overflow_flag, void *stream);
int n2v_walk_alias_packed(const uint64_t *packed_rows, const n2v_slot_t *node_slots, const void *recs,
                          const n2v_slot_t *edge_slots, const int32_t *starts, int64_t n_walks,
                          int32_t L, uint64_t seed, uint64_t walk_id_base, int32_t *walks,
                          int32_t *lens, void *stream);

/* Rejection-sampling walker (KnightKing style) for graphs whose edge tables do not fit: same
 * transition distribution as get_alias_edge (node2vec.py:142-150), no edge tables.
 * node_slots may be NULL when w is NULL (uniform candidate). symmetric as above.
 * counters: optional device uint64[4] accumulating {steps, trials, membership tests,
 * charged probes = sum ceil(log2(deg(prev)+1))}. */
int n2v_walk_reject(const int64_t *row_ptr, const int32_t *col, const double *w,
                    const n2v_slot_t *node_slots, double p, double q, int symmetric,
                    const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                    uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                    unsigned long long *counters, void *stream);

/* The same walkers with the reference's popularity laws (popwalk = "pop", node2vec.py:154-174,:206-237;
 * reached from main_link.py:206-219,:274-281). law == NULL: as above.
 *  first_slots: alias table of the FIRST step (prev absent), e.g. the popularity node tables of
 *      preprocess_transition_probs_popularity (:213-218: w / len(G[nbr]) for non-item nodes); later steps
 *      keep the plain get_alias_edge law with node_slots (weighted graphs) / a uniform candidate.
 *  pop_edges = 1: later steps follow get_alias_edge_pop (:154-174) -- the candidate is drawn from
 *      node_slots, which the caller builds over w / len(G[nbr]) for EVERY node, and accepted with
 *      alpha = 1/p on the return edge, 1 elsewhere (the reference does not use q there). */
typedef struct {
    const n2v_slot_t *first_slots;
    int32_t pop_edges;
} n2v_walk_law_t;
int n2v_walk_reject_law(const int64_t *row_ptr, const int32_t *col, const double *w,
                        const n2v_slot_t *node_slots, const n2v_walk_law_t *law, double p, double q,
                        int symmetric, const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                        uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                        unsigned long long *counters, void *stream);

/* Second form of the rejection walker: hashed distance-1 test + per-lane state machine (one
 * 8-byte load per lane per iteration; lanes never wait for each other's trials). Needs
 *  - packed_rows[v] = row_ptr[v] << 24 | deg(v)          (n2v_pack_rows; *overflow_flag set when
 *    a degree >= 2^24 or a row start >= 2^40 does not fit),
 *  - edge_hash: open-addressing set of all arcs, key = u << 32 | v, capacity = a power of two
 *    >= 2 * nnz slots of 8 bytes (n2v_edge_hash_capacity / n2v_edge_hash_build).
 * Same transition law and Philox addressing as n2v_walk_reject. strength (optional, weighted
 * symmetric graphs): float64[N] row sums of w -- lets the return edge be folded out as an outlier
 * for weighted graphs too (without it the dartboard is 1/p high and acceptance collapses for
 * small p). */
int n2v_pack_rows(const int64_t *row_ptr, int32_t n_nodes, uint64_t *packed, int *overflow_flag,
                  void *stream);
uint64_t n2v_edge_hash_capacity(int64_t nnz);
int n2v_edge_hash_build(const int64_t *row_ptr, const int32_t *col, int32_t n_nodes, int64_t nnz,
                        unsigned long long *table, uint64_t capacity, void *stream);
int n2v_walk_reject_indexed(const uint64_t *packed_rows, const int32_t *col, int64_t nnz, const double *w,
                            const double *strength, const n2v_slot_t *node_slots, const unsigned long long *edge_hash,
                            uint64_t hash_capacity, double p, double q, int symmetric,
                            const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                            uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                            unsigned long long *counters, void *stream);

int n2v_walk_reject_indexed_law(const uint64_t *packed_rows, const int32_t *col, int64_t nnz, const double *w,
                                const double *strength, const n2v_slot_t *node_slots, const n2v_walk_law_t *law,
                                const unsigned long long *edge_hash, uint64_t hash_capacity, double p, double q,
                                int symmetric, const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                                uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                                unsigned long long *counters, void *stream);

/* ---- (4) skip-gram negative sampling -----------------------------------------------------
 * replaces: gensim.models.Word2Vec(sentences, size, window, min_count=0, sg=1, ...) as called
 * by learn_embeddings (src/main.py:82-90, src/main_link.py:36-41,:304-349). */

/* scan_vocab: counts[id] += occurrences over tokens (negative ids = padding) */
int n2v_vocab_count(const int32_t *tokens, int64_t n_tokens, int32_t n_ids,
                    unsigned long long *counts, void *stream);

/* scale_vocab + make_cum_table over counts in vocabulary order (count descending):
 * keep_thr[i] = min(round(p_keep * 2^32), 2^32-1) (a token is dropped iff keep_thr < r32),
 * cum_table[i] = round(cumsum(count^0.75)/total * (2^31-1)), and the 2^bucket_bits-entry
 * index over the top bits of the 31-bit draw that makes bisect_left O(1) expected. */
size_t n2v_sgns_prepare_workspace_bytes(int32_t V);
int n2v_sgns_prepare(const unsigned long long *counts, int32_t V, double sample,
                     uint32_t *keep_thr, uint32_t *cum_table, int32_t *bucket_lo,
                     int32_t bucket_bits, void *workspace, size_t workspace_bytes, void *stream);

/* reset_weights: syn0[i] = (u - 0.5)/dim with u from Philox keyed (seed, row, column);
 * syn1neg = 0. */
int n2v_sgns_init(float *syn0, float *syn1neg, int32_t V, int32_t dim, uint64_t seed,
                  void *stream);

typedef struct {
    int32_t V;                 /* vocabulary size (rows of syn0 / syn1neg) */
    int32_t dim;               /* multiple of 4, <= 1024 */
    int32_t window;            /* gensim window (default 5; the reference passes 10) */
    int32_t negative;          /* <= 16 */
    int32_t bucket_bits;
    int32_t max_sentence_len;  /* tokens of a sentence beyond this are ignored (<= 10000) */
    float alpha0, min_alpha;   /* 0.025 -> 1e-4, linear in sentences dispatched */
    int64_t total_examples;    /* corpus sentences * epochs, all ranks */
    int64_t example_base;      /* global index of sentence 0 of this call (epoch*n + shard offset) */
    int64_t sent_per_job;      /* alpha is constant over this many sentences (gensim: 10000-word jobs) */
    uint32_t epoch;
    uint64_t seed;
    int32_t grid_warps;        /* concurrent warps (Hogwild width); 1 = sequential, deterministic */
    int32_t atomic_updates;    /* 0 = plain stores (gensim's racy Hogwild), 1 = red.global.add */
    int32_t negative_sharing;  /* 0 = fresh negatives for every (centre, context) pair (gensim);
                                  1 = one set per centre, shared by its context pairs (dim<=128, k=5) */
    int32_t tuning;            /* 0 = default. bits 0-1: resident blocks/SM of the d<=128,k=5 kernel
                                  (1: 4, 2: 8, else 6); bit 3: force the generic kernel; bit 4 (with
                                  negative_sharing): the tensor-core window-batch experiment (n2v_sgns_mma.cu) */
    int32_t hot_rows;          /* negative_sharing kernels: negatives among the first hot_rows vocabulary rows (the
                                  most frequent words) are NOT carried in registers across a centre's pairs but
                                  re-read and reduced pair by pair. A carried copy is stale by what the other warps
                                  holding the same row add meanwhile; for a hub word hundreds of warps hold it at
                                  once. Sequential semantics are unchanged. 0 = carry every row. */
} n2v_sgns_params_t;

/* One pass over sentences [0, n_sent): tokens int32 ids (node ids if vocab_of_id != NULL,
 * else vocabulary indices; negative = padding), sentence s = tokens[sent_off[s] ..
 * sent_off[s+1]) or, when sent_off == NULL, the fixed-stride row s of an [n_sent, stride]
 * buffer (a walk buffer). sent_id_base addresses the Philox draws (global sentence id).
 * Updates syn0/syn1neg in place; pairs_out (device uint64[2]): [0] += (centre, context) pairs
 * trained, [1] += centres whose output rows were carried (negative_sharing mode only). */
int n2v_sgns_train(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent,
                   int32_t stride, int64_t sent_id_base, const int32_t *vocab_of_id,
                   const uint32_t *keep_thr, const uint32_t *cum_table, const int32_t *bucket_lo,
                   const n2v_sgns_params_t *params, float *syn0, float *syn1neg,
                   unsigned long long *pairs_out, void *stream);

/* Sharded tables: ONE logical syn0 / syn1neg pair spread over n_parts (1, 2, 4 or 8) allocations,
 * vocabulary row i in part i % n_parts at local row i / n_parts. The parts may live on other GPUs
 * of the node (peer memory mapped over NVLink, e.g. cudaIpcOpenMemHandle): every GPU then trains
 * its own walks against the same tables with red.global.add over NVLink -- gensim's shared-memory
 * Hogwild across GPUs, no replicas to reconcile. syn0_parts / syn1neg_parts are HOST arrays of
 * n_parts device pointers. Shared-negative kernel only (negative_sharing = 1, dim <= 128, k = 5).
 * n2v_sgns_init_part initialises one part (what n2v_sgns_init does for part 0 of 1). */
int n2v_sgns_init_part(float *syn0_part, float *syn1neg_part, int32_t V, int32_t dim, uint64_t seed,
                       int32_t part, int32_t n_parts, void *stream);
int n2v_sgns_train_sharded(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent,
                           int32_t stride, int64_t sent_id_base, const int32_t *vocab_of_id,
                           const uint32_t *keep_thr, const uint32_t *cum_table, const int32_t *bucket_lo,
                           const n2v_sgns_params_t *params, float *const *syn0_parts,
                           float *const *syn1neg_parts, int32_t n_parts,
                           unsigned long long *pairs_out, void *stream);

/* Block-partitioned training (the multi-GPU form of learn_embeddings, src/main.py:82-90; same
 * per-pair arithmetic and the same negative law as n2v_sgns_train with negative_sharing = 1).
 * Tables in n_parts row sets as above. A pool of walks is expanded into (centre, context) pairs;
 * the pairs whose centre is in part `part` are written as n_parts streams, stream b = pairs whose
 * context is in part b. Stream (part, b) touches only syn1neg part `part` and syn0 part b, so GPU
 * k trains stream (k, (k + e) % n_parts) in sub-step e and the syn0 parts travel round a ring: no
 * row is replicated, nothing is averaged.
 * Stream format, uint32 words: the pairs of one centre occurrence inside a stream are a GROUP = 8
 * header words {0x80000000 | centre local row, sentence index within this call, token position |
 * pairs << 16, the centre's 5 negatives as local rows of `part`} followed by one word per pair (the
 * context's local row); only a group's first word has bit 31 set.
 *   n2v_sgns_groups_count: offsets int64[n_parts * n_sent + 1], exclusive scan of the per-(stream,
 *       sentence) word counts in stream-major order: stream b = words [offsets[b * n_sent],
 *       offsets[(b + 1) * n_sent]); the last entry is the total.
 *   n2v_sgns_groups_fill: writes the words (order: stream, sentence, centre, context); words beyond
 *       capacity_words are dropped and counted in *overflow (device uint64). It also draws every
 *       centre's negative set, once, and repeats it in the header of each stream the centre reaches:
 *       Philox ctr (sent_id_base + sentence index, position / neg_group << 16 | 0xFFFF, epoch) ->
 *       count^0.75 table, i.e. the very draws n2v_sgns_train makes for that centre (neg_group = 1),
 *       each mapped to the word of the same local row in `part`.
 *   n2v_sgns_train_groups: trains the stream words[first_word, first_word + n_words) -- or, when
 *       dev_first / dev_end are given, [*dev_first, *dev_end) read on the device (two entries of
 *       `offsets`; no host round trip), clamped to capacity_words -- against (syn0 part of the
 *       stream's contexts, syn1neg part `part`). Every warp takes one contiguous range of whole
 *       groups. Negatives: the set in the group's header (ONE per centre occurrence); a negative
 *       equal to the centre is skipped, a set with a repeated row runs pair by pair with gensim's
 *       sequential semantics. neg_group = G > 1 (the value given to n2v_sgns_groups_fill): the centres
 *       at G consecutive token positions of a sentence carry the same set, which is then kept in
 *       registers across them. alpha follows the sentence's job (params->alpha0, min_alpha,
 *       total_examples, example_base, sent_per_job) as in n2v_sgns_train. Also uses params->V, dim
 *       (<= 128, % 4), negative (5), grid_warps, atomic_updates. pairs_out[0] += pairs, [1] += output rows carried in registers (one centre
 *       row per group + 5 per negative set): algorithmic bytes = 1,024 * (pairs + carried rows). */
size_t n2v_sgns_groups_workspace_bytes(int64_t n_sent, int32_t n_parts);
int n2v_sgns_groups_count(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                          int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                          const n2v_sgns_params_t *params, int32_t part, int32_t n_parts,
                          int64_t *offsets, void *workspace, size_t workspace_bytes, void *stream);
int n2v_sgns_groups_fill(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                         int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                         const n2v_sgns_params_t *params, int32_t part, int32_t n_parts,
                         const uint32_t *cum_table, const int32_t *bucket_lo, int32_t neg_group,
                         const int64_t *offsets, uint32_t *words, int64_t capacity_words,
                         unsigned long long *overflow, void *stream);
int n2v_sgns_train_groups(const uint32_t *words, int64_t first_word, int64_t n_words,
                          const int64_t *dev_first, const int64_t *dev_end, int64_t capacity_words,
                          const n2v_sgns_params_t *params, int32_t neg_group, float *syn0_part,
                          float *syn1neg_part, int32_t part, int32_t n_parts,
                          unsigned long long *pairs_out, void *stream);

/* ---- link scoring ---------------------------------------------------------------------------
 * replaces: link_score(emb, a, b) with link_method "cos" (src/main_link.py:43-49) over a batch of
 * pairs, as looped by get_roc_score (:173-189). a/b: row indices (-1 = word not in vocabulary ->
 * score 0, the reference's except branch). emb float32[V, dim], dim % 4 == 0. */
int n2v_cosine_pairs(const float *emb, int32_t dim, const int32_t *a, const int32_t *b,
                     int64_t n_pairs, float *out, void *stream);

/* All-pairs similarity with the selection fused in -- link_prediction / make_links_and_score /
 * links_score (src/main_link.py:70-171: score every user x item pair, or every unordered pair, keep the
 * best k) and build_user_sim_matrx + get_add_edge_by_* (:368-453: user x user similarity, per-user
 * threshold / top share). The n_a x n_b score matrix is never written.
 *   n2v_row_norms: mean[i] (0 unless centered: pearsonr == cosine of centred rows, :363) and
 *       inv_norm[i] of emb[rows[i]] (0 for a zero row or rows[i] < 0, so that pair scores 0).
 *   n2v_sim_threshold: S[r, c] = cos(emb[rows_a[r]], emb[rows_b[c]]) on the fp32 pipes, 64 x 64 tiles;
 *       a pair is emitted iff S > thr_row[r] (thr_row != NULL) or S > thr, and it passes the masks:
 *       upper_only = 1 keeps c > r (unordered pairs of one node list, :72), skip_diagonal = 1 scores
 *       (r, r) as 0 (:386), exclude = sorted int64 keys r * n_b + c (train edges, :73,:84).
 *       Emitted pairs go to out_a / out_b / out_score (positions into rows_a / rows_b) in no
 *       particular order; *count (device uint64, zeroed by the caller) counts ALL of them, so a
 *       caller that sees *count > capacity re-runs with more room or a higher threshold. */
int n2v_row_norms(const float *emb, int32_t dim, const int32_t *rows, int64_t n, int centered,
                  float *mean, float *inv_norm, void *stream);
int n2v_sim_threshold(const float *emb, int32_t dim, const int32_t *rows_a, int32_t n_a,
                      const int32_t *rows_b, int32_t n_b, const float *mean_a, const float *inv_a,
                      const float *mean_b, const float *inv_b, const float *thr_row, float thr,
                      int upper_only, int skip_diagonal, const long long *exclude, int64_t n_exclude,
                      int32_t *out_a, int32_t *out_b, float *out_score, int64_t capacity,
                      unsigned long long *count, void *stream);

/* ---- walk-file formatter ---------------------------------------------------------------------
 * replaces: " ".join(map(str, walk)) per line of the walk file (src/main_link.py:237-239,:544-546)
 * that LineSentence (:340) reads back. Two steps because the byte count is data dependent:
 *  1. n2v_format_walks_offsets: tok_off int64[n_walks*L + 1] = byte offset of every token slot
 *     (tok_off[n_walks*L] = total bytes); labels int64[N] = label of compact id, or NULL = ids;
 *  2. n2v_format_walks_write into out[total bytes]. */
size_t n2v_format_workspace_bytes(int64_t n_walks, int32_t L);
int n2v_format_walks_offsets(const int32_t *walks, const int32_t *lens, int64_t n_walks, int32_t L,
                             const int64_t *labels, int64_t *tok_off, void *workspace,
                             size_t workspace_bytes, void *stream);
int n2v_format_walks_write(const int32_t *walks, const int32_t *lens, int64_t n_walks, int32_t L,
                           const int64_t *labels, const int64_t *tok_off, unsigned char *out,
                           void *stream);

/* Parser of the same file (what LineSentence, src/main_link.py:340,346, does for a walk file):
 * whitespace-separated integer tokens, one sentence per line.
 *  1. n2v_parse_walks_index: flags + exclusive scans over the n_bytes+1 positions;
 *     tok_idx[n_bytes] = number of tokens, line_idx[n_bytes] = number of lines;
 *  2. n2v_parse_walks_fill: labels int64[n_tokens], sent_off int64[n_lines] = tokens before each
 *     line (append n_tokens to close the last line); *bad_flag set on a non-integer token. */
size_t n2v_parse_workspace_bytes(int64_t n_bytes);
int n2v_parse_walks_index(const unsigned char *text, int64_t n_bytes, int32_t *tok_flag, int32_t *line_flag,
                          int64_t *tok_idx, int64_t *line_idx, void *workspace, size_t workspace_bytes,
                          void *stream);
int n2v_parse_walks_fill(const unsigned char *text, int64_t n_bytes, const int32_t *tok_flag,
                         const int32_t *line_flag, const int64_t *tok_idx, const int64_t *line_idx,
                         int64_t *labels, int64_t *sent_off, int *bad_flag, void *stream);

/* ---- measurement helpers -------------------------------------------------------------------
 * Random-access HBM roofline denominators: mode 0 = one random 32-byte sector
 * read per access; mode 1 = random 512-byte row read-modify-write. buf holds n_bytes;
 * accesses per launch = n_access. sink: device uint64. */
int n2v_random_gather_bench(void *buf, size_t n_bytes, int64_t n_access, int mode, uint64_t seed,
                            unsigned long long *sink, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* N2V_B200_H */
