/*
 * oracle/sgns_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (C + pthreads) of skip-gram negative sampling as gensim 3.2.0 runs it
 * behind the reference's learn_embeddings (src/main.py:82-90, src/main_link.py:36-41,
 * :304-349: Word2Vec(walks, size, window, min_count=0, sg=1, workers, iter), everything
 * else gensim defaults: negative=5, sample=1e-3, alpha=0.025, min_alpha=1e-4, hs=0,
 * seed=1, batch_words=10000).
 *
 * PARITY UNPINNED: gensim==3.2.0 (requirements.txt:17) is a third-party dependency that is
 * not vendored under /root/reference and cannot be installed here (no network); the
 * reference has no tests or golden vectors at this boundary. This file restates the
 * published algorithm of that release (gensim/models/word2vec.py: scan_vocab, scale_vocab,
 * make_cum_table, reset_weights, train/job_producer; gensim/models/word2vec_inner.pyx:
 * train_batch_sg, fast_sentence_sg_neg) from knowledge of the release. What is checked
 * end to end is the link-prediction AUC protocol of src/main_link.py:519-565.
 *
 * Two RNG modes:
 *   0 GENSIM  -- per job one 48-bit LCG stream (next_random*25214903917+11) used first for
 *               the sub-sampling draws then for the negatives, as train_batch_sg does; the
 *               per-position window shrink comes from a per-job splitmix64 stream (gensim
 *               uses model.random.randint; numpy's MT stream is not reproduced).
 *   1 PHILOX  -- every draw is a Philox4x32-10 word addressed by (epoch, sentence,
 *               position[, context position]), identical to the device kernel, so that a
 *               one-worker run here can be compared element-wise with a sequential device run.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void n2v_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#define EXP_TABLE_SIZE 1000
#define MAX_EXP 6
#define MAX_SENTENCE_LEN 10000

static float EXP_TABLE[EXP_TABLE_SIZE];
static pthread_once_t exp_once = PTHREAD_ONCE_INIT;

/* word2vec_inner.pyx init(): EXP_TABLE[i] = exp((i/1000*2-1)*6); then x/(x+1) */
static void build_exp_table(void)
{
    for (int i = 0; i < EXP_TABLE_SIZE; ++i) {
        float e = (float)exp(((float)i / (float)EXP_TABLE_SIZE * 2 - 1) * MAX_EXP);
        EXP_TABLE[i] = (float)(e / (e + 1));
    }
}

void sgns_oracle_exp_table(float *out)
{
    pthread_once(&exp_once, build_exp_table);
    memcpy(out, EXP_TABLE, sizeof(EXP_TABLE));
}

/* ---- vocabulary: scan_vocab + scale_vocab(min_count=0) + sort_vocab -------------------- */
typedef struct { int64_t count; int64_t first; int32_t id; } vrec_t;

static int vrec_cmp(const void *a, const void *b)
{
    const vrec_t *x = (const vrec_t *)a, *y = (const vrec_t *)b;
    if (x->count != y->count) return x->count > y->count ? -1 : 1;   /* count descending */
    if (x->first != y->first) return x->first < y->first ? -1 : 1;   /* stable: first seen */
    return 0;
}

/* tokens are ids in [0, n_ids) (negative = padding, ignored). Outputs: counts in vocab
 * order, index2id[V], id2index[n_ids] (-1 when the id never occurs). Returns V. */
int32_t sgns_oracle_vocab(const int32_t *tokens, int64_t n_tokens, int32_t n_ids,
                          int64_t *counts, int32_t *index2id, int32_t *id2index)
{
    vrec_t *rec = (vrec_t *)calloc((size_t)n_ids, sizeof(vrec_t));
    if (!rec) return -1;
    for (int32_t i = 0; i < n_ids; ++i) { rec[i].id = i; rec[i].first = -1; }
    for (int64_t t = 0; t < n_tokens; ++t) {
        int32_t id = tokens[t];
        if (id < 0 || id >= n_ids) continue;
        if (rec[id].count++ == 0) rec[id].first = t;
    }
    qsort(rec, (size_t)n_ids, sizeof(vrec_t), vrec_cmp);
    int32_t V = 0;
    for (int32_t i = 0; i < n_ids; ++i) id2index[i] = -1;
    for (int32_t i = 0; i < n_ids; ++i) {
        if (rec[i].count == 0) break;
        counts[V] = rec[i].count;
        index2id[V] = rec[i].id;
        id2index[rec[i].id] = V;
        ++V;
    }
    free(rec);
    return V;
}

/* scale_vocab's sample_int and make_cum_table (power 0.75, domain 2^31-1). */
int sgns_oracle_prepare(const int64_t *counts, int32_t V, double sample,
                        uint64_t *sample_int, uint32_t *cum_table)
{
    double retain_total = 0.0;
    for (int32_t i = 0; i < V; ++i) retain_total += (double)counts[i];
    double threshold;
    if (sample <= 0.0) threshold = retain_total;
    else if (sample < 1.0) threshold = sample * retain_total;
    else threshold = (double)(int64_t)(sample * (3.0 + sqrt(5.0)) / 2.0);
    for (int32_t i = 0; i < V; ++i) {
        double v = (double)counts[i];
        double prob = (sqrt(v / threshold) + 1.0) * (threshold / v);
        if (!(prob < 1.0)) prob = 1.0;
        sample_int[i] = (uint64_t)llround(prob * 4294967296.0);
    }
    const double power = 0.75, domain = 2147483647.0;
    double total = 0.0, cum = 0.0;
    for (int32_t i = 0; i < V; ++i) total += pow((double)counts[i], power);
    for (int32_t i = 0; i < V; ++i) {
        cum += pow((double)counts[i], power);
        cum_table[i] = (uint32_t)llround(cum / total * domain);
    }
    return 0;
}

/* reset_weights: syn0[i] = (rand(d) - 0.5) / d, one seeded stream per word. gensim seeds a
 * numpy RandomState with hash(word + str(seed)); here the stream is Philox keyed by
 * (seed, word index) -- the same function the device uses (n2v_sgns_init_rows). */
void sgns_oracle_init_syn0(float *syn0, int32_t V, int32_t dim, uint64_t seed)
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int32_t i = 0; i < V; ++i)
        for (int32_t c = 0; c < dim; c += 4) {
            uint32_t ctr[4] = { (uint32_t)i, (uint32_t)(c >> 2), 0x53594E30u, 0u }, r[4];
            n2v_oracle_philox4x32_10(ctr, key, r);
            for (int32_t k = 0; k < 4 && c + k < dim; ++k)
                syn0[(int64_t)i * dim + c + k] =
                    (float)(((double)r[k] * (1.0 / 4294967296.0) - 0.5) / (double)dim);
        }
}

/* ---- training ------------------------------------------------------------------------ */
typedef struct {
    const int32_t *tok; const int64_t *sent_off; int64_t n_sent;
    int32_t V, dim, window, negative;
    const uint64_t *sample_int; const uint32_t *cum_table;
    int32_t rng_mode; uint64_t seed;
    float *syn0, *syn1neg;
    /* jobs */
    int64_t n_jobs; const int64_t *job_begin; const int64_t *job_end; const float *job_alpha;
    int64_t next_job; pthread_mutex_t mu;
    int64_t pairs;
} train_ctx_t;

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* bisect_left(cum_table, r) */
static inline int32_t bisect_left_u32(const uint32_t *a, int32_t n, uint32_t r)
{
    int32_t lo = 0, hi = n;
    while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (a[mid] < r) lo = mid + 1; else hi = mid; }
    return lo;
}

/* fast_sentence_sg_neg, word2vec_inner.pyx: word_index = centre (positive target in
 * syn1neg), word2_index = context (input row in syn0). neg_idx[] are the already drawn
 * negative vocabulary indices. */
static void sg_neg_pair(train_ctx_t *c, int32_t word_index, int32_t word2_index,
                        const int32_t *neg_idx, float alpha, float *work)
{
    const int32_t d = c->dim;
    float *row1 = c->syn0 + (int64_t)word2_index * d;
    memset(work, 0, sizeof(float) * (size_t)d);
    for (int32_t k = 0; k < c->negative + 1; ++k) {
        int32_t target; float label;
        if (k == 0) { target = word_index; label = 1.0f; }
        else {
            target = neg_idx[k - 1];
            if (target == word_index) continue;
            label = 0.0f;
        }
        float *row2 = c->syn1neg + (int64_t)target * d;
        float f = 0.0f;
        for (int32_t i = 0; i < d; ++i) f += row1[i] * row2[i];          /* our_dot */
        if (f <= -MAX_EXP || f >= MAX_EXP) continue;
        f = EXP_TABLE[(int)((f + MAX_EXP) * (EXP_TABLE_SIZE / MAX_EXP / 2))];
        float g = (label - f) * alpha;
        for (int32_t i = 0; i < d; ++i) work[i] += g * row2[i];          /* saxpy */
        for (int32_t i = 0; i < d; ++i) row2[i] += g * row1[i];          /* saxpy */
    }
    for (int32_t i = 0; i < d; ++i) row1[i] += work[i];                  /* word_locks == 1 */
}

/* Philox word addressing shared with the device kernel (csrc/n2v_sgns.cu):
 *  sub-sample/window: ctr = (sent lo, sent hi, pos, epoch<<8 | 0) -> r[0] sub-sample draw,
 *                     r[1] % window = shrink b.
 *  negatives:         ctr = (sent lo, sent hi, pos_i<<16 | pos_j, epoch<<8 | 1+blk) ->
 *                     4 draws per block, r % cum_table[V-1]. */
static inline void philox_words(uint64_t seed, uint64_t sent, uint32_t c2, uint32_t c3, uint32_t r[4])
{
    uint32_t ctr[4] = { (uint32_t)sent, (uint32_t)(sent >> 32), c2, c3 };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    n2v_oracle_philox4x32_10(ctr, key, r);
}

/* one job == one train_batch_sg call */
static int64_t run_job(train_ctx_t *c, int64_t job, int32_t *indexes, int32_t *origpos,
                       int64_t *sidx, int32_t *redwin, float *work, int32_t *neg)
{
    const int64_t s0 = c->job_begin[job], s1 = c->job_end[job];
    const float alpha = c->job_alpha[job];
    const int64_t n_sent = c->n_sent;
    uint64_t next_random = 0, sm = c->seed * 0x9E3779B97F4A7C15ull + (uint64_t)job;
    if ((c->rng_mode & 1) == 0) next_random = splitmix64(&sm) & 281474976710655ull;   /* 2^24*r1+r2 */

    int64_t eff_words = 0, eff_sent = 0;
    sidx[0] = 0;
    for (int64_t s = s0; s < s1; ++s) {
        const int64_t gs = s % n_sent; const uint32_t epoch = (uint32_t)(s / n_sent);
        const int64_t b = c->sent_off[gs], e = c->sent_off[gs + 1];
        if (e <= b) continue;
        for (int64_t t = b; t < e; ++t) {
            int32_t w = c->tok[t];
            if (w < 0) continue;                                  /* not in vocab / padding */
            if (c->sample_int) {
                uint64_t r32;
                if ((c->rng_mode & 1) == 0) {
                    next_random = (next_random * 25214903917ull + 11ull) & 281474976710655ull;
                    r32 = next_random >> 16;
                } else {
                    uint32_t r[4]; philox_words(c->seed, (uint64_t)gs, (uint32_t)(t - b), epoch << 8, r);
                    r32 = r[0];
                }
                if (c->sample_int[w] < r32) continue;
            }
            indexes[eff_words] = w;
            origpos[eff_words] = (int32_t)(t - b);
            if ((c->rng_mode & 1) == 0) redwin[eff_words] = (int32_t)(splitmix64(&sm) % (uint64_t)c->window);
            else {
                uint32_t r[4]; philox_words(c->seed, (uint64_t)gs, (uint32_t)(t - b), epoch << 8, r);
                redwin[eff_words] = (int32_t)(r[1] % (uint32_t)c->window);
            }
            ++eff_words;
            if (eff_words == MAX_SENTENCE_LEN) break;
        }
        ++eff_sent;
        sidx[eff_sent] = eff_words;
        /* remember which global sentence this effective sentence is (for Philox addressing) */
        sidx[MAX_SENTENCE_LEN + 1 + eff_sent] = s;
        if (eff_words == MAX_SENTENCE_LEN) break;
    }

    int64_t pairs = 0;
    const uint32_t cum_last = c->cum_table[c->V - 1];
    for (int64_t si = 0; si < eff_sent; ++si) {
        const int64_t is = sidx[si], ie = sidx[si + 1];
        const int64_t s = sidx[MAX_SENTENCE_LEN + 1 + si + 1];
        const int64_t gs = s % n_sent; const uint32_t epoch = (uint32_t)(s / n_sent);
        for (int64_t i = is; i < ie; ++i) {
            int64_t j = i - c->window + redwin[i];
            if (j < is) j = is;
            int64_t k = i + c->window + 1 - redwin[i];
            if (k > ie) k = ie;
            int fresh = 1;
            for (; j < k; ++j) {
                if (j == i) continue;
                uint32_t rr[4] = {0, 0, 0, 0};
                /* rng_mode bit 1: ONE negative set per centre position, shared by all its context
                 * pairs (the device's shared-negative mode; not gensim's behaviour) */
                const int share = (c->rng_mode & 2) != 0;
                for (int32_t n = 0; n < c->negative && !(share && !fresh); ++n) {
                    uint32_t r32;
                    if ((c->rng_mode & 1) == 0) {
                        r32 = (uint32_t)(next_random >> 16);
                        next_random = (next_random * 25214903917ull + 11ull) & 281474976710655ull;
                    } else {
                        if ((n & 3) == 0)
                            philox_words(c->seed, (uint64_t)gs,
                                         ((uint32_t)origpos[i] << 16) | (share ? 0xFFFFu : (uint32_t)origpos[j]),
                                         (epoch << 8) | (uint32_t)(1 + (n >> 2)), rr);
                        r32 = rr[n & 3];
                    }
                    neg[n] = bisect_left_u32(c->cum_table, c->V, r32 % cum_last);
                }
                fresh = 0;
                sg_neg_pair(c, indexes[i], indexes[j], neg, alpha, work);
                ++pairs;
            }
        }
    }
    return pairs;
}

static void *worker(void *arg)
{
    train_ctx_t *c = (train_ctx_t *)arg;
    int32_t *indexes = (int32_t *)malloc(sizeof(int32_t) * MAX_SENTENCE_LEN);
    int32_t *origpos = (int32_t *)malloc(sizeof(int32_t) * MAX_SENTENCE_LEN);
    int32_t *redwin = (int32_t *)malloc(sizeof(int32_t) * MAX_SENTENCE_LEN);
    int64_t *sidx = (int64_t *)malloc(sizeof(int64_t) * 2 * (MAX_SENTENCE_LEN + 2));
    float *work = (float *)malloc(sizeof(float) * (size_t)c->dim);
    int32_t *neg = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c->negative > 0 ? c->negative : 1));
    int64_t pairs = 0;
    for (;;) {
        pthread_mutex_lock(&c->mu);
        int64_t job = c->next_job < c->n_jobs ? c->next_job++ : -1;
        pthread_mutex_unlock(&c->mu);
        if (job < 0) break;
        pairs += run_job(c, job, indexes, origpos, sidx, redwin, work, neg);
    }
    pthread_mutex_lock(&c->mu);
    c->pairs += pairs;
    pthread_mutex_unlock(&c->mu);
    free(indexes); free(origpos); free(redwin); free(sidx); free(work); free(neg);
    return NULL;
}

/* Word2Vec.train as called from __init__ (total_examples = corpus_count, epochs = iter):
 * job_producer batches whole sentences up to batch_words raw words; the job's alpha is fixed
 * when the job is created: alpha - (alpha-min_alpha) * pushed_examples/total_examples,
 * floored at min_alpha. tok[] holds vocabulary indices (or -1), sent_off[n_sent+1]. */
int sgns_oracle_train(const int32_t *tok, const int64_t *sent_off, int64_t n_sent,
                      int32_t V, int32_t dim, int32_t window, int32_t negative,
                      const uint64_t *sample_int, const uint32_t *cum_table,
                      float alpha0, float min_alpha, int32_t iter, int32_t batch_words,
                      int32_t workers, int32_t rng_mode, uint64_t seed,
                      float *syn0, float *syn1neg, int64_t *pairs_out)
{
    pthread_once(&exp_once, build_exp_table);
    if (n_sent <= 0 || V <= 0) { if (pairs_out) *pairs_out = 0; return 0; }
    const int64_t total_examples = n_sent * (int64_t)iter;
    if (batch_words > MAX_SENTENCE_LEN) batch_words = MAX_SENTENCE_LEN;
    /* job_producer */
    int64_t cap = total_examples + 1, n_jobs = 0;
    int64_t *jb = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    int64_t *je = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    float *ja = (float *)malloc(sizeof(float) * (size_t)cap);
    if (!jb || !je || !ja) { free(jb); free(je); free(ja); return -1; }
    int64_t batch_begin = 0, batch_size = 0, pushed = 0;
    float next_alpha = alpha0;
    for (int64_t s = 0; s < total_examples; ++s) {
        int64_t gs = s % n_sent, len = sent_off[gs + 1] - sent_off[gs];
        if (batch_size + len <= batch_words) batch_size += len;
        else {
            jb[n_jobs] = batch_begin; je[n_jobs] = s; ja[n_jobs] = next_alpha; ++n_jobs;
            if (min_alpha < next_alpha) {
                pushed += s - batch_begin;
                double progress = 1.0 * (double)pushed / (double)total_examples;
                double a = (double)alpha0 - ((double)alpha0 - (double)min_alpha) * progress;
                next_alpha = (float)(a > (double)min_alpha ? a : (double)min_alpha);
            }
            batch_begin = s; batch_size = len;
        }
    }
    if (batch_begin < total_examples) {
        jb[n_jobs] = batch_begin; je[n_jobs] = total_examples; ja[n_jobs] = next_alpha; ++n_jobs;
    }

    train_ctx_t c;
    memset(&c, 0, sizeof(c));
    c.tok = tok; c.sent_off = sent_off; c.n_sent = n_sent; c.V = V; c.dim = dim;
    c.window = window; c.negative = negative; c.sample_int = sample_int; c.cum_table = cum_table;
    c.rng_mode = rng_mode; c.seed = seed; c.syn0 = syn0; c.syn1neg = syn1neg;
    c.n_jobs = n_jobs; c.job_begin = jb; c.job_end = je; c.job_alpha = ja;
    pthread_mutex_init(&c.mu, NULL);
    if (workers < 1) workers = 1;
    if (workers == 1) worker(&c);
    else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)workers);
        for (int32_t t = 0; t < workers; ++t) pthread_create(&th[t], NULL, worker, &c);
        for (int32_t t = 0; t < workers; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    pthread_mutex_destroy(&c.mu);
    if (pairs_out) *pairs_out = c.pairs;
    free(jb); free(je); free(ja);
    return 0;
}

/* ---- block-partitioned training (csrc/n2v_sgns_block.cu) --------------------------------------
 * Not gensim's schedule: the multi-GPU form of the same per-pair arithmetic. Tables cut into
 * n_parts = 1 << lg row sets (row i -> part i % n_parts, local row i / n_parts).
 *
 * sgns_oracle_make_groups: the (centre, context) pairs of the sentences whose centre lies in `part`,
 * sub-sampling and window shrink addressed by Philox exactly as rng_mode bit 0 above, written as
 * n_parts streams (stream b = context in part b) in the order sentence, centre, context. The pairs
 * of one centre occurrence inside a stream form a group: 8 header words {0x80000000 | centre local
 * row, sentence index, position | pairs << 16, the centre's 5 negatives}, then one word per pair
 * (context local row). The negatives are rng_mode 3's draws for that centre (Philox ctr (sentence id,
 * position / G << 16 | 0xFFFF, epoch << 8 | 1 + n / 4) -> bisect_left(cum_table, r % cum[-1])), each
 * mapped to the word of the same local row in `part`. Two calls: words == NULL counts (stream_len[b],
 * in words), then fill with stream_off[b] = start of stream b. */
int sgns_oracle_make_groups(const int32_t *tok, const int64_t *sent_off, int64_t n_sent, int64_t sent_id_base,
                            int32_t window, const uint64_t *sample_int, uint64_t seed, uint32_t epoch,
                            int32_t part, int32_t lg, int32_t V, const uint32_t *cum_table, int32_t neg_group,
                            int64_t *stream_len, const int64_t *stream_off, uint32_t *words)
{
    enum { FN = 5, HDR = 8 };
    const uint32_t cum_last = cum_table[V - 1];
    const int32_t n_parts = 1 << lg, mask = n_parts - 1;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * MAX_SENTENCE_LEN);
    int32_t *red = (int32_t *)malloc(sizeof(int32_t) * MAX_SENTENCE_LEN);
    int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * MAX_SENTENCE_LEN);
    int64_t cur[64], hdr[64]; int32_t cnt[64];
    if (!idx || !red || !pos) { free(idx); free(red); free(pos); return -1; }
    for (int32_t b = 0; b < n_parts; ++b) cur[b] = words ? stream_off[b] : 0;
    for (int64_t s = 0; s < n_sent; ++s) {
        const uint64_t gs = (uint64_t)(sent_id_base + s);
        const int64_t b0 = sent_off[s], e0 = sent_off[s + 1];
        int64_t n = 0;
        for (int64_t t = b0; t < e0 && t - b0 < MAX_SENTENCE_LEN; ++t) {
            const int32_t w = tok[t];
            if (w < 0) continue;
            uint32_t r[4]; philox_words(seed, gs, (uint32_t)(t - b0), epoch << 8, r);
            if (sample_int && sample_int[w] < (uint64_t)r[0]) continue;
            idx[n] = w; red[n] = (int32_t)(r[1] % (uint32_t)window); pos[n] = (int32_t)(t - b0); ++n;
        }
        for (int64_t i = 0; i < n; ++i) {
            if ((idx[i] & mask) != part) continue;
            int64_t j = i - window + red[i]; if (j < 0) j = 0;
            int64_t k = i + window + 1 - red[i]; if (k > n) k = n;
            for (int32_t b = 0; b < n_parts; ++b) { hdr[b] = -1; cnt[b] = 0; }
            for (; j < k; ++j) {
                if (j == i) continue;
                const int32_t b = idx[j] & mask;
                if (hdr[b] < 0) { hdr[b] = cur[b]; cur[b] += HDR; }
                if (words) words[cur[b]] = (uint32_t)(idx[j] >> lg);
                ++cur[b]; ++cnt[b];
            }
            if (words) {
                uint32_t tg[FN], rr[4] = {0, 0, 0, 0};
                const uint32_t poskey = neg_group > 1 ? (uint32_t)pos[i] / (uint32_t)neg_group : (uint32_t)pos[i];
                for (int32_t n2 = 0; n2 < FN; ++n2) {
                    if ((n2 & 3) == 0) philox_words(seed, gs, (poskey << 16) | 0xFFFFu, (epoch << 8) | (uint32_t)(1 + (n2 >> 2)), rr);
                    int32_t t = bisect_left_u32(cum_table, V, rr[n2 & 3] % cum_last) >> lg;
                    if ((((int64_t)t << lg) | part) >= V) --t;
                    tg[n2] = (uint32_t)t;
                }
                for (int32_t b = 0; b < n_parts; ++b)
                    if (hdr[b] >= 0) {
                        words[hdr[b]] = 0x80000000u | (uint32_t)(idx[i] >> lg);
                        words[hdr[b] + 1] = (uint32_t)s;
                        words[hdr[b] + 2] = (uint32_t)pos[i] | ((uint32_t)cnt[b] << 16);
                        for (int32_t d = 0; d < FN; ++d) words[hdr[b] + 3 + d] = tg[d];
                    }
            }
        }
    }
    if (!words) for (int32_t b = 0; b < n_parts; ++b) stream_len[b] = cur[b];
    free(idx); free(red); free(pos);
    return 0;
}

/* sgns_oracle_train_groups: one stream against (syn0 part, syn1neg part `part`), group by group in
 * stream order: the centre's negatives come from the group header, alpha = the sentence's job alpha
 * (jobs of sent_per_job sentences, as the device's job_alpha). Pairs run through
 * fast_sentence_sg_neg's arithmetic in order, updates immediate -- which is what the device's carried
 * registers + one reduction per row compute when one warp runs alone. */
int sgns_oracle_train_groups(const uint32_t *words, int64_t n_words, float *syn0_part, float *syn1_part,
                             int32_t dim, float alpha0, float min_alpha, int64_t total_examples,
                             int64_t example_base, int64_t sent_per_job, int64_t *pairs_out)
{
    pthread_once(&exp_once, build_exp_table);
    enum { FN = 5, HDR = 8 };
    float *work = (float *)malloc(sizeof(float) * (size_t)dim);
    if (!work) return -1;
    int64_t pairs = 0, p = 0;
    while (p < n_words) {
        if (!(words[p] & 0x80000000u)) { free(work); return -2; }
        const int32_t centre = (int32_t)(words[p] & 0x7FFFFFFFu);
        const uint32_t s = words[p + 1];
        const int32_t cnt = (int32_t)(words[p + 2] >> 16);
        const uint32_t *tg = words + p + 3;
        const int64_t ex = example_base + (int64_t)s;
        const int64_t job_first = ex - ex % sent_per_job;
        const double prog = (double)job_first / (double)total_examples;
        const double al = (double)alpha0 - ((double)alpha0 - (double)min_alpha) * prog;
        const float alpha = (float)(al > (double)min_alpha ? al : (double)min_alpha);
        for (int32_t j = 0; j < cnt; ++j) {
            float *row1 = syn0_part + (int64_t)words[p + HDR + j] * dim;
            memset(work, 0, sizeof(float) * (size_t)dim);
            for (int32_t k = 0; k <= FN; ++k) {
                int32_t target; float label;
                if (k == 0) { target = centre; label = 1.0f; }
                else { target = (int32_t)tg[k - 1]; if (target == centre) continue; label = 0.0f; }
                float *row2 = syn1_part + (int64_t)target * dim;
                float f = 0.0f;
                for (int32_t i = 0; i < dim; ++i) f += row1[i] * row2[i];
                if (f <= -MAX_EXP || f >= MAX_EXP) continue;
                f = EXP_TABLE[(int)((f + MAX_EXP) * (EXP_TABLE_SIZE / MAX_EXP / 2))];
                const float g = (label - f) * alpha;
                for (int32_t i = 0; i < dim; ++i) work[i] += g * row2[i];
                for (int32_t i = 0; i < dim; ++i) row2[i] += g * row1[i];
            }
            for (int32_t i = 0; i < dim; ++i) row1[i] += work[i];
            ++pairs;
        }
        p += HDR + cnt;
    }
    if (pairs_out) *pairs_out = pairs;
    free(work);
    return 0;
}
