/* Fast seeded generator for the synthetic DNA pair workloads of BASELINE.json / SURVEY.md 8d.
 * Bench/test input infrastructure only -- not on the alignment path.  Pair p of a batch is a
 * pure function of (seed, p), so shards generated on different ranks are reproducible.
 *
 * Model: ancestor i.i.d. uniform over {A,C,G,T} of length L (optionally jittered); each child is
 * the ancestor with `subst` substitutions per site and indel events at rate `indel` per site,
 * geometric length with mean 3; a fraction `decorated` of the pairs additionally gets
 * internal-node-like symbols (ambiguity codes p=5%/site, gap-bit codes p=3%/site).
 * Every sequence is prefixed with the gap code 16. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint8_t base2(uint64_t r) { return (uint8_t)(1u << (r & 3)); }

typedef struct { uint32_t subst, indel, amb, gapb; } thresholds; /* 16-bit fixed point probabilities */

static int64_t child(uint64_t *st, const uint8_t *anc, int L, const thresholds *th, int decorate, uint8_t *out) {
    int64_t n = 0;
    int i = 0;
    out[n++] = 16;
    while (i < L) {
        uint64_t r = splitmix(st);
        if ((r & 0xFFFF) < th->indel) {
            int len = 1;
            uint64_t g = splitmix(st);
            while ((g % 3) != 0 && len < 36) { g /= 3; len++; }     /* P(continue) = 2/3: geometric, mean 3 */
            if (r & 0x10000) { i += len; continue; }               /* deletion */
            for (int k = 0; k < len; k++) out[n++] = base2(splitmix(st)); /* insertion */
        }
        uint8_t c = anc[i++];
        if (((r >> 20) & 0xFFFF) < th->subst) c = base2(r >> 40);
        if (decorate) {
            uint64_t q = splitmix(st);
            if ((q & 0xFFFF) < th->amb) c |= base2(q >> 16);
            if (((q >> 20) & 0xFFFF) < th->gapb) c |= 16;
        }
        out[n++] = c;
    }
    return n;
}

static int64_t one_pair(uint64_t seed, int64_t p, int L, double jitter, const thresholds *th, uint32_t dec_thr,
                        uint8_t *anc, uint8_t *out, int64_t *len_a) {
    uint64_t s0 = seed ^ 0x5851f42d4c957f2dull, st;
    st = splitmix(&s0) ^ ((uint64_t)p * 0xd1342543de82ef95ull);
    splitmix(&st); splitmix(&st);
    int len = L;
    if (jitter > 0) {
        double u = (double)(splitmix(&st) >> 11) * (1.0 / 9007199254740992.0);
        len = (int)(L * (1.0 + jitter * (2.0 * u - 1.0)));
        if (len < 1) len = 1;
    }
    for (int i = 0; i < len; i += 16) {
        uint64_t r = splitmix(&st);
        for (int k = 0; k < 16 && i + k < len; k++) { anc[i + k] = base2(r); r >>= 2; }
    }
    int dec = (splitmix(&st) & 0xFFFF) < dec_thr;
    int64_t na = child(&st, anc, len, th, dec, out);
    int64_t nb = child(&st, anc, len, th, dec, out + na);
    *len_a = na;
    return na + nb;
}

static int64_t pair_cap(int L, double jitter) { return 2 * ((int64_t)(L * (1.0 + jitter)) * 2 + 64); }

/* Upper bound on bytes needed for n pairs of nominal length L. */
int64_t synth_pairs_capacity(int64_t n, int L, double jitter) { return n * pair_cap(L, jitter); }

typedef struct {
    uint64_t seed; int64_t first, lo, hi; int L; double jitter; thresholds th; uint32_t dec_thr;
    uint8_t *data; int64_t *la, *lb; int64_t cap;
} job;

static void *worker(void *arg) {
    job *j = (job *)arg;
    uint8_t *anc = (uint8_t *)malloc((size_t)(j->L * (1.0 + j->jitter)) + 32);
    for (int64_t p = j->lo; p < j->hi; p++) {
        int64_t na, tot = one_pair(j->seed, j->first + p, j->L, j->jitter, &j->th, j->dec_thr, anc, j->data + p * j->cap, &na);
        j->la[p] = na; j->lb[p] = tot - na;
    }
    free(anc);
    return NULL;
}

/* Writes 2n sequences (pair p = sequences 2p, 2p+1) into data (capacity synth_pairs_capacity),
 * offsets[0..2n].  Returns total bytes. */
int64_t synth_pairs(uint64_t seed, int64_t first_pair, int64_t n, int L, double jitter, double subst, double indel,
                    double decorated, uint8_t *data, int64_t *offsets, int nthreads) {
    pthread_t th[64];
    job jobs[64];
    int64_t *la = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1)), *lb = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    thresholds t = { (uint32_t)(subst * 65536), (uint32_t)(indel * 65536), (uint32_t)(0.05 * 65536), (uint32_t)(0.03 * 65536) };
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (nthreads > n) nthreads = n > 0 ? (int)n : 1;
    for (int k = 0; k < nthreads; k++) {
        job jb = { seed, first_pair, n * k / nthreads, n * (k + 1) / nthreads, L, jitter, t, (uint32_t)(decorated * 65536),
                   data, la, lb, pair_cap(L, jitter) };
        jobs[k] = jb;
        pthread_create(&th[k], NULL, worker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    int64_t pos = 0, cap = pair_cap(L, jitter);
    offsets[0] = 0;
    for (int64_t p = 0; p < n; p++) {   /* compact */
        int64_t tot = la[p] + lb[p];
        if (pos != p * cap) memmove(data + pos, data + p * cap, (size_t)tot);
        offsets[2 * p + 1] = pos + la[p];
        offsets[2 * p + 2] = pos + tot;
        pos += tot;
    }
    free(la); free(lb);
    return pos;
}
