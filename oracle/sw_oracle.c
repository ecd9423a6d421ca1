/*
 * sw_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the scoring path of bmwoolf/mini_parallel, used only as the
 * checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference legs).  Nothing under mini_parallel_b200/ may
 * call into this file.
 *
 * Parity status: the reference ships no tests, fixtures or golden vectors for
 * this path -- "parity unpinned" by the reference's own tests.
 * It IS pinned against the reference's own kernel source executed here:
 * oracle/Makefile compiles /root/reference/smith_waterman/src/smith_waterman.cl
 * unmodified with gcc through an OpenCL-C work-item emulator (oracle/_ref/) and
 * tests/test_oracle_vs_ref.py + tests/golden/ check
 *     sw_last_row_max()   == smith_waterman_detailed  (dead kernel, cl:74-151)
 *     sw_linear().score   == max over prefixes of smith_waterman_detailed
 *     ref_compat_align()  == smith_waterman_align     (live kernel, cl:11-71)
 * End coordinates do not exist in the reference (gpu_align returns one i32,
 * aligner.rs:410,531); the tie-break below is this repository's definition
 * (SURVEY.md 8c): first cell reaching the maximum in a row-major scan.
 *
 * Reference lines followed:
 *   constants            smith_waterman.cl:5-7    (+2 / -1 / -2, linear gap)
 *   substitution         smith_waterman.cl:114    (raw byte equality)
 *   recurrence           smith_waterman.cl:116-125 (diag/left/up with zero borders, max with 0)
 *   last-row reduction   smith_waterman.cl:130-134
 *   live kernel          smith_waterman.cl:26-53 + launch geometry aligner.rs:422-424
 *   empty input -> 0     aligner.rs:413-416
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define SW_MATCH     2   /* smith_waterman.cl:5 */
#define SW_MISMATCH (-1) /* smith_waterman.cl:6 */
#define SW_GAP      (-2) /* smith_waterman.cl:7 */

typedef struct { int32_t score, end_i, end_j; } sw_result;

static inline int32_t max2(int32_t a, int32_t b) { return a > b ? a : b; }

/* Full Smith-Waterman, linear gap.  i indexes s1 (rows), j indexes s2 (columns),
 * exactly like seq1[i] / seq2[j] at smith_waterman.cl:114.  Row-major scan,
 * update on strict '>' => (max score, then smallest i, then smallest j).
 * Returns 0 on success, -1 on allocation failure. */
int sw_linear(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, sw_result* out)
{
    out->score = 0; out->end_i = -1; out->end_j = -1;
    if (n1 == 0 || n2 == 0) return 0;                 /* aligner.rs:413-416 */
    int32_t* row = (int32_t*)calloc(n2 + 1, sizeof(int32_t));   /* row[j+1] = H[i-1][j] before update */
    if (!row) return -1;
    int32_t best = 0, bi = -1, bj = -1;
    for (uint64_t i = 0; i < n1; ++i) {
        int32_t diag = 0;      /* H[i-1][j-1], zero border (cl:116) */
        int32_t left = 0;      /* H[i][j-1],   zero border (cl:117) */
        const uint8_t a = s1[i];
        for (uint64_t j = 0; j < n2; ++j) {
            const int32_t up = row[j + 1];                          /* cl:118 */
            const int32_t s  = (a == s2[j]) ? SW_MATCH : SW_MISMATCH; /* cl:114 */
            int32_t h = max2(max2(diag + s, left + SW_GAP), max2(up + SW_GAP, 0)); /* cl:120-125 */
            diag = up;
            row[j + 1] = h;
            left = h;
            if (h > best) { best = h; bi = (int32_t)i; bj = (int32_t)j; }
        }
    }
    free(row);
    out->score = best; out->end_i = bi; out->end_j = bj;
    return 0;
}

/* Alignment behind sw_linear()'s result: start cell and CIGAR.  Like the end cell, this does not
 * exist in the reference (gpu_align returns one i32, aligner.rs:410); the rule is this repository's definition:
 * full H matrix of the recurrence above (cl:114-125), then walk back from (end_i, end_j) while H > 0, at every cell taking
 * the FIRST predecessor that explains its value in the order
 *     diagonal  H == H[i-1][j-1] + s     '=' (7) on equal bytes, 'X' (8) otherwise
 *     up        H == H[i-1][j] + gap     'I' (1): a base of s1 (the read) against a gap
 *     left      H == H[i][j-1] + gap     'D' (2): a base of s2 (the window) against a gap
 * ops[] receives BAM-style run-length operations (len << 4 | op) in alignment order (start -> end).
 * Returns the number of operations, -1 on allocation failure, -2 when cap is too small. */
int sw_traceback(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, int32_t end_i, int32_t end_j,
                 int32_t* start_i, int32_t* start_j, uint32_t* ops, uint32_t cap)
{
    *start_i = -1; *start_j = -1;
    if (end_i < 0 || end_j < 0 || (uint64_t)end_i >= n1 || (uint64_t)end_j >= n2) return 0;
    const uint64_t R = (uint64_t)end_i + 1, W = (uint64_t)end_j + 1, stride = W + 1;
    int32_t* H = (int32_t*)calloc((R + 1) * stride, sizeof(int32_t));      /* H[(i+1)*stride + j+1], zero borders */
    if (!H) return -1;
    for (uint64_t i = 0; i < R; ++i)
        for (uint64_t j = 0; j < W; ++j) {
            const int32_t s = (s1[i] == s2[j]) ? SW_MATCH : SW_MISMATCH;
            H[(i + 1) * stride + j + 1] = max2(max2(H[i * stride + j] + s, H[(i + 1) * stride + j] + SW_GAP),
                                               max2(H[i * stride + j + 1] + SW_GAP, 0));
        }
    uint32_t* rev = (uint32_t*)malloc((R + W + 1) * sizeof(uint32_t));
    if (!rev) { free(H); return -1; }
    uint32_t n_ops = 0;
    int64_t i = end_i, j = end_j;
    while (i >= 0 && j >= 0 && H[(i + 1) * stride + j + 1] > 0) {
        const int32_t h = H[(i + 1) * stride + j + 1];
        const int32_t s = (s1[i] == s2[j]) ? SW_MATCH : SW_MISMATCH;
        uint32_t op;
        *start_i = (int32_t)i; *start_j = (int32_t)j;
        if (h == H[i * stride + j] + s) { op = (s1[i] == s2[j]) ? 7u : 8u; --i; --j; }
        else if (h == H[i * stride + j + 1] + SW_GAP) { op = 1u; --i; }
        else { op = 2u; --j; }
        if (n_ops && (rev[n_ops - 1] & 15u) == op) rev[n_ops - 1] += 16u;
        else rev[n_ops++] = (1u << 4) | op;
    }
    free(H);
    if (n_ops > cap) { free(rev); return -2; }
    for (uint32_t k = 0; k < n_ops; ++k) ops[k] = rev[n_ops - 1 - k];
    free(rev);
    return (int)n_ops;
}

/* Same recurrence, reduced the way the dead kernel reduces it: max over the LAST
 * row only (smith_waterman.cl:130-134). */
int32_t sw_last_row_max(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2)
{
    if (n1 == 0 || n2 == 0) return 0;
    int32_t* row = (int32_t*)calloc(n2 + 1, sizeof(int32_t));
    if (!row) return -1;
    for (uint64_t i = 0; i < n1; ++i) {
        int32_t diag = 0, left = 0;
        const uint8_t a = s1[i];
        for (uint64_t j = 0; j < n2; ++j) {
            const int32_t up = row[j + 1];
            const int32_t s  = (a == s2[j]) ? SW_MATCH : SW_MISMATCH;
            int32_t h = max2(max2(diag + s, left + SW_GAP), max2(up + SW_GAP, 0));
            diag = up; row[j + 1] = h; left = h;
        }
    }
    int32_t m = 0;
    for (uint64_t j = 0; j < n2; ++j) m = max2(m, row[j + 1]);
    free(row);
    return m;
}

/* What gpu_align() returns today: the live kernel smith_waterman_align under the
 * host's launch geometry (aligner.rs:422-424), result buffer taken as
 * zero-initialised (the reference leaves it uninitialised, aligner.rs:494-499).
 *   L = min(n1,n2); wgs = min(dev_max_wg,1024); groups = min(ceil(L/wgs), 1e6)
 *   chunk = ceil(L/groups); work-item (g,lid) walks i = g*chunk+lid; i<end; i+=wgs
 *   running clamp-sum of +2/-1, max over everything (cl:26-53, :56-69). */
int32_t ref_compat_align(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, uint32_t dev_max_wg)
{
    const uint64_t L = n1 < n2 ? n1 : n2;
    if (L == 0) return 0;                                            /* aligner.rs:413-416 */
    const uint64_t wgs = dev_max_wg < 1024u ? dev_max_wg : 1024u;    /* aligner.rs:422, gpu.rs:9 */
    uint64_t groups = (L + wgs - 1) / wgs;                           /* aligner.rs:423 */
    if (groups > 1000000u) groups = 1000000u;                        /* aligner.rs:424, gpu.rs:10 */
    const uint64_t chunk = (L + groups - 1) / groups;                /* cl:26 */
    int32_t result = 0;
    for (uint64_t g = 0; g < groups; ++g) {
        const uint64_t start = g * chunk;                            /* cl:27 */
        if (start >= L) break;                                       /* cl:30-32 */
        const uint64_t end = (start + chunk < L) ? start + chunk : L;/* cl:28 */
        for (uint64_t lid = 0; lid < wgs; ++lid) {
            int32_t mx = 0, cur = 0;                                 /* cl:35-36 */
            for (uint64_t i = start + lid; i < end; i += wgs) {      /* cl:39 */
                const int32_t s = (s1[i] == s2[i]) ? SW_MATCH : SW_MISMATCH;   /* cl:43-47 */
                cur = max2(cur + s, 0);                              /* cl:50 */
                mx  = max2(mx, cur);                                 /* cl:51 */
            }
            result = max2(result, mx);                               /* cl:56-69 */
        }
    }
    return result;
}

/* ---- batch form (CSR offsets), optionally multi-threaded: the scalar CPU port ---- */
typedef struct {
    const uint8_t* q; const uint64_t* qo; const uint8_t* r; const uint64_t* ro;
    uint64_t lo, hi; sw_result* out; int rc;
} batch_job;

static void* batch_worker(void* p)
{
    batch_job* jb = (batch_job*)p;
    for (uint64_t k = jb->lo; k < jb->hi; ++k)
        if (sw_linear(jb->q + jb->qo[k], jb->qo[k + 1] - jb->qo[k],
                      jb->r + jb->ro[k], jb->ro[k + 1] - jb->ro[k], &jb->out[k]) != 0) jb->rc = -1;
    return NULL;
}

int sw_linear_batch(const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro,
                    uint64_t n_pairs, sw_result* out, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_pairs) n_threads = n_pairs ? (int)n_pairs : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * n_threads);
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (batch_job){ q, qo, r, ro, n_pairs * t / n_threads, n_pairs * (t + 1) / n_threads, out, 0 };
        if (t) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    int rc = jobs[0].rc;
    for (int t = 1; t < n_threads; ++t) { pthread_join(th[t], NULL); rc |= jobs[t].rc; }
    free(th); free(jobs);
    return rc;
}
