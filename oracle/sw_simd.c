/*
 * sw_simd.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The "reference CPU SIMD path" BASELINE.json asks to be timed beside the GPU does not
 * exist upstream (no std::arch / rayon use anywhere; every scoring route
 * goes through OpenCL, aligner.rs:410-532).  This file is the CPU port that stands in for
 * it: the same scoring function as sw_oracle.c (constants smith_waterman.cl:5-7, recurrence
 * cl:114-125, global max + first-in-row-major end cell), vectorised ACROSS pairs
 * (inter-sequence): one int16 lane per pair, 32 pairs per AVX-512BW vector, 16 per AVX2
 * vector, runtime dispatch, pthreads over pair blocks.  bench.py times it as
 * cpu_baseline / --impl reference (kind "port"); tests check it bit-for-bit against
 * sw_linear().
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <immintrin.h>

typedef struct { int32_t score, end_i, end_j; } sw_result;
int sw_linear(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, sw_result* out);

#define PAD_Q 256   /* lane sentinels outside the byte range: never equal to anything real */
#define PAD_R 257

/* ---------------- AVX-512BW: 32 pairs per vector ---------------- */
__attribute__((target("avx512bw,avx512f")))
static void block_avx512(const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro,
                         const uint64_t* idx, int cnt, sw_result* out, int16_t* qt, int16_t* rt, __m512i* hrow,
                         uint32_t maxq, uint32_t maxr)
{
    enum { L = 32 };
    for (uint32_t i = 0; i < maxq; ++i) for (int l = 0; l < L; ++l) qt[(size_t)i * L + l] = PAD_Q;
    for (uint32_t j = 0; j < maxr; ++j) for (int l = 0; l < L; ++l) rt[(size_t)j * L + l] = PAD_R;
    for (int l = 0; l < cnt; ++l) {
        const uint64_t k = idx[l];
        const uint8_t* a = q + qo[k]; const uint64_t n1 = qo[k + 1] - qo[k];
        const uint8_t* b = r + ro[k]; const uint64_t n2 = ro[k + 1] - ro[k];
        for (uint64_t i = 0; i < n1; ++i) qt[i * L + l] = a[i];
        for (uint64_t j = 0; j < n2; ++j) rt[j * L + l] = b[j];
    }
    const __m512i vmatch = _mm512_set1_epi16(2), vmis = _mm512_set1_epi16(-1), vgap = _mm512_set1_epi16(-2);
    const __m512i zero = _mm512_setzero_si512(), one = _mm512_set1_epi16(1);
    for (uint32_t j = 0; j <= maxr; ++j) hrow[j] = zero;
    __m512i best = zero, bi = _mm512_set1_epi16(-1), bj = _mm512_set1_epi16(-1), vi = zero;
    for (uint32_t i = 0; i < maxq; ++i) {
        const __m512i qa = _mm512_load_si512((const void*)(qt + (size_t)i * L));
        __m512i diag = zero, left = zero, vj = zero;
        for (uint32_t j = 0; j < maxr; ++j) {
            const __m512i up = hrow[j + 1];
            const __m512i rb = _mm512_load_si512((const void*)(rt + (size_t)j * L));
            const __mmask32 eq = _mm512_cmpeq_epi16_mask(qa, rb);
            const __m512i s = _mm512_mask_mov_epi16(vmis, eq, vmatch);
            __m512i h = _mm512_max_epi16(_mm512_add_epi16(diag, s), _mm512_add_epi16(_mm512_max_epi16(left, up), vgap));
            h = _mm512_max_epi16(h, zero);
            const __mmask32 gt = _mm512_cmpgt_epi16_mask(h, best);
            best = _mm512_max_epi16(best, h);
            bi = _mm512_mask_mov_epi16(bi, gt, vi);
            bj = _mm512_mask_mov_epi16(bj, gt, vj);
            diag = up; hrow[j + 1] = h; left = h;
            vj = _mm512_add_epi16(vj, one);
        }
        vi = _mm512_add_epi16(vi, one);
    }
    int16_t sb[L], si[L], sj[L];
    _mm512_storeu_si512((void*)sb, best); _mm512_storeu_si512((void*)si, bi); _mm512_storeu_si512((void*)sj, bj);
    for (int l = 0; l < cnt; ++l) { out[idx[l]].score = sb[l]; out[idx[l]].end_i = si[l]; out[idx[l]].end_j = sj[l]; }
}

/* ---------------- AVX2: 16 pairs per vector ---------------- */
__attribute__((target("avx2")))
static void block_avx2(const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro,
                       const uint64_t* idx, int cnt, sw_result* out, int16_t* qt, int16_t* rt, __m256i* hrow,
                       uint32_t maxq, uint32_t maxr)
{
    enum { L = 16 };
    for (uint32_t i = 0; i < maxq; ++i) for (int l = 0; l < L; ++l) qt[(size_t)i * L + l] = PAD_Q;
    for (uint32_t j = 0; j < maxr; ++j) for (int l = 0; l < L; ++l) rt[(size_t)j * L + l] = PAD_R;
    for (int l = 0; l < cnt; ++l) {
        const uint64_t k = idx[l];
        const uint8_t* a = q + qo[k]; const uint64_t n1 = qo[k + 1] - qo[k];
        const uint8_t* b = r + ro[k]; const uint64_t n2 = ro[k + 1] - ro[k];
        for (uint64_t i = 0; i < n1; ++i) qt[i * L + l] = a[i];
        for (uint64_t j = 0; j < n2; ++j) rt[j * L + l] = b[j];
    }
    const __m256i vgap = _mm256_set1_epi16(-2), zero = _mm256_setzero_si256(), one = _mm256_set1_epi16(1);
    const __m256i vmis = _mm256_set1_epi16(-1), vthree = _mm256_set1_epi16(3);
    for (uint32_t j = 0; j <= maxr; ++j) hrow[j] = zero;
    __m256i best = zero, bi = _mm256_set1_epi16(-1), bj = _mm256_set1_epi16(-1), vi = zero;
    for (uint32_t i = 0; i < maxq; ++i) {
        const __m256i qa = _mm256_load_si256((const __m256i*)(qt + (size_t)i * L));
        __m256i diag = zero, left = zero, vj = zero;
        for (uint32_t j = 0; j < maxr; ++j) {
            const __m256i up = hrow[j + 1];
            const __m256i rb = _mm256_load_si256((const __m256i*)(rt + (size_t)j * L));
            const __m256i eq = _mm256_cmpeq_epi16(qa, rb);                         /* 0xFFFF where equal */
            const __m256i s = _mm256_add_epi16(vmis, _mm256_and_si256(eq, vthree)); /* -1 + 3*eq */
            __m256i h = _mm256_max_epi16(_mm256_add_epi16(diag, s), _mm256_add_epi16(_mm256_max_epi16(left, up), vgap));
            h = _mm256_max_epi16(h, zero);
            const __m256i gt = _mm256_cmpgt_epi16(h, best);
            best = _mm256_max_epi16(best, h);
            bi = _mm256_blendv_epi8(bi, vi, gt);
            bj = _mm256_blendv_epi8(bj, vj, gt);
            diag = up; hrow[j + 1] = h; left = h;
            vj = _mm256_add_epi16(vj, one);
        }
        vi = _mm256_add_epi16(vi, one);
    }
    int16_t sb[L], si[L], sj[L];
    _mm256_storeu_si256((__m256i*)sb, best); _mm256_storeu_si256((__m256i*)si, bi); _mm256_storeu_si256((__m256i*)sj, bj);
    for (int l = 0; l < cnt; ++l) { out[idx[l]].score = sb[l]; out[idx[l]].end_i = si[l]; out[idx[l]].end_j = sj[l]; }
}

static int g_isa = -1;   /* 2 = avx512bw, 1 = avx2, 0 = scalar */
static int detect_isa(void)
{
    if (g_isa < 0) {
        __builtin_cpu_init();
        g_isa = __builtin_cpu_supports("avx512bw") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    }
    return g_isa;
}
const char* sw_simd_isa(void) { int i = detect_isa(); return i == 2 ? "avx512bw" : (i == 1 ? "avx2" : "scalar"); }
void sw_simd_force_isa(int isa) { detect_isa(); if (isa >= 0 && isa <= g_isa) g_isa = isa; }

typedef struct {
    const uint8_t* q; const uint64_t* qo; const uint8_t* r; const uint64_t* ro;
    uint64_t n_pairs, grain; uint64_t* next;     /* shared cursor: workers take `grain` pairs at a time (no static split: */
    sw_result* out; int rc;                      /*  a thread the OS de-schedules does not hold the whole batch back)      */
} simd_job;

static void* simd_worker(void* p)
{
    simd_job* jb = (simd_job*)p;
    const int isa = detect_isa();
    const int L = isa == 2 ? 32 : 16;
    uint64_t idx[32];
    size_t cap_q = 0, cap_r = 0;
    int16_t *qt = NULL, *rt = NULL; void* hrow = NULL;
    uint64_t k = 0, hi = 0;
    for (;;) {
        if (k >= hi) {
            k = __atomic_fetch_add(jb->next, jb->grain, __ATOMIC_RELAXED);
            if (k >= jb->n_pairs) break;
            hi = k + jb->grain < jb->n_pairs ? k + jb->grain : jb->n_pairs;
        }
        int cnt = 0; uint32_t maxq = 0, maxr = 0;
        while (k < hi && cnt < L) {
            const uint64_t n1 = jb->qo[k + 1] - jb->qo[k], n2 = jb->ro[k + 1] - jb->ro[k];
            const uint64_t mn = n1 < n2 ? n1 : n2;
            if (isa == 0 || n1 > 32000 || n2 > 32000 || 2 * mn > 32000) {   /* int16 lanes cannot hold it */
                if (sw_linear(jb->q + jb->qo[k], n1, jb->r + jb->ro[k], n2, &jb->out[k]) != 0) jb->rc = -1;
            } else if (mn == 0) {
                jb->out[k].score = 0; jb->out[k].end_i = -1; jb->out[k].end_j = -1;
            } else {
                idx[cnt++] = k;
                if (n1 > maxq) maxq = (uint32_t)n1;
                if (n2 > maxr) maxr = (uint32_t)n2;
            }
            ++k;
        }
        if (!cnt) continue;
        if (maxq > cap_q || maxr > cap_r) {
            free(qt); free(rt); free(hrow);
            cap_q = maxq > cap_q ? maxq : cap_q; cap_r = maxr > cap_r ? maxr : cap_r;
            qt = (int16_t*)aligned_alloc(64, cap_q * 32 * sizeof(int16_t));
            rt = (int16_t*)aligned_alloc(64, cap_r * 32 * sizeof(int16_t));
            hrow = aligned_alloc(64, (cap_r + 1) * 64);
            if (!qt || !rt || !hrow) { jb->rc = -1; break; }
        }
        if (isa == 2) block_avx512(jb->q, jb->qo, jb->r, jb->ro, idx, cnt, jb->out, qt, rt, (__m512i*)hrow, maxq, maxr);
        else          block_avx2(jb->q, jb->qo, jb->r, jb->ro, idx, cnt, jb->out, qt, rt, (__m256i*)hrow, maxq, maxr);
    }
    free(qt); free(rt); free(hrow);
    return NULL;
}

int sw_simd_batch(const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro,
                  uint64_t n_pairs, sw_result* out, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    const uint64_t blocks = (n_pairs + 31) / 32;
    if ((uint64_t)n_threads > blocks) n_threads = blocks ? (int)blocks : 1;
    /* pieces of whole 32-pair vectors, about 16 per thread, at most 2048 pairs */
    uint64_t grain = (blocks / ((uint64_t)n_threads * 16) + 1) * 32;
    if (grain > 2048) grain = 2048;
    uint64_t next = 0;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    simd_job* jobs = (simd_job*)malloc(sizeof(simd_job) * n_threads);
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (simd_job){ q, qo, r, ro, n_pairs, grain, &next, out, 0 };
        if (t) pthread_create(&th[t], NULL, simd_worker, &jobs[t]);
    }
    simd_worker(&jobs[0]);
    int rc = jobs[0].rc;
    for (int t = 1; t < n_threads; ++t) { pthread_join(th[t], NULL); rc |= jobs[t].rc; }
    free(th); free(jobs);
    return rc;
}
