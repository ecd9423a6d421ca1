/*
 * ref_cl_driver.c -- TEST INFRASTRUCTURE.  A minimal OpenCL work-group emulator, just
 * enough to EXECUTE the reference's own kernel source (compiled unmodified through
 * ref_cl_shim.h) on the CPU:
 *   smith_waterman_align     (smith_waterman.cl:11-71,  the kernel gpu_align launches)
 *   smith_waterman_detailed  (smith_waterman.cl:74-151, never launched; holds the DP)
 *
 * Execution model: work-groups run one after another; inside a group every work-item
 * is a ucontext fiber; the scheduler resumes the live work-items in ascending local id
 * and a work-item runs until it calls barrier() or returns.  OpenCL leaves the order
 * of work-items between two barriers undefined, so this is ONE legal schedule.  For
 * smith_waterman_detailed it is the schedule under which the unsynchronised read of
 * row_scores[j-1] (cl:117) sees the value written by work-item j-1 in the same row,
 * i.e. the sequential Smith-Waterman recurrence (needs local_size >= len2 so that each
 * work-item owns exactly one column, and len2 <= 256 because of cl:93-94).
 *
 * Limits inherited from the source: local arrays are int[256], so local_size <= 256
 * here (the reference indexes local_scores[local_id] with up to 1024 work-items,
 * cl:23/56 vs gpu.rs:9 -- out of bounds, not reproduced).
 */
#define _GNU_SOURCE
#include <ucontext.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char uchar;
typedef unsigned int  uint;

void smith_waterman_align(const uchar* seq1, const uchar* seq2, int* result, uint length);
void smith_waterman_detailed(const uchar* seq1, const uchar* seq2, int* result, uint len1, uint len2);

#define FIBER_STACK (64 * 1024)

static struct {
    size_t local_size, num_groups, group_id, cur_lid;
    ucontext_t sched;
    ucontext_t* fib;
    char* stacks;
    int* done;
    /* kernel arguments */
    int which;
    const uchar* a; const uchar* b; int* result; uint n1, n2;
} E;

size_t refcl_get_global_id(uint d)  { (void)d; return E.group_id * E.local_size + E.cur_lid; }
size_t refcl_get_local_id(uint d)   { (void)d; return E.cur_lid; }
size_t refcl_get_local_size(uint d) { (void)d; return E.local_size; }
size_t refcl_get_group_id(uint d)   { (void)d; return E.group_id; }
size_t refcl_get_num_groups(uint d) { (void)d; return E.num_groups; }
int    refcl_atomic_max(int* p, int v) { int old = *p; if (v > old) *p = v; return old; }

void refcl_barrier(int flags)
{
    (void)flags;
    size_t me = E.cur_lid;
    swapcontext(&E.fib[me], &E.sched);      /* yield; resumed with cur_lid == me */
}

static void fiber_main(void)
{
    if (E.which == 0) smith_waterman_align(E.a, E.b, E.result, E.n1);
    else              smith_waterman_detailed(E.a, E.b, E.result, E.n1, E.n2);
    E.done[E.cur_lid] = 1;
    /* returns to uc_link == scheduler */
}

static int run_ndrange(size_t local_size, size_t num_groups)
{
    if (local_size == 0 || local_size > 256) return -1;
    E.local_size = local_size; E.num_groups = num_groups;
    E.fib    = (ucontext_t*)calloc(local_size, sizeof(ucontext_t));
    E.stacks = (char*)malloc(local_size * FIBER_STACK);
    E.done   = (int*)calloc(local_size, sizeof(int));
    if (!E.fib || !E.stacks || !E.done) return -1;
    for (size_t g = 0; g < num_groups; ++g) {
        E.group_id = g;
        for (size_t l = 0; l < local_size; ++l) {
            getcontext(&E.fib[l]);
            E.fib[l].uc_stack.ss_sp = E.stacks + l * FIBER_STACK;
            E.fib[l].uc_stack.ss_size = FIBER_STACK;
            E.fib[l].uc_link = &E.sched;
            makecontext(&E.fib[l], fiber_main, 0);
            E.done[l] = 0;
        }
        size_t live = local_size;
        while (live) {
            live = 0;
            for (size_t l = 0; l < local_size; ++l) {
                if (E.done[l]) continue;
                E.cur_lid = l;
                swapcontext(&E.sched, &E.fib[l]);
                if (!E.done[l]) ++live;
            }
        }
    }
    free(E.fib); free(E.stacks); free(E.done);
    return 0;
}

/* smith_waterman_align with an explicit NDRange; result buffer zero-initialised. */
int refcl_run_align(const uint8_t* s1, const uint8_t* s2, uint32_t length,
                    uint32_t local_size, uint32_t num_groups, int32_t* out)
{
    int result = 0;
    E.which = 0; E.a = s1; E.b = s2; E.result = &result; E.n1 = length; E.n2 = 0;
    if (run_ndrange(local_size, num_groups) != 0) return -1;
    *out = result;
    return 0;
}

/* gpu_align's launch geometry (aligner.rs:412-424, :510-525) around smith_waterman_align. */
int refcl_gpu_align(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2,
                    uint32_t dev_max_wg, int32_t* out)
{
    uint64_t len = n1 < n2 ? n1 : n2;
    if (len == 0) { *out = 0; return 0; }                    /* aligner.rs:413-416 */
    uint64_t wgs = dev_max_wg < 1024 ? dev_max_wg : 1024;    /* aligner.rs:422 */
    uint64_t groups = (len + wgs - 1) / wgs;                 /* aligner.rs:423 */
    if (groups > 1000000) groups = 1000000;                  /* aligner.rs:424 */
    return refcl_run_align(s1, s2, (uint32_t)len, (uint32_t)wgs, (uint32_t)groups, out);
}

/* smith_waterman_detailed as ONE work-group (row == 0 initialises row_scores, cl:97-102);
 * every further group would only repeat the same whole-matrix computation (cl:105). */
int refcl_run_detailed(const uint8_t* s1, uint32_t len1, const uint8_t* s2, uint32_t len2,
                       uint32_t local_size, int32_t* out)
{
    int result = 0;
    if (len1 == 0 || len2 == 0 || len2 > 256 || local_size < len2) return -1;
    E.which = 1; E.a = s1; E.b = s2; E.result = &result; E.n1 = len1; E.n2 = len2;
    if (run_ndrange(local_size, 1) != 0) return -1;
    *out = result;
    return 0;
}
