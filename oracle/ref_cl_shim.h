/*
 * ref_cl_shim.h -- TEST INFRASTRUCTURE.  Force-included (gcc -include) in front of the
 * reference's own OpenCL-C source so that
 *     /root/reference/smith_waterman/src/smith_waterman.cl
 * compiles UNMODIFIED as plain C.  Work-item built-ins are served by the work-group
 * emulator in ref_cl_driver.c (one ucontext fiber per work-item; barrier() yields to
 * the scheduler, which resumes the work-items of a group in ascending local id).
 * The reference source is read where it lies; nothing of it is copied into this repo.
 */
#ifndef REF_CL_SHIM_H
#define REF_CL_SHIM_H
#include <stdint.h>
#include <stddef.h>

typedef unsigned char uchar;
typedef unsigned int  uint;

#define __kernel
#define __global
#define __local static          /* one copy per work-group: groups run one after another */
#define CLK_LOCAL_MEM_FENCE 1

size_t refcl_get_global_id(uint d);
size_t refcl_get_local_id(uint d);
size_t refcl_get_local_size(uint d);
size_t refcl_get_group_id(uint d);
size_t refcl_get_num_groups(uint d);
void   refcl_barrier(int flags);
int    refcl_atomic_max(int* p, int v);

#define get_global_id  refcl_get_global_id
#define get_local_id   refcl_get_local_id
#define get_local_size refcl_get_local_size
#define get_group_id   refcl_get_group_id
#define get_num_groups refcl_get_num_groups
#define barrier        refcl_barrier
#define atomic_max     refcl_atomic_max

/* OpenCL's generic min/max on same-typed integer operands */
#define max(a, b) __extension__({ __typeof__(a) a_ = (a); __typeof__(b) b_ = (b); a_ > b_ ? a_ : b_; })
#define min(a, b) __extension__({ __typeof__(a) a_ = (a); __typeof__(b) b_ = (b); a_ < b_ ? a_ : b_; })

#endif
