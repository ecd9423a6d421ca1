// link-only stubs: the TSan run exercises rsm_debug_bgzf_segments, which touches none of them
#include <cstdlib>
extern "C" {
#define STUB(name) int name() { abort(); return 0; }
STUB(swb_bind_thread) STUB(swb_create) STUB(swb_destroy) STUB(swb_device_count) STUB(swb_device_info) STUB(swb_fastq_bgzf_cancel)
STUB(swb_fastq_bgzf_prefetch) STUB(swb_fastq_bgzf_score) STUB(swb_free_pinned) STUB(swb_malloc_pinned) STUB(swb_memory_info)
STUB(swb_ref_compat_align) STUB(swb_score_batch) STUB(swb_score_batch_vs_reference) STUB(swb_score_pair) STUB(swb_set_reference)
const char* swb_last_error() { return ""; }
}
