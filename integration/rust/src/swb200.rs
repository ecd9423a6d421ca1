//! Binding of include/swb200.h.  Return code 0 = Ok, anything else = Err with the text of swb_last_error()
//! (the Err(String) arm of the reference's Result<i32, String>).  One context per device, one thread at a time.
use std::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
pub struct SwbCtx {
    _private: [u8; 0],
}

/// swb_multi: several devices behind one handle (the reference uses devices[0] only, gpu.rs:117-131, main.rs:95).
#[repr(C)]
pub struct SwbMulti {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct SwbResult {
    pub score: i32,
    pub end_i: i32,
    pub end_j: i32,
}

/// swb_alignment: start cell and CIGAR slice behind a result (operations are BAM-style words, length << 4 | op).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct SwbAlignment {
    pub start_i: i32,
    pub start_j: i32,
    pub cigar_len: u32,
    pub status: u32,
    pub cigar_off: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SwbBgzfBlock {
    pub in_off: u64,
    pub in_len: u32,
    pub out_len: u32,
}

extern "C" {
    pub fn swb_device_count() -> c_int;
    pub fn swb_device_info(id: c_int, name: *mut c_char, cap: usize, memory_gb: *mut f64, max_wg: *mut c_int) -> c_int;
    pub fn swb_create(out: *mut *mut SwbCtx, device_id: c_int, params: *const c_void) -> c_int;
    pub fn swb_destroy(ctx: *mut SwbCtx);
    pub fn swb_score_pair(ctx: *mut SwbCtx, s1: *const u8, n1: u64, s2: *const u8, n2: u64, out: *mut SwbResult) -> c_int;
    pub fn swb_score_batch(ctx: *mut SwbCtx, q: *const u8, q_off: *const u64, r: *const u8, r_off: *const u64, n_pairs: u64,
                           out: *mut SwbResult) -> c_int;
    pub fn swb_traceback_batch(ctx: *mut SwbCtx, q: *const u8, q_off: *const u64, r: *const u8, r_off: *const u64, n_pairs: u64,
                               results: *const SwbResult, out: *mut SwbAlignment, cigar: *mut u32, cigar_cap: u64,
                               cigar_used: *mut u64) -> c_int;
    pub fn swb_set_chunk_ramp(ctx: *mut SwbCtx, ramp: c_int) -> c_int;
    pub fn swb_set_reference(ctx: *mut SwbCtx, reference: *const u8, n: u64) -> c_int;
    pub fn swb_score_batch_vs_reference(ctx: *mut SwbCtx, q: *const u8, q_off: *const u64, n_pairs: u64, win_start: *const u64,
                                        win_len: *const u32, out: *mut SwbResult) -> c_int;
    pub fn swb_score_batch_ranges(ctx: *mut SwbCtx, q: *const u8, q_off: *const u64, n_pairs: u64, w_bytes: *const u8, w_total: u64,
                                  win_start: *const u64, win_len: *const u32, out: *mut SwbResult) -> c_int;
    pub fn swb_create_multi(out: *mut *mut SwbMulti, device_ids: *const c_int, n_devices: c_int, params: *const c_void) -> c_int;
    pub fn swb_destroy_multi(m: *mut SwbMulti);
    pub fn swb_multi_device_count(m: *mut SwbMulti) -> c_int;
    pub fn swb_multi_score_batch(m: *mut SwbMulti, q: *const u8, q_off: *const u64, r: *const u8, r_off: *const u64, n_pairs: u64,
                                 out: *mut SwbResult) -> c_int;
    pub fn swb_multi_set_reference(m: *mut SwbMulti, reference: *const u8, n: u64) -> c_int;
    pub fn swb_multi_score_batch_vs_reference(m: *mut SwbMulti, q: *const u8, q_off: *const u64, n_pairs: u64, win_start: *const u64,
                                              win_len: *const u32, out: *mut SwbResult) -> c_int;
    pub fn swb_fastq_bgzf_cancel(ctx: *mut SwbCtx, comp: *const u8) -> c_int;
    pub fn swb_fastq_bgzf_prefetch(ctx: *mut SwbCtx, comp: *const u8, comp_bytes: u64, blocks: *const SwbBgzfBlock, n_blocks: u64) -> c_int;
    pub fn swb_fastq_bgzf_score(ctx: *mut SwbCtx, comp: *const u8, comp_bytes: u64, blocks: *const SwbBgzfBlock, n_blocks: u64,
                                carry: *const u8, carry_len: u64, final_segment: c_int, file_index: u64, first_read: u64,
                                window_len: u32, score_sum: *mut i64, n_reads: *mut u64, n_bases: *mut u64, n_lines: *mut u64,
                                carry_out: *mut u8, carry_cap: u64, carry_out_len: *mut u64, status: *mut c_int) -> c_int;
    pub fn swb_ref_compat_align(ctx: *mut SwbCtx, s1: *const u8, n1: u64, s2: *const u8, n2: u64, max_wg: u32, out: *mut i32) -> c_int;
    pub fn swb_last_row_max(ctx: *mut SwbCtx, s1: *const u8, n1: u64, s2: *const u8, n2: u64, out: *mut i32) -> c_int;
    pub fn swb_set_chunking(ctx: *mut SwbCtx, chunk_bytes: u64, min_chunk_pairs: u64) -> c_int;
    pub fn swb_last_routing(ctx: *mut SwbCtx, counts: *mut u64) -> c_int;
    pub fn swb_malloc_pinned(bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn swb_bind_thread(ctx: *mut SwbCtx) -> c_int;          // a helper thread that page-locks buffers for `ctx` calls this first
    pub fn swb_debug_guard_check(ctx: *mut SwbCtx, report: *mut c_char, report_cap: u64, n_arenas: *mut u64) -> c_int;   // SWB_GUARD=1
    pub fn swb_free_pinned(p: *mut c_void) -> c_int;
    pub fn swb_last_error() -> *const c_char;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(swb_last_error()) }.to_string_lossy().into_owned()
}

/// One GPU.  Replaces the OpenCL (Context, Queue, Device) singleton of gpu.rs:13-14.
pub struct Engine {
    ctx: *mut SwbCtx,
}

unsafe impl Send for Engine {}

impl Engine {
    pub fn new(device_id: i32) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        if unsafe { swb_create(&mut ctx, device_id, std::ptr::null()) } != 0 {
            return Err(last_error()); // "error: gpu acceleration is required and no compatible gpu was found" without a GPU
        }
        Ok(Self { ctx })
    }

    /// Best local score and its end cell; (0, -1, -1) for empty input (aligner.rs:413-416 returns Ok(0)).
    pub fn score_pair(&mut self, seq1: &[u8], seq2: &[u8]) -> Result<SwbResult, String> {
        let mut r = SwbResult::default();
        let rc = unsafe { swb_score_pair(self.ctx, seq1.as_ptr(), seq1.len() as u64, seq2.as_ptr(), seq2.len() as u64, &mut r) };
        if rc != 0 { Err(last_error()) } else { Ok(r) }
    }

    /// The C ABI trusts CSR offsets to stay inside the byte slice: a safe function checks that before it hands out pointers.
    fn check_csr(bytes: &[u8], off: &[u64], what: &str) -> Result<(), String> {
        if off.is_empty() || off[0] != 0 { return Err(format!("{what}: offsets must start at 0")); }
        if off.windows(2).any(|w| w[1] < w[0]) { return Err(format!("{what}: offsets must be non-decreasing")); }
        if *off.last().unwrap() > bytes.len() as u64 { return Err(format!("{what}: the last offset is past the end of the byte slice")); }
        Ok(())
    }

    /// One chunk of reads (what process_fastq_file_in_chunks hands its callback) against their windows.
    pub fn score_batch(&mut self, reads: &[u8], read_off: &[u64], windows: &[u8], window_off: &[u64]) -> Result<Vec<SwbResult>, String> {
        if read_off.len() != window_off.len() { return Err("read_off and window_off must describe the same number of pairs".into()); }
        Self::check_csr(reads, read_off, "reads")?;
        Self::check_csr(windows, window_off, "windows")?;
        let n = read_off.len().saturating_sub(1);
        let mut out = vec![SwbResult::default(); n];
        let rc = unsafe {
            swb_score_batch(self.ctx, reads.as_ptr(), read_off.as_ptr(), windows.as_ptr(), window_off.as_ptr(), n as u64, out.as_mut_ptr())
        };
        if rc != 0 { Err(last_error()) } else { Ok(out) }
    }

    /// Reads against windows that are ranges of ONE buffer (candidate windows of a genome: they overlap); the covered part of
    /// the buffer is uploaded once per call.  Same results as `score_batch` on the materialised windows.
    pub fn score_batch_ranges(&mut self, reads: &[u8], read_off: &[u64], buffer: &[u8], win_start: &[u64], win_len: &[u32])
                              -> Result<Vec<SwbResult>, String> {
        Self::check_csr(reads, read_off, "reads")?;
        let n = read_off.len() - 1;
        if win_start.len() != n || win_len.len() != n { return Err("read_off, win_start and win_len must describe the same number of pairs".into()); }
        let mut out = vec![SwbResult::default(); n];
        let rc = unsafe {
            swb_score_batch_ranges(self.ctx, reads.as_ptr(), read_off.as_ptr(), n as u64, buffer.as_ptr(), buffer.len() as u64,
                                   win_start.as_ptr(), win_len.as_ptr(), out.as_mut_ptr())
        };
        if rc != 0 { Err(last_error()) } else { Ok(out) }
    }

    /// Start cell and CIGAR behind the results of `score_batch` on the same pairs; grows the operation buffer once if the
    /// first guess was too small (the library reports how much the batch needs).
    pub fn traceback_batch(&mut self, reads: &[u8], read_off: &[u64], windows: &[u8], window_off: &[u64], results: &[SwbResult])
                           -> Result<(Vec<SwbAlignment>, Vec<u32>), String> {
        let n = results.len();
        if read_off.len() != n + 1 || window_off.len() != n + 1 { return Err("offsets and results must describe the same number of pairs".into()); }
        Self::check_csr(reads, read_off, "reads")?;
        Self::check_csr(windows, window_off, "windows")?;
        let mut out = vec![SwbAlignment::default(); n];
        let mut cigar = vec![0u32; 8 * n + 1024];
        let mut used = 0u64;
        for _ in 0..2 {
            let rc = unsafe {
                swb_traceback_batch(self.ctx, reads.as_ptr(), read_off.as_ptr(), windows.as_ptr(), window_off.as_ptr(), n as u64,
                                    results.as_ptr(), out.as_mut_ptr(), cigar.as_mut_ptr(), cigar.len() as u64, &mut used)
            };
            if rc == 0 {
                cigar.truncate(used as usize);
                return Ok((out, cigar));
            }
            if used as usize <= cigar.len() { break; }
            cigar.resize(used as usize, 0);
        }
        Err(last_error())
    }

    /// What the reference's gpu_align returns today (its live kernel): 2 if any aligned position matches, else 0.
    pub fn ref_compat_align(&mut self, seq1: &[u8], seq2: &[u8], max_work_group: u32) -> Result<i32, String> {
        let mut v = 0i32;
        let rc = unsafe {
            swb_ref_compat_align(self.ctx, seq1.as_ptr(), seq1.len() as u64, seq2.as_ptr(), seq2.len() as u64, max_work_group, &mut v)
        };
        if rc != 0 { Err(last_error()) } else { Ok(v) }
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { swb_destroy(self.ctx) }
    }
}

/// Same signature as the reference (aligner.rs:410); `engine` replaces `&GpuDevice` + the global OpenCL context.
pub fn gpu_align(seq1: &str, seq2: &str, engine: &mut Engine) -> Result<i32, String> {
    if seq1.is_empty() || seq2.is_empty() {
        return Ok(0);
    }
    engine.score_pair(seq1.as_bytes(), seq2.as_bytes()).map(|r| r.score)
}


/// Every GPU of the box behind one handle: a batch is cut into contiguous slices, one per device, each written by its own
/// host thread into its slice of the result vector.
pub struct MultiEngine {
    m: *mut SwbMulti,
}

unsafe impl Send for MultiEngine {}

impl MultiEngine {
    /// `devices` empty = every visible device.
    pub fn new(devices: &[i32]) -> Result<Self, String> {
        let mut m: *mut SwbMulti = std::ptr::null_mut();
        let ids = if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() };
        if unsafe { swb_create_multi(&mut m, ids, devices.len() as c_int, std::ptr::null()) } != 0 { return Err(last_error()); }
        Ok(Self { m })
    }

    pub fn n_devices(&self) -> usize { unsafe { swb_multi_device_count(self.m) as usize } }

    pub fn set_reference(&mut self, reference: &[u8]) -> Result<(), String> {
        if unsafe { swb_multi_set_reference(self.m, reference.as_ptr(), reference.len() as u64) } != 0 { Err(last_error()) } else { Ok(()) }
    }

    pub fn score_batch(&mut self, reads: &[u8], read_off: &[u64], windows: &[u8], window_off: &[u64]) -> Result<Vec<SwbResult>, String> {
        if read_off.len() != window_off.len() { return Err("read_off and window_off must describe the same number of pairs".into()); }
        Engine::check_csr(reads, read_off, "reads")?;
        Engine::check_csr(windows, window_off, "windows")?;
        let n = read_off.len() - 1;
        let mut out = vec![SwbResult::default(); n];
        let rc = unsafe {
            swb_multi_score_batch(self.m, reads.as_ptr(), read_off.as_ptr(), windows.as_ptr(), window_off.as_ptr(), n as u64, out.as_mut_ptr())
        };
        if rc != 0 { Err(last_error()) } else { Ok(out) }
    }

    pub fn score_batch_vs_reference(&mut self, reads: &[u8], read_off: &[u64], win_start: &[u64], win_len: &[u32]) -> Result<Vec<SwbResult>, String> {
        Engine::check_csr(reads, read_off, "reads")?;
        let n = read_off.len() - 1;
        if win_start.len() != n || win_len.len() != n { return Err("read_off, win_start and win_len must describe the same number of pairs".into()); }
        let mut out = vec![SwbResult::default(); n];
        let rc = unsafe {
            swb_multi_score_batch_vs_reference(self.m, reads.as_ptr(), read_off.as_ptr(), n as u64, win_start.as_ptr(), win_len.as_ptr(),
                                               out.as_mut_ptr())
        };
        if rc != 0 { Err(last_error()) } else { Ok(out) }
    }
}

impl Drop for MultiEngine {
    fn drop(&mut self) {
        unsafe { swb_destroy_multi(self.m) }
    }
}
