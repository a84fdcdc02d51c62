//! Drop-in replacement for the reference's src/alignments.rs on top of libapd_b200.
//!
//! Same public surface (file:line in the reference):
//!   AlignmentWorkers { data, result }, ::new, ::align_all      src/alignments.rs:11-68
//!   AlignmentParams { .. }, ::default                          src/alignments.rs:77-94
//!   Alignment { n, m, sparse }, ::new, ::score, ::construct_alignment   src/alignments.rs:99-181
//! so that src/main.rs:187-200 and src/discovery.rs:38-45 compile unchanged.
//!
//! SOURCE ONLY: no cargo/rustc in this repository's build image -- reviewed, not compiled.
//! Differences a caller can observe, all deliberate:
//!   * align_all no longer prints one line per row (src/alignments.rs:43-49); it prints one
//!     summary line (pairs, cells, seconds, GCUPS) instead;
//!   * `Alignment.sparse` is kept for source compatibility but is not filled with every
//!     visited cell (the reference never reads them); the traced warping path is in `path`;
//!   * a failure of the library is a panic with the library's message (the reference's
//!     worker panics are swallowed by `let _ = child.join()` and leave zeros behind).
use crate::discovery::Discovery;
use crate::spectrogram::NDSequence;
use apd_sys as ffi;
use std::collections::HashMap;
use std::ffi::CStr;
use std::ptr;
use std::sync::{Arc, Mutex};

fn check(ctx: *mut ffi::apd_ctx, status: i32, what: &str) {
    if status != ffi::APD_OK {
        let msg = unsafe { CStr::from_ptr(ffi::apd_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("{} failed (status {}): {}", what, status, msg);
    }
}

struct Context(*mut ffi::apd_ctx);

// The handle is only ever used by one thread at a time (behind the Mutex below).
unsafe impl Send for Context {}

impl Context {
    /// Every visible GPU of the box in ONE context (apd_create_multi with n_dev = 0): the
    /// reference's one blocking `align_all` call (src/main.rs:189-195) fans out inside the library.
    fn all_devices() -> Context {
        let mut raw: *mut ffi::apd_ctx = ptr::null_mut();
        let st = unsafe { ffi::apd_create_multi(ptr::null(), 0, &mut raw) };
        check(ptr::null_mut(), st, "apd_create_multi");
        Context(raw)
    }

    /// One GPU is plenty for single pairs.
    fn one_device() -> Context {
        let mut raw: *mut ffi::apd_ctx = ptr::null_mut();
        let st = unsafe { ffi::apd_create(0, &mut raw) };
        check(ptr::null_mut(), st, "apd_create");
        Context(raw)
    }

    /// The spectrogram.rs glue: hands the library one pointer + length per NDSequence
    /// (frames are row-major T x n_bins, src/spectrogram.rs:13-24, len() 152-154).  The
    /// reference's euclidean() silently assumes equal widths (src/numerics.rs:114-120 iterates
    /// x.len()); a mismatch would read past a buffer here, so it is a panic.
    fn set_sequences(&self, data: &[&NDSequence]) {
        let dim = data.first().map(|s| s.n_bins).unwrap_or(1);
        for (k, s) in data.iter().enumerate() {
            assert!(s.n_bins == dim, "sequence {} has n_bins {} but sequence 0 has {}", k, s.n_bins, dim);
            assert!(s.frames.len() >= s.len() * s.n_bins, "sequence {} is shorter than len() * n_bins", k);
        }
        let ptrs: Vec<*const f32> = data.iter().map(|s| s.frames.as_ptr()).collect();
        let lens: Vec<u32> = data.iter().map(|s| s.len() as u32).collect();
        let st = unsafe { ffi::apd_set_sequences(self.0, ptrs.as_ptr(), lens.as_ptr(), data.len() as u32, dim as u32) };
        check(self.0, st, "apd_set_sequences");
    }
}

/// One single-GPU context shared by every `Alignment::construct_alignment` call of the process
/// (creating a CUDA context per pair would cost more than the pair).
fn pair_context() -> &'static Mutex<Context> {
    use std::sync::Once;
    static INIT: Once = Once::new();
    static mut CTX: Option<Mutex<Context>> = None;
    unsafe {
        INIT.call_once(|| CTX = Some(Mutex::new(Context::one_device())));
        CTX.as_ref().unwrap()
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::apd_destroy(self.0) }
    }
}

/// Aligns all sequences and saves the results in a flat matrix (src/alignments.rs:11-14).
pub struct AlignmentWorkers {
    pub data: Arc<Vec<NDSequence>>,
    pub result: Arc<Mutex<Vec<f32>>>,
}

impl AlignmentWorkers {
    pub fn new(data: Vec<NDSequence>) -> AlignmentWorkers {
        let n = data.len();
        AlignmentWorkers { data: Arc::from(data), result: Arc::from(Mutex::from(vec![0.0; n * n])) }
    }

    /// `alignment_workers` is accepted and ignored (the GPU spreads the work); 0 still
    /// panics like the reference's division at src/alignments.rs:33.
    pub fn align_all(&mut self, params: &Discovery) {
        let n = self.data.len();
        let _batch_size = (n / params.alignment_workers) + 1;
        let ctx = Context::all_devices();
        let refs: Vec<&NDSequence> = self.data.iter().collect();
        ctx.set_sequences(&refs);
        let p = ffi::apd_params {
            warping_band_percentage: params.warping_band_percentage,
            insertion_penalty: params.insertion_penalty,
            deletion_penalty: params.deletion_penalty,
            match_penalty: params.match_penalty,
            mode: ffi::APD_MODE_STRICT,
        };
        let mut result = self.result.lock().unwrap();
        let st = unsafe { ffi::apd_align_all(ctx.0, &p, result.as_mut_ptr()) };
        check(ctx.0, st, "apd_align_all");
        let mut stats = ffi::apd_stats::default();
        unsafe { ffi::apd_get_stats(ctx.0, &mut stats) };
        println!(
            "Aligned {} ordered pairs, {} cells in {:.3} s on the GPU ({:.1} GCUPS)",
            stats.ordered_pairs,
            stats.cells_reference,
            stats.kernel_ms / 1e3,
            stats.cells_reference as f64 / (stats.kernel_ms as f64 * 1e6)
        );
    }
}

#[derive(Clone, Debug)]
pub struct AlignmentParams {
    pub warping_band: usize,
    pub insertion_penalty: f32,
    pub deletion_penalty: f32,
    pub match_penalty: f32,
}

impl AlignmentParams {
    pub fn default(len: usize) -> AlignmentParams {
        AlignmentParams { warping_band: len, insertion_penalty: 1.0, deletion_penalty: 1.0, match_penalty: 1.0 }
    }
}

#[derive(Debug)]
pub struct Alignment {
    pub n: usize,
    pub m: usize,
    pub sparse: HashMap<(usize, usize), f32>,
    /// warping path (i, j), 1-based, end to start
    pub path: Vec<(usize, usize)>,
    /// the score computed on the device (already divided by n + m, bit-exact)
    device_score: Option<f32>,
}

impl Alignment {
    pub fn new() -> Alignment {
        let mut sparse = HashMap::new();
        sparse.insert((0, 0), 0.0);
        Alignment { n: 0, m: 0, sparse, path: vec![], device_score: None }
    }

    pub fn score(&self) -> f32 {
        if let Some(s) = self.device_score {
            return s;
        }
        if self.m == 0 && self.n == 0 {
            std::f32::INFINITY
        } else {
            match self.sparse.get(&(self.n.wrapping_sub(1), self.m.wrapping_sub(1))) {
                Some(score) => score / (self.n + self.m) as f32,
                None => std::f32::INFINITY,
            }
        }
    }

    pub fn construct_alignment(&mut self, x: &NDSequence, y: &NDSequence, params: &AlignmentParams) {
        self.n = x.len();
        self.m = y.len();
        let ctx = pair_context().lock().unwrap();
        ctx.set_sequences(&[x, y]);
        let p = ffi::apd_params {
            warping_band_percentage: 0.0,
            insertion_penalty: params.insertion_penalty,
            deletion_penalty: params.deletion_penalty,
            match_penalty: params.match_penalty,
            mode: ffi::APD_MODE_STRICT,
        };
        let cap = (self.n + self.m + 2) as u64;
        let mut path = vec![0u32; 2 * cap as usize];
        let (mut score, mut plen) = (0f32, 0u64);
        let pair = [0u32, 1u32];
        let st = unsafe {
            ffi::apd_align_pairs_band(ctx.0, &p, params.warping_band as u64, pair.as_ptr(), 1, &mut score,
                                      path.as_mut_ptr(), cap, &mut plen)
        };
        check(ctx.0, st, "apd_align_pairs_band");
        self.device_score = Some(score);
        self.path = (0..plen as usize).map(|k| (path[2 * k] as usize, path[2 * k + 1] as usize)).collect();
    }
}
