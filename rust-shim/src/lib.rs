//! GPU drop-in for `prover::ExecutionProver::prove` (reference: prover/src/lib.rs:40-77, called at vm/src/lib.rs:26).
//!
//! `GpuExecutionProver::new` takes exactly the arguments of `ExecutionProver::new` (prover/src/lib.rs:25-37) and
//! `prove` takes the `TraceTable<BaseElement>` that `vm::prove` builds (vm/src/lib.rs:18); the returned
//! `winterfell::Proof` is parsed from the bytes the library wrote, so the unchanged verifier call
//! (vm/src/lib.rs:91-98, examples/linear_regression/src/main.rs:81-85) keeps working.
//!
//! In `vm/src/lib.rs`, line 24-26 become:
//! ```ignore
//! let prover = prover_gpu::GpuExecutionProver::new(options, hash.to_elements(), output, inputs.server_key());
//! let proof = prover.prove(trace).unwrap();
//! ```
pub mod ffi;

use fhe::ServerKey;
use winterfell::{
    math::{fields::f128::BaseElement, StarkField},
    Proof, ProofOptions, ProverError, Trace, TraceTable,
};

pub struct GpuExecutionProver<'a> {
    options: ProofOptions,
    program_hash: [BaseElement; 2],
    stack_outputs: [BaseElement; 16],
    server_key: &'a ServerKey,
    handle: *mut ffi::ezk_prover,
}

impl<'a> GpuExecutionProver<'a> {
    /// Same signature as `ExecutionProver::new`; the prover lives on CUDA device 0 and keeps its workspace.
    pub fn new(options: ProofOptions, program_hash: [BaseElement; 2], stack_outputs: [BaseElement; 16],
               server_key: &'a ServerKey) -> Self {
        let mut handle = core::ptr::null_mut();
        let rc = unsafe { ffi::ezk_prover_create(0, &mut handle) };
        assert_eq!(rc, ffi::EZK_OK, "{}", last_error()); // EZK_ERR_NO_DEVICE: there is no CPU fallback
        Self { options, program_hash, stack_outputs, server_key, handle }
    }

    fn public_inputs(&self) -> ffi::ezk_public_inputs {
        let mut pi = ffi::ezk_public_inputs {
            program_hash: [[0; 16]; 2],
            stack_outputs: [[0; 16]; 16],
            // fhe/src/parameters.rs:4-21 (consumed by air/src/constrains.rs:113,129-130,151)
            lwe_k: self.server_key.parameters.k as u32,
            lwe_delta: self.server_key.parameters.delta as u32,
        };
        for (d, s) in pi.program_hash.iter_mut().zip(self.program_hash) {
            *d = s.as_int().to_le_bytes();
        }
        for (d, s) in pi.stack_outputs.iter_mut().zip(self.stack_outputs) {
            *d = s.as_int().to_le_bytes();
        }
        pi
    }

    fn c_options(&self) -> ffi::ezk_options {
        let o = &self.options;
        ffi::ezk_options {
            num_queries: o.num_queries() as u32,
            blowup_factor: o.blowup_factor() as u32,
            grinding_factor: o.grinding_factor(),
            field_extension: o.field_extension() as u32, // FieldExtension::None = 1
            fri_folding_factor: 8,                       // vm/src/lib.rs:20
            fri_remainder_max_degree: 127,
        }
    }

    /// `winterfell::Prover::prove` for `ProcessorAir`, on the GPU.
    pub fn prove(&self, trace: TraceTable<BaseElement>) -> Result<Proof, ProverError> {
        // BaseElement is a transparent u128: a column is `length` x 16 little-endian bytes, canonical (< M).  The
        // columns are ordinary pageable Vecs; the library stages them through its own page-locked ring.
        let cols: Vec<*const u8> = (0..trace.width()).map(|c| trace.get_column(c).as_ptr() as *const u8).collect();
        let t = ffi::ezk_trace { columns: cols.as_ptr(), width: trace.width() as u32, length: trace.length() as u64 };
        let (pi, opt) = (self.public_inputs(), self.c_options());
        let (mut p, mut n) = (core::ptr::null_mut(), 0usize);
        let rc = unsafe { ffi::ezk_prover_prove(self.handle, &t, &pi, &opt, &mut p, &mut n) };
        if rc != ffi::EZK_OK {
            return Err(match rc {
                ffi::EZK_ERR_UNSUPPORTED_FIELD_EXTENSION => {
                    ProverError::UnsupportedFieldExtension(self.options.field_extension().degree() as usize)
                }
                ffi::EZK_ERR_CONSTRAINT_DEGREE => {
                    ProverError::MismatchedConstraintPolynomialDegree(7 * trace.length() - 1, 8 * trace.length())
                }
                // EZK_ERR_DEEP_DEGREE is winterfell's `assert_eq!(trace_length - 2, deep_poly.degree())` (a panic there)
                _ => panic!("{}", last_error()),
            });
        }
        let bytes = unsafe { std::slice::from_raw_parts(p, n) }.to_vec();
        unsafe { ffi::ezk_free(p as *mut _) };
        Ok(Proof::from_bytes(&bytes).expect("libezkvm returned a malformed proof"))
    }

    /// Shards every following `prove` over `world` GPUs of one box (one process per GPU; `unique_id` comes from rank
    /// 0's `ezk_comm_unique_id`).  All ranks must then call `prove` with the same trace; all get the same proof.
    pub fn join(&self, rank: i32, world: i32, unique_id: &[u8; 128]) {
        let rc = unsafe { ffi::ezk_prover_join(self.handle, rank, world, unique_id.as_ptr()) };
        assert_eq!(rc, ffi::EZK_OK, "{}", last_error());
    }
}

impl Drop for GpuExecutionProver<'_> {
    fn drop(&mut self) {
        unsafe { ffi::ezk_prover_destroy(self.handle) }
    }
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(ffi::ezk_last_error()) }.to_string_lossy().into_owned()
}
