//! `extern "C"` declarations of include/ezkvm_prover.h (the part a Rust caller needs).  Plain pointers and sizes.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct ezk_options {
    pub num_queries: u32,
    pub blowup_factor: u32,
    pub grinding_factor: u32,
    pub field_extension: u32,
    pub fri_folding_factor: u32,
    pub fri_remainder_max_degree: u32,
}

#[repr(C)]
pub struct ezk_public_inputs {
    pub program_hash: [[u8; 16]; 2],
    pub stack_outputs: [[u8; 16]; 16],
    pub lwe_k: u32,
    pub lwe_delta: u32,
}

#[repr(C)]
pub struct ezk_trace {
    pub columns: *const *const u8, // `width` pointers, each `length` x 16 little-endian bytes, canonical (< M)
    pub width: u32,
    pub length: u64,
}

/// Byte-level switches of the winterfell 0.9.0 proof format; defaults = current reading.
#[repr(C)]
pub struct ezk_wire_compat {
    pub ood_interleaved: u32,
    pub remainder_low_to_high: u32,
    pub trace_info_aux_rands_byte: u32,
    pub reserved: u32,
    pub first_nonce: u64,
}

#[repr(C)]
pub struct ezk_prover {
    _private: [u8; 0],
}

pub const EZK_OK: c_int = 0;
pub const EZK_ERR_INVALID_ARGUMENT: c_int = -1;
pub const EZK_ERR_UNSUPPORTED_FIELD_EXTENSION: c_int = -2;
pub const EZK_ERR_CONSTRAINT_DEGREE: c_int = -3;
pub const EZK_ERR_DEEP_DEGREE: c_int = -4;
pub const EZK_ERR_NO_DEVICE: c_int = -5;
pub const EZK_ERR_VERIFICATION: c_int = -9;

extern "C" {
    pub fn ezk_last_error() -> *const c_char;
    pub fn ezk_device_count() -> c_int;
    pub fn ezk_free(p: *mut c_void);
    pub fn ezk_set_wire_compat(c: *const ezk_wire_compat);

    /// one-shot: lazily created prover on device 0
    pub fn ezk_prove(trace: *const ezk_trace, public: *const ezk_public_inputs, options: *const ezk_options,
                     proof: *mut *mut u8, proof_len: *mut usize) -> c_int;

    pub fn ezk_prover_create(device: c_int, out: *mut *mut ezk_prover) -> c_int;
    pub fn ezk_prover_destroy(p: *mut ezk_prover);
    pub fn ezk_prover_prove(p: *mut ezk_prover, trace: *const ezk_trace, public: *const ezk_public_inputs,
                            options: *const ezk_options, proof: *mut *mut u8, proof_len: *mut usize) -> c_int;
    pub fn ezk_prover_verify(p: *mut ezk_prover, proof: *const u8, proof_len: usize, public: *const ezk_public_inputs,
                             min_conjectured_security: u32) -> c_int;

    /// one proof sharded over the GPUs of a box: one process per GPU, rank 0 creates the id
    pub fn ezk_comm_unique_id(out: *mut u8) -> c_int; // 128 bytes
    pub fn ezk_prover_join(p: *mut ezk_prover, rank: c_int, world: c_int, unique_id: *const u8) -> c_int;
}
