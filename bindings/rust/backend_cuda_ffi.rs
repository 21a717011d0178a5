//! hephaestus-jit/src/backend/cuda/ffi.rs — `extern "C"` declarations of include/hj.h, hand-written
//! (no bindgen).  One item per C item the backend uses; layouts are `#[repr(C)]` mirrors of the
//! structs in hj.h (field order and widths identical — `hj_abi_version()` is checked at device
//! creation, see mod.rs).
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no cargo / rustc).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_void};

pub const HJ_ABI_VERSION: u32 = 1; // keep in step with hj_abi_version() in csrc/runtime.cpp

#[repr(C)] pub struct hj_device { _private: [u8; 0] }
#[repr(C)] pub struct hj_buffer { _private: [u8; 0] }

// hj_type_kind: VarType in declaration order (vartype.rs:89-104), then the composite kinds
pub const HJ_VOID: u32 = 0;
pub const HJ_BOOL: u32 = 1;
pub const HJ_I8: u32 = 2;
pub const HJ_U8: u32 = 3;
pub const HJ_I16: u32 = 4;
pub const HJ_U16: u32 = 5;
pub const HJ_I32: u32 = 6;
pub const HJ_U32: u32 = 7;
pub const HJ_I64: u32 = 8;
pub const HJ_U64: u32 = 9;
pub const HJ_F16: u32 = 10;
pub const HJ_F32: u32 = 11;
pub const HJ_F64: u32 = 12;
pub const HJ_VEC: u32 = 13;
pub const HJ_ARRAY: u32 = 14;
pub const HJ_MAT: u32 = 15;
pub const HJ_STRUCT: u32 = 16;

// hj_kernel_op: KernelOp in declaration order (op.rs:49-88)
pub const HJ_OP_NOP: u32 = 0;
pub const HJ_OP_SCATTER: u32 = 1;
pub const HJ_OP_SCATTER_REDUCE: u32 = 2;
pub const HJ_OP_SCATTER_ATOMIC: u32 = 3;
pub const HJ_OP_ATOMIC_INC: u32 = 4;
pub const HJ_OP_GATHER: u32 = 5;
pub const HJ_OP_INDEX: u32 = 6;
pub const HJ_OP_LITERAL: u32 = 7;
pub const HJ_OP_EXTRACT: u32 = 8;
pub const HJ_OP_DYN_EXTRACT: u32 = 9;
pub const HJ_OP_CONSTRUCT: u32 = 10;
pub const HJ_OP_SELECT: u32 = 11;
pub const HJ_OP_LOOP_START: u32 = 12;
pub const HJ_OP_LOOP_END: u32 = 13;
pub const HJ_OP_IF_START: u32 = 14;
pub const HJ_OP_IF_END: u32 = 15;
pub const HJ_OP_TEX_LOOKUP: u32 = 16;
pub const HJ_OP_TRACE_RAY: u32 = 17;
pub const HJ_OP_BOP: u32 = 18;
pub const HJ_OP_UOP: u32 = 19;
pub const HJ_OP_FMA: u32 = 20;
pub const HJ_OP_BUFFER_REF: u32 = 21;
pub const HJ_OP_TEXTURE_REF: u32 = 22;
pub const HJ_OP_ACCEL_REF: u32 = 23;

// hj_pass_kind
pub const HJ_PASS_KERNEL: u32 = 0;
pub const HJ_PASS_REDUCE: u32 = 1;
pub const HJ_PASS_PREFIX_SUM: u32 = 2;
pub const HJ_PASS_COMPRESS: u32 = 3;

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct hj_type_desc { pub kind: u32, pub elem: u32, pub num: u32, pub cols: u32, pub rows: u32, pub first_field: u32 }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct hj_ir_var { pub ty: u32, pub op: u32, pub arg: u32, pub dep_start: u32, pub dep_end: u32, pub _pad: u32, pub data: u64 }

#[repr(C)]
pub struct hj_ir {
    pub vars: *const hj_ir_var, pub n_vars: u32,
    pub deps: *const u32, pub n_deps: u32,
    pub types: *const hj_type_desc, pub n_types: u32,
    pub struct_fields: *const u32, pub n_struct_fields: u32,
    pub n_buffers: u32,
}

#[repr(C)]
pub struct hj_pass {
    pub kind: u32, pub arg: u32,
    pub resources: *const u32, pub n_resources: u32,
    pub size_buffer: i32,
    pub ir: *const hj_ir,
    pub size: u64,
}

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct hj_buffer_desc { pub size: u64, pub ty: u32, pub elem_bytes: u32 }

#[repr(C)] #[derive(Clone, Copy)]
pub struct hj_pass_report { pub name: [c_char; 64], pub start_us: f64, pub duration_us: f64 }

#[repr(C)]
pub struct hj_report { pub cpu_duration_us: f64, pub n_passes: u32, pub passes: *mut hj_pass_report, pub passes_capacity: u32 }

extern "C" {
    pub fn hj_last_error() -> *const c_char;
    pub fn hj_abi_version() -> u32;
    pub fn hj_device_create(ordinal: i32, out: *mut *mut hj_device) -> i32;
    pub fn hj_device_retain(dev: *mut hj_device) -> i32;
    pub fn hj_device_release(dev: *mut hj_device) -> i32;
    pub fn hj_buffer_create(dev: *mut hj_device, bytes: usize, out: *mut *mut hj_buffer) -> i32;
    pub fn hj_buffer_create_from_slice(dev: *mut hj_device, data: *const c_void, bytes: usize, out: *mut *mut hj_buffer) -> i32;
    pub fn hj_buffer_create_from_host_async(dev: *mut hj_device, src: *const c_void, bytes: usize, elem_bytes: usize,
                                            out: *mut *mut hj_buffer) -> i32;
    pub fn hj_host_alloc(bytes: usize, out: *mut *mut c_void) -> i32;
    pub fn hj_host_free(p: *mut c_void) -> i32;
    pub fn hj_buffer_retain(buf: *mut hj_buffer) -> i32;
    pub fn hj_buffer_release(buf: *mut hj_buffer) -> i32;
    pub fn hj_buffer_to_host(buf: *mut hj_buffer, offset_bytes: usize, nbytes: usize, dst: *mut c_void) -> i32;
    pub fn hj_execute_graph(dev: *mut hj_device, passes: *const hj_pass, n_passes: u32,
                            env: *const *mut hj_buffer, descs: *const hj_buffer_desc, n_resources: u32,
                            report: *mut hj_report) -> i32;
    pub fn hj_execute_graph_cached(dev: *mut hj_device, graph_key: u64, passes: *const hj_pass, n_passes: u32,
                                   env: *const *mut hj_buffer, descs: *const hj_buffer_desc, n_resources: u32,
                                   how: *mut u32) -> i32;
    pub fn hj_graph_cache_drop(dev: *mut hj_device, graph_key: u64) -> i32;
}
