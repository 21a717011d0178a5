//! hephaestus-jit/src/backend/cuda/mod.rs — the CUDA backend over libhj_b200.so.
//!
//! Replaces the `todo!()` stub of the reference (backend/cuda/mod.rs:16-73): `CudaDevice` /
//! `CudaBuffer` implement `BackendDevice` / `BackendBuffer` (backend/mod.rs:26-49) by calling the C
//! ABI of include/hj.h.  Nothing above the backend traits changes; `Device::cuda(id)` already
//! exists (backend/mod.rs:73-75).  Three edits outside this directory complete the slot:
//!   * backend/mod.rs:58-62   add `#[error("CUDA backend: {msg} (status {code})")] CudaError { code: i32, msg: String }`
//!   * backend/mod.rs:118     `Device::CudaDevice(device) => device.execute_graph(graph, env),`
//!   * backend/mod.rs:104,112 stay `todo!()` (textures / acceleration structures are out of scope)
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no cargo / rustc): this is the
//! source a maintainer drops into the reference tree.  Every call it makes is exercised here through
//! the ctypes mirror (hephaestus-jit_b200/__init__.py) and the C++ restatement of `launch_with`
//! (csrc/tgraph.cpp), which flattens passes exactly as `execute_graph` below does.
mod ffi; // = bindings/rust/backend_cuda_ffi.rs, installed as backend/cuda/ffi.rs

use std::collections::HashMap;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::time::Duration;

use crate::backend::{self, AccelDesc, BackendBuffer, BackendDevice, ExecReport, PassReport, Report};
use crate::graph::{Env, Graph, PassOp, ResourceId};
use crate::ir::IR;
use crate::op::{DeviceOp, KernelOp};
use crate::vartype::{AsVarType, VarType};

fn check(code: i32) -> backend::Result<()> {
    if code == 0 {
        return Ok(());
    }
    // hj_last_error() is thread-local and valid until the next call on this thread
    let msg = unsafe { CStr::from_ptr(ffi::hj_last_error()) }.to_string_lossy().into_owned();
    Err(backend::Error::CudaError { code, msg })
}

// ---- handles ------------------------------------------------------------------------------------
// hj_device / hj_buffer are intrusively ref-counted and internally synchronised (one stream per
// device, calls serialised by the library), which is what `Clone + Send + Sync` asks for.

#[derive(Debug)]
pub struct CudaDevice(*mut ffi::hj_device);
unsafe impl Send for CudaDevice {}
unsafe impl Sync for CudaDevice {}
impl Clone for CudaDevice {
    fn clone(&self) -> Self {
        unsafe { ffi::hj_device_retain(self.0) };
        Self(self.0)
    }
}
impl Drop for CudaDevice {
    fn drop(&mut self) {
        unsafe { ffi::hj_device_release(self.0) };
    }
}
impl CudaDevice {
    pub fn create(id: usize) -> backend::Result<Self> {
        assert_eq!(unsafe { ffi::hj_abi_version() }, ffi::HJ_ABI_VERSION, "libhj_b200.so has another ABI version");
        let mut dev = std::ptr::null_mut();
        check(unsafe { ffi::hj_device_create(id as i32, &mut dev) })?;
        Ok(Self(dev))
    }
}

#[derive(Debug)]
pub struct CudaBuffer {
    buf: *mut ffi::hj_buffer,
    device: CudaDevice,
}
unsafe impl Send for CudaBuffer {}
unsafe impl Sync for CudaBuffer {}
impl Clone for CudaBuffer {
    fn clone(&self) -> Self {
        unsafe { ffi::hj_buffer_retain(self.buf) };
        Self { buf: self.buf, device: self.device.clone() }
    }
}
impl Drop for CudaBuffer {
    fn drop(&mut self) {
        unsafe { ffi::hj_buffer_release(self.buf) }; // back to the stream-ordered pool, contents stale (core/pool.rs:36-41)
    }
}

// ---- IR flattening: ir::IR (ir.rs:40-46) -> hj_ir ----------------------------------------------------
// Owned arrays the hj_ir view points into; must outlive the hj_execute_graph call.
#[derive(Default)]
struct FlatIr {
    vars: Vec<ffi::hj_ir_var>,
    deps: Vec<u32>,
    types: Vec<ffi::hj_type_desc>,
    struct_fields: Vec<u32>,
    n_buffers: u32,
}
impl FlatIr {
    fn view(&self) -> ffi::hj_ir {
        ffi::hj_ir {
            vars: self.vars.as_ptr(), n_vars: self.vars.len() as u32,
            deps: self.deps.as_ptr(), n_deps: self.deps.len() as u32,
            types: self.types.as_ptr(), n_types: self.types.len() as u32,
            struct_fields: self.struct_fields.as_ptr(), n_struct_fields: self.struct_fields.len() as u32,
            n_buffers: self.n_buffers,
        }
    }
}

fn scalar_kind(ty: &VarType) -> u32 {
    match ty {
        VarType::Void => ffi::HJ_VOID,
        VarType::Bool => ffi::HJ_BOOL,
        VarType::I8 => ffi::HJ_I8,
        VarType::U8 => ffi::HJ_U8,
        VarType::I16 => ffi::HJ_I16,
        VarType::U16 => ffi::HJ_U16,
        VarType::I32 => ffi::HJ_I32,
        VarType::U32 => ffi::HJ_U32,
        VarType::I64 => ffi::HJ_I64,
        VarType::U64 => ffi::HJ_U64,
        VarType::F16 => ffi::HJ_F16,
        VarType::F32 => ffi::HJ_F32,
        VarType::F64 => ffi::HJ_F64,
        VarType::Vec { .. } => ffi::HJ_VEC,
        VarType::Array { .. } => ffi::HJ_ARRAY,
        VarType::Mat { .. } => ffi::HJ_MAT,
        VarType::Struct { .. } => ffi::HJ_STRUCT,
    }
}

// VarType tree -> types[] (children before parents), interned per IR by address: VarTypes are
// leaked singletons (vartype.rs:20-85), so pointer identity is type identity.
fn flatten_type(ty: &'static VarType, flat: &mut FlatIr, seen: &mut HashMap<*const VarType, u32>) -> u32 {
    if let Some(&i) = seen.get(&(ty as *const VarType)) {
        return i;
    }
    let mut d = ffi::hj_type_desc { kind: scalar_kind(ty), ..Default::default() };
    match ty {
        VarType::Vec { ty: elem, num } | VarType::Array { ty: elem, num } => {
            d.elem = flatten_type(elem, flat, seen);
            d.num = *num as u32;
        }
        VarType::Mat { ty: elem, rows, cols } => {
            d.elem = flatten_type(elem, flat, seen);
            d.cols = *cols as u32;
            d.rows = *rows as u32;
        }
        VarType::Struct { tys } => {
            let fields: Vec<u32> = tys.iter().map(|t| flatten_type(t, flat, seen)).collect();
            d.num = fields.len() as u32;
            d.first_field = flat.struct_fields.len() as u32;
            flat.struct_fields.extend(fields);
        }
        _ => {}
    }
    let i = flat.types.len() as u32;
    flat.types.push(d);
    seen.insert(ty as *const VarType, i);
    i
}

// KernelOp -> (tag, payload); discriminants follow the declaration order of op.rs
fn flatten_op(op: KernelOp) -> (u32, u32) {
    match op {
        KernelOp::Nop => (ffi::HJ_OP_NOP, 0),
        KernelOp::Scatter => (ffi::HJ_OP_SCATTER, 0),
        KernelOp::ScatterReduce(r) => (ffi::HJ_OP_SCATTER_REDUCE, r as u32),
        KernelOp::ScatterAtomic(r) => (ffi::HJ_OP_SCATTER_ATOMIC, r as u32),
        KernelOp::AtomicInc => (ffi::HJ_OP_ATOMIC_INC, 0),
        KernelOp::Gather => (ffi::HJ_OP_GATHER, 0),
        KernelOp::Index => (ffi::HJ_OP_INDEX, 0),
        KernelOp::Literal => (ffi::HJ_OP_LITERAL, 0),
        KernelOp::Extract(elem) => (ffi::HJ_OP_EXTRACT, elem),
        KernelOp::DynExtract => (ffi::HJ_OP_DYN_EXTRACT, 0),
        KernelOp::Construct => (ffi::HJ_OP_CONSTRUCT, 0),
        KernelOp::Select => (ffi::HJ_OP_SELECT, 0),
        KernelOp::LoopStart => (ffi::HJ_OP_LOOP_START, 0),
        KernelOp::LoopEnd => (ffi::HJ_OP_LOOP_END, 0),
        KernelOp::IfStart => (ffi::HJ_OP_IF_START, 0),
        KernelOp::IfEnd => (ffi::HJ_OP_IF_END, 0),
        KernelOp::TexLookup => (ffi::HJ_OP_TEX_LOOKUP, 0),   // rejected by the library (out of scope)
        KernelOp::TraceRay => (ffi::HJ_OP_TRACE_RAY, 0),     // rejected by the library (out of scope)
        KernelOp::Bop(b) => (ffi::HJ_OP_BOP, b as u32),
        KernelOp::Uop(u) => (ffi::HJ_OP_UOP, u as u32),
        KernelOp::FMA => (ffi::HJ_OP_FMA, 0),
        KernelOp::BufferRef => (ffi::HJ_OP_BUFFER_REF, 0),
        KernelOp::TextureRef { dim } => (ffi::HJ_OP_TEXTURE_REF, dim),
        KernelOp::AccelRef => (ffi::HJ_OP_ACCEL_REF, 0),
    }
}

fn flatten_ir(ir: &IR) -> FlatIr {
    let mut flat = FlatIr { n_buffers: ir.n_buffers as u32, ..Default::default() };
    let mut seen = HashMap::new();
    flat.deps = ir.deps.iter().map(|d| d.0 as u32).collect();
    for id in ir.var_ids() {
        let var = ir.var(id);
        let (op, arg) = flatten_op(var.op);
        let ty = flatten_type(var.ty, &mut flat, &mut seen);
        flat.vars.push(ffi::hj_ir_var {
            ty, op, arg,
            dep_start: var.deps.0 as u32, dep_end: var.deps.1 as u32,
            _pad: 0, data: var.data,
        });
    }
    flat
}

// ---- BackendDevice ------------------------------------------------------------------------------
impl BackendDevice for CudaDevice {
    type Buffer = CudaBuffer;
    type Texture = CudaTexture;
    type Accel = CudaAccel;

    fn create_buffer(&self, size: usize) -> backend::Result<Self::Buffer> {
        // no power-of-two rounding (vulkan/mod.rs:130-132 needs it for its over-reading scan and
        // compress kernels; the CUDA kernels mask their tails)
        let mut buf = std::ptr::null_mut();
        check(unsafe { ffi::hj_buffer_create(self.0, size, &mut buf) })?;
        Ok(CudaBuffer { buf, device: self.clone() })
    }

    fn create_buffer_from_slice(&self, slice: &[u8]) -> backend::Result<Self::Buffer> {
        let mut buf = std::ptr::null_mut();
        check(unsafe { ffi::hj_buffer_create_from_slice(self.0, slice.as_ptr() as *const c_void, slice.len(), &mut buf) })?;
        Ok(CudaBuffer { buf, device: self.clone() })
    }

    // (inherent method, next to the trait: `tr::array` over pinned memory.  The upload returns at once; a launch of
    // ONE fused kernel over the buffer and `to_host` of its result then follow it chunk by chunk — DESIGN.md 3.8.
    // `src` must stay alive and unchanged until a blocking call on a dependent result has returned, which the
    // `PinnedSlice` owner type below guarantees by keeping the allocation until it is dropped.)
    // pub fn create_buffer_from_pinned<T: Copy>(&self, src: &PinnedSlice<T>) -> backend::Result<CudaBuffer> {
    //     let mut buf = std::ptr::null_mut();
    //     check(unsafe { ffi::hj_buffer_create_from_host_async(self.0, src.as_ptr() as *const c_void,
    //                    src.len() * std::mem::size_of::<T>(), std::mem::size_of::<T>(), &mut buf) })?;
    //     Ok(CudaBuffer { buf, device: self.clone() })
    // }

    fn create_texture(&self, _desc: &backend::TextureDesc) -> backend::Result<Self::Texture> {
        todo!("textures are outside the CUDA backend's scope")
    }

    fn create_accel(&self, _desc: &AccelDesc) -> backend::Result<Self::Accel> {
        todo!("acceleration structures are outside the CUDA backend's scope")
    }

    /// `vulkan/mod.rs:151-383` restated over the C ABI: the pass list goes down as is, resource
    /// ids stay resource ids, the library interprets the passes in order on the device stream.
    fn execute_graph(&self, graph: &Graph, env: &Env) -> backend::Result<Report> {
        let passes = graph.passes();
        // 1. owned storage first (the hj_pass / hj_ir views below point into these)
        let flat_irs: Vec<Option<FlatIr>> = passes
            .iter()
            .map(|p| match &p.op {
                PassOp::Kernel { ir, .. } => Some(flatten_ir(ir)),
                _ => None,
            })
            .collect();
        let ir_views: Vec<Option<ffi::hj_ir>> = flat_irs.iter().map(|f| f.as_ref().map(|f| f.view())).collect();
        let resource_lists: Vec<Vec<u32>> = passes.iter().map(|p| p.resources.iter().map(|r| r.0 as u32).collect()).collect();

        // 2. passes
        let mut flat_passes = Vec::with_capacity(passes.len());
        for (i, pass) in passes.iter().enumerate() {
            let (kind, arg, size) = match &pass.op {
                PassOp::None => continue,
                PassOp::Kernel { size, .. } => (ffi::HJ_PASS_KERNEL, 0, *size as u64),
                PassOp::DeviceOp(DeviceOp::ReduceOp(op)) => (ffi::HJ_PASS_REDUCE, *op as u32, 0),
                PassOp::DeviceOp(DeviceOp::PrefixSum { inclusive }) => (ffi::HJ_PASS_PREFIX_SUM, *inclusive as u32, 0),
                PassOp::DeviceOp(DeviceOp::Compress) => (ffi::HJ_PASS_COMPRESS, 0, 0),
                PassOp::DeviceOp(other) => todo!("{other:?} is outside the CUDA backend's scope"),
            };
            flat_passes.push(ffi::hj_pass {
                kind, arg,
                resources: resource_lists[i].as_ptr(), n_resources: resource_lists[i].len() as u32,
                size_buffer: pass.size_buffer.map_or(-1, |r| r.0 as i32),
                ir: ir_views[i].as_ref().map_or(std::ptr::null(), |v| v as *const ffi::hj_ir),
                size,
            });
        }

        // 3. environment: one slot per resource id that any pass names
        let n_resources = passes
            .iter()
            .flat_map(|p| p.resources.iter().chain(p.size_buffer.iter()))
            .map(|r| r.0 + 1)
            .max()
            .unwrap_or(0);
        let mut env_ptrs = vec![std::ptr::null_mut::<ffi::hj_buffer>(); n_resources];
        let mut descs = vec![ffi::hj_buffer_desc::default(); n_resources];
        for id in 0..n_resources {
            if let Some(buffer) = env.buffer(ResourceId(id)) {
                let cuda: &CudaBuffer = buffer.cuda().expect("a Vulkan buffer reached the CUDA backend");
                env_ptrs[id] = cuda.buf;
                let desc = graph.buffer_desc(ResourceId(id));
                descs[id] = ffi::hj_buffer_desc { size: desc.size as u64, ty: scalar_kind(desc.ty), elem_bytes: desc.ty.size() as u32 };
            }
        }

        // 4. launch; per-pass timings come back as PassReport (backend/report.rs:2-19)
        let mut pass_reports = vec![ffi::hj_pass_report { name: [0; 64], start_us: 0.0, duration_us: 0.0 }; flat_passes.len()];
        let mut rep = ffi::hj_report {
            cpu_duration_us: 0.0, n_passes: 0,
            passes: pass_reports.as_mut_ptr(), passes_capacity: pass_reports.len() as u32,
        };
        let cpu_start = std::time::SystemTime::now();
        check(unsafe {
            ffi::hj_execute_graph(self.0, flat_passes.as_ptr(), flat_passes.len() as u32, env_ptrs.as_ptr(),
                                  descs.as_ptr(), n_resources as u32, &mut rep)
        })?;
        let passes = pass_reports[..rep.n_passes as usize]
            .iter()
            .map(|p| PassReport {
                name: unsafe { CStr::from_ptr(p.name.as_ptr()) }.to_string_lossy().into_owned(),
                start: Duration::from_secs_f64(p.start_us * 1e-6),
                duration: Duration::from_secs_f64(p.duration_us * 1e-6),
            })
            .collect();
        Ok(Report {
            exec: ExecReport { cpu_start: Some(cpu_start), cpu_duration: Duration::from_secs_f64(rep.cpu_duration_us * 1e-6), passes },
        })
    }
}

impl BackendBuffer for CudaBuffer {
    type Device = CudaDevice;

    /// Blocks until everything enqueued before it on the device stream is visible
    /// (vulkan/mod.rs:478-509 blocks on a fence; same observable behaviour).
    fn to_host<T: AsVarType>(&self, range: std::ops::Range<usize>) -> backend::Result<Vec<T>> {
        let sz = std::mem::size_of::<T>();
        let mut out = Vec::<T>::with_capacity(range.len());
        check(unsafe { ffi::hj_buffer_to_host(self.buf, range.start * sz, range.len() * sz, out.as_mut_ptr() as *mut c_void) })?;
        unsafe { out.set_len(range.len()) };
        Ok(out)
    }

    fn device(&self) -> &Self::Device {
        &self.device
    }
}

#[derive(Debug, Clone)]
pub struct CudaTexture;
impl backend::BackendTexture for CudaTexture {
    type Device = CudaDevice;
}

#[derive(Debug, Clone)]
pub struct CudaAccel;
impl backend::BackendAccel for CudaAccel {
    type Device = CudaDevice;
}
