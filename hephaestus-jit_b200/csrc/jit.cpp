// jit.cpp — NVRTC compilation of fused-kernel IR to sm_100a cubins, kernel caches, launch.
//
// Plays the role of VulkanDevice::compile_ir + Pipeline::create
// (hephaestus-jit/src/backend/vulkan/mod.rs:83-91, vulkan_core/pipeline.rs:30-46: an
// in-memory pipeline cache keyed by the hash of the definition) and of the Kernel arm of
// execute_graph (vulkan/mod.rs:194-257).  Additions: an on-disk cubin cache (the reference
// recompiles every process start) and a vectorised entry point (codegen.cpp).
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <unordered_set>

#include "hj_internal.h"
#include "ir.h"

struct hj_kernel {
    std::atomic<int> rc{1};
    uint64_t hash = 0;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t scalar = nullptr;
    cudaKernel_t vec = nullptr;
    uint32_t vec_width = 1, unroll = 1, threads = 256, n_buffers = 0;
    std::vector<uint8_t> slot_flags;        // codegen.cpp: 1 = streamed input, 2 = streamed output
    std::vector<uint32_t> slot_elem_bytes;
    std::vector<hj::SlotAccess> slot_access;  // which slots may legally be bound to one buffer
};

namespace hj {

// The per-device pipeline cache.  A miss compiles OUTSIDE the lock: the hash is parked in
// `in_flight`, other threads asking for the same IR wait on the condition variable, and every
// other lookup (hits, misses of other IRs) goes on meanwhile — concurrent execute_graph calls
// (Device is Send + Sync) never queue behind somebody else's ~300 ms NVRTC run.
struct KernelCache {
    std::mutex mu;
    std::condition_variable cv;
    std::unordered_map<uint64_t, hj_kernel*> kernels;
    std::unordered_set<uint64_t> in_flight;
    uint64_t n_compiled = 0, n_hits = 0, n_disk_hits = 0;
};

namespace {

std::string cache_dir() {
    if (const char* d = getenv("HJ_CACHE_DIR")) return *d ? std::string(d) : std::string();
    const char* home = getenv("HOME");
    return std::string(home && *home ? home : "/tmp") + "/.cache/hj_b200";
}

void mkdirs(const std::string& p) {
    std::string cur;
    for (size_t i = 0; i < p.size(); i++) {
        cur += p[i];
        if (p[i] == '/' || i + 1 == p.size()) mkdir(cur.c_str(), 0755);
    }
}

const char* kArch = "sm_100a";

hj_status nvrtc_compile(const CodegenResult& cg, std::vector<char>* cubin) {
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, cg.source.c_str(), "hj_fused.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return fail(HJ_ERR_NVRTC, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo"};
    if (cg.uses_f16) opts.push_back("--include-path=/usr/local/cuda/include");
    if (getenv("HJ_FAST_MATH")) opts.push_back("--use_fast_math");
    r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        if (getenv("HJ_LOG")) fprintf(stderr, "[hj] NVRTC source:\n%s\n", cg.source.c_str());
        return fail(HJ_ERR_NVRTC, "NVRTC compilation failed: %s\n%.700s", nvrtcGetErrorString(r), log.c_str());
    }
    size_t sz = 0;
    r = nvrtcGetCUBINSize(prog, &sz);
    if (r != NVRTC_SUCCESS || sz == 0) {
        nvrtcDestroyProgram(&prog);
        return fail(HJ_ERR_NVRTC, "nvrtcGetCUBINSize: %s", nvrtcGetErrorString(r));
    }
    cubin->resize(sz);
    r = nvrtcGetCUBIN(prog, cubin->data());
    nvrtcDestroyProgram(&prog);
    if (r != NVRTC_SUCCESS) return fail(HJ_ERR_NVRTC, "nvrtcGetCUBIN: %s", nvrtcGetErrorString(r));
    return HJ_OK;
}

// cubin for `ir`: on-disk cache keyed by (source hash, arch, NVRTC version), else NVRTC.
hj_status get_cubin(const hj_ir* ir, CodegenResult* cg, std::vector<char>* cubin, bool* from_disk) {
    std::string err;
    if (!codegen_cuda(ir, cg, &err)) return fail(HJ_ERR_INVALID, "IR rejected: %s", err.c_str());
    *from_disk = false;
    int maj = 0, min = 0;
    nvrtcVersion(&maj, &min);
    uint64_t key = hash_bytes(cg->source.data(), cg->source.size());
    key = hash_bytes(kArch, strlen(kArch), key);
    int ver[3] = {maj, min, getenv("HJ_FAST_MATH") ? 1 : 0};
    key = hash_bytes(ver, sizeof(ver), key);
    std::string dir = cache_dir();
    char name[64];
    snprintf(name, sizeof(name), "/%016llx.cubin", (unsigned long long)key);
    if (!dir.empty()) {
        std::ifstream f(dir + name, std::ios::binary);
        if (f) {
            cubin->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
            if (cubin->size() > 64) { *from_disk = true; return HJ_OK; }
        }
    }
    HJ_TRY(nvrtc_compile(*cg, cubin));
    if (!dir.empty()) {
        mkdirs(dir);
        std::string tmp = dir + name + "." + std::to_string((long)getpid()) + ".tmp";
        std::ofstream f(tmp, std::ios::binary);
        if (f) {
            f.write(cubin->data(), (std::streamsize)cubin->size());
            f.close();
            rename(tmp.c_str(), (dir + name).c_str());  // atomic publish
        }
    }
    return HJ_OK;
}

}  // namespace
}  // namespace hj

using namespace hj;

namespace hj {
// A kernel pass behind asynchronous uploads: when at least one buffer is still arriving chunk by chunk
// (AsyncProgress with a schedule) and the kernel touches every buffer through the bare Index only, the
// pass is launched once per chunk on the kernel side stream, each launch behind the upload events of
// its chunk, and the buffers it writes inherit the schedule with one event per chunk — hj_buffer_to_host
// then drains them chunk by chunk.  `*done` = false: not applicable, the caller launches normally
// (hj_kernel_launch settles the buffers first).
hj_status kernel_launch_streamed(hj_device* dev, hj_kernel* k, size_t size, hj_buffer* const* buffers, uint32_t n_buffers,
                                 bool* done) {
    *done = false;
    static const bool debug = getenv("HJ_DEBUG_STREAMED") != nullptr;
    auto why = [&](int reason) {
        if (debug) fprintf(stderr, "[hj] kernel pass not streamed: reason %d\n", reason);
        return HJ_OK;
    };
    static const bool off = getenv("HJ_NO_STREAMED_LAUNCH") != nullptr;
    if (off || !k->vec || n_buffers != k->n_buffers || size == 0 || size > 0xffffffffull) return why(1);
    if (dev->async_live.load(std::memory_order_acquire) == 0) return why(2);
    DeviceGuard g(dev);
    std::shared_ptr<AsyncProgress> ref;
    for (uint32_t i = 0; i < n_buffers; i++) {
        hj_buffer* b = buffers[i];
        if (!b || b->dev != dev) return why(3);
        if (k->slot_flags[i] != 1 && k->slot_flags[i] != 2) return why(4);
        if (((uintptr_t)b->ptr & 15u) != 0 || b->bytes < size * k->slot_elem_bytes[i]) return why(5);
        for (uint32_t j = 0; j < i; j++)
            if (buffers[j] == b || buffers[j]->ptr == b->ptr) return why(6);  // aliased slots: the ordinary launch checks them
        if (!b->progress) continue;
        if (k->slot_flags[i] == 2) return why(7);  // a destination that is itself still being written
        if (b->progress->first.empty()) return why(8);  // being read elsewhere: settle and launch normally
        if (b->progress_elem_bytes != k->slot_elem_bytes[i]) return why(9);
        if (!ref) ref = b->progress;
        else if (ref->first != b->progress->first || ref->count != b->progress->count) return why(10);
    }
    if (!ref || ref->first.back() + ref->count.back() != size) return why(11);
    HJ_TRY(ensure_side_streams(dev));
    cudaStream_t sk = dev->side_kernel;
    // outputs were allocated on the device stream; whatever was enqueued there comes first
    cudaEvent_t start = nullptr;
    HJ_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(start, dev->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(sk, start, 0);
    cudaEventDestroy(start);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(HJ_ERR_CUDA, "streamed launch: %s", cudaGetErrorString(e)); }
    // every event this pass needs exists before its first launch: an error below never leaves work in flight
    // that nothing waits for
    auto out = std::make_shared<AsyncProgress>();
    auto fence = std::make_shared<AsyncProgress>();
    out->first = ref->first;
    out->count = ref->count;
    for (size_t c = 0; c <= ref->first.size(); c++) {
        cudaEvent_t ev = nullptr;
        HJ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        (c < ref->first.size() ? out : fence)->done.push_back(ev);
    }
    const size_t per_block = (size_t)k->threads * k->vec_width * k->unroll;
    std::vector<void*> ptrs(n_buffers);
    for (size_t c = 0; c < ref->first.size(); c++) {
        const size_t first = ref->first[c], count = ref->count[c];
        for (uint32_t i = 0; i < n_buffers; i++) {
            if (buffers[i]->progress) cudaStreamWaitEvent(sk, buffers[i]->progress->done[c], 0);
            ptrs[i] = (char*)buffers[i]->ptr + first * k->slot_elem_bytes[i];
        }
        const uint32_t* size_ptr = nullptr;
        uint32_t size_static = (uint32_t)count, index_base = (uint32_t)first;
        std::vector<void*> args = {(void*)&size_ptr, (void*)&size_static, (void*)&index_base};
        for (uint32_t i = 0; i < n_buffers; i++) args.push_back((void*)&ptrs[i]);
        const unsigned grid = (unsigned)((count + per_block - 1) / per_block);
        e = cudaLaunchKernel((const void*)k->vec, dim3(grid), dim3(k->threads), args.data(), 0, sk);
        if (e == cudaSuccess) e = cudaEventRecord(out->done[c], sk);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaStreamSynchronize(sk);  // nothing of this pass may outlive the error return
            return fail(HJ_ERR_CUDA, "streamed launch of fused kernel failed: %s", cudaGetErrorString(e));
        }
        dev->launches.fetch_add(1, std::memory_order_relaxed);
    }
    // the buffers this pass wrote arrive chunk by chunk; the ones it read are busy until its last chunk
    e = cudaEventRecord(fence->done[0], sk);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(sk);
        return fail(HJ_ERR_CUDA, "streamed launch: %s", cudaGetErrorString(e));
    }
    for (uint32_t i = 0; i < n_buffers; i++) {
        if (k->slot_flags[i] == 2) attach_progress(buffers[i], out, k->slot_elem_bytes[i]);
        else attach_progress(buffers[i], fence, k->slot_elem_bytes[i]);
    }
    *done = true;
    return HJ_OK;
}
}  // namespace hj

extern "C" {

hj_status hj_ir_codegen(const hj_ir* ir, char** out_source) {
    HJ_REQUIRE(ir && out_source, "null argument");
    CodegenResult cg;
    std::string err;
    if (!codegen_cuda(ir, &cg, &err)) return fail(HJ_ERR_INVALID, "IR rejected: %s", err.c_str());
    char* s = (char*)malloc(cg.source.size() + 1);
    if (!s) return fail(HJ_ERR_OOM, "out of host memory");
    memcpy(s, cg.source.c_str(), cg.source.size() + 1);
    *out_source = s;
    return HJ_OK;
}

void hj_free_string(char* s) { free(s); }

hj_status hj_ir_compile_cubin(const hj_ir* ir, void** out_cubin, size_t* out_size) {
    HJ_REQUIRE(ir && out_cubin && out_size, "null argument");
    CodegenResult cg;
    std::vector<char> cubin;
    bool disk = false;
    HJ_TRY(get_cubin(ir, &cg, &cubin, &disk));
    void* p = malloc(cubin.size());
    if (!p) return fail(HJ_ERR_OOM, "out of host memory");
    memcpy(p, cubin.data(), cubin.size());
    *out_cubin = p;
    *out_size = cubin.size();
    return HJ_OK;
}

hj_status hj_kernel_get(hj_device* dev, const hj_ir* ir, hj_kernel** out) {
    HJ_REQUIRE(dev && ir && out, "null argument");
    uint64_t h = hj_ir_hash(ir);
    {
        std::lock_guard<std::recursive_mutex> g(dev->mu);
        if (!dev->kcache) dev->kcache = new KernelCache();
    }
    KernelCache* kc = dev->kcache;
    {
        std::unique_lock<std::mutex> g(kc->mu);
        while (true) {
            auto it = kc->kernels.find(h);
            if (it != kc->kernels.end()) {
                kc->n_hits++;
                it->second->rc.fetch_add(1);
                *out = it->second;
                return HJ_OK;
            }
            if (!kc->in_flight.count(h)) break;
            kc->cv.wait(g);  // another thread is compiling this very IR
        }
        kc->in_flight.insert(h);
    }
    // ---- compile and load without the cache lock
    CodegenResult cg;
    std::vector<char> cubin;
    bool disk = false;
    hj_kernel* k = nullptr;
    hj_status st = get_cubin(ir, &cg, &cubin, &disk);
    if (st == HJ_OK) {
        cudaSetDevice(dev->ordinal);
        k = new hj_kernel();
        k->hash = h;
        cudaError_t e = cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->scalar, k->lib, "hj_kernel_scalar");
        if (e == cudaSuccess && cg.has_vec_entry) e = cudaLibraryGetKernel(&k->vec, k->lib, "hj_kernel_vec");
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (k->lib) cudaLibraryUnload(k->lib);
            delete k;
            k = nullptr;
            st = fail(HJ_ERR_CUDA, "loading the compiled kernel failed: %s", cudaGetErrorString(e));
        }
    }
    if (k) {
        k->vec_width = cg.vec;
        k->unroll = cg.unroll;
        k->threads = cg.threads;
        k->n_buffers = ir->n_buffers;
        k->slot_flags = cg.slot_flags;
        k->slot_elem_bytes = cg.slot_elem_bytes;
        analyse_slot_access(ir, &k->slot_access);
        k->rc.store(2);  // cache + caller
    }
    {
        std::lock_guard<std::mutex> g(kc->mu);
        kc->in_flight.erase(h);
        if (k) {
            if (disk) kc->n_disk_hits++;
            else kc->n_compiled++;
            kc->kernels[h] = k;
        }
    }
    kc->cv.notify_all();  // waiters of a failed compile retry it themselves and report their own error
    if (st != HJ_OK) return st;
    *out = k;
    return HJ_OK;
}

hj_status hj_kernel_release(hj_kernel* k) {
    HJ_REQUIRE(k, "null kernel");
    k->rc.fetch_sub(1);  // the per-device cache keeps compiled kernels for the process lifetime
    return HJ_OK;
}

hj_status hj_device_kernel_cache_stats(hj_device* dev, uint64_t* n_compiled, uint64_t* n_hits,
                                       uint64_t* n_disk_hits) {
    HJ_REQUIRE(dev, "null device");
    KernelCache* kc = dev->kcache;
    if (n_compiled) *n_compiled = kc ? kc->n_compiled : 0;
    if (n_hits) *n_hits = kc ? kc->n_hits : 0;
    if (n_disk_hits) *n_disk_hits = kc ? kc->n_disk_hits : 0;
    return HJ_OK;
}

hj_status hj_kernel_launch(hj_device* dev, hj_kernel* k, size_t size, hj_buffer* size_buf,
                           hj_buffer* const* buffers, uint32_t n_buffers, uint32_t index_base) {
    return hj::kernel_launch_shifted(dev, k, size, size_buf, buffers, n_buffers, index_base, nullptr);
}

}  // extern "C"

// `shift_bytes[i]` (optional): slot i is bound to buffers[i]->ptr - shift_bytes[i].  A segment kernel of
// the sharded pass interpreter addresses the rank's BLOCK of a sharded array with GLOBAL indices that all
// fall into that block: the base moves back by the block's start, every access lands inside the buffer.
hj_status hj::kernel_launch_shifted(hj_device* dev, hj_kernel* k, size_t size, hj_buffer* size_buf, hj_buffer* const* buffers,
                                    uint32_t n_buffers, uint32_t index_base, const uint64_t* shift_bytes) {
    HJ_REQUIRE(dev && k, "null argument");
    HJ_REQUIRE(n_buffers == k->n_buffers, "kernel expects %u buffers, got %u", k->n_buffers, n_buffers);
    HJ_REQUIRE(size <= 0xffffffffull, "kernel size does not fit the u32 index type (trace.rs:552-562)");
    HJ_REQUIRE(index_base != 0xffffffffu || (size_buf && size_buf->bytes >= 8),
               "index_base 0xffffffff reads the base from size_buf[1]: a size buffer of two u32 is needed");
    if (size == 0) return HJ_OK;
    std::vector<void*> ptrs(n_buffers);
    bool aligned = true;
    for (uint32_t i = 0; i < n_buffers; i++) {
        HJ_REQUIRE(buffers && buffers[i], "buffer %u is null", i);
        const uint64_t shift = shift_bytes ? shift_bytes[i] : 0;
        ptrs[i] = (void*)((uintptr_t)buffers[i]->ptr - (uintptr_t)shift);  // an address, not a pointer into the allocation
        // a shifted slot is only ever addressed through computed indices: no vector access touches it
        if (!shift && ((uintptr_t)ptrs[i] & 15u)) aligned = false;
    }
    // One buffer under two slots: every slot is declared __restrict__ and read-only slots are loaded
    // through the non-coherent path, so a duplicate is only legal when no thread can observe another
    // thread's store — both read-only, or an Index-addressed input whose element is read before the
    // same element of an Index-addressed output is written (the in-place reuse Graph::launch_with's
    // lifetime aliasing produces, graph.rs:237-296).
    for (uint32_t i = 0; i < n_buffers; i++)
        for (uint32_t j = i + 1; j < n_buffers; j++) {
            const char *a0 = (const char*)buffers[i]->ptr, *b0 = (const char*)buffers[j]->ptr;
            if (a0 + buffers[i]->bytes <= b0 || b0 + buffers[j]->bytes <= a0) continue;
            const SlotAccess &a = k->slot_access[i], &b = k->slot_access[j];
            if (!a.written && !b.written) continue;
            HJ_REQUIRE(slots_may_share(a, b) || slots_may_share(b, a),
                       "buffer slots %u and %u overlap in memory and the kernel does not access both through the bare Index "
                       "(read before write): the result would depend on thread order", i, j);
        }
    const uint32_t* size_ptr = size_buf ? (const uint32_t*)size_buf->ptr : nullptr;
    uint32_t size_static = (uint32_t)size;
    std::vector<void*> args;
    args.push_back((void*)&size_ptr);
    args.push_back((void*)&size_static);
    args.push_back((void*)&index_base);
    for (uint32_t i = 0; i < n_buffers; i++) args.push_back((void*)&ptrs[i]);

    DeviceGuard g(dev);
    for (uint32_t i = 0; i < n_buffers; i++) settle_locked(buffers[i]);
    const bool use_vec = k->vec && aligned && !getenv("HJ_JIT_SCALAR");
    const size_t per_block = use_vec ? (size_t)k->threads * k->vec_width * k->unroll : k->threads;
    const unsigned grid = (unsigned)((size + per_block - 1) / per_block);
    // programmatic stream serialization: the generated kernels wait (griddepcontrol.wait) before their
    // first global access, so their launch may overlap the tail of the kernel in front
    static const bool no_pdl = getenv("HJ_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(k->threads);
    cfg.stream = dev->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)(use_vec ? k->vec : k->scalar), args.data());
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HJ_ERR_CUDA, "launch of fused kernel failed: %s", cudaGetErrorString(e));
    }
    dev->launches.fetch_add(1, std::memory_order_relaxed);
    return HJ_OK;
}

extern "C" {

// Out-of-core elementwise map: every buffer of the kernel lives in HOST memory (pinned for full
// speed) and is addressed by the bare Index only.  The array is cut into chunks that flow
// through three streams — upload, kernel, download — with device-side chunk buffers reused
// round-robin, so the PCIe link runs in both directions while the kernel works on the chunk in
// between.  KernelOp::Index keeps its global value (index_base = chunk offset).  This is the
// pipelined form of `tr::array(..)` -> launch -> `to_vec(..)` (trace.rs:647-663, 1404-1438),
// which the reference runs as three blocking steps.
hj_status hj_kernel_map_host(hj_device* dev, hj_kernel* k, size_t n, void* const* host_arrays, uint32_t n_arrays,
                             size_t chunk_elems) {
    HJ_REQUIRE(dev && k && host_arrays, "null argument");
    HJ_REQUIRE(n_arrays == k->n_buffers, "kernel expects %u buffers, got %u", k->n_buffers, n_arrays);
    HJ_REQUIRE(n <= 0xffffffffull, "size does not fit the u32 index type");
    HJ_REQUIRE(k->vec, "kernel has no streamed entry point");
    for (uint32_t i = 0; i < n_arrays; i++) {
        HJ_REQUIRE(host_arrays[i], "host array %u is null", i);
        HJ_REQUIRE(k->slot_flags[i] == 1 || k->slot_flags[i] == 2,
                   "buffer %u is not accessed through the bare Index only: the kernel cannot be streamed", i);
    }
    if (n == 0) return HJ_OK;
    const size_t per_block = (size_t)k->threads * k->vec_width * k->unroll;
    if (chunk_elems == 0) chunk_elems = (size_t)1 << 24;
    chunk_elems = (chunk_elems + per_block - 1) / per_block * per_block;
    constexpr int DEPTH = 3;
    // Chunk schedule: the pipeline's fill (first upload) and drain (last download) cannot overlap
    // with traffic in the other direction, so the first and last chunks are small (1/8, 1/4, 1/2 of
    // a full chunk on the way in, mirrored on the way out) and only the middle runs at full size.
    std::vector<std::pair<size_t, size_t>> chunks;  // (first element, count)
    size_t max_chunk = 0;
    {
        std::vector<size_t> first, count;
        chunk_schedule(n, chunk_elems, per_block, &first, &count);
        for (size_t c = 0; c < first.size(); c++) {
            chunks.emplace_back(first[c], count[c]);
            max_chunk = std::max(max_chunk, count[c]);
        }
    }
    // the device lock is only needed for the stream-ordered allocations on the device stream: the chunks
    // run on this call's own three streams, so several host threads may stream at once
    std::unique_ptr<DeviceGuard> g(new DeviceGuard(dev));
    cudaStream_t s_up, s_k, s_down;
    HJ_CUDA(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    HJ_CUDA(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
    HJ_CUDA(cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking));
    cudaEvent_t up_done[DEPTH], k_done[DEPTH], down_done[DEPTH], start;
    for (int d = 0; d < DEPTH; d++) {
        HJ_CUDA(cudaEventCreateWithFlags(&up_done[d], cudaEventDisableTiming));
        HJ_CUDA(cudaEventCreateWithFlags(&k_done[d], cudaEventDisableTiming));
        HJ_CUDA(cudaEventCreateWithFlags(&down_done[d], cudaEventDisableTiming));
    }
    // order after everything already enqueued on the device stream
    HJ_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    HJ_CUDA(cudaEventRecord(start, dev->stream));
    HJ_CUDA(cudaStreamWaitEvent(s_up, start, 0));
    HJ_CUDA(cudaStreamWaitEvent(s_k, start, 0));
    HJ_CUDA(cudaStreamWaitEvent(s_down, start, 0));
    std::vector<char*> dbuf((size_t)DEPTH * n_arrays, nullptr);
    hj_status st = HJ_OK;
    for (size_t i = 0; i < dbuf.size() && st == HJ_OK; i++) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, max_chunk * k->slot_elem_bytes[i % n_arrays], dev->stream);
        if (e != cudaSuccess) { cudaGetLastError(); st = fail(HJ_ERR_OOM, "chunk buffer allocation failed: %s", cudaGetErrorString(e)); }
        dbuf[i] = (char*)p;
    }
    if (st == HJ_OK) {
        HJ_CUDA(cudaEventRecord(start, dev->stream));  // allocations are stream-ordered
        HJ_CUDA(cudaStreamWaitEvent(s_up, start, 0));
        HJ_CUDA(cudaStreamWaitEvent(s_k, start, 0));
        HJ_CUDA(cudaStreamWaitEvent(s_down, start, 0));
    }
    g.reset();
    const size_t n_chunks = chunks.size();
    for (size_t c = 0; c < n_chunks && st == HJ_OK; c++) {
        const int d = (int)(c % DEPTH);
        const size_t first = chunks[c].first;
        const size_t count = chunks[c].second;
        char** bufs = &dbuf[(size_t)d * n_arrays];
        // the slot is free once the chunk that used it DEPTH steps ago has been downloaded
        if (c >= DEPTH) cudaStreamWaitEvent(s_up, down_done[d], 0);
        for (uint32_t i = 0; i < n_arrays; i++)
            if (k->slot_flags[i] == 1)
                cudaMemcpyAsync(bufs[i], (const char*)host_arrays[i] + first * k->slot_elem_bytes[i],
                                count * k->slot_elem_bytes[i], cudaMemcpyHostToDevice, s_up);
        cudaEventRecord(up_done[d], s_up);
        cudaStreamWaitEvent(s_k, up_done[d], 0);
        if (c >= DEPTH) cudaStreamWaitEvent(s_k, down_done[d], 0);
        std::vector<void*> ptrs(n_arrays);
        for (uint32_t i = 0; i < n_arrays; i++) ptrs[i] = bufs[i];
        const uint32_t* size_ptr = nullptr;
        uint32_t size_static = (uint32_t)count, index_base = (uint32_t)first;
        std::vector<void*> args = {(void*)&size_ptr, (void*)&size_static, (void*)&index_base};
        for (uint32_t i = 0; i < n_arrays; i++) args.push_back((void*)&ptrs[i]);
        const unsigned grid = (unsigned)((count + per_block - 1) / per_block);
        cudaError_t e = cudaLaunchKernel((const void*)k->vec, dim3(grid), dim3(k->threads), args.data(), 0, s_k);
        if (e != cudaSuccess) { cudaGetLastError(); st = fail(HJ_ERR_CUDA, "launch of fused kernel failed: %s", cudaGetErrorString(e)); break; }
        dev->launches.fetch_add(1, std::memory_order_relaxed);
        cudaEventRecord(k_done[d], s_k);
        cudaStreamWaitEvent(s_down, k_done[d], 0);
        for (uint32_t i = 0; i < n_arrays; i++)
            if (k->slot_flags[i] == 2)
                cudaMemcpyAsync((char*)host_arrays[i] + first * k->slot_elem_bytes[i], bufs[i],
                                count * k->slot_elem_bytes[i], cudaMemcpyDeviceToHost, s_down);
        cudaEventRecord(down_done[d], s_down);
    }
    // results are visible on return, like BackendBuffer::to_host
    cudaStreamSynchronize(s_up);
    cudaStreamSynchronize(s_k);
    cudaError_t fin = cudaStreamSynchronize(s_down);
    {
        DeviceGuard g2(dev);
        for (char* p : dbuf)
            if (p) cudaFreeAsync(p, dev->stream);
    }
    for (int d = 0; d < DEPTH; d++) { cudaEventDestroy(up_done[d]); cudaEventDestroy(k_done[d]); cudaEventDestroy(down_done[d]); }
    cudaEventDestroy(start);
    cudaStreamDestroy(s_up); cudaStreamDestroy(s_k); cudaStreamDestroy(s_down);
    if (st == HJ_OK && fin != cudaSuccess) { cudaGetLastError(); st = fail(HJ_ERR_CUDA, "streamed map failed: %s", cudaGetErrorString(fin)); }
    return st;
}

}  // extern "C"
