// runtime.cpp — device / buffer / pool runtime and the C-ABI entry points of the device ops.
//
// Host-side counterpart of the reference's backend glue
// (hephaestus-jit/src/backend/vulkan/mod.rs:94-148,427-509 and vulkan_core/{device,buffer,pool}.rs):
// process-global devices per ordinal, ref-counted buffers leased from a caching pool, staged
// host<->device copies.  CUDA-native choices: one non-blocking stream per device (stream order
// replaces the reference's render-graph barriers and its blocking fence per submit), the
// stream-ordered allocator `cudaMallocAsync` with an unbounded release threshold as the pool.
#include <algorithm>
#include <cstdlib>
#include <map>

#include <sched.h>

#include <cctype>
#include <string>

#include "hj_internal.h"

namespace hj {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
hj_status fail(hj_status code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (getenv("HJ_LOG")) fprintf(stderr, "[hj] error %d: %s\n", code, buf);
    return code;
}

size_t type_size(hj_type_kind ty) {
    switch (ty) {
    case HJ_BOOL: case HJ_I8: case HJ_U8: return 1;
    case HJ_I16: case HJ_U16: case HJ_F16: return 2;
    case HJ_I32: case HJ_U32: case HJ_F32: return 4;
    case HJ_I64: case HJ_U64: case HJ_F64: return 8;
    default: return 0;
    }
}
const char* type_name(hj_type_kind ty) {
    static const char* names[] = {"Void", "Bool", "I8", "U8", "I16", "U16", "I32", "U32", "I64",
                                  "U64", "F16", "F32", "F64", "Vec", "Array", "Mat", "Struct"};
    return (unsigned)ty < sizeof(names) / sizeof(names[0]) ? names[ty] : "?";
}
const char* reduce_op_name(hj_reduce_op op) {
    static const char* names[] = {"Max", "Min", "Sum", "Prod", "Or", "And", "Xor"};
    return (unsigned)op < 7 ? names[op] : "?";
}

hj_status ensure_reduce_scratch(hj_device* dev, size_t bytes) {
    if (dev->reduce_scratch_bytes >= bytes) return HJ_OK;
    if (dev->reduce_scratch) HJ_CUDA(cudaFree(dev->reduce_scratch));  // implicit sync: safe
    dev->reduce_scratch = nullptr;
    dev->lookback.generation++;
    HJ_CUDA(cudaMalloc(&dev->reduce_scratch, bytes));
    HJ_CUDA(cudaMemsetAsync(dev->reduce_scratch, 0, bytes, dev->stream));
    dev->reduce_scratch_bytes = bytes;
    return HJ_OK;
}

hj_status ensure_hist_scratch(hj_device* dev, size_t bytes) {
    if (dev->hist_scratch_bytes >= bytes) return HJ_OK;
    if (dev->hist_scratch) HJ_CUDA(cudaFree(dev->hist_scratch));  // implicit sync: safe
    dev->hist_scratch = nullptr;
    dev->hist_scratch_bytes = 0;
    dev->lookback.generation++;
    HJ_CUDA(cudaMalloc(&dev->hist_scratch, bytes));
    dev->hist_scratch_bytes = bytes;
    return HJ_OK;
}

hj_status ensure_lookback_scratch(hj_device* dev, size_t n_tiles) {
    LookbackScratch& lb = dev->lookback;
    if (lb.capacity_tiles >= n_tiles) return HJ_OK;
    size_t cap = lb.capacity_tiles ? lb.capacity_tiles : 4096;
    while (cap < n_tiles) cap *= 2;
    if (lb.base) HJ_CUDA(cudaFree(lb.base));
    lb.base = nullptr;
    lb.capacity_tiles = 0;
    size_t bytes = 64 + cap * 24;
    lb.generation++;
    lb.epoch = 0;
    HJ_CUDA(cudaMalloc(&lb.base, bytes));
    HJ_CUDA(cudaMemsetAsync(lb.base, 0, bytes, dev->stream));
    lb.capacity_tiles = cap;
    lb.bytes = bytes;
    // epochs already handed out stay unique: the new buffer is all-zero (= epoch 0 = invalid)
    return HJ_OK;
}

// The epoch itself lives on the device (lookback.cuh); the host only counts how many epochs the
// launches it has enqueued (graph replays included) will consume, and clears the scratch — epoch
// counter included — before the 30-bit epoch could wrap onto a stale status word.
hj_status count_epoch(hj_device* dev, uint32_t n) {
    LookbackScratch& lb = dev->lookback;
    if (lb.base && (uint64_t)lb.epoch + n >= (1u << 30) - 1) {
        HJ_CUDA(cudaMemsetAsync(lb.base, 0, lb.bytes, dev->stream));
        lb.epoch = 0;
    }
    lb.epoch += n;
    return HJ_OK;
}

hj_status ensure_dynamic_smem(hj_device* dev, const void* kernel, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_pair(dev->ordinal, kernel);
    auto it = done.find(key);
    if (it != done.end() && it->second >= smem) return HJ_OK;
    HJ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done[key] = smem;
    return HJ_OK;
}

static std::mutex g_devices_mu;
static std::map<int, hj_device*> g_devices;

}  // namespace hj

using namespace hj;

extern "C" {

const char* hj_last_error(void) { return g_last_error.c_str(); }
uint32_t hj_abi_version(void) { return 1; }

int32_t hj_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

hj_status hj_device_create(int32_t ordinal, hj_device** out) {
    HJ_REQUIRE(out, "hj_device_create: out is null");
    std::lock_guard<std::mutex> g(g_devices_mu);
    auto it = g_devices.find(ordinal);
    if (it != g_devices.end()) {  // process-global singleton per ordinal (vulkan/mod.rs:94-113)
        it->second->rc.fetch_add(1);
        *out = it->second;
        return HJ_OK;
    }
    int n = hj_device_count();
    if (n == 0) return fail(HJ_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
    HJ_REQUIRE(ordinal >= 0 && ordinal < n, "device ordinal %d out of range (have %d)", ordinal, n);
    HJ_CUDA(cudaSetDevice(ordinal));
    auto dev = new hj_device();
    dev->ordinal = ordinal;
    cudaDeviceProp prop;
    HJ_CUDA(cudaGetDeviceProperties(&prop, ordinal));
    dev->sm_count = prop.multiProcessorCount;
    dev->cc_major = prop.major;
    dev->cc_minor = prop.minor;
    dev->total_mem = prop.totalGlobalMem;
    dev->l2_bytes = (size_t)prop.l2CacheSize;
    dev->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (const char* e = getenv("HJ_L2_FETCH")) {  // experiment: L2 fetch granularity in bytes (32 / 64 / 128)
        const int gran = atoi(e);
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        cudaGetLastError();
    }
    HJ_CUDA(cudaStreamCreateWithFlags(&dev->own_stream, cudaStreamNonBlocking));
    dev->stream = dev->own_stream;
    HJ_CUDA(cudaDeviceGetDefaultMemPool(&dev->pool, ordinal));
    uint64_t threshold = UINT64_MAX;  // keep freed blocks cached: this IS the resource pool
    HJ_CUDA(cudaMemPoolSetAttribute(dev->pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    dev->rc.store(2);  // one for the registry, one for the caller
    g_devices[ordinal] = dev;
    *out = dev;
    return HJ_OK;
}

hj_status hj_device_retain(hj_device* dev) {
    HJ_REQUIRE(dev, "null device");
    dev->rc.fetch_add(1);
    return HJ_OK;
}
hj_status hj_device_release(hj_device* dev) {
    HJ_REQUIRE(dev, "null device");
    dev->rc.fetch_sub(1);  // the registry keeps devices alive for the process lifetime
    return HJ_OK;
}
hj_status hj_device_sync(hj_device* dev) {
    HJ_REQUIRE(dev, "null device");
    DeviceGuard g(dev);
    HJ_CUDA(cudaStreamSynchronize(dev->stream));
    for (cudaStream_t s : {dev->side_up, dev->side_kernel, dev->side_down})
        if (s) HJ_CUDA(cudaStreamSynchronize(s));
    return HJ_OK;
}
hj_status hj_device_stream(hj_device* dev, void** out_stream) {
    HJ_REQUIRE(dev && out_stream, "null argument");
    *out_stream = (void*)dev->stream;
    return HJ_OK;
}
hj_status hj_device_set_stream(hj_device* dev, void* stream) {
    HJ_REQUIRE(dev, "null device");
    DeviceGuard g(dev);
    HJ_CUDA(cudaStreamSynchronize(dev->stream));  // scratch buffers are shared across streams
    dev->stream = stream ? (cudaStream_t)stream : dev->own_stream;
    return HJ_OK;
}
hj_status hj_device_info(hj_device* dev, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                         uint64_t* total_mem, uint64_t* l2_bytes) {
    HJ_REQUIRE(dev, "null device");
    if (sm_count) *sm_count = dev->sm_count;
    if (cc_major) *cc_major = dev->cc_major;
    if (cc_minor) *cc_minor = dev->cc_minor;
    if (total_mem) *total_mem = dev->total_mem;
    if (l2_bytes) *l2_bytes = dev->l2_bytes;
    return HJ_OK;
}
hj_status hj_device_pool_stats(hj_device* dev, uint64_t* bytes_live, uint64_t* bytes_cached,
                               uint64_t* n_alloc, uint64_t* n_free) {
    HJ_REQUIRE(dev, "null device");
    DeviceGuard g(dev);
    uint64_t used = 0, reserved = 0;
    HJ_CUDA(cudaMemPoolGetAttribute(dev->pool, cudaMemPoolAttrUsedMemCurrent, &used));
    HJ_CUDA(cudaMemPoolGetAttribute(dev->pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
    if (bytes_live) *bytes_live = used;
    if (bytes_cached) *bytes_cached = reserved - used;
    if (n_alloc) *n_alloc = dev->n_alloc.load();
    if (n_free) *n_free = dev->n_free.load();
    return HJ_OK;
}
hj_status hj_device_pool_trim(hj_device* dev) {
    HJ_REQUIRE(dev, "null device");
    DeviceGuard g(dev);
    HJ_CUDA(cudaStreamSynchronize(dev->stream));
    HJ_CUDA(cudaMemPoolTrimTo(dev->pool, 0));
    return HJ_OK;
}
hj_status hj_device_launch_count(hj_device* dev, uint64_t* out) {
    HJ_REQUIRE(dev && out, "null argument");
    *out = dev->launches.load();
    return HJ_OK;
}

// ---- chunk-wise asynchronous buffers ----------------------------------------------------------
}  // extern "C"
namespace hj {
hj_status ensure_side_streams(hj_device* dev) {
    if (dev->side_down) return HJ_OK;
    HJ_CUDA(cudaStreamCreateWithFlags(&dev->side_up, cudaStreamNonBlocking));
    HJ_CUDA(cudaStreamCreateWithFlags(&dev->side_kernel, cudaStreamNonBlocking));
    HJ_CUDA(cudaStreamCreateWithFlags(&dev->side_down, cudaStreamNonBlocking));
    return HJ_OK;
}
void attach_progress(hj_buffer* b, std::shared_ptr<AsyncProgress> p, uint32_t elem_bytes) {
    if (!b->progress) b->dev->async_live.fetch_add(1, std::memory_order_release);
    b->progress = std::move(p);
    b->progress_elem_bytes = elem_bytes;
}
void settle_locked(hj_buffer* b) {
    if (!b || !b->progress) return;
    if (!b->progress->done.empty()) cudaStreamWaitEvent(b->dev->stream, b->progress->done.back(), 0);
    b->progress.reset();
    b->dev->async_live.fetch_sub(1, std::memory_order_release);
}
void chunk_schedule(size_t n, size_t chunk_elems, size_t align_elems, std::vector<size_t>* first, std::vector<size_t>* count) {
    first->clear();
    count->clear();
    if (align_elems == 0) align_elems = 1;
    auto up = [&](size_t v) { return std::max(align_elems, (v + align_elems - 1) / align_elems * align_elems); };
    const size_t chunk = up(chunk_elems);
    size_t pos = 0;
    auto push = [&](size_t c) {
        if (c) {
            first->push_back(pos);
            count->push_back(c);
            pos += c;
        }
    };
    // every boundary is a multiple of `align_elems`; what is left of a ragged n joins the last chunk
    const size_t n_al = n / align_elems * align_elems;
    static const bool ramp = !getenv("HJ_MAP_NO_RAMP");
    const size_t r[3] = {up(chunk / 8), up(chunk / 4), up(chunk / 2)};
    const size_t ramp_total = r[0] + r[1] + r[2];
    if (ramp && n_al >= 4 * chunk && n_al >= 2 * ramp_total + chunk) {
        for (int i = 0; i < 3; i++) push(r[i]);
        for (size_t mid = n_al - 2 * ramp_total; mid;) {
            const size_t c = std::min(chunk, mid);
            push(c);
            mid -= c;
        }
        for (int i = 2; i >= 0; i--) push(r[i]);
    } else {
        for (size_t left = n_al; left;) {
            const size_t c = std::min(chunk, left);
            push(c);
            left -= c;
        }
    }
    if (n > n_al) {
        if (count->empty()) push(n - n_al);
        else count->back() += n - n_al;
    }
}
}  // namespace hj
extern "C" {

// ---- buffers --------------------------------------------------------------------------------

hj_status hj_buffer_create(hj_device* dev, size_t bytes, hj_buffer** out) {
    HJ_REQUIRE(dev && out, "null argument");
    DeviceGuard g(dev);
    void* p = nullptr;
    // zero-sized buffers are legal handles (the reference rounds 0 up to 1 byte: round_pow2(0))
    HJ_CUDA(cudaMallocAsync(&p, bytes ? bytes : 1, dev->stream));
    dev->n_alloc.fetch_add(1);
    auto b = new hj_buffer();
    b->dev = dev;
    b->ptr = p;
    b->bytes = bytes;
    b->owned = true;
    dev->rc.fetch_add(1);
    *out = b;
    return HJ_OK;
}

hj_status hj_buffer_create_from_slice(hj_device* dev, const void* data, size_t bytes, hj_buffer** out) {
    HJ_REQUIRE(dev && out && (data || bytes == 0), "null argument");
    HJ_TRY(hj_buffer_create(dev, bytes, out));
    if (bytes) {
        DeviceGuard g(dev);
        // pageable source: the copy is staged by the driver before the call returns, so the
        // caller may free `data` immediately (same contract as the reference's staging copy)
        cudaError_t e = cudaMemcpyAsync((*out)->ptr, data, bytes, cudaMemcpyHostToDevice, dev->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(dev->stream);
        if (e != cudaSuccess) {
            g.lock.unlock();
            hj_buffer_release(*out);
            *out = nullptr;
            return fail(HJ_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
        }
    }
    return HJ_OK;
}

// Asynchronous chunk-wise upload (`tr::array -> launch -> to_vec` without three blocking steps).
// Returns at once; the copy runs on the upload side stream in the chunks of chunk_schedule(), one event
// per chunk.  A kernel pass over the bare Index that reads such buffers is launched chunk by chunk behind
// those events (jit.cpp: kernel_launch_streamed) and hj_buffer_to_host of its outputs drains chunk by
// chunk too, so upload, kernels and download of one array -> launch -> to_vec sequence overlap like in
// hj_kernel_map_host.  Contract: `src` stays valid and unchanged until the buffer has been consumed by
// something that blocks (hj_buffer_to_host of a dependent result, hj_device_sync); pinned memory
// (hj_host_alloc) is needed for the copy to be asynchronous at all.
static size_t default_async_chunk() {
    static const size_t chunk = getenv("HJ_ASYNC_CHUNK_ELEMS") ? (size_t)atoll(getenv("HJ_ASYNC_CHUNK_ELEMS")) : ((size_t)1 << 24);
    return chunk ? chunk : ((size_t)1 << 24);
}
hj_status hj_async_chunk_schedule(uint64_t n, uint64_t chunk_elems, uint64_t* first, uint64_t* count, uint32_t capacity,
                                  uint32_t* n_chunks) {
    HJ_REQUIRE(n_chunks && ((first && count) || capacity == 0), "hj_async_chunk_schedule: null argument");
    std::vector<size_t> f, c;
    chunk_schedule((size_t)n, chunk_elems ? (size_t)chunk_elems : default_async_chunk(), 4096, &f, &c);
    for (size_t i = 0; i < f.size() && i < capacity; i++) {
        first[i] = f[i];
        count[i] = c[i];
    }
    *n_chunks = (uint32_t)f.size();
    return HJ_OK;
}
hj_status hj_buffer_create_from_host_async(hj_device* dev, const void* src, size_t bytes, size_t elem_bytes,
                                           hj_buffer** out) {
    HJ_REQUIRE(dev && out && (src || bytes == 0), "null argument");
    HJ_REQUIRE(elem_bytes > 0 && bytes % elem_bytes == 0, "size %zu is not a multiple of the element size %zu", bytes, elem_bytes);
    HJ_TRY(hj_buffer_create(dev, bytes, out));
    if (!bytes) return HJ_OK;
    DeviceGuard g(dev);
    hj_status st = ensure_side_streams(dev);
    auto pr = std::make_shared<AsyncProgress>();
    if (st == HJ_OK) {
        chunk_schedule(bytes / elem_bytes, default_async_chunk(), 4096, &pr->first, &pr->count);
        cudaEvent_t alloc_done = nullptr;  // the allocation is ordered on the device stream
        cudaError_t e = cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(alloc_done, dev->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(dev->side_up, alloc_done, 0);
        for (size_t c = 0; c < pr->first.size() && e == cudaSuccess; c++) {
            cudaEvent_t ev = nullptr;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) break;
            pr->done.push_back(ev);
            e = cudaMemcpyAsync((char*)(*out)->ptr + pr->first[c] * elem_bytes, (const char*)src + pr->first[c] * elem_bytes,
                                pr->count[c] * elem_bytes, cudaMemcpyHostToDevice, dev->side_up);
            if (e == cudaSuccess) e = cudaEventRecord(ev, dev->side_up);
        }
        if (alloc_done) cudaEventDestroy(alloc_done);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaStreamSynchronize(dev->side_up);
            st = fail(HJ_ERR_CUDA, "asynchronous upload failed: %s", cudaGetErrorString(e));
        }
    }
    if (st != HJ_OK) {
        g.lock.unlock();
        hj_buffer_release(*out);
        *out = nullptr;
        return st;
    }
    attach_progress(*out, std::move(pr), (uint32_t)elem_bytes);
    return HJ_OK;
}

hj_status hj_buffer_wrap(hj_device* dev, void* device_ptr, size_t bytes, hj_buffer** out) {
    HJ_REQUIRE(dev && out && device_ptr, "null argument");
    auto b = new hj_buffer();
    b->dev = dev;
    b->ptr = device_ptr;
    b->bytes = bytes;
    b->owned = false;
    dev->rc.fetch_add(1);
    *out = b;
    return HJ_OK;
}

hj_status hj_buffer_retain(hj_buffer* buf) {
    HJ_REQUIRE(buf, "null buffer");
    buf->rc.fetch_add(1);
    return HJ_OK;
}

hj_status hj_buffer_release(hj_buffer* buf) {
    HJ_REQUIRE(buf, "null buffer");
    if (buf->rc.fetch_sub(1) != 1) return HJ_OK;
    hj_device* dev = buf->dev;
    settle(buf);  // a side stream may still be filling or reading it: the free below is ordered behind that
    if (buf->owned && buf->ptr) {
        DeviceGuard g(dev);
        // stream-ordered free: memory returns to the pool after all enqueued work that may
        // still use it; contents are NOT cleared (pool.rs:36-41)
        cudaFreeAsync(buf->ptr, dev->stream);
        dev->n_free.fetch_add(1);
    }
    dev->rc.fetch_sub(1);
    delete buf;
    return HJ_OK;
}

hj_status hj_buffer_to_host(hj_buffer* buf, size_t offset_bytes, size_t nbytes, void* dst) {
    HJ_REQUIRE(buf && (dst || nbytes == 0), "null argument");
    HJ_REQUIRE(offset_bytes + nbytes <= buf->bytes, "to_host range [%zu, %zu) exceeds buffer of %zu bytes",
               offset_bytes, offset_bytes + nbytes, buf->bytes);
    if (!nbytes) return HJ_OK;
    hj_device* dev = buf->dev;
    std::unique_lock<std::recursive_mutex> lock(dev->mu);
    cudaSetDevice(dev->ordinal);
    if (buf->progress && !buf->progress->first.empty() && dev->side_down) {
        // The buffer is still being produced chunk by chunk on a side stream: every chunk goes down as soon as
        // its event fires, on the download stream — the copy overlaps the uploads and kernels of later chunks.
        std::shared_ptr<AsyncProgress> pr = buf->progress;
        const size_t es = buf->progress_elem_bytes;
        for (size_t c = 0; c < pr->first.size(); c++) {
            const size_t c0 = pr->first[c] * es, c1 = c0 + pr->count[c] * es;
            const size_t lo = std::max(c0, offset_bytes), hi = std::min(c1, offset_bytes + nbytes);
            if (lo >= hi) continue;
            HJ_CUDA(cudaStreamWaitEvent(dev->side_down, pr->done[c], 0));
            HJ_CUDA(cudaMemcpyAsync((char*)dst + (lo - offset_bytes), (const char*)buf->ptr + lo, hi - lo,
                                    cudaMemcpyDeviceToHost, dev->side_down));
        }
        cudaStream_t down = dev->side_down;
        lock.unlock();  // other threads may enqueue while this one waits for its data
        HJ_CUDA(cudaStreamSynchronize(down));
        return HJ_OK;
    }
    settle_locked(buf);
    HJ_CUDA(cudaMemcpyAsync(dst, (const char*)buf->ptr + offset_bytes, nbytes, cudaMemcpyDeviceToHost,
                            dev->stream));
    HJ_CUDA(cudaStreamSynchronize(dev->stream));
    return HJ_OK;
}

hj_status hj_buffer_upload(hj_buffer* buf, size_t offset_bytes, const void* src, size_t nbytes) {
    HJ_REQUIRE(buf && (src || nbytes == 0), "null argument");
    HJ_REQUIRE(offset_bytes + nbytes <= buf->bytes, "upload range exceeds buffer");
    if (!nbytes) return HJ_OK;
    DeviceGuard g(buf->dev);
    settle_locked(buf);
    HJ_CUDA(cudaMemcpyAsync((char*)buf->ptr + offset_bytes, src, nbytes, cudaMemcpyHostToDevice,
                            buf->dev->stream));
    return HJ_OK;
}

hj_status hj_buffer_fill_zero(hj_buffer* buf) {
    HJ_REQUIRE(buf, "null buffer");
    if (!buf->bytes) return HJ_OK;
    DeviceGuard g(buf->dev);
    settle_locked(buf);
    HJ_CUDA(cudaMemsetAsync(buf->ptr, 0, buf->bytes, buf->dev->stream));
    return HJ_OK;
}

hj_status hj_buffer_device_ptr(hj_buffer* buf, void** out) {
    HJ_REQUIRE(buf && out, "null argument");
    settle(buf);  // the pointer leaves the library: whatever uses it is ordered on the device stream
    *out = buf->ptr;
    return HJ_OK;
}
hj_status hj_buffer_size(hj_buffer* buf, size_t* out_bytes) {
    HJ_REQUIRE(buf && out_bytes, "null argument");
    *out_bytes = buf->bytes;
    return HJ_OK;
}
hj_status hj_buffer_device(hj_buffer* buf, hj_device** out) {
    HJ_REQUIRE(buf && out, "null argument");
    *out = buf->dev;
    return HJ_OK;
}

hj_status hj_host_alloc(size_t bytes, void** out) {
    HJ_REQUIRE(out, "null argument");
    HJ_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return HJ_OK;
}
// Pinned memory on the NUMA node the GPU hangs off: cudaHostAlloc places the pages where the calling
// thread runs (first touch), so the thread is moved onto the CPUs of the device's node for the
// duration of the allocation.  With 8 ranks streaming at once, buffers that all sit on node 0 send
// half of the PCIe traffic across the socket interconnect.  Best effort: without sysfs topology or
// with a cpuset that excludes those CPUs this is plain hj_host_alloc.
hj_status hj_host_alloc_near(hj_device* dev, size_t bytes, void** out) {
    HJ_REQUIRE(dev && out, "null argument");
    cpu_set_t old_set, want;
    bool moved = false;
    if (!getenv("HJ_NO_NUMA") && sched_getaffinity(0, sizeof(old_set), &old_set) == 0) {
        char bus[32] = {0};
        int node = -1;
        if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev->ordinal) == cudaSuccess) {
            for (char* c = bus; *c; c++) *c = (char)tolower(*c);
            std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
            if (FILE* f = fopen(path.c_str(), "r")) {
                if (fscanf(f, "%d", &node) != 1) node = -1;
                fclose(f);
            }
        } else {
            cudaGetLastError();
        }
        if (node >= 0) {
            std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
            CPU_ZERO(&want);
            int n_cpus = 0;
            if (FILE* f = fopen(path.c_str(), "r")) {
                int a, b;
                char sep;
                while (fscanf(f, "%d", &a) == 1) {  // "0-23,48-71"
                    b = a;
                    if (fscanf(f, "%c", &sep) == 1 && sep == '-') {
                        if (fscanf(f, "%d", &b) != 1) b = a;
                        if (fscanf(f, "%c", &sep) != 1) sep = 0;
                    }
                    for (int c = a; c <= b && c < CPU_SETSIZE; c++)
                        if (CPU_ISSET(c, &old_set)) { CPU_SET(c, &want); n_cpus++; }  // stay inside the cpuset
                    if (sep != ',') break;
                }
                fclose(f);
            }
            if (n_cpus > 0 && sched_setaffinity(0, sizeof(want), &want) == 0) moved = true;
        }
    }
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (moved) sched_setaffinity(0, sizeof(old_set), &old_set);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HJ_ERR_OOM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return HJ_OK;
}
hj_status hj_host_free(void* ptr) {
    if (ptr) HJ_CUDA(cudaFreeHost(ptr));
    return HJ_OK;
}

// ---- device ops -----------------------------------------------------------------------------

hj_status hj_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n, hj_buffer* src,
                    hj_buffer* dst) {
    HJ_REQUIRE(dev && src && dst, "hj_reduce: null argument");
    size_t es = type_size(ty);
    HJ_REQUIRE(es, "hj_reduce: %s is not a scalar type", type_name(ty));
    // the reference panics for n == 0 (and n == 1, defect D6; we define n == 1 as the identity)
    HJ_REQUIRE(n >= 1, "hj_reduce: empty input");
    HJ_REQUIRE(n * es <= src->bytes, "hj_reduce: src holds %zu bytes, need %zu", src->bytes, n * es);
    HJ_REQUIRE(es <= dst->bytes, "hj_reduce: dst too small");
    DeviceGuard g(dev);
    settle_all(dev, {src, dst});
    return launch_reduce(dev, op, ty, n, src->ptr, dst->ptr);
}

hj_status hj_prefix_sum(hj_device* dev, hj_type_kind ty, size_t n, int32_t inclusive, hj_buffer* src,
                        hj_buffer* dst, hj_buffer* seed) {
    HJ_REQUIRE(dev && src && dst, "hj_prefix_sum: null argument");
    size_t es = type_size(ty);
    HJ_REQUIRE(es, "hj_prefix_sum: %s is not a scalar type", type_name(ty));
    HJ_REQUIRE(n >= 1, "hj_prefix_sum: empty input");
    HJ_REQUIRE(n * es <= src->bytes && n * es <= dst->bytes, "hj_prefix_sum: buffer too small for %zu elements", n);
    HJ_REQUIRE(!seed || seed->bytes >= es, "hj_prefix_sum: seed too small");
    if (getenv("HJ_REF_COMPAT")) inclusive = 1;  // reference defect D10: always inclusive
    DeviceGuard g(dev);
    settle_all(dev, {src, dst, seed});
    return launch_prefix_sum(dev, ty, n, inclusive != 0, src->ptr, dst->ptr, seed ? seed->ptr : nullptr);
}

hj_status hj_compress(hj_device* dev, size_t n, hj_buffer* size_buf, hj_buffer* out_count,
                      hj_buffer* src_mask, hj_buffer* index_out, uint32_t index_base) {
    HJ_REQUIRE(dev && out_count && src_mask && index_out, "hj_compress: null argument");
    HJ_REQUIRE(n >= 1, "hj_compress: empty input");
    HJ_REQUIRE(n <= src_mask->bytes, "hj_compress: mask holds %zu bytes, need %zu", src_mask->bytes, n);
    HJ_REQUIRE(n * 4 <= index_out->bytes, "hj_compress: index_out too small");
    HJ_REQUIRE(out_count->bytes >= 4 && (!size_buf || size_buf->bytes >= 4), "hj_compress: count buffer too small");
    DeviceGuard g(dev);
    settle_all(dev, {size_buf, out_count, src_mask, index_out});
    return launch_compress(dev, n, size_buf ? (const uint32_t*)size_buf->ptr : nullptr,
                           (uint32_t*)out_count->ptr, (const uint8_t*)src_mask->ptr,
                           (uint32_t*)index_out->ptr, index_base);
}

hj_status hj_scatter_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n, hj_buffer* idx,
                            hj_buffer* src, uint64_t literal, hj_buffer* dst, size_t n_dst) {
    HJ_REQUIRE(dev && idx && dst, "hj_scatter_reduce: null argument");
    size_t es = type_size(ty);
    HJ_REQUIRE(es, "hj_scatter_reduce: not a scalar type");
    HJ_REQUIRE(n * 4 <= idx->bytes, "hj_scatter_reduce: idx too small");
    HJ_REQUIRE(!src || n * es <= src->bytes, "hj_scatter_reduce: src too small");
    HJ_REQUIRE(n_dst * es <= dst->bytes, "hj_scatter_reduce: dst too small");
    DeviceGuard g(dev);
    settle_all(dev, {idx, src, dst});
    return launch_scatter_reduce(dev, op, ty, n, (const uint32_t*)idx->ptr, src ? src->ptr : nullptr,
                                 literal, dst->ptr, n_dst);
}

hj_status hj_gather(hj_device* dev, size_t elem_bytes, size_t n, hj_buffer* src, hj_buffer* idx,
                    hj_buffer* dst) {
    HJ_REQUIRE(dev && src && idx && dst, "hj_gather: null argument");
    HJ_REQUIRE(n * 4 <= idx->bytes && n * elem_bytes <= dst->bytes, "hj_gather: buffer too small");
    DeviceGuard g(dev);
    settle_all(dev, {src, idx, dst});
    return launch_gather(dev, elem_bytes, n, src->ptr, (const uint32_t*)idx->ptr, dst->ptr);
}

}  // extern "C"
