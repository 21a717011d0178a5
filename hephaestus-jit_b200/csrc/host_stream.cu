// host_stream.cu — the device ops over HOST arrays: reduce / prefix-sum / compress of data that
// lives in (pinned) host memory, streamed through the GPU chunk by chunk.
//
// The reference's call sequence for data that starts and ends on the host is three blocking steps:
// tr::array (staged upload + fence, trace.rs:647-663, backend/vulkan/mod.rs:427-452), Graph::launch
// (blocking submit, vulkan_core/device.rs:297-334), to_vec (staged download + fence,
// trace.rs:1404-1438, backend/vulkan/mod.rs:478-509).  On a B200 every device op of this backend
// runs two orders of magnitude faster than the PCIe link feeds it, so the end-to-end cost of that
// sequence is the SUM of the transfer times.  These entry points overlap them instead, like
// hj_kernel_map_host does for fused elementwise kernels (jit.cpp): the array is cut into chunks,
// chunk c+1 is uploaded while chunk c is processed and chunk c-1 is downloaded (three streams:
// upload | the device stream | download; DEPTH device-side chunk buffers reused round-robin), and
// what must cross chunk boundaries stays on the device:
//   reduce      one partial per chunk, folded at the end            (host <- 1 element)
//   prefix sum  the running total is the `seed` of the next chunk's scan (scan.cu)
//   compress    `index_base` = first element of the chunk; the compacted indices of a chunk are
//               downloaded behind those of the chunks before it (the host reads each chunk's count
//               while the next chunks are already in flight)
// Results are complete on return, like BackendBuffer::to_host.  The device lock is only held while a
// chunk's work is being enqueued, never while waiting: several host threads may stream different
// arrays at once (bench.py's end-to-end step runs its four ops that way), their copies share the two
// PCIe directions and their kernels take turns on the device stream.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "hj_internal.h"

namespace hj {
namespace {

constexpr int DEPTH = 3;

// running total after a chunk: the last output (which already carries the seed), plus the last
// input for an exclusive scan
template <typename T>
__global__ void carry_kernel(T* carry, const T* __restrict__ out, const T* __restrict__ in, size_t n, int inclusive) {
    if (threadIdx.x == 0 && blockIdx.x == 0) carry[0] = inclusive ? out[n - 1] : (T)(out[n - 1] + in[n - 1]);
}

struct Pipe {
    hj_device* dev;
    cudaStream_t up = nullptr, down = nullptr;
    cudaEvent_t up_done[DEPTH] = {}, comp_done[DEPTH] = {}, down_done[DEPTH] = {}, start = nullptr;
    std::vector<void*> device_allocs;
    void* pinned = nullptr;

    explicit Pipe(hj_device* d) : dev(d) {}
    ~Pipe() {
        // (an error path may leave copies in flight on the side streams)
        if (up) cudaStreamSynchronize(up);
        if (down) cudaStreamSynchronize(down);
        {
            DeviceGuard g(dev);
            for (void* p : device_allocs) cudaFreeAsync(p, dev->stream);
        }
        for (int d = 0; d < DEPTH; d++) {
            if (up_done[d]) cudaEventDestroy(up_done[d]);
            if (comp_done[d]) cudaEventDestroy(comp_done[d]);
            if (down_done[d]) cudaEventDestroy(down_done[d]);
        }
        if (start) cudaEventDestroy(start);
        if (up) cudaStreamDestroy(up);
        if (down) cudaStreamDestroy(down);
        if (pinned) cudaFreeHost(pinned);
    }
    hj_status init(bool need_pinned) {
        HJ_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        HJ_CUDA(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
        for (int d = 0; d < DEPTH; d++) {
            HJ_CUDA(cudaEventCreateWithFlags(&up_done[d], cudaEventDisableTiming));
            HJ_CUDA(cudaEventCreateWithFlags(&comp_done[d], cudaEventDisableTiming));
            HJ_CUDA(cudaEventCreateWithFlags(&down_done[d], cudaEventDisableTiming));
        }
        HJ_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
        if (need_pinned) HJ_CUDA(cudaHostAlloc(&pinned, 256, cudaHostAllocDefault));
        return HJ_OK;
    }
    hj_status alloc(size_t bytes, void** out) {
        HJ_CUDA(cudaMallocAsync(out, bytes ? bytes : 16, dev->stream));
        device_allocs.push_back(*out);
        return HJ_OK;
    }
    // the side streams start after everything already enqueued on the device stream (incl. the allocations)
    hj_status fork() {
        HJ_CUDA(cudaEventRecord(start, dev->stream));
        HJ_CUDA(cudaStreamWaitEvent(up, start, 0));
        HJ_CUDA(cudaStreamWaitEvent(down, start, 0));
        return HJ_OK;
    }
    // Blocks until this op's own work is done: everything enqueued on the device stream so far (its
    // last kernel; recorded under the lock) and both side streams.  Does not wait for work other
    // threads enqueue afterwards.
    hj_status finish() {
        {
            DeviceGuard g(dev);
            HJ_CUDA(cudaEventRecord(start, dev->stream));
        }
        HJ_CUDA(cudaEventSynchronize(start));
        HJ_CUDA(cudaStreamSynchronize(up));
        HJ_CUDA(cudaStreamSynchronize(down));
        return HJ_OK;
    }
};

// (first element, count) of every chunk.  The fill of the pipeline (first upload) and its drain (last
// download) cannot overlap with traffic in the other direction, so the first chunks — and, when there
// is a download leg, the last ones — are small.
std::vector<std::pair<size_t, size_t>> make_chunks(size_t n, size_t chunk, size_t align, bool ramp_out) {
    chunk = std::max(align, (chunk + align - 1) / align * align);
    std::vector<std::pair<size_t, size_t>> out;
    size_t left = n;
    auto take = [&](size_t c) {
        c = std::min(std::max(c / align * align, align), left);
        if (c) {
            out.emplace_back(n - left, c);
            left -= c;
        }
    };
    static const bool ramp = !getenv("HJ_MAP_NO_RAMP");
    if (ramp && n >= 4 * chunk) {
        const size_t tail_total = ramp_out ? chunk / 2 + chunk / 4 + chunk / 8 : 0;
        for (size_t div = 8; div >= 2; div /= 2) take(chunk / div);
        while (left > tail_total + chunk) take(chunk);
        if (ramp_out) {
            while (left > tail_total) take(std::min(chunk, left - tail_total));
            for (size_t div = 2; div <= 8; div *= 2) take(chunk / div);
        }
    }
    while (left) take(chunk);
    return out;
}

// carry[0] = running total after the chunk `in` -> `out` of `count` elements (on the device stream)
void launch_carry(hj_device* dev, hj_type_kind ty, size_t es, void* carry, const void* out, const void* in, size_t count,
                  int inclusive) {
    switch (es) {
    case 1: carry_kernel<uint8_t><<<1, 32, 0, dev->stream>>>((uint8_t*)carry, (const uint8_t*)out, (const uint8_t*)in, count, inclusive); break;
    case 2: carry_kernel<uint16_t><<<1, 32, 0, dev->stream>>>((uint16_t*)carry, (const uint16_t*)out, (const uint16_t*)in, count, inclusive); break;
    case 4:
        if (ty == HJ_F32) carry_kernel<float><<<1, 32, 0, dev->stream>>>((float*)carry, (const float*)out, (const float*)in, count, inclusive);
        else carry_kernel<uint32_t><<<1, 32, 0, dev->stream>>>((uint32_t*)carry, (const uint32_t*)out, (const uint32_t*)in, count, inclusive);
        break;
    default:
        if (ty == HJ_F64) carry_kernel<double><<<1, 32, 0, dev->stream>>>((double*)carry, (const double*)out, (const double*)in, count, inclusive);
        else carry_kernel<unsigned long long><<<1, 32, 0, dev->stream>>>((unsigned long long*)carry, (const unsigned long long*)out, (const unsigned long long*)in, count, inclusive);
        break;
    }
}

size_t default_chunk(size_t chunk_elems) { return chunk_elems ? chunk_elems : ((size_t)1 << 24); }

}  // namespace

// A PrefixSum pass whose source is still arriving chunk by chunk (hj_buffer_create_from_host_async): every chunk
// is scanned on the device stream as soon as its upload event fires, with the running total of the chunks in front
// as its seed, and `dst` inherits the schedule so that hj_buffer_to_host drains it chunk by chunk — the download of
// chunk c overlaps the upload of chunk c + 2.  Integer types only: a float scan cut at other places rounds
// differently, and the result must not depend on how the source got to the device.
hj_status prefix_sum_arriving(hj_device* dev, hj_type_kind ty, size_t n, bool inclusive, hj_buffer* src, hj_buffer* dst,
                              bool* done) {
    *done = false;
    const size_t es = type_size(ty);
    const bool integer = ty != HJ_F16 && ty != HJ_F32 && ty != HJ_F64 && ty != HJ_BOOL;
    static const bool off = getenv("HJ_NO_STREAMED_LAUNCH") != nullptr;
    if (off || !es || !integer || !src || !dst || src == dst || n == 0) return HJ_OK;
    DeviceGuard g(dev);
    std::shared_ptr<AsyncProgress> pr = src->progress;
    if (!pr || pr->first.empty() || src->progress_elem_bytes != es || pr->first.back() + pr->count.back() != n) return HJ_OK;
    if (src->dev != dev || dst->dev != dev || n * es > src->bytes || n * es > dst->bytes || src->ptr == dst->ptr) return HJ_OK;
    settle_locked(dst);
    auto out = std::make_shared<AsyncProgress>();
    out->first = pr->first;
    out->count = pr->count;
    for (size_t c = 0; c < pr->first.size(); c++) {
        cudaEvent_t ev = nullptr;
        HJ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        out->done.push_back(ev);
    }
    void* carry = nullptr;
    HJ_CUDA(cudaMallocAsync(&carry, 16, dev->stream));
    hj_status st = HJ_OK;
    for (size_t c = 0; c < pr->first.size() && st == HJ_OK; c++) {
        const size_t first = pr->first[c], count = pr->count[c];
        const char* in = (const char*)src->ptr + first * es;
        char* o = (char*)dst->ptr + first * es;
        cudaError_t e = cudaStreamWaitEvent(dev->stream, pr->done[c], 0);
        if (e != cudaSuccess) { cudaGetLastError(); st = fail(HJ_ERR_CUDA, "streamed scan: %s", cudaGetErrorString(e)); break; }
        st = launch_prefix_sum(dev, ty, count, inclusive, in, o, c ? carry : nullptr);
        if (st != HJ_OK) break;
        if (c + 1 < pr->first.size()) {
            launch_carry(dev, ty, es, carry, o, in, count, inclusive ? 1 : 0);
            st = check_launch(dev, "carry_kernel");
            if (st != HJ_OK) break;
        }
        e = cudaEventRecord(out->done[c], dev->stream);
        if (e != cudaSuccess) { cudaGetLastError(); st = fail(HJ_ERR_CUDA, "streamed scan: %s", cudaGetErrorString(e)); }
    }
    cudaFreeAsync(carry, dev->stream);
    // the device stream has waited for every upload event (or the error path below waits for the last one)
    settle_locked(src);
    HJ_TRY(st);
    attach_progress(dst, std::move(out), (uint32_t)es);
    *done = true;
    return HJ_OK;
}
}  // namespace hj

using namespace hj;

extern "C" {

hj_status hj_reduce_host(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n, const void* host_src,
                         void* host_dst, size_t chunk_elems) {
    HJ_REQUIRE(dev && host_src && host_dst, "hj_reduce_host: null argument");
    const size_t es = type_size(ty);
    HJ_REQUIRE(es && n >= 1, "hj_reduce_host: bad type or empty input");
    const auto chunks = make_chunks(n, default_chunk(chunk_elems), 16 / std::min<size_t>(es, 16), false);
    size_t max_chunk = 0;
    for (auto& c : chunks) max_chunk = std::max(max_chunk, c.second);
    Pipe p(dev);
    void *slot[DEPTH], *partials = nullptr, *result = nullptr;
    {
        DeviceGuard g(dev);
        HJ_TRY(p.init(false));
        for (int d = 0; d < DEPTH; d++) HJ_TRY(p.alloc(max_chunk * es, &slot[d]));
        HJ_TRY(p.alloc(chunks.size() * es, &partials));
        HJ_TRY(p.alloc(16, &result));
        HJ_TRY(p.fork());
    }
    for (size_t c = 0; c < chunks.size(); c++) {
        const int d = (int)(c % DEPTH);
        DeviceGuard g(dev);  // per chunk: other threads' ops interleave between chunks
        if (c >= DEPTH) HJ_CUDA(cudaStreamWaitEvent(p.up, p.comp_done[d], 0));  // the slot has been reduced
        HJ_CUDA(cudaMemcpyAsync(slot[d], (const char*)host_src + chunks[c].first * es, chunks[c].second * es,
                                cudaMemcpyHostToDevice, p.up));
        HJ_CUDA(cudaEventRecord(p.up_done[d], p.up));
        HJ_CUDA(cudaStreamWaitEvent(dev->stream, p.up_done[d], 0));
        HJ_TRY(launch_reduce(dev, op, ty, chunks[c].second, slot[d], (char*)partials + c * es));
        HJ_CUDA(cudaEventRecord(p.comp_done[d], dev->stream));
    }
    {
        DeviceGuard g(dev);
        const void* final_src = partials;
        if (chunks.size() > 1) {
            HJ_TRY(launch_reduce(dev, op, ty, chunks.size(), partials, result));
            final_src = result;
        }
        // the result leaves on the download stream, behind the last kernel
        HJ_CUDA(cudaEventRecord(p.comp_done[0], dev->stream));
        HJ_CUDA(cudaStreamWaitEvent(p.down, p.comp_done[0], 0));
        HJ_CUDA(cudaMemcpyAsync(host_dst, final_src, es, cudaMemcpyDeviceToHost, p.down));
    }
    return p.finish();
}

hj_status hj_prefix_sum_host(hj_device* dev, hj_type_kind ty, size_t n, int32_t inclusive, const void* host_src,
                             void* host_dst, size_t chunk_elems) {
    HJ_REQUIRE(dev && host_src && host_dst, "hj_prefix_sum_host: null argument");
    const size_t es = type_size(ty);
    HJ_REQUIRE(es && n >= 1, "hj_prefix_sum_host: bad type or empty input");
    if (getenv("HJ_REF_COMPAT")) inclusive = 1;  // reference defect D10: always inclusive
    const auto chunks = make_chunks(n, default_chunk(chunk_elems), 16 / std::min<size_t>(es, 16), true);
    size_t max_chunk = 0;
    for (auto& c : chunks) max_chunk = std::max(max_chunk, c.second);
    Pipe p(dev);
    void *in[DEPTH], *out[DEPTH], *carry = nullptr;
    {
        DeviceGuard g(dev);
        HJ_TRY(p.init(false));
        for (int d = 0; d < DEPTH; d++) {
            HJ_TRY(p.alloc(max_chunk * es, &in[d]));
            HJ_TRY(p.alloc(max_chunk * es, &out[d]));
        }
        HJ_TRY(p.alloc(16, &carry));
        HJ_CUDA(cudaMemsetAsync(carry, 0, 16, dev->stream));
        HJ_TRY(p.fork());
    }
    for (size_t c = 0; c < chunks.size(); c++) {
        const int d = (int)(c % DEPTH);
        const size_t first = chunks[c].first, count = chunks[c].second;
        DeviceGuard g(dev);  // per chunk
        if (c >= DEPTH) HJ_CUDA(cudaStreamWaitEvent(p.up, p.down_done[d], 0));  // the slot's previous chunk has left
        HJ_CUDA(cudaMemcpyAsync(in[d], (const char*)host_src + first * es, count * es, cudaMemcpyHostToDevice, p.up));
        HJ_CUDA(cudaEventRecord(p.up_done[d], p.up));
        HJ_CUDA(cudaStreamWaitEvent(dev->stream, p.up_done[d], 0));
        if (c >= DEPTH) HJ_CUDA(cudaStreamWaitEvent(dev->stream, p.down_done[d], 0));
        HJ_TRY(launch_prefix_sum(dev, ty, count, inclusive != 0, in[d], out[d], carry));
        if (c + 1 < chunks.size()) {
            launch_carry(dev, ty, es, carry, out[d], in[d], count, inclusive);
            HJ_TRY(check_launch(dev, "carry_kernel"));
        }
        HJ_CUDA(cudaEventRecord(p.comp_done[d], dev->stream));
        HJ_CUDA(cudaStreamWaitEvent(p.down, p.comp_done[d], 0));
        HJ_CUDA(cudaMemcpyAsync((char*)host_dst + first * es, out[d], count * es, cudaMemcpyDeviceToHost, p.down));
        HJ_CUDA(cudaEventRecord(p.down_done[d], p.down));
    }
    return p.finish();
}

hj_status hj_compress_host(hj_device* dev, size_t n, const uint8_t* host_mask, uint32_t* host_index_out,
                           uint32_t* host_count, uint32_t index_base, size_t chunk_elems) {
    HJ_REQUIRE(dev && host_mask && host_index_out && host_count, "hj_compress_host: null argument");
    HJ_REQUIRE(n >= 1 && n <= 0xffffffffull, "hj_compress_host: n must be in [1, 2^32)");
    const auto chunks = make_chunks(n, default_chunk(chunk_elems), 16, true);
    size_t max_chunk = 0;
    for (auto& c : chunks) max_chunk = std::max(max_chunk, c.second);
    Pipe p(dev);
    void *mask[DEPTH], *idx[DEPTH];
    {
        DeviceGuard g(dev);
        HJ_TRY(p.init(true));
        for (int d = 0; d < DEPTH; d++) {
            HJ_TRY(p.alloc(max_chunk, &mask[d]));
            HJ_TRY(p.alloc(max_chunk * 4, &idx[d]));
        }
        HJ_TRY(p.fork());
    }
    uint32_t* h_cnt = reinterpret_cast<uint32_t*>(p.pinned);
    size_t total = 0, finalized = 0;
    // the host reads a chunk's count (it decides where the next chunk's indices go) and enqueues the
    // download; by then DEPTH - 1 later chunks are already in flight
    auto finalize = [&](size_t k) -> hj_status {
        const int d = (int)(k % DEPTH);
        HJ_CUDA(cudaEventSynchronize(p.comp_done[d]));
        const uint32_t c = h_cnt[d * 4];
        HJ_CUDA(cudaStreamWaitEvent(p.down, p.comp_done[d], 0));
        if (c) HJ_CUDA(cudaMemcpyAsync(host_index_out + total, idx[d], (size_t)c * 4, cudaMemcpyDeviceToHost, p.down));
        HJ_CUDA(cudaEventRecord(p.down_done[d], p.down));
        total += c;
        return HJ_OK;
    };
    for (size_t c = 0; c < chunks.size(); c++) {
        const int d = (int)(c % DEPTH);
        const size_t first = chunks[c].first, count = chunks[c].second;
        if (c >= DEPTH)
            while (finalized + DEPTH <= c) HJ_TRY(finalize(finalized++));  // waits on the host: no lock held
        DeviceGuard g(dev);  // per chunk
        if (c >= DEPTH) HJ_CUDA(cudaStreamWaitEvent(p.up, p.down_done[d], 0));
        HJ_CUDA(cudaMemcpyAsync(mask[d], host_mask + first, count, cudaMemcpyHostToDevice, p.up));
        HJ_CUDA(cudaEventRecord(p.up_done[d], p.up));
        HJ_CUDA(cudaStreamWaitEvent(dev->stream, p.up_done[d], 0));
        if (c >= DEPTH) HJ_CUDA(cudaStreamWaitEvent(dev->stream, p.down_done[d], 0));
        // the kernel writes the chunk's count straight into pinned host memory (unified addressing: the
        // pinned pointer is valid on the device), so no small copy queues behind the index downloads
        HJ_TRY(launch_compress(dev, count, nullptr, h_cnt + d * 4, (const uint8_t*)mask[d], (uint32_t*)idx[d],
                               index_base + (uint32_t)first));
        HJ_CUDA(cudaEventRecord(p.comp_done[d], dev->stream));
    }
    while (finalized < chunks.size()) HJ_TRY(finalize(finalized++));
    HJ_TRY(p.finish());
    *host_count = (uint32_t)total;
    return HJ_OK;
}

}  // extern "C"
