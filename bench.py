#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native hephaestus-jit backend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-suite]

For N > 1 it is launched by torchrun (one rank per GPU); rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1], "C2"): the fused elementwise chain
    t = fma(x, 1.5, 0.25);  y = select(x > 0, sin(t), exp2(t))
over 2^28 f32 per GPU, traced into one kernel, executed through the reference-facing boundary
(hj_execute_graph: IR -> CUDA C++ -> NVRTC sm_100a cubin, cached by IR hash).  A "step" is one
pass over the 2^28-element array.  Algorithmic bytes: 8 per element.

  value      GB/s of algorithmic bytes with inputs resident in HBM, K steps timed with CUDA
             events on the stream the kernels run on, max over ranks (N ranks: weak scaling,
             every rank owns its own 2^28 elements, Index is global = rank * 2^28 + i).
  e2e        same metric through the C ABI with HOST buffers: every step uploads the input
             from pinned host memory (hj_buffer_upload), runs the graph, and reads the full
             result back (hj_buffer_to_host) inside the timed region.
  roofline   of the fused kernel against the measured copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the oracle's C restatement of the same chain on all host cores, bounded sample.
  suite      the other hot-path kernels of BASELINE.json (C3 reduce, C4 scan + compress,
             C5 histogram) at their named sizes, each with GB/s and fraction of the peak.

`--impl reference` times the reference's path on the host CPU.  The reference itself (nightly
Rust + Vulkan on lavapipe) cannot be built or run in this image, so this arm
runs the oracle port (oracle/hj_oracle.c) with all host threads on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GB/s & % of HBM roofline for fused map/reduce/scan/compress"
LOG2N = 28
BYTES_PER_ELEM = 8  # 1 f32 read + 1 f32 written


def measured_peak():
    """(GB/s, source) — the driver-measured copy bandwidth, else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic():
    """DRAM bytes per launch of the fused kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons = index, period_s, [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host CPU
# --------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: ask for every core
    cores = oracle.get_threads()
    n = 1 << 24  # bounded sample of the 2^28-element workload per step
    rng = np.random.Generator(np.random.PCG64(0))
    x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    for _ in range(max(args.warmup, 1)):
        oracle.c2_chain(x, fast=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.c2_chain(x, fast=True)
    dt = (time.perf_counter() - t0) / args.steps
    gbs = BYTES_PER_ELEM * n / dt / 1e9
    sample = f"2^24 of the 2^28 elements per step, oracle C port (libm sinf/exp2f), {cores} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            # the own arm's config (same workload, same metric and unit); what one CPU step covers is in `sample`
            "workload": "C2: fused elementwise chain fma -> sin/exp2 -> select over 2^28 f32 per GPU, one NVRTC kernel",
            "elements_per_gpu": 1 << 28, "bytes_per_element": BYTES_PER_ELEM,
            "sample": "each step = 2^24 of the 2^28 elements on the host cores",
            "note": "reference (nightly Rust + Vulkan/lavapipe) is not buildable in this image; oracle port timed instead"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------

def cpu_baseline(check_inputs=None):
    """The one place of the own arm that executes anything under oracle/: times the CPU port on a
    bounded sample and, while it is at it, evaluates `check_inputs` ({name: f32 array}) so that the
    caller can compare the GPU results with them.  Returns (report, {name: oracle output})."""
    import oracle
    oracle.set_threads(os.cpu_count() or 1)
    cores = oracle.get_threads()
    wants = {k: oracle.c2_chain(np.ascontiguousarray(v)) for k, v in (check_inputs or {}).items()}
    n = 1 << 24
    rng = np.random.Generator(np.random.PCG64(0))
    x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    oracle.c2_chain(x, fast=True)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 5.0 or reps < 3:  # bounded: a few seconds of CPU work
        oracle.c2_chain(x, fast=True)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": BYTES_PER_ELEM * n / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"2^24 of the 2^28 elements, {reps} reps, oracle C port with libm sinf/exp2f on {cores} threads"}, wants


def timed_events(torch, fn, iters, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def run_suite(torch, hj, dev, peak, world, rank, comm):
    """Per-kernel GB/s at the sizes BASELINE.json names (per-GPU share of the array on N GPUs)."""
    out = {}
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
    n30 = (1 << 30) // world
    iters = 10

    def entry(nbytes, ms, extra=None):
        gbs = nbytes / ms / 1e6
        e = {"GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4), "ms": round(ms, 4)}
        if extra:
            e.update(extra)
        return e

    xf = torch.rand(n30, device="cuda", generator=g, dtype=torch.float32)
    xu = torch.randint(0, 4, (n30,), device="cuda", generator=g, dtype=torch.int32)
    o1 = torch.zeros(16, device="cuda", dtype=torch.float32)
    on = torch.empty(n30, device="cuda", dtype=torch.int32)
    bf, bu, bo1, bon = wrap(xf), wrap(xu), wrap(o1), wrap(on)
    for name, op, ty, buf in (("C3 reduce sum f32 2^30", hj.SUM, hj.F32, bf), ("C3 reduce max f32 2^30", hj.MAX, hj.F32, bf),
                              ("C3 reduce sum u32 2^30", hj.SUM, hj.U32, bu), ("C3 reduce min u32 2^30", hj.MIN, hj.U32, bu)):
        ms = timed_events(torch, lambda: dev.reduce(op, ty, n30, buf, bo1), iters, 3)
        out[name] = entry(4 * n30, ms, {"elements_per_s": n30 / ms * 1e3, "bytes_per_elem": 4})
    ms = timed_events(torch, lambda: dev.prefix_sum(hj.U32, n30, True, bu, bon), iters, 3)
    out["C4 inclusive scan u32 2^30"] = entry(8 * n30, ms, {"elements_per_s": n30 / ms * 1e3, "bytes_per_elem": 8})
    for p in (0.5, 0.01, 0.99):
        mask = (torch.rand(n30, device="cuda", generator=g) < p).to(torch.uint8)
        cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
        bm, bc = wrap(mask), wrap(cnt)
        ms = timed_events(torch, lambda: dev.compress(n30, bc, bm, bon), iters, 3)
        c = int(cnt.item())
        out[f"C4 compress p={p} 2^30"] = entry(n30 + 4 * c, ms, {"elements_per_s": n30 / ms * 1e3,
                                                                   "bytes_per_elem": round(1 + 4 * c / n30, 3)})
        del mask
    n28 = (1 << 28) // world
    keys = torch.randint(0, 1 << 16, (n28,), device="cuda", generator=g, dtype=torch.int32)
    hist = torch.zeros(1 << 16, device="cuda", dtype=torch.int32)
    bk, bh = wrap(keys), wrap(hist)
    ms = timed_events(torch, lambda: dev.scatter_reduce(hj.SUM, hj.U32, n28, bk, None, 1, bh, 1 << 16), iters, 3)
    out["C5 histogram 2^28 keys -> 2^16 bins"] = entry(4 * n28, ms, {"elements_per_s": n28 / ms * 1e3, "bytes_per_elem": 4,
                                                                      "bound": "shared-memory atomic throughput, not HBM"})
    # C5b: gather (12 B/elem nominal: u32 index + f32 value + f32 out) and the traced Monte-Carlo loop
    del keys, hist
    idx = torch.randint(0, 1 << 20, (n28,), device="cuda", generator=g, dtype=torch.int32)
    outf = torch.empty(n28, device="cuda", dtype=torch.float32)
    bi, bof = wrap(idx), wrap(outf)
    for log_t in (20, 28):
        table = torch.rand(1 << log_t, device="cuda", generator=g, dtype=torch.float32)
        if log_t == 28:
            idx = torch.randint(0, 1 << 28, (n28,), device="cuda", generator=g, dtype=torch.int32)
            bi = wrap(idx)
        bt = wrap(table)
        ms = timed_events(torch, lambda: dev.gather(4, n28, bt, bi, bof), iters, 3)
        out[f"C5 gather f32 2^28 indices, table 2^{log_t}"] = entry(12 * n28, ms, {
            "elements_per_s": n28 / ms * 1e3, "bytes_per_elem": 12,
            "bound": "L2-resident table: HBM streams of idx/out" if log_t == 20 else "DRAM sectors: 32 B fetched per 4 B used"})
        del table
    del idx, outf
    if world == 1:
        tr = importlib.import_module("hephaestus-jit_b200.tr")
        n_l, k_it, log_t = 1 << 24, 16, 20
        table = tr.from_buffer(wrap(torch.rand(1 << log_t, device="cuda", generator=g, dtype=torch.float32)), hj.F32, 1 << log_t)

        def mc_graph():
            s0 = tr.sized_index(n_l).mul(tr.literal(2654435761, hj.U32)).add(tr.literal(12345, hj.U32))
            acc0, it0, c0 = tr.sized_literal(0.0, n_l, hj.F32), tr.sized_literal(0, n_l, hj.U32), tr.literal(True)

            def body(c, vs):
                s_, acc, it = vs
                s_ = s_.mul(tr.literal(1664525, hj.U32)).add(tr.literal(1013904223, hj.U32))
                acc = acc.add(table.gather(s_.shr(tr.literal(32 - log_t, hj.U32))))
                it = it.add(tr.literal(1, hj.U32))
                return c.and_(it.lt(tr.literal(k_it, hj.U32))), [s_, acc, it]

            c1, (s1, acc1, it1) = tr.loop_record(c0, [s0, acc0, it0], body)
            total = acc1.reduce_sum()
            total.schedule()
            return tr.compile(), total

        graph, total = mc_graph()
        ms = timed_events(torch, lambda: graph.launch(dev), 5, 2)
        out["C5 traced Monte-Carlo loop: 2^24 lanes x 16 gathers (table 2^20) + reduce_sum"] = {
            "ms": round(ms, 4), "lanes_per_s": n_l / ms * 1e3, "gathers_per_s": n_l * k_it / ms * 1e3,
            "passes": graph.n_passes(), "sum": float(total.item())}
        del graph, total, table
        # C4 through the trace layer: `mask.compress()` as a user of the reference writes it — the
        # scheduler's zero-fill of the index buffer + the Compress pass (the backend
        # leaves the zero-fill to the compaction).  5 B/elem = 1 (mask) + 4 (every index word is written once)
        try:
            n_c = 1 << 28
            mask_t = (torch.rand(n_c, device="cuda", generator=g) < 0.5).to(torch.uint8)
            mvar = tr.from_buffer(wrap(mask_t), hj.BOOL, n_c)
            count, index = mvar.compress()
            cgraph = tr.compile()
            ms = timed_events(torch, lambda: cgraph.launch(dev), 5, 3)
            c = int(count.to_vec(np.uint32)[0])
            ok = c == int(mask_t.sum().item()) and bool((index.to_vec(np.uint32, c, min(c + 4096, n_c)) == 0).all())
            out["C4 traced mask.compress() p=0.5 2^28 (zero-fill + Compress passes)"] = entry(5 * n_c, ms, {
                "elements_per_s": n_c / ms * 1e3, "bytes_per_elem": 5, "passes": cgraph.n_passes(), "count_and_tail_ok": ok})
            del count, index, cgraph, mvar, mask_t
        except Exception as exc:  # an extra line of the suite must never take the bench down
            out["C4 traced mask.compress() p=0.5 2^28 (zero-fill + Compress passes)"] = {"error": str(exc)[:200]}
    if world > 1:
        out["_note"] = f"per-GPU share (1/{world}) of each array, local kernels only; collectives reported under 'sharded'"
    return out


def run_sharded(torch, dist, hj, dev, peak, world, rank):
    """The sharded ops of BASELINE.json (C3 / C4 / C5) INCLUDING their NCCL exchange: 2^30 (2^28 for
    the histogram) elements split contiguously over the ranks, timed with CUDA events on the launch
    stream, max over ranks.  GB/s = GLOBAL algorithmic bytes / time; the scan moves 12 B/elem when
    sharded (shard totals first) but is credited with 8."""
    sharded = importlib.import_module("hephaestus-jit_b200.sharded")
    comm = sharded.Comm.from_torch(dev)
    out = {}
    g = torch.Generator(device="cuda").manual_seed(4321 + rank)
    wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
    n30, n28 = (1 << 30) // world, (1 << 28) // world

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in ev:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[iters // 2]
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def entry(nbytes_global, ms):
        gbs = nbytes_global / ms / 1e6
        return {"GB/s": round(gbs, 1), "frac_of_measured_peak_x_gpus": round(gbs / (peak * world), 4), "ms": round(ms, 4)}

    xf = torch.rand(n30, device="cuda", generator=g, dtype=torch.float32)
    xu = torch.randint(0, 4, (n30,), device="cuda", generator=g, dtype=torch.int32)
    o1 = torch.zeros(16, device="cuda", dtype=torch.float32)
    on = torch.empty(n30, device="cuda", dtype=torch.int32)
    bf, bu, bo1, bon = wrap(xf), wrap(xu), wrap(o1), wrap(on)
    out["C3 sharded reduce sum f32 2^30 + all-gather"] = entry(4 * n30 * world, timed(lambda: comm.reduce(hj.SUM, hj.F32, n30, bf, bo1)))
    out["C3 sharded reduce max u32 2^30 + all-gather"] = entry(4 * n30 * world, timed(lambda: comm.reduce(hj.MAX, hj.U32, n30, bu, bo1)))
    out["C4 sharded inclusive scan u32 2^30 (totals pass + all-gather + seeded scan)"] = entry(
        8 * n30 * world, timed(lambda: comm.prefix_sum(hj.U32, n30, True, bu, bon)))
    mask = (torch.rand(n30, device="cuda", generator=g) < 0.5).to(torch.uint8)
    cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
    counts = torch.zeros(world, device="cuda", dtype=torch.int32)
    bm, bc, bcs = wrap(mask), wrap(cnt), wrap(counts)
    ms = timed(lambda: comm.compress(n30, (rank * n30) & 0xFFFFFFFF, bm, bon, bc, bcs))
    out["C4 sharded compress p=0.5 2^30 + all-gather of counts"] = entry(n30 * world + 4 * int(cnt.item()), ms)
    keys = torch.randint(0, 1 << 16, (n28,), device="cuda", generator=g, dtype=torch.int32)
    hist = torch.zeros(1 << 16, device="cuda", dtype=torch.int32)
    bk, bh = wrap(keys), wrap(hist)
    out["C5 sharded histogram 2^28 keys -> 2^16 bins + all-reduce"] = entry(
        4 * n28 * world, timed(lambda: comm.scatter_reduce(hj.SUM, hj.U32, n28, bk, None, 1, bh, 1 << 16)))
    comm.destroy()
    return out


def run_own(args):
    import torch
    import torch.distributed as dist

    # stdout carries exactly ONE line, the JSON: anything native libraries print on fd 1 meanwhile
    # (NCCL's version banner, for one) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    hj = importlib.import_module("hephaestus-jit_b200")  # raises if libhj_b200.so is missing
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    dev = hj.Device.cuda(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev.set_stream(stream.cuda_stream)  # kernels and torch's timing events share one stream

    peak, peak_src = measured_peak()
    n = 1 << LOG2N
    nbytes_step = BYTES_PER_ELEM * n

    # synthetic input, resident in HBM: f32 uniform [-4, 4), seeded
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.rand(n, device="cuda", generator=g, dtype=torch.float32) * 8 - 4
    y = torch.empty_like(x)
    bx = dev.wrap(x.data_ptr(), 4 * n)
    by = dev.wrap(y.data_ptr(), 4 * n)

    # the graph Graph::launch hands to BackendDevice::execute_graph for this trace: one Kernel pass
    ir = irm.c2_chain_ir()
    passes = [{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": ir, "size": n}]
    descs = [(n, hj.F32, 4), (n, hj.F32, 4)]
    kernel = dev.kernel(ir)  # compile outside the timed region (cached by IR hash afterwards)
    index_base = (rank * n) & 0xFFFFFFFF

    def step():
        if world > 1:
            dev.launch(kernel, n, [bx, by], index_base=index_base)  # global Index on a shard
        else:
            dev.execute_graph(passes, [bx, by], descs)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = dev.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0.record()
    for a, b in per:
        a.record()
        step()
        b.record()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    launches = dev.launch_count() - launches0
    # keep the sampler alive long enough to have seen the load even for very short runs
    if total_ms < 50:
        t_end = time.perf_counter() + 0.3
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in per]))

    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * nbytes_step / (ms_per_step * 1e-3) / 1e9

    # ---- correctness spot check (outside the timed region): samples for the oracle, which only the
    # cpu_baseline leg below executes
    check_in, check_got = {}, {}
    if rank == 0:
        m = 1 << 16
        check_in["resident"], check_got["resident"] = x[:m].cpu().numpy(), y[:m].cpu().numpy()

    # ---- end-to-end through the C ABI with HOST buffers (pinned), rank-local.  Two public paths:
    #  (a) pipelined: hj_kernel_map_host streams chunks through upload / kernel / download streams
    #      (the headline e2e: every input byte crosses PCIe in, every output byte out, per step);
    #  (b) blocking, the reference's own call sequence: upload -> execute_graph -> to_host.
    import ctypes
    L = importlib.import_module("hephaestus-jit_b200._lib")
    hx, hy = ctypes.c_void_p(), ctypes.c_void_p()
    L.check(L.lib.hj_host_alloc(4 * n, ctypes.byref(hx)))
    L.check(L.lib.hj_host_alloc(4 * n, ctypes.byref(hy)))
    host_x = np.ctypeslib.as_array(ctypes.cast(hx, ctypes.POINTER(ctypes.c_float)), shape=(n,))
    host_y = np.ctypeslib.as_array(ctypes.cast(hy, ctypes.POINTER(ctypes.c_float)), shape=(n,))
    host_x[:] = np.random.Generator(np.random.PCG64(rank)).random(n, dtype=np.float32) * 8 - 4
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_pipelined():
        dev.map_host(kernel, n, [hx.value, hy.value], 1 << 23)

    dx, dy = dev.create_buffer(4 * n), dev.create_buffer(4 * n)

    def e2e_blocking():
        L.check(L.lib.hj_buffer_upload(dx.handle, 0, hx, 4 * n))            # H2D, pinned source
        dev.execute_graph(passes, [dx, dy], descs)
        L.check(L.lib.hj_buffer_to_host(dy.handle, 0, 4 * n, hy))           # D2H, blocks until visible

    def time_e2e(fn):
        fn()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        dev.sync()
        dt = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    launches_e2e0 = dev.launch_count()
    pipe_s = time_e2e(e2e_pipelined)
    launches_e2e = (dev.launch_count() - launches_e2e0) // (e2e_steps + 1)
    if rank == 0:
        m = 1 << 16
        check_in["e2e"], check_got["e2e"] = host_x[-m:].copy(), host_y[-m:].copy()
    block_s = time_e2e(e2e_blocking)
    e2e = {"value": world * nbytes_step / pipe_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * n,
           "d2h_bytes_per_step": 4 * n, "ms_per_step": pipe_s * 1e3, "steps": e2e_steps,
           "path": "pinned host arrays -> hj_kernel_map_host (8 Mi-element chunks, ramped at both ends: upload | kernel | download streams)",
           "kernel_launches_per_step": int(launches_e2e), "oracle_check": None,
           "blocking_path": {"value": world * nbytes_step / block_s / 1e9, "ms_per_step": block_s * 1e3,
                             "path": "hj_buffer_upload (pinned) -> hj_execute_graph -> hj_buffer_to_host"}}
    L.lib.hj_host_free(hx)
    L.lib.hj_host_free(hy)
    del dx, dy

    suite = None
    shard_suite = None
    if not args.no_suite:
        del x, y
        torch.cuda.empty_cache()
        suite = run_suite(torch, hj, dev, peak, world, rank, None)
        if world > 1:
            torch.cuda.empty_cache()
            shard_suite = run_sharded(torch, dist, hj, dev, peak, world, rank)

    if rank == 0:
        cpu_report, wants = cpu_baseline(check_in)
        ok = {k: bool(np.allclose(check_got[k], wants[k], rtol=4e-7, atol=1e-7)) for k in wants}
        check, e2e["oracle_check"] = ok.get("resident"), ok.get("e2e")
        achieved = nbytes_step / (kernel_ms * 1e-3) / 1e9
        traffic = ncu_traffic().get("fused_c2_kernel_dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "C2: fused elementwise chain fma -> sin/exp2 -> select over 2^28 f32 per GPU, one NVRTC kernel",
                "elements_per_gpu": n, "bytes_per_element": BYTES_PER_ELEM,
                "l2_policy": "inputs (1 GiB in + 1 GiB out per step) are larger than the 126 MB L2; no flush needed",
                "boundary": "hj_execute_graph (N=1) / hj_kernel_launch with index_base = rank*2^28 (N>1)",
                "parallelism": f"{world} independent shards, no data-path collective",
            },
            "elements_per_s": world * n / (ms_per_step * 1e-3),
            "frac_of_measured_peak": value / (world * peak),
            "frac_of_8TBps_nominal": value / (world * 8000.0),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "hj_kernel_vec (NVRTC, C2 IR)", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": nbytes_step},
            "cpu_baseline": cpu_report,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "oracle_check": check,
            "kernel_cache": dev.kernel_cache_stats(),
        }
        if suite is not None:
            line["suite"] = suite
        if shard_suite is not None:
            line["sharded"] = shard_suite
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-suite", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
