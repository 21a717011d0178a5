#!/usr/bin/env python3
"""Host-side cost of re-launching a compiled Graph (Graph::launch_with, graph.rs:192-400) on tiny
arrays, where nothing but launch overhead matters: the aliasing1 graph of the reference
(test.rs:1638-1670: three chained kernels, two aliased temporaries) and a reduce + scan graph."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
tr = importlib.import_module("hephaestus-jit_b200.tr")
dev = hj.Device.cuda(0)

def build_chain():
    x = tr.sized_literal(1, 100, hj.I32); x.schedule(); tr.schedule_eval()
    y = x.add(tr.literal(1, hj.I32)); y.schedule(); tr.schedule_eval()
    z = y.add(tr.literal(1, hj.I32)); z.schedule(); tr.schedule_eval()
    g = tr.compile()
    return g, (x, y, z)

def build_ops():
    x = tr.sized_index(4096).cast(hj.F32)
    s = x.reduce_sum()
    p = tr.sized_index(4096).prefix_sum(True)
    s.schedule(); p.schedule()
    return tr.compile(), (x, s, p)

for name, build in (("3 chained kernels (aliasing1)", build_chain), ("kernel + reduce + kernel + scan", build_ops)):
    g, keep = build()
    for _ in range(20): g.launch(dev)
    dev.sync()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n): g.launch(dev)
    t1 = time.perf_counter()
    dev.sync()
    t2 = time.perf_counter()
    mode = "pass by pass (HJ_NO_CUDA_GRAPHS)" if os.environ.get("HJ_NO_CUDA_GRAPHS") else "captured CUDA graph replay"
    print(f"{name} [{mode}]: {g.n_passes()} passes, {(t1-t0)/n*1e6:.1f} us per launch_with call (host), "
          f"{(t2-t0)/n*1e6:.1f} us incl. GPU drain; graph cache (captured, replayed, plain) = {dev.graph_cache_stats()}")
    del g, keep
