#!/usr/bin/env python3
"""Host-side counterpart of the reference's criterion bench `compile` (benches/compile.rs:6-27):
trace a chain of 10 000 dependent adds, schedule the last one and time trace + graph compile
(trace -> schedule groups -> IR lowering).  No device is needed."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
tr = importlib.import_module("hephaestus-jit_b200.tr")

def once(n):
    t0 = time.perf_counter()
    x = tr.sized_literal(1, 2, hj.I32)
    for i in range(n):
        x = x.add(tr.literal(i, hj.I32))
    x.schedule()
    t1 = time.perf_counter()
    g = tr.compile()
    t2 = time.perf_counter()
    assert g.n_passes() == 1
    return t1 - t0, t2 - t1

for n in (100, 1000, 10000):
    best = min((once(n) for _ in range(5)), key=lambda p: p[0] + p[1])
    print(f"n={n:6d} chained adds: trace {best[0]*1e3:8.2f} ms (python + ctypes + C++), graph compile {best[1]*1e3:8.2f} ms (C++)")
