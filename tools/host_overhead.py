#!/usr/bin/env python3
"""Host cost of one step (the 4-pass list of bench.py) — tiny arrays, so the GPU is never the limit:
hj_execute_graph vs hj_execute_graph_sharded (world-1 communicator) vs the captured-graph relaunch."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200"); irm = importlib.import_module("hephaestus-jit_b200.ir")
L = importlib.import_module("hephaestus-jit_b200._lib"); sh = importlib.import_module("hephaestus-jit_b200.sharded")
torch.cuda.set_device(0); dev = hj.Device.cuda(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
x = torch.rand(n, device="cuda"); y = torch.empty_like(x); f = torch.rand(n, device="cuda"); s = torch.zeros(4, device="cuda")
u = torch.randint(0, 4, (n,), device="cuda", dtype=torch.int32); scan = torch.empty_like(u); seed = torch.zeros(4, device="cuda", dtype=torch.int32)
mask = (torch.rand(n, device="cuda") < 0.5).to(torch.uint8); index = torch.zeros(n, device="cuda", dtype=torch.int32); count = torch.zeros(4, device="cuda", dtype=torch.int32)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
env = [wrap(t) for t in (x, y, f, s, u, scan, mask, index, count)]
passes = [{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": irm.c2_chain_ir(), "size": n},
          {"kind": hj.PASS_REDUCE, "arg": hj.SUM, "resources": [3, 2]},
          {"kind": hj.PASS_PREFIX_SUM, "arg": 1, "resources": [5, 4]},
          {"kind": hj.PASS_COMPRESS, "resources": [7, 8, 6]}]
descs = [(n, hj.F32, 4), (n, hj.F32, 4), (n, hj.F32, 4), (1, hj.F32, 4), (n, hj.U32, 4), (n, hj.U32, 4), (n, hj.BOOL, 1), (n, hj.U32, 4), (1, hj.U32, 4)]
S, R = L.RES_SHARDED, L.RES_REPLICATED
comm = sh.Comm.local(dev, 0, 1, lambda h: [h])
plain = hj.PreparedGraph(dev, passes, env, descs)
shard = hj.PreparedGraph(dev, passes, env, descs, comm, [S, S, S, R, S, S, S, S, R], [None] * 5 + [wrap(seed)] + [None] * 3)
def bench(name, fn, reps=2000):
    for _ in range(20): fn()
    dev.sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    t_host = time.perf_counter() - t0
    dev.sync()
    t_all = time.perf_counter() - t0
    print(f"{name:46s} host enqueue {t_host / reps * 1e6:7.1f} us/step   incl. GPU drain {t_all / reps * 1e6:7.1f} us/step", flush=True)
print(f"n = {n}, PDL {'off' if os.environ.get('HJ_NO_PDL') else 'on'}")
bench("hj_execute_graph (PreparedGraph.run)", plain.run)
bench("hj_execute_graph_sharded, world 1", shard.run)
c_p, npass, c_env, c_desc, keep = hj.marshal_graph(passes, env, descs)
how = __import__("ctypes").c_uint32()
bench("hj_execute_graph_cached (CUDA graph replay)", lambda: L.lib.hj_execute_graph_cached(dev.handle, 77, c_p, npass, c_env, c_desc, len(env), None))
for name, fn in (("reduce alone", lambda: dev.reduce(hj.SUM, hj.F32, n, env[2], env[3])), ("scan alone", lambda: dev.prefix_sum(hj.U32, n, True, env[4], env[5])),
                 ("compress alone", lambda: dev.compress(n, env[8], env[6], env[7])), ("kernel alone", lambda k=dev.kernel(irm.c2_chain_ir()): dev.launch(k, n, [env[0], env[1]]))):
    bench(name, fn)
comm.destroy()
