#!/usr/bin/env python3
"""Session-4 kernels under compute-sanitizer: histogram paths (dummy counters, bin copies, windows,
keys out of range), the traced compress with the index zero-fill left to the Compress pass, graph round trip."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200"); tr = importlib.import_module("hephaestus-jit_b200.tr")
dev = hj.Device.cuda(0)
rng = np.random.Generator(np.random.PCG64(0))
n = (1 << 20) + 4097
for nb, lit, oob in ((1 << 16, 1, False), (40001, 1, True), (1 << 16, 2, False), (1000, 1, True), (1024, 3, False),
                     (3000, 1, True), (16383, 1, False), (100000, 1, True)):
    keys = rng.integers(0, 2 * nb if oob else nb, size=n).astype(np.uint32)
    keys[: n // 3] = 7
    if oob:
        keys[::97] = 0xFFFFFFFF
    bh = dev.create_buffer_from_slice(np.zeros(nb, np.uint32))
    dev.scatter_reduce(hj.SUM, hj.U32, n, dev.create_buffer_from_slice(keys), None, lit, bh, nb)
    want = (np.bincount(keys[keys < nb], minlength=nb) * lit).astype(np.uint32)
    assert np.array_equal(bh.to_host(np.uint32), want), (nb, lit, oob)
for m in (5, 70_001, n):
    vals = rng.random(m, dtype=np.float32)
    x = tr.array(vals, dev)
    for evaluated in (False, True):
        mask = x.lt(tr.literal(0.3, hj.F32))
        if evaluated:
            mask.schedule(); tr.compile().launch(dev)
        count, index = mask.compress()
        g = tr.compile()
        g2 = tr.Graph.deserialize(g.serialize(), dev)
        for graph in (g, g2):
            graph.launch(dev)
            want = np.zeros(m, np.uint32); sel = np.nonzero(vals < np.float32(0.3))[0]; want[: sel.size] = sel
            assert int(count.to_vec(np.uint32)[0]) == sel.size and np.array_equal(index.to_vec(np.uint32), want)
        del count, index, mask, g, g2
    del x
dev.sync()
print("sanitize workload ok")
