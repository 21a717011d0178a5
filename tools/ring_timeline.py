#!/usr/bin/env python3
"""Per-tile timeline of the compress ring kernel (development aid): where does a tile spend its life,
how long do the consumer warps wait for prefixes, how late are the predecessors' aggregates?
Usage: python tools/ring_timeline.py [log2 n] ; """
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
lib = importlib.import_module("hephaestus-jit_b200._lib").lib

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
TILE, AHEAD, W = 57344, 3, 10
n_tiles = (n + TILE - 1) // TILE
torch.cuda.set_device(0)
dev = hj.Device.cuda(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); dev.set_stream(side.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
lib.hj_debug_compress_trace.argtypes = [ctypes.c_void_p]
g = torch.Generator(device="cuda").manual_seed(0)
for p in (0.5, 0.01, 0.99):
    m = (torch.rand(n, device="cuda", generator=g) < p).to(torch.uint8)
    idx = torch.zeros(n, device="cuda", dtype=torch.int32); cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
    for _ in range(3): dev.compress(n, wrap(cnt), wrap(m), wrap(idx))
    tr = torch.zeros(n_tiles * W, device="cuda", dtype=torch.int64)
    lib.hj_debug_compress_trace(ctypes.c_void_p(tr.data_ptr()))
    for _ in range(2): dev.compress(n, wrap(cnt), wrap(m), wrap(idx))   # first traced launch warms the instantiation
    torch.cuda.synchronize()
    lib.hj_debug_compress_trace(None)
    t = tr.cpu().numpy().reshape(n_tiles, W).astype(np.int64)
    t0 = t[:, 0].min()
    s = [t[:, k] - t0 for k in range(8)]
    cta = t[:, 8]
    print(f"=== compress p={p} n=2^{int(np.log2(n))}: tiles {n_tiles}, span {s[7].max() / 1e3:.1f} us (traced kernel)")
    def row(name, a):
        print(f"  {name:46s} mean {a.mean():8.0f} ns  p50 {np.median(a):8.0f}  p90 {np.percentile(a, 90):8.0f}  max {a.max():8.0f}")
    row("draw -> phase 1 starts (load + ring queue)", s[1] - s[0])
    row("phase 1 (warp 0)", s[2] - s[1])
    row("phase 1 done -> aggregate published", s[3] - s[2])
    row("published -> sweep starts (prefix warp busy)", s[4] - s[3])
    row("sweep (status words -> prefix handed over)", s[5] - s[4])
    row("prefix handed over -> phase 2 starts", s[6] - s[5])
    row("phase 2 (warp 0)", s[7] - s[6])
    row("lifetime draw -> phase 2 done", s[7] - s[0])
    # lateness of predecessors: when was the last aggregate of the tile's sweep range published,
    # relative to the tile's own publication (G = 160 tickets per round)
    G = 160
    late = np.empty(n_tiles)
    for k in range(0, n_tiles, G):
        blk = s[3][k:k + G]
        run = np.maximum.accumulate(blk)
        late[k:k + G] = np.concatenate(([0], run[:-1] - blk[1:]))
    row("last predecessor published - own published", late)
    # per-CTA sequences: consumer wait and prefix slack
    waits, slack, period = [], [], []
    for c in np.unique(cta):
        ids = np.nonzero(cta == c)[0]
        ids = ids[np.argsort(s[0][ids])]
        if len(ids) <= AHEAD + 2: continue
        p1done = s[2][ids]; pref = s[5][ids]; p2start = s[6][ids]; p2end = s[7][ids]
        waits.append(p2start[:-AHEAD] - p1done[AHEAD:])      # after phase 1 of tile it: wait for pref(it - AHEAD)
        slack.append(p1done[AHEAD:] - pref[:-AHEAD])         # > 0: the prefix was there before it was needed
        period.append(np.diff(p2end))
    waits, slack, period = map(np.concatenate, (waits, slack, period))
    row("consumer warp 0: wait for the prefix", waits)
    row("prefix ready before needed (slack, <0 = late)", slack)
    row("tile period per CTA", period)
    tiles_per_cta = np.bincount(cta.astype(np.int64))
    print(f"  tiles per CTA: min {tiles_per_cta.min()} max {tiles_per_cta.max()};  sum of consumer waits per CTA: "
          f"{waits.sum() / len(np.unique(cta)) / 1e3:.1f} us of {s[7].max() / 1e3:.1f} us")
    # fill and drain: the k-th tile of every CTA from the start, and from the end (times relative to the kernel's
    # first draw / to the last phase-2-done of the whole kernel; means over CTAs)
    end = s[7].max()
    seqs = []
    for c in np.unique(cta):
        ids = np.nonzero(cta == c)[0]
        seqs.append(ids[np.argsort(s[0][ids])])
    names = ("draw", "ph1 start", "ph1 done", "published", "sweep start", "prefix", "ph2 start", "ph2 done")
    for label, pick, ref in (("first tiles of a CTA, us after the kernel's first draw", lambda q, k: q[k], 0),
                             ("last tiles of a CTA, us before the kernel's last phase-2-done", lambda q, k: q[-1 - k], end)):
        print(f"  -- {label}")
        for k in range(5):
            ids = np.array([pick(q, k) for q in seqs if len(q) > k])
            cells = "  ".join(f"{nm} {abs(s[j][ids] - ref).mean() / 1e3:6.2f}" for j, nm in enumerate(names))
            print(f"     tile {'+' if ref == 0 else '-'}{k}: {cells}")
            if k == 0:
                cells = "  ".join(f"{nm} {abs(s[j][ids] - ref).max() / 1e3:6.2f}" for j, nm in enumerate(names))
                print(f"       (max): {cells}")
    last_done = np.array([s[7][q].max() for q in seqs])
    print(f"  CTA finish times before the kernel end: mean {(end - last_done).mean() / 1e3:.2f} us, max {(end - last_done).max() / 1e3:.2f} us")
    del m, idx, tr
