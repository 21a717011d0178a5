#!/usr/bin/env python3
"""What the host link of this box sustains: pinned H2D alone, D2H alone, both at once (1 GiB each,
CUDA events, best of 5) — the ceiling of bench.py's e2e number (8 bytes cross the link per element)."""
import torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, chunks=1):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        step = n // chunks
        for c in range(chunks):
            sl = slice(c * step, (c + 1) * step)
            if up:
                with torch.cuda.stream(s1): d_in[sl].copy_(h_in[sl], non_blocking=True)
            if down:
                with torch.cuda.stream(s2): h_out[sl].copy_(d_out[sl], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for name, up, down in (("H2D alone", True, False), ("D2H alone", False, True), ("both at once", True, True)):
    for chunks in (1, 32):
        ms = run(up, down, chunks)
        total = n * (up + down)
        print(f"{name:14s} chunks={chunks:3d}: {ms:7.2f} ms  {total / ms / 1e6:6.1f} GB/s total", flush=True)
