#!/usr/bin/env python3
"""Development sweep over the scan / compress kernel configurations (HJ_SCAN_CFG / HJ_COMPRESS_CFG),
one subprocess per configuration; checks the result against torch on the device and prints GB/s."""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(which, log2n):
    import torch
    hj = importlib.import_module("hephaestus-jit_b200")
    torch.cuda.set_device(0)
    dev = hj.Device.cuda(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    dev.set_stream(side.cuda_stream)
    wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
    g = torch.Generator(device="cuda").manual_seed(0)
    res = {}

    def timeit(fn, iters=20, warmup=3):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return ts[len(ts) // 2]

    for n in ((1 << log2n), (1 << log2n) - 12345, (1 << 20) + 7, 70000):
        if which == "scan":
            x = torch.randint(0, 1 << 31, (n,), device="cuda", generator=g, dtype=torch.int32)
            y = torch.empty_like(x)
            for incl in (True, False):
                y.zero_()
                dev.prefix_sum(hj.U32, n, incl, wrap(x), wrap(y))
                torch.cuda.synchronize()
                want = torch.cumsum(x.to(torch.int64), 0)
                if not incl:
                    want = want - x
                ok = bool(torch.equal(y.to(torch.int64) & 0xffffffff, want & 0xffffffff))
                res[f"ok n={n} incl={incl}"] = ok
            if n >= 1 << 24:
                ms = timeit(lambda: dev.prefix_sum(hj.U32, n, True, wrap(x), wrap(y)))
                res[f"GB/s n={n}"] = round(8 * n / ms / 1e6, 1)
                xd = torch.rand(n // 2, device="cuda", generator=g, dtype=torch.float64)
                yd = torch.empty_like(xd)
                ms = timeit(lambda: dev.prefix_sum(hj.F64, n // 2, True, wrap(xd), wrap(yd)))
                res[f"f64 GB/s n={n // 2}"] = round(16 * (n // 2) / ms / 1e6, 1)
                res["f64 ok"] = bool(torch.allclose(yd, torch.cumsum(xd, 0), rtol=1e-9))
                del xd, yd
            del x, y
        else:
            for p in (0.5, 0.01, 0.99):
                m = (torch.rand(n, device="cuda", generator=g) < p).to(torch.uint8)
                idx = torch.zeros(n, device="cuda", dtype=torch.int32)
                cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
                dev.compress(n, wrap(cnt), wrap(m), wrap(idx))
                torch.cuda.synchronize()
                want = torch.nonzero(m).flatten().to(torch.int32)
                c = int(cnt.item())
                ok = c == want.numel() and bool(torch.equal(idx[:c], want)) and bool((idx[c:] == 0).all())
                res[f"ok n={n} p={p}"] = ok
                if n >= 1 << 24:
                    ms = timeit(lambda: dev.compress(n, wrap(cnt), wrap(m), wrap(idx)))
                    res[f"GB/s n={n} p={p}"] = round((n + 4 * c) / ms / 1e6, 1)
                del m, idx, want
    print(json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(sys.argv[2], int(sys.argv[3]))
    which = sys.argv[1] if len(sys.argv) > 1 else "scan"
    cfgs = [int(c) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,3,4,5,6,7".split(","))]
    log2n = sys.argv[3] if len(sys.argv) > 3 else "28"
    var = "HJ_SCAN_CFG" if which == "scan" else "HJ_COMPRESS_CFG"
    for c in cfgs:
        env = dict(os.environ)
        env[var] = str(c)
        try:
            r = subprocess.run([sys.executable, __file__, "--child", which, log2n], env=env, capture_output=True,
                               text=True, timeout=240)
            out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("ERR " + r.stderr[-600:])
        except subprocess.TimeoutExpired:
            out = "TIMEOUT (hang?)"
        print(f"{var}={c}: {out}", flush=True)


if __name__ == "__main__":
    main()
