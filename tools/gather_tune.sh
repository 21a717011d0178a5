#!/bin/bash
# Gather with a DRAM-resident table: how the table load is issued decides how many sectors L1 asks L2 for.
python - <<'PY'
import torch
n = (1 << 28)
g = torch.Generator(device="cuda").manual_seed(0)
table = torch.rand(1 << 28, device="cuda", generator=g); idx = torch.randint(0, 1 << 28, (n,), device="cuda", generator=g, dtype=torch.int64)
for _ in range(2): out = table[idx]
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = table[idx]; b.record(); torch.cuda.synchronize()
print(f"torch table[idx] (int64 indices): {a.elapsed_time(b):.3f} ms  {n / a.elapsed_time(b) / 1e6:.1f} Gelem/s")
PY
for cfg in 20 22 23 24 25 26 42 43 82; do
  echo "== HJ_GATHER_CFG=$cfg"
  HJ_GATHER_CFG=$cfg timeout 120 python tools/gather_time.py 2>&1 | grep "gather table"
done
