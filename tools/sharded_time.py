"""torchrun target: time the sharded ops (with and without the peer-memory exchange: HJ_NO_P2P=1)."""
import importlib, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
sharded = importlib.import_module("hephaestus-jit_b200.sharded")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
dev = hj.Device.cuda(lr)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); dev.set_stream(st.cuda_stream)
comm = sharded.Comm.from_torch(dev)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(rank)
n = (1 << 30) // world
xu = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
o1 = torch.zeros(16, device="cuda", dtype=torch.int32); on = torch.empty(n, device="cuda", dtype=torch.int32)
bu, bo1, bon = wrap(xu), wrap(o1), wrap(on)
def timed(fn, iters=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev: a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[iters // 2]
    t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
r = {}
HIST_ONLY = bool(os.environ.get("HIST_ONLY"))
if HIST_ONLY:
    ok = True
if not HIST_ONLY:
  r["reduce_sum_u32_ms"] = timed(lambda: comm.reduce(hj.SUM, hj.U32, n, bu, bo1))
if not HIST_ONLY:
  want = torch.tensor([int(xu.to(torch.int64).sum().item())], device="cuda", dtype=torch.int64); dist.all_reduce(want)
  ok = (int(o1[0].item()) & 0xFFFFFFFF) == (int(want.item()) & 0xFFFFFFFF)
  r["local_reduce_ms"] = timed(lambda: dev.reduce(hj.SUM, hj.U32, n, bu, bo1))
  r["scan_ms"] = timed(lambda: comm.prefix_sum(hj.U32, n, True, bu, bon))
  r["local_scan_ms"] = timed(lambda: dev.prefix_sum(hj.U32, n, True, bu, bon))
# histogram: 2^28 keys over the ranks -> 2^16 bins, private per GPU + all-reduce of the bins
nk = (1 << 28) // world
keys = torch.randint(0, 1 << 16, (nk,), device="cuda", generator=g, dtype=torch.int32)
hist = torch.zeros(1 << 16, device="cuda", dtype=torch.int32)
bk, bh = wrap(keys), wrap(hist)
def hist_step():
    hist.zero_()
    comm.scatter_reduce(hj.SUM, hj.U32, nk, bk, None, 1, bh, 1 << 16)
def hist_local():
    hist.zero_()
    dev.scatter_reduce(hj.SUM, hj.U32, nk, bk, None, 1, bh, 1 << 16)
r["hist_ms"] = timed(hist_step)
hist_step(); torch.cuda.synchronize()
ok = ok and (int(hist.to(torch.int64).sum().item()) == nk * world)
r["local_hist_ms"] = timed(hist_local)
r["bins_allreduce_only_ms"] = timed(lambda: comm.scatter_reduce(hj.SUM, hj.U32, 0, bk, None, 1, bh, 1 << 16))
if rank == 0:
    print(f"world={world} peer_array={'off' if os.environ.get('HJ_NO_PEER_ARRAY') else 'on'} p2p={'off' if os.environ.get('HJ_NO_P2P') else 'on'} ok={ok} " + " ".join(f"{k}={v*1e3:.1f}us" for k, v in r.items()), flush=True)
comm.destroy(); dist.barrier(); dist.destroy_process_group()
