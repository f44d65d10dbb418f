"""BASELINE config 5: TF-IDF top-10 on n x n synthetic 8..32-char strings, to_list row-sharded across the GPUs of one
node (torchrun).  Prints the end-to-end match_arrays time (host packing + H2D + K1 + index + K2 + all-gather + merge)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import polyfuzz_b200
from polyfuzz_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
to = synth.uniform_strings(n, seed=0); frm = synth.uniform_strings(n, seed=1)
m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=10, distributed=world > 1)
ts = []
for rep in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, val, k = m.match_arrays(frm, to)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
chk = int(idx.to(torch.int64).sum().item()); vs = float(val.sum().item())
tmax = torch.tensor([min(ts[1:])], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
print(f"rank {rank}: idx checksum {chk} score sum {vs:.12f}", flush=True)
if rank == 0:
    t = float(tmax.item())
    print(f"C5 n={n} x {n} on {world} GPU(s): variant={m._index.variant} tile={m._index.tile} V={m.vectorizer.n_vocab} "
          f"match_arrays {t*1e3:.1f} ms  pairs/s={float(n)*n/t:.3e}", flush=True)
if world > 1: dist.destroy_process_group()
