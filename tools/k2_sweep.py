"""Developer tool: time K1 stages and K2 over tile sizes on the company-names workload (real fixture when present)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
from polyfuzz_b200 import datasets, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
tiles = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [512, 1024, 1536, 2048, 2560, 2816]
names, _kind = datasets.load_company_names(n)
torch.cuda.init()


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, float(np.median(ts)), float(np.min(ts))


t0 = time.time(); v = engine.NgramTfidf((3, 3), True, True)
(rows,), t_fit, _ = timed(lambda: v.fit_rows([names]), 3, 1)
csr, t_emit, _ = timed(lambda: v.emit(rows), 3, 1)
torch.cuda.synchronize()
ip = csr.indptr.cpu().numpy(); nnz = int(ip[-1])
df = np.bincount(csr.indices[:nnz].cpu().numpy(), minlength=v.n_vocab).astype(np.float64)
P = float((df * df).sum())
print(f"n={n} V={v.n_vocab} nnz={nnz} P={P:.4g} fit(stageA+vocab incl. H2D, host idf)={t_fit:.2f} ms emit={t_emit:.2f} ms wall={time.time()-t0:.2f}s")
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["dense32", "dense", "list"]
ref = None
for variant in variants:
  for tile in tiles:
    idx_obj, t_ix, _ = timed(lambda: engine.SparseIndex(csr, tile=tile, variant=variant), 3, 1)
    (oi, ov), t_k2, t_min = timed(lambda: engine.spcos_topk(csr, idx_obj, 10, 0.0, self_match=True, n_splits=1, variant=variant), 5, 2)
    if ref is None:
        ref = (oi.clone(), ov.clone())
    same = bool(torch.equal(oi, ref[0]) and torch.equal(ov, ref[1]))
    pairs = float(n) * n - n
    print(f"{variant:5s} same={same} tile={tile:5d} n_tiles={idx_obj.n_tiles:4d} index_build={t_ix:7.2f} ms  K2 median={t_k2:8.2f} ms min={t_min:8.2f} ms  "
          f"pairs/s={pairs / (t_k2 * 1e-3):.3e}  postings/s={P / (t_k2 * 1e-3):.3e}  B_alg GB/s={(P * 12 + nnz * 12 + n * 120) / (t_k2 * 1e-3) / 1e9:.1f}")
