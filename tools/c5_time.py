"""Developer tool: uniform-text workload (BASELINE config 5 shape) on one GPU: n x n, list variant."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import polyfuzz_b200
from polyfuzz_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
to = synth.uniform_strings(n, seed=0); frm = synth.uniform_strings(n, seed=1)
m = polyfuzz_b200.TFIDF(min_similarity=0.0, top_n=10)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, val, k = m.match_arrays(frm, to); torch.cuda.synchronize(); t = time.perf_counter() - t0
print(f"uniform {n} x {n}: variant={m._index.variant} tile={m._index.tile} V={m.vectorizer.n_vocab} density={m.vectorizer.density():.4f} "
      f"match_arrays {t*1e3:.1f} ms  pairs/s={n*n/t:.3e}")
