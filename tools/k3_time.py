"""Developer tool: time K3 on the movie-titles-like grid (BASELINE config 3 shape: 6172 x 80852)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from polyfuzz_b200 import editdist, synth
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 6172
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 80852
frm = synth.titles(nf, seed=1); to = synth.titles(nt, seed=2)
fl = np.array([len(s) for s in frm]); tl = np.array([len(s) for s in to])
cells = float(fl.sum()) * float(tl.sum())
print(f"n_from={nf} n_to={nt} mean len from {fl.mean():.1f} to {tl.mean():.1f} max {fl.max()} {tl.max()} cells={cells:.4g}")
for metric in ("norm_lev", "ratio"):
    for _ in range(2):
        editdist.edit_argbest(frm, to, metric)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); editdist.edit_argbest(frm, to, metric); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    print(f"{metric:9s} end-to-end (host packing + H2D + kernels) median {t*1e3:8.2f} ms  pairs/s={nf*nt/t:.3e}  GCUPS={cells/t/1e9:.1f}")
