"""Developer tool: per-phase SM cycles of the block kernel (library built with PFZ_NVCC_EXTRA=-DPFZ_B3_TIMING)."""
import sys, os, ctypes
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from polyfuzz_b200 import datasets, engine, _lib
names, _ = datasets.load_company_names(100_000)
v = engine.NgramTfidf((3, 3), True, True)
(rows,) = v.fit_rows([names]); csr = v.emit(rows)
ix = engine.SparseIndex(csr, variant="block")
lib = _lib.load()
out = (ctypes.c_ulonglong * 8)()
for _ in range(2):
    engine.spcos_topk(csr, ix, 10, 0.0, self_match=True, n_splits=1)
lib.pfz_debug_b3_cycles(out, 1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); engine.spcos_topk(csr, ix, 10, 0.0, self_match=True, n_splits=1); e1.record(); torch.cuda.synchronize()
lib.pfz_debug_b3_cycles(out, 1)
c = np.array(list(out), dtype=np.float64)
names_ = ["table", "wait A", "items", "wait B", "scan (+prologue)", "exact at block end", "wait block end", "-"]
print(f"kernel+prep {e0.elapsed_time(e1):.2f} ms; tile {ix.tile} rows {engine.BLOCK_ROWS} bits {ix.acc_bits}")
for n_, x in zip(names_, c):
    print(f"{n_:22s} {100 * x / c.sum():5.1f}%  {x / 1e9:8.2f} G warp-cycles")
