"""Developer tool: time K4 on BASELINE config 4 (100k x 100k x 768 random unit vectors, top-10)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from polyfuzz_b200 import dense
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device("cuda")
torch.manual_seed(0); X = torch.randn(n, d, device=dev); torch.manual_seed(1); Y = torch.randn(n, d, device=dev)
x, _ = dense.to_bf16_rows(X, True); y, _ = dense.to_bf16_rows(Y, True)
for splits in (None, 1, 3):
    for _ in range(2):
        dense.dense_topk(x, y, 10, 0.0, n_splits=splits)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dense.dense_topk(x, y, 10, 0.0, n_splits=splits); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts)) * 1e-3
    print(f"n={n} d={d} splits={splits}: {t*1e3:.2f} ms  pairs/s={n*n/t:.3e}  TFLOP/s={2.0*n*n*d/t/1e12:.1f}")
