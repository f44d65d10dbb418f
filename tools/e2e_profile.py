"""Developer tool: wall-clock breakdown of TFIDF.match (host + device) on the 100k synthetic workload."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import polyfuzz_b200
from polyfuzz_b200 import engine, synth
from polyfuzz_b200.matchers._utils import assemble_matches, prepare_strings
names = synth.company_names(100_000, seed=0)
def sync(): torch.cuda.synchronize()
for rep in range(4):
    T = {}
    t0 = time.perf_counter()
    vec = engine.NgramTfidf((3, 3), True, True)
    S = vec.stage(names); sync(); t1 = time.perf_counter(); T["stage(pack+H2D)"] = t1 - t0
    (rows,) = vec.fit_staged([S]); sync(); t2 = time.perf_counter(); T["fit(K1 A/B + sync + host idf)"] = t2 - t1
    csr = vec.emit(rows); sync(); t3 = time.perf_counter(); T["emit"] = t3 - t2
    ix = engine.SparseIndex(csr, variant="dense"); sync(); t4 = time.perf_counter(); T["index"] = t4 - t3
    oi, ov = engine.spcos_topk(csr, ix, 10, 0.0, self_match=True); sync(); t5 = time.perf_counter(); T["K2"] = t5 - t4
    hi, hv = oi.cpu().numpy(), ov.cpu().numpy(); t6 = time.perf_counter(); T["D2H"] = t6 - t5
    prep = prepare_strings(names, None); t7 = time.perf_counter(); T["arrow prep"] = t7 - t6
    df = assemble_matches(names, names, hi, hv, prepared=prep); t8 = time.perf_counter(); T["assemble"] = t8 - t7
    t9 = time.perf_counter(); m = polyfuzz_b200.TFIDF(min_similarity=0, top_n=10); d2 = m.match(names); t10 = time.perf_counter()
    if rep >= 2:
        print(" | ".join(f"{k} {v*1e3:.1f}" for k, v in T.items()), f"| sum {1e3*(t8-t0):.1f} | match() {1e3*(t10-t9):.1f} ms")
