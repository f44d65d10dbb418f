#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 pairwise string-similarity hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): char-trigram TF-IDF top-10, company-names self-match,
n = 100 000 -- on SYNTHETIC names calibrated to the reference dataset (polyfuzz_b200/synth.py; the
GPU box has no copy of /root/reference/data).  One step = one pass of the whole hot path:
vectorise (K1) -> inverted index -> sparse cosine + top-10 (K2) [-> all-gather + merge for N > 1].

N > 1 (torchrun, one rank per GPU, NCCL): weak scaling.  The to_list grows to N x 100 000 names and is
row-sharded (rank r owns block r); the from_list stays the first block (100 000 names), scored
against all shards with the global diagonal excluded -- i.e. one from-row-block of the N*100k
self-match.  Per-GPU work is fixed; one all-reduce (df) + one all-gather (top-k) per step.

Prints ONE JSON line on rank 0 (see the repository README / DESIGN.md for the keys).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scored string pairs/sec (TF-IDF char-trigram top-10, top-k index bit-exact vs CPU ref)"
UNIT = "pairs/s"
N_PER_SHARD = 100_000
TOP_N = 10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_SHARD, help="rows per shard (default 100000)")
    ap.add_argument("--cpu-sample-rows", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(n, n_gpus, extra=None):
    cfg = {"workload": "TF-IDF char-trigram top-10 self-match, synthetic company names (BASELINE configs[1] stand-in)",
           "n_from": n, "n_to": n * n_gpus, "top_n": TOP_N, "min_similarity": 0.0, "n_gram_range": [3, 3],
           "parallelism": "to_list row-sharded x%d, all-reduce(df) + all-gather(top-k)" % n_gpus if n_gpus > 1 else "single GPU",
           "l2": "flushed between steps (512 MiB write)"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference path (sklearn TfidfVectorizer with the restated
# analyzer + C restatement of awesome_cossim_topn + the reference's assembly tail), timed on a
# bounded sample and extrapolated to the whole job.
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(names, sample_rows, threads):
    from oracle import native, tfidf
    from oracle.assemble import assemble
    n = len(names)
    t0 = time.perf_counter()
    a, _, _ = tfidf.fit_transform_sklearn(names)                 # whole list (fit + transform), 1 thread
    t1 = time.perf_counter()
    inv = native.InvertedIndex(a)                                # to_vector.T as awesome_cossim_topn takes it
    t2 = time.perf_counter()
    s = min(sample_rows, n)
    idx, val = native.spdot_topn(a[:s], inv, TOP_N, 0.0, self_match=True, n_threads=threads)
    t3 = time.perf_counter()
    assemble(names[:s], names, idx, val)
    t4 = time.perf_counter()
    t_vec, t_inv, t_cos, t_asm = t1 - t0, t2 - t1, t3 - t2, t4 - t3
    whole = t_vec + t_inv + (t_cos + t_asm) * (n / s)
    pairs = float(n) * n - n
    return {"pairs_per_s": pairs / whole, "t_step_measured_s": t4 - t0, "t_whole_job_extrapolated_s": whole,
            "t_vectorise_s": t_vec, "t_transpose_s": t_inv, "t_cos_sample_s": t_cos, "t_assemble_sample_s": t_asm,
            "sample_rows": s}


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import native
    native.build()
    from polyfuzz_b200 import synth
    names = synth.company_names(args.n, seed=0)
    thr = cpu_threads()
    res = []
    for it in range(args.warmup + args.steps):
        r = cpu_reference_step(names, args.cpu_sample_rows, thr)
        if it >= args.warmup:
            res.append(r)
    v = float(np.mean([r["pairs_per_s"] for r in res]))
    ms = float(np.mean([r["t_step_measured_s"] for r in res])) * 1e3
    sample = ("per step: scikit-learn TfidfVectorizer (reference analyzer restated) on all %d names, 1 thread; "
              "C restatement of awesome_cossim_topn top-10 + reference assembly on the first %d from-rows x %d to-rows, "
              "%d OpenMP threads; value = n(n-1) / (t_vec + t_transpose + (t_cos + t_asm) * n / sample)"
              % (args.n, res[0]["sample_rows"], args.n, thr))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args.n, args.gpus, {"reference_block": "100k x 100k block (CPU throughput is per pair)"}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample,
                             "detail": {k: float(np.mean([r[k] for r in res])) for k in res[0] if k != "sample_rows"}},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  Sampled in-process through NVML
    (nvidia_ml_py) every 50 ms: an external `nvidia-smi -lms` loop takes the driver lock for milliseconds per query and
    showed up as +4 ms outliers in 14 ms steps.  Falls back to nvidia-smi if NVML is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.samples, self.maxs, self.reasons = [], [], set()
        self._stop = threading.Event()
        self.thread = None
        self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None
            self._start_smi(gpu_index)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for name, bit in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def _start_smi(self, gpu_index):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "250"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml" if self.nv else "nvidia-smi"}
        if self.nv is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                           samples=len(self.samples))
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import polyfuzz_b200
    from polyfuzz_b200 import _lib, engine, synth
    from polyfuzz_b200.distributed import get_comm, tfidf_topk_sharded
    comm = get_comm()
    n = args.n

    # ---- data: rank r owns to-block r (seed r); the from-block is block 0 -------------------------
    from_list = synth.company_names(n, seed=0)
    shard = from_list if rank == 0 else synth.company_names(n, seed=rank)
    full_to = None                                           # for the public-API e2e leg at N > 1
    if world > 1:
        full_to = from_list + [s for r in range(1, world) for s in synth.company_names(n, seed=r)]

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident inputs (value) -----------------------------------------------------
    vec0 = engine.NgramTfidf((3, 3), True, True)
    staged_from = vec0.stage(from_list)
    staged_to = staged_from if rank == 0 and world == 1 else vec0.stage(shard)
    if world > 1 and rank == 0:
        staged_to = vec0.stage(shard)                        # symmetric work on every rank
    k2_events = []
    result = {}

    def device_step(record_k2):
        vec = engine.NgramTfidf((3, 3), True, True)
        idx, val, csr_to, index = tfidf_topk_sharded(vec, staged_from, staged_to, rank * n, TOP_N, 0.0, self_match=True,
                                                     from_index_base=0, fit=True, fit_on_from=False, comm=comm,
                                                     timings=k2_events if record_k2 else None)
        result["idx"], result["val"], result["vec"], result["csr"], result["index"] = idx, val, vec, csr_to, index

    import gc
    for _ in range(args.warmup):
        flush.zero_(); device_step(False)
    barrier()
    gc.collect(); gc.disable()                               # no collector pauses inside the timed steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.launch_count()
    step_ms = []
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); device_step(True); e1.record()
        barrier()
        step_ms.append(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    total_ms = torch.tensor([float(np.sum(step_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    k2_ms = [a.elapsed_time(b) for a, b in k2_events]

    keep = {k: result[k] for k in ("vec", "csr", "index")}
    result.clear(); result.update(keep)
    gc.enable(); gc.collect(); gc.disable()
    # ---- leg 2: end to end through the public matcher API with HOST lists (e2e) -------------------
    def e2e_step():
        m = polyfuzz_b200.TFIDF(n_gram_range=(3, 3), min_similarity=0.0, top_n=TOP_N, distributed=world > 1)
        if world > 1:
            # block 0 of the N*100k self-match: from-rows [0, n) against the full sharded list,
            # global diagonal excluded
            return m.match(full_to, from_block=(0, n))
        return m.match(from_list)

    e2e_ms = []
    e2e_steps = max(3, args.steps)
    for it in range(2 + e2e_steps):
        flush.zero_()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); df = e2e_step(); e1.record()
        barrier()
        if it >= 2:
            e2e_ms.append(e0.elapsed_time(e1))
    e2e_total = torch.tensor([float(np.sum(e2e_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(e2e_total.item()) / len(e2e_ms)
    gc.enable()
    del df
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- numbers ------------------------------------------------------------------------------------
    pairs = float(n) * (n * world) - n                       # from-block x all to-rows minus the diagonal
    ms_per_step = total_ms / args.steps
    value = pairs / (ms_per_step * 1e-3)
    vec, csr = result["vec"], result["csr"]
    nnz = int(csr.indptr[-1].item())
    # algorithmic bytes of the dominant kernel: P*(4+8) + nnz_from*(4+8) + n_from*k*(4+8),
    # P = postings visited = sum_t df_from(t) * df_to_shard(t)
    cols = csr.indices[:nnz].cpu().numpy()
    df_to = np.bincount(cols, minlength=vec.n_vocab).astype(np.float64)
    if world == 1:
        df_from = df_to; nnz_from = nnz
    else:
        f_csr = vec.emit(vec.rows(staged_from)); nf = int(f_csr.indptr[-1].item())
        df_from = np.bincount(f_csr.indices[:nf].cpu().numpy(), minlength=vec.n_vocab).astype(np.float64); nnz_from = nf
    P = float((df_from * df_to).sum())
    b_alg = P * 12 + nnz_from * 12 + n * TOP_N * 12
    k2_avg_ms = float(np.mean(k2_ms))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k2_ncu_summary.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    achieved = b_alg / (k2_avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "pfz::spcos_dense_kernel (K2, variant %s)" % result["index"].variant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_alg, "postings_per_launch": P, "kernel_ms_avg": k2_avg_ms,
                "kernel_share_of_step": k2_avg_ms / ms_per_step,
                "note": "index (~16 MB) is L2-resident: DRAM traffic << algorithmic bytes by design (SURVEY 8d)"}

    h2d = staged_from.h2d_bytes * (1 if world == 1 else 2) + vec.n_vocab * 8
    d2h = n * TOP_N * 12 + vec.n_vocab * 12 + 4
    e2e = {"value": pairs / (e2e_ms_per_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_ms_per_step, "steps": len(e2e_ms), "ms_each": [round(x, 2) for x in e2e_ms],
           "what": "TFIDF.match(list[str]) -> pandas.DataFrame: UTF-32 packing, H2D, K1, index, K2, D2H, frame assembly"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import native
        native.build()
        thr = cpu_threads()
        r = cpu_reference_step(from_list, args.cpu_sample_rows, thr)
        cpu = {"value": r["pairs_per_s"], "unit": UNIT, "cores": thr, "kind": "port",
               "sample": "sklearn TfidfVectorizer (1 thread) on all %d names + C restatement of awesome_cossim_topn top-10 and the "
                         "reference assembly tail on the first %d from-rows x %d to-rows with %d OpenMP threads, extrapolated to the "
                         "whole job" % (n, r["sample_rows"], n, thr),
               "detail": {k: v for k, v in r.items() if k != "pairs_per_s"}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(n, world, {"tile": result["index"].tile, "k2_variant": result["index"].variant, "V": vec.n_vocab, "nnz": nnz}),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "step_ms_each": [round(x, 3) for x in step_ms]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
