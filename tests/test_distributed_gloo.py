"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard bounds, the single all-gather of
packed top-k lists, the df all-reduce, and that merge(shard results) == unsharded oracle result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from polyfuzz_b200.distributed import Comm, get_comm, shard_bounds
        from polyfuzz_b200 import synth
        from oracle import native, tfidf
        comm = get_comm()
        assert isinstance(comm, Comm) and comm.world_size == world and comm.rank == rank
        to = synth.company_names(600, seed=1); frm = synth.company_names(150, seed=2)
        f, t, _ = tfidf.fit_transform_sklearn(frm, to)
        lo, hi = shard_bounds(len(to), world, rank)
        # df all-reduce: per-shard document frequencies sum to the global ones
        df_local = torch.from_numpy(np.bincount(t[lo:hi].indices, minlength=t.shape[1]).astype(np.int32))
        comm.all_reduce_sum(df_local)
        assert np.array_equal(df_local.numpy(), np.bincount(t.indices, minlength=t.shape[1]))
        # local top-k with global indices, one all-gather, canonical merge
        li, lv = native.spdot_topn(f, t[lo:hi], 5, 0.0, to_index_base=lo)
        gi, gv = comm.all_gather_topk(torch.from_numpy(li), torch.from_numpy(lv))
        assert gi.shape == (world, 150, 5) and gi.dtype == torch.int32 and gv.dtype == torch.float64
        mi, mv = native.topk_merge(gi.numpy(), gv.numpy(), 5)
        fi, fv = native.spdot_topn(f, t, 5, 0.0)
        assert np.array_equal(mi, fi) and np.array_equal(mv, fv)
        # top_n > 32: the sort-based canonical merge (same key) against the oracle's merge
        from polyfuzz_b200.distributed import merge_topk_any
        li, lv = native.spdot_topn(f, t[lo:hi], 40, 0.0, to_index_base=lo)
        gi, gv = comm.all_gather_topk(torch.from_numpy(li), torch.from_numpy(lv))
        mi2, mv2 = merge_topk_any(gi, gv, 40)
        fi, fv = native.spdot_topn(f, t, 40, 0.0)
        assert np.array_equal(mi2.numpy(), fi) and np.array_equal(mv2.numpy(), fv)
        # edit-distance exchange: per-shard best (global index, score, distance) -> one all-gather -> first maximum
        bi, bs, bd = native.editdist_argbest(frm[:40], to[lo:hi], "ratio")
        bi = np.where(bi >= 0, bi + lo, bi).astype(np.int32)
        gi, gs, gd = comm.all_gather_best(torch.from_numpy(bi), torch.from_numpy(bs), torch.from_numpy(bd))
        assert gi.shape == (world, 40) and gi.dtype == torch.int32 and gs.dtype == torch.float64 and gd.dtype == torch.int32
        ref_i, ref_s, ref_d = native.editdist_argbest(frm[:40], to, "ratio")
        best = np.full(40, -1, np.int32); bsc = np.zeros(40); bdd = np.full(40, -1, np.int32)
        for r in range(world):
            for q in range(40):
                j, sc = int(gi[r, q]), float(gs[r, q])
                if j >= 0 and (best[q] < 0 or sc > bsc[q] or (sc == bsc[q] and j < best[q])):
                    best[q], bsc[q], bdd[q] = j, sc, int(gd[r, q])
        assert np.array_equal(best, ref_i) and np.array_equal(bsc, ref_s) and np.array_equal(bdd, ref_d)
        # frame-tail exchange: the ranks' string shards (bytes + offsets) -> the whole list on every rank
        from types import SimpleNamespace
        from polyfuzz_b200.distributed import gather_string_shards
        mine = to[lo:hi]
        blob = np.frombuffer("".join(mine).encode("ascii"), dtype=np.uint8).astype(np.int32)
        offs = np.zeros(len(mine) + 1, dtype=np.int64); np.cumsum([len(x) for x in mine], out=offs[1:])
        S = SimpleNamespace(n=len(mine), n_chars=int(blob.size), ascii=True, d_blob=torch.from_numpy(blob), d_off=torch.from_numpy(offs))
        gb, go, ok_ascii, _ = gather_string_shards(comm, S)
        assert ok_ascii and go.numel() == len(to) + 1
        whole = bytes(gb.to(torch.uint8).numpy()).decode("ascii")
        assert [whole[int(go[i]):int(go[i + 1])] for i in range(len(to))] == to
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from polyfuzz_b200.distributed import shard_bounds
    for n in (0, 1, 7, 100, 101):
        for g in (1, 2, 3, 8):
            cover = []
            for r in range(g):
                lo, hi = shard_bounds(n, g, r)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi))
            assert cover == list(range(n))


def test_pack_unpack():
    from polyfuzz_b200.distributed import pack_topk, unpack_topk
    idx = torch.tensor([[3, -1], [0, 7]], dtype=torch.int32)
    val = torch.tensor([[0.5, 0.0], [1.0, 1e-300]], dtype=torch.float64)
    i2, v2 = unpack_topk(pack_topk(idx, val))
    assert torch.equal(i2, idx) and torch.equal(v2, val)


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
