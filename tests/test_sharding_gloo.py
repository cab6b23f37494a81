"""CPU-only, world_size 2 over gloo: the multi-GPU plan of bench.py / the CLI --
contiguous shards of the read batch per rank, no data-path collective, one
all-reduce(sum) of the mapping statistics -- gives the single-process answer."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def shard_bounds(n, world, rank):
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def stats_vector(res, b1, b2):
    """The 18 integers of paired_end_mapping_statistics (abismal.cpp:1034-1057) that
    do not need CIGAR text: pairs {total, unique, ambig, skipped, edits} + per-end totals."""
    pe_valid = res.pe_r1["pos"] != 0
    pe_ambig = (res.pe_r1["flags"] & 0x100) != 0
    l1 = np.diff(b1.off.astype(np.int64))
    l2 = np.diff(b2.off.astype(np.int64))
    rep = pe_valid & ~pe_ambig
    v = [b1.n, int((pe_valid & ~pe_ambig).sum()), int((pe_valid & pe_ambig).sum()), int(((l1 == 0) | (l2 == 0)).sum()),
         int(res.pe_r1["diffs"][rep].astype(np.int64).sum() + res.pe_r2["diffs"][rep].astype(np.int64).sum())]
    for se, ln in ((res.se1, l1), (res.se2, l2)):
        sv = (se["pos"] != 0) & ~rep
        sa = (se["flags"] & 0x100) != 0
        v += [int((~rep).sum()), int((sv & ~sa).sum()), int((sv & sa).sum()), int(((ln == 0) & ~rep).sum()),
              int(se["diffs"][sv & ~sa].astype(np.int64).sum())]
    return np.array(v, np.int64)


def _worker(rank, world, port, idx_path, fq1, fq2, out_dir):
    sys.path.insert(0, helpers.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from abismal_b200 import IndexFile, load_fastq
    ixf = IndexFile(idx_path)
    b1, b2 = load_fastq(fq1, 600), load_fastq(fq2, 600)
    lo, hi = shard_bounds(b1.n, world, rank)
    o = helpers.OracleMapper(ixf, mode=1)
    res = o.map_batch(b1.slice(lo, hi), b2.slice(lo, hi))
    vec = torch.from_numpy(stats_vector(res, b1.slice(lo, hi), b2.slice(lo, hi)))
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, "pos_%d.npy" % rank), res.pe_r1["pos"])
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), vec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(workspace, tmp_path):
    workspace.need_trex()
    from abismal_b200 import IndexFile, load_fastq
    idx, fq1, fq2 = workspace.path("tRex1.idx"), workspace.path("reads_pe_1.fq"), workspace.path("reads_pe_2.fq")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, idx, fq1, fq2, str(tmp_path)), nprocs=2, join=True)
    ixf = IndexFile(idx)
    b1, b2 = load_fastq(fq1, 600), load_fastq(fq2, 600)
    o = helpers.OracleMapper(ixf, mode=1)
    res = o.map_batch(b1, b2)
    want = stats_vector(res, b1, b2)
    got = np.load(str(tmp_path / "stats.npy"))
    assert np.array_equal(got, want)
    pos = np.concatenate([np.load(str(tmp_path / ("pos_%d.npy" % r))) for r in range(2)])
    assert np.array_equal(pos, res.pe_r1["pos"])  # shards concatenate to the input order


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(n, world, r)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _ranksync_worker(rank, world, port, out_dir, failing_rank):
    """bench.py's RankSync on gloo: a phase that raises on one rank must end every rank (exit status 1), with the
    traceback in the failing rank's log and in the JSON error line rank 0 prints -- nobody waits in a barrier."""
    import io
    import json
    sys.path.insert(0, helpers.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    import bench
    importlib.reload(bench)  # RANK is read at import
    sync = bench.RankSync(dist, True, world)
    assert sync.run("phase one", lambda: 40 + rank) == 40 + rank  # a phase that succeeds everywhere goes on
    out = io.StringIO()
    real_stdout, sys.stdout = sys.stdout, out
    code = None
    try:
        def phase_two():
            if rank == failing_rank:
                raise RuntimeError("boom on rank %d" % rank)
            return "fine"
        sync.run("phase two", phase_two)
    except SystemExit as e:
        code = e.code
    finally:
        sys.stdout = real_stdout
    with open(os.path.join(out_dir, "ranksync_%d.json" % rank), "w") as f:
        json.dump({"code": code, "stdout": out.getvalue(), "error": sync.error}, f)


def test_bench_rank_failure_ends_all_ranks_with_the_traceback(tmp_path):
    import json
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_ranksync_worker, args=(2, port, str(tmp_path), 1), nprocs=2, join=True)
    r0 = json.load(open(str(tmp_path / "ranksync_0.json")))
    r1 = json.load(open(str(tmp_path / "ranksync_1.json")))
    assert r0["code"] == 1 and r1["code"] == 1            # both ranks leave through sys.exit(1)
    assert r0["error"] is None and "boom on rank 1" in r1["error"]
    line = json.loads(r0["stdout"].strip().splitlines()[-1])
    assert line["error"] == "bench failed" and "boom on rank 1" in line["ranks"]["1"] and "0" not in line["ranks"]
    assert r1["stdout"].strip() == ""                     # only rank 0 prints
