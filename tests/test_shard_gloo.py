"""N>1 path on CPU: two gloo ranks deal the chunks of one read stream (bwa_mem2_b200.shard.ShardedAligner), each runs the hot path on
its own chunks with no data-path collective (the CPU oracle stands in for the GPU context here; tests/test_shard_gpu.py runs the same
flow on the GPU with the broadcast index), rank 0 gathers in input order; the result must equal the single-process result read for
read.  Also: chunk_bounds cuts the stream where the reference's bseq_read_orig does."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, golden_dir, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load_package
    pkg = load_package()
    import importlib
    shard = importlib.import_module("bwa_mem2_b200.shard")
    import oracle_lib as ol
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa"); opt = pkg.capi.default_opt()
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    n, L = reads.shape
    codes = reads.reshape(-1); offs = (np.arange(n + 1) * L).astype(np.int64)

    def oracle(c, o):
        regs, ro, _, rc = ol.seed_chain_extend(idx, opt, c, o)
        assert rc == 0
        return regs, ro
    sa = shard.ShardedAligner(pkg.capi, golden_dir + "/c0_index/ref.fa", compute=oracle)
    bounds = shard.chunk_bounds((n, L), 40_000, paired=True)           # 133 pairs per chunk: 4 chunks, not aligned to 512-read blocks
    res = sa.align_chunks(codes, offs, bounds)
    table = sa.chunk_table(res)
    assert [t[0] for t in table] == list(range(len(bounds))) and all(t[1] == t[0] % world for t in table)
    regs, ro = sa.gather_in_order(res, dst=0)
    cnt = torch.tensor([sum(len(r[3]) - 1 for r in res)]); dist.all_reduce(cnt)     # the only data-path collective: a counter
    if rank == 0:
        assert int(cnt.item()) == n
        np.save(out + ".off.npy", ro); open(out + ".regs.bin", "wb").write(regs.tobytes())
        np.save(out + ".bounds.npy", np.array(bounds))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(pkg, golden_dir, tmp_path):
    import oracle_lib as ol
    out = str(tmp_path / "gather")
    mp.spawn(_worker, args=(2, 29571, golden_dir, out), nprocs=2, join=True)
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa"); opt = pkg.capi.default_opt()
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    n, L = reads.shape
    # the single-process result, chunk by chunk as the reference would process the stream with the same -K
    bounds = [tuple(b) for b in np.load(out + ".bounds.npy")]
    assert len(bounds) == 4 and bounds[0] == (0, 266)
    parts = []; counts = []
    for s, e in bounds:
        regs, ro, _, rc = ol.seed_chain_extend(idx, opt, reads[s:e].reshape(-1), (np.arange(e - s + 1) * L).astype(np.int64))
        parts.append(regs.tobytes()); counts.append(np.diff(ro))
    assert np.array_equal(np.diff(np.load(out + ".off.npy")), np.concatenate(counts))
    assert open(out + ".regs.bin", "rb").read() == b"".join(parts)


def test_chunk_bounds_follow_the_reference_reader(pkg):
    import importlib
    shard = importlib.import_module("bwa_mem2_b200.shard")
    # uniform 151-bp pairs, -K 10 Mbp: ceil(1e7 / 302) = 33113 pairs per chunk (src/bwa.cpp:170-216)
    b = shard.chunk_bounds((200_000, 151), 10_000_000, paired=True)
    assert b[0] == (0, 66_226) and b[-1][1] == 200_000 and all(s % 2 == 0 for s, _ in b)
    # ragged lengths: a chunk ends at the first even read count whose bases reach the task size
    rng = np.random.default_rng(5)
    lens = rng.integers(30, 300, 5000)
    bb = shard.chunk_bounds(lens, 50_000, paired=True)
    assert bb[0][0] == 0 and bb[-1][1] == 5000 and all(bb[i][1] == bb[i + 1][0] for i in range(len(bb) - 1))
    for s, e in bb[:-1]:
        tot = int(lens[s:e].sum())
        assert (e - s) % 2 == 0 and tot >= 50_000 and int(lens[s:e - 2].sum()) < 50_000
    assert shard.chunk_bounds((0, 151), 1000) == [] and shard.chunk_bounds((10, 100), 250, paired=False) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard.task_size(10_000_000, 8) == 80_000_000 and shard.task_size(10_000_000, 8, fixed_k=123) == 123
    r = shard.rank_chunks(bb, 1, 3)
    assert [i for i, _ in r] == list(range(1, len(bb), 3))
