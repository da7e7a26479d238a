"""N>1 path on CPU: two gloo ranks shard the chunks of one read set (bwa_mem2_b200.shard), each runs the hot path
(the CPU oracle stands in for the GPU here) on its own chunks with no data-path collective, rank 0 gathers; the
result must equal the single-process result read for read."""
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
    mine = []
    for ci, (s, e) in shard.rank_chunks(n, 512, rank, world):
        codes = reads[s:e].reshape(-1); offs = (np.arange(e - s + 1) * L).astype(np.int64)
        regs, ro, _, rc = ol.seed_chain_extend(idx, opt, codes, offs)
        mine.append((ci, s, regs.tobytes(), ro.tolist()))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    cnt = torch.tensor([sum(len(m[3]) - 1 for m in mine)]); dist.all_reduce(cnt)     # the only collective: a counter
    if rank == 0:
        allc = sorted([c for g in gathered for c in g])
        blob = b"".join(c[2] for c in allc)
        counts = np.concatenate([np.diff(np.array(c[3])) for c in allc])
        np.save(out + ".counts.npy", counts); open(out + ".regs.bin", "wb").write(blob)
        assert int(cnt.item()) == n
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(pkg, golden_dir, tmp_path):
    import oracle_lib as ol
    out = str(tmp_path / "gather")
    mp.spawn(_worker, args=(2, 29571, golden_dir, out), nprocs=2, join=True)
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa"); opt = pkg.capi.default_opt()
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    codes = reads.reshape(-1); offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    regs, ro, _, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    assert np.array_equal(np.load(out + ".counts.npy"), np.diff(ro))
    assert open(out + ".regs.bin", "rb").read() == regs.tobytes()


def test_chunk_ranges_are_block_aligned(pkg):
    import importlib
    shard = importlib.import_module("bwa_mem2_b200.shard")
    r = shard.chunk_ranges(100_000, 33_000)
    assert all(s % 512 == 0 for s, _ in r) and r[-1][1] == 100_000 and r[0] == (0, 33_280)
    a = shard.rank_chunks(100_000, 33_000, 0, 2); b = shard.rank_chunks(100_000, 33_000, 1, 2)
    assert sorted(a + b) == list(enumerate(r))
