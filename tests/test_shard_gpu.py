"""The sharded start-up and dealing on the GPU: two ranks (gloo rendezvous; both on cuda:0 when the box has one GPU), rank 0
reads the index files, rank 1 receives the four big arrays by broadcast and adopts them in place (bm2_create_resident), each aligns its
chunks on the GPU, rank 0 gathers in input order; the result must equal one context aligning the same chunks."""
import os, sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, golden_dir, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load_package
    pkg = load_package()
    import importlib
    shard = importlib.import_module("bwa_mem2_b200.shard")
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    sa = shard.ShardedAligner(pkg.capi, golden_dir + "/c0_index/ref.fa", device=dev)
    assert "broadcast_s" in sa.startup and (rank != 0 or sa.startup["index_load_s"] > 0)
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    n, L = reads.shape
    codes = reads.reshape(-1); offs = (np.arange(n + 1) * L).astype(np.int64)
    bounds = shard.chunk_bounds((n, L), 40_000, paired=True)
    res = sa.align_chunks(codes, offs, bounds)
    regs, ro = sa.gather_in_order(res, dst=0)
    if rank == 0:
        np.save(out + ".off.npy", ro); open(out + ".regs.bin", "wb").write(regs.tobytes()); np.save(out + ".bounds.npy", np.array(bounds))
    sa.close()
    dist.barrier(); dist.destroy_process_group()


def test_two_ranks_with_broadcast_index_equal_one_context(pkg, golden_dir, tmp_path):
    out = str(tmp_path / "gather")
    mp.spawn(_worker, args=(2, 29573, golden_dir, out), nprocs=2, join=True)
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa")
    ctx = pkg.capi.Context(0, index=idx)
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    n, L = reads.shape
    parts = []; counts = []
    for s, e in [tuple(b) for b in np.load(out + ".bounds.npy")]:
        regs, ro = ctx.seed_chain_extend(reads[s:e].reshape(-1), (np.arange(e - s + 1) * L).astype(np.int64))
        parts.append(regs.tobytes()); counts.append(np.diff(ro))
    assert np.array_equal(np.diff(np.load(out + ".off.npy")), np.concatenate(counts))
    assert open(out + ".regs.bin", "rb").read() == b"".join(parts)
    ctx.close(); idx.close()
