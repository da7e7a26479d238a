#!/usr/bin/env python
"""Profiling driver: one context, the bench's 1 M-read batch UNSPLIT (one sub-batch: deterministic launch order), one warm-up
step and one profiled step.  Run under ncu with a kernel-name filter and -s <launches per step> -c <launches per step>.
Usage: prof_step.py <bench work dir> [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package


def main():
    import torch
    pkg = load_package(); capi = pkg.capi
    work = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    if not os.path.exists(os.path.join(work, "reads.npy")):      # same inputs as bench.py's default workload
        import bench
        bench.prepare_pipeline_inputs(work, 3_000_000_000, 500_000, seed=21)
    reads = np.load(os.path.join(work, "reads.npy"))
    n, L = reads.shape
    codes = reads.reshape(-1); offs = np.arange(n + 1, dtype=np.int64) * L
    index = capi.Index(os.path.join(work, "ref.fa"))
    ctx = capi.Context(0, index=index)
    ctx.set_sub_batches(1)
    d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
    torch.cuda.synchronize()
    for _ in range(steps):
        ctx.seed_chain_extend_resident(codes, offs, d_codes.data_ptr(), d_offs.data_ptr(), False)
    torch.cuda.synchronize()
    print({k: round(v, 2) for k, v in ctx.stage_ms().items()}, ctx.counters())
    ctx.close(); index.close()


if __name__ == "__main__":
    main()
