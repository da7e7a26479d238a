#!/usr/bin/env python
"""Experiment: N contexts (own stream, own scratch, own index copy) on ONE GPU, each aligning 1/N of the batch from its own
host thread.  The SMEM stage is latency-bound and BSW is ALU-bound, so co-running sub-batches can fill each other's
stalls and the host-sync gaps.  Prints wall-clock ms per full batch for N = 1..4 (inputs: bench.py's cached work dir)."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package


def main():
    import torch
    pkg = load_package(); capi = pkg.capi
    work = sys.argv[1]
    maxn = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reads = np.load(os.path.join(work, "reads.npy"))
    n, L = reads.shape
    index = capi.Index(os.path.join(work, "ref.fa"))
    for N in range(1, maxn + 1):
        cuts = [((n * k // N) // 512) * 512 for k in range(N)] + [n]
        parts = []
        for k in range(N):
            r = reads[cuts[k]:cuts[k + 1]]
            codes = np.ascontiguousarray(r.reshape(-1)); offs = np.arange(len(r) + 1, dtype=np.int64) * L
            ctx = capi.Context(0, index=index)
            st = torch.cuda.Stream()
            ctx.set_stream(st.cuda_stream)
            parts.append(dict(ctx=ctx, st=st, codes=codes, offs=offs, d_codes=torch.from_numpy(codes).cuda(), d_offs=torch.from_numpy(offs).cuda()))
        def run(p):
            p["ctx"].seed_chain_extend_resident(p["codes"], p["offs"], p["d_codes"].data_ptr(), p["d_offs"].data_ptr(), False)
        for p in parts:       # sequential warm-up (buffer growth, one-time attribute calls)
            run(p); run(p)
        torch.cuda.synchronize()
        times = []
        for it in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            th = [threading.Thread(target=run, args=(p,)) for p in parts]
            for t in th: t.start()
            for t in th: t.join()
            torch.cuda.synchronize(); times.append((time.perf_counter() - t0) * 1e3)
        print(f"N={N}: ms per {n}-read batch: " + " ".join(f"{t:.1f}" for t in times) + f"  -> {n / (min(times) * 1e-3) / 1e6:.2f} M reads/s", flush=True)
        for p in parts: p["ctx"].close()
        del parts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
