#!/usr/bin/env python
"""Experiment: the 1 M-read bench step under different tuning knobs (all read from the environment at every launch), one
process, one index upload.  Prints ms per step (CUDA events around the resident seam-2 call, L2 flushed) per setting.
Usage: exp_knobs.py <bench work dir> [steps]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package

KNOBS = ("BM2_TAIL_COOP", "BM2_LIGHT_SORTED", "BM2_BSW_UNROLL8", "BM2_TAIL_HEAVY", "BM2_CHAIN_HEAVY", "BM2_BSW_DYN", "BM2_SMEM_TEXT", "BM2_LANE_SKEW", "BM2_STAGE_TOKENS", "BM2_BSW_REGSHRINK", "BM2_CHAIN_COOP_MIN", "BM2_BSW_NTHR", "BM2_BSW_COL2", "BM2_BSW_SMEM_KB", "BM2_BSW_MAX_CTAS", "BM2_SMEM_CTAS", "BM2_SMEM_P3_CTAS", "BM2_STAGE_TOKENS")
CONFIGS = [
    dict(name="default, sub 1", sub=1),
    dict(name="default, sub 4", sub=4),
    dict(name="default, sub 1 (again)", sub=1),
    dict(name="default, sub 4 (again)", sub=4),
]


def main():
    import torch
    pkg = load_package(); capi = pkg.capi
    work = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if not os.path.exists(os.path.join(work, "reads.npy")):      # same inputs as bench.py's default workload
        import bench
        bench.prepare_pipeline_inputs(work, 3_000_000_000, 500_000, seed=21)
    reads = np.load(os.path.join(work, "reads.npy"))
    n, L = reads.shape
    codes = reads.reshape(-1); offs = np.arange(n + 1, dtype=np.int64) * L
    index = capi.Index(os.path.join(work, "ref.fa"))
    ctx = capi.Context(0, index=index)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    h_codes = torch.from_numpy(codes.copy()).pin_memory(); h_offs = torch.from_numpy(offs.copy()).pin_memory()      # the e2e leg's pinned inputs
    results = []
    ref_bytes = None
    for cfg in CONFIGS:
        for k in KNOBS:
            os.environ.pop(k, None)
        for k, v in cfg.items():
            if k.startswith("BM2_"):
                os.environ[k] = v
        ctx.set_sub_batches(cfg["sub"])
        e2e = bool(cfg.get("e2e"))          # host buffers in, regs out to host memory (wall clock), as bench.py's e2e leg
        ctx.set_stream(None if e2e else stream.cuda_stream)
        if e2e:
            ctx.seed_chain_extend(h_codes.numpy(), h_offs.numpy(), copy=False)
        else:
            ctx.seed_chain_extend_resident(codes, offs, d_codes.data_ptr(), d_offs.data_ptr(), False)      # warm-up / buffer growth
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            if e2e:
                t0 = time.perf_counter()
                n_regs = len(ctx.seed_chain_extend(h_codes.numpy(), h_offs.numpy(), copy=False)[0])
                ms.append((time.perf_counter() - t0) * 1e3)
                continue
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(stream)
            n_regs = ctx.seed_chain_extend_resident(codes, offs, d_codes.data_ptr(), d_offs.data_ptr(), False)
            b.record(stream); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        st = ctx.stage_ms()
        row = dict(name=cfg["name"], ms=[round(x, 2) for x in ms], best_ms=round(min(ms), 2), mreads_s=round(n / (min(ms) * 1e-3) / 1e6, 3),
                   n_regs=int(n_regs), stages={k: round(v, 2) for k, v in st.items()})
        results.append(row)
        print(json.dumps(row), flush=True)
    # parity of one knob setting against another on a slice (regs must be byte-identical whatever the knobs)
    ns = 65536
    outs = []
    for env in (dict(BM2_BSW_COL2="0", BM2_STAGE_TOKENS="0"), dict(), dict(BM2_BSW_MAX_CTAS="3", BM2_SMEM_CTAS="5"), dict(BM2_BSW_REGSHRINK="1", BM2_CHAIN_COOP_MIN="64"), dict(BM2_SMEM_TEXT="0", BM2_BSW_DYN="0", BM2_LIGHT_SORTED="1", BM2_TAIL_COOP="0"), dict(BM2_TAIL_HEAVY="2", BM2_CHAIN_HEAVY="16", BM2_BSW_UNROLL8="1"), dict(BM2_BSW_REGSHRINK="0")):
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        ctx.set_sub_batches(4)
        ctx.set_stream(None)
        r, o = ctx.seed_chain_extend(codes[:ns * L], offs[:ns + 1])
        outs.append((r.tobytes(), o.tobytes()))
    print("parity across knob settings:", all(x == outs[0] for x in outs), flush=True)
    ctx.close(); index.close()


if __name__ == "__main__":
    main()
