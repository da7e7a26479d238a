#!/usr/bin/env python
"""Round-2 experiment: the unsplit 1 M-read step with the Occ table in the file layout (BM2_OCC_LAYOUT=0: four 8-byte loads over both
sectors of a checkpoint) and in the device layout (1: one 256-bit load of one sector), per-stage ms from the library's CUDA events,
plus a knob sweep of the SMEM kernels' CTAs per SM under the new layout.  Usage: exp_layout.py <bench work dir> [steps]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package


def main():
    import torch
    pkg = load_package(); capi = pkg.capi
    work = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    reads = np.load(os.path.join(work, "reads.npy"))
    n, L = reads.shape
    codes = reads.reshape(-1); offs = np.arange(n + 1, dtype=np.int64) * L
    index = capi.Index(os.path.join(work, "ref.fa"))
    d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ref = None
    for layout, knobs in (("0", {}), ("1", {}), ("1", {"BM2_SMEM_CTAS": "4"}), ("1", {"BM2_SMEM_CTAS": "6"}), ("1", {"BM2_SMEM_CTAS": "10"}),
                          ("1", {"BM2_SMEM_CTAS": "12", "BM2_SMEM_P3_CTAS": "12"})):
        os.environ["BM2_OCC_LAYOUT"] = layout
        for k in ("BM2_SMEM_CTAS", "BM2_SMEM_P3_CTAS"):
            os.environ.pop(k, None)
        os.environ.update(knobs)
        ctx = capi.Context(0, index=index)
        ctx.set_sub_batches(1)
        best = None
        for _ in range(steps + 1):
            flush.fill_(1)
            ctx.seed_chain_extend_resident(codes, offs, d_codes.data_ptr(), d_offs.data_ptr(), False)
            torch.cuda.synchronize()
            st = ctx.stage_ms()
            if best is None or st["smem"] < best["smem"]:
                best = dict(st)
        ns = 100_000
        ctx.set_stream(None)
        r, o = ctx.seed_chain_extend(codes[:ns * L], offs[:ns + 1])
        sig = (r.tobytes(), o.tobytes())
        if ref is None:
            ref = sig
        print(json.dumps({"occ_layout": layout, "knobs": knobs, "stages_ms": {k: round(v, 2) for k, v in best.items()},
                          "counters": ctx.counters(), "same_regs_as_first": sig == ref}), flush=True)
        ctx.close()
    index.close()


if __name__ == "__main__":
    main()
