"""Read sharding across the GPUs of one box: one process per GPU (torch.distributed), no data-path collective.

Reads are independent through seeding / chaining / extension, so the path shards by CHUNKS: rank g of N takes chunks g, g+N, g+2N, ...
A chunk is what the reference hands to one mem_process_seqs call, because (a) mem_pestat is per chunk (src/bwamem.cpp:1368-1378),
(b) hash_64(id + i) uses the read's index in the whole input (src/bwamem.cpp:1327; `id_base` below) and (c) the 512-read-block quirk
(src/bwamem.cpp:835) is relative to the chunk start.  `chunk_bounds` therefore cuts the stream exactly where bseq_read_orig does
(src/bwa.cpp:170-216: read (pair) by read (pair) until the running base count reaches the task size, pairs kept together), with
the task size of src/fastmap.cpp:943-949 (`-K`, else chunk_size x n_threads): a sharded run and `bwa-mem2 mem` with the same -K see
the same chunks.

Start-up: rank 0 reads the index files, every other rank receives the four big arrays (Occ table, sampled SA x 2, reference) by
broadcast - NCCL over NVLink when the group is NCCL - and adopts them in place (bm2_create_resident): one disk read and one PCIe upload
per box instead of one per GPU.  Results stay in each rank's pinned host buffers; `chunk_table` (one small all_gather) tells every rank
which rank holds which chunk, `gather_in_order` brings regs to one rank in input order for consumers that want a single stream.
"""
from __future__ import annotations
import ctypes as C
import time
import numpy as np


def task_size(chunk_size: int = 10_000_000, n_threads: int = 1, fixed_k: int = 0) -> int:
    """Bases per chunk as main_mem sets it (src/fastmap.cpp:943-949): -K if given, else opt->chunk_size * opt->n_threads."""
    return int(fixed_k) if fixed_k and fixed_k > 0 else int(chunk_size) * int(n_threads)


def chunk_bounds(read_lens, chunk_bases: int, paired: bool = True):
    """[(start, end)) read ranges of the chunks bseq_read_orig forms (src/bwa.cpp:170-216).  read_lens: array of read lengths in
    input order (mates adjacent when paired), or (n_reads, uniform_length)."""
    if isinstance(read_lens, tuple):
        n, L = int(read_lens[0]), int(read_lens[1])
        if n == 0:
            return []
        unit = 2 if paired else 1
        per = max(1, -(-int(chunk_bases) // (unit * max(L, 1)))) * unit if L > 0 else n       # units until size >= chunk_bases
        return [(s, min(n, s + per)) for s in range(0, n, per)]
    lens = np.asarray(read_lens, dtype=np.int64)
    n = len(lens)
    cs = np.concatenate([[0], np.cumsum(lens)])
    out = []
    s = 0
    while s < n:
        e = int(np.searchsorted(cs, cs[s] + int(chunk_bases), side="left"))      # first e with sum(lens[s:e]) >= chunk_bases
        e = max(e, s + 1)
        if paired and (e - s) % 2:
            e += 1
        e = min(e, n)
        out.append((s, e))
        s = e
    return out


def rank_chunks(bounds, rank: int, world: int):
    """(chunk id, (start, end)) of the chunks rank `rank` of `world` aligns: g, g + N, g + 2N, ..."""
    return [(i, r) for i, r in enumerate(bounds) if i % world == rank]


# ---- compatibility with round 1's helpers (block-aligned fixed-size chunks; used by the unit test of the dealing) ----------------
def chunk_ranges(n_reads: int, chunk_reads: int):
    """[(start, end)) ranges of fixed-size chunks, the size rounded up to a multiple of 512 reads.  NOT the reference's chunking
    (that is chunk_bounds): parity with a reference run holds only when its chunks coincide."""
    c = max(512, (chunk_reads + 511) // 512 * 512)
    return [(s, min(n_reads, s + c)) for s in range(0, n_reads, c)]


class ShardedAligner:
    """Per-rank front end of the hot path for one box: shared start-up (index read once, broadcast), chunk dealing, ordered results.

    compute: callable(codes, offsets) -> (regs, read_off) that stands in for the GPU context (tests on machines without a GPU pass
    the CPU oracle); default: a capi.Context on `device` created from the broadcast index."""

    def __init__(self, capi, prefix: str, device: int = 0, opt=None, compute=None, keep_host_index: bool = False):
        import torch
        import torch.distributed as dist
        self.capi = capi
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        self.device = device
        self.startup = {}
        self._keep = []
        self.ctx = None
        self.index = None
        self.compute = compute
        if compute is not None:
            return
        use_cuda = torch.cuda.is_available()
        if not use_cuda:
            raise RuntimeError("ShardedAligner: no CUDA device and no compute stand-in (the product has no CPU fallback)")
        if self.world == 1:
            t0 = time.perf_counter()
            self.index = capi.Index(prefix)
            self.startup["index_load_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            self.ctx = capi.Context(device, index=self.index, opt=opt)
            self.startup["upload_s"] = time.perf_counter() - t0
            return
        nccl = dist.get_backend() == "nccl"
        dev = torch.device("cuda", device)
        meta = [None]
        host = None
        t0 = time.perf_counter()
        if self.rank == 0:
            self.index = capi.Index(prefix)
            d = self.index.desc
            N = int(d.reference_seq_len); l_pac = int(d.l_pac); ns = int(d.n_seqs)
            n_occ = (N >> 6) + 1; n_sa = (N >> 3) + 1

            def view(p, nbytes):
                return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))
            host = [view(d.cp_occ, n_occ * 64), view(d.sa_ms_byte, n_sa), view(d.sa_ls_word, n_sa * 4), view(d.ref_string, 2 * l_pac)]
            i32 = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n,)).copy()
            i64 = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), shape=(n,)).copy()
            meta[0] = dict(N=N, l_pac=l_pac, n_seqs=ns, count=[int(x) for x in d.count], sentinel=int(d.sentinel_index),
                           ann_offset=i64(d.ann_offset, ns), ann_len=i32(d.ann_len, ns),
                           ann_is_alt=i32(d.ann_is_alt, ns) if d.ann_is_alt else np.zeros(ns, np.int32),
                           sizes=[int(h.nbytes) for h in host])
        self.startup["index_load_s"] = time.perf_counter() - t0          # (rank 0 only; the others wait in the broadcast)
        t0 = time.perf_counter()
        self.dist.broadcast_object_list(meta, src=0)
        m = meta[0]
        big = []
        for k, nbytes in enumerate(m["sizes"]):
            if nccl:
                t = torch.from_numpy(host[k]).to(dev) if self.rank == 0 else torch.empty(nbytes, dtype=torch.uint8, device=dev)
                self.dist.broadcast(t, src=0)
            else:           # gloo (tests): through host tensors, then one upload per rank
                t = torch.from_numpy(host[k].copy()) if self.rank == 0 else torch.empty(nbytes, dtype=torch.uint8)
                self.dist.broadcast(t, src=0)
                t = t.to(dev)
            big.append(t)
        torch.cuda.synchronize(dev)
        self.startup["broadcast_s"] = time.perf_counter() - t0
        self.meta = m
        self.big = big                          # [cp_occ, sa_ms, sa_ls, ref] as device byte tensors; owned here, adopted by the context
        desc = capi.IndexDesc()
        desc.reference_seq_len = m["N"]; desc.l_pac = m["l_pac"]; desc.n_seqs = m["n_seqs"]; desc.sentinel_index = m["sentinel"]
        for i in range(5):
            desc.count[i] = m["count"][i]
        desc.cp_occ, desc.sa_ms_byte, desc.sa_ls_word, desc.ref_string = (t.data_ptr() for t in big)
        ao = np.ascontiguousarray(m["ann_offset"], np.int64); al = np.ascontiguousarray(m["ann_len"], np.int32)
        aa = np.ascontiguousarray(m["ann_is_alt"], np.int32)
        self._keep += [ao, al, aa, desc]
        desc.ann_offset = ao.ctypes.data; desc.ann_len = al.ctypes.data; desc.ann_is_alt = aa.ctypes.data
        t0 = time.perf_counter()
        self.ctx = capi.Context(device, index=desc, opt=opt, resident=True)
        self.startup["adopt_s"] = time.perf_counter() - t0
        if self.rank == 0 and not keep_host_index:
            self.index.close(); self.index = None          # the host copy is not needed any more

    # ---- the dealing --------------------------------------------------------------------------------------------------------------
    def my_chunks(self, bounds):
        return rank_chunks(bounds, self.rank, self.world)

    def align_chunks(self, codes, offsets, bounds):
        """Aligns this rank's chunks of the stream (codes, offsets) -> [(chunk id, start read, regs, read_off)] (copies)."""
        f = self.compute if self.compute is not None else (lambda c, o: self.ctx.seed_chain_extend(c, o))
        out = []
        offsets = np.asarray(offsets, np.int64)
        for ci, (s, e) in self.my_chunks(bounds):
            o = offsets[s:e + 1] - offsets[s]
            regs, ro = f(codes[offsets[s]:offsets[e]], o)
            out.append((ci, s, regs, ro))
        return out

    def chunk_table(self, results):
        """Every rank learns (chunk id, owner rank, n_reads, n_regs) of all chunks: the index a writer needs to emit chunks in input
        order from the ranks' buffers.  One small all_gather."""
        mine = [(ci, self.rank, len(ro) - 1, int(ro[-1])) for ci, _, _, ro in results]
        if not self.dist:
            return sorted(mine)
        allm = [None] * self.world
        self.dist.all_gather_object(allm, mine)
        return sorted(x for m in allm for x in m)

    def gather_in_order(self, results, dst: int = 0):
        """Regs of ALL chunks on rank `dst`, concatenated in input order, with per-read offsets: the single-stream view of the sharded
        run (equal to one unsharded pass chunk by chunk).  Returns (regs, read_off) on dst, (None, None) elsewhere."""
        if not self.dist:
            allr = [results]
        else:
            payload = [(ci, s, np.asarray(regs).tobytes(), np.asarray(ro, np.int64)) for ci, s, regs, ro in results]
            allr = [None] * self.world if self.rank == dst else None
            self.dist.gather_object(payload, allr, dst=dst)
            if self.rank != dst:
                return None, None
        chunks = sorted((c for r in allr for c in r), key=lambda c: c[0])
        dt = self.capi.REG_DT
        regs = [np.frombuffer(c[2], dtype=dt) if isinstance(c[2], (bytes, bytearray)) else np.asarray(c[2]) for c in chunks]
        counts = [np.diff(np.asarray(c[3], np.int64)) for c in chunks]
        allregs = np.concatenate(regs) if regs else np.zeros(0, dt)
        off = np.concatenate([[0], np.cumsum(np.concatenate(counts))]).astype(np.int64) if counts else np.zeros(1, np.int64)
        return allregs, off

    def close(self):
        if self.ctx is not None:
            self.ctx.close(); self.ctx = None
        if self.index is not None:
            self.index.close(); self.index = None
        self.big = None
