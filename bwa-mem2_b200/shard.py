"""Read sharding across GPUs.  Reads are independent through seeding/chaining/extension, so the
path shards by CHUNKS with no data-path collective: rank g of N takes chunks g, g+N, g+2N, ...  Chunk boundaries
are the reference's `-K` units kept at multiples of 512 reads, because (a) mem_pestat is per chunk
(src/bwamem.cpp:1368-1378), (b) hash_64(id+i) uses the global read index (src/bwamem.cpp:1327) and (c) the
reference's 512-read-block quirk (src/bwamem.cpp:835) is relative to the chunk start."""
from __future__ import annotations


def chunk_ranges(n_reads: int, chunk_reads: int):
    """[(start, end)) read ranges; chunk_reads is rounded up to a multiple of 512 (and of 2: pairs stay together)."""
    c = max(512, (chunk_reads + 511) // 512 * 512)
    return [(s, min(n_reads, s + c)) for s in range(0, n_reads, c)]


def rank_chunks(n_reads: int, chunk_reads: int, rank: int, world: int):
    return [(i, r) for i, r in enumerate(chunk_ranges(n_reads, chunk_reads)) if i % world == rank]
