"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/*/ref_driver).

Run here (container with /root/reference):  python tests/golden/make_golden.py
  1. synthesises the C0 input (200 kbp reference with repeats + N runs, 500 pairs 2x151)
  2. indexes it with the reference binary, aligns with ref_driver -t 1 and stage dumps
  3. writes  bsw_c0.npz (4000 extension jobs incl. the reference's outputs),
             c0_index/ (the reference-built index files), c0_reads.npz, c0_stages.npz (SMEMs,
             chains, regs), c0.sam (reference SAM without @PG)
"""
import os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import refdump  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    synth = __import__("importlib").import_module("bwa_mem2_b200.synth") if load_package() else None
    work = "/tmp/bm2_golden_c0"
    os.makedirs(work, exist_ok=True)
    ctg = synth.make_reference(200_000, seed=1, n_contigs=4)
    synth.write_fasta(f"{work}/ref.fa", ctg)
    r1, r2 = synth.make_pairs(ctg, 500, seed=2)
    synth.write_fastq(f"{work}/r1.fq", r1, "p"); synth.write_fastq(f"{work}/r2.fq", r2, "p")
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    bindir = os.path.join(ROOT, "oracle", "_ref", isa)
    subprocess.check_call([f"{bindir}/bwa-mem2", "index", f"{work}/ref.fa"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    env = dict(os.environ, BM2_DUMP_PREFIX=f"{work}/dump")
    with open(f"{work}/out.sam", "w") as f:
        subprocess.check_call([f"{bindir}/ref_driver", "mem", "-t", "1", "-K", "100000000", f"{work}/ref.fa",
                               f"{work}/r1.fq", f"{work}/r2.fq"], stdout=f, stderr=subprocess.DEVNULL, env=env)
    # --- BSW golden
    g = refdump.merge_bsw(refdump.read_bsw(f"{work}/dump.bsw.bin"))
    assert len(g) == 1
    g = g[0]
    rng = np.random.default_rng(7)
    n = len(g["h0"])
    sel = np.sort(rng.choice(n, size=min(4000, n), replace=False))
    len1 = g["len1"][sel]; len2 = g["len2"][sel]
    ref = np.concatenate([g["ref"][g["idr"][i]:g["idr"][i] + g["len1"][i]] for i in sel])
    qer = np.concatenate([g["qer"][g["idq"][i]:g["idq"][i] + g["len2"][i]] for i in sel])
    idr = np.concatenate([[0], np.cumsum(len1[:-1])]).astype(np.int32)
    idq = np.concatenate([[0], np.cumsum(len2[:-1])]).astype(np.int32)
    out = g["out"][sel]
    np.savez_compressed(os.path.join(HERE, "bsw_c0.npz"), w=g["w"], len1=len1, len2=len2, h0=g["h0"][sel], idr=idr, idq=idq,
                        ref=ref, qer=qer, kind=g["kind"][sel], out_score=out[:, 0], out_tle=out[:, 1], out_gtle=out[:, 2],
                        out_qle=out[:, 3], out_gscore=out[:, 4], out_max_off=out[:, 5],
                        **{"p_" + k: v for k, v in g["params"].items()})
    # --- index + reads + stages
    idx_dir = os.path.join(HERE, "c0_index"); os.makedirs(idx_dir, exist_ok=True)
    for suf in (".bwt.2bit.64", ".0123", ".ann", ".amb", ".pac"):
        subprocess.check_call(["cp", f"{work}/ref.fa{suf}", os.path.join(idx_dir, "ref.fa" + suf)])
    reads = np.empty((1000, 151), np.uint8); reads[0::2] = r1; reads[1::2] = r2
    np.savez_compressed(os.path.join(HERE, "c0_reads.npz"), reads=reads)
    smems = refdump.read_smems(f"{work}/dump.smem.bin")
    chains = refdump.read_chains(f"{work}/dump.chains.bin")
    regs, reg_off = refdump.read_regs(f"{work}/dump.regs.bin")
    ch_rows = []; seed_rows = []; ch_off = [0]
    for rd in chains:
        for c in rd:
            ch_rows.append((c["pos"], c["seqid"], c["rid"], c["n"], len(seed_rows), c["w"], c["kept"], c["first"], c["is_alt"], c["frac_rep"]))
            for s in c["seeds"]:
                seed_rows.append(tuple(int(x) for x in s))
        ch_off.append(len(ch_rows))
    np.savez_compressed(os.path.join(HERE, "c0_stages.npz"), smems=smems,
                        chains=np.array(ch_rows, dtype=[("pos", "<i8"), ("seqid", "<i4"), ("rid", "<i4"), ("n", "<i4"),
                                                        ("seed_off", "<i4"), ("w", "<i4"), ("kept", "<i4"), ("first", "<i4"),
                                                        ("is_alt", "<i4"), ("frac_rep", "<f4")]),
                        seeds=np.array(seed_rows, dtype=[("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4"), ("score", "<i4")]),
                        chain_off=np.array(ch_off, np.int64), regs=regs, reg_off=reg_off)
    with open(f"{work}/out.sam") as f, open(os.path.join(HERE, "c0.sam"), "w") as o:
        for line in f:
            if not line.startswith("@PG"):
                o.write(line)
    print("golden written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
