#!/usr/bin/env python
"""BASELINE configs 1 and 2 end to end, from FILES, through the CLI drivers:
   C1: `composition` on 10 synthetic MAGs (~3 Mbp, ~300 contigs each) written as FASTA;
   C2: `coverage` on the same MAGs with 3 synthetic coordinate-sorted BAMs (~4-5 M reads each).
Wall-clock per stage is printed as JSON (host decode included: BAM inflate + record filter run on the
host cores, as north_star prescribes).  Usage: python tools/run_c1_c2.py [workdir]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpurify2_b200 import _bam, bamio, cli, synth  # noqa: E402


def main():
    work = Path(sys.argv[1] if len(sys.argv) > 1 else "/tmp/mp2_c1c2")
    work.mkdir(parents=True, exist_ok=True)
    d = synth.make_mags(10, seed=1)
    names, genomes = [], []
    t0 = time.perf_counter()
    for m in range(10):
        p = work / f"mag{m:02d}.fna"
        with open(p, "w") as f:
            for c in range(d["mag_off"][m], d["mag_off"][m + 1]):
                nm = f"mag{m:02d}_contig{c}"
                names.append(nm)
                seq = d["bases"][d["offsets"][c]:d["offsets"][c + 1]].tobytes().decode()
                f.write(f">{nm}\n")
                f.write("\n".join(seq[o:o + 80] for o in range(0, len(seq), 80)) + "\n")
        genomes.append(p)
    bams, n_reads = [], []
    rng = np.random.default_rng(1)
    for s in range(3):
        st, en, ev_off = synth.make_events(d["lengths"], d["mag_off"], s, seed=1)
        tid = np.repeat(np.arange(len(d["lengths"])), np.diff(ev_off))
        nm = rng.binomial(150, 0.01, len(st))
        flag = rng.choice([0, 1, 0x100, 0x800], len(st), p=[0.93, 0.05, 0.01, 0.01])
        p = work / f"sample{s}.bam"
        bamio.write_bam_fixed(str(p), names, d["lengths"], tid, st, nm, flag=flag)
        bams.append(p)
        n_reads.append(len(st))
    t_write = time.perf_counter() - t0
    out = work / "out"
    L = int(d["offsets"][-1])
    res = {"contigs": len(names), "bases": L, "reads_per_bam": n_reads, "write_inputs_s": round(t_write, 2),
           "host_cores": os.cpu_count()}
    for rep in ("cold", "warm"):
        for f in (out / "tnfs.pickle.gz", out / "coverages.pickle.gz"):
            if f.exists():
                f.unlink()
        t0 = time.perf_counter()
        cli.cli(["composition", *map(str, genomes), str(out), "-q"])
        t1 = time.perf_counter()
        cli.cli(["coverage", *map(str, genomes), str(out), "--bam_files", *map(str, bams), "-q"])
        t2 = time.perf_counter()
        res[f"composition_cli_s_{rep}"] = round(t1 - t0, 3)
        res[f"coverage_cli_s_{rep}"] = round(t2 - t1, 3)
    res["composition_gbp_s_from_fasta"] = round(L / res["composition_cli_s_warm"] / 1e9, 4)
    res["coverage_reads_s_from_bam"] = round(sum(n_reads) / res["coverage_cli_s_warm"], 1)
    t0 = time.perf_counter()
    for p in bams:
        with _bam.BamEvents(p, 0.97, threads=os.cpu_count() or 1) as b:
            kept = b.n_kept
    res["bam_decode_only_s"] = round(time.perf_counter() - t0, 3)
    res["bam_decode_reads_s"] = round(sum(n_reads) / res["bam_decode_only_s"], 1)
    res["bam_bytes"] = [os.path.getsize(p) for p in bams]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
