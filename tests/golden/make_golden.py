#!/usr/bin/env python
"""Generate tests/golden/jf_*.json by running the REFERENCE's own engine
(/root/reference/library/jellyfish-linux, Jellyfish 2.3.0) with the reference's literal argv
(library/identify.py:82-87, library/Vote_Strain_L2_Lasso_new_sp.py:357-372).

Run in the build container (the binary cannot be fetched on the GPU box):
    make -C oracle ref && python tests/golden/make_golden.py
Each golden file holds the inputs (k-mer FASTA text, read-file texts, k, how the reads were fed)
and the parsed `dump -c` output {KMER: count}.  Deterministic: NumPy default_rng with fixed seeds.
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
JF = os.path.join(ROOT, "oracle", "_ref", "jellyfish-linux")

COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def revcomp(s):
    return s.translate(COMP)[::-1]


def rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def fastq(reads, crlf=False, qual_at=False):
    nl = "\r\n" if crlf else "\n"
    out = []
    for i, r in enumerate(reads):
        q = "I" * len(r)
        if qual_at and len(q) > 0:
            q = "@" + q[1:]
        out.append("@r%d%s%s%s+%s%s%s" % (i, nl, r, nl, nl, q, nl))
    return "".join(out)


def fasta_db(kmers, header="1"):
    return "".join(">%s\n%s\n" % (header if header else str(i + 1), k) for i, k in enumerate(kmers))


def run_jf(fasta_text, read_texts, k, mode, threads):
    """mode: 'files' (plain argv), 'zcat' (zcat a b | jellyfish count /dev/fd/0)."""
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "kmer.fa")
        with open(fa, "w", newline="") as f:
            f.write(fasta_text)
        paths = []
        for i, t in enumerate(read_texts):
            if mode == "zcat":
                p = os.path.join(td, "r%d.fq.gz" % i)
                with gzip.open(p, "wb") as f:
                    f.write(t.encode())
            else:
                p = os.path.join(td, "r%d.fq" % i)
                with open(p, "w", newline="") as f:
                    f.write(t)
            paths.append(p)
        jf = os.path.join(td, "o.jf")
        if mode == "zcat":
            cmd = "zcat %s | %s count /dev/fd/0 -m %d -s 100M -t %d --if %s -o %s" % (
                " ".join(paths), JF, k, threads, fa, jf)
        else:
            cmd = "%s count -m %d -s 100M -t %d --if %s -o %s %s" % (JF, k, threads, fa, jf, " ".join(paths))
        subprocess.check_call(cmd, shell=True)
        txt = subprocess.check_output([JF, "dump", "-c", jf]).decode()
    dump = {}
    for line in txt.splitlines():
        a, b = line.split(" ")
        dump[a] = int(b)
    return dump


def cases():
    rng = np.random.default_rng(20260117)
    k = 31
    g = rand_seq(rng, 400)
    km = [g[i:i + k] for i in range(0, 300, 7)]
    out = []

    # 1 strand specificity: fwd records + one rc record; reads fwd and rc
    db = km[:10] + [revcomp(km[3])]
    out.append(("strand", fasta_db(db), [fastq([g[0:150], revcomp(g[0:150]), g[100:250]])], k, "files", 8))
    # 2 N / IUPAC break, lowercase reads
    r = g[0:150]
    rn = r[:20] + "N" + r[21:]
    ri = r[:60] + "R" + r[61:]
    out.append(("nbreak_lower", fasta_db(km[:20]), [fastq([rn, ri, r.lower(), r[:75] + r[75:].lower()])], k, "files", 8))
    # 3 --if oddities: lowercase duplicate, N record, short/long record, absent poly-G, exact duplicates
    db = km[:6] + [km[2].lower(), km[4][:10] + "N" + km[4][11:], km[5][:20], g[50:50 + 40], "G" * k, km[1], km[0]]
    out.append(("if_oddities", fasta_db(db), [fastq([g[0:200], g[0:200], g[30:120]])], k, "files", 8))
    # 4 reads shorter than / equal to / one longer than k
    out.append(("short_reads", fasta_db(km[:12]), [fastq([g[0:30], g[0:31], g[7:39], g[14:20], ""])], k, "files", 4))
    # 5 CRLF + quality line starting with '@'
    out.append(("crlf", fasta_db(km[:12]), [fastq([g[0:150], g[50:200]], crlf=True)], k, "files", 8))
    out.append(("qual_at", fasta_db(km[:12]), [fastq([g[0:150], g[50:200], g[7:90]], qual_at=True)], k, "files", 8))
    # 6 paired files (argv concat) and the zcat pipe
    out.append(("pe_files", fasta_db(km[:25]), [fastq([g[0:150], g[40:190]]), fastq([revcomp(g[100:250]), g[5:155]])], k, "files", 8))
    out.append(("pe_zcat", fasta_db(km[:25]), [fastq([g[0:150], g[40:190]]), fastq([revcomp(g[100:250]), g[5:155]])], k, "zcat", 8))
    # 7 wrapped FASTQ and FASTA reads (Jellyfish joins lines)
    wrapped = "@w1\n%s\n%s\n+\n%s\n%s\n@w2\n%s\n+\n%s\n" % (g[0:80], g[80:170], "I" * 80, "I" * 90, g[20:120], "I" * 100)
    out.append(("wrapped_fastq", fasta_db(km[:30]), [wrapped], k, "files", 4))
    fa_reads = ">a\n%s\n%s\n>b\n%s\n" % (g[0:70], g[70:200], g[150:300])
    out.append(("fasta_reads", fasta_db(km), [fa_reads], k, "files", 4))
    # 8 L2-style: k=21, '>kid' headers, counts 0/1/2/3 (remove_1 matters)
    k2 = 21
    g2 = rand_seq(rng, 600)
    km2 = [g2[i:i + k2] for i in range(0, 500, 5)]
    reads2 = [g2[0:150], g2[0:150], g2[100:250], g2[100:250], g2[100:250], g2[300:450]]
    out.append(("l2_k21", fasta_db(km2, header=None), [fastq(reads2)], k2, "files", 10))
    # 9 extreme k
    g3 = rand_seq(rng, 300)
    out.append(("k32", fasta_db([g3[i:i + 32] for i in range(0, 200, 3)]), [fastq([g3[0:150], g3[100:300]])], 32, "files", 4))
    out.append(("k11", fasta_db([g3[i:i + 11] for i in range(0, 200, 3)]), [fastq([g3[0:150], g3[100:300]])], 11, "files", 4))
    # 9b k = 32 poly-T packs to all ones in 2 bits/base: a legal key, not an empty marker
    out.append(("polyT_k32", fasta_db(["T" * 32, "A" * 32, "T" * 31 + "G"]),
                [fastq(["T" * 40, "A" * 33 + "T" * 31 + "G"])], 32, "files", 4))
    # 10 homopolymers / low complexity (many identical windows in one read)
    out.append(("lowcomplex", fasta_db(["A" * 31, "AC" * 15 + "A", "CA" * 15 + "C", "T" * 31]),
                [fastq(["A" * 150, "AC" * 75, "T" * 40 + "N" + "T" * 40])], 31, "files", 8))
    # 11 medium random: both strands as separate records, errors, Ns, random strand
    rng = np.random.default_rng(7)
    G = rand_seq(rng, 30000)
    pos = rng.choice(30000 - 31, 3000, replace=False)
    recs = []
    for p in pos:
        recs.append(G[p:p + 31]); recs.append(revcomp(G[p:p + 31]))
    perm = rng.permutation(len(recs))
    recs = [recs[i] for i in perm]
    reads = []
    for _ in range(1500):
        s = int(rng.integers(0, 30000 - 150))
        r = list(G[s:s + 150])
        for j in np.nonzero(rng.random(150) < 0.01)[0]:
            r[j] = "ACGT"[int(rng.integers(0, 4))]
        for j in np.nonzero(rng.random(150) < 0.002)[0]:
            r[j] = "N"
        r = "".join(r)
        if rng.random() < 0.5:
            r = revcomp(r)
        reads.append(r)
    out.append(("medium_random", fasta_db(recs), [fastq(reads[:800]), fastq(reads[800:])], 31, "zcat", 8))
    return out


def main():
    if not os.path.exists(JF):
        sys.exit("oracle/_ref/jellyfish-linux missing: run `make -C oracle ref` in the build container")
    for name, fa, reads, k, mode, t in cases():
        dump = run_jf(fa, reads, k, mode, t)
        path = os.path.join(HERE, "jf_%s.json" % name)
        with open(path, "w") as f:
            json.dump({"name": name, "k": k, "mode": mode, "threads": t, "engine": "jellyfish 2.3.0",
                       "fasta": fa, "reads": reads, "dump": dump}, f, sort_keys=True)
        print("%-16s k=%-2d records=%-5d dumped=%-5d total=%d" % (
            name, k, fa.count(">"), len(dump), sum(dump.values())))


if __name__ == "__main__":
    main()
