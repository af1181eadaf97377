"""Tiny synthetic StrainScan database written directly in the reference's on-disk layout
(writers: Build_tree.py:494-698, Build_kmer_sets_..._sp.py:397-410,
Recls_withR_new.py:110-115, Build_overlap_matrix_sp.py:88-98), plus read sets drawn from its strains.
The reference's own builder cannot run in this image (dashing / R / sibeliaz absent), so the
end-to-end parity tests use this stand-in for test_run.sh's DB_Small (BASELINE.json configs[0]).
Deterministic: NumPy default_rng(seed) only.
"""
import gzip
import os
import pickle

import numpy as np
import scipy.sparse as sp

K = 31
_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(b):
    return b.translate(_COMP)[::-1]


def _rand_seq(rng, n):
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes()


class SynthDB:
    """4 clusters with 3/2/1/2 strains; 7-node search tree: leaves 1..4, internal 5=(1,2) 6=(3,4), root 7."""

    def __init__(self, seed=20260117, seg_len=5000, n_seg=12, snp_rate=0.004, prefix="GCF"):
        rng = np.random.default_rng(seed)
        self.seed = seed
        self.n_strains = {1: 3, 2: 2, 3: 1, 4: 2}
        self.parent = {1: 5, 2: 5, 3: 6, 4: 6, 5: 7, 6: 7, 7: None}
        self.children = {5: (1, 2), 6: (3, 4), 7: (5, 6)}
        self.seg_len, self.n_seg = seg_len, n_seg
        # segment i belongs to the root (i%3==0), the cluster's internal parent (1) or the leaf itself (2)
        segs = {}
        self.owner = {}
        self.genome = {}
        for c in (1, 2, 3, 4):
            parts = []
            for i in range(n_seg):
                own = (7, self.parent[c], c)[i % 3]
                if (own, i) not in segs:
                    segs[(own, i)] = _rand_seq(rng, seg_len)
                parts.append(segs[(own, i)])
                self.owner[(c, i)] = own
            self.genome[c] = b"".join(parts)
        # strains: SNPs inside the leaf-owned segments only
        self.strain_name, self.strain_genome, self.snps = {}, {}, {}
        for c in (1, 2, 3, 4):
            for s in range(1, self.n_strains[c] + 1):
                g = bytearray(self.genome[c])
                pos = []
                if self.n_strains[c] > 1:
                    for i in range(n_seg):
                        if i % 3 != 2:
                            continue
                        lo = i * seg_len
                        for p in np.nonzero(rng.random(seg_len) < snp_rate)[0]:
                            p = lo + int(p)
                            if p < lo + K or p >= lo + seg_len - K:
                                continue
                            g[p] = b"ACGT"[(b"ACGT".index(g[p]) + 1 + int(rng.integers(0, 3))) % 4]
                            pos.append(p)
                self.strain_name[(c, s)] = "%s_C%d_S%d" % (prefix, c, s)
                self.strain_genome[(c, s)] = bytes(g)
                self.snps[(c, s)] = pos
        self.rng = rng

    # ---- Tree_database ------------------------------------------------------------------------
    def _node_kmers(self):
        """node -> list of k-mer strings (both strands as separate records, Build_tree.py:101-109)."""
        out = {v: [] for v in range(1, 8)}
        done = set()
        for c in (1, 2, 3, 4):
            snp_all = sorted(set(p for s in range(1, self.n_strains[c] + 1) for p in self.snps[(c, s)]))
            bad = np.zeros(len(self.genome[c]) + 1, dtype=bool)
            for p in snp_all:
                bad[max(0, p - K + 1):p + 1] = True           # windows covering a SNP are not cluster-wide
            for i in range(self.n_seg):
                own = self.owner[(c, i)]
                if (own, i) in done:
                    continue
                done.add((own, i))
                g = self.genome[c]
                lo = i * self.seg_len
                for w in range(lo, lo + self.seg_len - K + 1):
                    if own == c and bad[w]:
                        continue
                    km = g[w:w + K]
                    out[own].append(km)
                    out[own].append(revcomp(km))
        return out

    def write(self, db_dir, low_mem=False, overlap=False):
        """overlap: leaf 4 is a *reconstructed* node (reconstructed_nodes.txt, label 'o2'): its kmers/4 list also
        holds 3000 ordinals of leaf 1's k-mers, and overlapping_info/1 (+ 1_supple) lists their positions in that
        list, so that once cluster 1 is identified the search corrects node 4's profile (adjust_profile,
        identify.py:167-191) instead of taking it at face value.
        low_mem: the `-e 1` memory-efficient layout -- kmer.fa keeps only the lexicographically smaller
        strand of every k-mer (Build_tree_mem.py:107-110) and <DB>/Memory_DB makes StrainScan.py:188-196 route
        the search through identify_low_mem (reads are still counted without -C: only same-strand hits)."""
        rng = np.random.default_rng(self.seed + 1)
        tdb = os.path.join(db_dir, "Tree_database")
        os.makedirs(os.path.join(tdb, "kmers"), exist_ok=True)
        os.makedirs(os.path.join(db_dir, "Cluster_Result"), exist_ok=True)
        nk = self._node_kmers()
        if low_mem:
            nk = {v: [km for km in kms if km <= revcomp(km)] for v, kms in nk.items()}
            open(os.path.join(db_dir, "Memory_DB"), "w").close()
        allk = [(v, km) for v in range(1, 8) for km in dict.fromkeys(nk[v])]
        order = rng.permutation(len(allk))
        node_ord = {v: [] for v in range(1, 8)}
        with open(os.path.join(tdb, "kmer.fa"), "wb") as f:
            for idx, j in enumerate(order):
                v, km = allk[j]
                f.write(b">1\n" + km + b"\n")
                node_ord[v].append(idx)
        if overlap:
            shared = node_ord[1][:3000]
            first = len(node_ord[4])
            node_ord[4] = node_ord[4] + shared
            os.makedirs(os.path.join(tdb, "overlapping_info"), exist_ok=True)
            with open(os.path.join(tdb, "overlapping_info", "1"), "w") as f:
                f.write(" ".join(str(first + i) for i in range(len(shared))) + "\n")
            with open(os.path.join(tdb, "overlapping_info", "1_supple"), "w") as f:
                f.write("4 0\n")
        for v in range(1, 8):
            with open(os.path.join(tdb, "kmers", str(v)), "w") as f:
                f.write("".join("%d " % x for x in node_ord[v]))
        with open(os.path.join(tdb, "node_length.txt"), "w") as f:
            for v in range(1, 8):
                f.write("%d\t%d\n" % (v, len(node_ord[v])))
        with open(os.path.join(tdb, "reconstructed_nodes.txt"), "w") as f:
            f.write("4\n" if overlap else "")
        with open(os.path.join(tdb, "tree_structure.txt"), "w") as f:
            for v in range(1, 8):
                par = "N" if self.parent[v] is None else str(self.parent[v])
                ch = "N" if v not in self.children else "%d %d" % self.children[v]
                line = "%d\t%s\t%s" % (v, par, ch)
                if v in self.n_strains and self.n_strains[v] == 1:
                    line += "\t" + self.strain_name[(v, 1)]
                f.write(line + "\n")
        for path in (os.path.join(tdb, "hclsMap_95_recls.txt"), os.path.join(db_dir, "Cluster_Result", "hclsMap_95_recls.txt")):
            with open(path, "w") as f:
                for c in (1, 2, 3, 4):
                    names = ",".join(self.strain_name[(c, s)] for s in range(1, self.n_strains[c] + 1))
                    f.write("%d\t%d\t%s\n" % (c, self.n_strains[c], names))
        # ---- Kmer_Sets_L2 ---------------------------------------------------------------------
        for c in (1, 2, 4):
            cdir = os.path.join(db_dir, "Kmer_Sets_L2", "Kmer_Sets", "C%d" % c)
            os.makedirs(cdir, exist_ok=True)
            S = self.n_strains[c]
            snp_all = sorted(set(p for s in range(1, S + 1) for p in self.snps[(c, s)]))
            kmap = {}
            for p in snp_all:
                for w in range(p - K + 1, p + 1):
                    for s in range(1, S + 1):
                        km = self.strain_genome[(c, s)][w:w + K]
                        for x in (km, revcomp(km)):
                            kmap.setdefault(x, set()).add(s - 1)
            kid = {km: i + 1 for i, km in enumerate(kmap)}
            with open(os.path.join(cdir, "all_kmer.fasta"), "wb") as f:
                for km, i in kid.items():
                    f.write(b">%d\n" % i + km + b"\n")
            with open(os.path.join(cdir, "all_kid.pkl"), "wb") as f:
                pickle.dump({km.decode(): i for km, i in kid.items()}, f, pickle.HIGHEST_PROTOCOL)
            rows, cols = [], []
            for km, i in kid.items():
                for s in sorted(kmap[km]):
                    rows.append(i - 1)
                    cols.append(s)
            X = sp.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(len(kid), S), dtype=np.int8)
            sp.save_npz(os.path.join(cdir, "all_strains_re.npz"), X)
            with open(os.path.join(cdir, "id2strain_re.pkl"), "wb") as f:
                pickle.dump([self.strain_name[(c, s)] for s in range(1, S + 1)], f, pickle.HIGHEST_PROTOCOL)
            om = sp.csr_matrix((np.ones(len(kid), dtype=np.int8), (np.arange(len(kid)), np.full(len(kid), c - 1))),
                               shape=(len(kid), 4), dtype=np.int8)
            sp.save_npz(os.path.join(cdir, "overlap_matrix.npz"), om)
        return db_dir

    # ---- reads ----------------------------------------------------------------------------------
    def reads(self, mix, read_len=150, p_sub=0.005, seed=1):
        """mix: list of ((cluster, strain), depth).  Returns a list of read byte strings."""
        rng = np.random.default_rng(self.seed * 1000 + seed)
        out = []
        for (c, s), depth in mix:
            g = self.strain_genome[(c, s)]
            n = int(round(depth * len(g) / read_len))
            for _ in range(n):
                p = int(rng.integers(0, len(g) - read_len))
                r = bytearray(g[p:p + read_len])
                for j in np.nonzero(rng.random(read_len) < p_sub)[0]:
                    r[j] = b"ACGT"[int(rng.integers(0, 4))]
                r = bytes(r)
                out.append(revcomp(r) if rng.random() < 0.5 else r)
        order = rng.permutation(len(out))
        return [out[i] for i in order]


def write_fastq(path, reads, start=0, bgzf=False):
    data = b"".join(b"@syn.%d\n%s\n+\n%s\n" % (start + i, r, b"I" * len(r)) for i, r in enumerate(reads))
    if bgzf:                                  # blocked gzip: `zcat` reads it as concatenated members, the GPU path
        from tests import util                # inflates the members on the device
        with open(path, "wb") as f:
            f.write(util.bgzf_compress(data, block=20_000, level=6))
    elif path.endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=4) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)
    return path


# the parity cases: name -> (mix, StrainScan flags, paired/gz)
CASES = {
    "two_clusters_se": dict(mix=[((1, 2), 12), ((4, 1), 6)], flags=[], pe=False, gz=False),
    "two_clusters_pe_gz": dict(mix=[((1, 2), 12), ((4, 1), 6)], flags=[], pe=True, gz=True),
    "same_cluster_two_strains": dict(mix=[((1, 1), 10), ((1, 3), 8)], flags=[], pe=False, gz=False),
    "singleton_cluster": dict(mix=[((3, 1), 8)], flags=[], pe=False, gz=True),
    "low_depth_prob": dict(mix=[((2, 1), 0.6), ((4, 2), 0.4)], flags=["-l", "2", "-b", "1"], pe=False, gz=False),
    "extra_region": dict(mix=[((1, 2), 10), ((2, 2), 5)], flags=["-e", "1"], pe=False, gz=False),
    "low_mem_db": dict(mix=[((1, 3), 14), ((2, 1), 9)], flags=[], pe=False, gz=False, low_mem=True),
    "two_strains_pe_bgzf": dict(mix=[((4, 1), 9), ((4, 2), 7), ((1, 1), 5)], flags=[], pe=True, gz=True, bgzf=True),
    # plasmid_mode (StrainScan.py:225-266): the run-time database DB_plasmid is the pre-made PLASMID_DB below
    # (baseline/run_pipeline.py --plasmid-db stands in for the StrainScan_build.py call); `pmix` = reads drawn
    # from ITS strains, added to the sample.  -p 2: reference genomes given by -r; -p 1: short contigs of the
    # strains of the identified multi-strain clusters (load_db_cls, StrainScan.py:47-94), default_cov = 0.
    # a reconstructed node whose k-mer list overlaps an identified cluster's (adjust_profile, identify.py:167-191)
    "reconstructed_node": dict(mix=[((1, 2), 10), ((4, 1), 8)], flags=[], pe=False, gz=False, db="overlap"),
    "plasmid_mode_2": dict(mix=[((1, 2), 10)], pmix=[((2, 1), 12), ((4, 2), 7)], flags=["-p", "2"], pe=False, gz=False),
    "plasmid_mode_1": dict(mix=[((1, 1), 9), ((2, 2), 6)], pmix=[((1, 3), 10), ((1, 1), 5)], flags=["-p", "1"], pe=True,
                           gz=True),
}


# read-sampling seeds, fixed per case (the committed golden reports depend on them)
CASE_SEEDS = {"extra_region": 1, "low_depth_prob": 2, "same_cluster_two_strains": 3, "singleton_cluster": 4,
              "two_clusters_pe_gz": 5, "two_clusters_se": 6, "low_mem_db": 7, "two_strains_pe_bgzf": 8,
              "plasmid_mode_2": 9, "plasmid_mode_1": 10, "reconstructed_node": 11}


def plasmid_db():
    """The database the plasmid_mode cases search: same layout, other sequences (shorter replicons)."""
    return SynthDB(seed=20260331, seg_len=1500, n_seg=12, snp_rate=0.006, prefix="PLS")


def make_case_inputs(db, name, out_dir):
    case = CASES[name]
    reads = db.reads(case["mix"], seed=CASE_SEEDS[name])
    extra = []
    if "pmix" in case:
        pdb = plasmid_db()
        more = pdb.reads(case["pmix"], seed=CASE_SEEDS[name])
        rng = np.random.default_rng(CASE_SEEDS[name])
        reads = reads + more
        reads = [reads[i] for i in rng.permutation(len(reads))]
        rdir = os.path.join(out_dir, name + "_refs")      # -r: one FASTA per strain (two contigs each)
        os.makedirs(rdir, exist_ok=True)
        src = db if case["flags"][1] == "1" else pdb
        for (c, s_), nm in src.strain_name.items():
            g = src.strain_genome[(c, s_)]
            with open(os.path.join(rdir, nm + ".fna"), "wb") as f:
                h = len(g) // 2
                f.write(b">" + nm.encode() + b"_ctg1\n" + g[:h] + b"\n>" + nm.encode() + b"_ctg2\n" + g[h:] + b"\n")
        extra = ["-r", rdir]
    ext = ".fq.gz" if case["gz"] else ".fq"
    bgzf = case.get("bgzf", False)
    if case["pe"]:
        h = len(reads) // 2
        a = write_fastq(os.path.join(out_dir, name + "_1" + ext), reads[:h], bgzf=bgzf)
        b = write_fastq(os.path.join(out_dir, name + "_2" + ext), reads[h:], start=h, bgzf=bgzf)
        return ["-i", a, "-j", b] + extra
    return ["-i", write_fastq(os.path.join(out_dir, name + ext), reads, bgzf=bgzf)] + extra


def write_plasmid_db(base_dir):
    """The pre-made DB_plasmid of the plasmid_mode cases (run_pipeline.py --plasmid-db)."""
    return plasmid_db().write(os.path.join(base_dir, "DB_plasmid_src"))


def write_dbs(base_dir):
    """Both database layouts the cases use: {low_mem flag: DB directory}."""
    db = SynthDB()
    return {False: db.write(os.path.join(base_dir, "DB")), True: db.write(os.path.join(base_dir, "DB_mem"), low_mem=True),
            "overlap": db.write(os.path.join(base_dir, "DB_ovl"), overlap=True)}
