"""Seeded synthetic genomes / taxonomy / reads for the parity tests.

Pure numpy; no oracle, no product code.  Everything is a deterministic
function of the seed so that golden fixtures can be regenerated.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, np.uint8)
COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[a] = b


def random_genome(rng: np.random.Generator, n: int) -> np.ndarray:
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(seq: np.ndarray) -> np.ndarray:
    return COMP[seq[::-1]]


# (ext_id, parent_ext_id, name, rank) — NCBI-like lineage, depth >= 5, two
# species sharing a genus so that build-time LCA values appear in the table.
TAXONOMY_CFG1 = [
    (1, 1, "root", "no rank"),
    (131567, 1, "cellular organisms", "no rank"),
    (2759, 131567, "Eukaryota", "superkingdom"),
    (7711, 2759, "Chordata", "phylum"),
    (40674, 7711, "Mammalia", "class"),
    (9443, 40674, "Primates", "order"),
    (9604, 9443, "Hominidae", "family"),
    (9605, 9604, "Homo", "genus"),
    (9606, 9605, "Homo sapiens", "species"),
    (2, 131567, "Bacteria", "superkingdom"),
    (1224, 2, "Pseudomonadota", "phylum"),
    (1236, 1224, "Gammaproteobacteria", "class"),
    (91347, 1236, "Enterobacterales", "order"),
    (543, 91347, "Enterobacteriaceae", "family"),
    (561, 543, "Escherichia", "genus"),
    (562, 561, "Escherichia coli", "species"),
    (564, 561, "Escherichia fergusonii", "species"),
    (1239, 2, "Bacillota", "phylum"),
    (91061, 1239, "Bacilli", "class"),
    (1385, 91061, "Bacillales", "order"),
    (186817, 1385, "Bacillaceae", "family"),
    (1386, 186817, "Bacillus", "genus"),
    (1423, 1386, "Bacillus subtilis", "species"),
]


def cfg1_genomes(seed: int = 1, scale: float = 1.0):
    """Synthetic human chr (5 Mbp) + 3 bacterial genomes (2/3/4 Mbp) at scale=1;
    the two Escherichia share a common segment (200 kb at scale=1)."""
    rng = np.random.default_rng(seed)
    n = lambda x: max(2000, int(x * scale))
    human = random_genome(rng, n(5_000_000))
    ecoli = random_genome(rng, n(2_000_000))
    eferg = random_genome(rng, n(3_000_000))
    shared = random_genome(rng, n(200_000))
    ecoli[len(ecoli) // 4: len(ecoli) // 4 + len(shared)] = shared
    eferg[len(eferg) // 2: len(eferg) // 2 + len(shared)] = shared
    bsub = random_genome(rng, n(4_000_000))
    return [(9606, human), (562, ecoli), (564, eferg), (1423, bsub)]


def mutate(rng, seq: np.ndarray, sub_rate: float) -> np.ndarray:
    if sub_rate <= 0:
        return seq
    seq = seq.copy()
    m = rng.random(len(seq)) < sub_rate
    k = int(m.sum())
    if k:
        seq[m] = ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]
    return seq


def illumina_reads(genomes, n_reads: int, read_len: int = 150, seed: int = 2,
                   fractions=(0.5, 0.25, 0.25), sub_rate: float = 0.005,
                   n_rate: float = 0.01, paired: bool = False, insert_mean: float = 350,
                   insert_sd: float = 50):
    """Reads: fractions = (human, bacterial, random).  genomes[0] is 'human'.
    Returns list of sequences (np.uint8 arrays); paired -> mates interleaved."""
    rng = np.random.default_rng(seed)
    out = []
    src = rng.choice(3, size=n_reads, p=np.array(fractions) / sum(fractions))
    for i in range(n_reads):
        if src[i] == 2:
            frag_len = read_len if not paired else max(read_len, int(rng.normal(insert_mean, insert_sd)))
            frag = random_genome(rng, frag_len)
        else:
            g = genomes[0][1] if src[i] == 0 else genomes[1 + rng.integers(0, len(genomes) - 1)][1]
            frag_len = read_len if not paired else max(read_len, int(rng.normal(insert_mean, insert_sd)))
            frag_len = min(frag_len, len(g))
            p = rng.integers(0, len(g) - frag_len + 1)
            frag = g[p:p + frag_len]
            if rng.random() < 0.5:
                frag = revcomp(frag)
        mates = [frag[:read_len]] if not paired else [frag[:read_len], revcomp(frag)[:read_len]]
        for m in mates:
            m = mutate(rng, m, sub_rate)
            if n_rate > 0 and rng.random() < n_rate:
                m = m.copy()
                m[rng.integers(0, len(m))] = ord("N")
            out.append(m)
    return out


def ont_reads(genomes, n_reads: int, seed: int = 4, n50: float = 10_000, err: float = 0.05,
              min_len: int = 200, max_len: int = 100_000, human_frac: float = 0.5):
    """Long reads, log-normal length, err split sub:ins:del 1:1:1."""
    rng = np.random.default_rng(seed)
    sigma = 0.8
    mu = np.log(n50) - sigma ** 2  # rough N50 tuning
    out = []
    for _ in range(n_reads):
        L = int(np.clip(rng.lognormal(mu, sigma), min_len, max_len))
        if rng.random() < human_frac:
            g = genomes[0][1]
        else:
            g = genomes[1 + rng.integers(0, len(genomes) - 1)][1]
        L = min(L, len(g))
        p = rng.integers(0, len(g) - L + 1)
        frag = g[p:p + L]
        if rng.random() < 0.5:
            frag = revcomp(frag)
        r = rng.random(L)
        sub = r < err / 3
        ins = (r >= err / 3) & (r < 2 * err / 3)
        dele = (r >= 2 * err / 3) & (r < err)
        frag = frag.copy()
        frag[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
        keep = ~dele
        reps = np.where(ins, 2, 1)[keep]
        frag = np.repeat(frag[keep], reps)
        out.append(frag)
    return out


def pack(seqs):
    """Concatenate -> (bases uint8, offsets uint64[n+1])."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, np.uint64)
    np.cumsum(lens, out=offsets[1:])
    bases = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases, dtype=np.uint8), offsets


def write_fastq(path, seqs, prefix="r", paired_suffix=None, qual=b"I"):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            name = f"@{prefix}{i}" + (paired_suffix or "")
            f.write(name.encode() + b"\n" + bytes(s) + b"\n+\n" + qual * len(s) + b"\n")
