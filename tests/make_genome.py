"""Deterministic synthetic genomes for parity tests (no reference data needed).

`repeat_genome` is built to hit the hard paths of the mapper: duplicated
segments with small divergence (ambiguous mappers, many-to-many concordant
pairs), microsatellites and homopolymers (seed buckets > max_candidates, so
the binary-search narrowing runs and PE candidate heaps grow), N runs of both
kinds (<=256: replaced by random bases at index time; longer: stay N), IUPAC
codes, and several chromosomes (reads crossing chromosome ends).
`repeat_genome(iupac=False)` is the same genome without the IUPAC codes (real
assemblies hold few or none): no compare window there takes the exact 4-bit route.
"""
import numpy as np


def _rand_seq(rng, n):
    return "".join(np.array(list("ACGT"))[rng.integers(0, 4, size=n)])


def _mutate(rng, s, rate):
    a = np.array(list(s))
    m = rng.random(len(a)) < rate
    a[m] = np.array(list("ACGT"))[rng.integers(0, 4, size=int(m.sum()))]
    return "".join(a)


def repeat_genome(seed=7, scale=1.0, iupac=True):
    rng = np.random.default_rng(seed)
    chroms = []
    unit = _rand_seq(rng, int(30000 * scale))
    # chrA: random + diverged copies of one segment
    parts = [_rand_seq(rng, int(400000 * scale))]
    for div in (0.0, 0.002, 0.01, 0.03):
        parts.append(_mutate(rng, unit, div))
        parts.append(_rand_seq(rng, int(20000 * scale)))
    chroms.append(("chrA", "".join(parts)))
    # chrB: low complexity
    parts = [_rand_seq(rng, int(200000 * scale))]
    parts.append("AC" * 4000)
    parts.append(_rand_seq(rng, 5000))
    parts.append("A" * 3000)
    parts.append(_rand_seq(rng, 5000))
    parts.append("TTAGGG" * 1500)
    parts.append(_rand_seq(rng, 5000))
    parts.append(_mutate(rng, "CAG" * 3000, 0.02))
    parts.append(_rand_seq(rng, int(100000 * scale)))
    parts.append(_mutate(rng, unit, 0.005))
    chroms.append(("chrB", "".join(parts)))
    # chrC: N runs and IUPAC codes
    parts = [_rand_seq(rng, 50000), "N" * 100, _rand_seq(rng, 30000), "N" * 2000, _rand_seq(rng, 40000)]
    s = list("".join(parts))
    for p in rng.integers(0, len(s), size=40):  # (drawn in both variants: the rest of the genome stays the same)
        code = "RYKMSW"[int(rng.integers(0, 6))]
        if s[p] != "N" and iupac:
            s[p] = code
    chroms.append(("chrC", "".join(s)))
    # a few short chromosomes (reads run off their ends)
    for k in range(4):
        chroms.append(("chrS%d description text" % k, _rand_seq(rng, 700 + 300 * k)))
    return chroms


def random_genome(n_bases, n_chroms=4, seed=11):
    rng = np.random.default_rng(seed)
    per = n_bases // n_chroms
    return [("chr%d" % (i + 1), _rand_seq(rng, per)) for i in range(n_chroms)]


def write_fasta(chroms, path, width=80):
    with open(path, "w") as f:
        for name, seq in chroms:
            f.write(">%s\n" % name)
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width])
                f.write("\n")


if __name__ == "__main__":
    import sys
    kind, out = sys.argv[1], sys.argv[2]
    if kind == "repeat":
        write_fasta(repeat_genome(), out)
    else:
        write_fasta(random_genome(int(sys.argv[3])), out)
