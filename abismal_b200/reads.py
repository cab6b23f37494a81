"""FASTQ loading with the reference's ReadLoader rules (src/abismal.cpp:164-201)
into the flat arena + offsets layout of abg_batch (tests / bench only)."""
import gzip

import numpy as np

MIN_READ_LENGTH = 44


class ReadBatch:
    def __init__(self, names, seqs):
        self.names = names
        self.n = len(seqs)
        lens = np.fromiter((len(s) for s in seqs), np.int64, self.n)
        self.off = np.zeros(self.n + 1, np.uint32)
        np.cumsum(lens, out=self.off[1:])
        self.seq = np.frombuffer(("".join(seqs) or "\0").encode(), np.uint8).copy()
        self.max_len = int(lens.max()) if self.n else 0

    def slice(self, lo, hi):
        b = ReadBatch.__new__(ReadBatch)
        b.names = self.names[lo:hi] if self.names is not None else None
        b.n = hi - lo
        b.off = (self.off[lo:hi + 1] - self.off[lo]).astype(np.uint32)
        b.seq = self.seq[int(self.off[lo]):max(int(self.off[hi]), int(self.off[lo]) + 1)].copy()
        b.max_len = int(np.diff(self.off[lo:hi + 1].astype(np.int64)).max()) if hi > lo else 0
        return b

    def to_pinned(self):
        """Copy of this batch whose arrays live in page-locked memory (abg_host_alloc)."""
        from .capi import pinned_zeros
        b = ReadBatch.__new__(ReadBatch)
        b.names, b.n, b.max_len = self.names, self.n, self.max_len
        b.off = pinned_zeros(self.off.shape, np.uint32)
        b.off[...] = self.off
        b.seq = pinned_zeros(self.seq.shape, np.uint8)
        b.seq[...] = self.seq
        return b

    def sequence(self, i):
        return self.seq[int(self.off[i]):int(self.off[i + 1])].tobytes().decode()

    @property
    def h2d_bytes(self):
        return int(self.off[-1]) + self.off.nbytes


def trim_read(line, min_read_length=MIN_READ_LENGTH):
    if sum(1 for c in line if c != "N") < min_read_length:
        return ""
    line = line.rstrip("N")
    for i, c in enumerate(line):
        if c in "ACGT":
            return line[i:]
    raise ValueError("read without A/C/G/T")


def load_fastq(path, limit=None, min_read_length=MIN_READ_LENGTH):
    opener = gzip.open if path.endswith(".gz") else open
    names, seqs = [], []
    with opener(path, "rt") as f:
        for k, line in enumerate(f):
            line = line.rstrip("\n").rstrip("\r")
            if k % 4 == 0:
                ws = min([p for p in (line.find(" "), line.find("\t")) if p >= 0], default=len(line))
                names.append(line[1:ws])
            elif k % 4 == 1:
                seqs.append(trim_read(line, min_read_length))
                if limit is not None and len(seqs) >= limit:
                    break
    return ReadBatch(names[:len(seqs)], seqs)
