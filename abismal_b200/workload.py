"""Synthetic benchmark workload: an i.i.d.
uniform ACGT genome, its AbismalIndex (built on the GPU, byte-compatible with
`abismal idx`), and bisulfite reads simulated by the reference's own `sim`.
Everything is cached under a scratch directory so that the 1/2/4/8-GPU runs
of one session reuse it.  Used by bench.py only."""
import os
import subprocess
import time

import numpy as np

from . import index_build
from .reads import ReadBatch

N_CHROMS = 24
LINE = 100


def cache_dir():
    d = os.environ.get("ABISMAL_B200_CACHE", "/tmp/abismal_b200_bench")
    os.makedirs(d, exist_ok=True)
    return d


def _chrom_len(genome_bases):
    per = genome_bases // N_CHROMS
    return per - per % LINE


def generate_genome(genome_bases, seed, fasta_path=None, device=0):
    """-> PreparedGenome-like object (names, starts, words, exclude, genome_size).
    Bases are drawn on the GPU (torch, plumbing only); optionally also written as FASTA."""
    import torch
    per = _chrom_len(genome_bases)
    n = per * N_CHROMS
    dev = torch.device("cuda", device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    x = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev, generator=gen)
    pad = index_build.PADDING
    total = n + 2 * pad
    codes = torch.zeros(total + (total & 1), dtype=torch.uint8, device=dev)
    codes[pad:pad + n] = torch.bitwise_left_shift(torch.ones_like(x), x)  # A1 C2 G4 T8
    packed = (codes[0::2] | (codes[1::2] << 4)).contiguous()
    n_words = (total + 15) // 16
    words = np.zeros(n_words, "<u8")
    pb = packed.cpu().numpy()
    words.view(np.uint8)[:pb.size] = pb
    del codes, packed

    class G:
        pass
    g = G()
    g.names = ["pad_start"] + ["chr%d" % (i + 1) for i in range(N_CHROMS)] + ["pad_end"]
    starts = [0] + [pad + i * per for i in range(N_CHROMS)] + [pad + n, total]
    g.starts = np.array(starts, "<u4")
    g.genome_size = total
    g.words = words
    g.exclude = np.array([[0, pad], [pad + n, total]], "<u8")
    if fasta_path is not None:
        lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
        tmp = fasta_path + ".tmp"
        with open(tmp, "wb") as f:
            for c in range(N_CHROMS):
                asc = lut[x[c * per:(c + 1) * per].long()].view(-1, LINE)
                nl = torch.full((asc.shape[0], 1), 10, dtype=torch.uint8, device=dev)
                f.write((">chr%d\n" % (c + 1)).encode())
                f.write(torch.cat([asc, nl], dim=1).cpu().numpy().tobytes())
        os.replace(tmp, fasta_path)
    del x
    torch.cuda.empty_cache()
    return g


def get_index(genome_bases, seed, device=0, need_files=True, log=print):
    """Build (or load from the cache) the index of the synthetic genome.
    -> (index arrays object usable with capi.Index, paths dict)"""
    from .index_file import IndexFile
    key = "g%d_s%d" % (genome_bases, seed)
    d = os.path.join(cache_dir(), key)
    os.makedirs(d, exist_ok=True)
    paths = {"dir": d, "fasta": os.path.join(d, "genome.fa"), "index": os.path.join(d, "genome.idx")}
    if os.path.exists(paths["index"]) and os.path.exists(paths["fasta"]):
        t = time.time()
        ix = IndexFile(paths["index"])
        log("index loaded from cache %s in %.1fs" % (paths["index"], time.time() - t))
        return ix, paths
    t = time.time()
    g = generate_genome(genome_bases, seed, paths["fasta"] if need_files else None, device)
    log("genome of %d bases generated in %.1fs" % (g.genome_size, time.time() - t))
    t = time.time()
    built = index_build.BuiltIndex(g, device)
    log("index built on GPU in %.1fs (two-letter entries %d, three-letter %d)"
        % (time.time() - t, built.index_size, built.index_size_three))
    if need_files:
        t = time.time()
        tmp = paths["index"] + ".tmp"
        built.write(tmp)
        os.replace(tmp, paths["index"])
        log("index file written in %.1fs" % (time.time() - t))
    return built, paths


def simulate_reads(sim_bin, fasta, out_prefix, n_pairs, seed, paired=True, mode_flag=None, read_len=150,
                   n_procs=4, log=print):
    """Run the reference's `sim` in n_procs processes (distinct seeds) and
    concatenate.  -> (fq1, fq2 or None)"""
    fq1, fq2 = out_prefix + "_1.fq", out_prefix + "_2.fq"
    if os.path.exists(fq1) and (not paired or os.path.exists(fq2)):
        return fq1, (fq2 if paired else None)
    t = time.time()
    per = (n_pairs + n_procs - 1) // n_procs
    procs = []
    for p in range(n_procs):
        n = min(per, n_pairs - p * per)
        if n <= 0:
            break
        cmd = [sim_bin, "sim", "-seed", str(seed * 1000 + p), "-l", str(read_len), "-min-fraglen", str(read_len),
               "-max-fraglen", "400", "-n", str(n), "-m", "0.01", "-b", "0.98", "-o", "%s.part%d" % (out_prefix, p)]
        if not paired:
            cmd.append("-single")
        if mode_flag:
            cmd.append(mode_flag)
        cmd.append(fasta)
        procs.append((p, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)))
    for p, pr in procs:
        _, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("sim failed: " + err.decode()[-500:])
    for end, dst in ((1, fq1), (2, fq2)):
        if end == 2 and not paired:
            continue
        with open(dst + ".tmp", "wb") as fo:
            for p, _ in procs:
                part = "%s.part%d_%d.fq" % (out_prefix, p, end)
                with open(part, "rb") as fi:
                    # read names must stay unique across parts
                    data = fi.read()
                fo.write(data.replace(b"@read", b"@p%dread" % p))
                os.remove(part)
        os.replace(dst + ".tmp", dst)
    log("simulated %d %s with %d sim processes in %.1fs" % (n_pairs, "pairs" if paired else "reads", len(procs),
                                                            time.time() - t))
    return fq1, (fq2 if paired else None)


def load_fastq_fast(path):
    """Vectorised FASTQ -> ReadBatch for simulator output (4-line records).
    Applies the ReadLoader rules when no read contains N (checked); falls back
    to the general loader otherwise."""
    buf = np.fromfile(path, np.uint8)
    nl = np.flatnonzero(buf == 10)
    if nl.size % 4 != 0:
        raise ValueError("FASTQ with a partial record: " + path)
    starts = np.concatenate(([0], nl[:-1] + 1))
    seq_s, seq_e = starts[1::4], nl[1::4]
    lens = (seq_e - seq_s).astype(np.int64)
    n = lens.size
    off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    # gather all sequence bytes
    idx = np.repeat(seq_s - off[:-1], lens) + np.arange(off[-1])
    seq = buf[idx]
    if (seq == ord("N")).any() or lens.min() < 44 or off[-1] >= 2 ** 32:
        from .reads import load_fastq
        return load_fastq(path)
    b = ReadBatch.__new__(ReadBatch)
    b.names = None
    b.n = int(n)
    b.off = off.astype(np.uint32)
    b.seq = seq
    b.max_len = int(lens.max())
    return b
