"""`abismal idx` on the GPU: host-side genome preparation (the reference's
load_genome / contiguous_n / replace_included_n / encode_dna_four_bit,
src/AbismalIndex.cpp:125-175, :1322-1360) around abg_build_index, plus a writer
of the on-disk AbismalIndex format (src/AbismalIndex.cpp:1037-1072)."""
import ctypes as C
import gzip
import struct

import numpy as np

from . import capi

PADDING = 32767
MAX_N_COUNT = 256

# dna_four_bit_encoding (src/dna_four_bit_bisulfite.hpp:156-165)
_ENC = np.zeros(256, np.uint8)
for _ch, _v in dict(A=1, B=14, C=2, D=13, G=4, H=11, K=12, M=3, R=5, S=6, T=8, V=7, W=9, Y=10).items():
    _ENC[ord(_ch)] = _v
    _ENC[ord(_ch.lower())] = _v


class abg_built_index(C.Structure):
    _fields_ = [
        ("counter", C.POINTER(C.c_uint32)), ("counter_t", C.POINTER(C.c_uint32)), ("counter_a", C.POINTER(C.c_uint32)),
        ("index", C.POINTER(C.c_uint32)), ("index_t", C.POINTER(C.c_uint32)), ("index_a", C.POINTER(C.c_uint32)),
        ("counter_size", C.c_uint64), ("counter_size_three", C.c_uint64), ("index_size", C.c_uint64),
        ("index_size_three", C.c_uint64), ("max_candidates", C.c_uint32), ("reserved", C.c_uint32),
    ]


def read_fasta(path):
    """-> (names, list of uint8 arrays), names cut at the first blank (load_genome :1346)."""
    opener = gzip.open if path.endswith(".gz") else open
    names, chunks, cur = [], [], []
    with opener(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\n").rstrip(b"\r")
            if line[:1] == b">":
                if names:
                    chunks.append(np.frombuffer(b"".join(cur), np.uint8))
                cur = []
                hdr = line[1:].decode()
                for k, ch in enumerate(hdr):
                    if ch in " \t":
                        hdr = hdr[:k]
                        break
                names.append(hdr)
            else:
                cur.append(line)
    if names:
        chunks.append(np.frombuffer(b"".join(cur), np.uint8))
    return names, chunks


def contiguous_n(genome):
    """[first, second) runs of 'N' (contiguous_n :125-145)."""
    is_n = genome == ord("N")
    d = np.diff(np.concatenate(([0], is_n.view(np.int8), [0])))
    return np.stack([np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]], axis=1).astype(np.uint64)


def lcg_bases(n):
    """random_base_generator (src/AbismalIndex.hpp:39-61): x0 = 1."""
    out = np.empty(n, np.uint8)
    x = 1
    acgt = b"ACGT"
    for i in range(n):
        x = (1103515245 * x + 12345) & 0x7FFFFFFF
        out[i] = acgt[x & 3]
    return out


def pack_four_bit(codes):
    """16 nibbles per little-endian uint64 word, base j at bits 4j..4j+3."""
    n = codes.size
    n_words = (n + 15) // 16
    padded = np.zeros(n_words * 16, np.uint8)
    padded[:n] = codes
    b = padded.reshape(-1, 2)
    packed_bytes = (b[:, 0] | (b[:, 1] << 4)).astype(np.uint8)
    return packed_bytes.view("<u8").copy()


class PreparedGenome:
    """names (with pad_start/pad_end), starts, 4-bit words, exclude intervals."""

    def __init__(self, names, seqs):
        self.names = ["pad_start"] + list(names) + ["pad_end"]
        starts = [0]
        pos = PADDING
        for s in seqs:
            starts.append(pos)
            pos += s.size
        starts.append(pos)          # pad_end
        pos += PADDING
        starts.append(pos)          # end of everything
        self.starts = np.array(starts, "<u4")
        self.genome_size = pos
        g = np.full(pos, ord("N"), np.uint8)
        off = PADDING
        for s in seqs:
            g[off:off + s.size] = s
            off += s.size
        runs = contiguous_n(g)
        self.exclude = runs[(runs[:, 1] - runs[:, 0]) > MAX_N_COUNT].copy()
        # replace_included_n (:164-175): Ns outside the excluded runs, in genome order
        mask = g == ord("N")
        for a, b in self.exclude:
            mask[int(a):int(b)] = False
        where = np.nonzero(mask)[0]
        if where.size:
            g[where] = lcg_bases(where.size)
        self.words = pack_four_bit(_ENC[g])


def prepare_fasta(path):
    names, seqs = read_fasta(path)
    if not names:
        raise ValueError("no names found in genome file")
    return PreparedGenome(names, seqs)


class BuiltIndex:
    """Result of abg_build_index as numpy arrays + everything IndexFile exposes."""

    def __init__(self, prepared, device=0, window_size=20):
        lib = capi.load_library()
        self.window_size = window_size
        lib.abg_index_build_last_error.restype = C.c_char_p
        lib.abg_build_index_w.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                          C.POINTER(abg_built_index)]
        lib.abg_built_index_free.argtypes = [C.POINTER(abg_built_index)]
        lib.abg_built_index_free.restype = None
        out = abg_built_index()
        ex = np.ascontiguousarray(prepared.exclude, "<u8")
        rc = lib.abg_build_index_w(prepared.words.ctypes.data_as(C.c_void_p), prepared.genome_size,
                                   ex.ctypes.data_as(C.c_void_p), ex.shape[0], window_size, device, C.byref(out))
        if rc != 0:
            raise capi.AbgError(lib.abg_index_build_last_error().decode())
        try:
            def arr(p, n):
                return np.ctypeslib.as_array(p, shape=(max(int(n), 1),))[:int(n)].copy()
            self.counter_size = int(out.counter_size)
            self.counter_size_three = int(out.counter_size_three)
            self.index_size = int(out.index_size)
            self.index_size_three = int(out.index_size_three)
            self.max_candidates = int(out.max_candidates)
            self.counter = arr(out.counter, self.counter_size + 1)
            self.counter_t = arr(out.counter_t, self.counter_size_three + 1)
            self.counter_a = arr(out.counter_a, self.counter_size_three + 1)
            self.index = arr(out.index, self.index_size)
            self.index_t = arr(out.index_t, self.index_size_three)
            self.index_a = arr(out.index_a, self.index_size_three)
        finally:
            lib.abg_built_index_free(C.byref(out))
        self.names = prepared.names
        self.starts = prepared.starts
        self.genome_size = prepared.genome_size
        n_words = (self.genome_size + 15) // 16
        self.genome = np.zeros(n_words + 1, "<u8")
        self.genome[:n_words] = prepared.words[:n_words]
        for k in ("index", "index_t", "index_a"):
            if getattr(self, k).size == 0:
                setattr(self, k, np.zeros(1, "<u4"))

    def write(self, path):
        n_words = (self.genome_size + 15) // 16
        with open(path, "wb") as f:
            f.write(b"AbismalIndex")
            f.write(struct.pack("<3I", 25, self.window_size, 256))
            f.write(struct.pack("<I", len(self.names)))
            for nm in self.names:
                b = nm.encode()
                f.write(struct.pack("<I", len(b)))
                f.write(b)
            self.starts.astype("<u4").tofile(f)
            self.genome[:n_words].tofile(f)
            f.write(struct.pack("<I", self.max_candidates))
            f.write(struct.pack("<4Q", self.counter_size, self.counter_size_three, self.index_size,
                                self.index_size_three))
            self.counter.tofile(f)
            self.counter_t.tofile(f)
            self.counter_a.tofile(f)
            self.index[:self.index_size].tofile(f)
            self.index_t[:self.index_size_three].tofile(f)
            self.index_a[:self.index_size_three].tofile(f)


def build_index_file(fasta_path, index_path, device=0, window_size=20):
    built = BuiltIndex(prepare_fasta(fasta_path), device, window_size)
    built.write(index_path)
    return built
