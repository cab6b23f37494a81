"""Reader of the on-disk AbismalIndex format (src/AbismalIndex.cpp:1037-1146 of
the reference): numpy views for the ctypes binding, used by tests and bench.py.
The C++ front end has its own reader (csrc/host/index_file.cpp)."""
import struct

import numpy as np


class IndexFile:
    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            if f.read(12) != b"AbismalIndex":
                raise ValueError("index file format problem: " + path)
            kw, ws, nsp = struct.unpack("<3I", f.read(12))
            # window_size 12 = an index of the reference configured with --enable-short (AbismalIndex.hpp:73-77)
            if (kw, nsp) != (25, 256) or ws not in (12, 20):
                raise ValueError("inconsistent seed parameters in " + path)
            self.window_size = ws
            (n_chroms,) = struct.unpack("<I", f.read(4))
            self.names = []
            for _ in range(n_chroms):
                (ln,) = struct.unpack("<I", f.read(4))
                self.names.append(f.read(ln).decode())
            self.starts = np.frombuffer(f.read(4 * (n_chroms + 1)), "<u4").copy()
            self.genome_size = int(self.starts[-1])
            n_words = (self.genome_size + 15) // 16
            # one spare zero word for the compare's look-ahead at the very end
            self.genome = np.zeros(n_words + 1, "<u8")
            self.genome[:n_words] = np.fromfile(f, "<u8", n_words)
            (self.max_candidates,) = struct.unpack("<I", f.read(4))
            (self.counter_size, self.counter_size_three, self.index_size,
             self.index_size_three) = struct.unpack("<4Q", f.read(32))
            self.counter = np.fromfile(f, "<u4", self.counter_size + 1)
            self.counter_t = np.fromfile(f, "<u4", self.counter_size_three + 1)
            self.counter_a = np.fromfile(f, "<u4", self.counter_size_three + 1)
            self.index = np.fromfile(f, "<u4", self.index_size)
            self.index_t = np.fromfile(f, "<u4", self.index_size_three)
            self.index_a = np.fromfile(f, "<u4", self.index_size_three)
            if self.index_a.size != self.index_size_three:
                raise ValueError("failed loading index file")
        # ctypes needs non-empty buffers to take an address
        for k in ("index", "index_t", "index_a"):
            if getattr(self, k).size == 0:
                setattr(self, k, np.zeros(1, "<u4"))
