"""ctypes binding of include/abismal_b200.h (and, for tests, of the oracle's
identical data contract in oracle/abismal_oracle.h)."""
import ctypes as C
import os
import weakref

import numpy as np

MODE_PAIRED = 1
MODE_A_RICH = 2
MODE_RANDOM_PBAT = 4

FLAG_RC = 0x10
FLAG_AMBIG = 0x100
FLAG_A_RICH = 0x1000

FEATURE_SEED_CONTEXT = 1
FEATURE_COMPACT_COUNTERS = 2
FEATURE_GENOME_HAS_IUPAC = 4

HIT_DTYPE = np.dtype([("diffs", "<i2"), ("flags", "<u2"), ("pos", "<u4")])

_HERE = os.path.dirname(os.path.abspath(__file__))


class AbgError(RuntimeError):
    pass


class abg_index_view(C.Structure):
    _fields_ = [
        ("genome", C.c_void_p), ("genome_words", C.c_uint64), ("genome_size", C.c_uint64),
        ("counter", C.c_void_p), ("counter_size", C.c_uint64),
        ("counter_t", C.c_void_p), ("counter_a", C.c_void_p), ("counter_size_three", C.c_uint64),
        ("index", C.c_void_p), ("index_size", C.c_uint64),
        ("index_t", C.c_void_p), ("index_a", C.c_void_p), ("index_size_three", C.c_uint64),
        ("max_candidates", C.c_uint32), ("window_size", C.c_uint32),
    ]


class abg_params(C.Structure):
    _fields_ = [
        ("mode", C.c_uint32), ("allow_ambig", C.c_uint32), ("min_dist", C.c_uint32), ("max_dist", C.c_uint32),
        ("valid_frac", C.c_double), ("max_candidates", C.c_uint32), ("cigar_stride", C.c_uint32),
    ]


class abg_batch(C.Structure):
    _fields_ = [
        ("n", C.c_uint32), ("reserved", C.c_uint32),
        ("seq1", C.c_void_p), ("off1", C.c_void_p), ("seq2", C.c_void_p), ("off2", C.c_void_p),
    ]


class abg_results(C.Structure):
    _fields_ = [
        ("pe_r1", C.c_void_p), ("pe_r2", C.c_void_p), ("se1", C.c_void_p), ("se2", C.c_void_p),
        ("cigar1", C.c_void_p), ("cigar2", C.c_void_p), ("n_cigar1", C.c_void_p), ("n_cigar2", C.c_void_p),
    ]


class abg_work_counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_lookup", "n_entry", "n_cmp", "n_word", "n_align", "n_dpref")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def lib_path():
    # ABISMAL_B200_LIB: a differently tuned build of the same library (kernel tuning experiments only)
    return os.environ.get("ABISMAL_B200_LIB") or os.path.join(_HERE, "libabismal_b200.so")


_lib = None


def load_library():
    """Load libabismal_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise AbgError("%s not found: build it with `make -C abismal_b200/csrc` "
                       "(or __graft_entry__.build()); there is no CPU fallback" % p)
    lib = C.CDLL(p)
    lib.abg_last_error.restype = C.c_char_p
    lib.abg_device_count.restype = C.c_int
    lib.abg_index_create.argtypes = [C.POINTER(abg_index_view), C.c_int, C.POINTER(C.c_void_p)]
    lib.abg_index_destroy.argtypes = [C.c_void_p]
    lib.abg_index_destroy.restype = None
    lib.abg_index_device_bytes.argtypes = [C.c_void_p]
    lib.abg_index_device_bytes.restype = C.c_uint64
    lib.abg_index_features.argtypes = [C.c_void_p]
    lib.abg_index_features.restype = C.c_uint32
    lib.abg_mapper_create.argtypes = [C.c_void_p, C.POINTER(abg_params), C.c_uint32, C.c_uint32, C.c_int,
                                      C.POINTER(C.c_void_p)]
    lib.abg_mapper_destroy.argtypes = [C.c_void_p]
    lib.abg_mapper_destroy.restype = None
    lib.abg_map_batch.argtypes = [C.c_void_p, C.POINTER(abg_batch), C.POINTER(abg_results)]
    lib.abg_mapper_upload.argtypes = [C.c_void_p, C.POINTER(abg_batch)]
    lib.abg_mapper_run.argtypes = [C.c_void_p]
    lib.abg_mapper_sync.argtypes = [C.c_void_p]
    lib.abg_mapper_download.argtypes = [C.c_void_p, C.POINTER(abg_results)]
    lib.abg_mapper_last_kernel_ms.argtypes = [C.c_void_p]
    lib.abg_mapper_last_kernel_ms.restype = C.c_float
    lib.abg_mapper_last_phase_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.abg_mapper_last_phase_ms.restype = None
    lib.abg_mapper_last_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.abg_mapper_last_kernel_times.restype = None
    lib.abg_mapper_launches_per_run.argtypes = [C.c_void_p]
    lib.abg_mapper_launches_per_run.restype = C.c_uint32
    lib.abg_mapper_get_counters.argtypes = [C.c_void_p, C.POINTER(abg_work_counters)]
    lib.abg_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    lib.abg_host_free.argtypes = [C.c_void_p]
    lib.abg_host_free.restype = None
    lib.abg_mapper_last_seed_times.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.abg_mapper_last_seed_times.restype = None
    lib.abg_mapper_binned.argtypes = [C.c_void_p]
    lib.abg_mapper_binned.restype = C.c_int
    lib.abg_mapper_bin_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.abg_mapper_bin_stats.restype = C.c_int
    lib.abg_mapper_last_run_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    lib.abg_mapper_last_run_stats.restype = C.c_int
    lib.abg_mapper_chunk.argtypes = [C.c_void_p]
    lib.abg_mapper_chunk.restype = C.c_uint32
    _lib = lib
    return lib


class _PinnedBlock:
    """Page-locked host memory from abg_host_alloc; freed when the last array viewing it dies."""

    def __init__(self, nbytes):
        self.lib = load_library()
        self.ptr = C.c_void_p()
        if self.lib.abg_host_alloc(max(int(nbytes), 1), C.byref(self.ptr)) != 0:
            raise AbgError(self.lib.abg_last_error().decode())
        self.buf = (C.c_char * max(int(nbytes), 1)).from_address(self.ptr.value)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.abg_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


def pinned_zeros(shape, dtype):
    """numpy array over page-locked memory (DMA'd in place by abg_map_batch)."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if not np.isscalar(shape) else int(shape)
    blk = _PinnedBlock(n * dt.itemsize)
    root = np.frombuffer(blk.buf, dtype=dt, count=n)  # every view keeps `root` alive through .base
    root[...] = np.zeros((), dt)
    _keep[id(blk)] = blk
    weakref.finalize(root, _keep.pop, id(blk), None)
    return root.reshape(shape)


_keep = {}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_view(ix):
    """abg_index_view over the numpy arrays of an index_file.IndexFile."""
    v = abg_index_view()
    v.genome = _ptr(ix.genome)
    v.genome_words = ix.genome.size
    v.genome_size = ix.genome_size
    v.counter = _ptr(ix.counter)
    v.counter_size = ix.counter_size
    v.counter_t = _ptr(ix.counter_t)
    v.counter_a = _ptr(ix.counter_a)
    v.counter_size_three = ix.counter_size_three
    v.index = _ptr(ix.index)
    v.index_size = ix.index_size
    v.index_t = _ptr(ix.index_t)
    v.index_a = _ptr(ix.index_a)
    v.index_size_three = ix.index_size_three
    v.max_candidates = ix.max_candidates
    v.window_size = getattr(ix, "window_size", 20)
    return v


def make_params(mode=0, allow_ambig=False, min_dist=32, max_dist=3000, valid_frac=0.1, max_candidates=0,
                cigar_stride=64):
    p = abg_params()
    p.mode = mode
    p.allow_ambig = 1 if allow_ambig else 0
    p.min_dist = min_dist
    p.max_dist = max_dist
    p.valid_frac = valid_frac
    p.max_candidates = max_candidates
    p.cigar_stride = cigar_stride
    return p


class Results:
    """Host result arrays for one batch (numpy, caller owned)."""

    def __init__(self, n, paired, stride, pinned=False):
        self.n, self.paired, self.stride = n, paired, stride
        zeros = pinned_zeros if pinned else np.zeros
        self.se1 = zeros(n, HIT_DTYPE)
        self.cigar1 = zeros((n, stride), np.uint32)
        self.n_cigar1 = zeros(n, np.uint32)
        if paired:
            self.pe_r1 = zeros(n, HIT_DTYPE)
            self.pe_r2 = zeros(n, HIT_DTYPE)
            self.se2 = zeros(n, HIT_DTYPE)
            self.cigar2 = zeros((n, stride), np.uint32)
            self.n_cigar2 = zeros(n, np.uint32)
        else:
            self.pe_r1 = self.pe_r2 = self.se2 = self.cigar2 = self.n_cigar2 = None

    def struct(self):
        r = abg_results()
        for k in ("pe_r1", "pe_r2", "se1", "se2", "cigar1", "cigar2", "n_cigar1", "n_cigar2"):
            setattr(r, k, _ptr(getattr(self, k)))
        return r

    def d2h_bytes(self, inline_ops=16):
        """Bytes abg_map_batch copies device -> host: hit records, CIGAR lengths and the first
        `inline_ops` operations of every CIGAR row (longer CIGARs are fetched singly)."""
        tot = 0
        for k in ("pe_r1", "pe_r2", "se1", "se2", "n_cigar1", "n_cigar2"):
            a = getattr(self, k)
            if a is not None:
                tot += a.nbytes
        for k in ("cigar1", "cigar2"):
            a = getattr(self, k)
            if a is not None:
                tot += a.shape[0] * min(inline_ops, a.shape[1]) * 4
        return tot

    def cigars(self, end):
        cig, n = (self.cigar1, self.n_cigar1) if end == 1 else (self.cigar2, self.n_cigar2)
        return [cig[i, :n[i]].copy() for i in range(self.n)]


def batch_struct(b1, b2=None):
    s = abg_batch()
    s.n = b1.n
    s.seq1 = _ptr(b1.seq)
    s.off1 = _ptr(b1.off)
    if b2 is not None:
        s.seq2 = _ptr(b2.seq)
        s.off2 = _ptr(b2.off)
    return s


class Index:
    """The AbismalIndex arrays resident in the HBM of one GPU."""

    def __init__(self, index_file, device=0):
        self.lib = load_library()
        self.index_file = index_file
        self._h = C.c_void_p()
        v = make_view(index_file)
        rc = self.lib.abg_index_create(C.byref(v), device, C.byref(self._h))
        if rc != 0:
            raise AbgError(self.lib.abg_last_error().decode())

    @property
    def device_bytes(self):
        return int(self.lib.abg_index_device_bytes(self._h))

    @property
    def features(self):
        """FEATURE_* bits: which derived arrays (accelerators of the same lookups) the index holds."""
        return int(self.lib.abg_index_features(self._h))

    def close(self):
        if self._h:
            self.lib.abg_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Mapper:
    """One stream's worth of mapping state on the GPU holding `index`."""

    def __init__(self, index, mode=0, allow_ambig=False, min_dist=32, max_dist=3000, valid_frac=0.1,
                 max_candidates=0, cigar_stride=64, max_batch=65536, max_read_len=256, count_work=False):
        self.lib = index.lib
        self.index = index
        self.params = make_params(mode, allow_ambig, min_dist, max_dist, valid_frac, max_candidates, cigar_stride)
        self.paired = bool(mode & MODE_PAIRED)
        self.stride = cigar_stride
        self._h = C.c_void_p()
        rc = self.lib.abg_mapper_create(index._h, C.byref(self.params), max_batch, max_read_len,
                                        1 if count_work else 0, C.byref(self._h))
        if rc != 0:
            raise AbgError(self.lib.abg_last_error().decode())

    def _check(self, rc):
        if rc != 0:
            raise AbgError(self.lib.abg_last_error().decode())

    def map_batch(self, b1, b2=None, results=None):
        """Host buffers in, host buffers out (the call a user makes)."""
        res = results if results is not None else Results(b1.n, self.paired, self.stride)
        bs, rs = batch_struct(b1, b2), res.struct()
        self._check(self.lib.abg_map_batch(self._h, C.byref(bs), C.byref(rs)))
        return res

    def upload(self, b1, b2=None):
        bs = batch_struct(b1, b2)
        self._check(self.lib.abg_mapper_upload(self._h, C.byref(bs)))

    def run(self):
        self._check(self.lib.abg_mapper_run(self._h))

    def sync(self):
        self._check(self.lib.abg_mapper_sync(self._h))

    def download(self, n, results=None):
        res = results if results is not None else Results(n, self.paired, self.stride)
        rs = res.struct()
        self._check(self.lib.abg_mapper_download(self._h, C.byref(rs)))
        return res

    @property
    def last_kernel_ms(self):
        return float(self.lib.abg_mapper_last_kernel_ms(self._h))

    @property
    def last_phase_ms(self):
        """(seed_kernel, align_kernel, redo map_reads_kernel) CUDA-event ms of the last run()."""
        out = (C.c_float * 3)()
        self.lib.abg_mapper_last_phase_ms(self._h, out)
        return [float(x) for x in out]

    KERNELS = ("seed_kernel", "enum_kernel", "dp_kernel", "align_kernel", "map_reads_kernel(redo)")

    @property
    def last_kernel_times(self):
        """CUDA-event ms of the five kernels of the last run(), in the order of Mapper.KERNELS."""
        out = (C.c_float * 5)()
        self.lib.abg_mapper_last_kernel_times(self._h, out)
        return [float(x) for x in out]

    @property
    def launches_per_run(self):
        return int(self.lib.abg_mapper_launches_per_run(self._h))

    SEED_KERNELS = ("hash_kernel", "scatter_kernel", "filter_kernel", "seed_kernel")

    @property
    def last_seed_times(self):
        """CUDA-event ms of the seeding kernels of the last run(), in the order of Mapper.SEED_KERNELS (binned
        seeding; their sum is last_kernel_times[0])."""
        out = (C.c_float * 4)()
        self.lib.abg_mapper_last_seed_times(self._h, out)
        return [float(x) for x in out]

    @property
    def binned(self):
        """True when the mapper seeds through the binned kernels (seed_bins.cuh)."""
        return bool(self.lib.abg_mapper_binned(self._h))

    def bin_stats(self):
        out = (C.c_uint64 * 8)()
        self._check(self.lib.abg_mapper_bin_stats(self._h, out))
        keys = ("strands", "strands_direct", "tuples", "survivors", "bins", "tuple_cap", "variants", "filter_grab")
        d = dict(zip(keys, [int(x) for x in out]))
        v = d.pop("variants")
        d["scatter"] = "tile-sorted" if v & 1 else "direct"
        d["filter"] = ("pipelined, " if v & 2 else "") + ("static distribution" if d["filter_grab"] == 0 else "work cursor")
        return d

    def last_run_stats(self):
        """Diagnostics of the last run(): redo pairs, set-arena use, tasks per band class (abg_mapper_last_run_stats)."""
        out = (C.c_uint32 * 8)()
        self._check(self.lib.abg_mapper_last_run_stats(self._h, out))
        keys = ("redo_pairs", "set_arena_used", "set_arena_cap", "tasks_bw16", "tasks_bw32", "tasks_bw61", "tb_units", "error_flag")
        return dict(zip(keys, [int(x) for x in out]))

    def counters(self):
        c = abg_work_counters()
        self._check(self.lib.abg_mapper_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def close(self):
        if self._h:
            self.lib.abg_mapper_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
