"""Shared test plumbing: paths, builds, workspaces, the oracle's ctypes binding."""
import ctypes as C
import hashlib
import os
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "abismal")
# the same sources compiled with -DENABLE_SHORT (configure --enable-short: seed::window_size 12)
REF_BIN_SHORT = os.path.join(ORACLE_DIR, "_ref", "abismal_short")
REF_DATA = os.path.join(ORACLE_DIR, "_ref", "data")
ORACLE_MAP = os.path.join(ORACLE_DIR, "oracle_map")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libabismal_oracle.so")
CLI = os.path.join(ROOT, "abismal_b200", "bin", "abismal-b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(cmd, cwd=None, check=True):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if check and p.returncode != 0:
        raise RuntimeError("command failed (%d): %s\n%s" % (p.returncode, " ".join(cmd), p.stderr[-2000:]))
    return p


def ensure_built():
    if not (os.path.exists(ORACLE_LIB) and os.path.exists(ORACLE_MAP)):
        run(["make", "-C", ORACLE_DIR, "oracle", "oracle_map"])
    if not (os.path.exists(REF_BIN) and os.path.exists(REF_BIN_SHORT)) and os.path.isdir("/root/reference/src"):
        run(["make", "-C", ORACLE_DIR, "-j8", "ref"])
    lib = os.path.join(ROOT, "abismal_b200", "libabismal_b200.so")
    if not (os.path.exists(lib) and os.path.exists(CLI)):
        run(["make", "-C", os.path.join(ROOT, "abismal_b200", "csrc")])


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def sam_body(path):
    """SAM text without the @PG line (it embeds argv)."""
    with open(path) as f:
        return [ln for ln in f if not ln.startswith("@PG")]


class Workspace:
    """tests/<files> layout of the reference's own test scripts plus a repeat-rich genome."""

    def __init__(self, d):
        self.dir = d
        self.tests = os.path.join(d, "tests")
        os.makedirs(self.tests)
        self._made = set()

    def path(self, name):
        return os.path.join(self.tests, name)

    def ref(self, *args):
        return run([REF_BIN] + list(args), cwd=self.dir)

    def need_trex(self):
        if "trex" in self._made:
            return
        shutil.copy(os.path.join(REF_DATA, "tRex1.fa"), self.path("tRex1.fa"))
        self.ref("idx", "tests/tRex1.fa", "tests/tRex1.idx")
        common = ["-seed", "1", "-n", "10000", "-m", "0.01", "-b", "0.98", "tests/tRex1.fa"]
        self.ref("sim", "-single", "-o", "tests/reads", *common)
        self.ref("sim", "-o", "tests/reads_pe", *common)
        self.ref("sim", "-a", "-o", "tests/reads_pbat_pe", *common)
        self.ref("sim", "-R", "-o", "tests/reads_rpbat_pe", *common)
        self._made.add("trex")

    def need_repeat(self, kind="rep"):
        """kind "rep": the repeat genome with N runs and IUPAC codes; "plain": the same without IUPAC codes.
        Files are <kind>.fa, <kind>.idx, <kind>_pe_1.fq, ..."""
        if kind in self._made:
            return
        import make_genome
        k = kind
        make_genome.write_fasta(make_genome.repeat_genome(iupac=(kind == "rep")), self.path(k + ".fa"))
        self.ref("idx", "-t", "4", "tests/%s.fa" % k, "tests/%s.idx" % k)
        self.ref("sim", "-seed", "3", "-l", "150", "-min-fraglen", "150", "-max-fraglen", "400", "-n", "6000",
                 "-m", "0.02", "-b", "0.98", "-o", "tests/%s_pe" % k, "tests/%s.fa" % k)
        self.ref("sim", "-seed", "4", "-R", "-l", "120", "-min-fraglen", "120", "-max-fraglen", "300", "-n", "4000",
                 "-m", "0.03", "-b", "0.9", "-o", "tests/%s_rpe" % k, "tests/%s.fa" % k)
        self.ref("sim", "-seed", "5", "-single", "-l", "75", "-n", "5000", "-m", "0.05", "-b", "0.95",
                 "-o", "tests/%s_se" % k, "tests/%s.fa" % k)
        self.ref("sim", "-seed", "6", "-a", "-l", "100", "-min-fraglen", "100", "-max-fraglen", "250", "-n", "4000",
                 "-m", "0.01", "-b", "0.98", "-o", "tests/%s_pbat" % k, "tests/%s.fa" % k)
        self._made.add(kind)

    def need_short(self, kind="rep"):
        """Window-12 index (the --enable-short reference) of the repeat genome + reads down to 48 bases."""
        if "short_" + kind in self._made:
            return
        import make_genome
        k = kind
        if not os.path.exists(self.path(k + ".fa")):
            make_genome.write_fasta(make_genome.repeat_genome(iupac=(kind == "rep")), self.path(k + ".fa"))
        pre = "w12" if kind == "rep" else "w12" + kind
        run([REF_BIN_SHORT, "idx", "-t", "4", "tests/%s.fa" % k, "tests/%s_w12.idx" % k], cwd=self.dir)
        self.ref("sim", "-seed", "11", "-single", "-l", "50", "-n", "5000", "-m", "0.03", "-b", "0.95",
                 "-o", "tests/%s_se" % pre, "tests/%s.fa" % k)
        self.ref("sim", "-seed", "12", "-l", "60", "-min-fraglen", "60", "-max-fraglen", "250", "-n", "4000",
                 "-m", "0.02", "-b", "0.98", "-o", "tests/%s_pe" % pre, "tests/%s.fa" % k)
        self.ref("sim", "-seed", "13", "-R", "-l", "100", "-min-fraglen", "100", "-max-fraglen", "300", "-n", "3000",
                 "-m", "0.02", "-b", "0.9", "-o", "tests/%s_rpe" % pre, "tests/%s.fa" % k)
        self._made.add("short_" + kind)

    def map_with(self, tool, tag, args, pre=()):
        """Run `<tool> map <pre...> -s <stats> -o <sam> <args...>` from the workspace dir
        (argument order matters: the @PG line embeds argv)."""
        sam, st = "tests/%s.sam" % tag, "tests/%s.mstats" % tag
        p = run([tool, "map"] + list(pre) + ["-s", st, "-o", sam] + list(args), cwd=self.dir)
        return os.path.join(self.dir, sam), os.path.join(self.dir, st), p


# ---- oracle ctypes binding (same structs as the product's C ABI) ---------------
_olib = None


def oracle_lib():
    global _olib
    if _olib is None:
        from abismal_b200 import capi
        lib = C.CDLL(ORACLE_LIB)
        lib.abo_last_error.restype = C.c_char_p
        lib.abo_index_create.argtypes = [C.POINTER(capi.abg_index_view), C.POINTER(C.c_void_p)]
        lib.abo_index_destroy.argtypes = [C.c_void_p]
        lib.abo_index_destroy.restype = None
        lib.abo_map_batch.argtypes = [C.c_void_p, C.POINTER(capi.abg_params), C.POINTER(capi.abg_batch),
                                      C.POINTER(capi.abg_results), C.POINTER(capi.abg_work_counters)]
        _olib = lib
    return _olib


class OracleMapper:
    def __init__(self, index_file, **params):
        from abismal_b200 import capi
        self.capi = capi
        self.lib = oracle_lib()
        self.index_file = index_file
        self.view = capi.make_view(index_file)
        self.h = C.c_void_p()
        assert self.lib.abo_index_create(C.byref(self.view), C.byref(self.h)) == 0
        self.params = capi.make_params(**params)
        self.paired = bool(self.params.mode & capi.MODE_PAIRED)
        self.counters = capi.abg_work_counters()

    def map_batch(self, b1, b2=None):
        capi = self.capi
        res = capi.Results(b1.n, self.paired, self.params.cigar_stride)
        bs, rs = capi.batch_struct(b1, b2), res.struct()
        rc = self.lib.abo_map_batch(self.h, C.byref(self.params), C.byref(bs), C.byref(rs), C.byref(self.counters))
        if rc != 0:
            raise RuntimeError(self.lib.abo_last_error().decode())
        return res

    def close(self):
        if self.h:
            self.lib.abo_index_destroy(self.h)
            self.h = C.c_void_p()


def for_kind(args, kind):
    """File arguments written for the "rep" fixture, renamed for fixture `kind`."""
    if kind == "rep":
        return list(args)
    return [a.replace("tests/rep", "tests/" + kind).replace("tests/w12_", "tests/w12%s_" % kind) for a in args]


def assert_results_equal(a, b, paired):
    keys = ["se1", "n_cigar1"] + (["pe_r1", "pe_r2", "se2", "n_cigar2"] if paired else [])
    for k in keys:
        x, y = getattr(a, k), getattr(b, k)
        if not np.array_equal(x, y):
            bad = np.nonzero(x != y)[0]
            raise AssertionError("%s differs at %d of %d records, first %d: %r vs %r" %
                                 (k, bad.size, x.size, bad[0], x[bad[0]], y[bad[0]]))
    for e in (1, 2) if paired else (1,):
        ca, cb = a.cigars(e), b.cigars(e)
        for i, (p, q) in enumerate(zip(ca, cb)):
            if not np.array_equal(p, q):
                raise AssertionError("cigar%d differs at record %d: %r vs %r" % (e, i, p, q))


# ---- BAM -> SAM text (decoder used to check `-B` output record by record) -------------
def bam_to_sam_lines(path):
    """Decode a BAM file (BGZF blocks + BAM records, SAM spec 4.1/4.2) to SAM text lines.
    Also checks the BGZF framing: BC subfield, BSIZE, CRC32, ISIZE and the EOF marker."""
    import struct
    import zlib
    raw = open(path, "rb").read()
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"), "no BGZF EOF marker"
    data = bytearray()
    at = 0
    while at < len(raw):
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04", "bad BGZF block header at %d" % at
        xlen = struct.unpack_from("<H", raw, at + 10)[0]
        assert raw[at + 12:at + 16] == b"BC\x02\x00" and xlen == 6
        bsize = struct.unpack_from("<H", raw, at + 16)[0] + 1
        body = raw[at + 18:at + bsize - 8]
        crc, isize = struct.unpack_from("<II", raw, at + bsize - 8)
        blk = zlib.decompress(body, -15)
        assert len(blk) == isize and (zlib.crc32(blk) & 0xffffffff) == crc and isize <= 0xff00
        data += blk
        at += bsize
    data = bytes(data)
    assert data[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, p)[0]
        name = data[p + 4:p + 4 + l_name - 1].decode()
        l_ref = struct.unpack_from("<i", data, p + 4 + l_name)[0]
        refs.append((name, l_ref))
        p += 8 + l_name
    lines = [ln + "\n" for ln in text.split("\n") if ln]
    sq = [ln.rstrip("\n").split("\t") for ln in lines if ln.startswith("@SQ")]
    assert [(f[1][3:], int(f[2][3:])) for f in sq] == refs
    nt16 = "=ACMGRSVTWYHKDBN"
    ops = "MIDNSHP=X"
    while p < len(data):
        block_size = struct.unpack_from("<i", data, p)[0]
        rec = data[p + 4:p + 4 + block_size]
        p += 4 + block_size
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", rec, 0)
        q = 32
        name = rec[q:q + l_name - 1].decode()
        q += l_name
        cig = struct.unpack_from("<%dI" % n_cig, rec, q)
        q += 4 * n_cig
        sb = rec[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = "".join(nt16[(sb[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        qual = rec[q:q + l_seq]
        q += l_seq
        assert all(b == 0xff for b in qual)
        tags = []
        while q < len(rec):
            tag, typ = rec[q:q + 2].decode(), chr(rec[q + 2])
            q += 3
            if typ == "A":
                tags.append("%s:A:%s" % (tag, chr(rec[q])))
                q += 1
            else:
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[typ]
                tags.append("%s:i:%d" % (tag, struct.unpack_from(fmt, rec, q)[0]))
                q += struct.calcsize(fmt)
        rname = refs[tid][0] if tid >= 0 else "*"
        rnext = "*" if mtid < 0 else ("=" if mtid == tid else refs[mtid][0])
        cigar = "".join("%d%s" % (c >> 4, ops[c & 15]) for c in cig) or "*"
        lines.append("\t".join([name, str(flag), rname, str(pos + 1), str(mapq), cigar, rnext, str(mpos + 1), str(tlen),
                                seq or "*", "*"] + tags) + "\n")
    return lines
