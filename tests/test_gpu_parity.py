"""GPU parity tests (run on the B200 box): the CUDA path, called through the C
ABI (ctypes) and through the `abismal-b200 map` front end, against
 (a) the CPU restatement (oracle/libabismal_oracle.so) record by record, and
 (b) the unmodified reference binary (oracle/_ref/abismal) as SAM text,
 (c) the reference's golden md5s for its own four map tests.
Everything is integer work: the bar is bit-exact.
"""
import os

import numpy as np
import pytest

import helpers
from test_oracle_vs_ref import MAP_CMDS, MAP_PRE, UNPINNED, golden_md5

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from abismal_b200 import capi
    lib = capi.load_library()  # raises if the CUDA library is missing: no fallback
    if lib.abg_device_count() < 1:
        pytest.fail("no CUDA device visible to libabismal_b200.so")
    return capi


KINDS = ["rep", "plain"]  # repeat genome with / without IUPAC codes (helpers.Workspace.need_repeat)


@pytest.fixture(scope="module", params=KINDS)
def rep_index(request, workspace, gpu):
    """(index file, index in HBM, fixture kind).  Both variants must run the benchmarked compare: the
    seed-context prefilter is on (in the IUPAC variant only the windows near such a code take the exact route)."""
    from abismal_b200 import Index, IndexFile
    kind = request.param
    workspace.need_repeat(kind)
    ixf = IndexFile(workspace.path(kind + ".idx"))
    ix = Index(ixf, 0)
    f = ix.features
    assert f & gpu.FEATURE_SEED_CONTEXT, "seed-context records were not built"
    assert bool(f & gpu.FEATURE_GENOME_HAS_IUPAC) == (kind == "rep")
    yield ixf, ix, kind
    ix.close()


def _fq(workspace, name, limit=None, kind="rep"):
    from abismal_b200 import load_fastq
    return load_fastq(workspace.path(name.replace("rep_", kind + "_", 1)), limit)


RECORD_CASES = [
    # tag, mode bits, kwargs, fastq files
    ("se", 0, {}, ("rep_se_1.fq",)),
    ("se_arich", 2, {}, ("rep_se_1.fq",)),
    ("se_rpbat", 4, {}, ("rep_se_1.fq",)),
    ("se_ambig_m", 0, {"allow_ambig": True, "valid_frac": 0.2}, ("rep_pe_1.fq",)),
    ("se_c5", 0, {"max_candidates": 5}, ("rep_pe_2.fq",)),
    ("pe", 1, {}, ("rep_pe_1.fq", "rep_pe_2.fq")),
    ("pe_pbat", 1 | 2, {}, ("rep_pbat_1.fq", "rep_pbat_2.fq")),
    ("pe_rpbat", 1 | 4, {}, ("rep_rpe_1.fq", "rep_rpe_2.fq")),
    ("pe_rpbat_ambig", 1 | 4, {"allow_ambig": True}, ("rep_rpe_1.fq", "rep_rpe_2.fq")),
    ("pe_frag", 1, {"min_dist": 50, "max_dist": 300, "valid_frac": 0.2}, ("rep_pe_1.fq", "rep_pe_2.fq")),
    ("pe_c10", 1, {"max_candidates": 10}, ("rep_pe_1.fq", "rep_pe_2.fq")),
]


@pytest.mark.parametrize("tag,mode,kw,files", RECORD_CASES, ids=[c[0] for c in RECORD_CASES])
def test_records_equal_oracle(workspace, rep_index, gpu, tag, mode, kw, files):
    from abismal_b200 import Mapper
    ixf, ix, kind = rep_index
    b = [_fq(workspace, f, kind=kind) for f in files]
    max_len = max(x.max_len for x in b)
    m = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max(max_len, 64), **kw)
    o = helpers.OracleMapper(ixf, mode=mode, **kw)
    got = m.map_batch(*b)
    want = o.map_batch(*b)
    helpers.assert_results_equal(got, want, bool(mode & 1))
    mapped = want.pe_r1["pos"] != 0 if mode & 1 else want.se1["pos"] != 0
    assert mapped.sum() > 30  # the case is not vacuous
    # the benchmarked seeding ran: hash -> scatter -> filter -> replay of the listed survivors, for most strands
    assert m.binned
    st = m.bin_stats()
    assert st["tuples"] > 0 and st["survivors"] > 0 and st["strands_direct"] < 0.5 * st["strands"], st
    m.close()
    o.close()


@pytest.mark.parametrize("tag,mode,kw,files", [c for c in RECORD_CASES if c[0] in ("se", "se_rpbat", "pe", "pe_rpbat_ambig", "pe_c10")],
                         ids=["se", "se_rpbat", "pe", "pe_rpbat_ambig", "pe_c10"])
def test_records_equal_oracle_with_direct_seeding(workspace, rep_index, gpu, monkeypatch, tag, mode, kw, files):
    """ABISMAL_B200_BINS=0: every strand gathers its own seed-context records (process_seeds in one warp), the
    path strands outside the binned kernels' fast path take (reads with N, survivor or tuple overflow)."""
    from abismal_b200 import Mapper
    ixf, ix, kind = rep_index
    monkeypatch.setenv("ABISMAL_B200_BINS", "0")
    b = [_fq(workspace, f, kind=kind) for f in files]
    m = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max([x.max_len for x in b] + [64]), **kw)
    monkeypatch.delenv("ABISMAL_B200_BINS")
    assert not m.binned
    o = helpers.OracleMapper(ixf, mode=mode, **kw)
    helpers.assert_results_equal(m.map_batch(*b), o.map_batch(*b), bool(mode & 1))
    o.close()
    m.close()


PE_CASES = [c for c in RECORD_CASES if c[1] & 1]


@pytest.mark.parametrize("tag,mode,kw,files", PE_CASES, ids=[c[0] for c in PE_CASES])
def test_records_equal_oracle_best_pair_by_rows(workspace, rep_index, gpu, monkeypatch, tag, mode, kw, files):
    """ABISMAL_B200_HEAVY_MIN=0: every pair takes the row-parallel best_pair that large candidate sets (repeats)
    take by default -- scores from the task results, 32 rows per step, the in-order fold for chunks that can reach
    sure_ambig -- and gives the records of the two-pointer sweep."""
    from abismal_b200 import Mapper
    ixf, ix, kind = rep_index
    monkeypatch.setenv("ABISMAL_B200_HEAVY_MIN", "0")
    b = [_fq(workspace, f, kind=kind) for f in files]
    m = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max([x.max_len for x in b] + [64]), **kw)
    monkeypatch.delenv("ABISMAL_B200_HEAVY_MIN")
    o = helpers.OracleMapper(ixf, mode=mode, **kw)
    helpers.assert_results_equal(m.map_batch(*b), o.map_batch(*b), True)
    o.close()
    m.close()


def test_binned_seeding_overflow_paths(workspace, rep_index, gpu, monkeypatch):
    """Tuple memory too small for the batch, survivor lists too short: the strands that do not fit take the direct
    path, same records."""
    from abismal_b200 import Mapper
    ixf, ix, kind = rep_index
    b1, b2 = _fq(workspace, "rep_pe_1.fq", kind=kind), _fq(workspace, "rep_pe_2.fq", kind=kind)
    o = helpers.OracleMapper(ixf, mode=1)
    want = o.map_batch(b1, b2)
    o.close()
    for var, val in (("ABISMAL_B200_TUPLE_CAP", "16384"), ("ABISMAL_B200_SURV_CAP", "3")):
        monkeypatch.setenv(var, val)
        m = Mapper(ix, mode=1, max_batch=b1.n, max_read_len=160)
        monkeypatch.delenv(var)
        got = m.map_batch(b1, b2)
        st = m.bin_stats()
        assert m.binned and 0 < st["strands_direct"] < st["strands"], (var, st)
        helpers.assert_results_equal(got, want, True)
        m.close()


BIN_VARIANTS = [
    ("many_bins", {"ABISMAL_B200_BIN_SHIFT": "11"}),                      # hundreds of bins on the test genome
    ("direct_scatter", {"ABISMAL_B200_SCATTER_SORT": "0", "ABISMAL_B200_BIN_SHIFT": "11"}),
    ("static_filter", {"ABISMAL_B200_FILTER_GRAB": "0", "ABISMAL_B200_BIN_SHIFT": "12"}),
    ("small_grab_two_ctas", {"ABISMAL_B200_FILTER_GRAB": "32", "ABISMAL_B200_SCATTER_SORT": "0", "ABISMAL_B200_SCATTER_CTAS": "2",
                             "ABISMAL_B200_BIN_SHIFT": "12"}),
    ("interleaved_cursors", {"ABISMAL_B200_FILTER_GRAB": "32", "ABISMAL_B200_FILTER_CURSORS": "16", "ABISMAL_B200_BIN_SHIFT": "12"}),
    ("pipelined_filter", {"ABISMAL_B200_FILTER_PIPE": "1", "ABISMAL_B200_BIN_SHIFT": "12"}),
    ("cache_hints", {"ABISMAL_B200_FILTER_CACHE": "3", "ABISMAL_B200_BIN_SHIFT": "12"}),
    ("general_replay", {"ABISMAL_B200_ACC_CAP": "0", "ABISMAL_B200_BIN_SHIFT": "12"}),  # every strand ranks all its survivors
    ("short_replay_of_3", {"ABISMAL_B200_ACC_CAP": "3", "ABISMAL_B200_BIN_SHIFT": "12"}),
]


@pytest.mark.parametrize("variant,env", BIN_VARIANTS, ids=[v[0] for v in BIN_VARIANTS])
@pytest.mark.parametrize("tag,mode,kw,files", [c for c in RECORD_CASES if c[0] in ("se_rpbat", "pe_pbat")], ids=["se_rpbat", "pe_pbat"])
def test_binned_seeding_kernel_variants(workspace, rep_index, gpu, monkeypatch, tag, mode, kw, files, variant, env):
    """The tuning switches of the binned kernels (bin size, tile-sorted scatter, static / small-grab filter work
    distribution, pipelined filter) change the order in which tuples and survivors are produced, never the
    records.  The small bin sizes give the test genome hundreds of bins (the default puts it in a handful)."""
    from abismal_b200 import Mapper
    ixf, ix, kind = rep_index
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    b = [_fq(workspace, f, kind=kind) for f in files]
    m = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max([x.max_len for x in b] + [64]), **kw)
    for k in env:
        monkeypatch.delenv(k)
    assert m.binned
    st = m.bin_stats()
    o = helpers.OracleMapper(ixf, mode=mode, **kw)
    helpers.assert_results_equal(m.map_batch(*b), o.map_batch(*b), bool(mode & 1))
    st = m.bin_stats()
    assert st["bins"] >= 32, st
    assert (st["scatter"] == "tile-sorted") == (env.get("ABISMAL_B200_SCATTER_SORT", "1") == "1"), st
    assert st["filter"].startswith("pipelined") == ("ABISMAL_B200_FILTER_PIPE" in env), st
    assert (st["filter_grab"] == 0) == (env.get("ABISMAL_B200_FILTER_GRAB") == "0"), st
    o.close()
    m.close()


@pytest.mark.parametrize("tag,mode,kw,files", [c for c in RECORD_CASES if c[0] in ("se", "pe", "pe_rpbat_ambig")],
                         ids=["se", "pe", "pe_rpbat_ambig"])
def test_records_equal_oracle_with_alignments_in_the_warp(workspace, rep_index, gpu, monkeypatch, tag, mode, kw, files):
    """ABISMAL_B200_TASKS=0: every banded alignment runs in the warp of its read / pair (the path reads longer
    than 512 bases, the single-end fallback of a pair and the redo kernel always take) instead of the
    enum_kernel -> dp_kernel task lists."""
    from abismal_b200 import Mapper
    ixf, ix, kind = rep_index
    monkeypatch.setenv("ABISMAL_B200_TASKS", "0")
    b = [_fq(workspace, f, kind=kind) for f in files]
    m = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max([x.max_len for x in b] + [64]), **kw)
    monkeypatch.delenv("ABISMAL_B200_TASKS")
    mt = Mapper(ix, mode=mode, max_batch=b[0].n, max_read_len=max([x.max_len for x in b] + [64]), **kw)
    got, got_t = m.map_batch(*b), mt.map_batch(*b)
    assert m.launches_per_run == 3 + 4 * m.binned and mt.launches_per_run == 5 + 4 * mt.binned
    helpers.assert_results_equal(got, got_t, bool(mode & 1))
    o = helpers.OracleMapper(ixf, mode=mode, **kw)
    helpers.assert_results_equal(got, o.map_batch(*b), bool(mode & 1))
    o.close()
    m.close()
    mt.close()


def test_batch_split_and_repeat_invariance(workspace, rep_index, gpu):
    """Size-independent properties: results do not depend on how reads are
    batched, on their order, or on how often a batch is mapped."""
    from abismal_b200 import Mapper
    ixf, ix, kind = rep_index
    b1, b2 = _fq(workspace, "rep_pe_1.fq", kind=kind), _fq(workspace, "rep_pe_2.fq", kind=kind)
    m = Mapper(ix, mode=1, max_batch=b1.n, max_read_len=160)
    full = m.map_batch(b1, b2)
    again = m.map_batch(b1, b2)
    helpers.assert_results_equal(full, again, True)
    cut = b1.n // 3
    lo = m.map_batch(b1.slice(0, cut), b2.slice(0, cut))
    hi = m.map_batch(b1.slice(cut, b1.n), b2.slice(cut, b2.n))
    for k in ("pe_r1", "pe_r2", "se1", "se2", "n_cigar1", "n_cigar2"):
        assert np.array_equal(np.concatenate([getattr(lo, k), getattr(hi, k)]), getattr(full, k)), k
    m.close()


def test_empty_and_ragged_batches(workspace, rep_index, gpu):
    from abismal_b200 import Mapper, ReadBatch
    ixf, ix, kind = rep_index
    src = _fq(workspace, "rep_pe_1.fq", 64, kind=kind)
    seqs = [src.sequence(i) for i in range(src.n)]
    # ragged: skipped reads (empty), trimmed reads of different lengths, an all-N-masked read
    seqs[0] = ""
    seqs[5] = seqs[5][:44]
    seqs[6] = seqs[6][:47]
    seqs[7] = seqs[7][:48]
    seqs[8] = seqs[8][:60] + "N" * 20 + seqs[8][80:]
    seqs[9] = "ACGT" * 30
    seqs[10] = "A" * 100
    b = ReadBatch(["r%d" % i for i in range(len(seqs))], seqs)
    m = Mapper(ix, mode=0, max_batch=256, max_read_len=160)
    o = helpers.OracleMapper(ixf, mode=0)
    helpers.assert_results_equal(m.map_batch(b), o.map_batch(b), False)
    # paired with one or both ends empty
    seqs2 = list(reversed(seqs))
    seqs2[0] = ""
    seqs2[-1] = ""
    b2 = ReadBatch(b.names, seqs2)
    mp = Mapper(ix, mode=1, max_batch=256, max_read_len=160)
    op = helpers.OracleMapper(ixf, mode=1)
    helpers.assert_results_equal(mp.map_batch(b, b2), op.map_batch(b, b2), True)
    # an empty batch is legal
    e = ReadBatch([], [])
    r = m.map_batch(e)
    assert r.n == 0
    for x in (m, mp, o, op):
        x.close()


def test_bad_arguments_fail_loudly(rep_index, gpu):
    from abismal_b200 import AbgError, Mapper, ReadBatch
    ixf, ix, kind = rep_index
    m = Mapper(ix, mode=0, max_batch=8, max_read_len=100)
    with pytest.raises(AbgError):
        m.map_batch(ReadBatch(["x"], ["ACGT" * 50]))  # longer than max_read_len
    with pytest.raises(AbgError):
        m.map_batch(ReadBatch(["x"] * 9, ["ACGT" * 20] * 9))  # more than max_batch
    with pytest.raises(AbgError):
        m.map_batch(ReadBatch(["x"], ["ACGT" * 5]))  # shorter than 44: loader must empty it
    m.close()


@pytest.mark.parametrize("tag", sorted(MAP_CMDS))
def test_cli_golden_md5(workspace, gpu, tag):
    """The reference's own md5-pinned map tests, run through abismal-b200."""
    workspace.need_trex()
    g = golden_md5()
    sam, st, _ = workspace.map_with(helpers.CLI, tag, MAP_CMDS[tag], MAP_PRE[tag])
    assert helpers.md5(sam) == g["tests/%s.sam" % tag]
    assert helpers.md5(st) == g["tests/%s.mstats" % tag]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("tag,args", UNPINNED, ids=[u[0] for u in UNPINNED])
def test_cli_equals_reference_binary(workspace, gpu, tag, args, kind):
    workspace.need_repeat(kind)
    args = helpers.for_kind(args, kind)
    rsam, rst, _ = workspace.map_with(helpers.REF_BIN, "ref_%s_%s" % (kind, tag), args)
    gsam, gst, _ = workspace.map_with(helpers.CLI, "gpu_%s_%s" % (kind, tag), args)
    assert helpers.sam_body(rsam) == helpers.sam_body(gsam)
    assert open(rst).read() == open(gst).read()


def test_cli_small_gpu_batches_and_gz(workspace, gpu):
    """Output order and content do not depend on the GPU batch size; .gz input works."""
    workspace.need_repeat()
    import gzip
    import shutil
    for k in (1, 2):
        with open(workspace.path("rep_pe_%d.fq" % k), "rb") as fi, gzip.open(workspace.path("rep_pe_%d.fq.gz" % k), "wb") as fo:
            shutil.copyfileobj(fi, fo)
    a, ast_, _ = workspace.map_with(helpers.CLI, "b_default", ["-i", "tests/rep.idx", "tests/rep_pe_1.fq", "tests/rep_pe_2.fq"])
    b, bst, _ = workspace.map_with(helpers.CLI, "b_777", ["-gpu-batch", "777", "-i", "tests/rep.idx",
                                                        "tests/rep_pe_1.fq.gz", "tests/rep_pe_2.fq.gz"])
    assert helpers.sam_body(a)[3:] == helpers.sam_body(b)[3:]
    assert open(ast_).read() == open(bst).read()


def test_pipelined_chunks_and_pinned_buffers(workspace, rep_index, gpu, monkeypatch):
    """abg_map_batch pipelines sub-batches over three streams; results must not depend on the
    sub-batch size nor on whether the caller's buffers are pinned (DMA in place) or pageable (staged)."""
    from abismal_b200 import Mapper
    from abismal_b200.capi import Results
    ixf, ix, kind = rep_index
    b1, b2 = _fq(workspace, "rep_pe_1.fq", kind=kind), _fq(workspace, "rep_pe_2.fq", kind=kind)
    m = Mapper(ix, mode=1, max_batch=b1.n, max_read_len=160)
    want = m.map_batch(b1, b2)
    m.close()
    monkeypatch.setenv("ABISMAL_B200_CHUNK", "777")
    mc = Mapper(ix, mode=1, max_batch=b1.n, max_read_len=160)
    assert mc.lib.abg_mapper_chunk(mc._h) == 777
    got = mc.map_batch(b1, b2)
    helpers.assert_results_equal(got, want, True)
    p1, p2 = b1.to_pinned(), b2.to_pinned()
    res = Results(b1.n, True, mc.stride, pinned=True)
    for _ in range(2):
        mc.map_batch(p1, p2, res)
    helpers.assert_results_equal(res, want, True)
    # split API on the same mapper
    mc.upload(p1, p2)
    mc.run()
    helpers.assert_results_equal(mc.download(b1.n), want, True)
    mc.close()
