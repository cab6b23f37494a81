"""GPU tests of the --enable-short variant: seed::window_size 12 instead of 20
(src/AbismalIndex.hpp:73-77).  The reference needs a differently configured binary for it
(oracle/_ref/abismal_short); here the window is a property of the index file and a flag of the builder.
Bit-exact: index bytes, records, SAM text, stats."""
import numpy as np
import pytest

import helpers
from test_enable_short import CASES

pytestmark = pytest.mark.gpu


def test_cli_idx_enable_short_is_byte_identical(workspace):
    workspace.need_short()
    helpers.run([helpers.CLI, "idx", "-enable-short", "tests/rep.fa", "tests/rep_w12.cli.idx"], cwd=workspace.dir)
    assert helpers.md5(workspace.path("rep_w12.cli.idx")) == helpers.md5(workspace.path("rep_w12.idx"))


@pytest.mark.parametrize("tag,args", CASES, ids=[c[0] for c in CASES])
def test_cli_equals_short_reference(workspace, tag, args):
    workspace.need_short()
    ref = workspace.map_with(helpers.REF_BIN_SHORT, "ref_" + tag, args)
    got = workspace.map_with(helpers.CLI, "gpu_" + tag, args)
    assert helpers.sam_body(ref[0]) == helpers.sam_body(got[0])
    assert open(ref[1]).read() == open(got[1]).read()


def test_cli_map_g_enable_short(workspace):
    workspace.need_short()
    reads = ["tests/w12_pe_1.fq", "tests/w12_pe_2.fq"]
    ref = workspace.map_with(helpers.REF_BIN_SHORT, "ref_g12", ["-i", "tests/rep_w12.idx"] + reads)
    got = workspace.map_with(helpers.CLI, "gpu_g12", ["-g", "tests/rep.fa"] + reads, pre=["-enable-short"])
    assert helpers.sam_body(ref[0]) == helpers.sam_body(got[0])
    assert open(ref[1]).read() == open(got[1]).read()


@pytest.mark.parametrize("kind", ["rep", "plain"])
@pytest.mark.parametrize("window,mode,files", [(12, 0, ("w12_se_1.fq",)), (12, 1, ("w12_pe_1.fq", "w12_pe_2.fq")),
                                               (12, 1 | 4, ("w12_rpe_1.fq", "w12_rpe_2.fq")),
                                               (20, 0, ("w12_se_1.fq",)), (20, 1, ("w12_pe_1.fq", "w12_pe_2.fq"))])
def test_records_equal_oracle_shortest_reads(workspace, window, mode, files, kind):
    """Through the C ABI (abg_index_view.window_size), with every fifth read (pair) cut down to the shortest
    lengths the window admits: 36..47 bases for window 12, 44..55 for window 20.  Below 2 * window + 8 bases
    the specific phase visits seed offsets whose 25-mer overhangs the read (the reference reads past the end
    of its buffer there; the oracle and the kernels define the missing bases as 0) and which the sensitive
    phase does not revisit."""
    from abismal_b200 import Index, IndexFile, Mapper, load_fastq
    from abismal_b200.reads import ReadBatch
    from abismal_b200 import capi
    workspace.need_short(kind)
    workspace.need_repeat(kind)
    ixf = IndexFile(workspace.path(kind + ("_w12.idx" if window == 12 else ".idx")))
    assert ixf.window_size == window
    lo = 25 + window - 1
    files = [f.replace("w12_", "w12_" if kind == "rep" else "w12%s_" % kind) for f in files]
    b = [load_fastq(workspace.path(f), min_read_length=lo) for f in files]
    cut = []
    for x in b:
        seqs = [x.sequence(i) for i in range(x.n)]
        seqs = [s[:lo + (i % 12)] if i % 5 == 0 else s for i, s in enumerate(seqs)]
        cut.append(ReadBatch(None, seqs))
    ix = Index(ixf, 0)
    assert ix.features & capi.FEATURE_SEED_CONTEXT
    m = Mapper(ix, mode=mode, max_batch=cut[0].n, max_read_len=128)
    o = helpers.OracleMapper(ixf, mode=mode)
    got, want = m.map_batch(*cut), o.map_batch(*cut)
    helpers.assert_results_equal(got, want, bool(mode & 1))
    mapped = want.pe_r1["pos"] != 0 if mode & 1 else want.se1["pos"] != 0
    assert mapped.sum() > 50
    short = np.array([i % 5 == 0 for i in range(cut[0].n)])
    assert (mapped & short).sum() > 5  # some of the shortest reads do map
    m.close()
    o.close()
    ix.close()
