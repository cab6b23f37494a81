"""GPU tests of the rows either side of the hot path through the `abismal-b200` front end:
`idx` (golden md5 of tRex1.idx), `map -g` (index built on the fly), `map -B` (BAM decodes to the SAM of the
reference binary) and the host pipeline (threads, small batches) -- all bit-exact."""
import os

import pytest

import helpers
from test_oracle_vs_ref import golden_md5

pytestmark = pytest.mark.gpu


def test_cli_idx_matches_golden_md5_and_reference_binary(workspace):
    workspace.need_trex()
    workspace.need_repeat()
    helpers.run([helpers.CLI, "idx", "-v", "tests/tRex1.fa", "tests/tRex1.cli.idx"], cwd=workspace.dir)
    assert helpers.md5(workspace.path("tRex1.cli.idx")) == golden_md5()["tests/tRex1.idx"]
    # repeat-rich genome with N runs, IUPAC codes, several chromosomes; gz input
    import gzip
    import shutil
    with open(workspace.path("rep.fa"), "rb") as fi, gzip.open(workspace.path("rep.fa.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    helpers.run([helpers.CLI, "idx", "tests/rep.fa.gz", "tests/rep.cli.idx"], cwd=workspace.dir)
    assert helpers.md5(workspace.path("rep.cli.idx")) == helpers.md5(workspace.path("rep.idx"))


def test_cli_map_with_genome_equals_map_with_index(workspace):
    workspace.need_repeat()
    reads = ["tests/rep_pe_1.fq", "tests/rep_pe_2.fq"]
    ref = workspace.map_with(helpers.REF_BIN, "ref_g", ["-i", "tests/rep.idx"] + reads)
    got = workspace.map_with(helpers.CLI, "gpu_g", ["-g", "tests/rep.fa"] + reads)
    assert helpers.sam_body(ref[0]) == helpers.sam_body(got[0])
    assert open(ref[1]).read() == open(got[1]).read()


@pytest.mark.parametrize("tag,args", [("se", ["tests/rep_se_1.fq"]),
                                      ("pe", ["tests/rep_pe_1.fq", "tests/rep_pe_2.fq"]),
                                      ("rpbat_a", ["-R", "-a", "tests/rep_rpe_1.fq", "tests/rep_rpe_2.fq"])])
def test_cli_bam_decodes_to_reference_sam(workspace, tag, args):
    workspace.need_repeat()
    pre = [a for a in args if a.startswith("-")]
    files = [a for a in args if not a.startswith("-")]
    ref = workspace.map_with(helpers.REF_BIN, "ref_bam_" + tag, pre + ["-i", "tests/rep.idx"] + files)
    got = workspace.map_with(helpers.CLI, "gpu_bam_" + tag, pre + ["-i", "tests/rep.idx"] + files, pre=["-B", "-t", "3"])
    want = helpers.sam_body(ref[0])
    lines = [ln for ln in helpers.bam_to_sam_lines(got[0]) if not ln.startswith("@PG")]
    assert len(lines) > 500 and lines == want
    assert open(ref[1]).read() == open(got[1]).read()


def test_cli_threads_and_batches_keep_reference_order(workspace):
    workspace.need_repeat()
    reads = ["-i", "tests/rep.idx", "tests/rep_pe_1.fq", "tests/rep_pe_2.fq"]
    ref = workspace.map_with(helpers.REF_BIN, "ref_t", reads)
    for k, pre in enumerate((["-t", "1"], ["-t", "8", "-gpu-batch", "333"], ["-t", "2", "-gpu-batch", "100000"])):
        got = workspace.map_with(helpers.CLI, "gpu_t%d" % k, reads, pre=pre)
        assert helpers.sam_body(ref[0]) == helpers.sam_body(got[0])
        assert open(ref[1]).read() == open(got[1]).read()
