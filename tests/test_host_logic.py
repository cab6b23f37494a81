"""CPU-only: the host front end (option parsing, FASTQ rules, SAM/stats text)
exercised through oracle/oracle_map, which is the product's map_main.cpp and
host sources linked against the CPU restatement instead of the CUDA library."""
import gzip
import os
import subprocess

import pytest

import helpers


def run_tool(args, cwd):
    return subprocess.run([helpers.ORACLE_MAP] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_usage_errors_match_reference_exit_codes(workspace):
    workspace.need_trex()
    d = workspace.dir
    # missing -o: message + EXIT_SUCCESS (abismal.cpp:2360-2364)
    for tool in (helpers.REF_BIN, helpers.ORACLE_MAP):
        p = subprocess.run([tool, "map", "-i", "tests/tRex1.idx", "tests/reads_1.fq"], cwd=d, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0 and "Missing required argument" in p.stderr
    # -i and -g both absent
    for tool in (helpers.REF_BIN, helpers.ORACLE_MAP):
        p = subprocess.run([tool, "map", "-o", "tests/x.sam", "tests/reads_1.fq"], cwd=d, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0 and "Select one of index file" in p.stderr
    # missing FASTQ: EXIT_FAILURE
    for tool in (helpers.REF_BIN, helpers.ORACLE_MAP):
        p = subprocess.run([tool, "map", "-i", "tests/tRex1.idx", "-o", "tests/x.sam", "tests/nope.fq"], cwd=d,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 1 and "cannot open read 1 FASTQ file" in p.stderr
    # bad index file
    open(os.path.join(d, "tests", "bad.idx"), "wb").write(b"NotAnIndex--" + b"\0" * 64)
    for tool in (helpers.REF_BIN, helpers.ORACLE_MAP):
        p = subprocess.run([tool, "map", "-i", "tests/bad.idx", "-o", "tests/x.sam", "tests/reads_1.fq"], cwd=d,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 1 and "index file format problem" in p.stderr


def test_long_option_spellings(workspace):
    """OptionParser accepts -x, -long and the bare long name."""
    workspace.need_trex()
    a = workspace.map_with(helpers.ORACLE_MAP, "opt_a", ["-i", "tests/tRex1.idx", "tests/reads_1.fq"])[0]
    p = helpers.run([helpers.ORACLE_MAP, "map", "-index", "tests/tRex1.idx", "outfile", "tests/opt_b.sam",
                     "tests/reads_1.fq"], cwd=workspace.dir)
    b = os.path.join(workspace.dir, "tests/opt_b.sam")
    assert helpers.sam_body(a) == helpers.sam_body(b)
    p = subprocess.run([helpers.ORACLE_MAP, "map", "-i", "tests/tRex1.idx", "-o", "tests/x.sam", "-o", "tests/y.sam",
                        "tests/reads_1.fq"], cwd=workspace.dir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and "duplicate use of option" in p.stderr


def _write_fq(path, recs, gz=False):
    op = gzip.open if gz else open
    with op(path, "wt") as f:
        for name, seq in recs:
            f.write("@%s\n%s\n+\n%s\n" % (name, seq, "I" * len(seq)))


def test_fastq_rules_equal_reference(workspace):
    """Name cut at blank, N trimming, < 44 non-N bases skipped, CRLF, gz input, mate of a skipped read."""
    workspace.need_trex()
    from abismal_b200 import load_fastq
    src = load_fastq(workspace.path("reads_pe_1.fq"), 40)
    src2 = load_fastq(workspace.path("reads_pe_2.fq"), 40)
    r1 = [("r%d extra\tfield" % i, src.sequence(i)) for i in range(40) if src.sequence(i)]
    r2 = [("r%d" % i, src2.sequence(i)) for i in range(40) if src.sequence(i)]
    r1[0] = (r1[0][0], "NNNN" + r1[0][1][4:-3] + "NNN")
    r1[1] = (r1[1][0], "N" * 70 + r1[1][1][:30])          # fewer than 44 non-N: skipped, mate still mapped
    r1[2] = (r1[2][0], r1[2][1][:50] + "N" * 10 + r1[2][1][60:])
    r2[3] = (r2[3][0], "N" * 100)
    _write_fq(workspace.path("rules_1.fq"), r1)
    _write_fq(workspace.path("rules_2.fq"), r2)
    for tag, args in (("rules_pe", ["tests/rules_1.fq", "tests/rules_2.fq"]), ("rules_se", ["tests/rules_1.fq"])):
        ref = workspace.map_with(helpers.REF_BIN, "ref_" + tag, ["-i", "tests/tRex1.idx"] + args)
        got = workspace.map_with(helpers.ORACLE_MAP, "or_" + tag, ["-i", "tests/tRex1.idx"] + args)
        assert helpers.sam_body(ref[0]) == helpers.sam_body(got[0])
        assert open(ref[1]).read() == open(got[1]).read()
    # gz + CRLF input give the same records
    _write_fq(workspace.path("rules_1.fq.gz"), r1, gz=True)
    with open(workspace.path("rules_crlf_1.fq"), "w", newline="") as f:
        for name, seq in r1:
            f.write("@%s\r\n%s\r\n+\r\n%s\r\n" % (name, seq, "I" * len(seq)))
    base = helpers.sam_body(workspace.map_with(helpers.ORACLE_MAP, "g0", ["-i", "tests/tRex1.idx", "tests/rules_1.fq"])[0])
    for k, f in enumerate(("tests/rules_1.fq.gz", "tests/rules_crlf_1.fq")):
        out = helpers.sam_body(workspace.map_with(helpers.ORACLE_MAP, "g%d" % (k + 1), ["-i", "tests/tRex1.idx", f])[0])
        assert out[3:] == base[3:]
    # block mode of the FASTQ readers (-t 8 and up: several threads parse whole records of a mapped file or of
    # the zlib buffer): same records, same skips, for plain, gz and CRLF input, single and paired
    for k, f in enumerate(("tests/rules_1.fq", "tests/rules_1.fq.gz", "tests/rules_crlf_1.fq")):
        out = helpers.sam_body(workspace.map_with(helpers.ORACLE_MAP, "h%d" % k, ["-i", "tests/tRex1.idx", f],
                                                  pre=["-t", "16", "-gpu-batch", "7"])[0])
        assert out[3:] == base[3:]
    ref = workspace.map_with(helpers.REF_BIN, "ref_rules_pe", ["-i", "tests/tRex1.idx", "tests/rules_1.fq", "tests/rules_2.fq"])
    got = workspace.map_with(helpers.ORACLE_MAP, "or_rules_pe_t16",
                             ["-i", "tests/tRex1.idx", "tests/rules_1.fq", "tests/rules_2.fq"], pre=["-t", "16"])
    assert helpers.sam_body(ref[0]) == helpers.sam_body(got[0]) and open(ref[1]).read() == open(got[1]).read()


def test_pe_count_mismatch_is_an_error(workspace):
    workspace.need_trex()
    from abismal_b200 import load_fastq
    src = load_fastq(workspace.path("reads_pe_1.fq"), 10)
    recs = [("r%d" % i, src.sequence(i) or "A" * 100) for i in range(10)]
    _write_fq(workspace.path("mm_1.fq"), recs)
    _write_fq(workspace.path("mm_2.fq"), recs[:7])
    p = run_tool(["map", "-i", "tests/tRex1.idx", "-o", "tests/mm.sam", "tests/mm_1.fq", "tests/mm_2.fq"], workspace.dir)
    assert p.returncode == 1 and "paired-end batch sizes differ" in p.stderr


def test_bam_output_decodes_to_the_sam_records(workspace):
    """-B: BGZF framing is valid and every BAM record decodes to the SAM line of the same run (SE and PE)."""
    workspace.need_trex()
    for tag, args in (("bam_se", ["tests/reads_1.fq"]), ("bam_pe", ["tests/reads_pe_1.fq", "tests/reads_pe_2.fq"])):
        sam = workspace.map_with(helpers.ORACLE_MAP, tag, ["-i", "tests/tRex1.idx"] + args)
        bam = workspace.map_with(helpers.ORACLE_MAP, tag + "_b", ["-i", "tests/tRex1.idx"] + args, pre=["-B"])
        want = helpers.sam_body(sam[0])
        got = [ln for ln in helpers.bam_to_sam_lines(bam[0]) if not ln.startswith("@PG")]
        assert len(got) > 1000 and got == want
        assert open(sam[1]).read() == open(bam[1]).read()


def test_pipeline_batching_threads_and_sharding_do_not_change_the_output(workspace):
    """Small batches, several formatter threads and two mapper workers (the multi-GPU path of the front end,
    here two CPU engines) give byte-identical SAM and stats in input order."""
    workspace.need_trex()
    args = ["-i", "tests/tRex1.idx", "tests/reads_pe_1.fq", "tests/reads_pe_2.fq"]
    base = workspace.map_with(helpers.ORACLE_MAP, "pipe_a", args, pre=["-t", "1"])
    for k, pre in enumerate((["-t", "4", "-gpu-batch", "37"], ["-t", "3", "-gpu-batch", "501", "-gpus", "2"],
                             ["-gpu-batch", "10000"], ["-gpu-batch", "1", "-t", "2", "-gpus", "3"])):
        if k == 3:  # one read per batch: keep it short
            for e in (1, 2):
                with open(workspace.path("reads_pe_%d.fq" % e)) as f, open(workspace.path("few_%d.fq" % e), "w") as g:
                    g.writelines(f.readlines()[:4 * 300])
            few = ["-i", "tests/tRex1.idx", "tests/few_1.fq", "tests/few_2.fq"]
            a = workspace.map_with(helpers.ORACLE_MAP, "pipe_few_a", few)
            b = workspace.map_with(helpers.ORACLE_MAP, "pipe_few_b", few, pre=pre)
            assert helpers.sam_body(a[0]) == helpers.sam_body(b[0]) and open(a[1]).read() == open(b[1]).read()
            continue
        got = workspace.map_with(helpers.ORACLE_MAP, "pipe_%d" % k, args, pre=pre)
        assert helpers.sam_body(got[0]) == helpers.sam_body(base[0])
        assert open(got[1]).read() == open(base[1]).read()


def test_index_loaded_by_the_bulk_reader_gives_the_same_output(workspace, monkeypatch):
    """The product's index loader (the arrays of the file in one buffer, filled by several pread threads; the
    header parsed as before) behind the same host code: byte-identical SAM and statistics.  The test tool reads
    into vectors unless ABISMAL_B200_INDEX_BULK is set."""
    workspace.need_trex()
    args = ["-i", "tests/tRex1.idx", "tests/reads_pe_1.fq", "tests/reads_pe_2.fq"]
    base = workspace.map_with(helpers.ORACLE_MAP, "bulk_a", args)
    monkeypatch.setenv("ABISMAL_B200_INDEX_BULK", "1")
    got = workspace.map_with(helpers.ORACLE_MAP, "bulk_b", args)
    assert helpers.sam_body(got[0]) == helpers.sam_body(base[0])
    assert open(got[1]).read() == open(base[1]).read()
    # a truncated index file is refused with the reference's message, not read past its end
    with open(workspace.path("tRex1.idx"), "rb") as f:
        data = f.read()
    with open(workspace.path("cut.idx"), "wb") as f:
        f.write(data[:len(data) - 1000])
    p = helpers.run([helpers.ORACLE_MAP, "map", "-i", "tests/cut.idx", "-o", "tests/cut.sam", "tests/reads_pe_1.fq",
                     "tests/reads_pe_2.fq"], cwd=workspace.dir, check=False)
    assert p.returncode != 0 and "failed loading index file" in p.stderr
