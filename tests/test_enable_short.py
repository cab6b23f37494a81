"""CPU-only: the --enable-short variant of the reference (seed::window_size 12, src/AbismalIndex.hpp:73-77).

oracle/_ref/abismal_short is the unmodified reference compiled with -DENABLE_SHORT.  The CPU restatement
takes the window from the index file; here it must reproduce that binary's SAM and stats on a window-12 index
(reads of 50, 60 and 100 bases; the default binary would skip nothing here either, but seeds differently)."""
import pytest

import helpers

CASES = [
    ("w12_se", ["-i", "tests/rep_w12.idx", "tests/w12_se_1.fq"]),
    ("w12_se_A", ["-A", "-i", "tests/rep_w12.idx", "tests/w12_se_1.fq"]),
    ("w12_pe", ["-i", "tests/rep_w12.idx", "tests/w12_pe_1.fq", "tests/w12_pe_2.fq"]),
    ("w12_pe_R_a", ["-R", "-a", "-i", "tests/rep_w12.idx", "tests/w12_rpe_1.fq", "tests/w12_rpe_2.fq"]),
]


def test_reference_builds_differ_in_window(workspace):
    """The two reference builds refuse each other's index files: the window is part of the format."""
    workspace.need_short()
    p = helpers.run([helpers.REF_BIN, "map", "-i", "tests/rep_w12.idx", "-o", "tests/x.sam", "tests/w12_se_1.fq"],
                    cwd=workspace.dir, check=False)
    assert p.returncode != 0 and "window size" in p.stderr


@pytest.mark.parametrize("tag,args", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_short_reference(workspace, tag, args):
    workspace.need_short()
    rsam, rst, _ = workspace.map_with(helpers.REF_BIN_SHORT, "ref_" + tag, args)
    osam, ost, _ = workspace.map_with(helpers.ORACLE_MAP, "or_" + tag, args)
    assert helpers.sam_body(rsam) == helpers.sam_body(osam)
    assert open(rst).read() == open(ost).read()
    assert len(helpers.sam_body(rsam)) > 100


def test_window_changes_the_result(workspace):
    """Sanity: the window-12 and window-20 indexes of the same genome are different files, and mapping the
    same short reads through them does not give the same SAM (so the tests above do exercise the window)."""
    workspace.need_short()
    workspace.need_repeat()
    assert helpers.md5(workspace.path("rep_w12.idx")) != helpers.md5(workspace.path("rep.idx"))


def test_index_readers_follow_the_window_in_the_file(workspace, tmp_path):
    """Python and C++ readers accept window 12 and 20 and refuse anything else (seed::read,
    src/AbismalIndex.cpp:1005-1013, accepts only the window the binary was configured with)."""
    import struct
    from abismal_b200 import IndexFile
    workspace.need_short()
    assert IndexFile(workspace.path("rep_w12.idx")).window_size == 12
    bad = tmp_path / "w13.idx"
    with open(workspace.path("rep_w12.idx"), "rb") as f:
        blob = bytearray(f.read(4096))
    blob[16:20] = struct.pack("<I", 13)
    bad.write_bytes(bytes(blob))
    with pytest.raises(ValueError):
        IndexFile(str(bad))
    p = helpers.run([helpers.ORACLE_MAP, "map", "-i", str(bad), "-o", str(tmp_path / "x.sam"),
                     workspace.path("w12_se_1.fq")], check=False)
    assert p.returncode != 0 and "window size" in p.stderr
