"""GPU index construction must be byte-identical to the reference's `abismal idx`
(golden md5 of tRex1.idx in data/md5sum.txt, and the reference binary on the
repeat-rich genome with N runs, IUPAC codes and several chromosomes)."""
import os

import numpy as np
import pytest

import helpers
from test_oracle_vs_ref import golden_md5

pytestmark = pytest.mark.gpu


def test_trex_index_matches_golden_md5(workspace):
    from abismal_b200.index_build import build_index_file
    workspace.need_trex()
    out = workspace.path("tRex1.gpu.idx")
    build_index_file(workspace.path("tRex1.fa"), out)
    assert helpers.md5(out) == golden_md5()["tests/tRex1.idx"]


def test_repeat_genome_index_matches_reference_binary(workspace):
    from abismal_b200.index_build import build_index_file
    workspace.need_repeat()
    out = workspace.path("rep.gpu.idx")
    build_index_file(workspace.path("rep.fa"), out)
    assert helpers.md5(out) == helpers.md5(workspace.path("rep.idx"))


def test_random_genome_index_matches_reference_binary(workspace):
    import make_genome
    from abismal_b200.index_build import build_index_file
    fa = workspace.path("rnd.fa")
    make_genome.write_fasta(make_genome.random_genome(6_000_000, n_chroms=3, seed=3), fa)
    workspace.ref("idx", "-t", "2", "tests/rnd.fa", "tests/rnd.idx")
    out = workspace.path("rnd.gpu.idx")
    build_index_file(fa, out)
    assert helpers.md5(out) == helpers.md5(workspace.path("rnd.idx"))
