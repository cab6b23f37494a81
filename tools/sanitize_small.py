"""A small batch of every mode through the CUDA path, for compute-sanitizer (memcheck / racecheck / initcheck):
  compute-sanitizer --tool memcheck python tools/sanitize_small.py
Checks the records against the CPU oracle as well (checker only)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import make_genome  # noqa: E402
from abismal_b200 import Index, IndexFile, Mapper, load_fastq  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
with tempfile.TemporaryDirectory() as d:
    ws = helpers.Workspace(d)
    make_genome.write_fasta(make_genome.repeat_genome(scale=0.1), ws.path("g.fa"))
    ws.ref("idx", "tests/g.fa", "tests/g.idx")
    ws.ref("sim", "-seed", "9", "-l", "150", "-min-fraglen", "150", "-max-fraglen", "400", "-n", str(n),
           "-m", "0.02", "-b", "0.98", "-o", "tests/r", "tests/g.fa")
    ixf = IndexFile(ws.path("g.idx"))
    b1, b2 = load_fastq(ws.path("r_1.fq")), load_fastq(ws.path("r_2.fq"))
    ix = Index(ixf, 0)
    for mode in (1, 1 | 2, 1 | 4, 0, 4):
        m = Mapper(ix, mode=mode, max_batch=b1.n, max_read_len=160)
        o = helpers.OracleMapper(ixf, mode=mode)
        b = (b1, b2) if mode & 1 else (b1,)
        helpers.assert_results_equal(m.map_batch(*b), o.map_batch(*b), bool(mode & 1))
        m.upload(*b)
        m.run()
        m.sync()
        print("mode %d: %d items bit-exact vs oracle" % (mode, b1.n), flush=True)
        m.close()
        o.close()
    ix.close()
