import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers, make_genome
from abismal_b200 import Index, IndexFile, Mapper, load_fastq
from abismal_b200.index_build import build_index_file
which = sys.argv[1]
d = tempfile.mkdtemp(dir="/tmp"); ws = helpers.Workspace(d)
make_genome.write_fasta(make_genome.random_genome(2_000_000), ws.path("g.fa"))
if "torch" in which:
    import torch
    torch.cuda.set_device(0)
    x = torch.zeros(10, device="cuda")
    print("torch ok")
if "build" in which:
    build_index_file(ws.path("g.fa"), ws.path("g.idx"))
else:
    ws.ref("idx", "tests/g.fa", "tests/g.idx")
ws.ref("sim", "-seed", "1", "-l", "100", "-n", "1000", "-o", "tests/r", "tests/g.fa")
ixf = IndexFile(ws.path("g.idx"))
ix = Index(ixf, 0)
m = Mapper(ix, mode=1, max_batch=1000, max_read_len=128)
b1, b2 = load_fastq(ws.path("r_1.fq")), load_fastq(ws.path("r_2.fq"))
try:
    r = m.map_batch(b1, b2)
    print(which, "OK", int((r.pe_r1["pos"] != 0).sum()))
except Exception as e:
    print(which, "FAIL", e)
