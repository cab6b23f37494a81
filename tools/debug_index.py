import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers, make_genome
from abismal_b200 import IndexFile
from abismal_b200.index_build import build_index_file
d = tempfile.mkdtemp(dir="/tmp"); ws = helpers.Workspace(d)
make_genome.write_fasta(make_genome.repeat_genome(), ws.path("rep.fa"))
ws.ref("idx", "-t", "4", "tests/rep.fa", "tests/rep.idx")
build_index_file(ws.path("rep.fa"), ws.path("gpu.idx"))
a, b = IndexFile(ws.path("rep.idx")), IndexFile(ws.path("gpu.idx"))
print("names", a.names == b.names, "starts", np.array_equal(a.starts, b.starts))
for k in ("genome", "counter", "counter_t", "counter_a", "index", "index_t", "index_a"):
    x, y = getattr(a, k), getattr(b, k)
    if x.shape != y.shape:
        print(k, "shape", x.shape, y.shape); continue
    bad = np.nonzero(x != y)[0]
    print(k, "equal" if bad.size == 0 else "DIFF at %d of %d, first %s" % (bad.size, x.size, bad[:5]))
    if bad.size and k == "genome":
        for w in bad[:3]:
            print(hex(int(x[w])), hex(int(y[w])), "bases", w * 16)
    if bad.size and k.startswith("index"):
        cn = {"index": a.counter, "index_t": a.counter_t, "index_a": a.counter_a}[k]
        w = bad[0]; bk = np.searchsorted(cn, w, side="right") - 1
        s, e = cn[bk], cn[bk + 1]
        print(" bucket", bk, "size", e - s, "ref", x[s:s + 12], "gpu", y[s:s + 12], "same set", np.array_equal(np.sort(x[s:e]), np.sort(y[s:e])))
print("sizes", a.index_size, b.index_size, a.index_size_three, b.index_size_three)
