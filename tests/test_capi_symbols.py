"""CPU-only: the C-ABI library loads (nvcc cross-compiled, no GPU needed to
dlopen it) and exports every function the headers in include/ declare.  No
compute call is made here."""
import ctypes
import os
import re

import helpers


def declared_functions():
    names = []
    inc = os.path.join(helpers.ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(inc, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(abg_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_headers_declare_the_expected_entry_points():
    fns = declared_functions()
    for must in ("abg_index_create", "abg_mapper_create", "abg_map_batch", "abg_mapper_upload", "abg_mapper_run",
                 "abg_mapper_download", "abg_last_error", "abg_build_index"):
        assert must in fns


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(os.path.join(helpers.ROOT, "abismal_b200", "libabismal_b200.so"))
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, missing


def test_python_binding_struct_sizes_match_header(built, tmp_path):
    """sizeof() as gcc sees the header == sizeof() of the ctypes mirror."""
    import subprocess
    from abismal_b200 import capi
    from abismal_b200.index_build import abg_built_index
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "abismal_b200.h"\n#include "abismal_b200_index.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(abg_index_view), sizeof(abg_params),'
                   'sizeof(abg_batch), sizeof(abg_results), sizeof(abg_work_counters), sizeof(abg_hit),'
                   'sizeof(abg_built_index)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(helpers.ROOT, "include"), "-o", str(exe), str(src)])
    want = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    got = [ctypes.sizeof(capi.abg_index_view), ctypes.sizeof(capi.abg_params), ctypes.sizeof(capi.abg_batch),
           ctypes.sizeof(capi.abg_results), ctypes.sizeof(capi.abg_work_counters), capi.HIT_DTYPE.itemsize,
           ctypes.sizeof(abg_built_index)]
    assert got == want


def test_no_cuda_device_fails_loudly(built):
    """On a box without a GPU the product must refuse, not fall back."""
    from abismal_b200 import capi
    lib = capi.load_library()
    if lib.abg_device_count() > 0:
        return  # GPU box: nothing to check here
    import numpy as np

    class Fake:
        genome = np.zeros(4, "<u8"); genome_size = 64
        counter = np.zeros((1 << 25) + 1, "<u4"); counter_size = 1 << 25
        counter_t = counter_a = np.zeros(43046722, "<u4"); counter_size_three = 43046721
        index = index_t = index_a = np.zeros(1, "<u4"); index_size = 0; index_size_three = 0
        max_candidates = 100
    try:
        capi.Index(Fake(), 0)
    except capi.AbgError as e:
        assert "cuda" in str(e).lower() or "device" in str(e).lower()
    else:
        raise AssertionError("Index() succeeded without a CUDA device")
