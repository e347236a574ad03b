"""CPU: the C-ABI library loads, exports every symbol include/ugvc_b200.h declares, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import pytest

from variantcalling_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ugvc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ugvc_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    L = lib.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ugvc_b200.h but not exported"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in variantcalling_b200/lib.py"
    assert L.ugvc_version() >= 100


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "variantcalling_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
    for fn in ("ugvc/__main__.py",):
        assert "oracle" not in open(os.path.join(ROOT, fn)).read()


def test_no_cuda_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(lib.UgvcError) as ei:
        lib.Context(0)
    assert "no CPU fallback" in str(ei.value)
    # the --treat_multiallelics kernels and the model-apply step refuse the same way
    from variantcalling_b200 import multiallelics
    from variantcalling_b200.vcf_header import VcfHeader

    hdr = VcfHeader("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
    with pytest.raises(lib.UgvcError):
        multiallelics.DeviceSplitPlan(hdr, {}, "ACGT")


def test_synth_header_host_call():
    h = lib.synth_header(40)
    assert h.startswith("##fileformat=VCFv4.2") and h.rstrip().endswith("SAMPLE1")
    assert h.count("##INFO=<ID=ANN") == 35 and "##contig=<ID=chrY" in h


def test_learn_key_order():
    from variantcalling_b200 import synth

    _, lines, _ = synth.generate(synth.SynthSpec(n_records=1500, n_custom=40))
    info, fmt = lib.learn_key_order(("\n".join(lines) + "\n").encode())
    keys = info.split(";")
    assert keys[:5] == ["AC", "AF", "AN", "BaseQRankSum", "DP"] and fmt == "GT:AD:DP:GQ:PL"
    assert keys[28:33] == synth.BASE_CUSTOM and keys[-1] == "ANN34" and len(keys) == 68
    dense = lib.learn_key_order(("\n".join(lines) + "\n").encode(), min_presence=0.5)[0].split(";")
    assert dense[-1] == "X_RM" and len(dense) == 28        # the 15 %-present annotations can be left out
    # inconsistent orders are left to the generic path; valueless keys carry a '!' marker
    assert lib.learn_key_order(b"c\t1\t.\tA\tC\t1\t.\tB=1;A=2\tGT\t0/1\nc\t2\t.\tA\tC\t1\t.\tA=2;B=1\tGT\t0/1\n")[0] == ""
    assert lib.learn_key_order(b"c\t1\t.\tA\tC\t1\t.\tA=2;DB;Z=1\n")[0] == "A;DB!;Z"
