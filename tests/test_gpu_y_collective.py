"""GPU: the path's one collective through the C ABI alone (ugvc_nccl_unique_id / ugvc_nccl_comm_init /
ugvc_counts_allreduce, NCCL bound at run time): a one-rank communicator here (the driver's N > 1 runs go through the
same entry points in scripts/run_multi_gpu_cli.py); the reduced block equals the context's counters."""
import numpy as np
import pytest

from tests import util
from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu


def test_counts_allreduce_through_the_c_abi(gpu_ctx):
    ds = util.make_dataset(n_records=2000, n_custom=2, seed=8)
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    gpu_ctx.load_plan(MC.compile_plan(VcfHeader(ds["header_text"]), tr, model, ds["customs"]).blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    gpu_ctx.reserve(len(ds["text"]) + 4096, len(ds["lines"]) + 16, 1)
    gpu_ctx.counts_reset()
    res = gpu_ctx.filter_batch(ds["text"], 30.0)
    want = gpu_ctx.counts()
    uid = lib.Context.nccl_unique_id()
    assert len(uid) == 128 and any(uid)
    comm = gpu_ctx.nccl_comm_init(uid, 1, 0)
    try:
        got = gpu_ctx.counts_allreduce(comm)
    finally:
        gpu_ctx.nccl_comm_destroy(comm)
    assert got == want and got["n_records"] == res["n_records"] == len(ds["lines"])
    assert got["n_low_score"] == int(res["low_score"].sum()) and got["n_pass"] == got["n_records"] - got["n_low_score"]
