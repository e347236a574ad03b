"""GPU: device-resident batches on different lanes and streams (ugvc_filter_device_lane) give the results of the
serial lane-0 path: every lane has its own scratch, so two batches may be in flight at once."""
import numpy as np
import pytest

from tests import util
from variantcalling_b200 import lib
from variantcalling_b200 import model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader

pytestmark = pytest.mark.gpu


def test_two_lanes_in_flight_equal_the_serial_path(gpu_ctx):
    torch = pytest.importorskip("torch")
    ds = util.make_dataset(n_records=6000, n_custom=3, seed=21)
    _, tr, x = util.fit_transformer(ds)
    model = util.fit_model("gb_small", x, ds["labels"])
    gpu_ctx.load_plan(MC.compile_plan(VcfHeader(ds["header_text"]), tr, model, ds["customs"]).blob)
    gpu_ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    lines = ds["lines"]
    halves = [("\n".join(lines[:3000]) + "\n").encode(), ("\n".join(lines[3000:]) + "\n").encode()]
    gpu_ctx.reserve(max(len(h) for h in halves) + 4096, 3100, 2)
    K = gpu_ctx.n_classes
    dev = torch.device("cuda", 0)
    texts = [torch.from_numpy(np.frombuffer(h + b"\n" * 64, dtype=np.uint8).copy()).to(dev) for h in halves]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    low = [torch.zeros(3100, dtype=torch.uint8, device=dev) for _ in range(2)]
    probs = [torch.zeros((3100, K), dtype=torch.float32, device=dev) for _ in range(2)]
    qual = [torch.zeros(3100, dtype=torch.float64, device=dev) for _ in range(2)]
    nrec = torch.zeros(2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for rounds in range(3):  # both lanes enqueued before either is waited for
        for ln in range(2):
            gpu_ctx.filter_device(texts[ln].data_ptr(), len(halves[ln]), 30.0, low[ln].data_ptr(), probs[ln].data_ptr(),
                                  qual[ln].data_ptr(), 3100, d_n_records=nrec.data_ptr() + 8 * ln,
                                  stream=streams[ln].cuda_stream, lane=ln)
    for ln in range(2):
        gpu_ctx.device_status(streams[ln].cuda_stream, lane=ln)
    torch.cuda.synchronize()
    assert nrec.tolist() == [3000, 3000]
    for ln in range(2):
        want = gpu_ctx.filter_batch(halves[ln], 30.0)
        assert np.array_equal(low[ln][:3000].cpu().numpy(), want["low_score"])
        assert np.array_equal(probs[ln][:3000].cpu().numpy(), want["probs"])
        assert np.array_equal(qual[ln][:3000].cpu().numpy(), want["qual"])
