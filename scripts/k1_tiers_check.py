import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, os
from tests import util
from variantcalling_b200 import lib, model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader
ds = util.make_dataset(n_records=3000, n_custom=40, seed=5)
_, tr, x = util.fit_transformer(ds)
model = util.fit_model("gb_small", x, ds["labels"])
plan = MC.compile_plan(VcfHeader(ds["header_text"]), tr, model, ds["customs"])
ctx = lib.Context(0); ctx.load_plan(plan.blob); ctx.reserve(len(ds["text"])+1024, 4096, 1)
for mode in (["learn"] if os.environ.get("MODE") == "learn" else ["nolearn","learn"]):
    if mode=="learn": ctx.set_key_order(*lib.learn_key_order(ds["text"]))
    out = ctx.filter_batch(ds["text"])
    n = out["n_records"]
    raw = ctx.debug_raw(n)
    print(mode, "n", n, "slow", ctx.slow_records(0))
    if mode=="nolearn" or os.environ.get("MODE") == "learn": raw0=raw.copy()
print("raw equal", np.array_equal(raw0, raw))
np.save("/tmp/raw_%s.npy" % os.environ.get("TAG","x"), raw)
