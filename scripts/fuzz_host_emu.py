#!/usr/bin/env python
"""Differential fuzz of K1 / K2 against the oracle on the HOST EMULATION of the kernel sources
(tests/host_emu; build it with `make -C tests/host_emu`): per seed 4000 records with shuffled /
unknown INFO keys, permuted FORMAT columns, re-spelled and exotic numeric literals, compared feature
by feature in the generic and the learned-key-order mode.   python scripts/fuzz_host_emu.py 0 60"""
import os, sys, warnings, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["UGVC_LIB_PATH"] = os.environ.get("UGVC_EMU_LIB") or os.path.join(ROOT, "tests", "host_emu", "_build", "libugvc_emu.so")
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, pandas as pd
from oracle import ref_pipeline as R
from oracle.vcf_reader import OracleVariantFile
from tests import util
from tests.test_gpu_fuzz import mutate, FLOAT_KEYS
from variantcalling_b200 import lib, model_compiler as MC
from variantcalling_b200.vcf_header import VcfHeader
EXOTIC=["nan","NaN","1e-50","1e-46","1.4e-45","3.4028235e38","3.4028234e38","7e-46","0.1e1","1.e0",".5","5.","-.5e-1","+0.0","-0","123456789012345678","0.30000000000000004","16777217","1e+2","1E2","1e-0","00000000000000000001.5","0.000000000000000000015e20"]
def exotic(rng,line):
    c=line.split("\t"); info=c[7].split(";"); out=[]
    for kv in info:
        k,_,v=kv.partition("=")
        if k in FLOAT_KEYS and rng.random()<0.25: kv=k+"="+EXOTIC[rng.integers(0,len(EXOTIC))]
        out.append(kv)
    c[7]=";".join(out)
    if rng.random()<0.1: c[5]=EXOTIC[rng.integers(2,len(EXOTIC))]
    return "\t".join(c)
ctx=lib.Context(0)
tot=0; t0=time.time()
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng=np.random.default_rng(1000+seed)
    ds=util.make_dataset(n_records=4000,n_custom=8,seed=500+seed)
    _,tr,x=util.fit_transformer(ds)
    model=util.fit_model("lr",x,ds["labels"])
    lines=[exotic(rng,mutate(rng,l)) for l in ds["lines"]]
    text=("\n".join(lines)+"\n").encode()
    df=R.harness_float_columns(R.get_vcf_df(OracleVariantFile(ds["header_text"].encode()+text),None,ds["customs"]))
    with pd.option_context("future.infer_string", False):
        want=tr.transform(df).to_numpy(dtype=np.float64).astype(np.float32)
    plan=MC.compile_plan(VcfHeader(ds["header_text"]),tr,model,ds["customs"])
    ctx.load_plan(plan.blob); ctx.reserve(len(text)+64,len(lines)+8,1)
    for mode in ("generic","learned"):
        if mode=="learned": ctx.set_key_order(*lib.learn_key_order(text))
        else: ctx.set_key_order("", "")
        try:
            res=ctx.filter_batch(text)
        except lib.UgvcDataError as e:
            print("seed",seed,mode,"DATA ERROR",e, ctx.last_data_error()); 
            r=ctx.last_data_error()[0]; print(lines[r]); break
        got=ctx.debug_features(res["n_records"]).T
        eq=(got==want)|(np.isnan(got)&np.isnan(want))
        bad=np.argwhere(~eq)
        if bad.size:
            print("seed",seed,mode,len(bad),"mismatches; first", bad[0], plan.feature_names[bad[0][1]], got[tuple(bad[0])], want[tuple(bad[0])]); print(lines[bad[0][0]]); break
    tot+=len(lines)
print("checked",tot,"records in",round(time.time()-t0),"s")
