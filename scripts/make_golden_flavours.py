#!/usr/bin/env python
"""tests/golden/transformer_flavours.npz from the REFERENCE's own transformers module (run in the build container,
where /root/reference exists): the deep_variant and joint_callset flavours of get_transformer
(ugbio_filtering/transformers.py:221-245,278) fitted and applied to a DeepVariant-style synthetic call set
(tests/dv_data.py; the frame comes from the oracle's loader, as for transformer_single_sample.npz).

  vcf_text               header + records
  customs                custom annotation tags
  features_deep_variant  get_transformer(DEEP_VARIANT, annots).fit_transform(df)
  features_joint         get_transformer(JOINT, annots).fit_transform(df)
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/ugbio_utils/src"
sys.path.insert(0, os.path.join(REF, "filtering"))
sys.path.insert(0, os.path.join(REF, "core"))
pd.DataFrame.applymap = pd.DataFrame.map  # pandas-3 harness shim

from ugbio_filtering import transformers as ref_t  # noqa: E402
from ugbio_filtering.tprep_constants import VcfType as RefVcfType  # noqa: E402

from oracle import ref_pipeline as R  # noqa: E402
from oracle.vcf_reader import OracleVariantFile  # noqa: E402
from tests import dv_data  # noqa: E402


def main():
    ds = dv_data.generate(700, seed=23)
    text = ds["header_text"].encode() + ds["text"]
    df = R.harness_float_columns(R.get_vcf_df(OracleVariantFile(text), None, ds["customs"]))
    annots = [c.lower() for c in ds["customs"]]
    out = {}
    for name, vt in (("deep_variant", RefVcfType.DEEP_VARIANT), ("joint", RefVcfType.JOINT)):
        tr = ref_t.get_transformer(vt, annots)
        with pd.option_context("future.infer_string", False):
            out[name] = tr.fit_transform(df).to_numpy(dtype=np.float64)
        print(name, out[name].shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "transformer_flavours.npz"),
                        vcf_text=np.frombuffer(text, dtype=np.uint8), customs=np.array(ds["customs"]),
                        features_deep_variant=out["deep_variant"], features_joint=out["joint"])


if __name__ == "__main__":
    main()
