"""GPU: the classify kernel (csrc/concordance.cu conc_classify_gt through ugvc_conc_classify) against the rows the
reference's own functions produced (tests/golden/classify_rules.json) and against the oracle on random genotypes."""
import numpy as np
import pytest

from oracle import classify_ref as CR
from tests.test_classify_cpu import load
from variantcalling_b200 import concordance as CC

pytestmark = pytest.mark.gpu


def test_kernel_equals_the_reference_rows():
    gu, gt, base, want_c, want_g = load()
    ctx = CC.ConcordanceContext(0)
    c, g = ctx.classify(gu, gt, base)
    assert list(c) == want_c and list(g) == want_g
    ctx.close()


def test_kernel_equals_the_oracle_on_random_frames():
    rng = np.random.default_rng(17)
    vals = [None, 0, 1, 2, 3]
    n = 200_000
    def draw():
        return [tuple(vals[int(k)] for k in rng.integers(0, 5, size=int(rng.integers(1, 3)))) for _ in range(n)]
    gu, gt = draw(), draw()
    base = [("FN", "FN_CA", "TP", None)[int(k)] for k in rng.integers(0, 4, size=n)]
    ctx = CC.ConcordanceContext(0)
    c, g = ctx.classify(gu, gt, base)
    want_c, want_g = CR.classify_records(gu, gt, base)
    assert list(c) == want_c and list(g) == want_g
    ctx.close()
