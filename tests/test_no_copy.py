"""Copy hygiene: no product source may be a re-typed reference module.

``tools/overlap.py`` strips docstrings / comments / whitespace and counts the product lines that occur
verbatim anywhere in the reference package.  What remains shared after the host glue was rewritten are
protocol-forced lines (method signatures such as ``def make_irreducible(self, e, vec, mask=None):``,
``import numpy as np``, dictionary keys); the bound leaves room for those and nothing else.  Skipped
where the reference tree is absent (the GPU box).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = "/root/reference/pyqmc"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_product_python_sources_are_not_copies():
    import warnings

    import overlap

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rows = overlap.report(os.path.join(ROOT, "pyqmc_b200"), REF)
    worst = [(round(100 * frac, 1), name) for frac, shared, n, name in rows if n >= 20 and frac >= 0.15]
    assert not worst, f"files sharing >= 15 % of their code lines with the reference: {worst}"
