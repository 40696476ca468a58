"""Kept for ``make_golden.py`` and older imports: the loader lives in ``oracle/refload.py``."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.refload import *  # noqa: F401,F403,E402
from oracle.refload import REFERENCE_ROOT, available, build_reference_wf, load  # noqa: F401,E402
