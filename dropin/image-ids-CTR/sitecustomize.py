"""Put this directory on PYTHONPATH and the reference scripts of this subproject import the focr sm_100a backend for the hot-path
modules listed beside this file - `python main.py ...` / `python train.py` run unedited (see dropin/README.md)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(os.path.dirname(_here))
if _root not in sys.path:
    sys.path.append(_root)          # the repository root: `fudanocr_b200` itself
from fudanocr_b200.dropin import install  # noqa: E402

install(_here)
