"""model.tsrn of text-gestalt (model/tsrn.py) on the focr engine"""
from fudanocr_b200.model.tsrn import *  # noqa: F401,F403
from fudanocr_b200.model.tsrn import TSRN  # noqa: F401
