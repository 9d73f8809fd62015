"""model.tbsrn of text-gestalt (model/tbsrn.py) on the focr engine"""
from fudanocr_b200.model.tbsrn import *  # noqa: F401,F403
from fudanocr_b200.model.tbsrn import TBSRN  # noqa: F401
