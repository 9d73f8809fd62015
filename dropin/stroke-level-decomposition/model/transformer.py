"""model.transformer of stroke-level-decomposition (model/transformer.py) on the focr engine"""
from fudanocr_b200.model.transformer import *  # noqa: F401,F403
from fudanocr_b200.model.transformer import Transformer  # noqa: F401
