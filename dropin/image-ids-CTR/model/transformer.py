"""model.transformer of image-ids-CTR (model/transformer.py) on the focr engine"""
from fudanocr_b200.model.ids_transformer import *  # noqa: F401,F403
from fudanocr_b200.model.ids_transformer import Transformer  # noqa: F401
