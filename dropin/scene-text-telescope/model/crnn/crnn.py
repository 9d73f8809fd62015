"""model.crnn.crnn of scene-text-telescope (model/crnn/crnn.py) on the focr engine"""
from fudanocr_b200.model.crnn.crnn import *  # noqa: F401,F403
from fudanocr_b200.model.crnn.crnn import CRNN  # noqa: F401
