"""loss.stroke_focus_loss of text-gestalt (loss/stroke_focus_loss.py) on the focr engine"""
from fudanocr_b200.loss.stroke_focus_loss import *  # noqa: F401,F403
from fudanocr_b200.loss.stroke_focus_loss import StrokeFocusLoss, to_gray_tensor  # noqa: F401
