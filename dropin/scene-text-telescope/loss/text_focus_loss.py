"""loss.text_focus_loss of scene-text-telescope (loss/text_focus_loss.py) on the focr engine"""
from fudanocr_b200.loss.text_focus_loss import *  # noqa: F401,F403
from fudanocr_b200.loss.text_focus_loss import TextFocusLoss, str_filt  # noqa: F401
from fudanocr_b200.loss.stroke_focus_loss import to_gray_tensor  # noqa: F401
from fudanocr_b200.loss.weight_ce_loss import weight_cross_entropy  # noqa: F401
