"""loss.weight_ce_loss of scene-text-telescope (loss/weight_ce_loss.py) on the focr engine"""
from fudanocr_b200.loss.weight_ce_loss import *  # noqa: F401,F403
from fudanocr_b200.loss.weight_ce_loss import load_confuse_matrix, weight_cross_entropy, standard_alphebet  # noqa: F401
