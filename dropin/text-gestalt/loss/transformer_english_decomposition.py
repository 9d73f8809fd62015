"""loss.transformer_english_decomposition of text-gestalt on the focr engine (parameter container of the frozen recogniser)"""
from fudanocr_b200.loss.transformer_english_decomposition import *  # noqa: F401,F403
from fudanocr_b200.loss.transformer_english_decomposition import Transformer  # noqa: F401
