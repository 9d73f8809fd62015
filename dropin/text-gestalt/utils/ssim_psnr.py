"""utils.ssim_psnr of text-gestalt (utils/ssim_psnr.py) on the focr engine"""
from fudanocr_b200.utils.ssim_psnr import *  # noqa: F401,F403
from fudanocr_b200.utils.ssim_psnr import SSIM, calculate_psnr, ssim  # noqa: F401
