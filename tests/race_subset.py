import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from fennec_b200 import api, synth as S
a = S.noise_image(520, 90, 1, alpha="random"); b = S.perturb(a, 2, 9)
print(api.SSIM(a, b), api.SSIMFast(S.noise_image(1100, 600, 3), S.noise_image(1100, 600, 4)))
api.GaussianBlur(a, 2.0); api.box_downsample(a, 100, 30); api.lanczos_resize(a, 100, 30); api.Sharpen(a, 0.5)
print("race subset ok")
