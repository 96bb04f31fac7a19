import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from time_conv import time_conv
for n in (8, 70):
    time_conv(n, 256, 256, 66, 66, 3, False, iters=10)
    time_conv(n, 256, 256, 66, 66, 3, True, iters=10)
time_conv(8, 128, 128, 130, 130, 3, True)
time_conv(8, 256, 512, 34, 34, 4, True)
