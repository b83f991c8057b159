"""config #3 (PrefGP, 500 preferences over 1000 points, d = 4): fastUCBGallery of 4 under a few option settings."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ibo_oracle as orc
from ibo_b200 import _lib
from ibo_b200.acquisition import fastUCBGallery
from ibo_b200.gaussianprocess import PrefGaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
b3 = [[0., 10.]] * 4
P3 = np.array(orc.lhc_sample(b3, 1000, seed=2))
prefs = []
for i in range(500):
    a, b_ = P3[2 * i], P3[2 * i + 1]
    prefs.append((a, b_, 0) if -orc.shekel5(a) > -orc.shekel5(b_) else (b_, a, 0))
pg = PrefGaussianProcess(GaussianKernel_ard([5.146, 4.189, 4.622, 5.843]), prefs, noise=0.1)
def run(tag):
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fastUCBGallery(pg, b3, 4, seed=3); ts.append(1e3 * (time.perf_counter() - t0))
    print("%-28s %.2f ms (min of 5; all: %s)" % (tag, min(ts), " ".join("%.1f" % t for t in ts)), flush=True)
run("defaults")
_lib.set_option("tiny_server", 0); run("tiny_server=0"); _lib.set_option("tiny_server", 1)
_lib.set_option("i8_min_batch", 192); run("i8_min_batch=192"); _lib.set_option("i8_min_batch", 0); run("i8_min_batch=0"); _lib.set_option("i8_min_batch", -1)
run("defaults again")
_lib.set_option("direct_timing", 1)
fastUCBGallery(pg, b3, 4, seed=3)
