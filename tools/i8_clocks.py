import sys, subprocess, time, threading, numpy as np
sys.path.insert(0, '/root/repo')
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
N, d, M = 2048, 6, 1 << 20
rs = np.random.RandomState(0)
X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
gp = GaussianProcess(GaussianKernel_ard([0.5] * d), X, Y, noise=0.1)
m = gp.model
c = _lib.ResidentCandidates(m, np.ascontiguousarray(rs.rand(M, d)))
def sample(tag, fl, secs):
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    t0 = time.time(); ts = []
    while time.time() - t0 < secs:
        ts.append(c.score(_lib.ACQ_EI, 1.0, 0.01, fl)[2])
    p.terminate()
    lines = [l.strip() for l in p.stdout.read().splitlines() if l.strip()]
    clk = [float(l.split(",")[0]) for l in lines[2:]]
    pw = [float(l.split(",")[1]) for l in lines[2:]]
    cap = sum("Active" in l.split(",")[2] and "Not" not in l.split(",")[2] for l in lines[2:])
    print("%s: ms first %.2f  median %.2f  last %.2f | SM MHz median %.0f min %.0f | power median %.0f max %.0f W | sw_power_cap active in %d of %d samples"
          % (tag, ts[0], np.median(ts), ts[-1], np.median(clk), min(clk), np.median(pw), max(pw), cap, len(clk)), flush=True)
sample("int8", _lib.FLAG_MODE_CPP, 4.0)
sample("fp64", _lib.FLAG_MODE_CPP | _lib.FLAG_FP64, 4.0)
sample("int8", _lib.FLAG_MODE_CPP, 4.0)
