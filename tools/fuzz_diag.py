"""Per-quantity errors of chosen cases of tests/test_gpu_fuzz.py against both forms of the oracle's sigma^2: python tools/fuzz_diag.py 23 26"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from oracle import ibo_oracle as orc
from ibo_b200 import _lib
import importlib.util
spec=importlib.util.spec_from_file_location('f', os.path.join(ROOT, 'tests', 'test_gpu_fuzz.py')); f=importlib.util.module_from_spec(spec); spec.loader.exec_module(f)
for case in [int(a) for a in sys.argv[1:]] or [23, 26, 122, 152]:
    rs, kind, d, N, M, noise, hyper, prior, mode = f._draw(case)
    X = rs.rand(N, d); Y = np.sin(2.5 * X).sum(axis=1) + 0.1 * rs.randn(N)
    Xs = rs.rand(M, d); k = min(N, M // 3)
    if k: Xs[:k] = X[:k] + 1e-3 * rs.randn(k, d)
    op = None
    if prior: op = orc.PriorSpec(rs.rand(4, d), 0.3 * rs.randn(4), 3.0, np.zeros(d), np.ones(d))
    o = orc.GPOracle(orc.KernelSpec(kind, hyper, d), X, Y, noise, prior=op)
    m = _lib.Model(kind, hyper, X, Y, noise, prior=op)
    if mode == "py": mu_o, s2_o = o.posterior_batch(Xs); fl = _lib.FLAG_MODE_PY
    else: mu_o, sig = o.posterior_cpp(Xs); s2_o = sig ** 2; fl = _lib.FLAG_MODE_CPP
    ymax=float(Y.max())
    print("case", case, "kind", kind, "d", d, "N", N, "M", M, "noise", noise, "prior", prior, mode, "cond(R)=%.2e" % np.linalg.cond(o.R))
    for acq, parm in ((orc.ACQ_EI, 0.01), (orc.ACQ_PI, 0.01), (orc.ACQ_UCB, 1.7)):
        sc, mu, s2, best, bidx = m.score(Xs, acq, ymax, parm, flags=fl, want_posterior=True)
        want = orc.score(acq, mode, mu_o, s2_o, ymax, parm)
        floor = 1e-5 if acq != orc.ACQ_UCB else 1e-3
        e_mu=np.abs(mu - mu_o) / np.maximum(np.abs(mu_o), 1e-3); e_s=np.abs(np.sqrt(s2) - np.sqrt(s2_o)) / np.sqrt(s2_o); e_sc=np.abs(sc - want) / np.maximum(np.abs(want), floor)
        i=int(np.argmax(e_sc))
        print("  acq", acq, "mu %.2e sigma %.2e score %.2e at %d: want %.6e got %.6e sigma %.3e mu %.4f Z=%.2f; argmax ok %s" % (e_mu.max(), e_s.max(), e_sc.max(), i, want[i], sc[i], np.sqrt(s2_o[i]), mu_o[i], (mu_o[i]-ymax-parm)/np.sqrt(s2_o[i]), want[bidx] >= want.max() - 1e-10*max(abs(want.max()),floor)))
    mu_p, s2_p = o.posterior_batch(Xs, floor=1e-8)
    sc, mu, s2, best, bidx = m.score(Xs, 0, ymax, 0.01, flags=fl, want_posterior=True)
    print("  sigma^2: GPU vs triangular-solve oracle %.2e, GPU vs explicit-inverse oracle %.2e, the two oracles %.2e" % (
        np.max(np.abs(s2 - s2_p) / s2_p), np.max(np.abs(s2 - s2_o) / s2_o), np.max(np.abs(s2_p - s2_o) / s2_o)))
    m.close()
