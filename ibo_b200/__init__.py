"""
ibo_b200 -- B200-native drop-in for the acquisition hot path of misterwindupbird/IBO.

Package layout mirrors the reference's `ego` package for the modules on the hot path:

    ibo_b200.gaussianprocess   GaussianProcess, PrefGaussianProcess, kernels, RBFNMeanPrior, erf/CDF/PDF
    ibo_b200.acquisition       EI / PI / UCB, maximizeEI / maximizePI / maximizeUCB, cdirectGP, fastUCBGallery
    ibo_b200.utils.optimize    direct, cdirect
    ibo_b200.utils.latinhypercube  lhcSample

All numerics run in libibo_b200.so (hand-written sm_100a CUDA behind a C ABI, bound with ctypes over
NumPy buffers).  There is no PyTorch, no Triton and no CPU fallback in this package.
"""
__version__ = "0.1"
