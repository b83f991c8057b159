"""Latin hypercube sampling with the reference's exact random stream
(ego/utils/latinhypercube.py:27-46): per dimension one `rand(N)` draw, then one `shuffle`."""
import numpy as np
from numpy.random import RandomState


def lhcSample(bounds, N, seed=None):
    """
    @param bounds:  sequence of [min, max] bounds for the space
    @param N:       number of samples
    @return: list of sample points (arrays); a dimension with min == max is passed through.
    """
    rs = RandomState(seed)
    cols = []
    for bmin, bmax in bounds:
        if bmin == bmax:
            col = np.array([bmin] * N)
        else:
            col = (bmax - bmin) * rs.rand(N) / N + np.arange(bmin, bmax, (bmax - bmin) / N)
        rs.shuffle(col)
        cols.append(col)
    return list(np.vstack(cols).T)
