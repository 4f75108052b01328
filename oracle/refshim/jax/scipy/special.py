import scipy.special as _sp

from ..numpy import _wrap

gammaln = _wrap(_sp.gammaln)
erf = _wrap(_sp.erf)
erfinv = _wrap(_sp.erfinv)
logsumexp = _wrap(_sp.logsumexp)
