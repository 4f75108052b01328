import scipy.stats as _st

from ...numpy import _wrap

logpdf = _wrap(_st.norm.logpdf)
pdf = _wrap(_st.norm.pdf)
cdf = _wrap(_st.norm.cdf)
