"""datashader_b200 - datashader's projection + aggregation hot path, rebuilt for B200 (sm_100a).

Drop-in for `Canvas.points / Canvas.line` and the reductions count, any, sum, mean, min, max, first,
last, where, by/count_cat, summary (same call signatures, same DataArray results as
holoviz/datashader 0.19.1); the work is done by hand-written CUDA kernels in libdsb200.so behind a
C ABI (include/dsb200.h).  There is no CPU fallback.
"""
from .core import Canvas, bypixel  # noqa: F401
from . import config  # noqa: F401
from .frame import DeviceFrame, HostFrame, RaggedColumn  # noqa: F401
from .reductions import (any, by, category_binning, category_codes, category_modulo, count, count_cat,  # noqa: F401,A004
                         first, last, max, mean, min, sum, summary, where)
from . import palette  # noqa: F401
from . import transfer_functions as tf  # noqa: F401

__version__ = "0.1.0"
