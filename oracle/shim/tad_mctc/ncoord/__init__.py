from . import count, d3, d4, eeq  # noqa: F401
from .count import derf_count, dexp_count, dgfn2_count, erf_count, exp_count, gfn2_count  # noqa: F401
from .d3 import cn_d3, cn_d3_gradient, coordination_number  # noqa: F401
from .d4 import cn_d4  # noqa: F401
from .eeq import cn_eeq  # noqa: F401
