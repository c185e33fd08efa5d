"""fitnoise_b200 -- B200-native implementation of the Fitnoise hot path.

Drop-in for ``import fitnoise`` (reference fitnoise/__init__.py:1-7) for the noise-model fit,
linear-model coefficients, moderated t/F tests and BH adjustment:

    import fitnoise_b200 as fitnoise
    fitted = fitnoise.Model_t().fit(y, design)
    tested = fitted.test(coef=[1])

All per-gene computation runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/fitnoise_b200.h``; there is no CPU fallback.
"""
VERSION = "2.2"

from .env import Withable, as_jsonic  # noqa: F401,E402
from .p_adjust import fdr  # noqa: F401,E402
from .fittransform import Transform, Transform_polynomial, transform  # noqa: F401,E402
from .distributions import Mvnormal, Mvt  # noqa: F401,E402
from .models import *  # noqa: F401,F403,E402
from .models import Dataset, Model, as_dataset  # noqa: F401,E402
