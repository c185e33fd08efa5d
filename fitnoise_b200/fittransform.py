"""``fitnoise.transform`` (optimal moderated log transformation, reference fittransform.py) is a
different, dense objective that is not part of the per-gene hot path this package rebuilds
(SURVEY.md section 8f, rank 4).  The name is kept importable and fails loudly."""


def transform(*args, **kwargs):
    raise NotImplementedError(
        "fitnoise_b200 does not implement fitnoise.transform (fittransform.py:215-219); "
        "only the noise-model fit / coefficient / test / fdr path is built")
