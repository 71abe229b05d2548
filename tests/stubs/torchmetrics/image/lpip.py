from . import _Unavailable


class LearnedPerceptualImagePatchSimilarity(_Unavailable):
    pass
