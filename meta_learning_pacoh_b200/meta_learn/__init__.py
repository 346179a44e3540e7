"""Drop-in counterparts of the reference's ``meta_learn`` package for the PACOH meta-training hot path
(meta_learn/__init__.py:1-6): same class names, constructor signatures and methods, running on the B200 engine."""
from .GPR_meta_mll import GPRegressionMetaLearned
from .GPR_meta_vi import GPRegressionMetaLearnedVI
from .GPR_meta_svgd import GPRegressionMetaLearnedSVGD
from .GPR_mll import GPRegressionLearned

__all__ = ["GPRegressionMetaLearned", "GPRegressionMetaLearnedVI", "GPRegressionMetaLearnedSVGD", "GPRegressionLearned"]
