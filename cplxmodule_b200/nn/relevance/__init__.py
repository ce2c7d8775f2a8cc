from .base import BaseARD, penalties, named_penalties, named_relevance, compute_ard_masks
from .real import LinearVD, LinearARD, Conv1dVD, Conv2dVD, Conv1dARD, Conv2dARD
from .complex import (CplxLinearGaussian, CplxLinearVD, CplxLinearARD, CplxConv1dVD, CplxConv2dVD,
                      CplxConv1dARD, CplxConv2dARD)
from . import extensions
