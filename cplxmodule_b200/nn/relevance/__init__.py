from .base import BaseARD, penalties, named_penalties, named_relevance, compute_ard_masks
from .real import LinearVD, LinearARD
from .complex import (CplxLinearGaussian, CplxLinearVD, CplxLinearARD, CplxConv1dVD, CplxConv2dVD,
                      CplxConv1dARD, CplxConv2dARD)
