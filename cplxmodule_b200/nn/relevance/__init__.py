from .base import BaseARD, penalties, named_penalties, named_relevance, compute_ard_masks
from .real import (LinearGaussian, LinearVD, LinearARD, BilinearGaussian, BilinearVD, BilinearARD,
                   Conv1dGaussian, Conv2dGaussian, Conv1dVD, Conv2dVD, Conv1dARD, Conv2dARD)
from .complex import (CplxLinearGaussian, CplxLinearVD, CplxLinearARD, CplxBilinearGaussian,
                      CplxBilinearVD, CplxBilinearARD, CplxConv1dGaussian, CplxConv2dGaussian,
                      CplxConv1dVD, CplxConv2dVD, CplxConv1dARD, CplxConv2dARD)
from . import extensions
