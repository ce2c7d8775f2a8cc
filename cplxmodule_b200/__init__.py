"""B200-native (sm_100a) implementation of cplxmodule's complex linear/conv +
variational-dropout forward + KL ``penalties()`` hot path, behind the reference's
Python API (``Cplx``, ``nn.CplxLinear``, ``nn.CplxConv1d/2d``,
``nn.relevance.{LinearVD, LinearARD, CplxLinearVD, CplxLinearARD, penalties}``)."""
__version__ = "0.1.0"

from .cplx import Cplx, from_real, to_real
from .ops import (set_noise_mode, set_math_mode, set_operand_prepass, set_kl_fusion, set_kl_shard, set_sm_reserve, set_conv_vd_mode,
                  get_noise_mode,
                  get_math_mode)
from . import nn
