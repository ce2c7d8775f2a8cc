from .modules import *
from .modules.base import CplxParameter
from . import init
from . import relevance
from . import masked
from . import utils
