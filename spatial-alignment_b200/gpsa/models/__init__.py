from .gpsa import GPSA
from .vgpsa import VariationalGPSA
