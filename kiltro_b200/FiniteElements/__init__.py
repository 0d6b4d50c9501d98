from .Assembly import Assemble
from .ComputeSparsityPattern import ComputeSparsityPattern
