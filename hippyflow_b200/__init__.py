"""hippyflow_b200 -- B200-native reduced-basis hot path behind hIPPYflow's projector / operator /
collective API.  Importing the package does not load the CUDA library; the first
numerical call does, and raises if it is missing (no CPU fallback)."""
from .collectives import (CollectiveOperator, MatrixMultCollectiveOperator, MultipleSamePartitioningPDEsCollective,
                          MultipleSerialPDEsCollective, NcclCollective, NullCollective, TorchCollective, splitCommunicators)
from .multivector import DeviceMultiVector, DeviceVector, dense_to_mv_local, mv_to_dense, mv_to_dense_local
from .parameterList import ParameterList
from .randomized import doublePass, doublePassG
from .linalg import CsrCGSolver, CsrMatrix, SampleCovariance, b_orthonormalize
from .modeling import *
from .dataIO import compress_dataset

__version__ = "0.1.0"
