from ..multivector import dense_to_mv_local, mv_to_dense, mv_to_dense_local
from .affinity import bind_to_gpu_numa_node
