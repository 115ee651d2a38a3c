from .PODProjector import PODParameterList, PODProjector, PODProjectorFromData, StoredSnapshots, weighted_l2_norm_vector
from .KLEProjector import KLEParameterList, KLEProjector, MassPreconditionedCovarianceOperator, SampleCovariancePrior
from .activeSubspaceProjector import (ActiveSubspaceParameterList, ActiveSubspaceProjector, LowRankHessian, SparsePrior,
                                      StoredJacobians)
from .operators import (JJT, JTJ, LowRankRectangularOperator, MeanJJTfromDataOperator, MeanJTJfromDataOperator, SampleCovarianceOperator,
                        SandwichedCovarianceOperator, SummedListOperator, npToDolfinOperator)
from .projection import jacobian_action, jacobian_transpose_action, project_data, reduced_jacobians, stacked_jacobians
from .errors import PriorPreconditionedProjector, jacobian_truncated_svd, projection_errors
from .randomizedSVD import accuracyEnhancedSVD_batched
