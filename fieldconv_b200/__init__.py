"""fieldconv_b200 — B200-native (sm_100a) FieldConv hot path behind the reference's module API.

    from fieldconv_b200 import FieldConv, FCResNetBlock, FCPrecomp, build_plan
"""
from .plan import DensePlan, Plan, build_dense_plan, build_plan  # noqa: F401
from .partition import MeshPartition, allreduce_gradients, partition_mesh  # noqa: F401
from .transforms import FCPrecomp, SupportGraph, farthest_point_sample, radius_graph  # noqa: F401
from .echo import ECHO, ECHOBlock  # noqa: F401
from .lift import LiftBlock, TransField  # noqa: F401
from .nn import FCResNetBlock, FieldConv, TangentLin, TangentNonLin, TangentPerceptron, fold_weights, prefold  # noqa: F401

__version__ = "0.1.0"
