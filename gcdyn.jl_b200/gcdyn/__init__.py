"""gcdyn — host-side mirror of the Julia package `gcdyn` (reference: src/gcdyn.jl:1-40),
with the tree log-likelihoods executed by the B200 CUDA library `libgcdyn_b200.so`
through its C ABI (`include/gcdyn_b200.h`).  Names follow the reference's export list
(gcdyn.jl:15-33); `attach!`/`detach!`/`delete!`/`map_types!` drop the `!`.
"""
from .types import (EVENTS, EVENT_CODE, TreeNode, AbstractBranchingProcess,
                    SigmoidalBranchingProcess, ConstantBranchingProcess, DiscreteBranchingProcess,
                    ArgumentError, DimensionMismatch, BoundsError, DomainError)
from .treenode import (attach, detach, delete, map_types, map_types_inplace,
                       PostOrderTraversal, PreOrderTraversal, LeafTraversal)
from .multitype_branching_process import (expit, sigmoid, type_space_index, λ, μ, γ,
                                          lambda_, mu, gamma, rand_tree, mutate, sample_child)
from .flatten import FlatForest, flatten, map_types_flat
from .likelihood import (loglikelihood, naive_loglikelihood, stadler_appx_loglikelihood,
                         stadler_appx_unconditioned_loglikelihood, Forest, Context,
                         pack_params, DeviceParams, evaluate, evaluate_resident,
                         GcdynError, library_path, loglikelihood_and_gradient)
from .mcmc import (random_walk_metropolis, device_metropolis, loglikelihood_gradient, loglikelihood_gradient_exact,
                   MetropolisResult)
from .native_sim import simulate_flat

__all__ = [
    # reference exports (gcdyn.jl:15-33)
    "ConstantBranchingProcess", "DiscreteBranchingProcess", "SigmoidalBranchingProcess",
    "TreeNode", "LeafTraversal", "PostOrderTraversal", "PreOrderTraversal",
    "λ", "μ", "γ", "loglikelihood", "rand_tree", "map_types", "map_types_inplace",
    "attach", "detach", "delete",
    # unexported in the reference, reachable as `gcdyn.<name>` there too
    "naive_loglikelihood", "stadler_appx_loglikelihood", "stadler_appx_unconditioned_loglikelihood",
    "expit", "sigmoid", "type_space_index", "mutate", "sample_child", "EVENTS",
    # additions of this build
    "Forest", "Context", "FlatForest", "flatten", "map_types_flat", "pack_params", "DeviceParams", "evaluate",
    "evaluate_resident", "lambda_", "mu", "gamma",
    "ArgumentError", "DimensionMismatch", "BoundsError", "DomainError", "GcdynError", "library_path",
]
