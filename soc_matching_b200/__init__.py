"""soc_matching_b200 -- B200-native (sm_100a) implementation of the SOC-matching hot path:
Euler-Maruyama rollout -> SOCM matching target -> importance-weighted loss + backward.

Drop-in surface (same names / signatures as facebookresearch/SOC-matching):
    stochastic_trajectories            (SOC_matching/utils.py:17)
    SOC_Solver, NeuralSDE              (SOC_matching/method.py:146, :15)
    FullyConnectedUNet, SigmoidMLP, TwoBoundarySigmoidMLP   (SOC_matching/models.py)
    LinearControl, ConstantControlLinear, LowDimControl     (SOC_matching/models.py:10-150, tabulated controls)
    OU_Quadratic, OU_Linear, DoubleWell, MolecularDynamics  (SOC_matching/experiment_settings/)
All arithmetic of the path runs in hand-written CUDA kernels reached through the C ABI in
include/socm_b200.h (libsocm_b200.so); there is no CPU fallback.
"""
from .controls import ConstantControlLinear, LinearControl, LowDimControl  # noqa: F401
from .networks import FullyConnectedUNet, SigmoidMLP, TwoBoundarySigmoidMLP, WarmStartTable  # noqa: F401
from .sde import (DoubleWell, MolecularDynamics, NeuralSDE, OU_Linear, OU_Quadratic,  # noqa: F401
                  describe_setting, make_benchmark_sde)
from .simulate import control_objective, normalization_constant, rollout, stochastic_trajectories  # noqa: F401
from .solver import SOC_Solver  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .training import Trainer, TrainingStatistics  # noqa: F401

__all__ = [
    "FusedAdam", "Trainer", "TrainingStatistics", "stochastic_trajectories", "control_objective", "normalization_constant", "rollout", "SOC_Solver", "NeuralSDE",
    "FullyConnectedUNet", "SigmoidMLP", "TwoBoundarySigmoidMLP", "WarmStartTable",
    "OU_Quadratic", "OU_Linear", "DoubleWell", "MolecularDynamics", "describe_setting", "make_benchmark_sde",
    "LinearControl", "ConstantControlLinear", "LowDimControl",
]
