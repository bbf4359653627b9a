"""scfgp_b200 -- B200-native marginal-likelihood hot path of SCFGP behind the reference's API.

    from scfgp_b200 import SCFGP, Scaler, Optimizer        # reference: SCFGP/__init__.py:7-9
"""
from .model import SCFGP, CompiledFuncs  # noqa: F401
from .scaler import Scaler  # noqa: F401
from .gpu_scaler import GpuScaler  # noqa: F401
from .optimizer import Optimizer  # noqa: F401
from .engine import Engine, ScfgpError  # noqa: F401
from .sharding import shard_bounds, weighted_bounds, run_eval_phases, predict_sharded  # noqa: F401

__all__ = ["SCFGP", "Scaler", "GpuScaler", "Optimizer", "Engine", "ScfgpError", "CompiledFuncs", "shard_bounds", "weighted_bounds",
           "run_eval_phases", "predict_sharded"]
