"""B200-native (sm_100a) particle filter / fixed-lag smoother for qnmh-sysid2018's state-estimation
hot path.  `ParticleMethodsB200` is the drop-in estimator; `Engine` the handle-level API over
libpfb200.so (include/pfb200.h); `models` the host-side mirror of the reference's model surface."""
from .models import (LinearGaussianModel, StochasticVolatilityModel,  # noqa: F401
                     StochasticVolatilityModelLeverage, model_kind_of)
from .particle_methods import ParticleMethodsB200  # noqa: F401

__all__ = ["ParticleMethodsB200", "Engine", "LinearGaussianModel", "StochasticVolatilityModel",
           "StochasticVolatilityModelLeverage", "model_kind_of"]


def __getattr__(name):
    if name == "Engine":  # loads the CUDA library on first use (fails loudly if it is missing)
        from .engine import Engine
        return Engine
    raise AttributeError(name)
