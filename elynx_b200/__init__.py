"""elynx_b200 -- B200-native alignment simulator behind the ELynx `elynx-markov` simulator API.

Host-side mirror of ELynx.Simulate.MarkovProcessAlongTree over the C ABI of libelynx_b200.so
(include/elynx_b200.h).  All compute runs in hand-written CUDA kernels for sm_100a; there is no CPU path."""
from ._lib import ElxError, LIB_PATH  # noqa: F401
from .simulate import (Simulator, simulate, simulate_and_flatten_mixture_model_par, simulate_and_flatten_par,  # noqa: F401
                       simulate_mixture_model, simulateAndFlattenMixtureModelPar, simulateAndFlattenPar, simulateMixtureModel)

__version__ = "0.1.0"
