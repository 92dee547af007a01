"""bayesgm_b200 -- the posterior-sampling hot path of liuq-lab/bayesgm on B200.

Hand-written sm_100a CUDA behind the reference's Python method surface
(`CausalBGM.predict / metropolis_hastings_sampler / get_log_posterior /
infer_from_latent_posterior`, `BGM.predict / tfp_mcmc_sampler`).  See DESIGN.md.
"""
__version__ = "0.1.0"

from .causalbgm import CausalBGM  # noqa: F401
from .bgm import BGM  # noqa: F401
from .variants import IdentifiableCausalBGM, FullMCMCCausalBGM  # noqa: F401
from . import datasets  # noqa: F401

__all__ = ["CausalBGM", "BGM", "IdentifiableCausalBGM", "FullMCMCCausalBGM", "datasets"]
