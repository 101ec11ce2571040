"""mdt_policy_b200 -- B200-native (sm_100a) implementation of the MDT denoising hot path.

Drop-in replacements for the reference's Hydra targets:
    mdt.models.edm_diffusion.score_wrappers.GCDenoiser      -> mdt_policy_b200.score_wrappers.GCDenoiser
    mdt.models.networks.mdtv_transformer.MDTVTransformer    -> mdt_policy_b200.networks.MDTVTransformer
    mdt.models.networks.mdt_transformer.MDTTransformer      -> mdt_policy_b200.networks.MDTTransformer
    mdt.models.edm_diffusion.gc_sampling.*                  -> mdt_policy_b200.gc_sampling.*
    mdt.models.networks.transformers.perceiver_resampler.PerceiverResampler -> mdt_policy_b200.perceiver.PerceiverResampler
plus: DenoiseAgent (denoise_actions / step / reset), rollout.BatchedRollout (vectorised step()), optim.FusedAdamWEMA.
All arithmetic runs in libmdtb200.so (hand-written CUDA, C ABI in include/mdtb200.h); no CPU fallback.
"""
from .score_wrappers import GCDenoiser
from .networks import MDTVTransformer, MDTTransformer
from . import gc_sampling
from .agent import DenoiseAgent
from .perceiver import PerceiverResampler
from .rollout import BatchedRollout

__all__ = ["GCDenoiser", "MDTVTransformer", "MDTTransformer", "gc_sampling", "DenoiseAgent", "PerceiverResampler", "BatchedRollout"]
