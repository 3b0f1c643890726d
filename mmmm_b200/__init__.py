"""mmmm_b200 -- B200-native (sm_100a) visual-expert decoder layer for VividMed (function2-llx/MMMM).

Only the hot path named by BASELINE.json is here: ``CogVLMDecoderLayer`` and the kernels behind it.
Importing the package is cheap; the CUDA library (libvex.so) is built/loaded on first use and its
absence is a hard error (there is no CPU fallback).
"""
__all__ = ["CogVLMDecoderLayer", "VexConfig", "swap_decoder_layers", "build_plan", "VisualExpertDecoder",
           "fused_lm_head_loss", "LoraGradReducer"]


def __getattr__(name):
    if name in ("CogVLMDecoderLayer", "VexConfig", "swap_decoder_layers", "RMSNorm", "VisionExpertAttention",
                "VisionExpertMLP", "MLP", "RotaryEmbedding", "get_expert_mask", "VisualExpertDecoder",
                "masked_rms_norm"):
        from . import modeling_cogvlm as m
        return getattr(m, name)
    if name == "fused_lm_head_loss":
        from .lm_head import fused_lm_head_loss
        return fused_lm_head_loss
    if name in ("LoraGradReducer", "layer_forward_train"):
        from . import training as t
        return getattr(t, name)
    if name == "build_plan":
        from .plan import build_plan
        return build_plan
    raise AttributeError(name)
