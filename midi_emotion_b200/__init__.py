"""B200-native hot path of serkansulun/midi-emotion: the emotion-conditioned MIDI-token transformer
(train.py / generate.py model calls) on hand-written sm_100a CUDA behind a C-ABI library.

Public surface (mirrors the reference's models package):
    build_model(args, load_config_dict=None) -> (nn.Module, dict)
    MusicTransformer.forward(x[B, L] int64, condition[B, 2] float) -> logits[B, Ls, V]
"""
from .build_model import build_model, CONDITIONINGS  # noqa: F401
from .transformer import MusicTransformer, positional_table, set_dropout  # noqa: F401
from .regression import MusicRegression  # noqa: F401
from .decode import KVCacheDecoder  # noqa: F401
from .sampling import Sampler, generate  # noqa: F401
from .loss import cross_entropy  # noqa: F401
from .optim import ClipAdam  # noqa: F401

__all__ = ["build_model", "MusicTransformer", "MusicRegression", "KVCacheDecoder", "Sampler", "generate", "cross_entropy", "ClipAdam", "set_dropout",
           "positional_table", "CONDITIONINGS"]
