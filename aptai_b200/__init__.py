"""aptai_b200 — B200-native implementation of the APTAI hot path.

Public API mirrors the reference's model files:
    models/aptai.py        -> aptai_b200.APTAI
    models/w2v2_pr.py      -> aptai_b200.Wav2Vec2_PR
    models/force_aptai.py  -> aptai_b200.Force_APTAI
    models/modules.py      -> aptai_b200.LowPassFilterLayer, ForwardSumLoss, CrossAttention, RNN, PositionalEncoding
plus `forced_align` (CTC Viterbi) and the `Wav2Vec2Backbone` that stands in for transformers.Wav2Vec2Model.
Training: the modules' `forward` in train mode returns a loss whose `.backward()` runs the hand-written backward
kernels; `aptai_b200.train` has the flat gradient buffer, `FusedAdam` and the data-parallel gradient reducer.
`aptai_b200.postproc` mirrors the reference's caller-side helpers (collate, resample, segments, metrics) on the device.
"""
from .config import W2V2Config, frame_lengths
from .backbone import Wav2Vec2Backbone
from .modules import LowPassFilterLayer, ForwardSumLoss, CrossAttention, RNN, PositionalEncoding
from .aptai import APTAI
from .w2v2_pr import Wav2Vec2_PR
from .force_aptai import Force_APTAI
from .ops import ctc_viterbi as forced_align
from .train import FusedAdam, GradBuffer, GradReducer

__all__ = ["W2V2Config", "frame_lengths", "Wav2Vec2Backbone", "LowPassFilterLayer", "ForwardSumLoss",
           "CrossAttention", "RNN", "PositionalEncoding", "APTAI", "Wav2Vec2_PR", "Force_APTAI", "forced_align",
           "FusedAdam", "GradBuffer", "GradReducer"]
