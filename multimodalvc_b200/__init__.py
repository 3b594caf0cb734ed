"""multimodalvc_b200 — B200-native (sm_100a) implementation of one hot path of MMS-LLaMA / AV-HuBERT:
the batched encoder forward behind ``AVHubertModel.extract_finetune(source={'audio','video'}, padding_mask)``
(reference: avhubert/hubert.py:694-745) plus the audio featurisation that feeds it
(avhubert/hubert_dataset.py:286-296,317-346,351-353,430-456).

Thin PyTorch host code (device memory, streams, nn.Module surface) over ``libavh_b200.so``: hand-written
CUDA kernels (tcgen05/TMEM/TMA GEMM + implicit-GEMM convolutions, flash-style attention, fused log-fbank)
behind the C ABI declared in ``include/avh_b200.h``.
"""
from .hubert import AVHubertConfig, AVHubertModel  # noqa: F401
from .hubert_asr import HubertEncoder, HubertEncoderWrapper  # noqa: F401
from . import audio  # noqa: F401
from . import distributed  # noqa: F401
from . import fusion  # noqa: F401
from . import sr_predictor  # noqa: F401
from . import sharding  # noqa: F401
from . import video  # noqa: F401

__all__ = ["AVHubertConfig", "AVHubertModel", "HubertEncoder", "HubertEncoderWrapper", "audio", "distributed", "fusion", "sharding", "sr_predictor", "video"]
