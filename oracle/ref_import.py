"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference sources (this container only).

`/root/reference` ships a vendored fairseq that cannot be imported on Python 3.12
(`fairseq/fairseq/dataclass/configs.py:971` mutable-default dataclass; omegaconf/hydra absent).
This module installs *stubs for the glue only* (omegaconf, fairseq.dataclass, fairseq.models registry,
fairseq.data, avhubert.hubert_pretraining/decoder/utils, python_speech_features) and then lets Python
import the real, unmodified files that carry the arithmetic of the hot path:

  avhubert/hubert.py (AVHubertModel.extract_finetune, :694-745), avhubert/resnet.py,
  avhubert/hubert_dataset.py (stacker/add_noise/collater_audio, :259-456),
  fairseq/fairseq/models/wav2vec/wav2vec2.py (TransformerEncoder, :816-1014),
  fairseq/fairseq/modules/{multihead_attention,same_pad,gelu,layer_norm,grad_multiply,...}.py,
  fairseq/fairseq/utils.py (index_put, get_activation_fn).

It is used by oracle/make_golden.py to generate tests/golden/*.npz and by the "not gpu" tests
(when /root/reference exists) to pin oracle/avhubert_oracle.py against the real reference.
Nothing here is copied from the reference; nothing here runs on the GPU box.
"""
import dataclasses
import enum
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("AVH_REFERENCE_ROOT", "/root/reference")
FS = os.path.join(REF, "fairseq", "fairseq")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "avhubert", "hubert.py"))


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


_installed = None


def install():
    """Returns (avhubert.hubert module, wav2vec2 module). Idempotent."""
    global _installed
    if _installed is not None:
        return _installed
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    om = types.ModuleType("omegaconf")
    om.II = lambda s: None
    om.MISSING = "???"
    om.DictConfig = dict
    om.OmegaConf = object
    om.open_dict = None
    sys.modules.setdefault("omegaconf", om)

    fs = _pkg("fairseq", FS)
    dc = _pkg("fairseq.dataclass")

    def ChoiceEnum(choices):
        return enum.Enum("Choices", {c: c for c in choices})

    @dataclasses.dataclass
    class FairseqDataclass:
        _name: str = None

    dc.ChoiceEnum = ChoiceEnum
    dc.FairseqDataclass = FairseqDataclass
    fs.dataclass = dc

    md = _pkg("fairseq.models", os.path.join(FS, "models"))

    class BaseFairseqModel(nn.Module):
        def upgrade_state_dict_named(self, sd, name):
            return sd

    def register_model(name, dataclass=None):
        def deco(cls):
            return cls
        return deco

    md.BaseFairseqModel = BaseFairseqModel
    md.register_model = register_model
    md.FairseqEncoder = nn.Module
    fs.models = md
    _pkg("fairseq.models.wav2vec", os.path.join(FS, "models", "wav2vec"))

    da = _pkg("fairseq.data")
    du = types.ModuleType("fairseq.data.data_utils")
    du.compute_mask_indices = None
    sys.modules["fairseq.data.data_utils"] = du
    da.data_utils = du
    di = types.ModuleType("fairseq.data.dictionary")
    di.Dictionary = object
    sys.modules["fairseq.data.dictionary"] = di
    fd = types.ModuleType("fairseq.data.fairseq_dataset")
    fd.FairseqDataset = object
    sys.modules["fairseq.data.fairseq_dataset"] = fd

    mods = _pkg("fairseq.modules", os.path.join(FS, "modules"))
    fs.modules = mods
    for leaf in ["fairseq_dropout", "quant_noise", "gelu", "grad_multiply", "layer_norm", "same_pad",
                 "transpose_last", "fp32_group_norm", "gumbel_vector_quantizer"]:
        importlib.import_module("fairseq.modules." + leaf)
    importlib.import_module("fairseq.incremental_decoding_utils")
    utils = importlib.import_module("fairseq.utils")   # pulls in the real multihead_attention.py
    fs.utils = utils
    mha = importlib.import_module("fairseq.modules.multihead_attention")
    sm = sys.modules
    mods.MultiheadAttention = mha.MultiheadAttention
    mods.Fp32GroupNorm = sm["fairseq.modules.fp32_group_norm"].Fp32GroupNorm
    mods.Fp32LayerNorm = sm["fairseq.modules.layer_norm"].Fp32LayerNorm
    mods.LayerNorm = sm["fairseq.modules.layer_norm"].LayerNorm
    mods.GradMultiply = sm["fairseq.modules.grad_multiply"].GradMultiply
    mods.GumbelVectorQuantizer = sm["fairseq.modules.gumbel_vector_quantizer"].GumbelVectorQuantizer
    mods.gelu = sm["fairseq.modules.gelu"].gelu
    mods.gelu_accurate = sm["fairseq.modules.gelu"].gelu_accurate
    mods.SamePad = sm["fairseq.modules.same_pad"].SamePad
    mods.TransposeLast = sm["fairseq.modules.transpose_last"].TransposeLast
    # init_bert_params: run the reference's own function body (the rest of that file needs heavy deps)
    path = os.path.join(FS, "modules", "transformer_sentence_encoder.py")
    src = open(path).read()
    start = src.index("def init_bert_params")
    end = src.index("class TransformerSentenceEncoder")
    ns = {"nn": nn, "torch": torch, "MultiheadAttention": mha.MultiheadAttention}
    exec(compile("\n" * src[:start].count("\n") + src[start:end], path, "exec"), ns)
    tse = types.ModuleType("fairseq.modules.transformer_sentence_encoder")
    tse.init_bert_params = ns["init_bert_params"]
    sys.modules["fairseq.modules.transformer_sentence_encoder"] = tse
    w2v = importlib.import_module("fairseq.models.wav2vec.wav2vec2")

    _pkg("avhubert", os.path.join(REF, "avhubert"))
    hp = types.ModuleType("avhubert.hubert_pretraining")
    hp.AVHubertPretrainingConfig = object
    hp.AVHubertPretrainingTask = object
    sys.modules["avhubert.hubert_pretraining"] = hp
    dec = types.ModuleType("avhubert.decoder")
    dec.TransformerDecoder = object
    sys.modules["avhubert.decoder"] = dec
    ut = types.ModuleType("avhubert.utils")
    ut.compute_mask_indices = None
    sys.modules["avhubert.utils"] = ut
    argv = sys.argv
    sys.argv = ["x", "y"]  # the reference picks relative imports unless len(sys.argv)==1 (hubert.py:28)
    try:
        hub = importlib.import_module("avhubert.hubert")
    finally:
        sys.argv = argv
    _installed = (hub, w2v)
    return _installed


def install_dataset(logfbank_impl):
    """Import the real avhubert/hubert_dataset.py with `python_speech_features.logfbank` (third-party,
    absent here) supplied by `logfbank_impl`.  Returns the module (AVHubertDataset lives in it)."""
    install()
    psf = types.ModuleType("python_speech_features")
    psf.logfbank = logfbank_impl
    sys.modules["python_speech_features"] = psf
    argv = sys.argv
    sys.argv = ["x", "y"]
    try:
        return importlib.import_module("avhubert.hubert_dataset")
    finally:
        sys.argv = argv


def build_reference_model(size="base", audio_feat_dim=104, seed=1234, **overrides):
    """Instantiate the real AVHubertModel (fine-tuning form: dictionaries=[None]) with reference init."""
    hub, _ = install()
    cfg = hub.AVHubertConfig()
    shape = dict(base=(12, 768, 3072, 12), large=(24, 1024, 4096, 16), tiny=(2, 128, 256, 2))[size]
    cfg.encoder_layers, cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads = shape
    cfg.audio_feat_dim = audio_feat_dim
    cfg.modality_fuse = "concat"
    cfg.layer_norm_first = True        # every shipped config: avhubert/conf/pretrain/*.yaml
    cfg.label_rate = 25
    for k in ["dropout", "attention_dropout", "activation_dropout", "encoder_layerdrop", "dropout_input",
              "dropout_features"]:
        setattr(cfg, k, 0.0)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    task_cfg = types.SimpleNamespace(sample_rate=25)
    import warnings
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = hub.AVHubertModel(cfg, task_cfg, [None])
    return model.eval(), cfg
