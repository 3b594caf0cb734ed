"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference Q-Former (src/sub_model/Qformer.py) in this container.

The file targets transformers 4.15; under the installed 5.x its imports (``apply_chunking_to_forward`` & co. from
``transformers.modeling_utils``) and ``PreTrainedModel.init_weights`` plumbing no longer exist.  As with
oracle/ref_import.py, only the GLUE is stubbed: while the reference file is imported, ``transformers.modeling_utils`` is
replaced by a module whose ``PreTrainedModel`` is a bare ``nn.Module`` with the five members the reference's
``BertModel`` touches (config, dtype, init_weights -> the file's own ``_init_weights``, get_head_mask,
invert_attention_mask with the 4.15 arithmetic).  Every class that computes (BertEmbeddings, BertSelfAttention,
BertLayer, BertEncoder, BertModel.forward, get_extended_attention_mask) is the reference's own code.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

from oracle import ref_import

PATH = os.path.join(ref_import.REF, "src", "sub_model", "Qformer.py")


def available():
    return os.path.isfile(PATH)


class _PreTrainedModel(nn.Module):
    config_class = None
    base_model_prefix = ""

    def __init__(self, config, *a, **k):
        super().__init__()
        self.config = config

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def init_weights(self):
        self.apply(self._init_weights)

    def get_head_mask(self, head_mask, n, is_attention_chunked=False):
        assert head_mask is None
        return [None] * n

    def invert_attention_mask(self, m):          # transformers 4.15 modeling_utils.py: (1 - mask) * -1e4 (fp16) / -1e9
        ext = m[:, None, :, :] if m.dim() == 3 else m[:, None, None, :]
        ext = ext.to(dtype=self.dtype)
        return (1.0 - ext) * (-1e4 if self.dtype == torch.float16 else -1e9)


_mod = None


def load():
    """The reference module (cached)."""
    global _mod
    if _mod is not None:
        return _mod
    import transformers.pytorch_utils as pu
    real = sys.modules.get("transformers.modeling_utils")
    import transformers.modeling_utils  # noqa: F401  (make sure the real one is imported before it is shadowed)
    real = sys.modules["transformers.modeling_utils"]
    stub = types.ModuleType("transformers.modeling_utils")
    stub.PreTrainedModel = _PreTrainedModel
    stub.apply_chunking_to_forward = pu.apply_chunking_to_forward
    stub.find_pruneable_heads_and_indices = getattr(pu, "find_pruneable_heads_and_indices", None)
    stub.prune_linear_layer = pu.prune_linear_layer
    sys.modules["transformers.modeling_utils"] = stub
    try:
        spec = importlib.util.spec_from_file_location("avh_ref_qformer", PATH)
        m = importlib.util.module_from_spec(spec)
        sys.modules["avh_ref_qformer"] = m
        spec.loader.exec_module(m)
    finally:
        sys.modules["transformers.modeling_utils"] = real
    _mod = m
    return m


def build(hidden, heads, intermediate, layers, encoder_width, query_length, seed=0):
    """BertLMHeadModel as src/model.py:121-128 configures it (bert-large-uncased fields that matter given explicitly)."""
    m = load()
    cfg = m.BertConfig(hidden_size=hidden, num_attention_heads=heads, intermediate_size=intermediate,
                       num_hidden_layers=layers, vocab_size=64, max_position_embeddings=32, layer_norm_eps=1e-12,
                       hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    cfg.encoder_width = encoder_width
    cfg.add_cross_attention = True
    cfg.cross_attention_freq = 1
    cfg.query_length = query_length
    torch.manual_seed(seed)
    model = m.BertLMHeadModel(config=cfg).eval()
    return model, cfg
