"""Q-Former stage of MMS-LLaMA: oracle/qformer_oracle.py against goldens of the REAL classes
(src/sub_model/Qformer.py BertLMHeadModel driven by the real MMS_LLaMA.compression_using_qformer; made by
oracle/make_golden_qformer.py), live against the reference when present, and the device path (avh_qformer_forward:
tcgen05 GEMMs + the attention / LayerNorm kernels) against both."""
import os

import numpy as np
import pytest
import torch

from helpers import cosine, rel_err
from oracle import make_golden_qformer as mq
from oracle import qformer_oracle as qo
from oracle import ref_import


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "qformer_tiny.npz"))


@pytest.mark.parametrize("name", list(mq.CASES))
def test_qformer_oracle_reproduces_reference_golden(name, golden_dir):
    o = mq.seeded_oracle()
    av, len_feat, resized, len_queries = qo.synthetic_case(**mq.CASES[name])
    y = o.compression_using_qformer(len_queries, resized, len_feat, av)
    assert (y - torch.from_numpy(_golden(golden_dir)[name])).abs().max().item() < 1e-5


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present on this box")
def test_qformer_oracle_live_against_reference():
    import types
    from oracle import ref_qformer
    model, _ = ref_qformer.build(mq.SHAPE["hidden"], mq.SHAPE["heads"], mq.SHAPE["intermediate"], mq.SHAPE["layers"],
                                 mq.SHAPE["encoder_width"], mq.SHAPE["query_length"])
    o = mq.seeded_oracle(seed=23)
    missing = model.load_state_dict({k[len("Qformer."):]: v for k, v in o.state_dict().items() if k.startswith("Qformer.")},
                                    strict=False)
    assert not missing.unexpected_keys
    host = types.SimpleNamespace(Qformer=model, query_tokens=o.query_tokens.detach())
    av, len_feat, resized, len_queries = qo.synthetic_case(seed=9, B=3, T=77, C=192)
    with torch.no_grad():
        y_ref = mq.real_method()(host, len_queries, resized, len_feat, av)
    assert (o.compression_using_qformer(len_queries, resized, len_feat, av) - y_ref).abs().max().item() < 1e-5


def _device_model(dtype):
    from multimodalvc_b200.qformer import QFormerCompressor, QFormerConfig
    s = mq.SHAPE
    m = QFormerCompressor(QFormerConfig(hidden_size=s["hidden"], num_hidden_layers=s["layers"], num_attention_heads=s["heads"],
                                        intermediate_size=s["intermediate"], encoder_width=s["encoder_width"],
                                        query_length=s["query_length"]))
    missing = m.load_state_dict(mq.seeded_oracle().state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.to("cuda", dtype).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mq.CASES))
def test_device_qformer_matches_reference_golden(name, golden_dir):
    y_ref = torch.from_numpy(_golden(golden_dir)[name])
    av, len_feat, resized, len_queries = qo.synthetic_case(**mq.CASES[name])
    m = _device_model(torch.float32)                     # fp32 module: split-precision GEMMs, gate 2e-3 relative
    y = m.compression_using_qformer(len_queries, resized, len_feat, av.cuda())
    assert y.shape == y_ref.shape and y.dtype == torch.float32
    assert rel_err(y.cpu(), y_ref) < 2e-3, rel_err(y.cpu(), y_ref)
    y2 = m.compression_using_qformer(len_queries, resized, len_feat, av.cuda().clone())      # graph capture, fresh tensors
    y3 = m.compression_using_qformer(len_queries, resized, len_feat, av.cuda().clone())      # replay
    assert torch.equal(y, y2) and torch.equal(y, y3)
    mb = _device_model(torch.bfloat16)                   # bf16 mode: cosine gate
    yb = mb.compression_using_qformer(len_queries, resized, len_feat, av.cuda().bfloat16())
    assert yb.dtype == torch.bfloat16 and cosine(yb.float().cpu(), y_ref) > 0.999


@pytest.mark.gpu
def test_device_qformer_bert_masks_and_errors():
    m = _device_model(torch.float32)
    o = mq.seeded_oracle()
    g = torch.Generator().manual_seed(1)
    enc = torch.randn(2, 70, mq.SHAPE["encoder_width"], generator=g)
    mask = torch.ones(2, 70, dtype=torch.long)
    mask[1, 33:] = 0
    mask[0, 5:9] = 0                                     # a hole, not only a suffix
    lq = [37, 12]
    y_ref = o.bert(lq, enc, mask)
    y = m.bert(lq, enc.cuda(), mask.cuda())
    assert rel_err(y.cpu(), y_ref) < 2e-3
    y_nomask = m.bert(lq, enc.cuda(), None)
    assert rel_err(y_nomask.cpu(), o.bert(lq, enc, torch.ones(2, 70))) < 2e-3
    one = m.bert([1], enc[:1, :1].cuda(), None)                          # one query, one frame
    assert rel_err(one.cpu(), o.bert([1], enc[:1, :1], torch.ones(1, 1))) < 2e-3
    with pytest.raises(ValueError):
        m.bert([3], enc.cuda(), None)                                    # one length per sample
    with pytest.raises(ValueError):
        m.bert([mq.SHAPE["query_length"] + 1, 3], enc.cuda(), None)       # more queries than query_tokens holds
    with pytest.raises(ValueError):
        m.bert(lq, enc.cuda()[:, :, :64], None)                          # wrong encoder width
    with pytest.raises(RuntimeError):
        m.cpu().bert(lq, enc, None)                                      # no CPU path
