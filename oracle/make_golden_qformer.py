"""TEST INFRASTRUCTURE ONLY (this container; needs /root/reference) — golden vectors of the Q-Former stage from the REAL
reference code: ``BertLMHeadModel`` of src/sub_model/Qformer.py (loaded by oracle/ref_qformer.py) driven by the REAL
``MMS_LLaMA.compression_using_qformer`` (src/model.py:584-619), taken from the file by AST (the module as a whole
imports Whisper / LLaMA / peft) and bound to an object that only carries ``Qformer`` and ``query_tokens``.

  python -m oracle.make_golden_qformer    ->  tests/golden/qformer_tiny.npz (weights by seed through the oracle's keys)
"""
import ast
import os
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import qformer_oracle as qo
from . import ref_import, ref_qformer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPE = dict(hidden=128, heads=2, intermediate=256, layers=2, encoder_width=192, query_length=80)


CASES = {"short": dict(seed=3, B=3, T=40, C=192), "long": dict(seed=4, B=2, T=300, C=192), "one": dict(seed=5, B=1, T=26, C=192)}


def real_method():
    path = os.path.join(ref_import.REF, "src", "model.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MMS_LLaMA")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "compression_using_qformer")
    ns = {"torch": torch, "F": F, "nn": nn}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["compression_using_qformer"]


def seeded_oracle(seed=11):
    torch.manual_seed(seed)
    o = qo.OracleQFormer(SHAPE["hidden"], SHAPE["heads"], SHAPE["intermediate"], SHAPE["layers"], SHAPE["encoder_width"],
                         SHAPE["query_length"]).eval()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in o.parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        for n, p in o.named_parameters():
            if n.endswith("LayerNorm.weight"):
                p.add_(1.0)
        o.query_tokens.copy_(torch.randn(o.query_tokens.shape, generator=g) * 0.5)
    return o


def main():
    model, _ = ref_qformer.build(SHAPE["hidden"], SHAPE["heads"], SHAPE["intermediate"], SHAPE["layers"],
                                 SHAPE["encoder_width"], SHAPE["query_length"])
    oracle = seeded_oracle()
    sd = {k[len("Qformer."):]: v for k, v in oracle.state_dict().items() if k.startswith("Qformer.")}
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    host = types.SimpleNamespace(Qformer=model, query_tokens=oracle.query_tokens.detach())
    out = {}
    for name, kw in CASES.items():
        av, len_feat, resized, len_queries = qo.synthetic_case(**kw)
        with torch.no_grad():
            y_ref = real_method()(host, len_queries, resized, len_feat, av)
            y_or = oracle.compression_using_qformer(len_queries, resized, len_feat, av)
        print(name, "real vs oracle max abs diff", (y_ref - y_or).abs().max().item(), tuple(y_ref.shape), "keys",
              int(max(resized)), "|y| max", y_ref.abs().max().item())
        assert (y_ref - y_or).abs().max().item() < 1e-4
        out[name] = y_ref.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "qformer_tiny.npz"), **out)


if __name__ == "__main__":
    main()
