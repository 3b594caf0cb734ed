"""TEST INFRASTRUCTURE ONLY (this container; needs /root/reference) — golden vectors for the Speech_Rate_Predictor
mirror from the REAL reference class: the class statement is taken from src/sub_model/modules.py by AST (the module
as a whole imports peft / bitsandbytes / Whisper, absent here) and executed with the reference's own
fairseq TransformerEncoder (oracle/ref_import.py).  Nothing of the reference is written to the repository but the
class's numerical output.

  python -m oracle.make_golden_sr        ->  tests/golden/sr_predictor.npz
"""
import ast
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from . import ref_import, sr_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def real_class():
    _, w2v = ref_import.install()
    path = os.path.join(ref_import.REF, "src", "sub_model", "modules.py")
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Speech_Rate_Predictor")
    ns = {"nn": nn, "torch": torch, "SimpleNamespace": SimpleNamespace, "TransformerEncoder": w2v.TransformerEncoder}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["Speech_Rate_Predictor"]


def main():
    oracle = sr_oracle.build(2, seed=77)
    ref = real_class()(2).eval()
    missing = ref.load_state_dict(oracle.state_dict(), strict=True)
    x = sr_oracle.synthetic_features(3, 40, seed=5)
    with torch.no_grad():
        y_ref = ref(x)
        y_or = oracle(x)
    print("real vs oracle max abs diff", (y_ref - y_or).abs().max().item(), "outputs", y_ref.flatten().tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sr_predictor.npz"), y=y_ref.numpy(),
                        meta=np.array(["2", "77", "3", "40", "5"]))


if __name__ == "__main__":
    main()
