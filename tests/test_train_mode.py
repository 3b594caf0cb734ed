"""Training-mode forward (forward part): the module in .train() under no_grad — BatchNorm batch
statistics with running-stat updates, LayerDrop coins from np.random.random(), dropout.  Goldens come from the REAL
reference in .train() (oracle/make_golden_train.py, every dropout probability 0 so the output is deterministic)."""
import ast
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import avhubert_oracle as ao

from helpers import GOLDEN, cosine, rel_err

CASES = ["tiny_train", "tiny_train_layerdrop"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"enc_{name}.npz"))
    over, B, T, lengths, npseed = [str(v) for v in z["meta"]]
    over = ast.literal_eval(over)
    return z, over, int(B), int(T), ast.literal_eval(lengths), int(npseed)


def bn_flat(model):
    mods = model._bn_modules() if hasattr(model, "_bn_modules") else None
    if mods is None:
        from oracle.make_golden_train import bn_modules
        mods = bn_modules(model)
    return torch.cat([torch.cat([m.running_mean.float().cpu(), m.running_var.float().cpu()]) for m in mods])


@pytest.mark.parametrize("name", CASES)
def test_oracle_in_train_mode_matches_the_real_reference(name):
    z, over, B, T, lengths, npseed = load_case(name)
    o_over = {k: v for k, v in over.items() if k != "encoder_layerdrop"}
    oracle = ao.build_oracle("tiny", seed=1234, **o_over)
    src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=17)
    oracle.train()
    oracle.encoder.layer_skip = z["skip"].tolist()
    with torch.no_grad():
        y, _ = oracle.extract_finetune(src, pm)
    assert np.abs(y.numpy() - z["y"]).max() < 2e-4
    assert np.abs(bn_flat(oracle).numpy() - z["bn"]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_device_train_mode_forward_matches_reference_golden(name, dtype):
    from helpers import make_device_model, to_dev
    z, over, B, T, lengths, npseed = load_case(name)
    o_over = {k: v for k, v in over.items() if k != "encoder_layerdrop"}
    oracle = ao.build_oracle("tiny", seed=1234, **o_over)
    m = make_device_model(oracle, {}, "tiny", dtype, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                          dropout_input=0.0, **over)
    src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=17)
    d_src, d_pm = to_dev(src, pm, dtype=dtype)
    m.train()
    np.random.seed(npseed)                       # the reference draws its LayerDrop coins from numpy's global stream
    y, pm_out = m.extract_finetune(d_src, d_pm)
    assert torch.equal(pm_out, d_pm)
    y = y.float().cpu()
    y_ref = torch.from_numpy(z["y"])
    bn, bn_ref = bn_flat(m), torch.from_numpy(z["bn"])
    if dtype == torch.float32:
        assert rel_err(y, y_ref) < 2e-3
        assert (bn - bn_ref).abs().max() < 1e-3 * bn_ref.abs().max()
    else:
        assert cosine(y, y_ref) > 0.999
        assert (bn - bn_ref).abs().max() < 2e-2 * bn_ref.abs().max()       # bf16 buffers, bf16 conv outputs
    assert all(int(b.num_batches_tracked) == 1 for b in m._bn_modules())
    # back to eval: the folded BatchNorm must use the UPDATED running statistics
    oracle.train()
    oracle.encoder.layer_skip = z["skip"].tolist()
    with torch.no_grad():
        oracle.extract_finetune(src, pm)
    oracle.eval()
    oracle.encoder.layer_skip = None
    with torch.no_grad():
        y_eval_ref, _ = oracle.extract_finetune(src, pm)
    m.eval()
    y_eval, _ = m.extract_finetune(d_src, d_pm)
    if dtype == torch.float32:
        assert rel_err(y_eval.float().cpu(), y_eval_ref) < 2e-3
    else:
        assert cosine(y_eval.float().cpu(), y_eval_ref) > 0.999


@pytest.mark.gpu
def test_dropout_kernel_is_nn_dropout_with_its_own_philox_stream():
    from multimodalvc_b200 import _lib
    lib = _lib.load()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    n = 1 << 20
    for dt, code in ((torch.float32, 0), (torch.bfloat16, 2)):
        for p in (0.1, 0.5):
            x = torch.ones(n, device="cuda", dtype=dt)
            _lib.check(lib.avh_dropout(vp(x.data_ptr()), code, n, p, 1234, 7, vp(st)))
            kept = (x != 0)
            frac = kept.float().mean().item()
            assert abs(frac - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5 + 1e-4
            assert torch.allclose(x[kept].float(), torch.full_like(x[kept].float(), 1 / (1 - p)), rtol=1e-2)
            y = torch.ones(n, device="cuda", dtype=dt)
            _lib.check(lib.avh_dropout(vp(y.data_ptr()), code, n, p, 1234, 7, vp(st)))
            assert torch.equal(x, y)                                           # same (seed, site): same mask
            w = torch.ones(n, device="cuda", dtype=dt)
            _lib.check(lib.avh_dropout(vp(w.data_ptr()), code, n, p, 1234, 8, vp(st)))
            assert not torch.equal(x, w)                                       # another site: another mask
            agree = ((x != 0) == (w != 0)).float().mean().item()
            assert abs(agree - (p * p + (1 - p) * (1 - p))) < 0.01             # independent masks


@pytest.mark.gpu
def test_train_mode_with_dropout_is_seeded_and_finite():
    from helpers import make_device_model, to_dev
    oracle = ao.build_oracle("tiny", seed=1234)
    m = make_device_model(oracle, {}, "tiny", torch.bfloat16, dropout=0.2, attention_dropout=0.0, activation_dropout=0.3,
                          dropout_input=0.1, encoder_layerdrop=0.0)
    src, pm = ao.synthetic_inputs(2, 16, lengths=[16, 11], seed=2)
    d_src, d_pm = to_dev(src, pm, dtype=torch.bfloat16)
    m.train()
    outs = []
    for seed in (5, 5, 6):
        torch.manual_seed(seed)
        y, _ = m.extract_finetune(d_src, d_pm)
        assert torch.isfinite(y).all()
        outs.append(y.float().cpu())
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
    m.eval()
    y_eval, _ = m.extract_finetune(d_src, d_pm)
    assert cosine(outs[0], y_eval.float().cpu()) > 0.5            # a perturbation of the same function, not noise
    with pytest.raises(NotImplementedError):
        m2 = make_device_model(oracle, {}, "tiny", torch.bfloat16, attention_dropout=0.1)
        m2.train()
        m2.extract_finetune(d_src, d_pm)


@pytest.mark.gpu
def test_train_mode_large_shape_vs_oracle():
    """BASELINE config 5's forward at the Large shape (reduced batch): train-mode BatchNorm over 2 x 40 frames, LayerDrop
    with the shipped probability 0.1 (numpy seed chosen so that layers are dropped), bf16."""
    from helpers import make_device_model, to_dev
    oracle = ao.build_oracle("large", seed=1234)
    m = make_device_model(oracle, {}, "large", torch.bfloat16, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                          dropout_input=0.0, encoder_layerdrop=0.1)
    src, pm = ao.synthetic_inputs(2, 40, lengths=[40, 27], seed=9)
    d_src, d_pm = to_dev(src, pm, dtype=torch.bfloat16)
    np.random.seed(11)
    skip = [0 if np.random.random() > 0.1 else 1 for _ in range(24)]
    assert sum(skip) >= 1
    oracle.train()
    oracle.encoder.layer_skip = skip
    with torch.no_grad():
        y_ref, _ = oracle.extract_finetune(src, pm)
    m.train()
    np.random.seed(11)
    y, _ = m.extract_finetune(d_src, d_pm)
    assert cosine(y.float().cpu(), y_ref) > 0.999
    from oracle.make_golden_train import bn_modules
    ref_bn = torch.cat([torch.cat([b.running_mean, b.running_var]) for b in bn_modules(oracle)])
    assert (bn_flat(m) - ref_bn).abs().max() < 2e-2 * ref_bn.abs().max()
