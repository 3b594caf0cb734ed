"""The C-ABI library builds, loads without a GPU and exports every symbol include/avh_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "avh_b200.h")).read()
    return sorted(set(re.findall(r"AVH_API\s+[\w\s\*]+?\b(avh_\w+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ["avh_create", "avh_destroy", "avh_load_tensor", "avh_finalize_weights", "avh_forward",
              "avh_forward_host", "avh_fbank", "avh_add_noise", "avh_last_error", "avh_abi_version"]:
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    from multimodalvc_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in avh_b200.h but not exported"
    assert set(declared_symbols()) == set(_lib.EXPORTS)
    assert _lib.load().avh_abi_version() == 1


def test_config_struct_matches_header_size():
    from multimodalvc_b200 import _lib
    assert ctypes.sizeof(_lib.AvhConfig) == 16 * 4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from multimodalvc_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _lib.load()


def test_cpu_module_refuses_to_compute():
    import torch
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    m = AVHubertModel(AVHubertConfig.named("tiny")).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.extract_finetune({"audio": torch.zeros(1, 104, 4), "video": None})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multimodalvc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
