"""CPU-side checks: model-name expansion, blob reader, C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re
import numpy as np
import pytest
from gnina_b200 import model_blob, scorer, synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_expand_model_names_like_reference():
    # cnn_torch_scorer.cpp:28-62
    assert scorer.expand_model_names([]) == ["dense_1_3", "dense_1_3_PT_KD_3", "crossdock_default2018_KD_4"]
    assert scorer.expand_model_names(["fast"]) == ["all_default_to_default_1_3_1"]
    assert scorer.expand_model_names(["default1.0"]) == ["dense", "general_default2018_3", "dense_3",   # all packaged
                                                         "crossdock_default2018", "redock_default2018_2"]
    assert scorer.expand_model_names(["dense_1.3"]) == ["dense_1_3"]
    # an ensemble expands over the REFERENCE's table (64 models), not over whatever happens to be packaged
    assert len(scorer.REFERENCE_MODELS) == 64
    ens = scorer.expand_model_names(["crossdock_default2018_ensemble"], check=False)
    assert len(ens) == 15 and "crossdock_default2018" in ens and all(e.startswith("crossdock_default2018") for e in ens)
    assert len(scorer.expand_model_names(["dense_ensemble"], check=False)) == 20
    assert len(scorer.expand_model_names(["redock_default2018_ensemble"], check=False)) == 15
    # a member that is not packaged is an error, never a silently smaller ensemble
    missing = [m for m in ens if m not in scorer.builtin_models()]
    if missing:
        with pytest.raises(scorer.usage_error, match="not packaged"):
            scorer.expand_model_names(["crossdock_default2018_ensemble"])


def test_blob_reader_roundtrip():
    b = model_blob.load_model("crossdock_default2018")
    assert b.arch == "default2018" and b.resolution == 0.5 and b.dimension == 23.5
    assert b.tensors["unit1_conv.weight"].shape == (32, 28, 3, 3, 3)
    assert b.tensors["pose_output.weight"].shape == (2, 27648)
    d = model_blob.load_model("dense_1.3")
    assert d.arch == "dense" and d.tensors["data_enc_level1_bottleneck.weight"].shape == (160, 160, 1, 1, 1)
    with pytest.raises(FileNotFoundError, match="Invalid model name"):
        model_blob.load_model("nope")


def test_all_64_reference_models_are_packaged():
    """Every model the reference embeds (gninasrc/CMakeLists.txt:95-188) has its blob; architecture and head shapes are read
    from each (cheap header + shape checks; the numerics are the GPU tests')."""
    assert sorted(scorer.builtin_models()) == sorted(scorer.REFERENCE_MODELS)
    archs = {}
    for name in scorer.REFERENCE_MODELS:
        b = model_blob.load_model(name)
        archs[b.arch] = archs.get(b.arch, 0) + 1
        assert b.resolution == 0.5 and b.dimension == 23.5
        if b.arch == "default2018":
            assert b.tensors["unit1_conv.weight"].shape == (32, 28, 3, 3, 3) and b.tensors["affinity_output.weight"].shape == (1, 27648)
        elif b.arch == "dense":
            assert b.tensors["pose_output.weight"].shape == (2, 224)
    assert archs == {"default2018": 48, "dense": 15, "default2017": 1}
    assert len(scorer.expand_model_names(["crossdock_default2018_ensemble"])) == 15


def test_synth_is_deterministic_and_shaped():
    x1, t1 = synth.make_receptor(200, box=30)
    x2, t2 = synth.make_receptor(200, box=30)
    assert np.array_equal(x1, x2) and np.array_equal(t1, t2)
    d = np.linalg.norm(x1[:, None] - x1[None], axis=-1) + np.eye(200) * 10
    assert d.min() >= 1.2 - 1e-5
    lx, lt = synth.make_ligand()
    px, off = synth.make_poses(lx, 5)
    assert px.shape == (5 * 34, 3) and off[-1] == 5 * 34 and (lt == 1).sum() == 4
    sx, st, so = synth.make_screen(7)
    assert len(so) == 8 and so[-1] == len(st) == len(sx)


def test_capi_library_exports_every_declared_symbol():
    from gnina_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "gnina_b200.h")).read()
    declared = set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym



def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from gnina_b200 import CNNScorer, capi
    with pytest.raises(capi.GbError, match="no CPU fallback"):
        CNNScorer(["crossdock_default2018"])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gnina_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
