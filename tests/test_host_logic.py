"""CPU tests of the host-side logic: the boundary mirror, weight folding, sharding, FLOP accounting."""
import numpy as np
import pytest

from oracle import hift_oracle as H


def test_backend_registers_under_the_reference_id():
    from vocalie_tts_b200 import backend as B
    assert B.TTSBackend._REGISTRY["chatterbox"] is B.ChatterboxB200Backend
    cls = B.ChatterboxB200Backend
    # class attributes the pipeline reads (reference chatterbox_backend.py:21-25)
    assert (cls.supports_ref_audio, cls.uses_internal_voices, cls.supports_inter_chunk_gap) == (True, False, True)
    b = cls()
    assert b.map_language(None) == "fr" and b.map_language("en-US") == "en"
    assert b.default_language() == "fr-FR"
    schema = b.params_schema()
    assert set(schema) == {"chatterbox_mode", "multilang_cfg_weight", "exaggeration", "cfg_weight", "temperature",
                           "repetition_penalty"}
    for k, spec in schema.items():
        assert spec.key == k
        if spec.type in ("float", "int"):
            assert spec.min is not None and spec.max is not None and spec.step is not None


def test_unconfigured_backend_raises_the_single_error_type(tmp_path):
    from vocalie_tts_b200 import backend as B, BackendUnavailableError
    B.ChatterboxB200Backend.reset()
    assert not B.ChatterboxB200Backend.is_available()
    assert "configure" in B.ChatterboxB200Backend.unavailable_reason()
    b = B.ChatterboxB200Backend()
    with pytest.raises(BackendUnavailableError):
        b.synthesize_chunk("Bonjour.")
    with pytest.raises(BackendUnavailableError):
        b.synthesize("Bonjour.", str(tmp_path / "x.wav"))
    assert issubclass(BackendUnavailableError, RuntimeError)


def test_engine_params_match_reference_defaults_and_ignore_unknown_keys():
    from vocalie_tts_b200.backend import ChatterboxB200Backend as C
    p = C._engine_params({"voice": "x", "model_id": "y", "inter_chunk_gap_ms": 250, "temperature": 0.7,
                          "chatterbox_mode": "multilang"})
    assert p == {"tts_model_mode": "multilang", "multilang_cfg_weight": 0.5, "exaggeration": 0.5, "cfg_weight": 0.6,
                 "temperature": 0.7, "repetition_penalty": 1.35}
    assert C._engine_params({})["tts_model_mode"] == "fr_finetune"


def test_configure_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from vocalie_tts_b200 import backend as B, BackendUnavailableError
    with pytest.raises(BackendUnavailableError):
        B.ChatterboxB200Backend.configure(state_dict={"conv_pre.weight": np.zeros((512, 80, 7), np.float32)},
                                          mel_provider=lambda *a, **k: None)
    assert not B.ChatterboxB200Backend.is_available()
    B.ChatterboxB200Backend.reset()


def test_shard_chunks_balances_and_keeps_order():
    from vocalie_tts_b200.backend import shard_chunks
    rng = np.random.default_rng(1004)
    lens = (50 * rng.integers(1, 21, 2048)).tolist()
    for world in (1, 2, 4, 8):
        parts = shard_chunks(lens, world)
        assert sorted(i for p in parts for i in p) == list(range(len(lens)))
        assert all(p == sorted(p) for p in parts)
        loads = [sum(lens[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lens)      # LPT bound
    assert shard_chunks([], 4) == [[], [], [], []]
    assert shard_chunks([5], 2) == [[0], []]


def test_fold_weight_norm_matches_oracle():
    from vocalie_tts_b200.hift import fold_weight_norm
    sd = H.make_state_dict(0, "unit")
    ours = fold_weight_norm(sd)
    ref = H.fold_weight_norm(sd)
    assert set(ours) == set(ref)
    for k in ref:
        np.testing.assert_allclose(ours[k], ref[k].numpy(), rtol=1e-6, atol=1e-8, err_msg=k)
    # legacy weight_g / weight_v naming folds too
    legacy = {"conv_pre.weight_g": sd["conv_pre.parametrizations.weight.original0"],
              "conv_pre.weight_v": sd["conv_pre.parametrizations.weight.original1"], "conv_pre.bias": sd["conv_pre.bias"]}
    np.testing.assert_allclose(fold_weight_norm(legacy)["conv_pre.weight"], ref["conv_pre.weight"].numpy(), rtol=1e-6)


def test_random_state_dict_has_upstream_names_and_shapes():
    from vocalie_tts_b200.hift import random_state_dict, fold_weight_norm
    ours = fold_weight_norm(random_state_dict(0))
    ref = H.fold_weight_norm(H.make_state_dict(0, "init"))
    assert set(ours) == set(ref)
    for k in ref:
        assert tuple(ours[k].shape) == tuple(ref[k].shape), k


def test_algorithmic_flops_per_frame_matches_survey():
    from vocalie_tts_b200.hift import algorithmic_flops_per_frame
    # 612.451 MFLOP incl. 0.155 for STFT/iSTFT/SineGen which are not convs
    assert abs(algorithmic_flops_per_frame() / 1e6 - (612.451 - 0.155)) < 0.01
    assert abs((algorithmic_flops_per_frame() - algorithmic_flops_per_frame(False)) / 1e6 - 6.538) < 0.01
    # cross-check against the oracle's layer table
    total = 0.0
    steps = {"conv_pre": 1, "ups.0": 1, "ups.1": 8, "ups.2": 40, "source_downs.0": 8, "source_downs.1": 40,
             "source_downs.2": 120, "conv_post": 120}
    for name, kind, cin, cout, k, wn in H.layer_table():
        if name in steps:
            n = steps[name]
        elif "resblocks" in name:
            n = (8, 40, 120)[(int(name.split(".")[1]) // 3) if name.startswith("resblocks") else int(name.split(".")[1])]
        else:
            n = 1   # f0 predictor
        total += 2.0 * cin * cout * k * n
    total += 2 * 512   # classifier
    assert abs(total - algorithmic_flops_per_frame()) < 1.0


def test_bench_clock_summary_uses_timed_samples_and_all_reasons():
    """bench.py's clocks object: SM clock / power from the samples of the timed region only, throttle reasons from the
    warm-up too (the power-cap flag is not always up while a sample of a ~140 ms region is taken)."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("vt_bench", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    na, ac = "Not Active", "Active"
    rows = [["0", "1965", "1965", "210.0", na, na, na, na],        # warm-up: idle clock, no reason
            ["0", "1700", "1965", "900.0", na, na, na, ac],        # warm-up: capped
            ["0", "1650", "1965", "950.0", na, na, na, na],        # timed region: the flag happens to be down
            ["0", "1670", "1965", "960.0", na, na, na, na],
            ["0", "bad", "1965", "960.0", na, na, na, na],         # unparsable sample is skipped
            ["0", "1660", "1965", "940.0", na, na, na, na]]
    c = bench.ClockSampler.summarize(rows, 2, {"sync_boost"})
    assert c["sm_mhz"] == 1660.0 and c["sm_max_mhz"] == 1965.0 and c["samples"] == 3
    assert c["reasons"] == ["sw_power_cap", "sync_boost"] and c["power_w_max"] == 960.0
    assert bench.ClockSampler.summarize([], 0)["reasons"] == ["no samples"]
    # no mark (or a mark past the end): every sample counts
    assert bench.ClockSampler.summarize(rows[:2], 5)["samples"] == 2
