"""threestudio plugin boundary: registry semantics, registered names, Config parsing -- no GPU needed."""
import dataclasses

import pytest


def test_registry_matches_threestudio_semantics():
    from animate3d_b200 import registry

    @registry.register("a3d-test-main")
    class Main:
        tag = "main"

    @registry.register("a3d-test-mixin")
    class Mixin:
        extra = 1

    assert registry.find("a3d-test-main") is Main
    mixed = registry.find("a3d-test-main:a3d-test-mixin")            # threestudio/__init__.py:18-31
    assert issubclass(mixed, Main) and issubclass(mixed, Mixin) and mixed.__mro__[1] is Mixin
    with pytest.raises(ValueError):
        registry.register("a3d-test-main")(Main)
    with pytest.raises(KeyError):
        registry.find("a3d-no-such-module")


def test_reference_plugin_names_are_registered():
    import animate3d_b200.plugins  # noqa: F401  (what a threestudio user imports; mirrors launch.py:70-102 discovery)
    from animate3d_b200 import registry
    from animate3d_b200.guidance import AnimateMVDiffusionGuidance
    from animate3d_b200.renderer import DiffGaussian4D
    assert registry.find("animatemv-diffusion-guidance") is AnimateMVDiffusionGuidance
    assert registry.find("diff-gaussian-rasterizer-advanced-4d") is DiffGaussian4D
    ref_fields = {"invert_bg_prob": 1.0, "back_ground_color": (1, 1, 1), "first_frame_trainable": False}
    got = {f.name: f.default for f in dataclasses.fields(DiffGaussian4D.Config)}
    for k, v in ref_fields.items():                                   # diff_gaussian_rasterizer_advanced_4d.py:25-30
        assert got[k] == v
    gf = {f.name for f in dataclasses.fields(AnimateMVDiffusionGuidance.Config)}
    for k in ("pretrained_model_name_or_path", "motion_adapter_path", "ip_adapter_path", "pretrained_unet_path", "model_config",
              "guidance_scale", "grad_clip", "half_precision_weights", "min_step_percent", "max_step_percent", "sqrt_anneal",
              "trainer_max_steps", "camera_condition_type", "view_dependent_prompting", "i2v_cond_time_zero", "n_view", "n_frame",
              "image_size", "recon_loss", "recon_std_rescale", "noise_scheduler_kwargs"):   # animatemv_guidance.py:56-101
        assert k in gf, k


def test_schedule_helper_C():
    from animate3d_b200.registry import C
    assert C(0.3, 0, 10) == 0.3
    assert C([0, 1.0, 3.0, 100], 0, 50) == pytest.approx(2.0)        # int end_step -> global_step
    assert C([1.0, 3.0, 100], 0, 25) == pytest.approx(1.5)           # 3 entries: start_step 0
    assert C([0, 1.0, 3.0, 4.0], 2, 999) == pytest.approx(2.0)       # float end_step -> epoch
    assert C([0, 1.0, 3.0, 100], 0, 500) == 3.0
