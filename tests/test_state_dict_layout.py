"""State-dict key layout invariants of the reference (inference.py:214-223). CPU-only."""
import torch

from oracle import unet_oracle as O


def test_726_missing_keys_invariant():
    plan = O.key_plan(O.UNetConfig())
    released = [k for k in plan if "motion_modules." in k or "i2v." in k]   # train.py:350-356 trainable subset
    missing = [k for k in plan if k not in set(released)]
    assert len(missing) == 726
    stock = [k for k in missing if not (k.startswith("camera_embedding") or k.startswith("encoder_hid_proj")
                                        or "_ip." in k)]
    assert len(stock) == 686                     # tensors of a stock SD1.5 UNet2DConditionModel
    assert len([k for k in missing if k.startswith("camera_embedding")]) == 4
    assert len([k for k in missing if k.startswith("encoder_hid_proj")]) == 4
    assert len([k for k in missing if "_ip." in k]) == 32


def test_parameter_count_and_up_plan():
    cfg = O.UNetConfig()
    plan = O.key_plan(cfg)
    n = sum(torch.Size(s).numel() for s in plan.values())
    assert abs(n / 1e9 - 1.527) < 0.01
    cins = [c for (cl, _, _, _) in O.up_plan(cfg) for c in cl]
    assert cins == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]
    assert O.skip_channels(cfg) == [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
