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


def test_load_unet_checkpoint_rules(tmp_path):
    """`load_unet_checkpoint` = inference.py:213-223 on a width-reduced model (the key NAMES, hence the 726 rule, do not
    depend on channel widths): motion-only file -> 726 missing, full file -> 0, a stray key or a partial file -> error."""
    import pytest

    from animate3d_b200.unet import MVUNetMotionModel
    from animate3d_b200.unet_config import UNetConfig, key_plan
    from animate3d_b200.weights import load_unet_checkpoint
    cfg = UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=16, ip_image_embed_dim=16, sample_size=8)
    plan = key_plan(cfg)
    assert set(plan) == set(key_plan(UNetConfig()))           # same names as the released geometry
    full = {k: torch.zeros(s) for k, s in plan.items()}
    motion = {k: v for k, v in full.items() if "motion_modules." in k or "i2v." in k}
    torch.save({"state_dict": motion, "epoch": 3}, tmp_path / "motion.ckpt")
    torch.save(full, tmp_path / "full.ckpt")
    unet = MVUNetMotionModel(cfg, device="cpu")
    m, u = load_unet_checkpoint(unet, str(tmp_path / "motion.ckpt"))
    assert len(m) == 726 and not u
    m, u = load_unet_checkpoint(unet, str(tmp_path / "full.ckpt"))
    assert not m and not u and len(unet.state_dict()) == len(plan)
    torch.save({**motion, "bogus.weight": torch.zeros(1)}, tmp_path / "bad.ckpt")
    with pytest.raises(ValueError, match="unexpected"):
        load_unet_checkpoint(MVUNetMotionModel(cfg, device="cpu"), str(tmp_path / "bad.ckpt"))
    some = dict(list(motion.items())[:-5])
    torch.save(some, tmp_path / "partial.ckpt")
    with pytest.raises(ValueError, match="missing"):
        load_unet_checkpoint(MVUNetMotionModel(cfg, device="cpu"), str(tmp_path / "partial.ckpt"))
