"""FLOP bookkeeping used by bench.py's roofline / cpu_baseline figures vs SURVEY.md section 8(d). CPU-only."""
from animate3d_b200.flops import attention_call_flops, unet_forward_flops
from animate3d_b200.unet_config import UNetConfig


def test_forward_flops_match_survey():
    f = unet_forward_flops(UNetConfig(), 1, 4, 16)
    t = {k: v / 1e12 for k, v in f.items()}
    assert abs(t["total"] - 25.97) < 0.01
    assert abs(t["conv"] - 7.103) < 0.002
    assert abs(t["mv_qkpv"] - 1.960) < 0.001 and abs(t["i2v_qkpv"] - 1.960) < 0.001
    assert abs(t["spatial_qkpv"] - 3.923) < 0.001
    assert abs(t["spatial_ff"] - 2.456) < 0.001 and abs(t["motion_ff"] - 2.658) < 0.001
    assert abs(unet_forward_flops(UNetConfig(), 2, 4, 16)["total"] / 1e12 - 51.94) < 0.01
    assert abs(unet_forward_flops(UNetConfig(), 1, 1, 4)["total"] / 1e12 - 1.25) < 0.01


def test_attention_call_flops():
    assert abs(attention_call_flops(16, 4096, 320) / 1e9 - 343.6) < 0.1
    assert abs(attention_call_flops(16, 1024, 640) / 1e9 - 42.9) < 0.1
