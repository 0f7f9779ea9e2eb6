"""CPU: the oracle UNet restatement and the drop-in module layout against the golden vectors that
tests/golden/make_golden.py produced by running the reference's own model.py."""
import hashlib

import numpy as np
import torch

from conftest import randomize_bn
from oracle import unet_oracle
from svs_unet_pytorch_b200 import model as svs_model


def _net(seed=0):
    torch.manual_seed(seed)
    return svs_model.UNet()


def _x(seed, n=2):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 1, 512, 128, generator=g)


def test_state_dict_layout_matches_reference(golden):
    net = _net()
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in golden["sd_keys"]]          # 79 keys, same order
    assert len(sd) == 79
    assert sum(p.numel() for p in net.parameters()) == int(golden["n_params"]) == 9823313
    # same seed -> same random init as reference model.py (module creation order preserved)
    np.testing.assert_allclose([float(v.double().sum()) for v in sd.values()], golden["sd_sum"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([float(v.double().abs().sum()) for v in sd.values()], golden["sd_abs_sum"], rtol=1e-12)


def test_state_dict_shapes():
    sd = _net().state_dict()
    assert sd["conv1.0.weight"].shape == (16, 1, 5, 5)
    assert sd["conv6.0.weight"].shape == (512, 256, 5, 5)
    assert sd["deconv1.weight"].shape == (512, 256, 5, 5)        # ConvTranspose2d: (Cin, Cout, 5, 5)
    assert sd["deconv6.weight"].shape == (32, 1, 5, 5)
    assert "deconv6_BAD.0.weight" not in sd                       # deconv6 has no BatchNorm
    assert sd["deconv5_BAD.0.running_var"].shape == (16,)


def test_oracle_forward_matches_reference_default_bn(golden):
    net = _net().eval()
    x = _x(1)
    assert hashlib.sha1(x.numpy().tobytes()).hexdigest() == str(golden["x_sha1"])
    with torch.no_grad():
        mask, acts = unet_oracle.unet_forward(net.state_dict(), x, return_activations=True)
    assert mask.shape == (2, 1, 512, 128)
    np.testing.assert_allclose(mask[:, 0, ::4, ::4].numpy(), golden["mask_default_sub"], rtol=0, atol=2e-6)
    assert abs(float(mask.double().sum()) - float(golden["mask_default_sum"])) < 1e-1
    names = [str(n) for n in golden["act_names"]]
    for i, n in enumerate(names):
        a = acts[n.replace("_BAD", "")]
        assert abs(float(a.double().abs().mean()) - golden["act_default_absmean"][i]) < 1e-5 * max(1.0, golden["act_default_absmean"][i])


def test_oracle_forward_matches_reference_random_bn(golden):
    net = _net().eval()
    randomize_bn(net, seed=2)
    assert float(golden["mask_bn_fp32_vs_fp64"]) < 1e-5          # conditioning guard
    with torch.no_grad():
        mask = unet_oracle.unet_forward(net.state_dict(), _x(1))
    np.testing.assert_allclose(mask[:, 0, ::4, ::4].numpy(), golden["mask_bn_sub"], rtol=0, atol=5e-6)


def test_oracle_training_loss_and_grads_match_reference(golden):
    net = _net().train()
    g = torch.Generator().manual_seed(3)
    mix = torch.rand(2, 1, 512, 128, generator=g)
    voc = mix * torch.rand(2, 1, 512, 128, generator=g)
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k)
              for k, v in net.state_dict().items()}
    mask = unet_oracle.unet_forward(params, mix, training=True)
    loss = unet_oracle.l1_masked_loss(mask, mix, voc, two_term=True)
    assert abs(float(loss.detach()) - golden["train_loss"][0]) < 1e-6
    loss.backward()
    names = [str(n) for n in golden["train_param_names"]]
    got = np.array([float(params[n].grad.double().norm()) for n in names])
    # conv biases feeding a BatchNorm have an analytically zero gradient (noise ~1e-8): absolute floor
    np.testing.assert_allclose(got, golden["train_grad_l2"], rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(mask.detach()[:, 0, ::4, ::4].numpy(), golden["train_mask_sub"], atol=5e-6)


def test_masked_l1_loss_call_shapes():
    crit = svs_model.MaskedL1Loss()
    a, b, m = torch.rand(2, 1, 8, 8), torch.rand(2, 1, 8, 8), torch.rand(2, 1, 8, 8)
    assert torch.allclose(crit(a, b), (a - b).abs().mean())
    assert torch.allclose(crit(a, b, m), unet_oracle.l1_masked_loss(m, b, a))


def test_separate_spectrogram_patch_logic():
    # reference inference.py:75,88: T//128 + 1 segments, the empty one skipped; DC row re-inserted as zeros
    net = _net().eval()
    sd = net.state_dict()
    for t in (65, 128, 129):
        spec = np.random.default_rng(t).random((513, t), dtype=np.float32)
        out = unet_oracle.separate_spectrogram(sd, spec, vocal_solo=True)
        assert out.shape == (513, t) and out.dtype == np.float32
        assert np.all(out[0] == 0)
        assert np.all(out[1:] <= spec[1:] + 1e-6) and np.all(out >= 0)


def test_product_forward_refuses_cpu():
    import pytest
    net = _net().eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(torch.rand(1, 1, 512, 128))
