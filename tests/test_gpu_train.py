"""GPU: training step (train-mode forward, masked-L1 loss, backward) against the golden vectors produced
by the reference's model.py autograd on CPU (tests/golden/make_golden.py) and against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import unet_oracle  # noqa: E402
from svs_unet_pytorch_b200 import model as svs_model, training  # noqa: E402


def _net(p_drop=0.0):
    torch.manual_seed(0)
    net = svs_model.UNet().train().cuda()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = p_drop
    return net


def _data(seed=3, n=2):
    g = torch.Generator().manual_seed(seed)
    mix = torch.rand(n, 1, 512, 128, generator=g)
    voc = mix * torch.rand(n, 1, 512, 128, generator=g)
    return mix, voc


def test_train_forward_backward_match_reference_golden(golden):
    net = _net(0.0)
    mix, voc = _data()
    mix_d, voc_d = mix.cuda(), voc.cuda()
    mask = net(mix_d)                                               # UNet.forward in train mode (autograd)
    np.testing.assert_allclose(mask.detach()[:, 0, ::4, ::4].cpu().numpy(), golden["train_mask_sub"], atol=2e-5)
    crit = torch.nn.L1Loss()                                        # reference train.py:275-283
    loss = crit(mask * mix_d, voc_d) + crit((1 - mask) * mix_d, torch.clamp(mix_d - voc_d, min=0.0))
    assert abs(float(loss.detach()) - golden["train_loss"][0]) <= 1e-4 * golden["train_loss"][0]   # rel 1e-4
    loss.backward()
    names = [str(n) for n in golden["train_param_names"]]
    params = dict(net.named_parameters())
    got = np.array([float(params[n].grad.double().norm()) for n in names])
    np.testing.assert_allclose(got, golden["train_grad_l2"], rtol=1e-3, atol=1e-7)
    got_sum = np.array([float(params[n].grad.double().sum()) for n in names])
    np.testing.assert_allclose(got_sum, golden["train_grad_sum"], rtol=5e-3, atol=2e-5)
    # running statistics after one step (momentum 0.1, unbiased variance)
    bufs = dict(net.named_buffers())
    bn = [str(n) for n in golden["train_buffer_names"]]
    got_buf = np.array([float(bufs[n].double().sum()) for n in bn])
    np.testing.assert_allclose(got_buf, golden["train_buffer_sum"], rtol=1e-4, atol=1e-5)
    assert int(bufs["conv1.1.num_batches_tracked"]) == 1


def test_grads_match_oracle_elementwise_with_dropout_masks():
    net = _net(0.5)
    mix, voc = _data(seed=11, n=3)
    g = torch.Generator().manual_seed(5)
    masks = {f"deconv{i}": torch.rand(3, c, generator=g) >= 0.5 for i, c in zip(range(1, 6), [256, 128, 64, 32, 16])}
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    params = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    ref_mask = unet_oracle.unet_forward(params, mix, training=True, dropout_masks=masks)
    ref_loss = unet_oracle.l1_masked_loss(ref_mask, mix, voc, two_term=True)
    ref_loss.backward()
    loss = training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False, injected_masks=masks)
    assert abs(float(loss[0]) - float(ref_loss.detach())) <= 1e-4 * float(ref_loss.detach())
    assert abs(float(loss[0]) - float(loss[1]) - float(loss[2])) < 1e-6
    for name, p in net.named_parameters():
        ref = params[name].grad.float()
        got = p.grad.cpu()
        denom = max(float(ref.norm()), 1e-6)
        assert float((got - ref).norm()) / denom <= 1e-3 or float((got - ref).abs().max()) < 1e-7, name


def test_train_step_is_bit_reproducible_and_updates_weights():
    mix, voc = _data(seed=21, n=2)
    outs = []
    for _ in range(2):
        net = _net(0.0)
        w0 = net.conv3[0].weight.detach().clone()
        loss = training.train_step(net, mix.cuda(), voc.cuda())
        outs.append((loss.clone(), net._flat_grad.clone(), net.conv3[0].weight.detach().clone()))
        assert not torch.equal(w0, outs[-1][2])                     # Adam moved the weights (model.py:116)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2])


def test_fused_loss_kernel_matches_torch():
    g = torch.Generator().manual_seed(1)
    m = torch.rand(2, 1, 512, 128, generator=g).cuda().requires_grad_(True)
    mix, voc = (t.cuda() for t in _data(seed=2))
    for two in (True, False):
        ref = unet_oracle.l1_masked_loss(m, mix, voc, two_term=two)
        (gref,) = torch.autograd.grad(ref, m)
        loss, grad = training.masked_l1(m.detach(), mix, voc, two_term=two)
        assert abs(float(loss[0]) - float(ref)) < 1e-6
        assert torch.allclose(grad, gref, atol=1e-9)


def _oracle_step(net, mix, voc, masks=None, dtype=torch.float32):
    """Oracle autograd step.  dtype=float64 makes the tolerance measure OUR error only: at batch 64 a weight
    gradient is a sum over up to a million pixels with heavy cancellation, and the fp32 summation order of the CPU
    oracle itself is then worth ~1e-3 of the result."""
    sd = {k: (v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu().clone())
          for k, v in net.state_dict().items()}
    params = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    ref_mask = unet_oracle.unet_forward(params, mix.to(dtype), training=True, dropout_masks=masks)
    ref_loss = unet_oracle.l1_masked_loss(ref_mask, mix.to(dtype), voc.to(dtype), two_term=True)
    ref_loss.backward()
    return float(ref_loss.detach()), params


def _check_grads(net, params, rel=1e-3):
    for name, p in net.named_parameters():
        ref = params[name].grad.float()
        got = p.grad.cpu()
        denom = max(float(ref.norm()), 1e-6)
        assert float((got - ref).norm()) / denom <= rel or float((got - ref).abs().max()) < 1e-7, \
            (name, float((got - ref).norm()) / denom)


def test_train_step_at_the_benched_batch_64_matches_oracle_autograd():
    # BASELINE configs[4] / SURVEY 8(d) config 5: batch 64 per GPU, mix = rand, voc = mix * rand, dropout masks
    # injected (torch's Philox stream cannot be matched), BatchNorm in train mode
    net = _net(0.5)
    mix, voc = _data(seed=0, n=64)
    g = torch.Generator().manual_seed(9)
    masks = {f"deconv{i}": torch.rand(64, c, generator=g) >= 0.5 for i, c in zip(range(1, 6), [256, 128, 64, 32, 16])}
    ref_loss, params = _oracle_step(net, mix, voc, masks, dtype=torch.float64)
    loss = training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False, injected_masks=masks)
    assert abs(float(loss[0]) - ref_loss) <= 1e-4 * ref_loss          # loss rel 1e-4
    _check_grads(net, params, rel=1e-3)                               # grad rel-L2 1e-3 (TF32 bound of SURVEY 8d)


def test_train_step_batch_one_matches_oracle():
    # torch BatchNorm trains with B = 1 whenever H*W > 1 (every layer here has >= 16 pixels per channel)
    net = _net(0.0)
    mix, voc = _data(seed=4, n=1)
    ref_loss, params = _oracle_step(net, mix, voc)
    loss = training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False)
    assert abs(float(loss[0]) - ref_loss) <= 1e-4 * ref_loss
    _check_grads(net, params, rel=1e-3)


def test_stale_backward_raises():
    net = _net(0.0)
    mix, voc = _data(seed=6, n=2)
    m1 = net(mix.cuda())
    _ = net(mix.cuda())                                               # a second train-mode forward overwrites the workspace
    with pytest.raises(RuntimeError, match="most recent train-mode forward"):
        m1.sum().backward()
