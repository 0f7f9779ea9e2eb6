"""GPU: training step (train-mode forward, masked-L1 loss, backward) against the golden vectors produced
by the reference's model.py autograd on CPU (tests/golden/make_golden.py) and against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import unet_oracle  # noqa: E402
from svs_unet_pytorch_b200 import model as svs_model, training  # noqa: E402


def _net(p_drop=0.0, precision="fp32"):
    """precision: "fp32" = exact CUDA-core arithmetic (parity mode), "tf32" = tcgen05 kind::tf32 (the default, and
    what torch + cuDNN do for the reference's fp32 model on a GPU)."""
    torch.manual_seed(0)
    net = svs_model.UNet().train().cuda()
    net.train_precision = precision
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = p_drop
    return net


# gradient tolerances (rel-L2 per tensor against the oracle).  fp32: SURVEY 8(d).  tf32: products carry 10-bit
# mantissas and BatchNorm's backward cancels the two largest components of every gradient, so the reference's own
# GPU arithmetic (torch eager + cuDNN, allow_tf32 = True, the torch default for convolutions) is 1.5e-2 .. 7e-2 away
# from float64 on these tensors (tests/diag_train.py; test_tf32_step_is_as_accurate_as_torch_cudnn below pins ours to it)
GRAD_TOL = {"fp32": 1e-3, "tf32": 1e-1}


def _data(seed=3, n=2):
    g = torch.Generator().manual_seed(seed)
    mix = torch.rand(n, 1, 512, 128, generator=g)
    voc = mix * torch.rand(n, 1, 512, 128, generator=g)
    return mix, voc


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_train_forward_backward_match_reference_golden(golden, precision):
    net = _net(0.0, precision)
    tf32 = precision == "tf32"
    mix, voc = _data()
    mix_d, voc_d = mix.cuda(), voc.cuda()
    mask = net(mix_d)                                               # UNet.forward in train mode (autograd)
    np.testing.assert_allclose(mask.detach()[:, 0, ::4, ::4].cpu().numpy(), golden["train_mask_sub"],
                               atol=1e-3 if tf32 else 2e-5)           # north_star: fp32/TF32 mask within 1e-3
    crit = torch.nn.L1Loss()                                        # reference train.py:275-283
    loss = crit(mask * mix_d, voc_d) + crit((1 - mask) * mix_d, torch.clamp(mix_d - voc_d, min=0.0))
    assert abs(float(loss.detach()) - golden["train_loss"][0]) <= 1e-4 * golden["train_loss"][0]   # rel 1e-4
    loss.backward()
    names = [str(n) for n in golden["train_param_names"]]
    params = dict(net.named_parameters())
    got = np.array([float(params[n].grad.double().norm()) for n in names])
    np.testing.assert_allclose(got, golden["train_grad_l2"], rtol=GRAD_TOL[precision], atol=1e-7)
    got_sum = np.array([float(params[n].grad.double().sum()) for n in names])
    if not tf32:                                                    # sums cancel: meaningful for exact arithmetic only
        np.testing.assert_allclose(got_sum, golden["train_grad_sum"], rtol=5e-3, atol=2e-5)
    # running statistics after one step (momentum 0.1, unbiased variance)
    bufs = dict(net.named_buffers())
    bn = [str(n) for n in golden["train_buffer_names"]]
    got_buf = np.array([float(bufs[n].double().sum()) for n in bn])
    # (running means are near-zero sums of signed values: absolute tolerance under TF32)
    np.testing.assert_allclose(got_buf, golden["train_buffer_sum"], rtol=1e-4, atol=5e-4 if tf32 else 1e-5)
    assert int(bufs["conv1.1.num_batches_tracked"]) == 1


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_grads_match_oracle_elementwise_with_dropout_masks(precision):
    net = _net(0.5, precision)
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
        assert float((got - ref).norm()) / denom <= GRAD_TOL[precision] or float((got - ref).abs().max()) < 1e-7, name


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_train_step_is_bit_reproducible_and_updates_weights(precision):
    mix, voc = _data(seed=21, n=2)
    outs = []
    for _ in range(2):
        net = _net(0.0, precision)
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
        assert abs(float(loss[0]) - float(ref.detach())) < 1e-6
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


def _torch_cudnn_grads(net, mix, voc, allow_tf32):
    """The reference's own GPU arithmetic: the same torch modules through torch eager + cuDNN (no dropout)."""
    import bench
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    try:
        tnet = svs_model.UNet().train().cuda()
        tnet.load_state_dict(net.state_dict())
        for m in tnet.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p = 0.0
        x, v = mix.cuda(), voc.cuda()
        mask = bench.torch_eager_forward(tnet, x)
        loss = (mask * x - v).abs().mean() + ((1 - mask) * x - torch.clamp(x - v, min=0)).abs().mean()
        loss.backward()
        return {n: p.grad.detach().cpu() for n, p in tnet.named_parameters()}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(got, ref):
    return float((got.double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-12)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_train_step_at_the_benched_batch_64_matches_oracle_autograd(precision):
    # BASELINE configs[4] / SURVEY 8(d) config 5: batch 64 per GPU, mix = rand, voc = mix * rand, BatchNorm in train
    # mode, against the float64 oracle.  At this size a weight gradient is a sum over up to a million pixels with
    # heavy cancellation (BatchNorm's backward removes the two largest components): fp32 arithmetic itself is only
    # good to a few 1e-3 here (torch + cuDNN fp32: 3.8e-3), TF32 to ~6e-2 (torch + cuDNN TF32: 6.6e-2).  So the bar
    # is relative: per tensor no worse than 1.5 x the reference's own GPU arithmetic at the same precision, and an
    # absolute cap.
    net = _net(0.0, precision)
    mix, voc = _data(seed=0, n=64)
    ref_loss, params = _oracle_step(net, mix, voc, None, dtype=torch.float64)
    cudnn = _torch_cudnn_grads(net, mix, voc, allow_tf32=precision == "tf32")
    loss = training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False)
    assert abs(float(loss[0]) - ref_loss) <= 1e-4 * ref_loss          # loss rel 1e-4
    cap = 1e-2 if precision == "fp32" else 1e-1
    for name, p in net.named_parameters():
        ref = params[name].grad
        if float(ref.norm()) < 1e-9:                                  # conv biases under BatchNorm: identically zero
            assert float(p.grad.abs().max()) < 1e-6, name
            continue
        ours, theirs = _rel(p.grad.cpu(), ref), _rel(cudnn[name], ref)
        assert ours <= cap, (name, ours)
        assert ours <= max(1.5 * theirs, 1e-3), (name, ours, theirs)


def test_train_step_batch_64_with_injected_dropout_masks_matches_oracle():
    net = _net(0.5, "fp32")
    mix, voc = _data(seed=0, n=64)
    g = torch.Generator().manual_seed(9)
    masks = {f"deconv{i}": torch.rand(64, c, generator=g) >= 0.5 for i, c in zip(range(1, 6), [256, 128, 64, 32, 16])}
    ref_loss, params = _oracle_step(net, mix, voc, masks, dtype=torch.float64)
    loss = training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False, injected_masks=masks)
    assert abs(float(loss[0]) - ref_loss) <= 1e-4 * ref_loss
    _check_grads(net, params, rel=1e-2)                               # see the comment in the test above


def test_train_step_batch_one_matches_oracle():
    # torch BatchNorm trains with B = 1 whenever H*W > 1 (every layer here has >= 16 pixels per channel)
    net = _net(0.0, "fp32")
    mix, voc = _data(seed=4, n=1)
    ref_loss, params = _oracle_step(net, mix, voc)
    loss = training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False)
    assert abs(float(loss[0]) - ref_loss) <= 1e-4 * ref_loss
    _check_grads(net, params, rel=1e-3)


def test_conv_wgrad_tf32_op_matches_float64():
    # svs_conv_wgrad_tf32 alone: TF32-representable operands -> products are exact, only the fp32 accumulation differs
    from svs_unet_pytorch_b200 import _lib

    def tf32(x):
        return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

    g = torch.Generator().manual_seed(1)
    for (b, gh, gw, cs, cl, scoff, lcoff) in [(3, 16, 8, 64, 32, 0, 0), (8, 8, 2, 256, 32, 0, 0), (2, 32, 16, 32, 16, 0, 16),
                                              (5, 16, 4, 160, 48, 0, 0), (2, 64, 16, 32, 16, 0, 0),
                                              (3, 16, 8, 64, 64, 0, 0), (8, 8, 2, 256, 128, 0, 0), (9, 16, 4, 128, 64, 0, 64)]:
        S = tf32(torch.randn(b, gh, gw, cs, generator=g))
        L = tf32(torch.randn(b, 2 * gh, 2 * gw, cl + lcoff, generator=g))
        Lp = torch.nn.functional.pad(L[..., lcoff:].double(), (0, 0, 2, 2, 2, 2))
        ref = torch.zeros(cs, cl, 5, 5, dtype=torch.float64)
        for kh in range(5):
            for kw in range(5):
                ref[:, :, kh, kw] = torch.einsum("byxm,byxn->mn", S.double(), Lp[:, kh:kh + 2 * gh:2, kw:kw + 2 * gw:2, :])
        got = _lib.conv_wgrad_tf32(S.cuda(), L.cuda(), l_coff=lcoff, l_c=cl).cpu().double()
        assert _rel(got, ref) <= 2e-6, (b, gh, gw, cs, cl, _rel(got, ref))


def test_stale_backward_raises():
    net = _net(0.0)
    mix, voc = _data(seed=6, n=2)
    m1 = net(mix.cuda())
    _ = net(mix.cuda())                                               # a second train-mode forward overwrites the workspace
    with pytest.raises(RuntimeError, match="most recent train-mode forward"):
        m1.sum().backward()


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_graph_replayed_step_equals_the_eager_step(precision):
    # train_step replays forward / loss / backward as one CUDA graph: same kernels, same order -> same bits; the
    # capture and its warm-up must not advance the BatchNorm buffers
    mix, voc = _data(seed=31, n=4)
    res = []
    for use_graph in (False, True):
        net = _net(0.0, precision)
        losses = [training.train_step(net, mix.cuda(), voc.cuda(), use_graph=use_graph).clone() for _ in range(3)]
        res.append((torch.stack(losses), net._flat_grad.clone(), net.conv4[0].weight.detach().clone(),
                    net.conv2[1].running_var.clone(), int(net.conv1[1].num_batches_tracked)))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])
    assert res[0][4] == res[1][4] == 3


def test_mr_stft_objective_restatement():
    # losses.py: specific_istft must agree with the hand-written iSTFT kernel on the same
    # spectrogram, the MR-STFT loss is 0 for identical waveforms, positive otherwise, and the full objective of
    # reference train.py:274-296 back-propagates through the sm_100a UNet
    from svs_unet_pytorch_b200 import losses, spectral, synth, train as svs_train
    mix, voc_wav, _ = synth.synth_song(12.0, seed=4)
    batch = spectral.SongBatch.from_audio([mix[:768 * 127]])
    mag, phase, _ = batch.stft()
    assert mag.shape[0] == 128
    ang = torch.atan2(phase[..., 1], phase[..., 0])
    m4 = mag[:, 1:].T.reshape(1, 1, 512, 128).contiguous()
    p4 = ang[:, 1:].T.reshape(1, 1, 512, 128).contiguous()
    wav = losses.specific_istft(m4, p4)
    mag0 = mag.clone()
    mag0[:, 0] = 0                                                    # specific_istft re-inserts a ZERO DC row
    ref, _ = batch.istft(mag0, phase)
    assert wav.shape == (1, 1, 768 * 127)
    assert (wav[0, 0] - ref).abs().max().item() < 2e-5
    mr = losses.MultiResolutionSTFTLoss(sample_rate=8192).cuda()
    assert float(mr(wav, wav)) == 0.0
    assert float(mr(wav, 0.5 * wav)) > 0.1
    net = _net(0.5, "tf32")
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 1, 512, 128, generator=g).cuda()
    v = x * torch.rand(2, 1, 512, 128, generator=g).cuda()
    ph = (torch.rand(2, 1, 512, 128, generator=g).cuda() - 0.5) * 6.0
    total, l1, mrv = svs_train.full_objective(net, (x, v, ph, ph), mr)
    total.backward()
    assert torch.isfinite(total) and float(mrv.detach()) > 0 and float(l1.detach()) > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    assert float(net.conv3[0].weight.grad.abs().max()) > 0
