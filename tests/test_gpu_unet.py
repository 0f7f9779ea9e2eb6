"""GPU: UNet inference kernels against the CPU oracle (pinned to the reference's model.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import randomize_bn  # noqa: E402
from oracle import unet_oracle  # noqa: E402
from svs_unet_pytorch_b200 import model as svs_model  # noqa: E402

# north_star tolerances on the mask (max-abs)
TOL = {"fp32": 1e-3, "tf32": 1e-3, "bf16": 1e-2}
ACT_NAMES = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "deconv1", "deconv2", "deconv3", "deconv4",
             "deconv5"]


def _net(seed=0, bn_seed=None):
    torch.manual_seed(seed)
    net = svs_model.UNet().eval()
    if bn_seed is not None:
        randomize_bn(net, bn_seed)
    return net


def _x(seed, n):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 1, 512, 128, generator=g)


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
def test_mask_matches_golden_reference_vectors(golden, precision):
    net = _net().cuda()
    net.precision = precision
    with torch.no_grad():
        mask = net(_x(1, 2).cuda())
    assert mask.shape == (2, 1, 512, 128) and mask.dtype == torch.float32
    err = np.abs(mask[:, 0, ::4, ::4].cpu().numpy() - golden["mask_default_sub"]).max()
    assert err <= TOL[precision], err


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
@pytest.mark.parametrize("batch", [1, 3, 8])
def test_mask_and_activations_match_oracle_random_bn(precision, batch):
    net = _net(bn_seed=2)
    x = _x(5, batch)
    with torch.no_grad():
        ref, acts = unet_oracle.unet_forward(net.state_dict(), x, return_activations=True)
    net = net.cuda()
    net.precision = precision
    with torch.no_grad():
        mask = net(x.cuda())
    assert (mask.cpu() - ref).abs().max().item() <= TOL[precision]
    plan = net.plan()
    rel = 2e-2 if precision == "bf16" else (3e-3 if precision == "tf32" else 1e-4)
    for i, n in enumerate(ACT_NAMES):
        a = plan.read_activation(i, batch).cpu()
        scale = acts[n].abs().max().item()
        assert (a - acts[n]).abs().max().item() <= rel * scale, (n, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mask_application_flags(precision):
    net = _net().cuda()
    net.precision = precision
    x = _x(7, 2).cuda()
    with torch.no_grad():
        mask = net(x)
        voc = net.separate(x, vocal_solo=True)
        acc = net.separate(x, vocal_solo=False)
    assert torch.allclose(voc, x * mask, atol=1e-6)                   # inference.py:107
    assert torch.allclose(acc, x * (1 - mask), atol=1e-6)             # inference.py:102
    assert torch.allclose(voc + acc, x, atol=1e-5)


def test_plan_tracks_parameter_updates():
    net = _net().cuda()
    net.precision = "fp32"
    x = _x(9, 1).cuda()
    with torch.no_grad():
        m1 = net(x).clone()
        net.deconv6.bias.add_(1.0)                                     # in-place update must invalidate the plan
        m2 = net(x)
    assert (m2 - m1).abs().max().item() > 0.05
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sd, x.cpu())
    assert (m2.cpu() - ref).abs().max().item() < 1e-3


def test_checkpoint_round_trip(tmp_path):
    net = _net(bn_seed=4).cuda()
    path = str(tmp_path / "svs_test.pth")
    net.save(path)                                                    # reference model.py:140-152 layout
    state = torch.load(path, map_location="cpu")
    assert set(state.keys()) >= {"model_state_dict", "optim", "loss_list_total"}
    net2 = svs_model.UNet().cuda().eval()
    net2.load_state_dict(state["model_state_dict"], strict=True)      # reference inference.py:48
    net.precision = net2.precision = "fp32"
    x = _x(11, 1).cuda()
    with torch.no_grad():
        assert torch.equal(net(x), net2(x))


def _oracle_chunked(sd, x, chunk=64):
    out = []
    with torch.no_grad():
        for a in range(0, x.shape[0], chunk):
            out.append(unet_oracle.unet_forward(sd, x[a:a + chunk]))
    return torch.cat(out)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_mask_matches_oracle_at_the_benched_batch_64_under_a_cuda_graph(precision):
    # BASELINE configs[1] exactly as bench.py times it: batch 64, torch.rand patches, CUDA-graph replay
    net = _net(bn_seed=6)
    x = _x(64, 64)
    ref = _oracle_chunked(net.state_dict(), x)
    net = net.cuda()
    net.precision = precision
    plan = net.plan()
    xd, yd = x.cuda(), torch.empty(64, 1, 512, 128, device="cuda")
    plan.forward_dense(xd, 0, yd)                                     # warm-up (allocates the workspace)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            plan.forward_dense(xd, 0, yd)
    torch.cuda.current_stream().wait_stream(side)
    yd.zero_()
    g.replay()
    torch.cuda.synchronize()
    err = (yd.cpu() - ref).abs().max().item()
    assert err <= TOL[precision], err
    g.replay()                                                        # replays are bit-identical
    torch.cuda.synchronize()
    first = yd.clone()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(first, yd)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_mask_matches_oracle_at_the_pipeline_batch_512(precision):
    # pipeline.Separator's UNet batch (max_batch = 512): 296-/148-CTA grids, wide tiles, no split-K
    net = _net(bn_seed=8)
    x = _x(512, 512)
    ref = _oracle_chunked(net.state_dict(), x)
    net = net.cuda()
    net.precision = precision
    with torch.no_grad():
        mask = net(x.cuda())
    err = (mask.cpu() - ref).abs().max().item()
    assert err <= TOL[precision], err
