"""Generates tests/golden/unet_golden.npz by running the REFERENCE's own model.py (torch CPU fp32).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

`import model` fails at reference model.py:6 (`import auraloss`, unused in that file), so a stub
module is placed in sys.modules first.  Everything stored is small: seeds, per-tensor checksums of
the random-init state_dict, sub-sampled masks and per-layer activation statistics.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SVS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unet_golden.npz")


def import_reference_model():
    sys.modules.setdefault("auraloss", types.ModuleType("auraloss"))
    sys.path.insert(0, REF)
    try:
        import model as ref_model  # noqa
    finally:
        sys.path.pop(0)
    return ref_model


def randomize_bn(net, seed):
    """Moderate BatchNorm statistics (SURVEY.md section 4 caution ii) so the fold is exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            c = m.num_features
            m.running_var.copy_(torch.rand(c, generator=g) * 1.5 + 0.5)
            m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
            m.weight.data.copy_(torch.rand(c, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(c, generator=g) * 0.1)


def main():
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    ref_model = import_reference_model()
    out = {}
    torch.manual_seed(0)
    net = ref_model.UNet().eval()
    sd = net.state_dict()
    keys = list(sd.keys())
    out["sd_keys"] = np.array(keys)
    out["sd_sum"] = np.array([float(sd[k].double().sum()) for k in keys])
    out["sd_abs_sum"] = np.array([float(sd[k].double().abs().sum()) for k in keys])
    out["n_params"] = np.array(sum(p.numel() for p in net.parameters()))

    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 1, 512, 128, generator=g)
    out["x_sha1"] = np.array(hashlib.sha1(x.numpy().tobytes()).hexdigest())

    acts = {}
    hooks = []
    names = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "deconv1_BAD", "deconv2_BAD",
             "deconv3_BAD", "deconv4_BAD", "deconv5_BAD"]
    for n in names:
        hooks.append(getattr(net, n).register_forward_hook(lambda m, i, o, n=n: acts.__setitem__(n, o.detach().clone())))
    with torch.no_grad():
        mask = net(x)
    out["mask_default_sub"] = mask[:, 0, ::4, ::4].numpy()
    out["mask_default_sum"] = np.array(float(mask.double().sum()))
    out["mask_default_sqsum"] = np.array(float((mask.double() ** 2).sum()))
    out["act_names"] = np.array(names)
    out["act_default_mean"] = np.array([float(acts[n].double().mean()) for n in names])
    out["act_default_absmean"] = np.array([float(acts[n].double().abs().mean()) for n in names])

    randomize_bn(net, seed=2)
    with torch.no_grad():
        mask_bn = net(x)
        mask_bn64 = net.double()(x.double())
    net.float()
    out["mask_bn_sub"] = mask_bn[:, 0, ::4, ::4].numpy()
    out["mask_bn_sum"] = np.array(float(mask_bn.double().sum()))
    out["mask_bn_fp32_vs_fp64"] = np.array(float((mask_bn.double() - mask_bn64).abs().max()))
    out["act_bn_mean"] = np.array([float(acts[n].double().mean()) for n in names])
    out["act_bn_absmean"] = np.array([float(acts[n].double().abs().mean()) for n in names])
    for h in hooks:
        h.remove()

    # training-mode step (config 5 shape at batch 2): BN batch statistics, dropout disabled
    torch.manual_seed(0)
    net = ref_model.UNet().train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    g = torch.Generator().manual_seed(3)
    mix = torch.rand(2, 1, 512, 128, generator=g)
    voc = mix * torch.rand(2, 1, 512, 128, generator=g)
    mask = net(mix)
    crit = torch.nn.L1Loss()                                   # reference config.py:33,44
    loss_v = crit(mask * mix, voc)                             # reference train.py:274-283
    loss_a = crit((1 - mask) * mix, torch.clamp(mix - voc, min=0.0))
    loss = loss_v + loss_a
    loss.backward()
    pnames = [n for n, _ in net.named_parameters()]
    out["train_loss"] = np.array([float(loss), float(loss_v), float(loss_a)])
    out["train_param_names"] = np.array(pnames)
    out["train_grad_l2"] = np.array([float(p.grad.double().norm()) for _, p in net.named_parameters()])
    out["train_grad_sum"] = np.array([float(p.grad.double().sum()) for _, p in net.named_parameters()])
    bnames = [n for n, b in net.named_buffers() if "running" in n]
    out["train_buffer_names"] = np.array(bnames)
    out["train_buffer_sum"] = np.array([float(dict(net.named_buffers())[n].double().sum()) for n in bnames])
    out["train_mask_sub"] = mask.detach()[:, 0, ::4, ::4].numpy()

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
