"""GPU diagnostic (not a test): per-parameter gradient error of training.train_step against the float64 oracle for
both arithmetic modes, plus the standalone tcgen05 weight-gradient op.  Usage: python tests/diag_train.py [batch]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle  # noqa: E402
from svs_unet_pytorch_b200 import _lib, model as svs_model, training  # noqa: E402


def tf32_round(x):
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if x.dtype == torch.float32 else x


def wgrad_ref(S, L):
    b, gh, gw, _ = S.shape
    Lp = torch.nn.functional.pad(L.double(), (0, 0, 2, 2, 2, 2))
    out = torch.zeros(S.shape[3], L.shape[3], 5, 5, dtype=torch.float64)
    for kh in range(5):
        for kw in range(5):
            win = Lp[:, kh:kh + 2 * gh:2, kw:kw + 2 * gw:2, :]
            out[:, :, kh, kw] = torch.einsum("byxm,byxn->mn", S.double(), win)
    return out


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    torch.manual_seed(1)
    for (b, gh, gw, cs, cl) in [(3, 16, 8, 64, 32), (8, 8, 2, 256, 32), (2, 32, 16, 32, 16), (5, 16, 4, 160, 48)]:
        S = tf32_round(torch.randn(b, gh, gw, cs))
        L = tf32_round(torch.randn(b, 2 * gh, 2 * gw, cl))
        ref = wgrad_ref(S, L)
        got = _lib.conv_wgrad_tf32(S.cuda(), L.cuda()).cpu().double()
        err = (got - ref).norm() / ref.norm()
        print(f"wgrad_tc B{b} {gh}x{gw} Cs{cs} Cl{cl}: rel-L2 {err:.3e}  max|ref| {ref.abs().max():.3f} max|err| {(got-ref).abs().max():.3e}")
    g = torch.Generator().manual_seed(3)
    mix = torch.rand(batch, 1, 512, 128, generator=g)
    voc = mix * torch.rand(batch, 1, 512, 128, generator=g)
    torch.manual_seed(0)
    net = svs_model.UNet().train().cuda()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    sd_f32 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sd = {k: (v.detach().cpu().double() if v.is_floating_point() else v.detach().cpu().clone()) for k, v in net.state_dict().items()}
    params = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    ref_mask = unet_oracle.unet_forward(params, mix.double(), training=True)
    ref_loss = unet_oracle.l1_masked_loss(ref_mask, mix.double(), voc.double(), two_term=True)
    ref_loss.backward()
    for prec in ("fp32", "tf32"):
        net.train_precision = prec
        loss = training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            training.train_step(net, mix.cuda(), voc.cuda(), two_term=True, step=False)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        print(f"--- {prec}: loss {float(loss[0]):.8f} ref {float(ref_loss):.8f} rel {abs(float(loss[0]) - float(ref_loss)) / float(ref_loss):.2e}   {ms:.2f} ms/step (batch {batch})")
        worst = 0.0
        for name, p in net.named_parameters():
            ref = params[name].grad.float()
            got = p.grad.cpu()
            rel = float((got - ref).norm()) / max(float(ref.norm()), 1e-12)
            worst = max(worst, rel if float(ref.norm()) > 1e-9 else 0.0)
            print(f"   {name:28s} |ref| {float(ref.norm()):.3e}  rel-L2 {rel:.3e}")
        print(f"   worst rel-L2 {worst:.3e}")
    # what the reference itself does on a GPU: the same torch modules through torch eager + cuDNN, TF32 on / off
    import bench
    for allow in (True, False):
        torch.backends.cudnn.allow_tf32 = allow
        torch.backends.cuda.matmul.allow_tf32 = allow
        net.zero_grad(set_to_none=True)
        tnet = svs_model.UNet().train().cuda()
        tnet.load_state_dict({k: v for k, v in sd_f32.items()})
        for m in tnet.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p = 0.0
        x, v = mix.cuda(), voc.cuda()
        mask = bench.torch_eager_forward(tnet, x)
        loss = (mask * x - v).abs().mean() + ((1 - mask) * x - torch.clamp(x - v, min=0)).abs().mean()
        loss.backward()
        worst = 0.0
        rows = []
        for name, p in tnet.named_parameters():
            ref = params[name].grad.float()
            rel = float((p.grad.cpu() - ref).norm()) / max(float(ref.norm()), 1e-12)
            if float(ref.norm()) > 1e-9:
                worst = max(worst, rel)
            if name.endswith("weight") and "BAD" not in name and ".1." not in name:
                rows.append(f"{name.split('.')[0]} {rel:.2e}")
        print(f"--- torch eager + cuDNN allow_tf32={allow}: loss rel {abs(float(loss) - float(ref_loss)) / float(ref_loss):.2e}  worst rel-L2 {worst:.3e}")
        print("    " + "  ".join(rows))


if __name__ == "__main__":
    main()
