"""GPU: every kernel variant stays parity-green — the TMA-im2col kernels that the zero-copy kernels replaced,
the un-merged transposed convolutions, split-K on/off / across a cluster / through HBM, 128x256 tiles at large
batch, both tensor-core conv1 kernels, the CUDA-core edge layers, and launches without PDL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.environ["SVS_ROOT"])
from oracle import unet_oracle
from svs_unet_pytorch_b200 import model as svs_model
torch.manual_seed(0)
net = svs_model.UNet().eval()
g = torch.Generator().manual_seed(1)
x = torch.rand(int(os.environ.get("SVS_TEST_BATCH", "5")), 1, 512, 128, generator=g)
with torch.no_grad():
    ref = unet_oracle.unet_forward(net.state_dict(), x)
net = net.cuda()
for prec, tol in (("bf16", 1e-2), ("tf32", 1e-3)):
    net.precision = prec
    with torch.no_grad():
        err = (net(x.cuda()).cpu() - ref).abs().max().item()
    print(prec, err)
    assert err <= tol, (prec, err)
print("ok")
"""

VARIANTS = {
    "default": {},
    "im2col_kernels": {"SVS_ZC_DISABLE": "1"},
    "unmerged_deconv": {"SVS_ZC_DISABLE": "1", "SVS_TC_NO_MERGE": "1"},
    "no_split_k": {"SVS_TC_SPLITK": "1"},
    "split_k_4": {"SVS_TC_SPLITK": "4"},
    "cuda_core_edges": {"SVS_TC_DISABLE_MASK": str((1 << 0) | (1 << 1) | (1 << 10) | (1 << 11))},
    "no_pdl": {"SVS_NO_PDL": "1"},
    "split_k_through_hbm": {"SVS_TC_CLUSTER": "0"},                     # partial buffer + reduction kernel
    "split_k_2_cluster": {"SVS_TC_SPLITK": "2"},
    "conv1_im2col": {"SVS_C1Z_DISABLE": "1"},                           # conv1_tc_kernel instead of conv1_zc_kernel
    "wide_tiles_large_batch": {"SVS_TEST_BATCH": "160"},                # conv5 / conv6 / deconv1 on 128 x 256 tiles
    "narrow_tiles_large_batch": {"SVS_TEST_BATCH": "160", "SVS_TC_NO_WIDE": "1"},
    "full_width_slab_rows": {"SVS_ZC_NARROW": "0"},
    "phase_trimmed_merged_deconv": {"SVS_ZC_TRIM": "1"},                # deconv3 / deconv4 taps only over the phase blocks they reach
    "dual_tiles_per_weight_pass": {"SVS_ZC_DUAL": "1", "SVS_TEST_BATCH": "160"},   # conv4 / deconv4: two M tiles per CTA step
    "wide_cluster_tiles": {"SVS_CK_WIDE": "1"},                         # conv5 / conv6 / deconv1 on 128 x 256 cluster tiles
    "per_thread_store_epilogue": {"SVS_ZC_TMASTORE": "0"},              # zc layers without the TMA-store epilogue
    "weight_multicast_pairs": {"SVS_ZC_MCAST": "1"},                    # CTA pairs share streamed weight chunks (TMA multicast)                     # conv2 / conv3 on 128-byte rows (both concat halves)
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_kernel_variant_matches_oracle(name, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, SVS_ROOT=ROOT, **VARIANTS[name])
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("flag", ["SVS_WGRAD_N64=1", "SVS_WGRAD_GROUPED=0"])
def test_wgrad_variant_matches_float64(tmp_path, flag):
    # SVS_WGRAD_N64=1: the four-pass N = 64 form of wgrad_tc_kernel; SVS_WGRAD_GROUPED=0: its two-pass N = 32 form (one
    # instruction per tap) instead of the default grouped-tap kernel (selected once per process)
    script = tmp_path / "w.py"
    script.write_text("""
import os, sys, torch
sys.path.insert(0, os.environ["SVS_ROOT"])
from svs_unet_pytorch_b200 import _lib
def tf32(x): return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
g = torch.Generator().manual_seed(1)
for (b, gh, gw, cs, cl) in [(3, 16, 8, 64, 64), (8, 8, 2, 256, 128), (2, 32, 16, 32, 16)]:
    S = tf32(torch.randn(b, gh, gw, cs, generator=g)); L = tf32(torch.randn(b, 2 * gh, 2 * gw, cl, generator=g))
    Lp = torch.nn.functional.pad(L.double(), (0, 0, 2, 2, 2, 2))
    ref = torch.zeros(cs, cl, 5, 5, dtype=torch.float64)
    for kh in range(5):
        for kw in range(5):
            ref[:, :, kh, kw] = torch.einsum("byxm,byxn->mn", S.double(), Lp[:, kh:kh + 2 * gh:2, kw:kw + 2 * gw:2, :])
    got = _lib.conv_wgrad_tf32(S.cuda(), L.cuda()).cpu().double()
    err = float((got - ref).norm() / ref.norm())
    assert err <= 2e-6, err
print("ok")
""")
    env = dict(os.environ, SVS_ROOT=ROOT, **dict([flag.split("=")]))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("flag", ["SVS_TRAIN_ZC=0", "SVS_TRAIN_FORK=0", "SVS_WGRAD_GROUPED=0"])
def test_training_variants_match_reference_goldens(flag):
    # the training step's opt-out switches (im2col kernels for the forward of conv2-4 / deconv3-5, weight gradients on
    # the caller's stream, one-instruction-per-tap wgrad) run the reference-golden tests of test_gpu_train.py unchanged
    env = dict(os.environ, **dict([flag.split("=")]))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_train.py"), "-x", "-q",
                        "-k", "golden or bit_reproducible or batch_64 or equals_the_eager"], capture_output=True, text=True, env=env,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
