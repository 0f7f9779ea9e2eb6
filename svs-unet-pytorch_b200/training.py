"""Training step on the sm_100a kernels (reference train.py:265-300).

* ``train_forward(model, mix)`` is what ``UNet.forward`` calls in ``train()`` mode: an autograd
  Function whose forward is ``svs_unet_train_forward`` (batch-statistic BatchNorm, running-stat update,
  Dropout2d through explicit keep masks) and whose backward is ``svs_unet_train_backward``; the torch
  loss code of reference train.py:275-299 works on the returned mask unchanged.
* ``train_step(model, mix, voc)`` is the fused fast path: forward, ``svs_l1_masked_loss`` (loss + dL/dmask
  in one kernel), backward, one flat-buffer gradient all-reduce over NCCL when ``torch.distributed`` is
  initialised (data parallel, reference has none), then ``model.optim.step()`` (Adam, model.py:116).
There is no CPU path."""
from __future__ import annotations

import os
from ctypes import byref

import torch

from . import _lib

TRAIN_PRECISION = "tf32"       # default arithmetic of the training step (see train_precision)

_LAYERS = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6",
           "deconv1", "deconv2", "deconv3", "deconv4", "deconv5", "deconv6"]


def _layer_modules(model, i):
    name = _LAYERS[i]
    if i < 6:
        seq = getattr(model, name)
        return seq[0], seq[1], None
    conv = getattr(model, name)
    if i == 11:
        return conv, None, None
    bad = getattr(model, name + "_BAD")
    return conv, bad[0], bad[2]


def param_list(model):
    """The 46 trainable tensors in the order the C ABI's svs_train_layer array uses."""
    out = []
    for i in range(12):
        conv, bn, _ = _layer_modules(model, i)
        out += [conv.weight, conv.bias]
        if bn is not None:
            out += [bn.weight, bn.bias]
    return out


def _workspace(model, batch):
    """Per-model training workspace (saved activations, gradients, scratch) for `batch` patches; the two most recent
    batch sizes are kept (a captured step graph is tied to its workspace address)."""
    cache = model.__dict__.setdefault("_train_ws", {})
    ws = cache.get(batch)
    if ws is None:
        nbytes = _lib.load().svs_unet_train_workspace_bytes(batch)
        dev = next(model.parameters()).device
        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        off = (-raw.data_ptr()) % 1024
        ws = raw[off:off + nbytes]
        if len(cache) >= 2:
            cache.pop(next(iter(cache)))
        cache[batch] = ws
    return ws


def train_precision(model) -> str:
    """"tf32" (default: tcgen05 kind::tf32 for forward / dgrad / wgrad, what torch + cuDNN do for the reference's fp32
    model on a GPU) or "fp32" (exact CUDA-core arithmetic).  Set ``model.train_precision`` or SVS_B200_TRAIN_PRECISION."""
    p = getattr(model, "train_precision", None) or os.environ.get("SVS_B200_TRAIN_PRECISION", TRAIN_PRECISION)
    if p not in ("tf32", "fp32"):
        raise _lib.SvsError(f"train precision must be 'tf32' or 'fp32', got {p!r}")
    return p


def _plan_handle(model):
    if train_precision(model) == "fp32":
        return None
    dev = next(model.parameters()).device
    plan = getattr(model, "_train_plan", None)
    if plan is None or plan.device != dev:
        plan = _lib.TrainPlan(dev)
        model._train_plan = plan
    return plan.handle


def _flat_grads(model):
    fg = getattr(model, "_flat_grad", None)
    n = sum(p.numel() for p in param_list(model))
    dev = next(model.parameters()).device
    if fg is None or fg.numel() != n or fg.device != dev:
        model._flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
    views, off = [], 0
    for p in param_list(model):
        views.append(model._flat_grad[off:off + p.numel()].view_as(p))
        off += p.numel()
    return model._flat_grad, views


def dropout_masks(model, batch, injected=None):
    """Per decoder block keep masks uint8 (B, C) — Dropout2d zeroes whole channels (model.py:83)."""
    masks = {}
    dev = next(model.parameters()).device
    for i in range(6, 11):
        name = _LAYERS[i]
        _, bn, drop = _layer_modules(model, i)
        if injected is not None and name in injected:
            masks[name] = injected[name].to(device=dev, dtype=torch.uint8).contiguous()
            continue
        p = float(drop.p)
        if p == 0.0:
            continue
        if abs(p - 0.5) > 1e-12:
            raise _lib.SvsError("the training kernels implement Dropout2d(p=0.5) (reference model.py:83) or p=0")
        masks[name] = (torch.rand(batch, bn.num_features, device=dev) >= 0.5).to(torch.uint8)
    return masks


def _layer_structs(model, grads, masks, need_grads=True):
    arr = (_lib.TrainLayer * 12)()
    gi = 0
    for i in range(12):
        conv, bn, _ = _layer_modules(model, i)
        arr[i].weight = conv.weight.data_ptr()
        arr[i].bias = conv.bias.data_ptr()
        if need_grads:
            arr[i].grad_weight = grads[gi].data_ptr()
            arr[i].grad_bias = grads[gi + 1].data_ptr()
        gi += 2
        if bn is not None:
            arr[i].bn_weight = bn.weight.data_ptr()
            arr[i].bn_bias = bn.bias.data_ptr()
            arr[i].bn_running_mean = bn.running_mean.data_ptr()
            arr[i].bn_running_var = bn.running_var.data_ptr()
            if need_grads:
                arr[i].grad_bn_weight = grads[gi].data_ptr()
                arr[i].grad_bn_bias = grads[gi + 1].data_ptr()
            gi += 2
        m = masks.get(_LAYERS[i]) if masks else None
        if m is not None:
            arr[i].dropout_keep = m.data_ptr()
    return arr


def _check_inputs(model, mix):
    _lib.require_cuda(mix, "mix", torch.float32)
    if mix.dim() != 4 or tuple(mix.shape[1:]) != (1, 512, 128):
        raise _lib.SvsError(f"mix must have shape (B, 1, 512, 128); got {tuple(mix.shape)}")
    for p in param_list(model):
        _lib.require_cuda(p, "parameter", torch.float32)
        if not p.is_contiguous():
            raise _lib.SvsError("parameters must be contiguous")
    _lib.check_device(mix.device)


def _raw_forward(model, mix, masks, update_running=True):
    b = mix.shape[0]
    ws = _workspace(model, b)
    arr = _layer_structs(model, None, masks, need_grads=False)
    mask = torch.empty_like(mix)
    # the workspace holds everything the backward needs (saved activations, batch statistics, packed weights):
    # stamp it so that a backward belonging to an EARLIER train-mode forward fails loudly instead of
    # differentiating the wrong graph
    model._train_gen = getattr(model, "_train_gen", 0) + 1
    with torch.cuda.device(mix.device):
        _lib.check(_lib.load().svs_unet_train_forward(_plan_handle(model), arr, mix.data_ptr(), b,
                                                      1 if update_running else 0,
                                                      mask.data_ptr(), ws.data_ptr(), ws.numel(),
                                                      _lib.stream_ptr(mix.device)), "svs_unet_train_forward")
    if update_running:
        torch._foreach_add_([_layer_modules(model, i)[1].num_batches_tracked for i in range(11)], 1)
    return mask


def _raw_backward(model, mix, grad_mask, masks, grads, first_layer=0, last_layer=11):
    b = mix.shape[0]
    ws = _workspace(model, b)
    arr = _layer_structs(model, grads, masks, need_grads=True)
    with torch.cuda.device(mix.device):
        _lib.check(_lib.load().svs_unet_train_backward_layers(_plan_handle(model), arr, mix.data_ptr(),
                                                              grad_mask.data_ptr(), b, ws.data_ptr(), ws.numel(),
                                                              first_layer, last_layer, _lib.stream_ptr(mix.device)),
                   "svs_unet_train_backward")


class _TrainForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mix, model, masks, *params):
        ctx.model, ctx.masks = model, masks
        ctx.save_for_backward(mix)
        mask = _raw_forward(model, mix, masks)
        ctx.gen = model._train_gen
        return mask

    @staticmethod
    def backward(ctx, grad_mask):
        (mix,) = ctx.saved_tensors
        model = ctx.model
        if getattr(model, "_train_gen", 0) != ctx.gen:
            raise _lib.SvsError(
                "UNet backward: another train-mode forward ran on this model after the forward being "
                "differentiated.  The saved activations live in ONE per-model workspace (not in the autograd "
                "graph), so only the most recent train-mode forward can be back-propagated: call backward() "
                "before the next forward (for gradient accumulation sum the .grad tensors, not the losses).")
        grads = [torch.empty_like(p) for p in param_list(model)]
        _raw_backward(model, mix, grad_mask.contiguous().float(), ctx.masks, grads)
        return (None, None, None) + tuple(grads)


def train_forward(model, mix, injected_masks=None):
    """``UNet.forward`` in train mode: returns the mask with an autograd edge to every parameter."""
    mix = mix.contiguous()
    _check_inputs(model, mix)
    masks = dropout_masks(model, mix.shape[0], injected_masks if injected_masks is not None
                          else getattr(model, "_injected_dropout_masks", None))
    if not torch.is_grad_enabled():
        return _raw_forward(model, mix, masks)
    return _TrainForward.apply(mix, model, masks, *param_list(model))


def masked_l1(mask, mix, voc, two_term=True, grad_scale=1.0, want_grad=True):
    """svs_l1_masked_loss: (loss tensor [3] = total / vocal / accompaniment, dL/dmask or None)."""
    loss = torch.empty(3, dtype=torch.float32, device=mask.device)
    scratch = torch.empty(_lib.L1_SCRATCH_FLOATS, dtype=torch.float32, device=mask.device)   # per call: stream safe
    grad = torch.empty_like(mask) if want_grad else None
    with torch.cuda.device(mask.device):
        _lib.check(_lib.load().svs_l1_masked_loss(mask.data_ptr(), mix.data_ptr(), voc.data_ptr(), mask.numel(),
                                                  1 if two_term else 0, float(grad_scale), loss.data_ptr(),
                                                  grad.data_ptr() if grad is not None else None,
                                                  scratch.data_ptr(), _lib.stream_ptr(mask.device)),
                   "svs_l1_masked_loss")
    return loss, grad




_SPLIT_LAYER = 6       # data parallel: gradients of layers >= 6 (the decoder, 55 % of the parameters) form the first bucket


def _step_stages(model, mix, voc, two_term, loss_scale, injected_masks, split: bool):
    """The step as a list of stage callables: [forward + loss + backward] or, for data-parallel overlap,
    [forward + loss + backward(decoder)], [backward(encoder)].  Stage 0 returns the loss tensor."""
    state = {}

    def head(last_first):
        masks = dropout_masks(model, mix.shape[0], injected_masks)
        flat, views = _flat_grads(model)
        mask = _raw_forward(model, mix, masks)
        loss, grad_mask = masked_l1(mask, mix, voc, two_term, loss_scale)
        state.update(masks=masks, views=views, grad_mask=grad_mask)
        _raw_backward(model, mix, grad_mask, masks, views, last_first, 11)
        return loss

    if not split:
        return [lambda: head(0)]
    return [lambda: head(_SPLIT_LAYER),
            lambda: _raw_backward(model, mix, state["grad_mask"], state["masks"], state["views"], 0, _SPLIT_LAYER - 1)]


def _step_body(model, mix, voc, two_term, loss_scale, injected_masks):
    """forward -> fused loss + dL/dmask -> backward into the flat gradient buffer (all on the current stream)."""
    return _step_stages(model, mix, voc, two_term, loss_scale, injected_masks, False)[0]()


class _StepGraph:
    """CUDA graph(s) of the step for a fixed batch size: the step is ~250 small launches (pack, conv, BatchNorm
    reduce / finalize / apply, dgrad, wgrad per layer), i.e. launch-bound when issued one by one from Python.
    With ``split`` the backward of the encoder is a second graph, so that the all-reduce of the decoder's
    gradients can run under it."""

    def __init__(self, model, mix, voc, two_term, loss_scale, split):
        dev = mix.device
        self.mix = torch.empty_like(mix)
        self.voc = torch.empty_like(voc)
        self.mix.copy_(mix)
        self.voc.copy_(voc)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                                 # warm-up outside capture: lazy initialisation,
            for _ in range(2):                                        # workspace / flat-buffer allocation
                _step_body(model, self.mix, self.voc, two_term, loss_scale, None)
        torch.cuda.current_stream(dev).wait_stream(side)
        stages = _step_stages(model, self.mix, self.voc, two_term, loss_scale, None, split)
        self.graphs = []
        pool = None
        for i, stage in enumerate(stages):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                out = stage()
            if i == 0:
                self.loss = out
                pool = g.pool()                                       # later stages read tensors stage 0 allocated
            self.graphs.append(g)

    def load(self, mix, voc):
        self.mix.copy_(mix)
        self.voc.copy_(voc)


def _bn_buffers(model):
    out = []
    for i in range(11):
        _, bn, _ = _layer_modules(model, i)
        out += [bn.running_mean, bn.running_var, bn.num_batches_tracked]
    return out


def _world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size()
    return 1


def _get_graph(model, mix, voc, two_term, loss_scale, split):
    key = (mix.shape[0], bool(two_term), float(loss_scale), train_precision(model), str(mix.device), bool(split),
           tuple(float(_layer_modules(model, i)[2].p) for i in range(6, 11)),
           tuple(p.data_ptr() for p in param_list(model)), _workspace(model, mix.shape[0]).data_ptr())
    cache = model.__dict__.setdefault("_step_graphs", {})
    g = cache.get(key)
    if g is None:
        saved = [b.clone() for b in _bn_buffers(model)]               # the warm-up runs real steps on the buffers
        if len(cache) >= 2:
            cache.pop(next(iter(cache)))
        g = _StepGraph(model, mix, voc, two_term, loss_scale, split)
        for b, v in zip(_bn_buffers(model), saved):
            b.copy_(v)
        cache[key] = g
        model._train_gen = getattr(model, "_train_gen", 0) + 1
    return g


def _all_reduce_mean(t, world):
    """Asynchronous mean over ranks of a slice of the flat gradient buffer (NCCL: averaged inside the collective)."""
    if torch.distributed.get_backend() == "nccl":
        return [torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.AVG, async_op=True)], None
    return [torch.distributed.all_reduce(t, async_op=True)], (t, world)


def _split_offset(model):
    """First element of the decoder's gradients in the flat buffer (parameters are ordered conv1 .. deconv6)."""
    n = 0
    for i in range(_SPLIT_LAYER):
        conv, bn, _ = _layer_modules(model, i)
        n += conv.weight.numel() + conv.bias.numel() + bn.weight.numel() + bn.bias.numel()
    return n


def train_step(model, mix, voc, two_term: bool = True, loss_scale: float = 1.0, step: bool = True,
               injected_masks=None, sync_grads: bool = True, use_graph: bool | None = None):
    """One fused optimisation step (reference train.py:271-300 without the MR-STFT term).

    Returns the device tensor [total, vocal, accompaniment] of the UNSCALED L1 loss.  Gradients are
    written into one flat fp32 buffer (``model._flat_grad``, 9,823,313 floats); with an initialised
    ``torch.distributed`` process group they are averaged across ranks over NCCL.  ``use_graph`` (default: on unless
    dropout masks are injected or SVS_B200_TRAIN_GRAPH=0) replays the forward / loss / backward as one CUDA graph."""
    mix = mix.contiguous()
    voc = voc.contiguous()
    _check_inputs(model, mix)
    _lib.require_cuda(voc, "voc", torch.float32)
    if use_graph is None:
        use_graph = injected_masks is None and os.environ.get("SVS_B200_TRAIN_GRAPH", "1") != "0"
    world = _world() if sync_grads else 1
    with torch.no_grad():
        flat, views = _flat_grads(model)
        if world > 1:
            # two stages: the decoder's gradients (the tail of the flat buffer) are all-reduced on NCCL's stream while
            # the encoder's backward runs; only the encoder's bucket is exposed
            off = _split_offset(model)
            if use_graph and injected_masks is None:
                g = _get_graph(model, mix, voc, two_term, loss_scale, True)
                g.load(mix, voc)
                stages = [gr.replay for gr in g.graphs]
                loss = g.loss
            else:
                stages = _step_stages(model, mix, voc, two_term, loss_scale, injected_masks, True)
                loss = None
            out = stages[0]()
            loss = out if loss is None else loss
            works, fix = [], []
            for lo, hi, stage in ((off, flat.numel(), stages[1]), (0, off, None)):
                w, f = _all_reduce_mean(flat[lo:hi], world)
                works += w
                if f is not None:
                    fix.append(f)
                if stage is not None:
                    stage()
            for w in works:
                w.wait()
            for t, n in fix:
                t.div_(n)
        elif use_graph and injected_masks is None:
            g = _get_graph(model, mix, voc, two_term, loss_scale, False)
            g.load(mix, voc)
            g.graphs[0].replay()
            loss = g.loss
        else:
            loss = _step_body(model, mix, voc, two_term, loss_scale, injected_masks)
        for p, g_ in zip(param_list(model), views):
            p.grad = g_
        if step:
            model.optim.step()
    return loss
