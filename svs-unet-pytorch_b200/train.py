"""Drop-in for reference train.py (same flags, checkpoint dict, CKPT/ and LOG/ conventions) on the
sm_100a training kernels.

Differences, all outside the kernels' arithmetic and stated here so they are not silent:
* the loss is ``alpha_L1 * (L1(mask*mix, voc) + L1((1-mask)*mix, clamp(mix-voc, 0))) + alpha_MR * MRSTFT`` (train.py:275-296
  with crit = nn.L1Loss, reference config.py:33,44).  ``--mr_stft 1`` (default, the reference's objective) adds the
  MR-STFT term through ``losses.py`` (a restatement of auraloss 0.4.0 on torch.stft / torch.istft; the UNet under it is
  the sm_100a autograd path); ``--mr_stft 0`` trains on the L1 term alone with the fused CUDA-graph step
  (training.train_step — the configuration BASELINE configs[4] names and bench.py times).
* ``SpectrogramDataset`` keeps every song's spectrogram resident in HBM (frame-major, all songs back to back) and
  cuts the random 128-frame crops of a whole batch with two svs_patches_gather launches (the reference re-reads four
  .npy files per sample in 8 DataLoader workers, train.py:86-143,182); the phase files are only needed by the
  MR-STFT term.
* with WORLD_SIZE > 1 (torchrun) the step is data parallel: one NCCL all-reduce of the flat gradient
  buffer per step; rank 0 writes checkpoints and logs.  The reference is single device.
"""
from __future__ import annotations

import argparse
import os
import random

import numpy as np
import torch

from . import _lib, training
from .config import INPUT_LEN, N_BINS, SAMPLE_RATE, SAMPLES_PER_SONG
from .model import UNet

alpha_L1 = 166.66          # reference train.py:24
alpha_MR = 0.66            # reference train.py:25


class SpectrogramDataset:
    """GPU-resident mirror of reference train.py:65-143: ``<path>/mixture/*_spec.npy`` with a matching
    ``<path>/vocal/`` file; ``len = n_songs * samples_per_song``; item = random 128-frame crop (shared
    start for mixture and vocal), DC row dropped, zero padded when the song is shorter."""

    def __init__(self, path, samples_per_song=SAMPLES_PER_SONG, device="cuda", seed=None, with_phase=False):
        self.with_phase = with_phase
        self.mixture_path = os.path.join(path, "mixture")
        self.vocal_path = os.path.join(path, "vocal")
        self.samples_per_song = samples_per_song
        if not os.path.exists(self.mixture_path):
            raise FileNotFoundError(f"mixture folder not found: {self.mixture_path}")
        names = sorted(f for f in os.listdir(self.mixture_path) if f.endswith("_spec.npy"))
        self.file_names = [f for f in names if os.path.exists(os.path.join(self.vocal_path, f))]
        self.device = torch.device(device)
        # Every song's (513, T) spectrogram is kept in HBM in FRAME-MAJOR form [T][513] — the bytes of the
        # Fortran-ordered array data.py writes — all songs back to back, so a random 128-frame crop (DC row dropped,
        # zero padded when the song is shorter) is exactly one patch of svs_patches_gather and a whole batch is
        # two kernel launches instead of 2 x B Python slices and a stack.
        mix, voc, frames, mph, vph = [], [], [], [], []
        for f in self.file_names:
            m = np.load(os.path.join(self.mixture_path, f))
            v = np.load(os.path.join(self.vocal_path, f))
            t = min(m.shape[1], v.shape[1])
            mix.append(np.ascontiguousarray(m[:, :t].T, dtype=np.float32))
            voc.append(np.ascontiguousarray(v[:, :t].T, dtype=np.float32))
            frames.append(t)
            if with_phase:                                            # train.py:92-106: angle of the *_phase.npy arrays
                pn = f.replace("_spec.npy", "_phase.npy")
                mph.append(np.ascontiguousarray(np.angle(np.load(os.path.join(self.mixture_path, pn)))[:, :t].T, dtype=np.float32))
                vph.append(np.ascontiguousarray(np.angle(np.load(os.path.join(self.vocal_path, pn)))[:, :t].T, dtype=np.float32))
        self.frames = frames
        self.frame_off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
        if frames:
            self.mix_all = torch.from_numpy(np.concatenate(mix, axis=0)).to(self.device)
            self.voc_all = torch.from_numpy(np.concatenate(voc, axis=0)).to(self.device)
            if with_phase:
                self.mix_phase_all = torch.from_numpy(np.concatenate(mph, axis=0)).to(self.device)
                self.voc_phase_all = torch.from_numpy(np.concatenate(vph, axis=0)).to(self.device)
        else:
            self.mix_all = self.voc_all = torch.zeros((0, N_BINS), dtype=torch.float32, device=self.device)
        self.rng = random.Random(seed)
        print(f"[{os.path.basename(path)}] loaded {len(self.file_names)} songs, {samples_per_song} samples per "
              f"song per epoch, {len(self)} items.")

    def __len__(self):
        return len(self.file_names) * self.samples_per_song

    def crop_starts(self, indices):
        """(song, first frame, valid frames) of every item: reference train.py:112-131 (shared start for mixture and
        vocal, one randint per item that is longer than a patch)."""
        out = []
        for idx in indices:
            s = idx % len(self.file_names)
            cur = self.frames[s]
            if cur > INPUT_LEN:
                start = self.rng.randint(0, cur - INPUT_LEN)          # train.py:116
                out.append((s, start, INPUT_LEN))
            else:
                out.append((s, 0, cur))                               # zero padded on the right (train.py:124-131)
        return out

    def item(self, idx):
        """One item of shape (512, 128) per array — the reference's __getitem__ (phases only with with_phase)."""
        return tuple(t[0, 0] for t in self.crop_batch([idx]))

    def epoch_order(self, batch_size, shuffle=True, rank=0, world=1, epoch_seed=None):
        """Item indices this rank visits in one epoch.  Single process: the reference's DataLoader order
        (shuffle, last batch kept).  Data parallel: every rank shuffles with the SAME ``epoch_seed`` so that the
        strided slices ``order[rank::world]`` partition the epoch, and the order is first padded (wrapping around,
        like torch's DistributedSampler) to a multiple of ``world * batch_size`` so that all ranks run the same
        number of full steps — a rank with an extra step would block forever in the gradient all-reduce."""
        order = list(range(len(self)))
        if shuffle:
            (random.Random(epoch_seed) if epoch_seed is not None else self.rng).shuffle(order)
        if world > 1 and order:
            per_step = world * batch_size
            total = -(-len(order) // per_step) * per_step
            while len(order) < total:
                order += order[:total - len(order)]
            order = order[rank::world]
        return order

    def batches(self, batch_size, shuffle=True, rank=0, world=1, epoch_seed=None):
        order = self.epoch_order(batch_size, shuffle, rank, world, epoch_seed)
        for a in range(0, len(order), batch_size):
            yield self.crop_batch(order[a:a + batch_size])

    def crop_batch(self, indices):
        """(mix, voc) float32 (B,1,512,128) for the given items (reference train.py:100-143 per item), cut on the
        device by svs_patches_gather."""
        crops = self.crop_starts(indices)
        offs = np.asarray([(self.frame_off[s] + start) * N_BINS + 1 for s, start, _ in crops], dtype=np.int64)
        valid = np.asarray([n for _, _, n in crops], dtype=np.int32)
        d_off = torch.from_numpy(offs).to(self.device, non_blocking=True)
        d_valid = torch.from_numpy(valid).to(self.device, non_blocking=True)
        mix = _lib.patches_gather_raw(self.mix_all, d_off, d_valid, None)
        voc = _lib.patches_gather_raw(self.voc_all, d_off, d_valid, None)
        if self.with_phase:
            return (mix, voc, _lib.patches_gather_raw(self.mix_phase_all, d_off, d_valid, None),
                    _lib.patches_gather_raw(self.voc_phase_all, d_off, d_valid, None))
        return mix, voc

    def n_batches(self, batch_size, world=1):
        if world > 1:
            per_step = world * batch_size
            return -(-len(self) // per_step)
        return (len(self) + batch_size - 1) // batch_size


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--train_folder", type=str, default="./data/vocals")
    p.add_argument("--load_path", type=str, default="result.pth")
    p.add_argument("--label", type=str, required=True)
    p.add_argument("--epoch", type=int, default=2)
    p.add_argument("--batch_size", type=int, default=2)
    p.add_argument("--valid_folder", type=str, default="unet_spectrograms/valid")
    p.add_argument("--val_interval", type=int, default=20)
    p.add_argument("--mr_stft", type=int, default=1,
                   help="1: reference objective (L1 + MR-STFT, train.py:287-296); 0: L1 only, fused CUDA-graph step")
    return p


def make_checkpoint(model, epoch, scheduler=None):
    ckpt = {"epoch": epoch, "model_state_dict": model.state_dict(), "optim": model.optim.state_dict(),
            "scheduler": scheduler.state_dict() if scheduler is not None else None}
    for key in model.__dict__:                                       # train.py:377-379
        if key.startswith("loss_list"):
            ckpt[key] = getattr(model, key)
    return ckpt


def full_objective(model, batch, mr_loss_fn):
    """reference train.py:274-296 on one batch (mix, voc, mix_phase, voc_phase): returns (total, l1, mr)."""
    from . import losses
    mix, voc, mix_phase, voc_phase = batch
    mask = model(mix)                                                # train mode: autograd over the sm_100a kernels
    pred_vocal = mask * mix
    l1 = (pred_vocal - voc).abs().mean() + ((1 - mask) * mix - torch.clamp(mix - voc, min=0.0)).abs().mean()
    mr = mr_loss_fn(losses.specific_istft(pred_vocal, mix_phase), losses.specific_istft(voc, voc_phase))
    return alpha_L1 * l1 + alpha_MR * mr, l1, mr


@torch.no_grad()
def validate(model, dataset, batch_size, rank=0, world=1, mr_loss_fn=None):
    model.eval()
    total, n = 0.0, 0
    dev = next(model.parameters()).device
    for batch in dataset.batches(batch_size, shuffle=False, rank=rank, world=world):
        mix, voc = batch[0], batch[1]
        mask = model(mix)
        loss, _ = training.masked_l1(mask, mix, voc, two_term=True, want_grad=False)
        val = alpha_L1 * float(loss[0])
        if mr_loss_fn is not None:                                   # train.py:318-341
            from . import losses
            val += alpha_MR * float(mr_loss_fn(losses.specific_istft(mask * mix, batch[2]),
                                               losses.specific_istft(voc, batch[3])))
        total += val
        n += 1
    if world > 1:                                                    # every rank sees the same validation loss
        acc = torch.tensor([total, float(n)], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(acc)
        total, n = float(acc[0]), int(acc[1])
    return total / max(n, 1)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not torch.cuda.is_available():
        raise _lib.SvsError("training needs a B200: svs-unet-pytorch_b200 has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    print(f"Using device: {device} (rank {rank}/{world})")
    os.makedirs("CKPT", exist_ok=True)
    os.makedirs("LOG", exist_ok=True)
    log_file = f"LOG/log_{args.label}.txt"
    best_weight = f"CKPT/svs_best_{args.label}.pth"
    ckpt_weight = f"CKPT/svs_{args.label}.pth"

    mr = bool(args.mr_stft)
    mr_loss_fn = None
    if mr:
        from . import losses
        mr_loss_fn = losses.MultiResolutionSTFTLoss(sample_rate=SAMPLE_RATE).to(device)
    train_set = SpectrogramDataset(args.train_folder, device=device, seed=None if world == 1 else 1234 + rank,
                                   with_phase=mr)
    valid_set = None
    if os.path.exists(args.valid_folder):
        vs = SpectrogramDataset(args.valid_folder, device=device, with_phase=mr)
        valid_set = vs if len(vs) > 0 else None
    else:
        print(f"Warning: validation folder {args.valid_folder} not found, skipping validation.")

    model = UNet().to(device)
    start_epoch, scheduler = 0, None
    if os.path.exists(args.load_path):                               # train.py:216-237
        print(f"Loading checkpoint from {args.load_path}")
        ckpt = torch.load(args.load_path, map_location=device)
        model.load_state_dict(ckpt["model_state_dict"])
        if "optim" in ckpt:
            model.optim.load_state_dict(ckpt["optim"])
        start_epoch = ckpt.get("epoch", 0)
        for key in ckpt:
            if key.startswith("loss_list"):
                setattr(model, key, ckpt[key])
    if world > 1:                                                    # identical replicas
        for t in list(model.parameters()) + list(model.buffers()):
            torch.distributed.broadcast(t.data, 0)

    base_seed = torch.tensor([random.randrange(1 << 30)], dtype=torch.int64, device=device)
    if world > 1:                                                    # one shuffle seed for all ranks (crop offsets
        torch.distributed.broadcast(base_seed, 0)                    # keep their per-rank stream)
    base_seed = int(base_seed)
    best_val, log_buffer = 100.0, []
    print(f"Start training for {args.epoch - start_epoch} epochs...")
    for ep in range(start_epoch, args.epoch):
        model.train()
        if ep == 400:                                                # train.py:251-262
            for group in model.optim.param_groups:
                group["lr"] = 5e-4
            if rank == 0:
                torch.save(make_checkpoint(model, ep + 1), f"CKPT/svs_{args.label}_400.pth")
            print(f"\n[Info] Epoch {ep}: Learning rate manually changed to 5e-4!\n")
        loss_sum, n_it = 0.0, 0
        for batch in train_set.batches(args.batch_size, shuffle=True, rank=rank, world=world,
                                       epoch_seed=base_seed + ep if world > 1 else None):
            if mr:                                                   # reference objective: autograd through the kernels
                model.optim.zero_grad()
                total, _, _ = full_objective(model, batch, mr_loss_fn)
                total.backward()
                if world > 1:
                    for p_ in model.parameters():
                        torch.distributed.all_reduce(p_.grad, op=torch.distributed.ReduceOp.AVG)
                model.optim.step()
                loss_sum += float(total.detach())                    # train.py:303 (.item() per step)
            else:
                loss = training.train_step(model, batch[0], batch[1], two_term=True, loss_scale=alpha_L1)
                loss_sum += alpha_L1 * float(loss[0])
            n_it += 1
        avg = loss_sum / max(n_it, 1)
        log_buffer.append(f"{avg}\n")
        if valid_set is not None and (ep + 1) % args.val_interval == 0:
            val = validate(model, valid_set, args.batch_size, rank, world, mr_loss_fn)
            log_buffer.append(f"Val {val}\n")
            print(f"\n[Epoch {ep + 1}] Train Loss: {avg:.4e} | Val Loss: {val:.4e}")
            if val < best_val:                                       # val is all-reduced: same decision on every rank
                best_val = val
                if rank == 0:
                    model.save(best_weight)                          # train.py:353-355
            if rank == 0:
                with open(log_file, "a") as f:
                    f.writelines(log_buffer)
            log_buffer = []
        else:
            print(f"Epoch {ep + 1} Avg Loss: {avg:.4e}")
        if rank == 0:
            torch.save(make_checkpoint(model, ep + 1, scheduler), ckpt_weight)
    if log_buffer and rank == 0:
        with open(log_file, "a") as f:
            f.writelines(log_buffer)
    if world > 1:
        torch.distributed.destroy_process_group()
    print("Finish training!")


if __name__ == "__main__":
    main()
