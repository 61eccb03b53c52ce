"""Input pipeline of the training hot loop: uint8 HWC frames (as MineRL delivers them, vae_utility.py:324-343) live in
pinned host memory, a copy stream moves batch i + 1 to the device while step i runs, and the step graph converts
uint8 -> fp32 NCHW on the device (cvae_frames_u8_to_f32).  12 KB per frame cross PCIe instead of 48 KB.

    stager = FrameStager(step)                       # step: cvae_native.trainer.TrainStep
    for losses in stager.run(batches):               # batches: iterable of pinned uint8 (B, 64, 64, 3) tensors
        ...                                          # losses: pinned host tensor [total, recon, KLD] of that step

torch supplies streams, events and pinned memory; the arithmetic is in libcvae.so.
"""
from __future__ import annotations

import torch


def pin_frames_u8(frames):
    """uint8 (N, 64, 64, 3) array / tensor -> pinned host tensor (no copy when it already is one)."""
    t = torch.as_tensor(frames)
    if t.dtype != torch.uint8 or t.dim() != 4 or tuple(t.shape[1:]) != (64, 64, 3):
        raise ValueError(f"expected uint8 frames of shape (N, 64, 64, 3), got {t.dtype} {tuple(t.shape)}")
    t = t.contiguous()
    return t if t.is_pinned() else t.pin_memory()


class FrameStager:
    """Double-buffered host -> device staging around a TrainStep.  Two device staging buffers, two pinned loss
    buffers; the copy of batch i + 1 is issued before step i is launched and waits (on the copy stream) only for
    the step that last read its staging buffer."""

    def __init__(self, step, eps_generator=None):
        self.step = step
        dev = step.x.device
        B = step.B
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [torch.empty(B, 64, 64, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.loss_host = [torch.empty(3, pin_memory=True) for _ in range(2)]
        self.loss_done = [torch.cuda.Event() for _ in range(2)]
        self.gen = eps_generator
        self.h2d_bytes_per_step = B * 64 * 64 * 3
        self.d2h_bytes_per_step = 12

    def _issue_copy(self, i, host_batch):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i % 2])
            self.stage[i % 2].copy_(host_batch, non_blocking=True)
            self.ready[i % 2].record(self.copy_stream)

    def run(self, batches, eps=None):
        """Generator over the steps: yields the pinned host losses of step i - 1 while step i is in flight (and the
        last step's after the loop), so the host never stalls the device.  `eps`: optional iterable of (B, 32) noise
        tensors (host or device) for parity runs; default: drawn on the device."""
        step, cur = self.step, torch.cuda.current_stream()
        it = iter(batches)
        eps_it = iter(eps) if eps is not None else None
        nxt = next(it, None)
        if nxt is None:
            return
        for b in range(2):
            self.consumed[b].record(cur)
        self._issue_copy(0, nxt)
        i = 0
        while nxt is not None:
            following = next(it, None)
            if following is not None:
                self._issue_copy(i + 1, following)
            cur.wait_event(self.ready[i % 2])
            step.load(frames_u8=self.stage[i % 2])
            self.consumed[i % 2].record(cur)
            if eps_it is not None:
                step.load(eps=next(eps_it))
            elif self.gen is not None:
                step.eps.normal_(generator=self.gen)
            else:
                step.eps.normal_()
            out = step.run(from_u8=True)
            self.loss_host[i % 2].copy_(out, non_blocking=True)
            self.loss_done[i % 2].record(cur)
            if i > 0:
                self.loss_done[(i - 1) % 2].synchronize()
                yield self.loss_host[(i - 1) % 2]
            nxt = following
            i += 1
        self.loss_done[(i - 1) % 2].synchronize()
        yield self.loss_host[(i - 1) % 2]
