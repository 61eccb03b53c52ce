"""Training-step driver for the hot loop of the reference's train() (vae.py:40-58):

    preds = critic.evaluate(images); opt.zero_grad(); out = autoencoder(images, preds)
    losses = autoencoder.vae_loss(*out); losses['total_loss'].backward(); opt.step()

as ONE replayable CUDA graph per batch size (about 75 kernel launches, no host work in between).
Data-parallel training (one process per GPU, PyTorch-DDP semantics: per-rank loss/grads on the local
shard, Adam on the mean gradient; BatchNorm statistics stay per-rank) all-reduces the
gradient over NCCL in two buckets between a forward/backward graph and an Adam graph; the first bucket
(decoder + heads) starts beside the encoder's backward pass, triggered by an external event node inside
the graph.

torch supplies device memory, streams, graph capture and the process group; all arithmetic is in
libcvae.so.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import binding as L
from .engine import VAEEngine, KLD_WEIGHT


def shard_batch(idx, rank, world):
    """This rank's contiguous slice of one global batch of sample indices.  Every rank gets the SAME number of
    samples (the gradient all-reduce weights ranks equally and every rank has to reach it): up to world-1 samples of
    a batch that does not divide evenly are dropped, and a batch smaller than the world yields an empty slice on
    EVERY rank (callers skip it together)."""
    per = len(idx) // world
    return idx[rank * per:(rank + 1) * per]


class TrainStep:
    """Static-buffer training step.  Feed inputs with `load_*`, run with `run()`."""

    def __init__(self, vae, critic, batch, lr=5e-5, use_graph=True, process_group=None):
        self.vae, self.critic, self.B, self.lr = vae, critic, int(batch), float(lr)
        self.eng: VAEEngine = vae._bind()
        dev = self.eng.device
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        if self.world > 1:
            # every rank must start from rank 0's parameters, optimizer state and BatchNorm buffers: identical seeding is
            # not a contract (a checkpoint loaded on one rank, a different init order) and a mismatch diverges silently
            eng = self.eng
            if eng.exp_avg is None:
                eng.exp_avg, eng.exp_avg_sq = torch.zeros_like(eng.flat), torch.zeros_like(eng.flat)
            for t in [eng.flat, eng.exp_avg, eng.exp_avg_sq, eng.step] + list(eng.running_mean) + list(eng.running_var) + list(eng.nbt):
                torch.distributed.broadcast(t, src=torch.distributed.get_global_rank(process_group, 0), group=process_group)
        self.x = torch.zeros(self.B, 3, 64, 64, device=dev)
        self.x_u8 = torch.zeros(self.B, 64, 64, 3, dtype=torch.uint8, device=dev)
        self.eps = torch.zeros(self.B, 32, device=dev)
        self.pred = torch.zeros(self.B, device=dev)
        self.ws = self.eng.workspace(self.B, True)
        self.critic_w = critic._weights() if critic is not None else None
        self.losses = self.ws.losses
        self._graphs = None
        if os.environ.get("CVAE_NO_SIDE_STREAM") is None:
            self.eng.side_stream = torch.cuda.Stream()
            self.eng.fold_stream = torch.cuda.Stream()
        self._use_graph = use_graph
        self.launches_per_step = None
        # data parallel: all-reduce the early gradient bucket beside the encoder's backward pass (CVAE_DP_OVERLAP=0: one
        # all-reduce of the whole gradient after the pass, the round-1 scheme)
        self.overlap = self.world > 1 and self.eng.side_stream is not None and os.environ.get("CVAE_DP_OVERLAP", "1") != "0"
        if self.overlap:
            self.comm_stream = torch.cuda.Stream()
            self.early_event = torch.cuda.Event(external=True)
            # CVAE_DP_BUCKETS=3: encoder convs 3 and 2 as a bucket of their own (second external event).  Measured slower than
            # two buckets at 2 and at 8 GPUs (1.427 vs 1.417 ms, 1.459 vs 1.453 ms): another NCCL kernel in the middle of the
            # backward pass costs more than the smaller last all-reduce saves.  Off by default.
            self.mid_event = torch.cuda.Event(external=True) if os.environ.get("CVAE_DP_BUCKETS", "2") == "3" else None
        # CVAE_COMM=native: the all-reduce goes through libcvae's own NCCL communicator (cvae_comm_*, csrc/comm.cu), the path a
        # non-Python host would use; torch.distributed then only carries the 128 rendezvous bytes.  Default: torch.distributed.
        self.native_comm = self.world > 1 and os.environ.get("CVAE_COMM") == "native"
        if self.native_comm and L.lib.cvae_comm_world() == 0:
            rank = torch.distributed.get_rank(process_group)
            ident = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (ctypes.c_char * 128)()
                L.check(L.lib.cvae_comm_unique_id(buf))
                ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            ident = ident.to(dev)
            torch.distributed.broadcast(ident, src=torch.distributed.get_global_rank(process_group, 0), group=process_group)
            raw = bytes(ident.cpu().numpy().tobytes())
            L.check(L.lib.cvae_comm_init(rank, self.world, raw))

    # ---- the work ---------------------------------------------------------------------------------
    def _front(self, from_u8, stage="all"):
        eng, ws, s = self.eng, self.ws, L.stream_ptr()
        if from_u8:
            L.check(L.lib.cvae_frames_u8_to_f32(self.B, self.x_u8.data_ptr(), self.x.data_ptr(), s))
        side, joined = eng.side_stream, None
        # CVAE_PACK_FIRST=1: the side stream packs the forward operand forms first (encoder conv 1 waits for them ~80 us into
        # the step) and runs the critic after that (its value is needed ~250 us in).  Measured: device-timed step 1.343 ->
        # 1.338 ms, but the end-to-end figure through the frame stager 192.9 k -> 190.3 k frames/s (profiles/r02_pack_first.log):
        # the end-to-end number is the headline, so the critic stays in front by default.
        packed_here = side is not None and os.environ.get("CVAE_PACK_FIRST") == "1"
        if packed_here:
            eng.pack()
        if self.critic_w is not None:
            if side is None:
                L.check(L.lib.cvae_critic_fwd(self.B, self.x.data_ptr(), self.critic_w.data_ptr(), self.pred.data_ptr(), s))
            else:   # the critic value is only needed by the decoder: score the frames beside the encoder
                fork = torch.cuda.Event()
                fork.record()
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    L.check(L.lib.cvae_critic_fwd(self.B, self.x.data_ptr(), self.critic_w.data_ptr(), self.pred.data_ptr(),
                                                  L.stream_ptr()))
                    joined = torch.cuda.Event()
                    joined.record()
        fused = eng.fused_bottleneck
        eng.encode(self.x, True, ws, pack=not packed_here, fc=not fused)
        if joined is not None:
            torch.cuda.current_stream().wait_event(joined)
        if fused:
            eng.bottleneck_forward(self.pred, self.eps, ws)
        eng.decode(self.pred, self.eps, True, ws, head=not fused)
        # the KL term rides along with the latent kernels: partial sums in the forward (separate kernels only; the loss kernel
        # reduces mu / logvar itself behind the fused bottleneck), its gradient in the backward
        # CVAE_LOSS_SPLIT=1: loss scalars (one block) on the side stream, the backward kernel derives its coefficients from the
        # level sums.  Measured slower (1.362 vs 1.355 ms): the fork costs more than the 8 us kernel it takes off the chain.
        split = os.environ.get("CVAE_LOSS_SPLIT") == "1"
        eng.loss_forward(ws.recon, self.x, ws.ml, ws, fused_kld=not fused, split=split)
        eng.loss_backward(ws.recon, self.x, ws.ml, ws, fused_kld=True, split=split)
        eng.early_event = self.early_event if self.overlap else None      # (the engine is shared between TrainSteps)
        eng.mid_event = self.mid_event if self.overlap else None
        # CVAE_EARLY_ADAM=1 (one GPU): Adam for everything but the first conv block runs beside that block's weight gradient,
        # the last kernel of the backward pass.  Measured: 1.397 ms against 1.388 ms with one Adam launch behind it (the
        # update competes with the GEMM it was meant to hide behind), so it is off by default.
        self._adam_split = self.world == 1 and stage == "all" and eng.side_stream is not None and os.environ.get("CVAE_EARLY_ADAM") == "1"
        eng.early_adam = dict(lr=self.lr, grad_scale=1.0) if self._adam_split else None
        eng.backward(self.x, self.eps, ws, ws.d_recon, None, None, kld_grad_scale=KLD_WEIGHT / self.B, stage=stage)
        eng.early_event = None
        eng.mid_event = None
        eng.early_adam = None

    def _front_encoder(self):
        self.eng.backward(self.x, self.eps, self.ws, self.ws.d_recon, None, None, stage="encoder")

    def _back(self):
        if getattr(self, "_adam_split", False):
            self.eng.adam_step(self.lr, grad_scale=1.0, hi=self.eng.first_block_end())       # the rest went beside the backward pass
        else:
            self.eng.adam_step(self.lr, grad_scale=1.0 / self.world)

    def _buckets(self):
        """(early, mid, late) views of the flat gradient in the order they become final: heads + decoder, encoder convs 3 and 2,
        encoder convs 1 and 0 (VAEEngine.early_bucket_offset / mid_bucket_offset); mid is None with two buckets."""
        off = self.eng.early_bucket_offset()
        if self.overlap and self.mid_event is not None:
            mid = self.eng.mid_bucket_offset()
            return self.eng.gflat[off:], self.eng.gflat[mid:off], self.eng.gflat[:mid]
        return self.eng.gflat[off:], None, self.eng.gflat[:off]

    def _allreduce(self):
        """Sum the flat gradient over the ranks, after _front has been launched on the current stream: the early bucket
        on the communication stream as soon as the event inside the backward pass fires (beside the encoder's backward
        pass), the late bucket behind the whole pass; the current stream then waits for both."""
        early, mid, late = self._buckets()
        if self.native_comm:
            reduce_ = lambda t: L.check(L.lib.cvae_comm_allreduce_sum(t.data_ptr(), t.numel(), L.stream_ptr()))
            if self.overlap:
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(self.early_event)
                    reduce_(early)
                    if mid is not None:
                        self.comm_stream.wait_event(self.mid_event)
                        reduce_(mid)
                    done = torch.cuda.Event()
                    done.record()
                reduce_(late)
                torch.cuda.current_stream().wait_event(done)
            else:
                reduce_(self.eng.gflat)
        elif self.overlap:
            works = []
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(self.early_event)
                works.append(torch.distributed.all_reduce(early, group=self.pg, async_op=True))
                if mid is not None:
                    self.comm_stream.wait_event(self.mid_event)
                    works.append(torch.distributed.all_reduce(mid, group=self.pg, async_op=True))
            works.append(torch.distributed.all_reduce(late, group=self.pg, async_op=True))
            for w in works:
                w.wait()
        else:
            torch.distributed.all_reduce(self.eng.gflat, group=self.pg)

    def _eager(self, from_u8):
        self._front(from_u8)
        if self.world > 1:
            self._allreduce()
        self._back()

    def _capture(self, from_u8):
        # warm-up on a side stream (allocations, cudaFuncSetAttribute, Adam state) -- results are
        # discarded by restoring parameters, optimizer state and BN buffers afterwards
        eng = self.eng
        keep = [eng.flat.clone(), eng.step.clone()] + [t.clone() for t in eng.running_mean + eng.running_var] + \
               [t.clone() for t in eng.nbt]
        had_state = eng.exp_avg is not None
        if had_state:
            keep_m, keep_v = eng.exp_avg.clone(), eng.exp_avg_sq.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            n0 = L.lib.cvae_launch_count()
            self._eager(from_u8)
            self.launches_per_step = int(L.lib.cvae_launch_count() - n0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        eng.check_fault()
        # Single GPU: one graph for the whole step.  Data parallel: a forward/backward graph and an Adam graph with the
        # NCCL all-reduces issued from the host between them (capturing the collectives inside the graph hung an 8-rank
        # run on this pool in round 1).  The forward/backward graph carries an EXTERNAL event record node at the point
        # where the early gradient bucket (decoder + heads, 58 % of the floats) is final, so its all-reduce starts on the
        # communication stream while the graph is still walking the encoder's backward pass.
        if self.world == 1:
            graphs = [torch.cuda.CUDAGraph()]
            with torch.cuda.graph(graphs[0]):
                self._eager(from_u8)
        else:
            graphs = [torch.cuda.CUDAGraph() for _ in range(2)]
            with torch.cuda.graph(graphs[0]):
                self._front(from_u8)
            with torch.cuda.graph(graphs[1]):
                self._back()
        # undo the warm-up step
        eng.flat.copy_(keep[0]); eng.step.copy_(keep[1])
        n = len(eng.running_mean)
        for i in range(n):
            eng.running_mean[i].copy_(keep[2 + i]); eng.running_var[i].copy_(keep[2 + n + i]); eng.nbt[i].copy_(keep[2 + 2 * n + i])
        if had_state:
            eng.exp_avg.copy_(keep_m); eng.exp_avg_sq.copy_(keep_v)
        else:
            eng.exp_avg.zero_(); eng.exp_avg_sq.zero_()
        return graphs

    # ---- public -----------------------------------------------------------------------------------
    def load(self, frames=None, eps=None, frames_u8=None, non_blocking=True):
        """Copy one batch into the static buffers.  `frames`: fp32 (B,3,64,64) in [0,1] (host or device);
        `frames_u8`: uint8 (B,64,64,3) as MineRL delivers them, converted on the device."""
        if frames is not None:
            self.x.copy_(frames, non_blocking=non_blocking)
        if frames_u8 is not None:
            self.x_u8.copy_(frames_u8, non_blocking=non_blocking)
        if eps is not None:
            self.eps.copy_(eps, non_blocking=non_blocking)

    def run(self, from_u8=False):
        """One optimizer step on the loaded batch.  Returns the device tensor [total, recon, KLD]."""
        if not self._use_graph:
            self._eager(from_u8)
            return self.losses
        if self._graphs is None:
            self._graphs = {}
        gs = self._graphs.get(from_u8)
        if gs is None:
            gs = self._graphs[from_u8] = self._capture(from_u8)
        gs[0].replay()
        if self.world > 1:
            self._allreduce()
            gs[1].replay()
        return self.losses
