"""Drop-in for the reference's vae.py: same flags (-train -inject -dataset -second -evalsecond -video
-thresh), same outputs, with the hot loops running on the B200 kernels.

    python vae.py -train                                   # single GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 vae.py -train   # batch-sharded over 8 GPUs

Reference behaviour followed: train() vae.py:33-66, image_evaluate() :68-108, mode dispatch :111-166.
Data-parallel training (new): every rank holds the full parameter set, takes a
contiguous 1/world slice of each shuffled global batch, and the flat fp32 gradient is summed with one
NCCL all-reduce before a fused Adam step that applies the 1/world mean (PyTorch-DDP semantics).
"""
import argparse
import os
import pickle
import statistics
from time import time

import numpy as np
import torch

torch.manual_seed(0)

from vae_parameters import *  # noqa: E402,F401,F403
from vae_nets import *  # noqa: E402,F401,F403
from vae_utility import *  # noqa: E402,F401,F403
from cvae_native.trainer import TrainStep, shard_batch  # noqa: E402
from cvae_native.loader import FrameStager  # noqa: E402


def _dist():
    """(rank, world, process_group) -- initialises NCCL when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, None
    import torch.distributed as dist
    if not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")      # one box, NVLink / NVSwitch only
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        dist.init_process_group("nccl")
    return dist.get_rank(), world, dist.group.WORLD


def _as_u8_frames(frames_f32):
    """float32 (N,3,64,64) frames that are exactly k/255 (MineRL povs after adjust_values, vae_utility.py:324-328)
    back to uint8 HWC; None when they are not (e.g. the reconstruction dataset of -second)."""
    u8 = np.rint(frames_f32 * 255.0)
    if not np.array_equal(u8.astype(np.float32) / np.float32(255.0), frames_f32):
        return None
    return np.ascontiguousarray(u8.astype(np.uint8).transpose(0, 2, 3, 1))


def _train_streamed(autoencoder, critic, frames_u8, logger, rank, world, pg):
    """The same loop with the dataset in pinned host memory (uint8, 12 KB per frame) and double-buffered H2D staging
    (cvae_native.loader.FrameStager) instead of an fp32 copy resident in HBM.  Chosen with CVAE_STREAM_FRAMES=1 or
    when the fp32 dataset would not fit comfortably in device memory."""
    pinned = torch.as_tensor(frames_u8).pin_memory()
    num_samples = pinned.shape[0]
    seed_rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little") if world == 1 else 0)
    if world > 1:
        torch.manual_seed(torch.initial_seed() + rank)      # per-rank reparameterisation noise
    stagers = {}
    for ep in range(epochs):
        order = seed_rng.permutation(num_samples)
        starts = list(range(0, num_samples, batch_size))
        by_size = {}
        for batch_i in starts:                               # the last, shorter batch is kept (vae.py:44-46)
            idx = shard_batch(torch.as_tensor(order[batch_i:batch_i + batch_size]), rank, world)
            if idx.numel():
                by_size.setdefault(idx.numel(), []).append((batch_i, idx))
        for B, items in by_size.items():
            sg = stagers.get(B)
            if sg is None:
                sg = stagers[B] = FrameStager(TrainStep(autoencoder, critic, B, lr=lr, process_group=pg))
            gather = (pinned.index_select(0, idx).pin_memory() for _, idx in items)
            for (batch_i, _), losses in zip(items, sg.run(gather)):
                if batch_i % log_n == 0 and rank == 0:
                    print(f'    ep:{ep}, imgs:{num_samples * ep + (batch_i + 1)}', end='\r')
                    if logger is not None:
                        log_info({'total_loss': losses[0], 'recon_loss': losses[1], 'KLD': losses[2]}, logger, batch_i, ep, num_samples)
                    autoencoder._engine.check_fault()
    autoencoder._engine.check_fault()
    return autoencoder


def train(autoencoder, dset, logger=None, critic=None):
    """vae.py:33-66.  `dset`: list of (1,3,64,64) float32 frames (or an (N,3,64,64) array)."""
    critic = critic if critic is not None else globals().get("critic")
    rank, world, pg = _dist()
    host = np.stack(dset).squeeze().reshape(-1, ch, w, w).astype(np.float32, copy=False)
    autoencoder.train()
    free_bytes = torch.cuda.mem_get_info()[0] if torch.cuda.is_available() else 0
    if os.environ.get("CVAE_STREAM_FRAMES") == "1" or host.nbytes > 0.4 * free_bytes:
        u8 = _as_u8_frames(host)
        if u8 is not None:
            return _train_streamed(autoencoder, critic, u8, logger, rank, world, pg)
    data = torch.as_tensor(host).to(device)                  # resident in HBM
    num_samples = data.shape[0]
    steps = {}
    seed_rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little") if world == 1 else 0)
    if world > 1:
        torch.manual_seed(torch.initial_seed() + rank)       # per-rank reparameterisation noise (ranks share the data order only)
    losses = None
    for ep in range(epochs):
        order = torch.as_tensor(seed_rng.permutation(num_samples), device=data.device)
        for batch_i in range(0, num_samples, batch_size):
            idx = order[batch_i:batch_i + batch_size]          # the last, shorter batch is kept (vae.py:44-46)
            idx = shard_batch(idx, rank, world)                # same count on every rank; an empty slice is empty everywhere
            if idx.numel() == 0:
                continue
            B = idx.numel()
            st = steps.get(B)
            if st is None:
                st = steps[B] = TrainStep(autoencoder, critic, B, lr=lr, process_group=pg)
            torch.index_select(data, 0, idx, out=st.x)
            st.eps.normal_()
            losses = st.run()
            if batch_i % log_n == 0 and rank == 0:
                print(f'    ep:{ep}, imgs:{num_samples * ep + (batch_i + 1)}', end='\r')
                if logger is not None:
                    log_info({'total_loss': losses[0], 'recon_loss': losses[1], 'KLD': losses[2]}, logger, batch_i, ep, num_samples)
                autoencoder._engine.check_fault()      # a device-side pipeline fault must not train on garbage until the end
    autoencoder._engine.check_fault()
    return autoencoder


def image_evaluate(autoencoder, critic):
    """vae.py:68-108."""
    from PIL import Image
    print('evaluating source images...')
    os.makedirs(SAVE_PATH, exist_ok=True)
    if args.inject:
        os.makedirs(INJECT_PATH, exist_ok=True)
    imgs, diff_max_values = [], []
    for i, img_file in enumerate(os.listdir(SOURCE_IMAGES_PATH)):
        img_tensor = preprocess_observation(Image.open(f'{SOURCE_IMAGES_PATH}/{img_file}'))
        pred = critic.evaluate(img_tensor)
        if args.inject:
            get_injected_img(autoencoder, img_tensor, pred[0]).save(f'{INJECT_PATH}image-{i:03d}.png', format="png")
        else:
            ro, rz, diff, max_value = get_diff_image(autoencoder, img_tensor, pred[0])
            imgs.append([img_tensor, ro, rz, diff, pred[0]])
            diff_max_values.append(max_value)
    if not args.inject and imgs:
        diffs, _ = get_diff_and_thr_masks([m[3] for m in imgs], diff_max_values)
        for i, m in enumerate(imgs):
            get_final_frame(m[0], m[1], m[2], Image.fromarray(diffs[i]), m[4]).save(f'{SAVE_PATH}/image-{i:03d}.png', format="png")


def main():
    global args, critic, vae
    parser = argparse.ArgumentParser()
    for flag in ('-train', '-inject', '-dataset', '-second', '-evalsecond', '-video', '-thresh'):
        parser.add_argument(flag, action='store_true')
    args = parser.parse_args()
    rank, world, _ = _dist()
    vae = VariationalAutoencoder().to(device)

    if args.video:
        load_vae_network(vae)
        critic = load_critic(CRITIC_PATH)
        frames, gt_frames = load_textured_minerl()
        if args.thresh:
            print('testing thresholds (thr):')
            for t, iou in eval_threshold_sweep(frames, vae, critic, gt_frames).items():
                print(f'thr={t}, thr_iou={iou}')
        vae_frames, thr_iou, crf_iou = eval_textured_frames(frames, vae, critic, gt_frames)
        print(f'thr_iou={thr_iou}')
        print(f'crf_iou={crf_iou}')
        create_video(vae_frames)
    elif args.dataset:
        load_vae_network(vae)
        critic = load_critic(CRITIC_PATH)
        with open(SAVE_DATASET_PATH, 'wb') as file:
            pickle.dump(load_minerl_data(critic, recon_dset=True, vae=vae), file)
    elif args.second:
        print('training second vae...')
        critic = load_critic(CRITIC_PATH)
        with open(SAVE_DATASET_PATH, 'rb') as file:
            recon_dset = pickle.load(file)
        vae = train(vae, recon_dset, critic=critic)
        if rank == 0:
            torch.save(vae.encoder.state_dict(), SECOND_ENCODER_PATH)
            torch.save(vae.decoder.state_dict(), SECOND_DECODER_PATH)
    elif args.evalsecond:
        critic = load_critic(CRITIC_PATH)
        load_vae_network(vae, second_vae=True)
        image_evaluate(vae, critic)
    else:
        critic = load_critic(CRITIC_PATH)
        if args.train:
            logger = None
            if rank == 0:
                from logger import Logger
                logger = Logger('./logs/vae' + str(time())[-5::])
            vae = train(vae, load_minerl_data(critic), logger=logger, critic=critic)
            if rank == 0:
                torch.save(vae.encoder.state_dict(), ENCODER_PATH)
                torch.save(vae.decoder.state_dict(), DECODER_PATH)
        else:
            load_vae_network(vae)
            image_evaluate(vae, critic)


if __name__ == "__main__":
    main()
