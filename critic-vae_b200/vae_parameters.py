"""Configuration globals, star-imported by vae.py / vae_nets.py / vae_utility.py exactly like the
reference's vae_parameters.py (same names and values, vae_parameters.py:1-41 there), so user code that
reads or overrides them keeps working."""
import os as _os

import torch

# device the reference would pick (vae_parameters.py:2); under torchrun each rank uses its LOCAL_RANK GPU
if torch.cuda.is_available():
    device = f"cuda:{int(_os.environ.get('LOCAL_RANK', 0))}" if "LOCAL_RANK" in _os.environ else "cuda:0"
else:
    device = "cpu"

# frame geometry
w, ch = 64, 3

# optimisation
epochs, batch_size, lr = 7, 128, 0.00005
k, p, step = 5, 2, 1                  # conv kernel size, padding, stride
bottleneck, latent_dim = 4096, 32     # 4x4x256 conv bottleneck -> 32-d latent
kld_weight = 0.001
total_images = 50000
log_n = batch_size * 30               # samples between log lines
inject_n = 6

# files and folders
ENCODER_PATH, DECODER_PATH = 'saved-networks/vae_encoder.pt', 'saved-networks/vae_decoder.pt'
SECOND_ENCODER_PATH, SECOND_DECODER_PATH = 'vae2_encoder.pt', 'vae2_decoder.pt'
SOURCE_IMAGES_PATH, SAVE_PATH, INJECT_PATH, VIDEO_PATH = 'source-images/', 'images/', 'inject/', 'videos/'
SAVE_DATASET_PATH = 'recon-dataset.pickle'
MINERL_EPISODE_PATH = 'minerl-episode/'
_CRITIC_FMT = 'saved-networks/critic-rewidx=1-cepochs=15-datamode=trunk-datasize={}-shift=12-chfak=1-dropout=0.3.pt'
CRITIC_PATH, SECOND_CRITIC_PATH = _CRITIC_FMT.format(99999), _CRITIC_FMT.format(100000)
MINERL_DATA_ROOT_PATH = _os.environ.get(
    'MINERL_DATA_ROOT', '/homes/lcicek/anaconda3/envs/vae/lib/python3.6/site-packages/minerl')
