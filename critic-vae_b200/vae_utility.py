"""Drop-in for the reference's vae_utility.py: same function names, arguments and return values, with
the hot parts (critic scoring, the two decodes per frame, difference map, normalise / quantise /
threshold, IoU counts) moved onto the GPU and batched.  Reference lines are cited per function
(paths relative to the reference checkout).

What changes in behaviour (all additive):
  * eval_textured_frames / load_minerl_data process whole batches instead of one frame per call and
    run the encoder once per frame (the reference runs it twice, vae_nets.py:43 via vae_utility.py:264-265);
  * the IoU for every threshold of the `-thresh` sweep comes from one histogram pass
    (eval_threshold_sweep) instead of 13 complete re-runs (vae.py:121-123);
  * crf() needs the optional third-party `denseCRF` (pip SimpleCRF); without it crf_iou is None.
"""
import os
import statistics
from collections import defaultdict
from io import BytesIO

import numpy as np
import torch
from torch import Tensor
from PIL import Image, ImageDraw, ImageFont

from vae_parameters import *  # noqa: F401,F403
from vae_nets import *  # noqa: F401,F403
from cvae_native import binding as _L

THRESHOLD = 50


def _load_font():
    for path in ("/usr/share/fonts/truetype/ubuntu/Ubuntu-R.ttf", "/usr/share/fonts/truetype/dejavu/DejaVuSans.ttf"):
        try:
            return ImageFont.truetype(path, 10)
        except OSError:
            pass
    return ImageFont.load_default()


font = _load_font()
titles = ["orig img\n+crit val", "crit val\ninjected", "crit=0\ninjected", "difference\nmask",
          f"thr-mask\nthr={THRESHOLD}", "thr-mask +\ncrf", "ground\ntruth"]
_CHUNK = 1024          # frames per device batch in the eval drivers


# ------------------------------------------------------------------------------------------------
# device helpers
# ------------------------------------------------------------------------------------------------
def _dev():
    return torch.device(device)


def _p(t):
    return None if t is None else t.data_ptr()


def frames_to_device(frames_u8):
    """uint8 (N,64,64,3) -> fp32 (N,3,64,64) on the GPU; same arithmetic as adjust_values + transpose."""
    src = torch.as_tensor(np.ascontiguousarray(frames_u8), dtype=torch.uint8).to(_dev())
    dst = torch.empty(src.shape[0], ch, w, w, device=_dev())
    _L.check(_L.lib.cvae_frames_u8_to_f32(src.shape[0], _p(src), _p(dst), _L.stream_ptr()))
    return dst


def _recon_pair(vae, x, preds, one=False):
    """recon(high) and recon(0) for a batch: encoder once, decoder twice."""
    eng, ws = vae._prepare(x.shape[0])
    eng.encode(x, vae.encoder.training, ws)
    hi = torch.ones(x.shape[0], device=x.device) if one else preds.reshape(-1).to(torch.float32).contiguous()
    r_hi = eng.decode(hi, None, False, ws).clone()
    r_lo = eng.decode(torch.zeros(x.shape[0], device=x.device), None, False, ws).clone()
    return r_hi, r_lo


def _diff_grey(r_hi, r_lo):
    n = r_hi.shape[0]
    diff = torch.empty(n, w, w, dtype=torch.float64, device=r_hi.device)
    mx = torch.empty(n, dtype=torch.float64, device=r_hi.device)
    _L.check(_L.lib.cvae_diff_grey(n, _p(r_hi), _p(r_lo), _p(diff), _p(mx), _L.stream_ptr()))
    return diff, mx


def _mask_iou(diff_dev, gt_dev, mean_max, factor, thr, thr_list=()):
    n = diff_dev.shape[0]
    dev = diff_dev.device
    u8 = torch.empty(n, w, w, dtype=torch.uint8, device=dev)
    mask = torch.empty(n, w, w, dtype=torch.uint8, device=dev)
    hist = torch.empty(512, dtype=torch.int64, device=dev)
    thr_t = torch.tensor(list(thr_list) or [thr], dtype=torch.int32, device=dev)
    counts = torch.empty(thr_t.numel(), 3, dtype=torch.int64, device=dev)
    if gt_dev is None:
        gt_dev = torch.zeros(n, w, w, dtype=torch.uint8, device=dev)
    _L.check(_L.lib.cvae_mask_iou(n, _p(diff_dev), _p(gt_dev), float(mean_max), float(factor), int(thr), thr_t.numel(),
                                  _p(thr_t), _p(u8), _p(mask), _p(hist), _p(counts), _L.stream_ptr()))
    return u8, mask, counts


def _iou_from_counts(tp, fn, fp):
    iou = 1 if tp + fn + fp == 0 else tp / (tp + fn + fp)   # vae_utility.py:61-66
    return round(iou, 3)


# ------------------------------------------------------------------------------------------------
# reference API
# ------------------------------------------------------------------------------------------------
def get_iou(G, T):
    """vae_utility.py:56-68; the counting runs on the GPU, division and rounding as in the reference."""
    g = torch.as_tensor(np.ascontiguousarray(np.asarray(G, dtype=bool)).view(np.uint8)).to(_dev())
    t = torch.as_tensor(np.ascontiguousarray(np.asarray(T, dtype=bool)).view(np.uint8)).to(_dev())
    if g.shape != t.shape:
        g, t = torch.broadcast_tensors(g, t)
        g, t = g.contiguous(), t.contiguous()
    counts = torch.empty(3, dtype=torch.int64, device=_dev())
    _L.check(_L.lib.cvae_iou_counts(g.numel(), _p(g), _p(t), _p(counts), _L.stream_ptr()))
    tp, fn, fp = (int(v) for v in counts.cpu())
    return _iou_from_counts(tp, fn, fp)


def load_textured_minerl():
    """vae_utility.py:70-82."""
    text_dset = np.load(MINERL_EPISODE_PATH + "X.npy")
    gt_dset = np.expand_dims(np.all(np.load(MINERL_EPISODE_PATH + "Y.npy"), axis=-1), axis=-1)
    text_dset = text_dset[100:5000:2]
    gt_dset = gt_dset[100:5000:2].transpose(0, 3, 1, 2).squeeze()
    return text_dset, gt_dset


def create_video(frames):
    """vae_utility.py:85-104."""
    print('creating video...')
    os.makedirs(VIDEO_PATH, exist_ok=True)
    buffers = []
    for f in frames:
        b = BytesIO()
        f.save(b, format="GIF")
        buffers.append(b)
    imgs = [Image.open(b) for b in buffers]
    imgs[0].save(f"{VIDEO_PATH}video-threshold={THRESHOLD}.gif", format='GIF', duration=100, save_all=True, loop=0,
                 append_images=imgs[1:])


def get_diff_factor(max_values):
    """vae_utility.py:106-110 (exact-rational mean on the host, as the reference)."""
    mean_max = statistics.mean(max_values)
    diff_factor = 1.0 / mean_max if mean_max != 0 else 0
    return diff_factor, mean_max


def save_bin_info_file(bin_ious, bin_frames, bin_gts):
    """vae_utility.py:112-130."""
    total_gt = np.sum(list(bin_gts.values()))
    n_frames = max(1, sum(bin_frames.values()))
    with open('bin_info_vae1.txt', 'w') as f:
        f.write('ground truth pixels sorted by bin:\n')
        for b, count in bin_gts.items():
            f.write(f'bin: {b}, pixels = {count} = {round(count / total_gt, 2) * 100 if total_gt else 0}%\n')
        f.write('\nframes separated by bin:\n')
        for b, count in bin_frames.items():
            f.write(f'bin: {b}, frames = {count} = {round(count / n_frames, 2) * 100}%\n')
        f.write('\niou-mean and std:\n')
        for b, vals in bin_ious.items():
            std = round(statistics.stdev(vals), 2) if len(vals) > 1 else 0.0
            f.write(f'bin: {b}, iou_mean={round(statistics.mean(vals), 2)}, iou_std={std}\n')


def save_bin_info(preds, gt, thr_masks):
    """vae_utility.py:132-145: per critic-value bin IoU statistics (per-frame integer counts on the host:
    1200 tiny frames, reporting only)."""
    bin_ious, bin_frames, bin_gts = defaultdict(list), defaultdict(int), defaultdict(int)
    thr_masks, gt = np.asarray(thr_masks, dtype=bool), np.asarray(gt, dtype=bool)
    for i, pred in enumerate(preds):
        b = round(float(pred), 1)
        tp = int(np.sum(thr_masks[i] & gt[i])); fn = int(np.sum(thr_masks[i] & ~gt[i])); fp = int(np.sum(~thr_masks[i] & gt[i]))
        bin_ious[b].append(_iou_from_counts(tp, fn, fp))
        bin_frames[b] += 1
        bin_gts[b] += int(gt[i].sum())
    save_bin_info_file(bin_ious, bin_frames, bin_gts)


def get_diff_and_thr_masks(diff_masks, max_values, thr=THRESHOLD):
    """vae_utility.py:148-160 on the GPU: clamp to mean_max, scale, quantise to uint8, threshold."""
    diff_factor, mean_max = get_diff_factor(list(max_values))
    d = torch.as_tensor(np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.float64) for x in diff_masks]))).to(_dev())
    u8, mask, _ = _mask_iou(d, None, mean_max, diff_factor, thr)
    return u8.cpu().numpy(), mask.cpu().numpy().astype(bool)


def crf(imgs, mask, Y, skip=1):
    """vae_utility.py:22-54.  Dense-CRF refinement through the optional third-party `denseCRF`
    (pip SimpleCRF); parameters as in the reference."""
    import denseCRF  # noqa: deferred so the hot path does not depend on it
    param = (22, 12, 3.1, 8, 1.8, 10)     # w1, alpha, beta, w2, gamma, iterations
    mask = mask.copy()
    M = mask[::skip]
    for i, img in enumerate(imgs[::skip]):
        frame = M[i, 0]
        M[i, 0] = denseCRF.densecrf(img, np.stack((1 - frame, frame), axis=-1), param)
    mask[::skip] = M
    return mask >= 1


def _score_episode(trajectory, vae, critic):
    """Batched core of eval_textured_frames (vae_utility.py:171-181): per frame critic value,
    recon(pred), recon(0), fp64 difference map and its max -- all resident on the device."""
    traj = np.asarray(trajectory)
    xs, preds, r1s, r0s, diffs, maxes = [], [], [], [], [], []
    for s in range(0, traj.shape[0], _CHUNK):
        x = frames_to_device(traj[s:s + _CHUNK])
        p = critic.evaluate(x)
        r_hi, r_lo = _recon_pair(vae, x, p)
        d, m = _diff_grey(r_hi, r_lo)
        xs.append(x); preds.append(p); r1s.append(r_hi); r0s.append(r_lo); diffs.append(d); maxes.append(m)
    cat = lambda ts: torch.cat(ts) if ts else torch.empty(0, device=_dev())
    vae._engine.check_fault()        # a device-side pipeline fault must never turn into silently wrong masks
    return cat(xs), cat(preds), cat(r1s), cat(r0s), cat(diffs), cat(maxes)


def eval_threshold_sweep(trajectory, vae, critic, gt, thresholds=range(0, 130, 10)):
    """All thresholds of `vae.py -video -thresh` (vae.py:121-123) from ONE network pass and ONE
    histogram pass.  Returns {thr: thr_iou}."""
    _, _, _, _, diff, mx = _score_episode(trajectory, vae, critic)
    factor, mean_max = get_diff_factor([np.float64(v) for v in mx.cpu().numpy()])
    gt_dev = torch.as_tensor(np.ascontiguousarray(np.asarray(gt, dtype=bool)).view(np.uint8)).to(_dev())
    thr_list = list(thresholds)
    _, _, counts = _mask_iou(diff, gt_dev, mean_max, factor, thr_list[0], thr_list)
    return {t: _iou_from_counts(*(int(v) for v in c)) for t, c in zip(thr_list, counts.cpu())}


def eval_textured_frames(trajectory, vae, critic, gt, t=THRESHOLD):
    """vae_utility.py:162-212: returns (composited frames, thr_iou, crf_iou)."""
    print('processing frames...')
    x, preds, r_one, r_zero, diff, mx = _score_episode(trajectory, vae, critic)
    factor, mean_max = get_diff_factor([np.float64(v) for v in mx.cpu().numpy()])
    gt_np = np.asarray(gt, dtype=bool)
    gt_dev = torch.as_tensor(np.ascontiguousarray(gt_np).view(np.uint8)).to(_dev())
    u8, mask, counts = _mask_iou(diff, gt_dev, mean_max, factor, t)
    thr_iou = _iou_from_counts(*(int(v) for v in counts[0].cpu()))
    diff_masks, thr_masks = u8.cpu().numpy(), mask.cpu().numpy().astype(bool)

    crf_masks, crf_iou = None, None
    try:
        crf_masks = crf(np.asarray(trajectory)[:, np.newaxis, ...], thr_masks[:, np.newaxis, ...].astype(np.float32),
                        gt_np[..., np.newaxis]).squeeze()
        crf_iou = get_iou(gt_np, crf_masks)
    except ImportError:
        print('denseCRF (pip SimpleCRF) is not installed: skipping the CRF refinement, crf_iou=None')

    one_np, zero_np, preds_c = r_one.cpu().numpy(), r_zero.cpu().numpy(), preds.cpu()
    x_c = x.cpu()
    ret = []
    for i in range(x_c.shape[0]):
        ret.append(get_final_frame(
            x_c[i:i + 1], one_np[i], zero_np[i], Image.fromarray(diff_masks[i]), preds_c[i],
            gt_img=Image.fromarray(gt_np[i]), thr_img=Image.fromarray(thr_masks[i]),
            crf_img=Image.fromarray(crf_masks[i] if crf_masks is not None else np.zeros_like(thr_masks[i])),
            thr_iou=thr_iou, crf_iou=crf_iou))
    save_bin_info(preds_c.reshape(-1).tolist(), gt_np, thr_masks)
    return ret, thr_iou, crf_iou


def select_balanced(preds, collect=150):
    """The per-trajectory critic-bin selection of vae_utility.py:431-457: walk the frames in order,
    keep up to `collect` mid (0.4..0.6), high (>= 0.7) and low (<= 0.25) frames, stop once all three
    bins are full.  Returns a list of (frame index, bin) with bin in {'mid','high','low'}."""
    picked, c = [], {'mid': 0, 'high': 0, 'low': 0}
    for i, pr in enumerate(preds):
        if min(c.values()) >= collect:
            break
        if 0.4 <= pr <= 0.6 and c['mid'] < collect:
            b = 'mid'
        elif pr >= 0.7 and c['high'] < collect:
            b = 'high'
        elif pr <= 0.25 and c['low'] < collect:
            b = 'low'
        else:
            continue
        c[b] += 1
        picked.append((i, b))
    return picked


def collect_frames(trajectory_names):
    """vae_utility.py:214-238 (needs the `minerl` package and dataset)."""
    print('collecting frames...')
    import minerl
    os.environ['MINERL_DATA_ROOT'] = MINERL_DATA_ROOT_PATH
    data = minerl.data.make('MineRLTreechop-v0', num_workers=1)
    all_frames = []
    for name in trajectory_names:
        frames = []
        for obs, _, _, _, _ in data.load_data(name, skip_interval=0, include_metadata=False):
            frames.append(preprocess_observation(obs["pov"]))
            if len(frames) >= 1000:
                all_frames.append(frames)
                break
    del data
    return all_frames


def get_injected_img(autoencoder, img_tensor, pred):
    """vae_utility.py:240-254."""
    recons = autoencoder.inject(img_tensor)
    tiles = [to_np(img_tensor.view(-1, ch, w, w)[0])] + [to_np(r.view(-1, ch, w, w)[0]) for r in recons[:inject_n]]
    _, img = prepare_rgb_image(np.concatenate(tiles, axis=2))
    return img


def get_diff_image(autoencoder, img_tensor, pred, one=False):
    """vae_utility.py:256-277 for one frame (or a batch): returns recon_one, recon_zero (3,64,64) fp32,
    diff (64,64) float64 and its max, computed on the GPU."""
    x = img_tensor.to(_dev(), torch.float32).view(-1, ch, w, w).contiguous()
    r_hi, r_lo = _recon_pair(autoencoder, x, torch.as_tensor(pred).to(_dev()).reshape(-1).expand(x.shape[0]), one=one)
    d, m = _diff_grey(r_hi, r_lo)
    return to_np(r_hi[0]), to_np(r_lo[0]), d[0].cpu().numpy(), np.float64(m[0].item())


def prepare_diff(diff_img, diff_factor, mean_max):
    """vae_utility.py:279-284 (host version kept for callers that hold numpy arrays, e.g. image_evaluate)."""
    diff_img[diff_img > mean_max] = mean_max
    return diff_img * diff_factor


def get_final_frame(img_tensor, recon_one, recon_zero, diff_img, pred, gt_img=None, thr_img=None, crf_img=None,
                    thr_iou=None, crf_iou=None):
    """vae_utility.py:286-322: 4 (or 7, with masks) 64x64 tiles side by side, titles on top."""
    strip = np.concatenate((to_np(img_tensor.view(-1, ch, w, w)[0]), recon_one, recon_zero), axis=2)
    _, strip_img = prepare_rgb_image(strip)
    with_masks = gt_img is not None
    n_tiles, top = (7, w) if with_masks else (4, 0)
    img = Image.new('RGB', (w * n_tiles, top + w))
    draw = ImageDraw.Draw(img)
    img.paste(strip_img, (0, top))
    img.paste(diff_img, (w * 3, top))
    if with_masks:
        for j, tile in enumerate((thr_img, crf_img, gt_img)):
            img.paste(tile, (w * (4 + j), top))
        for i, title in enumerate(titles):
            if i == 4:
                title += f"\niou={thr_iou}"
            elif i == 5:
                title += f"\niou={crf_iou}"
            draw.text((w * i + 2, 0), title, (255, 255, 255), font=font)
    draw.text((2, top + 2), f'{float(pred):.1f}', (255, 255, 255), font=font)
    return img


def adjust_values(obs):
    """vae_utility.py:324-328."""
    return np.array(obs).astype(np.float32) / 255


def reverse_preprocess(recon):
    """vae_utility.py:330-335."""
    return (to_np(recon.view(-1, ch, w, w)[0]).transpose(1, 2, 0) * 255).astype(np.uint8)


def preprocess_observation(obs):
    """vae_utility.py:337-343."""
    return Tensor(adjust_values(obs).transpose(2, 0, 1)[np.newaxis, ...]).to(device)


def load_vae_network(vae, second_vae=False):
    """vae_utility.py:345-361 (missing weight files are tolerated exactly like the reference)."""
    enc_path, dec_path = (SECOND_ENCODER_PATH, SECOND_DECODER_PATH) if second_vae else (ENCODER_PATH, DECODER_PATH)
    try:
        vae.encoder.load_state_dict(torch.load(enc_path, map_location=device))
        vae.decoder.load_state_dict(torch.load(dec_path, map_location=device))
    except Exception as e:
        print(e)
    vae.eval()
    vae.encoder.eval()
    vae.decoder.eval()


def load_critic(path):
    """vae_utility.py:363-370."""
    from critic_net import Critic
    critic = Critic()
    critic.load_state_dict(torch.load(path, map_location='cpu'))
    critic.eval()
    critic.to(device)
    return critic


def log_info(losses, logger, batch_i, ep, num_samples):
    """vae_utility.py:372-380."""
    for tag, key in (('recon_loss', 'recon_loss'), ('kld', 'KLD'), ('total_loss', 'total_loss')):
        logger.scalar_summary(tag, losses[key].item(), batch_i + (num_samples * ep))


def to_np(x):
    return x.data.cpu().numpy()


def prepare_rgb_image(img_array):
    """vae_utility.py:385-390."""
    arr = (np.transpose(img_array, (1, 2, 0)) * 255).astype(np.uint8)
    return arr, Image.fromarray(arr, mode='RGB')


def load_minerl_data(critic, recon_dset=False, vae=None):
    """vae_utility.py:393-462 with the per-frame critic / VAE calls batched per trajectory (needs the
    `minerl` package and the MineRLTreechop-v0 dataset, neither of which ships with this repository)."""
    print("loading minerl-data...")
    import minerl
    os.environ['MINERL_DATA_ROOT'] = MINERL_DATA_ROOT_PATH
    data = minerl.data.make('MineRLTreechop-v0', num_workers=1)
    names = data.get_trajectory_names()
    np.random.default_rng(seed=0).shuffle(names)
    dset = []
    for name in names:
        if len(dset) >= total_images:
            break
        print(f'total images = {len(dset)}')
        povs = np.stack([o["pov"] for o, _, _, _, _ in data.load_data(name, skip_interval=0, include_metadata=False)])
        dset.extend(dataset_from_trajectory(povs, critic, recon_dset=recon_dset, vae=vae))
    del data
    return dset


def dataset_from_trajectory(povs, critic, recon_dset=False, vae=None):
    """The per-trajectory body of load_minerl_data (vae_utility.py:416-457), batched: critic values of every frame in
    one call (chunks of _CHUNK), the balanced mid / high / low selection on the host, then for the `-dataset` path
    (recon_dset) the reconstructions with the critic value and with 0 of the selected frames in one batch.  Returns
    the list of (1,3,64,64) float32 arrays the reference appends, in the reference's order."""
    povs = np.asarray(povs)
    x = torch.cat([frames_to_device(povs[s:s + _CHUNK]) for s in range(0, povs.shape[0], _CHUNK)]) if povs.shape[0] else None
    if x is None:
        return []
    preds = torch.cat([critic.evaluate(x[s:s + _CHUNK]).reshape(-1) for s in range(0, x.shape[0], _CHUNK)])
    picked = select_balanced(preds.cpu().tolist())
    out = []
    if not picked:
        return out
    idx = torch.tensor([i for i, _ in picked], device=x.device)
    if recon_dset:
        r_hi, r_lo = _recon_pair(vae, x[idx], preds[idx])
        vae._engine.check_fault()
        r_hi, r_lo = r_hi.cpu().numpy(), r_lo.cpu().numpy()
        for j, (_, b) in enumerate(picked):
            if b in ('mid', 'high'):
                out.append(r_hi[j:j + 1])
            if b in ('mid', 'low'):
                out.append(r_lo[j:j + 1])
    else:
        sel = x[idx].cpu().numpy()
        out.extend(sel[j:j + 1] for j in range(sel.shape[0]))
    return out
