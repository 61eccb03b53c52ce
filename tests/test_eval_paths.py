"""The callers either side of the hot path: the batched `-video` / `-thresh` drivers, the `-dataset`
generator, the uint8 input conversion and the double-buffered loader, against the oracle's per-frame restatement of
the reference loops (vae_utility.py:162-212, 256-284, 416-457, 324-343)."""
import numpy as np
import pytest
import torch

import critic_vae_oracle as O
import synth
from test_vae_module import _modules, PIXEL_ATOL

pytestmark = pytest.mark.gpu


def _episode(n, seed):
    frames = synth.make_frames(n, seed=seed)
    u8 = (frames.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).numpy()
    return frames, u8, synth.make_gt_masks(n, seed=seed + 1)


def _oracle_sweep(diff, mx, gt, thresholds):
    """vae_utility.py:148-160,56-68 per threshold, straight from the oracle's numpy restatement."""
    mxs = [np.float64(v) for v in mx]
    out = {}
    for t in thresholds:
        _, mask = O.diff_and_thr_masks(list(diff), mxs, thr=t)
        out[t] = O.iou(gt, mask)
    return out


@pytest.mark.parametrize("n,chunk", [(37, 16), (5, 16), (1100, None)])
def test_threshold_sweep_and_textured_frames_vs_oracle(critic_state, n, chunk, monkeypatch, tmp_path):
    """`vae.py -video -thresh`: chunked batches (sizes that are not multiples of the chunk, crossing the chunk boundary)
    give the per-frame result: reconstructions within the pixel tolerance of the oracle, difference maps bit-exact
    given those reconstructions, masks / IoU bit-exact given the difference maps, for all 13 thresholds at once."""
    import vae_utility as U
    if chunk is not None:
        monkeypatch.setattr(U, "_CHUNK", chunk)
    monkeypatch.chdir(tmp_path)            # eval_textured_frames writes bin_info_vae1.txt next to the caller, like the reference
    vae, critic = _modules(critic_state)
    vae.eval()
    frames, u8, gt = _episode(n, 300 + n)
    x, preds, r1, r0, diff, mx = U._score_episode(u8, vae, critic)
    assert torch.equal(x.cpu(), frames)                                        # uint8 -> fp32 /255 is exact
    pred_ref = O.critic_forward(critic_state, frames)
    np.testing.assert_allclose(preds.cpu().numpy().reshape(-1), pred_ref.numpy().reshape(-1), atol=5e-6)
    # reconstructions: the oracle's eval-mode network on the whole batch (the per-frame loop of vae_nets.py:42-46 is the
    # same arithmetic frame by frame)
    enc, dec = synth.make_vae_state(0)
    sel = slice(0, min(n, 64))
    mu, _ = O.encoder_forward(enc, frames[sel], training=False)
    ref1 = O.decoder_forward(dec, mu, pred_ref[sel].reshape(-1, 1))
    ref0 = O.decoder_forward(dec, mu, torch.zeros(mu.shape[0], 1))
    np.testing.assert_allclose(r1[sel].cpu().numpy(), ref1.numpy(), atol=PIXEL_ATOL)
    np.testing.assert_allclose(r0[sel].cpu().numpy(), ref0.numpy(), atol=PIXEL_ATOL)
    # difference maps: bit-exact against the ordered-sum restatement on the device's own reconstructions
    r1n, r0n = r1.cpu().numpy(), r0.cpu().numpy()
    probe = list(range(min(n, 8))) + ([n - 1] if n > 8 else []) + ([1023, 1024, 1025] if n > 1025 else [])
    for i in probe:
        d_ref, m_ref = O.diff_grey_ordered(r1n[i], r0n[i])
        assert np.array_equal(diff[i].cpu().numpy(), d_ref) and float(mx[i]) == float(m_ref), i
    # every threshold: integer counts / IoU equal to the oracle's mask pipeline on the same difference maps
    thresholds = list(range(0, 130, 10))
    sweep = U.eval_threshold_sweep(u8, vae, critic, gt, thresholds)
    want = _oracle_sweep(diff.cpu().numpy(), mx.cpu().numpy(), gt, thresholds)
    assert sweep == want
    if n <= 64:
        frames_out, thr_iou, crf_iou = U.eval_textured_frames(u8, vae, critic, gt)
        assert thr_iou == want[U.THRESHOLD] and len(frames_out) == n


def test_dataset_generation_matches_reference_loop(critic_state, monkeypatch):
    """`vae.py -dataset` (vae_utility.py:416-457): critic scores, the mid / high / low selection walk and the
    reconstructions appended per selected frame, against the reference's frame-by-frame loop restated with the oracle."""
    import vae_utility as U
    monkeypatch.setattr(U, "_CHUNK", 48)
    vae, critic = _modules(critic_state)
    vae.eval()
    n = 130
    frames, u8, _ = _episode(n, 500)
    got = U.dataset_from_trajectory(u8, critic, recon_dset=True, vae=vae)
    got_src = U.dataset_from_trajectory(u8, critic, recon_dset=False)
    preds_dev = critic.evaluate(frames.cuda()).reshape(-1).cpu()
    preds_ref = O.critic_forward(critic_state, frames).reshape(-1)
    np.testing.assert_allclose(preds_dev.numpy(), preds_ref.numpy(), atol=5e-6)
    # the reference loop (collect = 150 per bin), on the device's critic values so bin edges cannot flip
    enc, dec = synth.make_vae_state(0)
    want, want_src, c = [], [], {"mid": 0, "high": 0, "low": 0}
    for i in range(n):
        pr = float(preds_dev[i])
        if min(c.values()) >= 150:
            break
        if 0.4 <= pr <= 0.6 and c["mid"] < 150:
            b = "mid"
        elif pr >= 0.7 and c["high"] < 150:
            b = "high"
        elif pr <= 0.25 and c["low"] < 150:
            b = "low"
        else:
            continue
        c[b] += 1
        want_src.append(frames[i:i + 1].numpy())
        mu, _ = O.encoder_forward(enc, frames[i:i + 1], training=False)
        if b in ("mid", "high"):
            want.append(O.decoder_forward(dec, mu, preds_ref[i].reshape(1, 1)).numpy())
        if b in ("mid", "low"):
            want.append(O.decoder_forward(dec, mu, torch.zeros(1, 1)).numpy())
    assert len(got) == len(want) and len(got_src) == len(want_src) and len(want) > 0
    for a, b in zip(got_src, want_src):
        assert a.shape == (1, 3, 64, 64) and np.array_equal(a, b)
    for a, b in zip(got, want):
        assert a.shape == (1, 3, 64, 64) and a.dtype == np.float32
        np.testing.assert_allclose(a, b, atol=PIXEL_ATOL)


def test_frames_u8_to_f32_is_bit_exact():
    """cvae_frames_u8_to_f32 == astype(float32) / 255 + HWC -> CHW (vae_utility.py:324-343) for every byte value."""
    from cvae_native import binding as L
    n = 5
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 256, size=(n, 64, 64, 3), dtype=np.uint8)
    u8[0, 0, :, 0] = np.arange(64, dtype=np.uint8) * 4
    u8[0, 1, :, 1] = 255 - np.arange(64, dtype=np.uint8)
    u8.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
    src = torch.from_numpy(u8).cuda()
    dst = torch.empty(n, 3, 64, 64, device="cuda")
    L.check(L.lib.cvae_frames_u8_to_f32(n, src.data_ptr(), dst.data_ptr(), L.stream_ptr()))
    torch.cuda.synchronize()
    want = (u8.astype(np.float32) / np.float32(255.0)).transpose(0, 3, 1, 2)
    assert np.array_equal(dst.cpu().numpy(), want)
    L.check(L.lib.cvae_frames_u8_to_f32(0, None, None, L.stream_ptr()))       # empty input is a no-op


def test_frame_stager_equals_direct_steps(critic_state):
    """cvae_native.loader.FrameStager (pinned uint8 batches, copy stream, double buffering) runs the same optimizer
    steps as loading each batch directly, and hands back every step's loss in order."""
    from cvae_native.trainer import TrainStep
    from cvae_native.loader import FrameStager, pin_frames_u8
    B, steps = 16, 5
    batches = [pin_frames_u8(_episode(B, 700 + i)[1]) for i in range(steps)]
    eps = [synth.make_eps(B, seed=800 + i) for i in range(steps)]
    vae, critic = _modules(critic_state)
    vae.train()
    st = TrainStep(vae, critic, B)
    direct = []
    for b, e in zip(batches, eps):
        st.load(frames_u8=b.cuda(), eps=e.cuda())
        direct.append(st.run(from_u8=True).clone().cpu())
    flat_direct = st.eng.flat.clone()
    vae2, critic2 = _modules(critic_state)
    vae2.train()
    sg = FrameStager(TrainStep(vae2, critic2, B))
    staged = [loss.clone() for loss in sg.run(iter(batches), eps=iter(eps))]
    torch.cuda.synchronize()
    assert len(staged) == steps
    for a, b in zip(staged, direct):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-6)
    np.testing.assert_allclose(sg.step.eng.flat.cpu().numpy(), flat_direct.cpu().numpy(), atol=1e-7)
    assert list(sg.run(iter([]))) == []
    with pytest.raises(ValueError):
        pin_frames_u8(np.zeros((2, 3, 64, 64), dtype=np.uint8))
