"""CPU tests of the post-processing rows next to the hot path: orbit-video camera trajectory and
frame layout against the reference's own outputs (`oracle/make_golden_post.py` -> tests/golden/post.npz), the
`metric_log.txt` text byte for byte, the two-rank gather of per-image metrics (gloo), and the Motion-JPEG writer."""
import json
import os
import tempfile
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_orbit_cameras_match_reference_frame_by_frame():
    from spi_b200.utils.video_utils import orbit_cameras
    gold = np.load(os.path.join(GOLD, 'post.npz'))['orbit_poses']
    cams, poses = orbit_cameras(120, device='cpu')
    assert cams.shape == (120, 25) and poses.shape == (120, 4, 4)
    np.testing.assert_allclose(poses.numpy(), gold, rtol=0, atol=2e-7)          # batched vs one-by-one evaluation
    np.testing.assert_array_equal(cams[:, :16].numpy(), poses.reshape(120, 16).numpy())
    np.testing.assert_allclose(cams[:, 16:].numpy(), np.tile(np.float32([4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1]), (120, 1)))
    # every camera sits on the radius-2.7 sphere around the look-at point's origin
    np.testing.assert_allclose(np.linalg.norm(poses[:, :3, 3].numpy(), axis=1), 2.7, rtol=1e-6)


def test_layout_grid_matches_reference():
    from spi_b200.utils.video_utils import layout_grid
    g = np.load(os.path.join(GOLD, 'post.npz'))
    out = layout_grid(torch.from_numpy(g['grid_in']), grid_w=3, grid_h=2)
    assert out.dtype == np.uint8
    np.testing.assert_array_equal(out, g['grid_out'])


def test_metric_log_text_is_the_reference_text():
    from spi_b200.utils.metric_utils import format_metric_log
    g = json.load(open(os.path.join(GOLD, 'metric_log.json')))
    hp = types.SimpleNamespace(**g['hyperparameters'])
    assert format_metric_log(g['coach_name'], hp, g['metric_dic']) == g['text']


def _metric_worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from spi_b200.configs import hyperparameters, paths_config
    from spi_b200.training.coaches.base_coach import BaseCoach
    g = json.load(open(os.path.join(GOLD, 'metric_log.json')))
    for k, v in g['hyperparameters'].items():
        setattr(hyperparameters, k, v)
    paths_config.experiments_output_dir = out_dir
    # rank 0 holds images 0-2 (dataset block 1/2 of 5 images: len // W + 1 = 3), rank 1 images 3-4
    lo, hi = (0, 3) if rank == 0 else (3, 5)
    part = {mode: {k: v[lo:hi] for k, v in cur.items()} for mode, cur in g['metric_dic'].items()}
    fake = types.SimpleNamespace(coach_name=g['coach_name'], metric_dic=part)
    merged = BaseCoach.log_metric(fake)
    assert merged == g['metric_dic']
    dist.barrier()
    dist.destroy_process_group()


def test_metric_log_two_ranks_write_the_single_process_file():
    g = json.load(open(os.path.join(GOLD, 'metric_log.json')))
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_metric_worker, args=(2, 29541, d), nprocs=2, join=True)
        assert open(os.path.join(d, 'metric_log.txt')).read() == g['text']       # written once, by rank 0


def test_cal_metric_rows_and_mirror_flip():
    from spi_b200.training.coaches.base_coach import BaseCoach
    calls = []

    class M:
        def run(self, gt, fake):
            calls.append((gt.clone(), fake.clone()))
            return float(len(calls)), 0.5, 0.25
    fake_self = types.SimpleNamespace(metric_dic={}, metric=M())
    gt = torch.arange(8.).view(1, 1, 2, 4)
    BaseCoach.cal_metric(fake_self, gt + 1, gt, 'w_inv', fake_m=gt + 2)
    assert fake_self.metric_dic['w_inv'] == {'l2': [1.0], 'lpips': [0.5], 'id': [0.25], 'l2_m': [2.0], 'lpips_m': [0.5], 'id_m': [0.25]}
    assert torch.equal(calls[1][0], torch.flip(gt, dims=[3]))       # the mirrored render is compared with the flipped photo


def test_mjpeg_writer_round_trip():
    from spi_b200.utils.video_utils import MJPEGWriter
    cv2 = pytest.importorskip('cv2')
    rs = np.random.RandomState(0)
    base = np.kron(rs.randint(0, 255, (8, 8, 3)), np.ones((8, 8, 1))).astype(np.uint8)          # blocky: survives JPEG
    frames = [np.roll(base, 8 * i, axis=1) for i in range(5)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'clip.avi')
        w = MJPEGWriter(path, fps=60)
        for f in frames:
            w.append_data(f)
        w.close()
        cap = cv2.VideoCapture(path)
        assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 5
        assert abs(cap.get(cv2.CAP_PROP_FPS) - 60) < 1e-3
        for f in frames:
            ok, got = cap.read()
            assert ok and got.shape == f.shape
            assert np.abs(got[..., ::-1].astype(int) - f.astype(int)).mean() < 6
        cap.release()


def test_interpolated_keyframes_are_periodic_and_hit_the_keys():
    from spi_b200.utils.video_utils import interpolate_keyframes
    ws = torch.randn(3, 14, 8, generator=torch.Generator().manual_seed(0))
    assert interpolate_keyframes(ws[:1], 120) is None                        # one keyframe: broadcast, no copies
    w = interpolate_keyframes(ws, 4)
    assert w.shape == (12, 14, 8)
    for k in range(3):
        torch.testing.assert_close(w[4 * k], ws[k], rtol=1e-5, atol=1e-5)


class _FakeG(torch.nn.Module):
    """Stands in for TriPlaneGenerator on the CPU: records how `synthesis` is called and paints the camera into the image."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self._last_planes = 'kept'
        self.calls = []

    def synthesis(self, ws, c, noise_mode='const', need_image=True, cache_backbone=False, use_cached_backbone=False, **kw):
        self.calls.append(dict(n=ws.shape[0], stride0=ws.stride(0), noise_mode=noise_mode, need_image=need_image,
                               cache=cache_backbone, use_cached=use_cached_backbone))
        if cache_backbone:
            self._last_planes = 'planes of this clip'
        n = c.shape[0]
        img = c[:, 3].reshape(n, 1, 1, 1).expand(n, 3, 16, 16).clone() / 3.0           # cam2world[0, 3]: the camera's x position
        return {'image': img if need_image else None, 'image_raw': img[:, :, ::2, ::2], 'image_depth': img[:, :1, ::2, ::2] + torch.rand(n, 1, 8, 8)}


def test_render_orbit_uses_one_backbone_pass_and_broadcast_latents():
    from spi_b200.utils.video_utils import render_orbit
    G = _FakeG()
    ws = torch.randn(1, 14, 512)
    frames, poses = render_orbit(G, ws, w_frames=10, batch=4)
    assert frames.shape == (10, 16, 16, 3) and frames.dtype == torch.uint8 and poses.shape == (10, 4, 4)
    assert [c['n'] for c in G.calls] == [4, 4, 2]
    assert all(c['stride0'] == 0 and c['noise_mode'] == 'const' and c['need_image'] for c in G.calls)        # a broadcast view, never a copy
    assert [(c['cache'], c['use_cached']) for c in G.calls] == [(True, False), (False, True), (False, True)]
    assert G._last_planes == 'kept'                                        # the caller's cached planes are put back
    assert len({int(f[0, 0, 0]) for f in frames}) > 3                      # the orbit moves
    G.calls.clear()
    depth, _ = render_orbit(G, ws, w_frames=3, batch=8, image_mode='image_depth')
    assert depth.shape == (3, 8, 8, 3) and not G.calls[0]['need_image']
    assert int(depth.amin()) == 0 and int(depth.amax()) == 255
    G.calls.clear()
    two = torch.randn(2, 14, 512)                                          # two keyframes: interpolated latents, no backbone sharing
    frames, _ = render_orbit(G, two, w_frames=3, batch=4)
    assert frames.shape[0] == 6 and all(not c['cache'] and not c['use_cached'] and c['stride0'] != 0 for c in G.calls)


def test_coach_tail_writes_checkpoint_stills_clip_and_metrics():
    """`BaseCoach.finish_image` + `log_metric` with logging on (pti_coach.py:87-98): files of the reference's output tree."""
    cv2 = pytest.importorskip('cv2')
    from spi_b200.configs import global_config, hyperparameters, paths_config
    from spi_b200.training.coaches.base_coach import BaseCoach
    keys = ('checkpoints_dir', 'embedding_base_dir', 'experiments_output_dir', 'images_output_dir', 'mirror_images_output_dir', 'video_output_dir')
    saved = {k: getattr(paths_config, k) for k in keys}
    saved_step, saved_dev = hyperparameters.G_1_step, global_config.device

    class M:
        def run(self, gt, fake):
            return float((gt - fake).square().mean()), 0.25, 0.5
    try:
        with tempfile.TemporaryDirectory() as d:
            for k in keys:
                setattr(paths_config, k, os.path.join(d, k) + '/')
                os.makedirs(os.path.join(d, k, 'Coach_x'))
            hyperparameters.G_1_step, global_config.device = 3, 'cpu'
            coach = object.__new__(BaseCoach)
            coach.G, coach.use_wandb, coach.coach_name, coach.metric_dic, coach._metric = _FakeG(), True, 'Coach_x', {}, M()
            from spi_b200.utils.camera_utils import cal_canonical_c
            w, c, image = torch.randn(1, 14, 512), cal_canonical_c(0.3, 0, 1, 'cpu'), torch.rand(1, 3, 16, 16) * 2 - 1
            paths_config.experiments_output_dir = os.path.join(d, 'experiments_output_dir', 'Coach_x', 'img0')
            os.makedirs(paths_config.experiments_output_dir)
            coach.finish_image(w, image, c, 'img0')
            exp = paths_config.experiments_output_dir
            assert sorted(os.listdir(exp)) == ['img0_G1_inv.avi', 'img0_G1_inv.jpg', 'img0_G1_inv_m.jpg']
            assert os.path.isfile(os.path.join(d, 'images_output_dir', 'Coach_x', 'img0.jpg'))
            assert os.path.isfile(os.path.join(d, 'mirror_images_output_dir', 'Coach_x', 'img0.jpg'))
            clip = os.path.join(d, 'video_output_dir', 'Coach_x', 'img0.avi')
            cap = cv2.VideoCapture(clip)
            assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 120          # base_coach.py:236-237: the default 120-frame orbit
            cap.release()
            ckpt = torch.load(os.path.join(d, 'checkpoints_dir', 'Coach_x', 'img0.pt'))
            assert sorted(ckpt) == ['G', 'c', 'w'] and torch.equal(ckpt['w'], w)
            assert list(coach.metric_dic) == ['G1_inv'] and len(coach.metric_dic['G1_inv']['l2_m']) == 1
            paths_config.experiments_output_dir = os.path.join(d, 'experiments_output_dir', 'Coach_x')
            coach.log_metric()
            text = open(os.path.join(paths_config.experiments_output_dir, 'metric_log.txt')).read()
            assert text.startswith('Coach name: Coach_x\n') and 'Mode: G1_inv AVG' in text and 'Lpips: 0.250000; ID Sim: 0.500000' in text
    finally:
        for k, v in saved.items():
            setattr(paths_config, k, v)
        hyperparameters.G_1_step, global_config.device = saved_step, saved_dev


def test_inference_coach_reloads_checkpoints_and_writes_the_clips():
    """`--G_1_type Inference` (inference_coach.py:18-46): for every dataset item the checkpoint {w, c, G} written by the coach named in
    `load_embedding_coach_name` is reloaded into G and its 120-frame orbit clip goes to `video_output_dir/<name>.mp4|.avi`."""
    cv2 = pytest.importorskip('cv2')
    from spi_b200.configs import global_config, hyperparameters, paths_config
    from spi_b200.training.coaches.base_coach import BaseCoach
    from spi_b200.training.coaches.inference_coach import InferenceCoach
    from spi_b200.utils.camera_utils import cal_canonical_c
    keys = ('checkpoints_dir', 'embedding_base_dir', 'experiments_output_dir', 'images_output_dir', 'mirror_images_output_dir', 'video_output_dir')
    saved = {k: getattr(paths_config, k) for k in keys}
    saved_hp = (hyperparameters.load_embedding_coach_name, hyperparameters.max_images_to_invert, global_config.device)
    try:
        with tempfile.TemporaryDirectory() as d:
            for k in keys:
                setattr(paths_config, k, os.path.join(d, k) + '/')
                os.makedirs(os.path.join(d, k, 'Coach_src'))
            global_config.device = 'cpu'
            hyperparameters.load_embedding_coach_name, hyperparameters.max_images_to_invert = 'Coach_src', 10
            writer = object.__new__(BaseCoach)
            src = _FakeG()
            with torch.no_grad():
                src.p.fill_(3.5)
            names = ['a0', 'b1']
            for i, nm in enumerate(names):
                writer.save(torch.full((1, 14, 512), float(i)), cal_canonical_c(0.3 + 0.1 * i, 0, 1, 'cpu'), src, os.path.join(d, 'checkpoints_dir', 'Coach_src', nm + '.pt'))
            coach = object.__new__(InferenceCoach)
            coach.G, coach.use_wandb, coach.coach_name, coach.image_counter = _FakeG(), False, 'InferenceCoach_x', 0
            coach.data_loader = [{'name': [nm]} for nm in names]
            clips = coach.train()
            assert float(coach.G.p) == 3.5                                      # the checkpoint's generator state was loaded
            assert coach.image_counter == 2 and len(clips) == 2
            for nm, clip in zip(names, clips):
                assert os.path.dirname(clip).rstrip('/') == os.path.join(d, 'video_output_dir') and os.path.basename(clip).startswith(nm + '.')
                cap = cv2.VideoCapture(clip)
                assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 120
                cap.release()
                assert os.path.isdir(os.path.join(d, 'experiments_output_dir', 'InferenceCoach_x', nm))
    finally:
        for k, v in saved.items():
            setattr(paths_config, k, v)
        hyperparameters.load_embedding_coach_name, hyperparameters.max_images_to_invert, global_config.device = saved_hp
