"""CPU tests of the host logic either side of the hot path, against outputs of the reference's own classes
(`oracle/make_golden_post.py` -> tests/golden/host_logic.json): `PTIDataset` item formats and slicing rules
(spi/data/images_dataset.py:102-198), coach naming + output tree (base_coach.py:240-270), the CLI surface
(spi/run_inversion.py:18-56) and the checkpoint format (base_coach.py:204-217)."""
import json
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from oracle import toy_dataset

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'host_logic.json')))


def _names(ds):
    return [os.path.dirname(p).split('/')[-1] for p in ds.source_paths]


@pytest.fixture(scope='module')
def toy_root():
    with tempfile.TemporaryDirectory() as d:
        toy_dataset.write(d, n=7)
        yield d


def _kw(d):
    return dict(source_root=os.path.join(d, 'crop'), c_root=os.path.join(d, 'c'), mask_root=os.path.join(d, 'mask'),
                lm_root=os.path.join(d, 'lm'), mode='png')


def test_dataset_items_equal_the_reference_items(toy_root):
    from spi_b200.data.images_dataset import PTIDataset
    ds = PTIDataset(**_kw(toy_root))
    assert len(ds) == 7
    for gold, i in zip(GOLD['dataset']['items'], (0, 6)):
        got = toy_dataset.digest(ds[i])
        for k in ('img_shape', 'c_dtype', 'mask_sum', 'mask_dtype', 'mask_shape', 'lm_dtype', 'name', 'fname', 'c', 'lm'):
            assert got[k] == gold[k], k
        np.testing.assert_array_equal(np.float32(got['img_grid']), np.float32(gold['img_grid']))      # ToTensor + Normalize(0.5, 0.5), bit for bit
        assert got['img_sum'] == gold['img_sum']


def test_dataset_slicing_rules(toy_root):
    from spi_b200.data.images_dataset import PTIDataset
    kw = _kw(toy_root)
    g = GOLD['dataset']
    for block, names in g['blocks'].items():
        assert _names(PTIDataset(dataset_block=block, **kw)) == names, block
    assert _names(PTIDataset(select_range=3, **kw)) == g['select_3']
    assert _names(PTIDataset(filter_index=['00004', '00001'], **kw)) == g['filter']
    with tempfile.TemporaryDirectory() as out:
        for nm in ('00000', '00003'):
            open(os.path.join(out, nm + '.jpg'), 'w').close()
        assert _names(PTIDataset(output_root=out, **kw)) == g['resume']                           # finished images are skipped
        assert _names(PTIDataset(output_root=out, dataset_block='2/2', **kw)) == g['resume_block_2_2']


def test_every_image_lands_on_exactly_one_rank(toy_root):
    """The blocks of `len // W + 1` images cover the list without overlap for any world size (trailing ranks may idle)."""
    from spi_b200.data.images_dataset import PTIDataset
    kw = _kw(toy_root)
    for world in (1, 2, 3, 4, 8):
        got = sum((_names(PTIDataset(dataset_block=f'{r + 1}/{world}', **kw)) for r in range(world)), [])
        assert got == [f'{i:05d}' for i in range(7)], world


def test_coach_names_and_output_tree():
    from spi_b200.configs import hyperparameters, paths_config
    from spi_b200.training.coaches.base_coach import BaseCoach
    saved_hp = {k: getattr(hyperparameters, k) for k in GOLD['coach_names'][0]['hyperparameters']}
    keys = ('checkpoints_dir', 'embedding_base_dir', 'experiments_output_dir', 'images_output_dir', 'mirror_images_output_dir', 'video_output_dir')
    saved_paths = {k: getattr(paths_config, k) for k in keys}
    try:
        for case in GOLD['coach_names']:
            with tempfile.TemporaryDirectory() as d:
                for k in keys:
                    setattr(paths_config, k, os.path.join(d, k) + '/')
                for k, v in case['hyperparameters'].items():
                    setattr(hyperparameters, k, v)
                fake = types.SimpleNamespace(coach_name='RotBboxCoach')
                BaseCoach.build_name(fake)
                assert fake.coach_name == case['coach_name']
                dirs = sorted(os.path.relpath(os.path.join(r, x), d) for r, dd, _ in os.walk(d) for x in dd)
                assert dirs == case['dirs']
    finally:
        for k, v in saved_hp.items():
            setattr(hyperparameters, k, v)
        for k, v in saved_paths.items():
            setattr(paths_config, k, v)


# flag -> (default, type) of spi/run_inversion.py:18-42
REFERENCE_FLAGS = {
    'data_root': 'test/dataset/', 'data_mode': 'png', 'output_root': None, 'use_encoder': False, 'use_G_avg': False,
    'use_adapt_yaw_range': False, 'not_use_wandb': False, 'first_inv_type': 'pti', 'first_inv_steps': 500, 'G_1_step': 500,
    'G_1_type': 'space', 'G_2_step': 500, 'load_embedding_coach_name': None, 'pt_rot_lambda': 0, 'pt_mirror_rot_lambda': 0,
    'pt_depth_lambda': 0, 'pt_tv_lambda': 0, 'description': None, 'dataset_block': None, 'select_range': None, 'filter_index': None}


def test_cli_surface_defaults_and_config_globals():
    from spi_b200 import run_inversion
    from spi_b200.configs import hyperparameters, paths_config
    keys = ('root', 'checkpoints_dir', 'embedding_base_dir', 'experiments_output_dir', 'images_output_dir', 'mirror_images_output_dir',
            'video_output_dir', 'EG3D_PATH')
    saved_paths = {k: getattr(paths_config, k) for k in keys}
    saved_hp = dict(vars(hyperparameters))
    try:
        args = run_inversion.parse_args([])
        for flag, default in REFERENCE_FLAGS.items():
            assert getattr(args, flag) == default, flag
        with tempfile.TemporaryDirectory() as d:
            out = d + '/'
            args = run_inversion.parse_args(['--output_root', out, '--first_inv_type', 'mir', '--first_inv_steps', '7', '--G_1_type', 'RotBbox',
                                             '--G_1_step', '9', '--pt_rot_lambda', '0.1', '--pt_mirror_rot_lambda', '0.05', '--pt_depth_lambda',
                                             '1', '--not_use_wandb', '--dataset_block', '2/8', '--filter_index', '3,4'])
            assert (hyperparameters.first_inv_type, hyperparameters.first_inv_steps, hyperparameters.G_1_type, hyperparameters.G_1_step) == ('mir', 7, 'RotBbox', 9)
            assert (hyperparameters.pt_rot_lambda, hyperparameters.pt_mirror_rot_lambda, hyperparameters.pt_depth_lambda) == (0.1, 0.05, 1.0)
            assert sorted(os.listdir(d)) == ['checkpoints', 'embedding', 'experiments', 'image', 'image_m', 'video']     # run_inversion.py:60-79
            assert paths_config.video_output_dir == out + 'video/' and args.dataset_block == '2/8'
        # the parser defaults are not runnable, as in the reference (run_inversion.py:119-120): both flags are effectively required
        with tempfile.TemporaryDirectory() as d, pytest.raises(NotImplementedError):
            toy_dataset.write(d, n=1)
            run_inversion.run(['--not_use_wandb', '--data_root', d])
    finally:
        for k, v in saved_paths.items():
            setattr(paths_config, k, v)
        for k, v in saved_hp.items():
            if not k.startswith('__'):
                setattr(hyperparameters, k, v)


def test_checkpoint_format_round_trip():
    """base_coach.py:204-217: {'w', 'c', 'G': state_dict} on the CPU."""
    from spi_b200.training.coaches.base_coach import BaseCoach
    G = torch.nn.Linear(3, 2)
    w, c = torch.randn(1, 14, 512), torch.randn(1, 25)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'x.pt')
        BaseCoach.save(None, w, c, G, path)
        ckpt = torch.load(path, map_location='cpu')
        assert sorted(ckpt) == ['G', 'c', 'w'] and sorted(ckpt['G']) == ['bias', 'weight']
        assert torch.equal(ckpt['w'], w) and torch.equal(ckpt['c'], c) and torch.equal(ckpt['G']['weight'], G.weight.detach())
        holder = types.SimpleNamespace(G=torch.nn.Linear(3, 2))
        from spi_b200.configs import global_config
        dev, global_config.device = global_config.device, 'cpu'
        try:
            w2, c2, G2 = BaseCoach.load(holder, path)
        finally:
            global_config.device = dev
        assert torch.equal(w2, w) and torch.equal(c2, c) and torch.equal(G2.weight, G.weight)


def test_product_camera_helpers_against_reference_goldens(golden):
    """The product's camera helpers are device-agnostic host logic: on CPU tensors they must reproduce what the reference's
    spi/utils/camera_utils.py produced for the same uniform draws (tests/golden/geometry_losses.npz)."""
    from oracle import weights
    from spi_b200.utils import camera_utils as cu, rng
    from conftest import rel_l2
    g = golden('geometry_losses')
    c = weights.canonical_camera(0.3)
    r = torch.from_numpy(g['rand42'])
    rng.inject(r[:, 0:1].clone(), r[:, 1:2].clone())
    assert rel_l2(cu.sample_surrounding_camera(c, batch_size=4, yaw_range=0.2, pitch_range=0.1), g['surround']) < 1e-6
    rng.inject(r[:, 0:1].clone(), r[:, 1:2].clone())
    assert rel_l2(cu.sample_camera(4, yaw_range=0.7, pitch_range=0.4, device='cpu'), g['sampled']) < 1e-6
    assert rng.pending() == 0
    assert rel_l2(cu.cal_camera_weight(c), g['cam_weight']) < 1e-6
    assert rel_l2(cu.cal_canonical_c(0.3, 0, 1, 'cpu'), c) < 1e-6
    # |yaw| < 0.2: the mirror branch is switched off (camera_utils.py:407-408)
    assert float(cu.cal_camera_weight(cu.cal_canonical_c(0.1, 0, 1, 'cpu'))) == 0.0
    # mirror pose: entries [0,1],[0,2],[0,3],[1,0],[2,0] negated, an involution, yaw changes sign
    m = cu.cal_mirror_c(c)
    d = (m[:, :16].view(4, 4) != c[:, :16].view(4, 4)).nonzero().tolist()
    assert set(map(tuple, d)) <= {(0, 1), (0, 2), (0, 3), (1, 0), (2, 0)} and torch.equal(cu.cal_mirror_c(m), c)
    yaw = lambda cam: float(cu.rotation_to_angle(cam.view(25)[:16].view(4, 4)[:3, :3])[0])
    assert abs(yaw(m) + yaw(c)) < 1e-6 and abs(abs(yaw(c)) - 0.3) < 5e-2       # the orbit centre is the look-at point [0, 0, 0.2], not the origin


def test_projector_schedule_all_steps():
    """LR / w-noise schedule of the stage-1 projectors (mirror_projector.py:84-91): the product's host code against the oracle's and
    against the closed form, at every step of the 500-step run; fixed points: lr = 0 at step 0, peak lr 0.01 from 5 % to 75 %
    (`initial_learning_rate`, not `first_inv_lr`), noise off after 75 %."""
    import types
    from oracle import loops
    from spi_b200.training.projectors._common import LatentProjector
    fake = types.SimpleNamespace(num_steps=500, w_std=9.7, hp=dict(lr0=0.01, noise0=0.05, down=0.25, up=0.05, nramp=0.75, regw=1e5))
    for step in range(500):
        lr, wn = LatentProjector.schedule(fake, step)
        lr_o, wn_o = loops.lr_schedule(step, 500, 9.7)
        t = step / 500
        ramp = (0.5 - 0.5 * np.cos(min(1.0, (1.0 - t) / 0.25) * np.pi)) * min(1.0, t / 0.05)
        assert lr == lr_o == 0.01 * ramp and wn == wn_o == 9.7 * 0.05 * max(0.0, 1.0 - t / 0.75) ** 2
    assert LatentProjector.schedule(fake, 0)[0] == 0.0
    assert all(LatentProjector.schedule(fake, s)[0] == 0.01 for s in (25, 100, 375))
    assert LatentProjector.schedule(fake, 376)[0] < 0.01 and LatentProjector.schedule(fake, 375)[1] == 0.0


def test_bench_schedule_accepts_any_step_count():
    """bench.py must run under whatever `--steps K --warmup W` a harness passes (round 1 asserted K % 12 == 0 and produced
    nothing under `--steps 20 --warmup 5`): K iterations, ~1/3 mir, ~1/4 of the RotBbox iterations heavy, consecutive loop indices."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for k in (1, 2, 5, 12, 20, 24, 37):
        s = bench.schedule(k)
        assert len(s) == k
        n_mir, heavy, light = bench.mix_of(s)
        assert n_mir + heavy + light == k and abs(n_mir - k / 3) <= 0.67
        rot = [i for kind, i in s if kind == 'rot']
        assert rot == list(range(rot[0], rot[0] + len(rot))) if rot else True
        assert abs(heavy - len(rot) / 4) <= 0.5
        inter = bench.interleave(s)
        assert sorted(inter) == sorted(s)
        if k >= 12:             # every prefix of the interleaved order keeps the mix (the CPU arm may be cut by its budget)
            half = inter[:k // 2]
            hm, hh, hl = bench.mix_of(half)
            assert abs(hm - n_mir / 2) <= 1 and abs(hh - heavy / 2) <= 1
    assert bench.mix_of(bench.schedule(24)) == (8, 4, 12) and bench.mix_of(bench.schedule(20)) == (7, 3, 10)


def test_zero_arena_sizing_and_fallback():
    """ops/zero_arena.py host logic on CPU tensors: the first iteration of a kind only records its requests, the second gets one zeroed
    buffer of exactly that size, slices are 1 KiB-aligned, disjoint and shaped as asked, an over-ask returns None."""
    from spi_b200.ops import zero_arena as Z
    Z.reset()
    cpu = torch.device('cpu')
    assert Z.take((4, 4)) is None                       # outside an iteration
    with Z.iteration('k', cpu):
        assert not Z.active() and Z.take((3, 5)) is None and Z.take((2, 8, 4, 4), channels_last=True) is None
    assert Z._hint['k'] == 1024 + 1024
    with Z.iteration('k', cpu):
        assert Z.active()
        a = Z.take((3, 5))
        b = Z.take((2, 8, 4, 4), channels_last=True)
        c = Z.take((1,))
        assert a.shape == (3, 5) and a.dtype == torch.float32 and float(a.abs().sum()) == 0
        assert b.shape == (2, 8, 4, 4) and b.is_contiguous(memory_format=torch.channels_last)
        assert b.data_ptr() - a.data_ptr() == 1024 and c is None
    assert Z._hint['k'] == 3 * 1024                      # the over-ask is remembered: the next iteration's buffer has room for it
    with Z.iteration('other', cpu):
        assert not Z.active()
    Z.reset()


def test_modulation_plan_and_bank_rules(monkeypatch):
    """Host-side decisions of the grouped modulation (no kernels): layout / tap reversal / transposed copy per layer kind and engine,
    and which layer sets the grouped kernels accept."""
    from spi_b200.ops import conv as E
    from spi_b200.ops.modulate import bank_usable
    from spi_b200.training.networks_stylegan2 import modulation_plan
    w33 = torch.empty(128, 256, 3, 3)
    w11 = torch.empty(96, 512, 1, 1)
    wrgb = torch.empty(3, 128, 1, 1)
    monkeypatch.setattr(E, 'ENGINE', 'tc2')
    assert modulation_plan(w33, 1, True) == ('ohwi', False, 'rev')            # non-resampling layer: taps as stored, reversed copy for the data gradient
    assert modulation_plan(w33, 1, False) == ('ohwi', True, 'rev')            # flip_weight=False: modulate writes the taps reversed
    assert modulation_plan(w33, 2, False) == ('ohwi', False, 'keep')          # up-sampling layer on the engine: [O][taps][I], plain transposed copy
    assert modulation_plan(w11, 1, True) == ('ohwi', False, 'rev')
    assert modulation_plan(wrgb, 1, True) == ('ohwi', False, None)            # RGB head (3 channels): streaming kernels, no transposed copy
    monkeypatch.setattr(E, 'ENGINE', 'cudnn')
    assert modulation_plan(w33, 2, False) == ('ihwo', False, None)            # the library's transposed convolution wants [I][taps][O]
    assert modulation_plan(w33, 1, True) == ('ohwi', False, None)
    # bank_usable: CUDA fp32 only (a CPU tensor is refused without touching the library), layouts OIHW / OHWI, rows that fit in shared memory
    s = torch.empty(1, 256)
    assert not bank_usable([(w33, s, True, 'ohwi', False, 'rev')])
    assert not bank_usable([])
