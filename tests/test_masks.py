"""Undersampling masks (data/subsample.py) and apply_mask (data/transforms.py:66-92): SURVEY.md section 8f row 4."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import make_golden_masks as GM                                    # noqa: E402


def test_mask_functions_reproduce_the_reference_masks():
    """bit-identical to the masks the reference's own classes produced (committed fixture, made by make_golden_masks.py)"""
    from deep_cine_cardiac_mri_b200 import masks
    gold = np.load(ROOT / "tests" / "golden" / "golden_v3_masks.npz")
    for ci, (kind, cf, acc, shape) in enumerate(GM.CASES):
        for seed in GM.SEEDS:
            np.random.seed(seed)
            m = masks.create_mask_for_mask_type(kind, cf, acc)(shape, seed)
            assert m.dtype == torch.float32 and np.array_equal(m.numpy(), gold[f"c{ci}_s{seed}"]), (ci, seed)
    # explicit stream == global stream with the same seed; rows per frame as documented
    a = masks.RandomMaskFunc([10], [4], rng=np.random.RandomState(7))((15, 1, 200, 200, 2), 7)
    np.random.seed(7)
    b = masks.RandomMaskFunc([10], [4])((15, 1, 200, 200, 2), 7)
    assert torch.equal(a, b) and a.shape == (15, 1, 200, 1, 1)
    assert set(a.reshape(15, 200).sum(1).tolist()) == {50.0} and bool((a.reshape(15, 200)[:, 95:105] == 1).all())
    with pytest.raises(ValueError):
        masks.RandomMaskFunc([10], [4, 8])
    with pytest.raises(Exception, match="not supported"):
        masks.create_mask_for_mask_type("poisson", [10], [4])


def test_mask_functions_against_the_live_reference():
    from oracle import load_reference
    if not load_reference.available():
        pytest.skip("reference not shipped")
    import importlib
    load_reference.load(models=False)
    R = importlib.import_module("reconstruction.data.subsample")
    from deep_cine_cardiac_mri_b200 import masks
    for seed in (5, 99):
        shape = (15, 1, 200, 200, 2)
        np.random.seed(seed); a = R.RandomMaskFunc([10], [4])(shape, seed)
        np.random.seed(seed); b = masks.RandomMaskFunc([10], [4])(shape, seed)
        assert torch.equal(a, b)
        assert torch.equal(R.EquispacedMaskFunc([0.08], [4])(shape, seed), masks.EquispacedMaskFunc([0.08], [4])(shape, seed))


@pytest.mark.gpu
def test_apply_mask_on_the_gpu_is_bit_exact():
    from deep_cine_cardiac_mri_b200 import masks
    rng = np.random.default_rng(3)
    for shape in ((15, 10, 200, 200, 2), (4, 3, 18, 14, 2), (1, 6, 5, 256, 256, 2)):
        k = rng.standard_normal(shape, dtype=np.float32)
        kd = torch.from_numpy(k).cuda()
        np.random.seed(11)
        masked, mask = masks.apply_mask(kd, masks.RandomMaskFunc([4], [4]), seed=11)
        m = mask.numpy() if len(shape) == 5 else mask.numpy()[None]
        want = k * m + 0.0
        got = masked.cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want)
        assert not np.signbit(got[got == 0]).any()                 # the + 0.0 of transforms.py:90
        np.random.seed(11)
        masked2, m8 = masks.apply_mask_u8(kd, masks.RandomMaskFunc([4], [4]), seed=11)
        assert torch.equal(masked2, masked) and m8.dtype == torch.uint8 and m8.dim() == 6 and m8.shape[3] == shape[-3]
