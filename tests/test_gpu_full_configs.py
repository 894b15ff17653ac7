"""Oracle parity at BASELINE.json's FULL configurations (SURVEY.md section 8 sizes)

    A  b4 t15 c10 200x200   (bench.py's workload, configs[1])
    B  b1 t25 c20 200x200   (configs[2])
    C  b1 t30 c32 256x256   (largest point of the configs[4] sweep)

for every kernel family that can serve the size: the persistent loops, image strides and the tail hand-over between
the whole-image and the half-split kernel only show at these image counts (600 / 500 / 960 coil images per launch).
The numpy fp64 oracle is the arbiter; tolerance 1e-5 of max|ref| (north_star)."""
import numpy as np
import pytest
import torch

import make_golden as G
from oracle import sense_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
CONFIGS = {"A": (4, 15, 10, 200, 200), "B": (1, 25, 20, 200, 200), "C": (1, 30, 32, 256, 256)}
FAMILIES = ["auto", "half", "packed"]        # ops.set_fused_path


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(out, ref):
    out = out.detach().cpu().numpy().astype(np.float64)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    return float(np.abs(out - ref).max() / max(np.abs(ref).max(), 1e-30))


_cache = {}


def case(tag):
    """Inputs (fp32) and the fp64 oracle outputs for one configuration (computed once per session)."""
    if tag in _cache:
        return _cache[tag]
    _cache.clear()                                    # one configuration resident at a time (C is 0.5 GB per tensor)
    b, t, c, h, w = CONFIGS[tag]
    cs = G.sense_case(900 + c, b, t, c, h, w)
    d = {k: (np.asarray(v, np.float64) if getattr(v, "dtype", None) == np.float32 and v.ndim else v) for k, v in cs.items()}
    v = float(O.softplus(cs["lam"]))
    kx = O.sens_expand(d["img"], d["sens"])
    want = {
        "expand": kx,
        "expand_mask": O.apply_mask(kx, d["mask"]),
        "expand_dc": O.dc_blend(kx, d["ref"], d["mask"], v),
        "expand_res": O.apply_mask(kx, d["mask"]) - d["ref"],
        "reduce": O.sens_reduce(d["k"], d["sens"]),
        "reduce_mask": O.sens_reduce(O.apply_mask(d["k"], d["mask"]), d["sens"]),
        "fft2c": O.fft2c(d["k"][:, :2]),
        "ifft2c": O.ifft2c(d["k"][:, :2]),
    }
    if h in (200, 256):
        want["normal"] = O.normal_op(d["img"], d["mask"], d["sens"], v)
    _cache[tag] = (cs, v, want)
    return _cache[tag]


@pytest.fixture
def family(request):
    from deep_cine_cardiac_mri_b200 import ops
    ops.set_fused_path(request.param)
    yield request.param
    ops.set_fused_path(None)


@pytest.mark.parametrize("tag", ["A", "B", "C"])
@pytest.mark.parametrize("family", FAMILIES, indirect=True)
def test_operators_at_full_config(tag, family):
    from deep_cine_cardiac_mri_b200 import ops
    if tag == "C" and family != "auto":
        pytest.skip("256 x 256 has one kernel family (quarter split)")
    cs, v, want = case(tag)
    b, t, c, h, w = CONFIGS[tag]
    img, sens, k, ref = cu(cs["img"]), cu(cs["sens"]), cu(cs["k"]), cu(cs["ref"])
    mask = cu(cs["mask"])
    vd = torch.tensor([v], device="cuda")
    assert rel(ops.sens_expand(img, sens), want["expand"]) <= TOL
    assert rel(ops.sens_expand(img, sens, ops.EXPAND_MASK, mask=mask), want["expand_mask"]) <= TOL
    assert rel(ops.sens_expand(img, sens, ops.EXPAND_DC, ref=ref, mask=mask, v=vd), want["expand_dc"]) <= TOL
    assert rel(ops.sens_expand(img, sens, ops.EXPAND_RESIDUAL, ref=ref, mask=mask), want["expand_res"]) <= TOL
    assert rel(ops.sens_reduce(k, sens).unsqueeze(2), want["reduce"]) <= TOL
    assert rel(ops.sens_reduce(k, sens, mask=mask).unsqueeze(2), want["reduce_mask"]) <= TOL
    assert rel(ops.fft2c(k[:, :2].contiguous(), "ortho"), want["fft2c"]) <= TOL
    assert rel(ops.fft2c(k[:, :2].contiguous(), "ortho", inverse=True), want["ifft2c"]) <= TOL
    if "normal" in want and family == "auto":
        assert rel(ops.normal_op(img.squeeze(2), sens, mask, vd).unsqueeze(2), want["normal"]) <= TOL
    # deterministic coil sum at full size: bit-identical from run to run and equal to the oracle
    ops.set_deterministic(True)
    try:
        r1 = ops.sens_reduce(k, sens)
        r2 = ops.sens_reduce(k, sens)
        assert torch.equal(r1, r2)
        assert rel(r1.unsqueeze(2), want["reduce"]) <= TOL
    finally:
        ops.set_deterministic(False)


def test_two_stream_deterministic_and_generic_sizes():
    """The scratch buffer of the deterministic coil sum / of shapes without a fused plan is per call: two streams
    running the hot path concurrently (pipeline.varnet_hot_path_streams) must reproduce the single-stream result
    (a per-device scratch cache raced here: ADVICE r1)."""
    from deep_cine_cardiac_mri_b200 import ops, pipeline, synth
    for (b, t, c, h, w), det in (((4, 5, 6, 200, 200), True), ((4, 4, 5, 128, 128), False), ((4, 4, 5, 128, 128), True)):
        cases = [synth.cine_case(300 + i, 1, t, c, h, w) for i in range(b)]
        mk = torch.from_numpy(np.concatenate([q["masked_kspace"] for q in cases], 0)).cuda()
        mask = torch.from_numpy(np.concatenate([q["mask"] for q in cases], 0)).cuda()
        v = torch.ones(1, device="cuda")
        ops.set_deterministic(det)
        try:
            with torch.no_grad():
                one = pipeline.varnet_hot_path(mk, mask, v, 3, xf=True)
                for _ in range(3):
                    two = pipeline.varnet_hot_path_streams(mk, mask, v, 3, xf=True, n_streams=2)
                    torch.cuda.synchronize()
                    if det:
                        assert torch.equal(one, two), (h, det)
                    else:
                        assert float((one - two).abs().max()) <= 1e-5 * float(one.abs().max()), (h, det)
        finally:
            ops.set_deterministic(False)
