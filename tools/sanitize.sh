#!/bin/bash
# compute-sanitizer passes over the smallest shapes (run through gpurun); logs land in gpurun_out/.
mkdir -p gpurun_out
cat > /tmp/san_target.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from deep_cine_cardiac_mri_b200 import ops, metrics
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
for (b, t, c, h, w) in ((1, 2, 2, 200, 200), (1, 1, 2, 256, 256), (1, 2, 3, 18, 14)):
    k = torch.randn(b, t, c, h, w, 2, device=dev, generator=g); ref = torch.randn_like(k)
    s = torch.randn(b, c, h, w, 2, device=dev, generator=g); x = torch.randn(b, t, h, w, 2, device=dev, generator=g)
    m = (torch.rand(b, t, h, device=dev, generator=g) < 0.3).to(torch.uint8); v = torch.tensor([0.7], device=dev)
    for _ in range(2):                      # twice: the persistent loop's item boundary and re-launch
        ops.raw_fft2c(k, False, 1); ops.raw_fft2c(k, True, 1)
        img = ops.raw_sens_reduce(k, s); ops.raw_sens_reduce(k, s, 1, False, m, v); ops.raw_sens_reduce(k, x, 2, True, m, v)
        for mode in (0, 1, 2, 3):
            ops.raw_sens_expand(img, s, mode, ref, m, v)
        if ops.normal_op_supported(h, w):
            ops.raw_normal_op(x, s, m, v)
            ops.raw_normal_dc(x, s, m, v, s.pow(2).sum(dim=(1, 4)).contiguous(), img)
            ops.raw_normal_dc(x, s, m, v, s.pow(2).sum(dim=(1, 4)).contiguous(), img, magnitude=True)
            from deep_cine_cardiac_mri_b200 import blocks
            blocks._cg_inference(x, img, m.view(b, t, 1, h, 1, 1), s, v, 2)     # fused CG iteration (normal_op_dot, cg_update, cg_direction)
        for fam in ("half", "packed"):          # both kernel families of the fused plan sizes
            ops.set_fused_path(fam)
            ops.raw_sens_expand(img, s, 2, ref, m, v); ops.raw_sens_expand(img, s, 0, ref, m, v); ops.raw_sens_reduce(k, s)
        ops.set_fused_path(None)
        for norm, pad in ((True, True), (False, False)):   # regulariser-side plane glue
            xf, yf, ctx = ops.raw_planes_pack(x, norm, pad)
            ops.raw_planes_unpack(xf, yf, ctx)
    if ops.normal_op_supported(200, 36) and h == 200:
        ops.raw_normal_op(x[..., :36, :].contiguous(), s[..., :36, :].contiguous(), m, v)      # run-time width plan
    if h >= 7 and w >= 7:
        a = torch.rand(b, 1, t, h, w, device=dev, generator=g).requires_grad_(True); bb = torch.rand(b, 1, t, h, w, device=dev, generator=g)
        metrics.ssim_loss(a, bb).backward()
torch.cuda.synchronize(); print("sanitize target done")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
