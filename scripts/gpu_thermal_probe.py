"""Where the tensor-core mode's temperature offset comes from: ours (tc) against the fp32 oracle and against the
oracle with its field weights rounded to fp16, eval and training mode."""
import copy, sys
sys.path.insert(0, ".")
import torch
from oracle import OracleRays, make_synthetic_rays
from tests.helpers import make_pair
from thermo_nerf_b200 import _lib as L
from thermo_nerf_b200 import functional as F

for (lf, lp) in [(19, 17), (15, 12)]:
    oracle, model = make_pair(log2_field=lf, log2_prop=lp, num_images=100, trained_like=True, precision="tc_fp16",
                              camera_optimizer_mode="off")
    og = copy.deepcopy(oracle).to("cuda:0")
    og16 = copy.deepcopy(og)
    with torch.no_grad():
        for n, p in og16.field.named_parameters():
            if "hash_table" not in n and "embedding" not in n:
                p.copy_(p.half().float())
    R = 4096
    rays = make_synthetic_rays(R, num_images=100, seed=17)
    o, d, cam = rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda()
    jitter = torch.rand((3, R, 1), generator=torch.Generator().manual_seed(3)).cuda()
    for training in (False, True):
        with torch.no_grad():
            kw = dict(jitter=jitter) if training else {}
            ref = og.get_outputs(OracleRays(o, d, cam), training=training, **kw)
            ref16 = og16.get_outputs(OracleRays(o, d, cam), training=training, **kw)
            for prec in (L.PRECISION_TC_FP16, L.PRECISION_FP32):
                res = F.render_forward(model.tensors(), o, d, cam.reshape(-1), None, None, jitter.reshape(3, -1) if training else None,
                                       training=training, near_plane=0.05 if training else 0.0, far_plane=1000.0,
                                       appearance_mode=L.APPEARANCE_LOOKUP if training else L.APPEARANCE_MEAN, precision=prec)
                for name, r in (("fp32 oracle", ref), ("fp16-weights oracle", ref16)):
                    for k in ("thermal", "rgb"):
                        dd = res[k].reshape(r[k].shape) - r[k]
                        print(f"2^{lf} train={training} prec={prec} vs {name:20s} {k:8s}: mean {float(dd.mean()):+.2e} std {float(dd.std()):.2e} max {float(dd.abs().max()):.2e}  (image std {float(r[k].std()):.3f})")
