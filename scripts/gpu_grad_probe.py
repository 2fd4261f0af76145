"""Per-parameter and per-hash-level gradient error against the oracle's autograd run on the GPU."""
import copy, sys
sys.path.insert(0, ".")
import torch
from oracle import OracleRays, make_synthetic_rays
from tests.helpers import make_pair
from thermo_nerf_b200 import _lib as L
from thermo_nerf_b200 import functional as F

def run(log2_field, log2_prop, precision, R=4096, contrast=True):
    oracle, model = make_pair(log2_field=log2_field, log2_prop=log2_prop, num_images=100, trained_like=True,
                              precision=precision, thermal_contrast=contrast, camera_optimizer_mode="off")
    og = copy.deepcopy(oracle).to("cuda:0").train()
    og.anneal = 1.0
    rays = make_synthetic_rays(R, num_images=100, seed=17)
    g = torch.Generator().manual_seed(3)
    jitter = torch.rand((3, R, 1), generator=g).cuda()
    gt_rgb, gt_th = torch.rand((R, 3), generator=g).cuda(), torch.rand((R, 1), generator=g).cuda()
    o, d, cam = rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda()
    og.zero_grad()
    ref_out = og.get_outputs(OracleRays(o, d, cam), training=True, jitter=jitter)
    ref_ld = og.get_loss_dict(ref_out, gt_rgb, gt_th, training=True)
    sum(ref_ld.values()).backward()
    model.train(); model.zero_grad()
    prec = L.PRECISION_FP32 if precision == "fp32" else L.PRECISION_TC_FP16
    out = F.render(model.tensors(), o, d, cam.reshape(-1), None, None, jitter.reshape(3, -1), num_samples=(256, 96, 48),
                   near_plane=0.05, far_plane=1000.0, anneal=1.0, appearance_mode=L.APPEARANCE_LOOKUP, precision=prec)
    ld = F.losses(out, gt_rgb, gt_th)
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    print(f"== field 2^{log2_field} prop 2^{log2_prop} {precision} R={R} contrast={contrast}")
    print("  losses", {k: (float(ld[k]), float(ref_ld[k])) for k in ref_ld})
    ref_params = dict(og.named_parameters())
    for name, p in model.named_parameters():
        q = ref_params.get(name)
        if q is None or q.grad is None or p.grad is None:
            continue
        a, b = p.grad.flatten().double(), q.grad.flatten().double()
        if float(b.norm()) < 1e-12:
            continue
        rel = float((a - b).norm() / b.norm()); cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
        flag = " <<<" if rel > 3e-2 else ""
        print(f"  {name}: rel {rel:.3e} cos {cos:.5f} |ref| {float(b.norm()):.3e}{flag}")
    # per level of the field table
    a, b = model.field.mlp_base.encoder.hash_table.grad.double(), ref_params["field.mlp_base.encoder.hash_table"].grad.double()
    enc = oracle.field.mlp_base.encoder
    offs = getattr(enc, "offsets", None)
    if offs is None:
        n = a.shape[0] // 16
        offs = [i * n for i in range(17)]
    offs = [int(x) for x in offs]
    for l in range(len(offs) - 1):
        x, y = a[offs[l]:offs[l + 1]].flatten(), b[offs[l]:offs[l + 1]].flatten()
        print(f"    level {l}: entries {offs[l+1]-offs[l]} rel {float((x-y).norm()/y.norm().clamp_min(1e-30)):.3e} |ref| {float(y.norm()):.3e} |ours| {float(x.norm()):.3e} nnz ref {int((y!=0).sum())} ours {int((x!=0).sum())}")

run(19, 17, "tc_fp16")
run(19, 17, "fp32")
run(19, 17, "tc_fp16", contrast=False)
run(15, 12, "tc_fp16")
