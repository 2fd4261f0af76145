"""GPU parity tests of the training path: tnf_losses, tnf_render_backward (through the autograd node)
and tnf_adam_step against the CPU oracle's autograd / torch.optim.Adam on identical inputs.

Tolerances:
* losses: 1e-5 relative (fp32 both sides; summation order differs)
* gradients, precision="fp32": per-tensor rel-L2 error <= 2e-3 for the dense tensors (atomics reorder fp32
  sums; the oracle's autograd runs the same math on the CPU) and <= 1e-2 for the hash tables, the trunk
  MLP behind them (field.mlp_base.*) and the temperature MLP's first layer, which consumes the trunk's output: sample positions agree to ~1e-6 between the two implementations, which at
  the 2047-cell level moves the trilinear weights by ~1e-3 relative and occasionally moves a sample into the
  neighbouring cell; the test weights give the temperature head a gain of a few hundred on the geometry features
  (helpers.add_thermal_contrast), so that difference dominates the trunk's gradient (measured 4-6e-3)
* gradients, precision="tc_fp16" (fp16 forward operands, bf16 backward operands, fp32 accumulate):
  per-tensor rel-L2 <= 5e-2 and cosine >= 0.998 against the fp32 oracle
* Adam: 1e-6 relative after 5 steps
"""

import pytest
import torch

from oracle import OracleRays, make_synthetic_rays
from oracle import nerfstudio_math as M
from tests.helpers import make_pair

pytestmark = pytest.mark.gpu


def _train_pair(precision="fp32", R=192, seed=21, **kw):
    oracle, model = make_pair(trained_like=True, precision=precision, log2_field=12, log2_prop=10,
                              camera_optimizer_mode="off", **kw)
    rays = make_synthetic_rays(R, num_images=8, seed=seed)
    g = torch.Generator().manual_seed(seed)
    jitter = torch.rand((3, R, 1), generator=g)
    gt_rgb = torch.rand((R, 3), generator=g)
    gt_th = torch.rand((R, 1), generator=g)
    return oracle, model, rays, jitter, gt_rgb, gt_th


def _oracle_grads(oracle, rays, jitter, gt_rgb, gt_th, anneal, mults, prop_grad=True):
    oracle.anneal = anneal
    oracle.cfg.interlevel_loss_mult, oracle.cfg.distortion_loss_mult = mults
    oracle.zero_grad()
    out = oracle.get_outputs(rays, training=True, jitter=jitter, prop_grad=prop_grad)
    ld = oracle.get_loss_dict(out, gt_rgb, gt_th, training=True)
    sum(ld.values()).backward()
    return out, ld, {k: (p.grad.clone() if p.grad is not None else None) for k, p in oracle.named_parameters()}


def _ours(model, rays, jitter, gt_rgb, gt_th, anneal, mults, prop_grad=True):
    from thermo_nerf_b200 import functional as F
    from thermo_nerf_b200 import _lib as L

    model.train()
    model.zero_grad()
    prec = L.PRECISION_FP32 if model.config.precision == "fp32" else L.PRECISION_TC_FP16
    out = F.render(model.tensors(), rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda(),
                   None, None, jitter.cuda().reshape(3, -1), prop_grad=prop_grad, num_samples=(256, 96, 48),
                   near_plane=0.05, far_plane=1000.0, anneal=anneal, appearance_mode=L.APPEARANCE_LOOKUP,
                   precision=prec)
    ld = F.losses(out, gt_rgb.cuda(), gt_th.cuda(), interlevel_mult=mults[0], distortion_mult=mults[1])
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    return out, ld, {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in model.named_parameters()}


def _rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def test_losses_and_their_gradients_match_oracle():
    from thermo_nerf_b200 import functional as F

    oracle, model, rays, jitter, gt_rgb, gt_th = _train_pair()
    with torch.no_grad():
        out = oracle.get_outputs(rays, training=True, jitter=jitter)
    w = [x.detach().clone().requires_grad_(True) for x in out["weights_list"]]
    rgb = out["rgb"].detach().clone().requires_grad_(True)
    th = out["thermal"].detach().clone().requires_grad_(True)
    ref = {
        "rgb_loss": torch.nn.functional.mse_loss(gt_rgb, rgb),
        "interlevel_loss": 1.0 * M.interlevel_loss(w, out["sdist_list"]),
        "distortion_loss": 0.5 * M.distortion_loss(w, out["sdist_list"]),
        "thermal": torch.nn.functional.mse_loss(th, gt_th),
    }
    sum(ref.values()).backward()
    losses, g = F.losses_forward_backward([x.detach().cuda() for x in w], [s.cuda() for s in out["sdist_list"]],
                                          rgb.detach().cuda(), th.detach().cuda(), gt_rgb.cuda(), gt_th.cuda(),
                                          interlevel_mult=1.0, distortion_mult=0.5)
    torch.cuda.synchronize()
    for i, k in enumerate(F.LOSS_NAMES):
        assert losses[i].item() == pytest.approx(ref[k].item(), rel=1e-5, abs=1e-9), k
    assert _rel_l2(g["rgb"].cpu(), rgb.grad) <= 1e-5
    assert _rel_l2(g["thermal"].cpu(), th.grad.reshape(-1)) <= 1e-5
    for k in range(3):
        assert _rel_l2(g["weights_list"][k].cpu(), w[k].grad[..., 0]) <= 2e-4, k


@pytest.mark.parametrize("anneal", [1.0, 0.35])
def test_backward_fp32_matches_oracle_autograd(anneal):
    oracle, model, rays, jitter, gt_rgb, gt_th = _train_pair("fp32")
    mults = (1.0, 0.5)  # a distortion weight large enough for its gradient to matter in the comparison
    o_out, o_ld, o_g = _oracle_grads(oracle, rays, jitter, gt_rgb, gt_th, anneal, mults)
    out, ld, g = _ours(model, rays, jitter, gt_rgb, gt_th, anneal, mults)
    for k in o_ld:
        assert ld[k].item() == pytest.approx(o_ld[k].item(), rel=2e-3, abs=1e-7), k
    checked, bad = 0, []
    for k, ref in o_g.items():
        if k.startswith("camera_optimizer"):
            continue
        assert g[k] is not None, k
        if ref is None or ref.norm() == 0:
            assert g[k].abs().max().item() <= 1e-7, k
            continue
        err = _rel_l2(g[k], ref)
        tol = 1e-2 if k.endswith("hash_table") or k.startswith(("field.mlp_base", "field.mlp_thermal.layers.0")) else 2e-3
        print(f"{k}: rel-L2 {err:.2e} (tol {tol})")
        if err > tol:
            bad.append((k, err))
        checked += 1
    assert not bad, bad
    assert checked >= 27


def test_backward_without_proposal_update_step():
    """ProposalNetworkSampler `updated == False`: proposal nets get no gradient at all."""
    oracle, model, rays, jitter, gt_rgb, gt_th = _train_pair("fp32", R=64)
    _, _, o_g = _oracle_grads(oracle, rays, jitter, gt_rgb, gt_th, 1.0, (1.0, 0.002), prop_grad=False)
    _, _, g = _ours(model, rays, jitter, gt_rgb, gt_th, 1.0, (1.0, 0.002), prop_grad=False)
    for k, v in g.items():
        if k.startswith("proposal_networks"):
            assert v is None and o_g[k] is None, k
    # 64 rays touch few table entries, so the half-precision rounding of the features (the reference's encoder returns
    # fp16) weighs more in the relative error than in the 4096-ray comparisons above
    assert _rel_l2(g["field.mlp_base.encoder.hash_table"], o_g["field.mlp_base.encoder.hash_table"]) <= 1.5e-2
    assert _rel_l2(g["field.mlp_head.layers.1.weight"], o_g["field.mlp_head.layers.1.weight"]) <= 2e-3


def test_detached_thermal_gradients():
    """pass_thermal_gradients=False (thermal_field.py:173-175): the thermal head does not move the trunk."""
    oracle, model, rays, jitter, gt_rgb, gt_th = _train_pair("fp32", R=64, pass_thermal_gradients=False)
    from thermo_nerf_b200 import functional as F
    from thermo_nerf_b200 import _lib as L

    oracle.zero_grad()
    out = oracle.get_outputs(rays, training=True, jitter=jitter)
    torch.nn.functional.mse_loss(out["thermal"], gt_th).backward()
    ref = {k: p.grad for k, p in oracle.named_parameters()}
    model.train()
    model.zero_grad()
    o = F.render(model.tensors(), rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda(), None, None,
                 jitter.cuda().reshape(3, -1), near_plane=0.05, appearance_mode=L.APPEARANCE_LOOKUP,
                 precision=L.PRECISION_FP32, detach_thermal_geo=True)
    torch.nn.functional.mse_loss(o["thermal"], gt_th.cuda()).backward()
    torch.cuda.synchronize()
    g = {k: p.grad for k, p in model.named_parameters()}
    assert _rel_l2(g["field.mlp_thermal.layers.0.weight"].cpu(), ref["field.mlp_thermal.layers.0.weight"]) <= 2e-3
    assert _rel_l2(g["field.field_head_thermal.net.weight"].cpu(), ref["field.field_head_thermal.net.weight"]) <= 2e-3
    # the density still reaches the trunk through the weights; the oracle says how much
    assert _rel_l2(g["field.mlp_base.mlp.layers.1.weight"].cpu(), ref["field.mlp_base.mlp.layers.1.weight"]) <= 2e-3


def test_backward_tensor_core_matches_oracle_autograd():
    oracle, model, rays, jitter, gt_rgb, gt_th = _train_pair("tc_fp16", R=256)
    mults = (1.0, 0.5)
    _, o_ld, o_g = _oracle_grads(oracle, rays, jitter, gt_rgb, gt_th, 1.0, mults)
    _, ld, g = _ours(model, rays, jitter, gt_rgb, gt_th, 1.0, mults)
    for k in o_ld:
        assert ld[k].item() == pytest.approx(o_ld[k].item(), rel=3e-2, abs=1e-5), k
    bad = []
    for k, ref in o_g.items():
        if k.startswith("camera_optimizer") or ref is None or ref.norm() == 0:
            continue
        err = _rel_l2(g[k], ref)
        cos = torch.nn.functional.cosine_similarity(g[k].flatten(), ref.flatten(), dim=0).item()
        print(f"{k}: rel-L2 {err:.2e} cos {cos:.5f}")
        if not (err <= 5e-2 and cos >= 0.998):
            bad.append((k, err, cos))
    assert not bad, bad


def test_adam_matches_torch_adam():
    from thermo_nerf_b200 import FusedAdam

    g = torch.Generator().manual_seed(3)
    shapes = [(5 * 1024 + 3, 2), (64, 63), (1,), (4096 * 3,)]
    ps = [torch.randn(s, generator=g).cuda() for s in shapes]
    a = [p.clone().requires_grad_(True) for p in ps]
    b = [p.clone().requires_grad_(True) for p in ps]
    oa = FusedAdam(a, lr=1e-2, eps=1e-15)
    ob = torch.optim.Adam(b, lr=1e-2, eps=1e-15)
    for step in range(5):
        for x, y in zip(a, b):
            gr = torch.randn(x.shape, generator=g).cuda() * (10.0 ** (step - 3))
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-7), (x - y).abs().max()
        assert torch.allclose(oa.state[x]["exp_avg_sq"], ob.state[y]["exp_avg_sq"], rtol=1e-6, atol=1e-12)


def test_adam_grad_scaler_protocol_and_skip():
    from thermo_nerf_b200 import functional as F

    p = torch.ones(5000, device="cuda")
    gr = torch.full((5000,), 64.0, device="cuda")
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    F.adam_step([p], [gr], [m], [v], [0.1], step=1, grad_scale=torch.tensor([64.0], device="cuda"),
                found_inf=torch.zeros(1, device="cuda"), zero_grads=True)
    torch.cuda.synchronize()
    assert torch.allclose(p, torch.full_like(p, 0.9), atol=1e-6)  # first Adam step moves by lr * sign(g)
    assert gr.abs().max().item() == 0.0
    gr.fill_(float("inf"))
    F.adam_step([p], [gr], [m], [v], [0.1], step=2, found_inf=torch.ones(1, device="cuda"), zero_grads=True)
    torch.cuda.synchronize()
    assert torch.allclose(p, torch.full_like(p, 0.9), atol=1e-6)  # skipped
    assert gr.abs().max().item() == 0.0  # but the gradients were consumed


def test_model_api_training_steps_reduce_loss():
    """Plugin surface end to end: model(ray_bundle) -> get_metrics_dict -> get_loss_dict -> backward ->
    FusedAdam, with the training callbacks; the loss on a fixed batch must go down."""
    from thermo_nerf_b200 import FusedAdam, RayBundle

    _, model = make_pair(trained_like=False, precision="tc_fp16", log2_field=14, log2_prop=12)
    model.train()
    rays = make_synthetic_rays(1024, num_images=8, seed=5)
    gen = torch.Generator().manual_seed(1)
    batch = {"image": torch.rand((1024, 3), generator=gen).mul(0.2).add(0.4).cuda(),
             "thermal": torch.rand((1024, 1), generator=gen).mul(0.2).add(0.6)}  # thermal GT stays on the host
    groups = model.get_param_groups()
    opts = [FusedAdam(groups["proposal_networks"], lr=1e-2, eps=1e-15),
            FusedAdam(groups["fields"], lr=1e-2, eps=1e-15)]
    cbs = model.get_training_callbacks()
    first = last = None
    for step in range(30):
        for c in cbs:
            if c.where_to_run == ["BEFORE_TRAIN_ITERATION"]:
                c.run_callback(step)
        rb = RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(),
                       camera_indices=rays.camera_indices.cuda())
        for o in opts:
            o.zero_grad()
        out = model(rb)
        metrics = model.get_metrics_dict(out, batch)
        ld = model.get_loss_dict(out, batch, metrics)
        assert set(ld) == {"rgb_loss", "interlevel_loss", "distortion_loss", "thermal"}
        loss = sum(ld.values())
        loss.backward()
        for o in opts:
            o.step()
        for c in cbs:
            if c.where_to_run == ["AFTER_TRAIN_ITERATION"]:
                c.run_callback(step)
        val = (ld["rgb_loss"] + ld["thermal"]).item()
        first = val if first is None else first
        last = val
    assert last < 0.5 * first, (first, last)


@pytest.mark.parametrize("precision,tol,med_tol", [("fp32", 3e-2, 2e-3), ("tc_fp16", 8e-2, 3e-2)])
def test_pose_gradients_match_oracle_autograd(precision, tol, med_tol):
    """Camera-optimiser path (the reference's default camera_optimizer_mode="SO3xR3"): get_outputs applies the pose
    deltas to the ray bundle (thermal_nerf_model.py:218-219) and the loss back-propagates into them through
    dL/d origins and dL/d directions - checked against oracle autograd on the rays themselves and on
    camera_optimizer.pose_adjustment through the plugin surface.

    The position derivative of a hash grid is piecewise constant per cell and the inverse-CDF resampling divides by
    near-zero CDF increments in empty space, so a last-ulp difference in a sample position moves a handful of rays
    visibly (measured: median per-ray error 3e-4, ~4 % of the rays above 5 %).  The bounds are therefore a tight
    median per ray plus a looser global rel-L2 / cosine."""
    from thermo_nerf_b200 import RayBundle
    from thermo_nerf_b200 import functional as F
    from thermo_nerf_b200 import _lib as L

    oracle, model = make_pair(trained_like=True, precision=precision, log2_field=12, log2_prop=10,
                              camera_optimizer_mode="SO3xR3")
    R = 256
    rays = make_synthetic_rays(R, num_images=8, seed=31)
    g = torch.Generator().manual_seed(31)
    jitter = torch.rand((3, R, 1), generator=g)
    gt_rgb, gt_th = torch.rand((R, 3), generator=g), torch.rand((R, 1), generator=g)
    mults = (1.0, 0.5)
    # ---- 1. gradients w.r.t. the rays (camera optimiser bypassed on both sides)
    oracle.anneal = 0.6
    oracle.cfg.interlevel_loss_mult, oracle.cfg.distortion_loss_mult = mults
    oracle.cfg.camera_optimizer_mode, oracle.camera_optimizer.mode = "off", "off"
    ro = OracleRays(rays.origins.clone().requires_grad_(True), rays.directions.clone().requires_grad_(True),
                    rays.camera_indices)
    out = oracle.get_outputs(ro, training=True, jitter=jitter)
    sum(oracle.get_loss_dict(out, gt_rgb, gt_th, training=True).values()).backward()
    model.train()
    prec = L.PRECISION_FP32 if precision == "fp32" else L.PRECISION_TC_FP16
    o = rays.origins.cuda().requires_grad_(True)
    d = rays.directions.cuda().requires_grad_(True)
    res = F.render(model.tensors(), o, d, rays.camera_indices.cuda(), None, None, jitter.cuda().reshape(3, -1),
                   num_samples=(256, 96, 48), near_plane=0.05, far_plane=1000.0, anneal=0.6,
                   appearance_mode=L.APPEARANCE_LOOKUP, precision=prec)
    ld = F.losses(res, gt_rgb.cuda(), gt_th.cuda(), interlevel_mult=mults[0], distortion_mult=mults[1])
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    for name, ours, ref in (("origins", o.grad.cpu(), ro.origins.grad), ("directions", d.grad.cpu(), ro.directions.grad)):
        err = _rel_l2(ours, ref)
        cos = torch.nn.functional.cosine_similarity(ours.flatten(), ref.flatten(), dim=0).item()
        per_ray = (ours - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-12)
        med, frac = per_ray.median().item(), (per_ray > 0.05).float().mean().item()
        print(f"d loss / d {name}: rel-L2 {err:.2e} cos {cos:.5f} per-ray median {med:.2e}, {frac:.1%} of rays > 5%")
        assert err <= tol and cos >= 0.995 and med <= med_tol and frac <= 0.12, (name, err, cos, med, frac)
    # ---- 2. through the plugin surface into the pose deltas
    oracle.cfg.camera_optimizer_mode, oracle.camera_optimizer.mode = "SO3xR3", "SO3xR3"
    oracle.zero_grad()
    out = oracle.get_outputs(rays, training=True, jitter=jitter)
    sum(oracle.get_loss_dict(out, gt_rgb, gt_th, training=True).values()).backward()
    ref = oracle.camera_optimizer.pose_adjustment.grad.clone()
    assert ref.abs().sum() > 0
    model.zero_grad()
    rb = RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(), camera_indices=rays.camera_indices.cuda())
    model.camera_optimizer.apply_to_raybundle(rb)  # get_outputs would also draw its own jitter: call the pieces
    res = F.render(model.tensors(), rb.origins.contiguous(), rb.directions.contiguous(), rays.camera_indices.cuda(), None,
                   None, jitter.cuda().reshape(3, -1), num_samples=(256, 96, 48), near_plane=0.05, far_plane=1000.0,
                   anneal=0.6, appearance_mode=L.APPEARANCE_LOOKUP, precision=prec)
    ld = F.losses(res, gt_rgb.cuda(), gt_th.cuda(), interlevel_mult=mults[0], distortion_mult=mults[1])
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    ours = model.camera_optimizer.pose_adjustment.grad.cpu()
    err = _rel_l2(ours, ref)
    print(f"pose_adjustment.grad: rel-L2 {err:.2e}")
    assert err <= tol, err
    # and model(ray_bundle) itself keeps the pose deltas in the graph
    model.zero_grad()
    outm = model(RayBundle(origins=rays.origins.cuda(), directions=rays.directions.cuda(),
                           camera_indices=rays.camera_indices.cuda()))
    (outm["rgb"].sum() + outm["thermal"].sum()).backward()
    assert model.camera_optimizer.pose_adjustment.grad.abs().sum().item() > 0


@pytest.mark.parametrize("precision,tol", [("fp32", 0.02), ("tc_fp16", 0.06)])
def test_training_trajectory_tracks_oracle(precision, tol):
    """Whole iterations, many of them: TrainEngine (forward, fused losses, backward, fused Adam, the sampler's
    update schedule and weight annealing) against the oracle driven the way nerfstudio's Trainer drives the reference
    (autograd + torch.optim.Adam(lr=1e-2, eps=1e-15) per param group, proposal nets under no_grad on non-update
    steps), same initial weights, same ray batches, same jitter draws.  The loss trajectories must stay together."""
    from thermo_nerf_b200.engine import TrainEngine, exponential_decay_lr

    oracle, model = make_pair(trained_like=False, precision=precision, log2_field=12, log2_prop=10,
                              camera_optimizer_mode="off")
    model.train()
    eng = TrainEngine(model)
    field = [p for n, p in oracle.named_parameters() if n.startswith("field.")]
    props = [p for n, p in oracle.named_parameters() if n.startswith("proposal_networks.")]
    opt_f, opt_p = torch.optim.Adam(field, lr=1e-2, eps=1e-15), torch.optim.Adam(props, lr=1e-2, eps=1e-15)
    R, steps = 384, 24
    gen = torch.Generator().manual_seed(77)
    ours_hist, ref_hist = [], []
    since_update = 0
    for step in range(steps):
        rays = make_synthetic_rays(R, num_images=8, seed=100 + step % 3)  # three batches, cycled
        gt_rgb = (0.5 + 0.4 * torch.sin(rays.directions * 7.0)).float()
        gt_th = (0.5 + 0.4 * torch.cos(rays.directions[:, :1] * 5.0)).float()
        jitter = torch.rand((3, R, 1), generator=gen)
        # ---- oracle: what Trainer.train_iteration does around the reference model
        prev = max(step - 1, 0)  # the sampler's _step is set by step_cb AFTER an iteration: it lags by one
        updated = since_update > model.update_schedule(prev) or prev < 10
        oracle.set_anneal_for_step(step)
        opt_f.zero_grad(); opt_p.zero_grad()
        out = oracle.get_outputs(rays, training=True, jitter=jitter, prop_grad=updated)
        ld = oracle.get_loss_dict(out, gt_rgb, gt_th, training=True)
        sum(ld.values()).backward()
        for opt in (opt_f, opt_p):
            for grp in opt.param_groups:
                grp["lr"] = exponential_decay_lr(step)
        opt_f.step()
        if updated:
            opt_p.step()
            since_update = 0
        since_update += 1
        ref_hist.append(float(sum(ld.values())))
        # ---- ours
        ls = eng.step(rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda().reshape(-1), gt_rgb.cuda(),
                      gt_th.cuda().reshape(-1), jitter=jitter.cuda().reshape(3, -1))
        ours_hist.append(float(ls.sum()))
    print("oracle:", ["%.5f" % v for v in ref_hist[::4]])
    print("ours:  ", ["%.5f" % v for v in ours_hist[::4]])
    assert ref_hist[-1] < 0.6 * ref_hist[0]
    for a, b in zip(ours_hist, ref_hist):
        assert abs(a - b) <= tol * abs(b) + 1e-5, (ours_hist, ref_hist)


@pytest.mark.parametrize("precision", ["fp32", "tc_fp16"])
def test_backward_ragged_ray_and_sample_counts(precision):
    """193 rays x (64, 33, 17) samples: nothing is a multiple of a tile - ragged 16-sample field tiles (bulk stores of
    1 row), ragged 64-row blocks of the weight-gradient GEMMs (zero-filled tails), partial proposal chunks."""
    from thermo_nerf_b200 import functional as F
    from thermo_nerf_b200 import _lib as L

    ns = (64, 33, 17)
    oracle, model = make_pair(trained_like=True, precision=precision, log2_field=12, log2_prop=10,
                              camera_optimizer_mode="off", num_samples=ns)
    R = 193
    rays = make_synthetic_rays(R, num_images=8, seed=41)
    g = torch.Generator().manual_seed(41)
    jitter = torch.rand((3, R, 1), generator=g)
    gt_rgb, gt_th = torch.rand((R, 3), generator=g), torch.rand((R, 1), generator=g)
    mults = (1.0, 0.5)
    if precision != "fp32":
        # coherent temperature residuals (targets 0.25 above the rendered values): with random-sign residuals the
        # temperature chain's gradient cancels down to (tensor-core forward error) x Jacobian and the relative error
        # measures that forward error, not the backward kernel (see test_fullsize_gpu's gradient test)
        with torch.no_grad():
            oracle.anneal = 0.7
            gt_th = (oracle.get_outputs(rays, training=True, jitter=jitter)["thermal"] + 0.25).clamp(0.0, 1.0)
    _, o_ld, o_g = _oracle_grads(oracle, rays, jitter, gt_rgb, gt_th, 0.7, mults)
    model.train()
    model.zero_grad()
    prec = L.PRECISION_FP32 if precision == "fp32" else L.PRECISION_TC_FP16
    out = F.render(model.tensors(), rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda(), None, None,
                   jitter.cuda().reshape(3, -1), num_samples=ns, near_plane=0.05, far_plane=1000.0, anneal=0.7,
                   appearance_mode=L.APPEARANCE_LOOKUP, precision=prec)
    ld = F.losses(out, gt_rgb.cuda(), gt_th.cuda(), interlevel_mult=mults[0], distortion_mult=mults[1])
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    # fp32: 5e-3 instead of the 2e-3 of the 256/96/48 test - 17 coarse field samples per ray carry large
    # delta*sigma each, which amplifies last-ulp differences in the bin edges (measured 2.9e-3 on one tensor)
    tol_l, tol_g, tol_t = (2e-3, 5e-3, 1.5e-2) if precision == "fp32" else (3e-2, 5e-2, 5e-2)
    for k in o_ld:
        assert ld[k].item() == pytest.approx(o_ld[k].item(), rel=tol_l, abs=1e-5), k
    bad = []
    for k, p in model.named_parameters():
        ref = o_g.get(k)
        if k.startswith("camera_optimizer") or ref is None or ref.norm() == 0:
            continue
        err = _rel_l2(p.grad.detach().cpu(), ref)
        if err > (tol_t if k.endswith("hash_table") or k.startswith("field.mlp_base") else tol_g):
            bad.append((k, err))
    assert not bad, bad


def test_engine_step_from_host_batches_equals_the_device_step():
    """TrainEngine.step_host (pinned host batch -> persistent device buffers -> step) is the same iteration as
    TrainEngine.step on device tensors: identical losses and parameters for identical jitter draws."""
    from thermo_nerf_b200.engine import TrainEngine

    R = 256
    rays = make_synthetic_rays(R, num_images=8, seed=21)
    gt_rgb = (0.5 + 0.4 * torch.sin(rays.directions * 7.0)).float()
    gt_th = (0.5 + 0.4 * torch.cos(rays.directions[:, 0] * 5.0)).float()
    results = []
    for host in (False, True):
        _, model = make_pair(trained_like=False, precision="tc_fp16", log2_field=12, log2_prop=10,
                             camera_optimizer_mode="off")
        model.train()
        eng = TrainEngine(model)
        torch.manual_seed(5)
        for _ in range(3):
            if host:
                batch = [t.pin_memory() for t in (rays.origins, rays.directions, rays.camera_indices.reshape(-1), gt_rgb, gt_th)]
                ls = eng.step_host(*batch)
            else:
                ls = eng.step(rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda().reshape(-1),
                              gt_rgb.cuda(), gt_th.cuda())
        torch.cuda.synchronize()
        results.append((ls.cpu(), model.field.mlp_head.layers[1].weight.detach().cpu().clone()))
    # hash-table gradients are summed with atomics: last-bit differences between two runs are expected
    assert torch.allclose(results[0][0], results[1][0], rtol=1e-4, atol=1e-6)
    assert torch.allclose(results[0][1], results[1][1], rtol=1e-3, atol=1e-5)
    with pytest.raises(ValueError, match="camera optimiser"):
        _, m2 = make_pair(trained_like=False, log2_field=12, log2_prop=10, camera_optimizer_mode="SO3xR3")
        TrainEngine(m2)
