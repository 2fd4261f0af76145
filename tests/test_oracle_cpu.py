"""Oracle known-answer tests (SURVEY 8c list) and golden-vector pinning.  CPU only.

PARITY UNPINNED: no reference test or fixture pins this path, so these KATs check the
restatement against closed forms and hand-computed values, and the goldens pin the oracle
against drift (tests/golden/make_golden.py)."""

from pathlib import Path

import pytest
import torch

from oracle import nerfstudio_math as M
from tests.golden.make_golden import CASES, build_case, weights_checksum

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_hash_scalings_match_the_float32_expression():
    assert M.hash_scalings(16, 16, 2048).tolist() == [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776,
                                                     1072, 1482, 2047]
    assert M.hash_scalings(5, 16, 128).tolist() == [16, 26, 45, 76, 128]
    assert M.hash_scalings(5, 16, 256).tolist() == [16, 32, 64, 128, 256]


def test_hash_index_known_answers_and_uint32_equivalence():
    T, L = 19, 3
    triples = torch.tensor([[[0, 0, 0]] * L, [[1, 1, 1]] * L, [[2047, 2047, 2047]] * L, [[5, 17, 1023]] * L],
                           dtype=torch.int32)
    got = M.hash_indices(triples, T, L)
    for row, (x, y, z) in zip(got, [(0, 0, 0), (1, 1, 1), (2047, 2047, 2047), (5, 17, 1023)]):
        want = ((x * 1) ^ (y * 2654435761) ^ (z * 805459861)) % (1 << T)
        assert row.tolist() == [want + l * (1 << T) for l in range(L)]
    # the CUDA kernel uses 32-bit mul.lo/xor/and: identical low T bits for non-negative coordinates
    g = torch.Generator().manual_seed(0)
    c = torch.randint(0, 2049, (100000, 1, 3), generator=g, dtype=torch.int32)
    ref = M.hash_indices(c, T, 1)[:, 0]
    c64 = c[:, 0].to(torch.int64)
    u32 = ((c64[:, 0] & 0xFFFFFFFF) ^ ((c64[:, 1] * 2654435761) & 0xFFFFFFFF) ^ ((c64[:, 2] * 805459861) & 0xFFFFFFFF))
    assert torch.equal(ref, u32 & ((1 << T) - 1))


def test_hash_encode_integral_and_midpoint():
    L, T = 2, 6
    scal = torch.tensor([4.0, 8.0])
    table = torch.arange(L * (1 << T) * 2, dtype=torch.float32).view(-1, 2)
    # integral scaled coordinate: ceil == floor, so the value is the single corner entry
    x = torch.tensor([[0.25, 0.5, 0.75]])
    enc = M.hash_encode(x, table, scal, T)
    for l, s in enumerate(scal.tolist()):
        ix, iy, iz = int(0.25 * s), int(0.5 * s), int(0.75 * s)
        idx = ((ix ^ (iy * 2654435761) ^ (iz * 805459861)) % (1 << T)) + l * (1 << T)
        assert torch.allclose(enc[0, 2 * l:2 * l + 2], table[idx])
    # constant table -> any point encodes to the constant (weights sum to 1)
    enc = M.hash_encode(torch.rand(50, 3), torch.full_like(table, 3.5), scal, T)
    assert torch.allclose(enc, torch.full_like(enc, 3.5), atol=1e-5)


def test_contraction_known_answers():
    x = torch.tensor([[0.5, -0.25, 0.1], [1.0, 0.0, 0.0], [2.0, 0.0, 0.0], [0.0, -4.0, 2.0]])
    got = M.contract_linf(x)
    want = torch.tensor([[0.5, -0.25, 0.1], [1.0, 0.0, 0.0], [1.5, 0.0, 0.0], [0.0, -1.75, 0.875]])
    assert torch.allclose(got, want)
    far = M.contract_linf(torch.tensor([[1e6, 0.0, 0.0]]))
    assert 1.999 < far[0, 0] < 2.0


def test_normalise_positions_selector():
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    p, sel = M.normalise_positions(torch.tensor([[0.0, 0.0, 0.0], [3.0, 0.0, 0.0]]), aabb, True)
    assert sel.tolist() == [True, True] and torch.allclose(p[0], torch.tensor([0.5, 0.5, 0.5]))
    p, sel = M.normalise_positions(torch.tensor([[0.0, 0.0, 0.0], [3.0, 0.0, 0.0]]), aabb, False)
    assert sel.tolist() == [True, False] and p[1].abs().sum() == 0


def test_sh4_on_axes():
    c = M.sh4(torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [-1.0, 0, 0]]))
    assert torch.allclose(c[:, 0], torch.full((4,), 0.28209479177387814))
    assert c[0, 3].item() == pytest.approx(0.4886025119029199) and c[3, 3].item() == pytest.approx(-0.4886025119029199)
    assert c[1, 1].item() == pytest.approx(0.4886025119029199) and c[2, 2].item() == pytest.approx(0.4886025119029199)
    assert c[2, 6].item() == pytest.approx(0.9461746957575601 - 0.31539156525251999)
    assert c[0, 8].item() == pytest.approx(0.5462742152960396) and c[1, 8].item() == pytest.approx(-0.5462742152960396)
    assert c[2, 12].item() == pytest.approx(0.3731763325901154 * 2)
    assert c[0, 15].item() == pytest.approx(0.5900435899266435)


def test_piecewise_spacing_known_answers():
    x = torch.tensor([0.05, 0.5, 1.0, 2.0, 1000.0])
    s = M.spacing_fn(x)
    assert torch.allclose(s, torch.tensor([0.025, 0.25, 0.5, 0.75, 0.9995]))
    assert torch.allclose(M.spacing_fn_inv(s), x, rtol=1e-3)
    to_e = M.make_spacing_to_euclid(torch.tensor([[0.05]]), torch.tensor([[1000.0]]))
    e = to_e(torch.tensor([[0.0, 1.0]]))
    assert torch.allclose(e, torch.tensor([[0.05, 1000.0]]), rtol=1e-3)


def test_initial_bins_eval_and_jittered():
    b = M.piecewise_initial_bins(2, 4, None)
    assert torch.allclose(b, torch.tensor([[0, 0.25, 0.5, 0.75, 1.0]] * 2))
    j = M.piecewise_initial_bins(2, 4, torch.tensor([[0.0], [1.0]]))
    assert torch.allclose(j[0], torch.tensor([0, 0.125, 0.375, 0.625, 0.875]))
    assert torch.allclose(j[1], torch.tensor([0.125, 0.375, 0.625, 0.875, 1.0]))


def test_get_weights_constant_density_closed_form():
    sigma, S = 2.0, 8
    t = torch.linspace(0.0, 2.0, S + 1)
    deltas = (t[1:] - t[:-1]).view(1, S, 1)
    w = M.get_weights(deltas, torch.full((1, S, 1), sigma))[0, :, 0]
    want = torch.exp(-sigma * t[:-1]) * (1 - torch.exp(-sigma * (t[1:] - t[:-1])))
    assert torch.allclose(w, want, atol=1e-6)
    assert w.sum().item() == pytest.approx(1 - float(torch.exp(torch.tensor(-sigma * 2.0))), abs=1e-6)


def test_pdf_sampler_one_hot_weights():
    S_prev, S_new = 8, 4
    w = torch.zeros(1, S_prev)
    w[0, 3] = 1.0
    existing = torch.linspace(0, 1, S_prev + 1)[None]
    bins = M.pdf_resample_bins(w, existing, S_new, None)
    assert bins.shape == (1, S_new + 1)
    assert torch.all(bins[0, 1:] >= bins[0, :-1])
    # 1/(1+8*0.01) of the mass sits in [3/8, 4/8]: the central bins must fall inside it
    inside = ((bins[0] >= 3 / 8 - 1e-6) & (bins[0] <= 4 / 8 + 1e-6)).sum().item()
    assert inside >= 3
    # uniform weights reproduce a uniform resampling of the same interval
    bins_u = M.pdf_resample_bins(torch.ones(1, S_prev), existing, S_new, None)
    u = torch.linspace(0, 1 - 1 / (S_new + 1), S_new + 1) + 1 / (2 * (S_new + 1))
    assert torch.allclose(bins_u[0], u, atol=1e-6)


def test_renderers_simple_cases():
    w = torch.tensor([[[0.1], [0.6], [0.2]]])
    c = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]])
    rgb = M.render_rgb_last_sample(c, w, training=False)
    assert torch.allclose(rgb, torch.tensor([[0.1, 0.6, 0.2 + 0.1]]))  # last-sample background gets 1 - sum(w)
    starts = torch.tensor([[[0.0], [1.0], [2.0]]])
    ends = starts + 1.0
    assert M.render_depth_median(w, starts, ends).item() == pytest.approx(1.5)
    assert M.render_depth_expected(w, starts, ends).item() == pytest.approx((0.05 + 0.9 + 0.5) / 0.9, rel=1e-5)
    assert M.render_depth_median(w * 0.1, starts, ends).item() == pytest.approx(2.5)  # never reaches 0.5 -> last
    # eval clamps, training does not
    big = M.render_rgb_last_sample(c * 3, w, training=True)
    assert big.max() > 1.0
    assert M.render_rgb_last_sample(c * 3, w, training=False).max() <= 1.0


def test_losses_closed_forms():
    # a single unit-weight interval: distortion = (e - s)/3
    t = torch.tensor([[0.2, 0.5]])
    w = torch.tensor([[[1.0]]])
    assert M.distortion_loss([w], [t]).item() == pytest.approx(0.3 / 3)
    # proposal envelope that dominates the fine weights -> zero interlevel loss
    c = torch.tensor([[0.0, 0.5, 1.0]])
    wf = torch.tensor([[[0.3], [0.4]]])
    cp = torch.tensor([[0.0, 1.0]])
    wp = torch.tensor([[[0.9]]])
    assert M.interlevel_loss([wp, wf], [cp, c]).item() == pytest.approx(0.0)
    wp_small = torch.tensor([[[0.1]]])
    assert M.interlevel_loss([wp_small, wf], [cp, c]).item() > 0


def test_mae_thermal_denormalisation():
    # temperature bounds of the reference's own fixture tests/data/thermal/temperature_bounds.json
    tmax, tmin = 33.085, 13.896
    gt = torch.tensor([0.0, 0.5, 1.0])
    pred = torch.tensor([0.1, 0.5, 0.8])
    mae = M.mae_thermal(gt, pred, False, tmax, tmin)
    assert mae.item() == pytest.approx((0.1 + 0.0 + 0.2) / 3 * (tmax - tmin), rel=1e-5)
    fg = M.mae_thermal(gt, pred, False, tmax, tmin, threshold=0.4)
    assert fg.item() == pytest.approx((0.0 + 0.2) / 2 * (tmax - tmin), rel=1e-5)
    cold = M.mae_thermal(gt, pred, True, tmax, tmin, threshold=0.4)
    assert cold.item() == pytest.approx(0.1 * (tmax - tmin), rel=1e-5)


def test_trunc_exp_backward_is_clamped():
    x = torch.tensor([0.0, 20.0], requires_grad=True)
    M.trunc_exp(x).sum().backward()
    assert x.grad[0].item() == pytest.approx(1.0)
    assert x.grad[1].item() == pytest.approx(float(torch.exp(torch.tensor(15.0))), rel=1e-5)


def test_exp_map_so3xr3_small_rotation():
    t = torch.tensor([[0.1, 0.2, 0.3, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0, 0.0, torch.pi / 2]])
    m = M.exp_map_so3xr3(t)
    assert torch.allclose(m[0, :, :3], torch.eye(3), atol=1e-6) and torch.allclose(m[0, :, 3], t[0, :3])
    assert torch.allclose(m[1, :, :3], torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]]), atol=1e-6)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    blob = torch.load(GOLDEN / f"{name}.pt", weights_only=True)
    model, rays, jitter, training = build_case(name)
    if abs(weights_checksum(model) - blob["weights_checksum"]) > 1e-6 * abs(blob["weights_checksum"]):
        pytest.skip("torch CPU RNG stream differs from the one the golden was generated with")
    assert torch.equal(rays.origins, blob["origins"]) and torch.equal(rays.directions, blob["directions"])
    with torch.no_grad():
        out = model.get_outputs(rays, training=training, jitter=jitter)
    for k, v in blob["outputs"].items():
        if isinstance(v, torch.Tensor):
            assert torch.allclose(out[k], v, atol=1e-5, rtol=1e-4), k
    for a, b in zip(out["sdist_list"], blob["outputs"]["sdist_list"]):
        assert torch.allclose(a, b, atol=1e-6)


def test_oracle_output_contract_plumbing_config():
    """BASELINE.json configs[0]: synthetic 32-ray / 16-sample forward on CPU (plumbing)."""
    model, rays, _, _ = build_case("e2e_mini_r32")
    with torch.no_grad():
        out = model.get_outputs(rays, training=False)
    for k, c in (("rgb", 3), ("thermal", 1), ("depth", 1), ("expected_depth", 1), ("accumulation", 1),
                 ("prop_depth_0", 1), ("prop_depth_1", 1)):
        assert out[k].shape == (32, c) and out[k].dtype == torch.float32
    assert float(out["rgb"].min()) >= 0 and float(out["rgb"].max()) <= 1
    assert [w.shape[1] for w in out["weights_list"]] == [16, 8, 16]


def test_oracle_training_gradients_flow_to_all_parameter_groups():
    model, rays, jitter, _ = build_case("e2e_train_r128")
    out = model.get_outputs(rays, training=True, jitter=jitter)
    loss = model.get_loss_dict(out, torch.rand(128, 3), torch.rand(128, 1))
    assert set(loss) == {"rgb_loss", "interlevel_loss", "distortion_loss", "thermal"}
    sum(loss.values()).backward()
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert model.proposal_networks[0].encoding.hash_table.grad.abs().sum() > 0
    assert model.field.mlp_thermal.layers[0].weight.grad.abs().sum() > 0
    assert model.camera_optimizer.pose_adjustment.grad.abs().sum() > 0


def test_thermal_metrics_match_vectors_generated_by_the_reference_itself():
    """tests/golden/reference_thermal_metrics.pt was produced by running the reference's own
    thermo_nerf/thermal_nerf/thermal_metrics.py (make_reference_golden.py): the oracle's and the product model's
    mae_thermal must reproduce it bit for bit; the modality / model-type enums must carry the reference's values."""
    from thermo_nerf_b200 import RenderedImageModality, ThermalNerfModel, ThermalNerfModelConfig

    blob = torch.load(GOLDEN / "reference_thermal_metrics.pt", weights_only=True)
    args = [{"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 128, "use_linear": False}] * 2
    checked = 0
    for c in blob["cases"]:
        want = c["mae"]
        got = M.mae_thermal(c["gt"], c["pred"], c["cold"], c["tmax"], c["tmin"], threshold=c["threshold"])
        assert torch.equal(got, want) or (torch.isnan(got) and torch.isnan(want))
        cfg = ThermalNerfModelConfig(log2_hashmap_size=8, proposal_net_args_list=args, max_temperature=c["tmax"],
                                     min_temperature=c["tmin"], cold=c["cold"])
        if checked % 12 == 0:  # model construction is the slow part: every 12th case through the plugin surface
            model = ThermalNerfModel(cfg, {"thermal": []}, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), 2)
            got_m = model.mae_thermal(c["gt"], c["pred"], threshold=c["threshold"])
            assert torch.equal(got_m, want) or (torch.isnan(got_m) and torch.isnan(want))
        checked += 1
    assert checked == len(blob["cases"]) >= 72
    assert {m.name: m.value for m in RenderedImageModality} == blob["modalities"]
    assert blob["model_types"] == {"THERMALNERFACTO": 1, "THERMONERF": 2, "CONCATNERF": 3, "NERFACTO": 4}


# ---- compositing pinned by the reference's own renderer code (tests/golden/make_reference_renderer_golden.py) ----
def _reference_renderers():
    from pathlib import Path

    return torch.load(Path(__file__).parent / "golden" / "reference_renderers.pt", weights_only=True)


def test_thermal_compositing_matches_the_reference_renderer_bitwise():
    """ThermalRenderer.forward (thermal_renderer.py:113-149) executed from the reference: train and eval mode,
    out-of-range and non-finite samples, near-empty and opaque rays."""
    blob = _reference_renderers()
    assert len(blob["thermal"]) == 6
    for case in blob["thermal"]:
        got = M.render_rgb_last_sample(case["thermal"].clone(), case["weights"], case["training"])
        assert got.shape == case["out"].shape
        assert torch.equal(got, case["out"]), (case["training"], (got - case["out"]).abs().max())
    # eval output is clamped to [0,1]; training output is not
    assert all(((c["out"] >= 0) & (c["out"] <= 1)).all() for c in blob["thermal"] if not c["training"])
    assert any(((c["out"] < 0) | (c["out"] > 1)).any() for c in blob["thermal"] if c["training"])
    # the background argument is overridden to "last_sample" inside combine_thermal (thermal_renderer.py:49)
    f = blob["thermal_forced_background"]
    assert torch.equal(M.render_rgb_last_sample(f["thermal"], f["weights"], True), f["out"])


def test_concat_compositing_and_loss_blend_match_the_reference_renderer_bitwise():
    """RGBTRenderer.forward with its default "random" background (rgbt_renderer.py:63-71,163-174) and the loss-time
    blend of rgbt_renderer.py:134-140, executed from the reference."""
    blob = _reference_renderers()
    for case in blob["rgbt"]:
        got = M.render_rgbt_no_background(case["rgbt"].clone(), case["weights"], case["training"])
        assert torch.equal(got, case["out"]), (case["training"], (got - case["out"]).abs().max())
    for case in blob["blend"]:
        torch.manual_seed(case["seed"])
        noise = torch.rand_like(case["pred"])
        assert torch.equal(case["pred"] + noise * (1.0 - case["acc"]), case["pred_out"])
        assert torch.equal(case["gt"], case["gt_out"])  # the ground truth is left alone
