"""Backward pass of the feature-field branch (first slice).

Reference gradients: torch autograd through the oracle's ``sam_field`` + MeanRenderer (fp32 gradient of the rounded
forward pass; fp16 rounding points are straight-through).  On the CPU the kernel bodies run through tests/emu; on a
GPU the same comparison goes through ``snrf_feature_backward`` and a finite-difference-free linearity property."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import model_pair, test_rays

# stated tolerance: gradients agree to 2e-3 of the largest magnitude of the same tensor (fp32 sums in a different
# order; on the GPU additionally the forward's own 1-ulp fp16 flips of x and the order of the atomics)
GRAD_RTOL_OF_MAX = 2e-3


def _branch_inputs(cfg, orc, n, seed, which="sam"):
    """Rays + picked samples + weights as the march kernel would hand them over: taken from the oracle render."""
    o, d = test_rays(n, seed=seed)
    with torch.no_grad():
        full = orc.render_rays(o, d, get_feature=(which,), return_intermediates=True)
    eu = full["_eu1"]
    tm2 = eu[:, :-1] + eu[:, 1:]  # start + end per nerf sample; the kernels carry 2 x midpoint
    return o, d, torch.gather(tm2, 1, full["_best_ids"]), full["_sam_weights"]


def oracle_branch(orc, which, o, d, sam_t, sam_w):
    """out[N,n_out] = sum_k w_k * net(enc(pos_k)) with pos = o + d * (start + end) / 2 (rays.py:48-57)."""
    pos = o[:, None, :] + d[:, None, :] * sam_t[..., None] / 2.0
    f = orc.sam_field(pos, which=(which,))
    return (sam_w[..., None] * f[which]).sum(dim=-2), f


def oracle_branch_at(orc, which, o, d, sam_t, sam_w, enc_saved):
    """``oracle_branch`` evaluated at saved encoder outputs: the value of the hash-grid encoding is ``enc_saved``
    (what the device forward stored for its backward), its gradient is the oracle encoder's."""
    from oracle import tcnn_spec as T
    from oracle.samnerf_oracle import contract

    n = o.shape[0]
    pos = o[:, None, :] + d[:, None, :] * sam_t[..., None] / 2.0
    x = (contract(pos.reshape(-1, 3), None) + 2.0) / 4.0
    stem = "clip" if which == "sam" else "clipseg"
    enc = torch.cat([T.hash_grid_encode(x, orc.p[f"sam_field.{stem}_encs.{i}.params"], lv, g.n_features)
                     for i, (lv, g) in enumerate(zip(orc.sam_levels, orc.cfg.sam_grids))], dim=-1)
    enc = enc + (enc_saved.reshape(enc.shape).float() - enc).detach()
    f = T.mlp_forward(enc, orc.sam_w if which == "sam" else orc.clipseg_w).view(n, sam_t.shape[1], -1)
    return (sam_w[..., None] * f).sum(dim=-2)


def _fresh_oracle(cfg, params):
    from oracle.samnerf_oracle import Oracle

    p = {k: v.clone().requires_grad_(k.startswith("sam_field")) for k, v in params.items()}
    return Oracle(cfg, p), p


def _f16_bits(t):
    return np.ascontiguousarray(t.detach().to(torch.float16).numpy().view(np.uint16))


def _levels(cfg):
    lv = np.zeros((2, 12, 5), np.float64)
    for e, g in enumerate(cfg.sam_grids):
        for l, (scale, res, offset, size, hashed) in enumerate(g.levels()):
            lv[e, l] = (scale, res, size, offset, float(hashed))
    return lv


def _assert_grad_close(got, want, what):
    got, want = torch.as_tensor(got).flatten(), torch.as_tensor(want).flatten()
    scale = float(want.abs().max())
    assert scale > 0, what
    err = float((got - want).abs().max())
    print(f"{what}: max|err| {err:.3e}  max|grad| {scale:.3e}  ratio {err / scale:.2e}")
    assert err <= GRAD_RTOL_OF_MAX * scale, f"{what}: max|err| {err:.3e} vs max|grad| {scale:.3e}"
    # and the gradient must be where the reference has it (same sparsity pattern of the table scatter)
    # (entries below 1e-6 of the largest one may legitimately underflow to zero on one side only)
    nz_w, nz_g = want.abs() > 1e-6 * scale, got != 0
    assert float((nz_w & ~nz_g).float().mean()) < 1e-4, what


def _significant_prefix_mask(sam_w, cutoff):
    """Slots the kernels keep: everything up to the last slot whose weight is not below the cut-off (all 16 if < 0)."""
    if cutoff < 0:
        return torch.ones_like(sam_w, dtype=torch.bool)
    sig = (~(sam_w < cutoff)) & (sam_w != 0)
    k = torch.where(sig.any(-1), 16 - sig.flip(-1).float().argmax(-1), torch.zeros(sam_w.shape[0], dtype=torch.long))
    return torch.arange(16)[None, :] < k[:, None]


@pytest.mark.parametrize("which,clipseg,cutoff", [("sam", False, -1.0), ("clipseg", True, -1.0), ("sam", False, 1e-3)])
def test_kernel_bodies_match_autograd(which, clipseg, cutoff):
    from emu.build_emu import load

    cfg, params, orc0 = model_pair("tiny", "scene", 21, clipseg, 1)
    n = 40
    o, d, sam_t, sam_w = _branch_inputs(cfg, orc0, n, seed=9, which=which)
    ok = torch.isfinite(sam_w).all(-1)
    o, d, sam_t, sam_w = o[ok], d[ok], sam_t[ok], sam_w[ok]
    # the march kernel stores the picks in descending weight order (slot = rank); the oracle's topk is unsorted
    sam_w, order = sam_w.sort(dim=-1, descending=True)
    sam_t = torch.gather(sam_t, 1, order)
    n = o.shape[0]
    keep = _significant_prefix_mask(sam_w, cutoff)
    assert cutoff < 0 or 0.05 < float(keep.float().mean()) < 0.9  # the cut-off really drops rows in this test
    orc, p = _fresh_oracle(cfg, params)
    out, f = oracle_branch(orc, which, o, d, sam_t, sam_w * keep)  # reference: the dropped slots carry no weight
    g_out = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * g_out).sum().backward()
    enc_names = [f"sam_field.{'clip' if which == 'sam' else 'clipseg'}_encs.{i}.params" for i in range(2)]
    net_name = f"sam_field.{which}_net.params"
    n_out = out.shape[-1]

    # the encoder outputs the forward pass saves (fp16): recomputed here without grad
    with torch.no_grad():
        x = orc0.sam_field(o[:, None, :] + d[:, None, :] * sam_t[..., None] / 2.0, which=(which,))
        xs = x["hashgrid"] if which == "sam" else None
    if xs is None:  # the oracle only reports the hash-grid output for the SAM net; rebuild it for ClipSeg
        from oracle import tcnn_spec as T
        from oracle.samnerf_oracle import contract

        pts = (contract((o[:, None, :] + d[:, None, :] * sam_t[..., None] / 2.0).reshape(-1, 3), None) + 2.0) / 4.0
        xs = torch.cat([T.hash_grid_encode(pts, params[enc_names[i]], orc0.sam_levels[i], 8) for i in range(2)], -1).view(n, 16, 192)
    net = params[net_name]
    w1, w2 = net[: 256 * 192].view(256, 192), net[256 * 192:].view(n_out, 256)

    lib = load()
    g_w1 = np.zeros((256, 192), np.float32)
    g_w2 = np.zeros((n_out, 256), np.float32)
    g_t = [np.zeros(params[k].numel(), np.float32) for k in enc_names]
    hbar = np.zeros((n, 256), np.float32)
    arr = lambda t: np.ascontiguousarray(t.detach().numpy(), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    ins = [arr(o), arr(d), arr(sam_t), arr(sam_w), arr(g_out), _f16_bits(xs), _f16_bits(w1), _f16_bits(w2), _levels(cfg)]
    n_rows = C.c_int(0)
    lib.emu_feature_backward(ptr(ins[0]), ptr(ins[1]), ptr(ins[2]), ptr(ins[3]), C.c_longlong(n), ptr(ins[4]), n_out,
                             ptr(ins[5]), ptr(ins[6]), ptr(ins[7]), ptr(ins[8]), ptr(g_w1), ptr(g_w2), ptr(g_t[0]),
                             ptr(g_t[1]), ptr(hbar), C.c_float(cutoff), C.byref(n_rows))
    assert n_rows.value == int(keep.sum())  # one compact row per kept (ray, slot)
    want_net = p[net_name].grad
    _assert_grad_close(g_w1, want_net[: 256 * 192], "dW1")
    _assert_grad_close(g_w2, want_net[256 * 192:], "dW2")
    for i in range(2):
        _assert_grad_close(g_t[i], p[enc_names[i]].grad, f"d table {i}")
    # the recomputed weighted hidden sum is what the forward kernel feeds the output layer: W2 . hbar == out
    out2 = torch.from_numpy(hbar) @ w2.to(torch.float16).float().T
    assert torch.allclose(out2, out.detach(), rtol=2e-2, atol=2e-3)


def test_oracle_gradients_are_straight_through_fp16():
    """The gradient contract: rounding points do not round or zero the gradient."""
    from oracle import tcnn_spec as T

    x = torch.tensor([1e-9, 0.3333, -2.5], requires_grad=True)
    y = T.f16(x * 3.0)
    assert y.dtype == torch.float32 and float(y.detach()[1]) == float(torch.tensor(0.3333 * 3.0).half())
    y.sum().backward()
    assert torch.equal(x.grad, torch.full((3,), 3.0))


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("which,clipseg,cutoff", [("sam", False, -1.0), ("clipseg", True, -1.0), ("sam", False, 1e-3)])
def test_gpu_backward_matches_autograd(which, clipseg, cutoff):
    from helpers import assert_features_close, make_renderer

    cfg, params, orc0 = model_pair("tiny", "scene", 21, clipseg, 1)
    r = make_renderer(cfg, params)
    r.set_feature_cutoff(cutoff)  # the forward (bucketed kernel) and the backward (compact rows) share the rule
    o, d, sam_t, sam_w = _branch_inputs(cfg, orc0, 700, seed=9, which=which)
    ok = torch.isfinite(sam_w).all(-1)
    o, d, sam_t, sam_w = o[ok], d[ok], sam_t[ok], sam_w[ok]
    sam_w, order = sam_w.sort(dim=-1, descending=True)  # slot = rank, as the march kernel stores the picks
    sam_t = torch.gather(sam_t, 1, order)
    keep = _significant_prefix_mask(sam_w, cutoff)
    got_out, enc = r.feature_forward(which, o, d, sam_t, sam_w)
    with torch.no_grad():  # the training forward evaluates every slot (it saves all encoder outputs)
        out_full, _ = oracle_branch(orc0, which, o, d, sam_t, sam_w)
    assert_features_close(got_out, out_full, which + " forward", row_frac=0.99)
    # Reference gradient = autograd through the oracle AT THE FORWARD THE DEVICE COMPUTED: the saved fp16 encoder
    # outputs replace the oracle's own (value only - the gradient still flows into the oracle's tables).  Without this
    # the comparison is ill-posed at this size: a 1-ulp fp16 flip of one encoder output moves a hidden pre-activation
    # by ~1e-5, which flips the ReLU mask of a unit sitting that close to zero (expected a handful of times among
    # 700 x 16 x 256 units), and one flipped unit of a heavy sample changes a row of dW1 by O(|w g x|) ~ 0.5 - that
    # is a property of differentiating a rounded forward, not an error of either side.
    orc, p = _fresh_oracle(cfg, params)
    out = oracle_branch_at(orc, which, o, d, sam_t, sam_w * keep, enc.cpu())
    g_out = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * g_out).sum().backward()
    grads = r.feature_backward(which, o, d, sam_t, sam_w, enc, g_out)
    torch.cuda.synchronize()
    enc_names = [f"sam_field.{'clip' if which == 'sam' else 'clipseg'}_encs.{i}.params" for i in range(2)]
    _assert_grad_close(grads["net"].cpu(), p[f"sam_field.{which}_net.params"].grad, "d net")
    for i in range(2):
        _assert_grad_close(grads[f"grid{i}"].cpu(), p[enc_names[i]].grad, f"d table {i}")
    # properties that hold at any size: linear in d_out, accumulating (+=), frozen parameters untouched
    # (the sums are fp32 atomics in launch order: two runs agree to rounding of the largest terms, not of the result)
    def same(a, b):
        return float((a - b).abs().max()) <= 2e-5 * float(b.abs().max())

    g2 = r.feature_backward(which, o, d, sam_t, sam_w, enc, 2.0 * g_out)
    assert same(g2["net"], 2.0 * grads["net"]) and same(g2["grid0"], 2.0 * grads["grid0"])
    acc = {k: v.clone() for k, v in grads.items()}
    r.feature_backward(which, o, d, sam_t, sam_w, enc, g_out, grads=acc)
    assert same(acc["grid1"], 2.0 * grads["grid1"]) and same(acc["net"], 2.0 * grads["net"])
    only_net = r.feature_backward(which, o, d, sam_t, sam_w, enc, g_out, want=("net",))
    assert set(only_net) == {"net"} and same(only_net["net"], grads["net"])


@pytest.mark.gpu
def test_gpu_training_step_reduces_the_loss():
    """Autograd shim: one Adam step on the sam_field parameters through libsnrf lowers an MSE distillation loss."""
    from samnerf_b200.nerfstudio_api import SAMModel

    cfg, params, orc0 = model_pair("tiny", "scene", 21, False, 1)
    m = SAMModel(cfg)
    m.load_state_dict(params)
    m.train()
    m.proposal_sampler.train_stratified = False  # same samples every step, so that the loss is comparable
    o, d = test_rays(512, seed=4)
    from samnerf_b200.nerfstudio_api import RayBundle

    bundle = RayBundle(origins=o.cuda(), directions=d.cuda())
    target = torch.randn(512, 256, generator=torch.Generator().manual_seed(0)).cuda() * 0.1
    opt = torch.optim.Adam(m.get_param_groups()["sam_field"], lr=5e-3, eps=1e-15)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        out = m(bundle, get_feature=["sam"])
        loss = torch.nn.functional.mse_loss(out["sam"], target, reduction="none").mean(dim=-1).nanmean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


def test_training_shim_logic_on_cpu(monkeypatch):
    """SAMModel.train(): parameter groups, version-counter re-upload, autograd routing and state_dict, run on the CPU
    with the oracle as forward and the emulated kernel bodies as backward (tests/fake_renderer.py)."""
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg, params, _ = model_pair("tiny", "scene", 21, False, 4)
    m = api.SAMModel(cfg)
    m.load_state_dict({"_model." + k: v for k, v in params.items()})
    with pytest.raises(KeyError):
        m.load_state_dict({k: v for k, v in params.items() if "mlp_head" not in k})
    m.load_state_dict(params)
    m.train()
    groups = m.get_param_groups()
    assert len(groups["sam_field"]) == 3 and len(groups["conv"]) == 4
    assert len(groups["proposal_networks"]) == 1 and len(groups["fields"]) == 2
    assert m.collider.training
    o, d = test_rays(32, seed=4)  # 2 patches of 16 rays
    bundle = api.RayBundle(origins=o, directions=d)
    target = torch.randn(2, 256, generator=torch.Generator().manual_seed(0)) * 0.05
    m.proposal_sampler.train_stratified = False  # same samples every step, so that the loss is comparable
    opt = torch.optim.Adam(groups["sam_field"] + groups["conv"], lr=2e-4, eps=1e-15)
    losses = []
    for step in range(2):
        opt.zero_grad()
        out = m(bundle, get_feature=["sam"])
        assert out["sam"].shape == (2, 256) and out["rgb"].shape == (32, 3) and out["rgb"].requires_grad
        loss = torch.nn.functional.mse_loss(out["sam"], target, reduction="none").mean(dim=-1).nanmean()
        loss.backward()
        for p in groups["sam_field"]:
            assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
    # each optimiser step bumps the version counters -> the next forward re-uploads exactly the changed tensors
    assert m.renderer.uploads.count("sam_field.sam_net.params") == 1 and m.renderer.uploads.count("conv_head") == 1
    m.eval()
    assert not m.training and not m.collider.training and m.renderer.uploads[-1] == "conv_head"
    # eval after training renders with the trained feature field and conv head (same call as before training)
    ev = m(bundle, get_feature=["sam"])
    assert ev["sam"].shape == (2, 256) and not ev["sam"].requires_grad
    sd = m.state_dict()
    assert not torch.equal(sd["sam_field.sam_net.params"], params["sam_field.sam_net.params"])
    assert torch.equal(sd["field.mlp_base.params"], params["field.mlp_base.params"])
    assert sd["conv_head.0.weight"].shape == (256, 256, 3, 3)


# ---------------------------------------------------------------------------------------------------------
# density fields (nerfacto base + colour head, proposal)
# ---------------------------------------------------------------------------------------------------------
def _sample_positions(n, seed):
    """World positions along real rays (inside and outside the unit sphere) and their directions."""
    o, d = test_rays(n, seed=seed)
    g = torch.Generator().manual_seed(seed)
    t = torch.exp(torch.rand(n, generator=g) * 5.0 - 2.5)  # 0.08 .. 12 along the ray
    return (o + d * t[:, None]).contiguous(), d.contiguous()


def _field_activations(cfg, orc, params, which, pos, dirs):
    """The fp16 forward activations launch_field_backward recomputes on the device, here from the oracle's pieces."""
    from oracle import tcnn_spec as T

    x_pts, sel = orc._normalized(pos, float("inf"))
    if which == 1:
        table, levels, ws = orc.field_table, orc.field_levels, orc.base_w
    else:
        table, levels, ws = orc.prop_table, orc.prop_levels, orc.prop_w
    x = T.hash_grid_encode(x_pts, table, levels, 2)
    xin = x if which == 1 else torch.cat([x, torch.zeros(x.shape[0], 6)], -1)
    h1 = T.f16(torch.relu(xin @ T.f16(ws[0]).T))
    o = T.f16(h1 @ T.f16(ws[1]).T)
    acts = dict(x=x, h1=h1, o=o, sel=sel.float())
    if which == 1:
        hx = torch.cat([T.sh4((dirs + 1.0) / 2.0), o[:, 1:16], torch.ones(o.shape[0], 1)], -1)
        g1 = T.f16(torch.relu(hx @ T.f16(orc.head_w[0]).T))
        g2 = T.f16(torch.relu(g1 @ T.f16(orc.head_w[1]).T))
        acts.update(hx=hx, g1=g1, g2=g2, pre3=T.f16(g2 @ T.f16(orc.head_w[2]).T))
    return acts


def _grid_levels(g):
    lv = np.zeros((g.n_levels, 5), np.float64)
    for l, (scale, res, offset, size, hashed) in enumerate(g.levels()):
        lv[l] = (scale, res, size, offset, float(hashed))
    return lv


@pytest.mark.parametrize("which", [1, 0])
def test_field_kernel_bodies_match_autograd(which):
    from emu.build_emu import load
    from oracle.samnerf_oracle import Oracle

    cfg, params, orc0 = model_pair("tiny", "scene", 23, False, 1)
    n = 300
    pos, dirs = _sample_positions(n, seed=5)
    names = ["field.mlp_base.params", "field.mlp_head.params"] if which == 1 else ["proposal_networks.0.mlp_base.params"]
    p = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}
    orc = Oracle(cfg, p)
    gen = torch.Generator().manual_seed(2)
    g_d = torch.randn(n, generator=gen)
    if which == 1:
        dens, geo = orc.field_density(pos)
        rgb = orc.field_rgb(dirs, geo)
        g_rgb = torch.randn(n, 3, generator=gen)
        ((dens * g_d).sum() + (rgb * g_rgb).sum()).backward()
    else:
        dens = orc.proposal_density(pos)
        g_rgb = None
        (dens * g_d).sum().backward()
    with torch.no_grad():
        acts = _field_activations(cfg, orc0, params, which, pos, dirs)
    grid = cfg.field_grid if which == 1 else cfg.proposal_grid
    ws = orc0.base_w if which == 1 else orc0.prop_w
    g_base = np.zeros(params[names[0]].numel(), np.float32)
    g_head = np.zeros(params["field.mlp_head.params"].numel(), np.float32)
    bits = lambda t: None if t is None else _f16_bits(t)
    arr = lambda t: None if t is None else np.ascontiguousarray(t.detach().numpy(), np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    keep = [arr(pos), arr(g_d), arr(g_rgb), _grid_levels(grid), bits(ws[0]), bits(ws[1])]
    keep += [bits(w) for w in orc0.head_w]
    keep += [bits(acts[k]) if k in acts else None for k in ("x", "h1", "o", "hx", "g1", "g2", "pre3")]
    keep += [arr(acts["sel"])]
    load().emu_field_backward(which, ptr(keep[0]), C.c_longlong(n), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), grid.n_levels,
                              *[ptr(k) for k in keep[4:17]], ptr(g_base), ptr(g_head))
    want = p[names[0]].grad
    n_net = cfg.field_mlp_params if which == 1 else cfg.proposal_mlp_params
    _assert_grad_close(g_base[:n_net], want[:n_net], "d base MLP")
    _assert_grad_close(g_base[n_net:], want[n_net:], "d table")
    if which == 1:
        _assert_grad_close(g_head, p[names[1]].grad, "d head MLP")
        # rows 3..15 of the padded output layer are unused: exactly zero gradient
        assert not g_head.reshape(-1)[64 * 32 + 64 * 64 + 3 * 64:].any()


# ---------------------------------------------------------------------------------------------------------
# ray-wise ops: get_weights and the RGB composite
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S", [32, 64])
def test_ray_op_bodies_match_autograd(S):
    from emu.build_emu import load
    from oracle.samnerf_oracle import composite_rgb, get_weights

    lib = load()
    n = 200
    gen = torch.Generator().manual_seed(S)
    deltas = torch.rand(n, S, generator=gen) * 0.2 + 1e-3
    dens = torch.exp(torch.randn(n, S, generator=gen) * 2.0).requires_grad_(True)
    dens.data[::7, 3] = 0.0  # empty samples
    g_w = torch.randn(n, S, generator=gen)
    w = get_weights(deltas, dens)
    (w * g_w).sum().backward()
    arr = lambda t: np.ascontiguousarray(t.detach().numpy(), np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    d_dens = np.zeros((n, S), np.float32)
    a = [arr(deltas), arr(dens), arr(g_w)]
    lib.emu_weights_backward(ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(d_dens), C.c_longlong(n), S)
    _assert_grad_close(d_dens, dens.grad, "d densities")

    for bg in (None, (0.2, 0.7, 1.0)):
        rgb = torch.rand(n, S, 3, generator=gen).requires_grad_(True)
        wt = w.detach().clone().requires_grad_(True)
        g_out = torch.randn(n, 3, generator=gen)
        out = composite_rgb(rgb, wt, None if bg is None else torch.tensor(bg), training=True)
        (out * g_out).sum().backward()
        d_rgb, d_w = np.zeros((n, S, 3), np.float32), np.zeros((n, S), np.float32)
        b = [arr(rgb), arr(wt), arr(g_out), None if bg is None else np.asarray(bg, np.float32)]
        lib.emu_rgb_backward(ptr(b[0]), ptr(b[1]), ptr(b[2]), int(bg is not None), ptr(b[3]), ptr(d_rgb), ptr(d_w),
                             C.c_longlong(n), S)
        _assert_grad_close(d_rgb, rgb.grad, f"d rgb samples (bg {bg})")
        _assert_grad_close(d_w, wt.grad, f"d weights (bg {bg})")


def test_training_step_gradients_match_autograd_end_to_end(monkeypatch):
    """One full training step through the shim (proposal sampler -> fields -> get_weights -> RGB composite, every
    backward an emulated kernel body) against torch autograd through the oracle's own functions on the same detached
    sample positions: rgb loss + a loss on both weight lists, gradients of all three density-field tensors."""
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from oracle.samnerf_oracle import Oracle, composite_rgb, get_weights

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg, params, orc0 = model_pair("tiny", "scene", 25, False, 1)
    m = api.SAMModel(cfg)
    m.load_state_dict(params)
    m.train()
    o, d = test_rays(48, seed=6)
    gen = torch.Generator().manual_seed(3)
    target = torch.rand(48, 3, generator=gen)
    c0, c1 = torch.randn(48, 64, 1, generator=gen), torch.randn(48, 32, 1, generator=gen)

    torch.manual_seed(11)  # the sampler draws its stratified jitter from the global generator, like the reference
    out = m(api.RayBundle(origins=o, directions=d), get_feature=[])
    assert set(out) >= {"rgb", "depth", "accumulation", "prop_depth_0", "weights_list", "ray_samples_list"}
    loss = ((out["rgb"] - target) ** 2).mean() + (out["weights_list"][0] * c0).mean() + (out["weights_list"][1] * c1).mean()
    loss.backward()

    # the sample lists carry what the reference's interlevel / distortion losses read (losses.py:100-143)
    rs0, rs1 = out["ray_samples_list"]
    sd1 = torch.cat([rs1.spacing_starts[..., 0], rs1.spacing_ends[..., -1:, 0]], dim=-1)
    assert sd1.shape == (48, 33) and bool((sd1[:, 1:] >= sd1[:, :-1] - 1e-6).all()) and float(sd1.min()) >= -1e-6 and float(sd1.max()) <= 1 + 1e-6
    assert torch.allclose(rs1.spacing_to_euclidean_fn(sd1), torch.cat([rs1.frustums.starts[..., 0], rs1.frustums.ends[..., -1:, 0]], -1),
                          rtol=1e-4, atol=1e-5)
    # reference: same positions (training near plane 0.05, detached bins), oracle functions under autograd
    names = list(FakeRenderer.DENSITY_PARAMS)
    p = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}
    orc = Oracle(cfg, p)
    nears, fars = torch.full((48, 1), 0.05), torch.full((48, 1), float(cfg.far_plane))
    jit = m.proposal_sampler.last_jitter  # the stratified draws of this training step
    assert jit is not None and jit.shape == (48, 2)
    with torch.no_grad():
        res = orc0.render_rays(o, d, nears, fars, get_feature=(), return_intermediates=True, jitter=jit)
    eu0, eu1 = res["_eu0"], res["_eu1"]
    pos0 = o[:, None, :] + d[:, None, :] * ((eu0[:, :-1] + eu0[:, 1:]) / 2)[..., None]
    pos1 = o[:, None, :] + d[:, None, :] * ((eu1[:, :-1] + eu1[:, 1:]) / 2)[..., None]
    w0 = get_weights(eu0[:, 1:] - eu0[:, :-1], orc.proposal_density(pos0))
    dens, geo = orc.field_density(pos1)
    rgb_s = orc.field_rgb(d[:, None, :].expand(-1, 32, -1), geo)
    w1 = get_weights(eu1[:, 1:] - eu1[:, :-1], dens)
    rgb = composite_rgb(rgb_s, w1, None, training=True)
    ref_loss = ((rgb - target) ** 2).mean() + (w0[..., None] * c0).mean() + (w1[..., None] * c1).mean()
    ref_loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss.detach())) < 1e-5
    for name in names:
        _assert_grad_close(m.params[name].grad, p[name].grad, name)


# ---------------------------------------------------------------------------------------------------------
# the same comparisons through the C ABI on a GPU
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("which", ["field", "proposal"])
def test_gpu_field_backward_matches_autograd(which):
    from helpers import make_renderer
    from oracle.samnerf_oracle import Oracle

    cfg, params, orc0 = model_pair("tiny", "scene", 23, False, 1)
    r = make_renderer(cfg, params)
    n = 70000  # more than one internal block of 65 536 samples
    pos, dirs = _sample_positions(n, seed=5)
    names = ["field.mlp_base.params", "field.mlp_head.params"] if which == "field" else ["proposal_networks.0.mlp_base.params"]
    p = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}
    orc = Oracle(cfg, p)
    gen = torch.Generator().manual_seed(2)
    g_d = torch.randn(n, generator=gen)
    if which == "field":
        dens, geo = orc.field_density(pos)
        g_rgb = torch.randn(n, 3, generator=gen)
        ((dens * g_d).sum() + (orc.field_rgb(dirs, geo) * g_rgb).sum()).backward()
    else:
        g_rgb = None
        (orc.proposal_density(pos) * g_d).sum().backward()
    g = r.field_backward(which, pos, dirs if which == "field" else None, d_density=g_d, d_rgb=g_rgb)
    torch.cuda.synchronize()
    n_net = cfg.field_mlp_params if which == "field" else cfg.proposal_mlp_params
    want = p[names[0]].grad
    _assert_grad_close(g["base"][:n_net].cpu(), want[:n_net], "d base MLP")
    _assert_grad_close(g["base"][n_net:].cpu(), want[n_net:], "d table")
    if which == "field":
        _assert_grad_close(g["head"].cpu(), p[names[1]].grad, "d head MLP")
    # accumulation (+=) and density-only calls
    g2 = r.field_backward(which, pos, dirs if which == "field" else None, d_density=g_d, d_rgb=g_rgb,
                          grads={k: v.clone() for k, v in g.items()})
    assert torch.allclose(g2["base"], 2.0 * g["base"], rtol=1e-3, atol=1e-6 * float(g["base"].abs().max()))
    only_d = r.field_backward(which, pos, d_density=g_d)
    assert set(only_d) == {"base"} and torch.isfinite(only_d["base"]).all()


@pytest.mark.gpu
def test_gpu_ray_op_backward_matches_autograd():
    from helpers import make_renderer
    from oracle.samnerf_oracle import composite_rgb, get_weights

    cfg, params, _ = model_pair("tiny", "scene", 23, False, 1)
    r = make_renderer(cfg, params)
    for S in (32, 64):
        n = 3000
        gen = torch.Generator().manual_seed(S)
        deltas = torch.rand(n, S, generator=gen) * 0.2 + 1e-3
        dens = torch.exp(torch.randn(n, S, generator=gen) * 2.0).requires_grad_(True)
        g_w = torch.randn(n, S, generator=gen)
        w = get_weights(deltas, dens)
        (w * g_w).sum().backward()
        _assert_grad_close(r.ray_op_backward(0, deltas, dens.detach(), g_w).cpu(), dens.grad, "d densities")
        for bg in (None, (0.2, 0.7, 1.0)):
            rgb = torch.rand(n, S, 3, generator=gen).requires_grad_(True)
            wt = w.detach().clone().requires_grad_(True)
            g_out = torch.randn(n, 3, generator=gen)
            (composite_rgb(rgb, wt, None if bg is None else torch.tensor(bg), training=True) * g_out).sum().backward()
            d_rgb, d_w = r.ray_op_backward(3, rgb.detach(), wt.detach(), g_out, background=bg)
            _assert_grad_close(d_rgb.cpu(), rgb.grad, "d rgb samples")
            _assert_grad_close(d_w.cpu(), wt.grad, "d weights")
    with pytest.raises(RuntimeError, match="64 samples"):
        r.ray_op_backward(0, torch.rand(4, 65), torch.rand(4, 65), torch.rand(4, 65))


@pytest.mark.gpu
def test_gpu_full_training_step_lowers_rgb_and_feature_losses():
    from samnerf_b200.nerfstudio_api import RayBundle, SAMModel
    from samnerf_b200.training import interlevel_loss

    cfg, params, _ = model_pair("tiny", "scene", 25, False, 1)
    m = SAMModel(cfg)
    m.load_state_dict(params)
    m.train()
    m.proposal_sampler.train_stratified = False  # same samples every step, so that the losses are comparable
    o, d = test_rays(1024, seed=6)
    bundle = RayBundle(origins=o.cuda(), directions=d.cuda())
    gen = torch.Generator().manual_seed(3)
    image = torch.rand(1024, 3, generator=gen).cuda()
    feat = (torch.randn(1024, 256, generator=gen) * 0.1).cuda()
    groups = m.get_param_groups()
    opt = torch.optim.Adam([q for g in groups.values() for q in g], lr=1e-3, eps=1e-15)
    hist = []
    for _ in range(5):
        opt.zero_grad()
        out = m(bundle, get_feature=["sam"])
        rgb_loss = torch.nn.functional.mse_loss(out["rgb"], image)
        sam_loss = torch.nn.functional.mse_loss(out["sam"], feat, reduction="none").mean(dim=-1).nanmean()
        # the proposal network only learns from the interlevel loss (the sampler detaches its bins, ray_samplers.py:357;
        # nerfacto.py:324-327), exactly as in the reference: without it its gradient is None there too
        prop_loss = interlevel_loss(out["weights_list"], out["ray_samples_list"])
        (rgb_loss + sam_loss + prop_loss).backward()
        assert all(q.grad is not None and torch.isfinite(q.grad).all() for g in groups.values() for q in g)
        opt.step()
        hist.append((float(rgb_loss.detach()), float(sam_loss.detach())))
    assert hist[-1][0] < hist[0][0] and hist[-1][1] < hist[0][1], hist


def test_jittered_bins_are_the_torch_expression_bit_for_bit():
    """The closed form the march kernel uses for the training-mode spacing bins (common.cuh jittered_bin) against
    the reference's tensor expression (ray_samplers.py:100-111)."""
    from emu.build_emu import load

    n, nb = 4096, 64
    t_rand = torch.rand(n, 1, generator=torch.Generator().manual_seed(0))
    bins = torch.linspace(0.0, 1.0, nb + 1)[None, ...]
    centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
    upper, lower = torch.cat([centers, bins[..., -1:]], -1), torch.cat([bins[..., :1], centers], -1)
    want = lower + (upper - lower) * t_rand
    got = np.zeros((n, nb + 1), np.float32)
    tr = np.ascontiguousarray(t_rand.numpy().ravel(), np.float32)
    load().emu_jittered_bins(tr.ctypes.data_as(C.c_void_p), C.c_longlong(n), nb, got.ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, want.numpy())


# ---------------------------------------------------------------------------------------------------------
# conv head
# ---------------------------------------------------------------------------------------------------------
def _conv_reference(feat, w0, b0, w2, b2):
    """Conv head with the forward kernel's precision convention (fp16 operands and hidden activations, fp32
    accumulation; straight-through rounding) - sam_model.py:202-208,260-265."""
    import torch.nn.functional as F

    from oracle import tcnn_spec as T

    x = T.f16(feat).reshape(-1, 4, 4, 256).permute(0, 3, 1, 2)
    h = T.f16(F.relu(F.conv2d(x, T.f16(w0), b0, padding=1)))
    y = F.conv2d(h, T.f16(w2), b2, padding=1)
    return y.mean(dim=[2, 3])


def test_conv_head_kernel_bodies_match_autograd():
    from emu.build_emu import load

    cfg, params, orc = model_pair("tiny", "scene", 5, False, 4)
    P = 3
    gen = torch.Generator().manual_seed(7)
    feat = (torch.randn(P * 16, 256, generator=gen) * 0.3).requires_grad_(True)
    names = ["conv_head.0.weight", "conv_head.0.bias", "conv_head.2.weight", "conv_head.2.bias"]
    w0, b0, w2, b2 = [params[k].clone().requires_grad_(True) for k in names]
    out = _conv_reference(feat, w0, b0, w2, b2)
    # the precision convention stays within the stated tolerance of the reference's fp32 conv
    assert torch.allclose(out.detach(), orc.patch_aggregate(feat.detach()), rtol=2e-2, atol=3e-3)
    g_out = torch.randn(P, 256, generator=gen)
    (out * g_out).sum().backward()

    arr = lambda t: np.ascontiguousarray(t.detach().numpy(), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    g = [np.zeros(t.numel(), np.float32) for t in (w0, b0, w2, b2)]
    d_feat = np.zeros((P * 16, 256), np.float32)
    fwd = np.zeros((P, 256), np.float32)
    keep = [arr(feat), arr(g_out), _f16_bits(w0.reshape(256, -1)), arr(b0), _f16_bits(w2.reshape(256, -1)), arr(b2)]
    load().emu_conv_backward(ptr(keep[0]), C.c_longlong(P), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), ptr(keep[4]), ptr(keep[5]),
                             ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(d_feat), ptr(fwd))
    assert torch.allclose(torch.from_numpy(fwd), out.detach(), rtol=1e-4, atol=1e-5)
    for got, want, what in zip(g, (w0, b0, w2, b2), names):
        _assert_grad_close(got, want.grad, "d " + what)
    _assert_grad_close(d_feat, feat.grad, "d feat")


@pytest.mark.gpu
def test_gpu_conv_head_backward_matches_autograd():
    from helpers import make_renderer

    cfg, params, orc = model_pair("tiny", "scene", 5, False, 4)
    r = make_renderer(cfg, params)
    P = 300  # more than one internal block of 256 patches
    gen = torch.Generator().manual_seed(7)
    feat = (torch.randn(P * 16, 256, generator=gen) * 0.3).requires_grad_(True)
    names = ["conv_head.0.weight", "conv_head.0.bias", "conv_head.2.weight", "conv_head.2.bias"]
    ws = [params[k].clone().requires_grad_(True) for k in names]
    out = _conv_reference(feat, *ws)
    g_out = torch.randn(P, 256, generator=gen)
    (out * g_out).sum().backward()
    fwd = r.patch_aggregate(feat.detach())
    assert torch.allclose(fwd.cpu(), out.detach(), rtol=2e-2, atol=3e-3)
    g, d_feat = r.patch_aggregate_backward(feat.detach(), g_out)
    torch.cuda.synchronize()
    for name, w in zip(names, ws):
        _assert_grad_close(g[name].cpu(), w.grad, "d " + name)
    _assert_grad_close(d_feat.cpu(), feat.grad, "d feat")


def test_pick_samples_body_matches_reference_topk():
    """Stand-alone top-k + sharpen (sam_model.py:244-255) against torch.topk / pow / normalise; slot order is
    descending weight, which is what the bucketed forward kernel and the compact backward rows rely on."""
    from emu.build_emu import load

    n, S, k, T = 500, 32, 16, 10.0
    gen = torch.Generator().manual_seed(3)
    w = torch.rand(n, S, generator=gen) ** 6
    w[::5, 20:] = 0.0                       # saturated rays: exact ties at zero, broken by sample index
    starts = torch.rand(n, S, generator=gen)
    ends = starts + torch.rand(n, S, generator=gen)
    arr = lambda t: np.ascontiguousarray(t.numpy(), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    keep = [arr(w), arr(starts), arr(ends)]
    sam_t, sam_w = np.zeros((n, k), np.float32), np.zeros((n, k), np.float32)
    load().emu_pick_samples(ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), C.c_longlong(n), S, k, C.c_float(T), ptr(sam_t), ptr(sam_w))
    ref_w, ids = torch.topk(w, k, dim=-1, sorted=True)  # values descending; the reference uses sorted=False (a set)
    ref_w = ref_w ** T
    ref_w = ref_w / ref_w.sum(-1, keepdim=True)
    got_w, got_t = torch.from_numpy(sam_w), torch.from_numpy(sam_t)
    assert torch.allclose(got_w, ref_w, rtol=1e-5, atol=1e-12)
    assert bool((got_w[:, :-1] >= got_w[:, 1:]).all())
    # slot = rank by (weight descending, sample index ascending): ties - e.g. the zero-weight tail of a saturated ray -
    # go to the lower index, exactly like the march kernel; torch.topk agrees as a set wherever the k-th weight is untied
    ids_ranked = torch.sort(-w, dim=-1, stable=True).indices[:, :k]
    assert torch.equal(got_t, torch.gather(starts + ends, 1, ids_ranked))
    untied = ref_w[:, -1] > 0
    assert untied.any() and (~untied).any()
    assert torch.equal(ids_ranked[untied].sort(-1).values, ids[untied].sort(-1).values)
