"""The native multi-tensor AdamW (cpt_adamw_step through cpt_b200.optimization.AdamW) against the CPU oracle, in both
semantics, with parameter groups, tensors that are not a multiple of the chunk size, parameters without gradient, a
schedule changing lr between steps, and the gradient scale.  Tolerance: fp32 arithmetic in a different association
order (fused multiply-adds) — 2e-6 relative to the parameter scale after 6 steps."""
import pytest
import torch

from oracle import optim_oracle as OO

pytestmark = pytest.mark.gpu


def make_params(gen):
    shapes = [(1000, 33), (16384,), (16385,), (7,), (300, 128), (1,)]
    return [torch.randn(*s, generator=gen) for s in shapes]


@pytest.mark.parametrize("torch_semantics", [False, True])
@pytest.mark.parametrize("correct_bias", [True, False])
def test_adamw_against_oracle(torch_semantics, correct_bias):
    from cpt_b200.optimization import AdamW, WarmupLinearSchedule
    if torch_semantics and not correct_bias:
        pytest.skip("torch.optim.AdamW always corrects the bias")
    gen = torch.Generator().manual_seed(3)
    host = make_params(gen)
    params = [torch.nn.Parameter(t.clone().cuda()) for t in host]
    frozen = torch.nn.Parameter(torch.ones(5).cuda())  # never receives a gradient
    groups = [{"params": params[:3] + [frozen], "weight_decay": 0.05}, {"params": params[3:], "weight_decay": 0.0}]
    eps = 1e-8 if torch_semantics else 1e-6
    opt = AdamW(groups, lr=2e-3, betas=(0.9, 0.98), eps=eps, correct_bias=correct_bias,
                torch_semantics=torch_semantics)
    sched = WarmupLinearSchedule(opt, warmup_steps=2, t_total=10)
    ref, states = [t.clone() for t in host], [dict() for _ in host]
    for step in range(6):
        lr = opt.param_groups[0]["lr"]
        for i, p in enumerate(params):
            g = torch.randn(*host[i].shape, generator=gen) * (0.3 + step)
            p.grad = g.cuda()
            wd = 0.05 if i < 3 else 0.0
            if torch_semantics:
                OO.adamw_torch(ref[i], g, states[i], lr, (0.9, 0.98), eps, wd)
            else:
                OO.adamw_hf1(ref[i], g, states[i], lr, (0.9, 0.98), eps, wd, correct_bias)
        opt.step()
        sched.step()
    torch.cuda.synchronize()
    for i, p in enumerate(params):
        err = (p.detach().cpu() - ref[i]).abs().max().item()
        assert err <= 2e-6 * max(1.0, ref[i].abs().max().item()), (i, err)
        st = opt.state[p]
        assert st["step"] == 6
        assert (st["exp_avg"].cpu() - states[i]["exp_avg"]).abs().max().item() < 1e-5
        assert (st["exp_avg_sq"].cpu() - states[i]["exp_avg_sq"]).abs().max().item() < 1e-4
    assert torch.equal(frozen.detach().cpu(), torch.ones(5)) and len(opt.state[frozen]) == 0


def test_adamw_grad_scale_and_state_dict_roundtrip():
    from cpt_b200.optimization import AdamW
    gen = torch.Generator().manual_seed(5)
    p = torch.nn.Parameter(torch.randn(5000, generator=gen).cuda())
    q = torch.nn.Parameter(p.detach().clone())
    g = torch.randn(5000, generator=gen).cuda()
    a, b = AdamW([p], lr=1e-2), AdamW([q], lr=1e-2)
    p.grad = g * 128.0
    v0 = p._version
    a.step(grad_scale=torch.tensor(1.0 / 128.0, device="cuda"))
    assert p._version > v0  # the in-place update must be visible to version-based caches (engine weight refresh)
    q.grad = g.clone()
    b.step()
    assert (p - q).abs().max().item() < 1e-6
    sd = a.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    c = AdamW([p], lr=1e-2)
    c.load_state_dict(sd)
    p.grad = g.clone()
    c.step()
    q.grad = g.clone()
    b.step()
    assert (p - q).abs().max().item() < 1e-6


def test_adamw_rejects_cpu_parameters():
    from cpt_b200.optimization import AdamW
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        AdamW([p]).step()


def test_few_shot_loop_with_native_optimizer():
    """the GQA few-shot loop shape (gqa_cpt.py:428-462): grouped parameters, clip_grad_norm_, scheduler, zero_grad"""
    from cpt_b200 import config as C
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    from cpt_b200.optimization import AdamW, WarmupConstantSchedule
    from cpt_b200.synthetic import synth_batch, synth_state_dict
    cfg = C.oscar_tiny(num_hidden_layers=2)
    cfg.hidden_dropout_prob = 0.1
    sd = synth_state_dict(cfg, seed=12)
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    model = REC_MLM_CPT(cfg)
    model.copy_from_pretraining_model(pre.cuda())
    no_decay = ["bias", "LayerNorm.weight"]
    groups = [{"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)],
               "weight_decay": 0.05},
              {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)],
               "weight_decay": 0.0}]
    opt = AdamW(groups, lr=5e-4, eps=1e-8)
    sched = WarmupConstantSchedule(opt, warmup_steps=2)
    b = synth_batch(cfg, 8, 40, 20, seed=4)
    d = {k: v.cuda() for k, v in b.items()}
    labels = torch.full((8, 60), -1, dtype=torch.long)
    labels[torch.arange(8), b["mask_pos"]] = torch.arange(8) % 3 + 20
    labels = labels.cuda()
    torch.manual_seed(0)
    losses = []
    probe = dict(model.named_parameters())["bert.encoder.layer.1.output.dense.weight"]
    w0 = probe.detach().clone()
    for _ in range(12):
        model.train()
        loss, _ = model(input_ids=d["input_ids"], attention_mask=d["attention_mask"],
                        token_type_ids=d["token_type_ids"], masked_lm_labels=labels, img_feats=d["img_feats"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        sched.step()
        opt.step()
        model.zero_grad()
        losses.append(loss.item())
    assert all(l == l for l in losses) and losses[-1] < losses[0] - 0.3, losses  # finite and going down
    # the training handle must have picked up every optimizer step: its forward on the final weights equals a fresh
    # model's forward on a copy of them (a stale 16-bit copy would give the step-0 loss instead)
    assert (probe.detach() - w0).abs().max().item() > 0
    model.eval()
    cfg.hidden_dropout_prob = 0.0
    with torch.enable_grad():
        l_trained = model(input_ids=d["input_ids"], attention_mask=d["attention_mask"],
                          token_type_ids=d["token_type_ids"], masked_lm_labels=labels, img_feats=d["img_feats"])[0]
    with torch.no_grad():
        scores = model(input_ids=d["input_ids"], attention_mask=d["attention_mask"],
                       token_type_ids=d["token_type_ids"], img_feats=d["img_feats"])[0]
        l_infer = torch.nn.functional.cross_entropy(scores.view(-1, cfg.vocab_size), labels.view(-1), ignore_index=-1)
    assert abs(l_trained.item() - l_infer.item()) <= 5e-3 * abs(l_infer.item()), (l_trained.item(), l_infer.item())


class _DataMutatingAdamW(torch.optim.Optimizer):
    """pytorch-transformers 1.x AdamW as the reference's GQA / VCR loops import it (gqa_cpt.py:24,342): every update
    goes through `p.data`, which does NOT bump the autograd version counter of `p`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.data
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p.data)
                    state["exp_avg_sq"] = torch.zeros_like(p.data)
                b1, b2 = group["betas"]
                state["step"] += 1
                state["exp_avg"].mul_(b1).add_(grad, alpha=1.0 - b1)
                state["exp_avg_sq"].mul_(b2).addcmul_(grad, grad, value=1.0 - b2)
                denom = state["exp_avg_sq"].sqrt().add_(group["eps"])
                step_size = group["lr"] * (1.0 - b2 ** state["step"]) ** 0.5 / (1.0 - b1 ** state["step"])
                p.data.addcdiv_(state["exp_avg"], denom, value=-step_size)
                if group["weight_decay"] > 0.0:
                    p.data.add_(p.data, alpha=-group["lr"] * group["weight_decay"])


@pytest.mark.parametrize("how", ["optimizer", "by_hand"])
def test_weight_updates_through_p_data_are_picked_up(how):
    """`p.data.add_` leaves `p._version` untouched; both handles must still follow the update — after
    an optimizer step (global post-step hook) and after a hand-written update that follows a backward (dirty flag)."""
    from cpt_b200 import config as C
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids
    cfg = C.oscar_tiny(num_hidden_layers=2)
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.0
    sd = synth_state_dict(cfg, seed=21)
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    model = REC_MLM_CPT(cfg)
    model.copy_from_pretraining_model(pre.cuda())
    b = synth_batch(cfg, 4, 30, 12, seed=2)
    d = {k: v.cuda() for k, v in b.items()}
    vids = synth_vocab_ids(cfg, 4, seed=1).cuda()
    labels = torch.full((4, 42), -1, dtype=torch.long)
    labels[torch.arange(4), b["mask_pos"]] = torch.arange(4) % 3 + 20
    labels = labels.cuda()
    opt = _DataMutatingAdamW(model.parameters(), lr=5e-3) if how == "optimizer" else None

    def infer(m):
        m.eval()
        with torch.no_grad():
            return m(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                     mask_pos=d["mask_pos"], vocab_ids=vids)[0].clone()

    def loss_of(m):
        m.train()
        return m(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                 masked_lm_labels=labels)[0]

    before = infer(model)
    versions = [p._version for p in model.parameters()]
    for _ in range(3):
        loss = loss_of(model)
        loss.backward()
        if opt is not None:
            opt.step()
        else:
            for p in model.parameters():
                if p.grad is not None:
                    p.data.add_(p.grad, alpha=-0.05)
        model.zero_grad()
    assert versions == [p._version for p in model.parameters()]  # the premise: nothing bumped the counters
    after, after_loss = infer(model), float(loss_of(model))
    fresh_pre = BertImgForPreTraining(cfg)
    fresh_pre.load_state_dict({k: v.detach().cpu() for k, v in pre.state_dict().items()}, strict=False)
    fresh_pre.tie_weights()
    fresh = REC_MLM_CPT(cfg)
    fresh.copy_from_pretraining_model(fresh_pre.cuda())
    want, want_loss = infer(fresh), float(loss_of(fresh))
    assert (before - after).abs().max().item() > 1e-3          # the weights did move
    assert (after - want).abs().max().item() <= 1e-5 * want.abs().max().item()   # the inference handle followed them
    assert abs(after_loss - want_loss) <= 1e-6 * max(1.0, abs(want_loss))   # as did the training handle


@pytest.mark.parametrize("max_norm,gs", [(1.0, None), (0.05, None), (1e6, None), (1.0, 1.0 / 64.0)])
def test_fused_global_norm_clip_matches_clip_grad_norm(max_norm, gs):
    """optimizer.step(max_grad_norm=m) == torch.nn.utils.clip_grad_norm_(params, m); optimizer.step()
    (gqa_cpt.py:454-456): same norm, same update, gradients left untouched."""
    from cpt_b200.optimization import AdamW
    gen = torch.Generator().manual_seed(9)
    host = make_params(gen)
    grads = [torch.randn(*t.shape, generator=gen) * (0.01 + 0.3 * i) for i, t in enumerate(host)]

    def fresh():
        ps = [torch.nn.Parameter(t.clone().cuda()) for t in host]
        for p, g in zip(ps, grads):
            p.grad = (g / gs if gs else g).clone().cuda()
        groups = [{"params": ps[:3], "weight_decay": 0.05}, {"params": ps[3:], "weight_decay": 0.0, "betas": (0.8, 0.9)}]
        return ps, AdamW(groups, lr=1e-2)

    scale = torch.tensor(gs, device="cuda") if gs else None
    a_params, a = fresh()
    before = [p.grad.clone() for p in a_params]
    a.step(max_grad_norm=max_norm, grad_scale=scale)
    b_params, b = fresh()
    if gs:
        for p in b_params:
            p.grad.mul_(gs)
    want_norm = torch.nn.utils.clip_grad_norm_(b_params, max_norm)
    b.step()
    torch.cuda.synchronize()
    assert abs(float(a.last_grad_norm) - float(want_norm)) <= 2e-6 * float(want_norm)
    for pa, pb, g0 in zip(a_params, b_params, before):
        assert torch.equal(pa.grad, g0)   # the fused form never writes the gradients
        assert (pa.detach() - pb.detach()).abs().max().item() <= 2e-6 * max(1.0, pb.detach().abs().max().item())
    a.step(max_grad_norm=max_norm, grad_scale=scale)   # the scratch was left zeroed: a second step sees the same norm
    torch.cuda.synchronize()
    assert abs(float(a.last_grad_norm) - float(want_norm)) <= 2e-6 * float(want_norm)
