"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/mode_engine.h declares, the
product path fails loudly without a GPU (no fallback), the nn.Module mirror has the reference's state_dict contract,
and the schedule / sampler host logic matches the oracle (which is pinned to the reference by tests/test_oracle.py)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import mode_oracle as O
from mode_diffusion_policy_b200 import _lib, gc_sampling as S
from mode_diffusion_policy_b200.modedit import MoDeDiT, NoiseBlockMoE
from mode_diffusion_policy_b200.score_wrappers import GCDenoiser

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "mode_engine.h").read_text()
    declared = set(re.findall(r"\b(mode_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mode_engine.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_config_struct_layout_matches_header():
    header = (ROOT / "include" / "mode_engine.h").read_text()
    body = header[header.index("typedef struct mode_config {"):header.index("} mode_config_t;")]
    fields = re.findall(r"^\s*(?:int32_t|float)\s+([a-z_]+);", body, flags=re.M)
    assert fields == [f[0] for f in _lib.mode_config_t._fields_]
    assert ctypes.sizeof(_lib.mode_config_t) == 4 * len(fields)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    lib = _lib.load()
    cfg = _lib.mode_config_t(obs_dim=128, goal_dim=64, action_dim=7, embed_dim=256, n_layers=1, n_heads=4,
                             n_state_tokens=2, action_seq_len=10, num_experts=4, top_k=2, router_normalize=1,
                             max_batch=4, sigma_data=0.5, rms_eps=1e-6)
    h = ctypes.c_void_p()
    rc = lib.mode_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and not h.value  # MODE_ERR_CUDA, no engine
    assert lib.mode_last_error()
    from mode_diffusion_policy_b200.engine import EngineConfig, ModeEngine

    with pytest.raises(_lib.ModeError):
        ModeEngine(EngineConfig(max_batch=2))
    m = MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256,
                embed_pdrob=0, attn_pdrop=0.3, n_layers=1, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10,
                state_dim=7).eval()
    with pytest.raises(_lib.ModeError):  # no silent PyTorch path behind the module
        m({"state_images": torch.zeros(1, 2, 128)}, torch.zeros(1, 10, 7), torch.zeros(1, 1, 64), torch.ones(1))


def test_module_mirror_has_reference_state_dict_contract():
    cfg = O.ModeConfig(obs_dim=128, goal_dim=64, embed_dim=256, n_layers=2, n_heads=4, num_experts=4)
    m = MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256,
                embed_pdrob=0, attn_pdrop=0.3, n_layers=2, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10,
                state_dim=7, num_experts=4, top_k=2, init_style="olmoe")
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == [(n, tuple(s)) for n, s in O.state_dict_spec(cfg)]  # names, shapes AND order (EMA zips by position)
    den = GCDenoiser(m, sigma_data=0.5)
    assert [k for k in den.state_dict()] == ["inner_model." + n for n, _ in O.state_dict_spec(cfg)]
    blocks = [mod for mod in den.modules() if isinstance(mod, NoiseBlockMoE)]
    assert len(blocks) == 2 and blocks[0].total_tokens_processed == 0
    assert torch.equal(blocks[0].get_expert_usage(), torch.zeros(4))
    names = dict(m.named_parameters())
    assert all(n in names for n in ("blocks.0.router.router.mlp.3.bias", "pos_emb", "out.bias"))
    m.freeze_router()
    assert not any(p.requires_grad for p in m.blocks[0].router.parameters())
    assert m.get_router_states()[1]["frozen"]
    with pytest.raises(NotImplementedError):
        MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256,
                embed_pdrob=0, attn_pdrop=0.3, n_layers=1, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10,
                state_dim=7, use_proprio=True)


def test_schedules_match_oracle_and_reference_values():
    s = S.get_sigmas_exponential(10, 1e-3, 80.0).numpy()
    np.testing.assert_allclose(s, O.get_sigmas_exponential(10, 1e-3, 80.0), rtol=2e-6)
    # SURVEY.md §8a a3 [probe of the reference]
    np.testing.assert_allclose(s, [80, 22.82, 6.509, 1.857, 0.5296, 0.1511, 0.04309, 0.01229, 0.003506, 0.001, 0], rtol=2e-3)
    np.testing.assert_allclose(S.get_sigmas_karras(10, 1e-3, 80.0).numpy(), O.get_sigmas_karras(10, 1e-3, 80.0), rtol=1e-5)
    np.testing.assert_allclose(S.get_sigmas_linear(5, 1e-3, 80.0).numpy(), O.get_sigmas_linear(5, 1e-3, 80.0), rtol=1e-6)
    for fn in (S.get_sigmas_vp, S.cosine_beta_schedule):
        v = fn(8)
        assert v.shape == (9,) and v[-1] == 0


class _ToyDenoiser:
    """D(x; sigma) = x / (1 + sigma^2): the exact denoiser of N(0, 1) data — lets the sampler logic run on CPU."""

    def __call__(self, state, x, goal, sigma, **kw):
        return x / (1 + sigma.view(-1, 1, 1) ** 2)


def test_samplers_host_logic():
    torch.manual_seed(0)
    model = _ToyDenoiser()
    sig = S.get_sigmas_exponential(10, 1e-3, 80.0)
    x0 = torch.randn(4, 10, 7) * 80.0
    # ddim against the oracle's coefficient table (pinned to the reference through the goldens)
    x = x0.numpy().copy()
    for i, (ratio, em1) in enumerate(O.ddim_coefficients(sig.numpy())):
        den = x / (1 + sig[i].item() ** 2)
        x = (ratio * x - em1 * den).astype(np.float32)
    got = S.sample_ddim(model, None, x0, None, sig)
    np.testing.assert_allclose(got.numpy(), x, rtol=2e-5, atol=1e-6)
    # every deterministic sampler lands on (numerically) the same sample of the probability-flow ODE
    ref = S.sample_heun(model, None, x0, None, S.get_sigmas_exponential(200, 1e-3, 80.0)).numpy()
    for name in ("euler", "heun", "dpm", "dpmpp_2m", "dpmpp_2s", "lms", "ddim", "dpmpp_2_with_lms"):
        out = S.SAMPLERS[name](model, None, x0, None, S.get_sigmas_exponential(60, 1e-3, 80.0)).numpy()
        assert np.abs(out - ref).max() < 0.15 * np.abs(ref).max() + 1e-3, name
    for name in ("euler_ancestral", "ancestral", "dpmpp_2s_ancestral"):
        out = S.SAMPLERS[name](model, None, x0, None, sig)
        assert torch.isfinite(out).all() and out.shape == x0.shape
    calls = []
    S.sample_ddim(model, None, x0, None, sig, callback=lambda d: calls.append(d["i"]))
    assert calls == list(range(10))
    down, up = S.get_ancestral_step(2.0, 1.0)
    assert abs(down ** 2 + up ** 2 - 1.0) < 1e-6


def test_checkpoint_loader_selects_and_remaps_denoiser_keys(tmp_path):
    """model_cleaned.safetensors is keyed as MoDEAgent.state_dict() (reference mode_agent.py:134-265): denoiser tensors
    under model.inner_model.*, encoders / CLIP beside them. Round trip through the writer, non-strict skip of a shape
    mismatch, strict failure, and the directory form."""
    from safetensors.torch import save_file
    from mode_diffusion_policy_b200 import checkpoint as CK

    def make(seed):
        torch.manual_seed(seed)
        return MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256,
                       embed_pdrob=0, attn_pdrop=0.0, n_layers=2, n_heads=4, goal_seq_len=1, obs_seq_len=1,
                       action_seq_len=10, state_dim=7, num_experts=4, top_k=2, init_style="olmoe")

    src, dst = make(1), make(2)
    with torch.no_grad():
        src.pos_emb.normal_()
    agent_sd = {"model.inner_model." + k: v.detach().clone() for k, v in src.state_dict().items()}
    agent_sd["static_resnet.resnet.conv1.weight"] = torch.zeros(4, 3, 7, 7)    # encoder: not ours
    agent_sd["language_goal.clip_rn50.visual.proj"] = torch.zeros(8, 8)        # CLIP: skipped by the reference too
    agent_sd["model.inner_model.out.bias"] = torch.zeros(9)                    # incompatible shape
    agent_sd["model.inner_model.pos_emb"] = src.pos_emb.detach()[0].clone()    # same data, saved without the batch dim
    d = tmp_path / "ckpt"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in agent_sd.items()}, str(d / "model_cleaned.safetensors"))
    rep = CK.load_pretrained_parameters(dst, str(d))
    assert rep.ignored == 2 and rep.missing == ["out.bias"]
    assert rep.skipped_shape == [("out.bias", (9,), (7,))]
    for k, v in src.state_dict().items():
        if k != "out.bias":
            assert torch.equal(dst.state_dict()[k], v), k
    with pytest.raises(RuntimeError):
        CK.load_pretrained_parameters(make(3), str(d), strict=True)
    # writer -> loader round trip on a single file, strict
    f = tmp_path / "model_cleaned.safetensors"
    CK.save_denoiser(src, str(f))
    again = make(4)
    rep = CK.load_pretrained_parameters(again, str(f), strict=True)
    assert not rep.missing and all(torch.equal(again.state_dict()[k], v) for k, v in src.state_dict().items())
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        CK.load_pretrained_parameters(make(5), str(empty))


def test_train_rng_bookkeeping_and_stochastic_arguments():
    """Host side of the stochastic training mode: the constructor's regularisation settings reach the engine call,
    the (seed, step) position advances once per step, and `deterministic_training` switches everything off."""
    m = MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256,
                embed_pdrob=0.05, attn_pdrop=0.3, n_layers=2, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10,
                state_dim=7, mlp_pdrop=0.1, goal_drop=0.1, num_experts=4, top_k=2, use_argmax=False)
    m.set_train_rng(1234, 7)
    a = m._stochastic_args()
    assert a == dict(attn_pdrop=0.3, mlp_pdrop=0.1, goal_drop=0.1, embed_pdrop=0.05, multinomial=True, seed=1234, step=7)
    m._advance_train_rng()
    assert m._stochastic_args()["step"] == 8
    m.deterministic_training = True
    off = m._stochastic_args()
    assert not any(off[k] for k in ("attn_pdrop", "mlp_pdrop", "goal_drop", "embed_pdrop", "multinomial"))
    m.deterministic_training = False
    # default seed: derived from torch's seed (every process / rank its own stream), fixed once chosen
    m2 = MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256, embed_pdrob=0,
                 attn_pdrop=0.3, n_layers=1, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10, state_dim=7,
                 use_argmax=True)
    s1 = m2._stochastic_args()
    assert s1["seed"] == m2._stochastic_args()["seed"] and s1["step"] == 0 and s1["multinomial"] is False


def test_optimizer_groups_match_the_reducers_per_layer_buckets():
    """`mode_adamw_step_group` puts block l's tensors of >= 2^20 elements in group l and everything else in the last
    group; parallel.plan_grad_buckets (default min_bucket 2^20) must make exactly those tensors the per-layer buckets,
    so that a block's update never starts before all of its gradients have been exchanged. Checked on the reference
    layout of the CALVIN model: flat buffer by tensor kind, like the engine's."""
    from mode_diffusion_policy_b200 import parallel

    cfg = O.ModeConfig()
    spec = [(n, int(np.prod(s))) for n, s in O.state_dict_spec(cfg) if n != "gripper_embed.weight"]
    # kind-major layout: tensors of the same kind (name without the block index) are contiguous over the blocks
    def kind(n):
        parts = n.split(".")
        return ".".join(parts[2:]) if parts[0] == "blocks" else n
    order = sorted(range(len(spec)), key=lambda i: (kind(spec[i][0]), spec[i][0].startswith("blocks.") and int(spec[i][0].split(".")[1])))
    off, ranges = 0, {}
    for i in order:
        n, numel = spec[i]
        ranges[n] = (off, numel)
        off += (numel + 31) // 32 * 32
    per_layer = [[] for _ in range(cfg.n_layers)]
    for n, r in ranges.items():
        if n.startswith("blocks."):
            per_layer[int(n.split(".")[1])].append(r)
    buckets, tail = parallel.plan_grad_buckets(per_layer, off)
    for layer in range(cfg.n_layers):
        big = sorted(r for n, r in ranges.items() if n.startswith(f"blocks.{layer}.") and r[1] >= 1 << 20)
        covered = sorted(buckets[layer])
        assert sum(n for _, n in covered) == sum(n for _, n in big)
        for o, n in big:  # every large tensor of the block lies inside one of its buckets
            assert any(bo <= o and o + n <= bo + bn for bo, bn in covered)
        small = [r for n, r in ranges.items() if n.startswith(f"blocks.{layer}.") and r[1] < 1 << 20]
        for o, n in small:  # ...and no small one does: they ride in the tail / the last optimizer group
            assert not any(bo <= o < bo + bn for bo, bn in covered)
            assert any(to <= o and o + n <= to + tn for to, tn in tail)


def test_fused_sampler_dispatch_conditions():
    """sample_euler / sample_dpmpp_2m hand the whole loop to the engine only when nothing needs the per-step python
    loop (no churn, callback, scaler or extra_args)."""
    class Fused(_ToyDenoiser):
        def __init__(self):
            super().__init__()
            self.calls = []

        def sample_fused(self, name, state, action, goal, sigmas):
            self.calls.append(name)
            return action

        def sample_ddim(self, state, action, goal, sigmas):
            self.calls.append("ddim")
            return action

    sig = S.get_sigmas_exponential(5, 1e-3, 80.0)
    x0 = torch.randn(2, 10, 7)
    m = Fused()
    S.sample_euler(m, None, x0, None, sig)
    S.sample_dpmpp_2m(m, None, x0, None, sig)
    S.sample_ddim(m, None, x0, None, sig)
    assert m.calls == ["euler", "dpmpp_2m", "ddim"]
    S.sample_euler(m, None, x0, None, sig, s_churn=1.0)
    S.sample_euler(m, None, x0, None, sig, callback=lambda d: None)
    S.sample_dpmpp_2m(m, None, x0, None, sig, extra_args={"uncond": False})
    assert m.calls == ["euler", "dpmpp_2m", "ddim"]  # all three ran the python loop


def test_training_noise_densities_match_reference_draws():
    """every `sigma_sample_density_type` of MoDEAgent.make_sample_density (mode_agent.py:692-731) —
    the same torch seed must give the same bits as the reference's utils functions (tests/golden/sample_densities.npz,
    generated by tests/golden/make_density_goldens.py), and the policy wrapper dispatches to them with the reference's
    parameters."""
    import math
    from pathlib import Path

    from mode_diffusion_policy_b200 import utils as U
    from mode_diffusion_policy_b200.agent import DenoisingPolicy

    g = np.load(Path(__file__).resolve().parent / "golden" / "sample_densities.npz")
    cases = {
        "lognormal": lambda: U.rand_log_normal((64,), loc=-1.2, scale=1.2),
        "loglogistic": lambda: U.rand_log_logistic((64,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0),
        "loguniform": lambda: U.rand_log_uniform((64,), min_value=0.001, max_value=80.0),
        "uniform": lambda: U.rand_uniform((64,), min_value=0.001, max_value=80.0),
        "v-diffusion": lambda: U.rand_v_diffusion((64,), sigma_data=0.5, min_value=0.001, max_value=80.0),
        "split-lognormal": lambda: U.rand_split_log_normal((64,), loc=-1.2, scale_1=0.8, scale_2=1.4),
        "discrete": lambda: U.rand_discrete((64,), values=torch.linspace(0.001, 80.0, 1000)),
    }
    for name, fn in cases.items():
        torch.manual_seed(1234)
        assert np.array_equal(fn().numpy(), g[name]), name
    # the wrapper's dispatch (reference parameters: loc = ln sigma_data, scale 0.5, [sigma_min, sigma_max], ...)
    for kind in ("lognormal", "loglogistic", "loguniform", "uniform", "v-diffusion"):
        pol = DenoisingPolicy(model=None, sigma_sample_density_type=kind, device="cpu")
        torch.manual_seed(1234)
        got = pol.make_sample_density()(shape=(64,), device="cpu").numpy()
        assert np.array_equal(got, g[kind]), kind
    pol = DenoisingPolicy(model=None, sigma_sample_density_type="discrete", device="cpu", num_sampling_steps=1e-3)
    s = pol.make_sample_density()(shape=(16,), device="cpu")  # a table of 100 exponential noise levels (+ the final 0)
    assert s.shape == (16,) and float(s.max()) <= 80.0 + 1e-4
    with pytest.raises(KeyError):
        DenoisingPolicy(model=None, sigma_sample_density_type="split-lognormal", device="cpu").make_sample_density()
    with pytest.raises(ValueError):
        DenoisingPolicy(model=None, sigma_sample_density_type="nope", device="cpu").make_sample_density()
    # truncated log-logistic: inside [sigma_min, sigma_max], median at sigma_data
    torch.manual_seed(0)
    x = U.rand_log_logistic((20000,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0)
    assert float(x.min()) >= 0.001 and float(x.max()) <= 80.0 and abs(float(x.median()) - 0.5) < 0.02


def test_synthetic_workload_equals_the_oracles_generators():
    """bench.py's engine arm takes its model shape, random weights, inputs and noise schedule from synthetic_workload.py
    (it must not touch oracle/); the generators are the oracle's, bit for bit, so the benchmark runs on the numbers the
    parity tests check."""
    import synthetic_workload as W

    kw = dict(embed_dim=256, n_layers=2, n_heads=4, obs_dim=128, goal_dim=64, num_experts=4)
    cw, co = W.ModeConfig(**kw), O.ModeConfig(**kw)
    assert W.ModeConfig() == W.ModeConfig() and vars(W.ModeConfig()) == vars(O.ModeConfig()) and cw.seq_len == co.seq_len
    assert W.state_dict_spec(cw) == O.state_dict_spec(co)
    a, b = W.make_weights_fast(cw, seed=1234), O.make_weights_fast(co, seed=1234)
    assert list(a) == list(b) and all(np.array_equal(a[k], b[k]) for k in a)
    for x, y in zip(W.make_inputs(cw, 3, seed=7), O.make_inputs(co, 3, seed=7)):
        assert np.array_equal(x, y)
    assert np.array_equal(W.get_sigmas_exponential(10, 1e-3, 80.0), O.get_sigmas_exponential(10, 1e-3, 80.0))
    # the engine arm of bench.py does not import the oracle
    src = (Path(__file__).resolve().parents[1] / "bench.py").read_text()
    main_src = src[src.index("def main():"):]
    assert "oracle" not in main_src.replace("oracle/ (the checker)", "")


def test_sharded_optimizer_refuses_to_hand_out_stale_masters():
    """After `EngineAdamW.step_sharded(master_sync="lazy")` the fp32 masters of the sharded tensors are current only on
    their owning rank: `state_dict()` (checkpoints, the EMA callback, the engine's own re-pack) must raise until
    `synchronize_parameters()` has gathered them, instead of silently returning half-updated tensors."""
    from mode_diffusion_policy_b200.optim import EngineAdamW
    from mode_diffusion_policy_b200.score_wrappers import GCDenoiser

    m = MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256, embed_pdrob=0,
                attn_pdrop=0.0, n_layers=1, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10, state_dim=7,
                use_argmax=True)
    wrapper = GCDenoiser(m, sigma_data=0.5)
    opt = EngineAdamW(m, lr=1e-4)
    assert "pos_emb" in m.state_dict() and any(k.endswith("pos_emb") for k in wrapper.state_dict())
    opt._masters_stale = True  # what a lazy sharded step leaves behind
    with pytest.raises(RuntimeError, match="synchronize_parameters"):
        m.state_dict()
    with pytest.raises(RuntimeError, match="synchronize_parameters"):
        wrapper.state_dict()  # the reference checkpoints the wrapper / the whole agent
    opt._masters_stale = False
    assert "pos_emb" in m.state_dict()
