"""CPU tests of the wire formats around the hot path: the pickle-over-HTTP agent proxy
(reference mode/evaluation/agent_proxy.py:91-203) with its batching front-end, and the Hugging Face export
(reference mode/utils/save_to_hf.py:98-157)."""
import json
import pickle
import threading
import urllib.error
import urllib.request

import numpy as np
import pytest
import torch

from mode_diffusion_policy_b200 import agent_proxy as AP
from mode_diffusion_policy_b200 import checkpoint as CK
from mode_diffusion_policy_b200 import save_to_hf as HF


class _Agent:
    def __init__(self):
        self.steps = 0

    def __call__(self, obs, goal):
        return {"sum": float(np.sum(obs["x"])), "goal": goal}

    def step(self, obs, lang):
        self.steps += 1
        return np.asarray(obs, np.float32) * 2, lang

    def reset(self):
        self.steps = 0
        return None

    def boom(self):
        raise ValueError("agent failure")


def _serve(create):
    server = AP.make_server(create, "127.0.0.1", 0)
    t = threading.Thread(target=server.serve_forever, daemon=True)
    t.start()
    return server, t


def _raw_post(port, payload: bytes):
    req = urllib.request.Request(f"http://127.0.0.1:{port}", data=payload, method="POST")
    try:
        with urllib.request.urlopen(req, timeout=10) as r:
            return r.status, pickle.loads(r.read())
    except urllib.error.HTTPError as e:
        return e.code, pickle.loads(e.read())


def test_agent_proxy_speaks_the_reference_wire_format():
    created = []
    server, thread = _serve(lambda: created.append(1) or _Agent())
    port = server.server_address[1]
    try:
        # the exact bytes a reference client sends: pickle of {"method", "args", "kwargs"}
        assert _raw_post(port, pickle.dumps({"method": "__init__"})) == (200, {"result": "agent created"})
        code, reply = _raw_post(port, pickle.dumps({"method": "__call__", "args": [{"x": np.arange(4.0)}, "open the drawer"]}))
        assert code == 200 and reply == {"result": {"sum": 6.0, "goal": "open the drawer"}}
        code, reply = _raw_post(port, pickle.dumps({"method": "step", "args": [np.ones(3, np.float32)], "kwargs": {"lang": "go"}}))
        assert code == 200 and np.array_equal(reply["result"][0], 2 * np.ones(3, np.float32)) and reply["result"][1] == "go"
        client = AP.AgentClient(port=port)
        assert client.reset() is None
        t = torch.arange(6.0).reshape(2, 3)
        assert torch.equal(torch.as_tensor(client.step(t, "x")[0]), 2 * t)  # tensors survive the allow-listed unpickler
        # an agent exception: status 500, the reference's error payload, agent destroyed (agent_proxy.py:127-132)
        assert _raw_post(port, pickle.dumps({"method": "boom"})) == (500, {"error": "there was a problem"})
        assert _raw_post(port, pickle.dumps({"method": "reset"}))[0] == 500  # no agent any more
        client.init()
        assert len(created) == 2 and client.reset() is None
        # a request that names code to execute is refused by the allow-list (the reference would run it)
        evil = b"cos\nsystem\n(S'true'\ntR."
        assert _raw_post(port, evil) == (500, {"error": "there was a problem"})
        client.init()
        assert client.shutdown() == "shutdown"
        thread.join(timeout=10)
        assert not thread.is_alive()
    finally:
        server.server_close()


def test_batching_front_end_coalesces_concurrent_environments():
    calls = []

    def batch_fn(reqs):
        calls.append(len(reqs))
        return [r * 10 for r in reqs]

    pol = AP.BatchingPolicy(batch_fn, max_batch=8, window_s=0.05)
    out = {}
    barrier = threading.Barrier(12)

    def env(i):
        barrier.wait()
        out[i] = pol.step(i)

    threads = [threading.Thread(target=env, args=(i,)) for i in range(12)]
    [t.start() for t in threads]
    [t.join(timeout=10) for t in threads]
    assert out == {i: i * 10 for i in range(12)}
    assert sum(calls) == 12 and max(calls) <= 8 and len(calls) < 12  # batched, never beyond max_batch
    # a failing batch reaches every waiting client; the worker survives
    pol.batch_fn = lambda reqs: (_ for _ in ()).throw(RuntimeError("engine error"))
    with pytest.raises(RuntimeError, match="engine error"):
        pol(1)
    pol.batch_fn = batch_fn
    assert pol(3) == 30
    pol.close()
    with pytest.raises(RuntimeError):
        pol(1)


def test_batching_policy_behind_the_http_proxy():
    """n simulator clients -> HTTP threads -> one batched policy call (what a B200 serving many environments does)."""
    sizes = []

    def batch_fn(reqs):
        sizes.append(len(reqs))
        return [np.full((10, 7), float(np.sum(r["state_images"])), np.float32) for r in reqs]

    server, thread = _serve(lambda: AP.BatchingPolicy(batch_fn, max_batch=16, window_s=0.05))
    port = server.server_address[1]
    try:
        AP.AgentClient(port=port).init()
        res, go = {}, threading.Barrier(6)

        def env(i):
            go.wait()
            res[i] = AP.AgentClient(port=port).step({"state_images": np.full((2, 4), i, np.float32), "latent_goal": np.zeros(3)})

        ts = [threading.Thread(target=env, args=(i,)) for i in range(6)]
        [t.start() for t in ts]
        [t.join(timeout=20) for t in ts]
        assert all(res[i].shape == (10, 7) and res[i][0, 0] == 8.0 * i for i in range(6))
        assert sum(sizes) == 6 and len(sizes) < 6
        AP.AgentClient(port=port).shutdown()
        thread.join(timeout=10)
    finally:
        server.server_close()


def test_hf_export_matches_the_reference_layout_and_round_trips(tmp_path):
    from mode_diffusion_policy_b200.modedit import MoDeDiT

    m = MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256, embed_pdrob=0,
                attn_pdrop=0.3, n_layers=2, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10, state_dim=7)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g))
    agent_sd = {"model.inner_model." + k: v for k, v in m.state_dict().items()}
    agent_sd["static_resnet.film1.gamma.weight"] = torch.ones(3, 2)  # encoder tensors ride along untouched
    cfg = {"latent_dim": 512, "obs_enc_dim": 2048, "cond_dim": 512, "resnet_type": "50", "multistep": 10,
           "sampler_type": "ddim", "num_sampling_steps": 10, "sigma_data": 0.5, "sigma_min": 0.001, "sigma_max": 80,
           "noise_scheduler": "exponential", "sigma_sample_density_type": "loglogistic", "act_window_size": 10,
           "use_proprio": False, "model": {"inner_model": {"_target_": "mode.models.networks.modedit.MoDeDiT", "embed_dim": 256}},
           "optimizer": {"lr": 1e-4}}
    paths = HF.export(agent_sd, cfg, str(tmp_path / "hf"))
    assert sorted(paths) == ["README.md", "config.json", "model.pt", "model.safetensors"]
    # keys exactly as the reference's `k.replace('model.', '')` writes them (save_to_hf.py:121)
    raw = torch.load(paths["model.pt"], weights_only=True)["state_dict"]
    assert set(raw) == {k.replace("model.", "") for k in agent_sd}
    assert "inner_blocks.0.ln_1.g" in raw and "static_resnet.film1.gamma.weight" in raw
    conf = json.load(open(paths["config.json"]))
    assert list(conf) == ["model_config"] and "optimizer" not in conf["model_config"]
    assert conf["model_config"]["model"]["inner_model"]["embed_dim"] == 256 and conf["model_config"]["sampler_type"] == "ddim"
    # round trip: restored agent-level keys, and the checkpoint loader takes the exported file directly
    sd, mc = HF.load_export(str(tmp_path / "hf"))
    assert set(sd) == set(agent_sd) and all(torch.equal(sd[k], agent_sd[k]) for k in agent_sd) and mc["multistep"] == 10
    m2 = MoDeDiT(obs_dim=128, goal_dim=64, device="cpu", goal_conditioned=True, action_dim=7, embed_dim=256, embed_pdrob=0,
                 attn_pdrop=0.3, n_layers=2, n_heads=4, goal_seq_len=1, obs_seq_len=1, action_seq_len=10, state_dim=7)
    rep = CK.load_pretrained_parameters(m2, paths["model.safetensors"], freeze_routers=True)
    assert not rep.missing and not rep.skipped_shape and rep.ignored == 1
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert not any(p.requires_grad for p in m2.blocks[0].router.parameters())
    # exporting a module directly uses the agent-level names as well
    HF.export(m, cfg, str(tmp_path / "hf2"), safetensors=False)
    assert set(torch.load(tmp_path / "hf2" / "model.pt", weights_only=True)["state_dict"]) == {k.replace("model.", "") for k in agent_sd if "resnet" not in k}
