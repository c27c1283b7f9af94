"""Thin object wrapper over the C ABI (include/mode_engine.h). PyTorch is used only for device memory and streams."""
from __future__ import annotations

import ctypes as C
from typing import Optional
from dataclasses import dataclass, asdict

import numpy as np
import torch

from . import _lib


@dataclass
class EngineConfig:
    """Mirror of mode_config_t; field meaning follows MoDeDiT.__init__ (reference modedit.py:643-674)."""

    obs_dim: int = 2048
    goal_dim: int = 512
    action_dim: int = 7
    embed_dim: int = 1024
    n_layers: int = 12
    n_heads: int = 8
    n_state_tokens: int = 2
    action_seq_len: int = 10
    num_experts: int = 4
    top_k: int = 2
    router_normalize: bool = True
    max_batch: int = 256
    sigma_data: float = 0.5
    rms_eps: float = 1e-6

    @property
    def seq_len(self) -> int:
        return 2 + self.n_state_tokens + self.action_seq_len


def _f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.ModeError(f"{what} must be a CUDA tensor: the MoDE engine has no CPU path")
    return t.detach().to(torch.float32).contiguous()


class ModeEngine:
    """One engine = packed bf16 weights + workspace + CUDA graphs on the current CUDA device."""

    def __init__(self, cfg: EngineConfig):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.ModeError("no CUDA device: the MoDE engine only runs on B200 (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        c = _lib.mode_config_t(**{k: (int(v) if not isinstance(v, float) else v) for k, v in asdict(cfg).items()})
        h = C.c_void_p()
        _lib.check(self.lib.mode_create(C.byref(c), C.byref(h)))
        self._h = h
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.train_generation = 0      # number of train_step calls: each one overwrites the flat gradient buffer
        self.has_train_state = False   # optimizer moments / EMA / exported zero-copy views live in this engine

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mode_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- weights
    def load_state_dict(self, sd) -> None:
        """sd: mapping reference-name -> numpy array or torch tensor (fp32 master weights, reference shapes)."""
        for name, w in sd.items():
            if isinstance(w, torch.Tensor):
                t = w.detach().to(torch.float32).contiguous()
                is_dev = 1 if t.is_cuda else 0
                ptr, shape = t.data_ptr(), tuple(t.shape)
                keep = t
            else:
                keep = np.ascontiguousarray(w, dtype=np.float32)
                is_dev, ptr, shape = 0, keep.ctypes.data, keep.shape
            shp = (C.c_int64 * len(shape))(*shape)
            _lib.check(self.lib.mode_set_weight_on_stream(self._h, name.encode(), ptr, is_dev, shp, len(shape), self._stream()))
            del keep
        _lib.check(self.lib.mode_finalize_weights_on_stream(self._h, self._stream()))

    # ---------------------------------------------------------------- evaluations
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _prep(self, state, goal, actions, sigma):
        state = _f32_cuda(state, "state_images")
        goal = _f32_cuda(goal, "goal")
        actions = _f32_cuda(actions, "actions")
        B = actions.shape[0]
        c = self.cfg
        if tuple(state.shape) != (B, c.n_state_tokens, c.obs_dim):
            raise _lib.ModeError(f"state_images must be {(B, c.n_state_tokens, c.obs_dim)}, got {tuple(state.shape)}")
        if goal.numel() != B * c.goal_dim:
            raise _lib.ModeError(f"goal must hold {B}x{c.goal_dim} values, got shape {tuple(goal.shape)}")
        if tuple(actions.shape) != (B, c.action_seq_len, c.action_dim):
            raise _lib.ModeError(f"actions must be {(B, c.action_seq_len, c.action_dim)}, got {tuple(actions.shape)}")
        stride = 1
        if sigma is not None:
            sigma = _f32_cuda(sigma, "sigma").reshape(-1)
            if sigma.numel() == 1:
                stride = 0
            elif sigma.numel() != B:
                raise _lib.ModeError(f"sigma must have 1 or {B} elements, got {sigma.numel()}")
        return state, goal, actions, sigma, stride, B

    def forward(self, state, actions, goal, sigma) -> torch.Tensor:
        """MoDeDiT.forward (raw network output F)."""
        state, goal, actions, sigma, stride, B = self._prep(state, goal, actions, sigma)
        out = torch.empty_like(actions)
        _lib.check(self.lib.mode_forward(self._h, state.data_ptr(), goal.data_ptr(), actions.data_ptr(),
                                         sigma.data_ptr(), stride, out.data_ptr(), B, self._stream()))
        return out

    def denoise(self, state, actions, goal, sigma) -> torch.Tensor:
        """GCDenoiser.forward: D(x; sigma)."""
        state, goal, actions, sigma, stride, B = self._prep(state, goal, actions, sigma)
        out = torch.empty_like(actions)
        _lib.check(self.lib.mode_denoise(self._h, state.data_ptr(), goal.data_ptr(), actions.data_ptr(),
                                         sigma.data_ptr(), stride, out.data_ptr(), B, self._stream()))
        return out

    PROF_CLASSES = ["router+plan", "embed", "qkv_gemm", "attention", "proj_gemm", "ln2_permute", "up_gemm_swiglu",
                    "down_gemm", "combine_ln1", "head", "cond_embed"]

    def profile_eval(self, state, actions, goal, sigma, reps: int = 3):
        """Per-kernel-class device time (ms per evaluation) and launch counts of one denoiser call."""
        state, goal, actions, sigma, stride, B = self._prep(state, goal, actions, sigma)
        out = torch.empty_like(actions)
        ms = np.zeros(len(self.PROF_CLASSES), dtype=np.float32)
        n = np.zeros(len(self.PROF_CLASSES), dtype=np.int32)
        _lib.check(self.lib.mode_profile_eval(self._h, state.data_ptr(), goal.data_ptr(), actions.data_ptr(),
                                              sigma.data_ptr(), stride, out.data_ptr(), B, reps, self._stream(),
                                              ms.ctypes.data, n.ctypes.data))
        return {k: (float(m), int(c)) for k, m, c in zip(self.PROF_CLASSES, ms, n)}

    def loss(self, state, action, goal, noise, sigma):
        """GCDenoiser.loss forward (eval-mode network). Returns (loss scalar tensor, model output)."""
        state, goal, action, sigma, stride, B = self._prep(state, goal, action, sigma)
        noise = _f32_cuda(noise, "noise")
        if stride != 1 and B != 1:
            sigma = sigma.expand(B).contiguous()
        out = torch.empty_like(action)
        loss = torch.empty(1, dtype=torch.float32, device=action.device)
        _lib.check(self.lib.mode_loss(self._h, state.data_ptr(), goal.data_ptr(), action.data_ptr(), noise.data_ptr(),
                                      sigma.data_ptr(), loss.data_ptr(), out.data_ptr(), B, self._stream()))
        return loss[0], out

    # ---------------------------------------------------------------- training
    def train_step(self, state, action, goal, noise, sigma):
        """Forward + hand-written backward of GCDenoiser.loss (deterministic mode). Returns (loss, model output F); the
        gradients are left in the engine's flat buffer (see `grad`, `flat_grads`)."""
        state, goal, action, sigma, stride, B = self._prep(state, goal, action, sigma)
        noise = _f32_cuda(noise, "noise")
        if stride != 1 and B != 1:
            sigma = sigma.expand(B).contiguous()
        out = torch.empty_like(action)
        loss = torch.empty(1, dtype=torch.float32, device=action.device)
        _lib.check(self.lib.mode_train_step(self._h, state.data_ptr(), goal.data_ptr(), action.data_ptr(), noise.data_ptr(),
                                            sigma.data_ptr(), loss.data_ptr(), out.data_ptr(), B, self._stream()))
        self.train_generation += 1
        return loss[0], out

    def set_stochastic(self, attn_pdrop=0.0, mlp_pdrop=0.0, goal_drop=0.0, multinomial=False, seed=0, step=0, embed_pdrop=0.0):
        """Stochastic regularisation of the reference's train mode for the following `train_step` calls (attention /
        expert-MLP dropout, goal masking, per-token multinomial routing; reference modedit.py:149, :254, :882-893,
        :389-390). The masks are a pure function of (seed, step, ...): `step` advances by one per stochastic step."""
        _lib.check(self.lib.mode_train_set_stochastic(self._h, float(attn_pdrop), float(mlp_pdrop), float(goal_drop),
                                                      float(embed_pdrop), int(bool(multinomial)),
                                                      int(seed) & (2 ** 64 - 1), int(step)))

    def token_routing(self, layer: int, batch: int):
        """(expert indices in draw order, renormalised probabilities), each [batch*T, top_k], of the last multinomial
        training step."""
        n = batch * self.cfg.seq_len
        idx = np.zeros((n, self.cfg.top_k), dtype=np.int32)
        w = np.zeros((n, self.cfg.top_k), dtype=np.float32)
        _lib.check(self.lib.mode_train_get_token_routing(self._h, layer, batch, idx.ctypes.data, w.ctypes.data))
        return idx, w

    def flat_grads(self) -> torch.Tensor:
        """Zero-copy torch view of the engine-owned flat fp32 gradient buffer (one all-reduce synchronises DP ranks)."""
        if getattr(self, "_flat", None) is None:
            self.has_train_state = True
            ptr, n = C.c_void_p(), C.c_int64()
            _lib.check(self.lib.mode_grad_buffer(self._h, C.byref(ptr), C.byref(n)))

            class _Dev:  # __cuda_array_interface__ holder
                pass

            holder = _Dev()
            holder.__cuda_array_interface__ = {"shape": (n.value,), "typestr": "<f4", "data": (ptr.value, False), "version": 2}
            self._flat = torch.as_tensor(holder, device=self.device)
        return self._flat

    def grad(self, name: str, shape) -> torch.Tensor:
        """View of the gradient of reference parameter `name` (reference layout)."""
        off, n = C.c_int64(), C.c_int64()
        _lib.check(self.lib.mode_grad_offset(self._h, name.encode(), C.byref(off), C.byref(n)))
        return self.flat_grads()[off.value: off.value + n.value].view(*shape)

    def grad_range(self, name: str) -> tuple[int, int]:
        """(element offset, numel) of `name`'s gradient inside the flat buffer."""
        off, n = C.c_int64(), C.c_int64()
        _lib.check(self.lib.mode_grad_offset(self._h, name.encode(), C.byref(off), C.byref(n)))
        return off.value, n.value

    def input_grads(self, B: int, state_shape, goal_shape, want_state=True, want_goal=True):
        """d loss / d state_images and d loss / d goal of the last train_step (fp32 tensors, or None if not wanted)."""
        ds = torch.empty(state_shape, dtype=torch.float32, device=self.device) if want_state else None
        dg = torch.empty(goal_shape, dtype=torch.float32, device=self.device) if want_goal else None
        _lib.check(self.lib.mode_train_input_grads(self._h, ds.data_ptr() if want_state else None,
                                                   dg.data_ptr() if want_goal else None, B, self._stream()))
        return ds, dg

    # ---------------------------------------------------------------- fused optimizer
    def optimizer_bind(self, name: str, param: torch.Tensor, weight_decay: bool) -> None:
        """Register the fp32 master of reference parameter `name`; mode_adamw_step updates it in place."""
        if not (param.is_cuda and param.dtype == torch.float32 and param.is_contiguous()):
            raise ValueError(f"{name}: the fused optimizer needs a contiguous fp32 CUDA parameter")
        self.has_train_state = True
        _lib.check(self.lib.mode_optimizer_bind(self._h, name.encode(), C.c_void_p(param.data_ptr()), int(bool(weight_decay))))

    def optimizer_unbind_all(self) -> None:
        _lib.check(self.lib.mode_optimizer_unbind_all(self._h))

    def adamw_step(self, lr, beta1, beta2, eps, weight_decay, step: int, grad_scale: Optional[torch.Tensor] = None,
                   group: Optional[int] = None, stream: Optional["torch.cuda.Stream"] = None) -> None:
        """AdamW over every bound parameter + re-pack of the engine's weight copies, one launch (csrc/optimizer.cuh).
        `group` = l launches only block l's large tensors, `group` = n_layers the remaining ones (last): the pieces a
        data-parallel step pipelines with its gradient exchange (`optim.EngineAdamW.step_overlapped`)."""
        gs = None
        if grad_scale is not None:
            gs = grad_scale.detach().to(device=self.device, dtype=torch.float32).reshape(1).contiguous()
            self._gs_keepalive = gs  # the launch may run on a side stream after this call returns
        gs_ptr = C.c_void_p(gs.data_ptr()) if gs is not None else None
        st = C.c_void_p(stream.cuda_stream) if stream is not None else self._stream()
        if group is None:
            _lib.check(self.lib.mode_adamw_step(self._h, float(lr), float(beta1), float(beta2), float(eps),
                                                float(weight_decay), int(step), gs_ptr, st))
        else:
            _lib.check(self.lib.mode_adamw_step_group(self._h, float(lr), float(beta1), float(beta2), float(eps),
                                                      float(weight_decay), int(step), gs_ptr, int(group), st))

    def optimizer_state(self):
        """Zero-copy views (exp_avg, exp_avg_sq) of the engine-owned moment buffers (gradient-buffer layout)."""
        m, v, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self.lib.mode_optimizer_state(self._h, C.byref(m), C.byref(v), C.byref(n)))

        def view(ptr):
            class _Dev:
                pass

            h = _Dev()
            h.__cuda_array_interface__ = {"shape": (n.value,), "typestr": "<f4", "data": (ptr.value, False), "version": 2}
            return torch.as_tensor(h, device=self.device)

        return view(m), view(v)

    def _flat_view(self, ptr: int, numel: int) -> torch.Tensor:
        class _Dev:
            pass

        h = _Dev()
        h.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(h, device=self.device)

    def set_optimizer_sharding(self, rank: int, world: int) -> None:
        """Each data-parallel rank updates 1/world of every large per-block tensor (`mode_optimizer_set_sharding`);
        world <= 1 switches it off."""
        self.has_train_state = True
        _lib.check(self.lib.mode_optimizer_set_sharding(self._h, int(rank), max(int(world), 1)))

    def optimizer_shard_tensors(self, group: int):
        """(offset, numel) spans of the flat buffers that block `group`'s optimizer launch covers (the tensors a sharded
        step reduce-scatters / all-gathers)."""
        n = C.c_int()
        _lib.check(self.lib.mode_optimizer_shard_tensors(self._h, int(group), None, None, 0, C.byref(n)))
        offs, nums = (C.c_int64 * max(n.value, 1))(), (C.c_int64 * max(n.value, 1))()
        _lib.check(self.lib.mode_optimizer_shard_tensors(self._h, int(group), offs, nums, n.value, C.byref(n)))
        return [(int(offs[i]), int(nums[i])) for i in range(n.value)]

    def optimizer_staging(self) -> torch.Tensor:
        """Zero-copy bf16 view of the sharded optimizer's all-gather buffer (gradient-buffer layout)."""
        self.has_train_state = True
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.mode_optimizer_staging(self._h, C.byref(p), C.byref(n)))

        class _Dev:
            pass

        h = _Dev()
        h.__cuda_array_interface__ = {"shape": (n.value,), "typestr": "<i2", "data": (p.value, False), "version": 2}
        return torch.as_tensor(h, device=self.device).view(torch.bfloat16)

    def optimizer_pack_group(self, group: int, stream: Optional["torch.cuda.Stream"] = None) -> None:
        """Packed bf16 copies of block `group`'s tensors from the (all-gathered) staging buffer."""
        st = C.c_void_p(stream.cuda_stream) if stream is not None else self._stream()
        _lib.check(self.lib.mode_optimizer_pack_group(self._h, int(group), st))

    def weights_record_ready(self, layer: int, stream: "torch.cuda.Stream") -> None:
        """Block `layer`'s packed weights are current once `stream`'s work so far is done; the next engine call that reads
        them waits for that on its own stream (`mode_weights_record_ready`)."""
        _lib.check(self.lib.mode_weights_record_ready(self._h, int(layer), C.c_void_p(stream.cuda_stream)))

    def set_ema(self, decay: Optional[float]) -> None:
        """Moving average of the bound parameters inside the optimizer launch (reference mode/callbacks/ema.py);
        None switches it off."""
        _lib.check(self.lib.mode_optimizer_set_ema(self._h, -1.0 if decay is None else float(decay)))

    def ema_state(self) -> torch.Tensor:
        """Zero-copy view of the engine-owned EMA buffer (gradient-buffer layout; `grad_range(name)` gives spans)."""
        self.has_train_state = True
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.mode_optimizer_ema_state(self._h, C.byref(p), C.byref(n)))
        return self._flat_view(p.value, n.value)

    def ema_mark_seeded(self) -> None:
        """After restoring a checkpointed average into `ema_state()`: keep it instead of re-seeding from the weights."""
        _lib.check(self.lib.mode_optimizer_ema_mark_seeded(self._h))

    def grad_sumsq(self, spans) -> torch.Tensor:
        """Sum of squares of each (offset, numel) span of the flat gradient buffer, as one device tensor (two launches,
        no host synchronisation beyond the span-table upload)."""
        seg = np.ascontiguousarray(np.asarray(spans, dtype=np.int64).reshape(-1, 2))
        out = torch.empty(seg.shape[0], dtype=torch.float32, device=self.device)
        _lib.check(self.lib.mode_grad_segment_sumsq(self._h, seg.ctypes.data, seg.shape[0], out.data_ptr(), self._stream()))
        return out

    def wait_grads(self, layer: int, stream: "torch.cuda.Stream") -> None:
        """Make `stream` wait until the last train_step finished block `layer`'s gradients (-1: all gradients)."""
        _lib.check(self.lib.mode_train_wait_grads(self._h, layer, C.c_void_p(stream.cuda_stream)))

    FUSED_SAMPLERS = {"ddim": 0, "euler": 1, "dpmpp_2m": 2}  # MODE_SAMPLER_* of include/mode_engine.h

    def sample(self, sampler: str, state, x, goal, sigmas) -> torch.Tensor:
        """A whole k-diffusion sampler loop over GCDenoiser as one CUDA-graph launch: "ddim" (reference sample_ddim),
        "euler" (sample_euler with s_churn=0), "dpmpp_2m" (sample_dpmpp_2m). `sigmas` includes the trailing 0."""
        state, goal, x, _, _, B = self._prep(state, goal, x, None)
        x = x.clone()
        sig = np.ascontiguousarray(torch.as_tensor(sigmas).detach().float().cpu().numpy(), dtype=np.float32)
        _lib.check(self.lib.mode_sample(self._h, self.FUSED_SAMPLERS[sampler], state.data_ptr(), goal.data_ptr(), x.data_ptr(),
                                        sig.ctypes.data_as(C.POINTER(C.c_float)), len(sig), B, self._stream()))
        return x

    def sample_program(self, state, x, goal, sigma_eval, reads_probe, prog, noise=None) -> torch.Tensor:
        """A sampler program (include/mode_engine.h `mode_sample_program`): one CUDA-graph launch for the whole loop.
        sigma_eval (n,), reads_probe (n,) 0/1, prog (n, 16) fp32 coefficient rows, noise (n, B, A, adim) CUDA or None."""
        state, goal, x, _, _, B = self._prep(state, goal, x, None)
        x = x.clone()
        sig = np.ascontiguousarray(sigma_eval, dtype=np.float32)
        rp = np.ascontiguousarray(reads_probe, dtype=np.int32)
        pg = np.ascontiguousarray(prog, dtype=np.float32)
        n = len(sig)
        if pg.shape != (n, 16) or rp.shape != (n,):
            raise _lib.ModeError(f"sampler program: expected ({n}, 16) coefficients and ({n},) flags")
        if noise is not None:
            noise = _f32_cuda(noise, "noise")
            if tuple(noise.shape) != (n,) + tuple(x.shape):
                raise _lib.ModeError(f"sampler program: noise must be {(n,) + tuple(x.shape)}")
        _lib.check(self.lib.mode_sample_program(
            self._h, state.data_ptr(), goal.data_ptr(), x.data_ptr(), sig.ctypes.data_as(C.POINTER(C.c_float)),
            rp.ctypes.data_as(C.POINTER(C.c_int32)), pg.ctypes.data_as(C.POINTER(C.c_float)),
            noise.data_ptr() if noise is not None else None, n, B, self._stream()))
        return x

    def sample_ddim(self, state, x, goal, sigmas) -> torch.Tensor:
        """sample_ddim over GCDenoiser (whole loop = one CUDA graph). `sigmas` includes the trailing 0. Returns actions."""
        return self.sample("ddim", state, x, goal, sigmas)

    def sample_ddim_host(self, state: np.ndarray, x: np.ndarray, goal: np.ndarray, sigmas: np.ndarray) -> np.ndarray:
        """Host-buffer entry (numpy or pinned torch CPU tensors): H2D, sample, D2H, synchronised. Returns actions."""
        def host_ptr(a):
            if isinstance(a, torch.Tensor):
                assert not a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()
                return a.data_ptr()
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
            return a.ctypes.data

        B = x.shape[0]
        sig = np.ascontiguousarray(sigmas, dtype=np.float32)
        _lib.check(self.lib.mode_sample_ddim_host(self._h, host_ptr(state), host_ptr(goal), host_ptr(x),
                                                  sig.ctypes.data_as(C.POINTER(C.c_float)), len(sig), B, self._stream()))
        return x

    def block_forward(self, layer: int, x, c) -> torch.Tensor:
        """NoiseBlockMoE.forward(x, c) for one layer; x (B, T, d), c (B, 1, d) or (B, d)."""
        x = _f32_cuda(x, "x")
        c = _f32_cuda(c, "c").reshape(x.shape[0], -1)
        if x.shape[1] != self.cfg.seq_len or x.shape[2] != self.cfg.embed_dim:
            raise _lib.ModeError(f"x must be (B, {self.cfg.seq_len}, {self.cfg.embed_dim}), got {tuple(x.shape)}")
        out = torch.empty_like(x)
        _lib.check(self.lib.mode_block_forward(self._h, layer, x.data_ptr(), c.data_ptr(), out.data_ptr(), x.shape[0],
                                               self._stream()))
        return out

    # ---------------------------------------------------------------- introspection
    def routing(self, layer: int, B: int, step: int = -1):
        """(top-k indices in torch.topk order, renormalised weights, clamped probabilities) of layer `layer` for the most
        recent evaluation, or for evaluation `step` of the most recent fused sampler call."""
        k, E = self.cfg.top_k, self.cfg.num_experts
        idx = np.empty((B, k), dtype=np.int32)
        w = np.empty((B, k), dtype=np.float32)
        probs = np.empty((B, E), dtype=np.float32)
        _lib.check(self.lib.mode_get_routing_at(self._h, step, layer, B, idx.ctypes.data, w.ctypes.data, probs.ctypes.data))
        return idx, w, probs

    def expert_usage(self, layer: int):
        usage = np.zeros(self.cfg.num_experts, dtype=np.int64)
        total = np.zeros(1, dtype=np.int64)
        _lib.check(self.lib.mode_get_expert_usage(self._h, layer, usage.ctypes.data, total.ctypes.data))
        return usage, int(total[0])

    def reset_expert_usage(self):
        _lib.check(self.lib.mode_reset_expert_usage(self._h))

    def last_launch_count(self) -> int:
        return int(self.lib.mode_last_launch_count(self._h))
