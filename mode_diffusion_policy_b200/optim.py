"""Fused optimizer for the engine's training path.

`EngineAdamW` is a `torch.optim.Optimizer` (LR schedulers, Lightning's `configure_optimizers` and `state_dict()` see
an ordinary optimizer with one param group), but `step()` is a single engine launch: AdamW with torch's arithmetic over
the flat gradient buffer written by `GCDenoiser.loss`, updating the fp32 master parameters in place AND the engine's
packed bf16 copies in the same pass (csrc/optimizer.cuh). It replaces three passes over 686 M parameters in the
reference setup — autograd handing out gradient copies, `torch.optim.AdamW` (mode_agent.py:267-301) and this engine's
weight re-pack.

Weight-decay groups follow MoDEAgent.get_optim_groups (mode_agent.py:362-384): parameters whose name contains 'bias',
'LayerNorm' or 'embedding' are not decayed.
"""
from __future__ import annotations

import torch

from .modedit import MoDeDiT

NO_DECAY_SUBSTRINGS = ("bias", "LayerNorm", "embedding")  # reference mode_agent.py:365-366


def use_weight_decay(name: str) -> bool:
    return all(x not in name for x in NO_DECAY_SUBSTRINGS)


class EngineAdamW(torch.optim.Optimizer):
    def __init__(self, inner_model: MoDeDiT, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 materialize_grads: bool = False, ema_decay=None):
        """materialize_grads=False: `loss.backward()` no longer copies the engine's gradients into `.grad` (nothing reads
        them); set True to keep `.grad` populated for gradient logging (mode_agent.py:304-359; or use
        `MoDeDiT.grad_norms()`, which needs no `.grad`).

        ema_decay: None (off), a float, or a callable step -> decay (the reference callback's warm-up,
        mode/callbacks/ema.py:84-92): an exponential moving average of the updated weights is kept by the same launch
        (`ema_state_dict()`, `swap_ema_weights()`)."""
        self.inner = inner_model
        named = [(n, p) for n, p in inner_model.named_parameters() if p.requires_grad and n != "gripper_embed.weight"]
        super().__init__([p for _, p in named], dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._names = [n for n, _ in named]
        self._ema_decay = ema_decay
        self._step = 0
        self._bound = {}
        self._consumed_generation = None  # engine.train_generation of the gradients the last step() used
        self._masters_stale = False   # sharded steps with master_sync="lazy": fp32 masters current only on their owners
        self._masters_ready = None    # event behind the last all-gather of the masters
        inner_model.register_state_dict_pre_hook(self._state_dict_guard)
        inner_model._skip_param_grads = not materialize_grads
        inner_model._engine_keeps_sync = True  # step() re-packs what it updates: GCDenoiser.loss need not
        inner_model._loss_grad_scale = None

    def _state_dict_guard(self, module, prefix, keep_vars):
        """`state_dict()` of the model (checkpoints, the EMA callback, and the engine's own re-pack from the masters) must
        not read masters that only their owning rank has updated."""
        if self._masters_stale:
            raise RuntimeError(
                "EngineAdamW.step_sharded(master_sync='lazy'): the fp32 master parameters of the sharded tensors are only "
                "current on their owning rank. Call optimizer.synchronize_parameters() on EVERY rank before reading "
                "state_dict() / modifying parameters (it all-gathers them), or use master_sync='step'.")
        if self._masters_ready is not None and torch.cuda.is_available():
            torch.cuda.current_stream().wait_event(self._masters_ready)

    def _set_ema(self, eng):
        d = self._ema_decay
        eng.set_ema(None if d is None else float(d(self._step) if callable(d) else d))

    def ema_state_dict(self):
        """name -> EMA tensor (reference layout, views of the engine's buffer) for every parameter this optimizer updates."""
        eng = self.inner._engine
        self._gather_sharded_state(eng)
        flat = eng.ema_state()
        params = dict(self.inner.named_parameters())
        out = {}
        for name in self._names:
            off, n = eng.grad_range(name)
            out[name] = flat[off: off + n].view(params[name].shape)
        return out

    def swap_ema_weights(self):
        """Context manager: evaluate with the averaged weights (EMA.replace_model_weights / restore_original_weights,
        ema.py:184-195), then put the training weights back."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            self.synchronize_parameters()  # after sharded steps: a collective, enter on every rank
            params = dict(self.inner.named_parameters())
            ema = self.ema_state_dict()
            saved = {n: params[n].detach().clone() for n in ema}
            with torch.no_grad():
                for n, v in ema.items():
                    params[n].copy_(v)  # bumps the version counters: the engine re-packs on the next call
            try:
                yield
            finally:
                with torch.no_grad():
                    for n, v in saved.items():
                        params[n].copy_(v)
        return ctx()

    def _check_fresh_gradients(self, eng):
        """step() reads the engine's flat buffer, which every training-mode loss() overwrites: exactly one loss() per
        step() (no gradient accumulation across calls) unless `.grad` tensors are materialised and accumulated by autograd."""
        if not self.inner._skip_param_grads:
            return
        gen = eng.train_generation
        if self._consumed_generation is not None and gen > self._consumed_generation + 1:
            raise RuntimeError(
                f"EngineAdamW(materialize_grads=False): {gen - self._consumed_generation} training-mode loss() calls since "
                "the last step(), but only the last one's gradients are in the engine's buffer. Gradient accumulation "
                "(e.g. Lightning accumulate_grad_batches > 1) needs a larger per-call batch instead.")
        self._consumed_generation = gen

    def _bind(self, eng):
        if self._bound.get("engine") is not eng:
            eng.optimizer_unbind_all()
            self._bound = {"engine": eng}
        params = dict(self.inner.named_parameters())
        for name in self._names:
            p = params[name]
            key = (p.data_ptr(), p.requires_grad)
            if self._bound.get(name) != key and p.requires_grad:
                eng.optimizer_bind(name, p.data, use_weight_decay(name))
                self._bound[name] = key

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng = getattr(self.inner, "_engine", None)
        if eng is None:
            raise RuntimeError("EngineAdamW.step before any GCDenoiser.loss call")
        self._bind(eng)
        self._check_fresh_gradients(eng)
        g = self.param_groups[0]
        self._step += 1
        self._set_ema(eng)
        eng.adamw_step(g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._step,
                       self.inner._loss_grad_scale)
        return loss

    @torch.no_grad()
    def step_overlapped(self, reducer, timeline=None, loss_scale=None) -> None:
        """Data-parallel step with the gradient exchange and the optimizer pipelined layer by layer.

        `reducer` is a `parallel.GradAllReduce` over the same engine. Block l's large gradient buckets are all-reduced on
        the reducer's stream as soon as its backward is done (last block first); the optimizer launch for that block
        (`mode_adamw_step_group`) follows on a second stream behind an event, so it runs while the next blocks' buckets
        are still on the wire: the HBM-bound update hides behind the NVLink-bound exchange instead of waiting for all of
        it. The small tensors and non-block parameters go last (tail buckets, group n_layers). Call after
        `loss.backward()`; the caller's stream continues only when every group has finished. `timeline`: optional list
        that receives (label, cuda event) pairs recorded on the streams involved (measurement aid).

        `loss_scale`: None = use the gradient autograd fed into the loss (`loss.backward()` must have run; the optimizer
        stream then waits for everything the caller's stream has enqueued, i.e. the whole backward). A float = the loss
        was (or would have been) scaled by this constant: no `loss.backward()` is needed and block l's update may start
        as soon as its gradients are exchanged, while the backward of blocks l-1..0 is still running — the small
        HBM-bound optimizer CTAs co-reside with the persistent tensor-core GEMM CTAs."""
        eng = getattr(self.inner, "_engine", None)
        if eng is None:
            raise RuntimeError("EngineAdamW.step_overlapped before any GCDenoiser.loss call")
        if getattr(self, "_exchange", None) is not None and reducer is not self._exchange:
            raise RuntimeError("this optimizer's state is sharded over the ranks (step_sharded was used): keep using step_sharded")
        self._bind(eng)
        self._check_fresh_gradients(eng)
        g = self.param_groups[0]
        self._step += 1
        self._set_ema(eng)
        dev = eng.device
        if loss_scale is None:
            scale = self.inner._loss_grad_scale
        elif float(loss_scale) == 1.0:
            scale = None
        else:
            if getattr(self, "_const_scale", None) is None or float(self._const_scale[0]) != float(loss_scale):
                self._const_scale = (float(loss_scale), torch.full((1,), float(loss_scale), dtype=torch.float32, device=dev))
                torch.cuda.current_stream(dev).synchronize()
            scale = self._const_scale[1]
        args = (g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._step, scale)
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_opt_stream", None) is None:
            self._opt_stream = torch.cuda.Stream(device=dev)
        opt_stream = self._opt_stream
        def mark(label, stream):
            if timeline is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                timeline.append((label, ev))

        mark("backward_done", main)
        if loss_scale is None:
            opt_stream.wait_stream(main)  # the autograd-provided scale is produced on the caller's stream, after the backward
        L = reducer.n_layers
        exchange = reducer.active()
        for layer in range(L - 1, -1, -1):
            if exchange:
                reducer.reduce_layer(layer)
                mark(f"exchange_{layer}", reducer.stream)
                opt_stream.wait_stream(reducer.stream)  # everything enqueued so far on the exchange stream, incl. this block
            else:
                eng.wait_grads(layer, opt_stream)
            eng.adamw_step(*args, group=layer, stream=opt_stream)
            mark(f"adamw_{layer}", opt_stream)
        if exchange:
            reducer.reduce_tail()
            mark("exchange_tail", reducer.stream)
            opt_stream.wait_stream(reducer.stream)
        else:
            eng.wait_grads(-1, opt_stream)
        eng.adamw_step(*args, group=L, stream=opt_stream)
        mark("adamw_rest", opt_stream)
        main.wait_stream(opt_stream)

    @torch.no_grad()
    def step_sharded(self, exchange, loss_scale=None, master_sync: str = "step", overlap_forward: bool = True,
                     timeline=None) -> None:
        """Data-parallel step with the optimizer state sharded over the ranks (`parallel.ShardedGradExchange`).

        Replaces DDP's all-reduce + replicated AdamW (reference mode/training_calvin.py:92-103, mode_agent.py:267-301):
        per block, last block first, (1) reduce-scatter(mean) of the block's large gradient tensors as soon as its
        backward is done, (2) this rank's 1/world of the fused AdamW (+ EMA) launch, which writes the new weights as bf16
        into the staging buffer, (3) all-gather of the staging spans, (4) re-pack of the block's bf16 GEMM operands from
        the gathered values; collectives on the exchange stream, updates and re-packs on a second one, so block l's
        HBM-bound update runs while block l+1's weights are on the wire. The remaining (small / non-block) tensors are all-reduced
        and updated on every rank as in `step_overlapped`. The same elements are averaged and updated by the same
        arithmetic as in the replicated step — each by exactly one rank — so every rank ends the step with identical
        packed weights; the wire carries 6 instead of 8 bytes per parameter and the HBM-bound optimizer pass shrinks by
        the number of ranks.

        `loss_scale`: as in `step_overlapped` (None = wait for autograd's incoming gradient of the loss, i.e. for the
        end of the backward, before the first update; a float = the caller guarantees that constant, updates start as
        soon as a block's gradients are reduced).

        `overlap_forward`: the caller's stream only waits for the small replicated tensors; each block's new weights are
        handed to the engine with a per-block event (`mode_weights_record_ready`), so the next `loss()` starts its forward
        while later blocks are still being updated / gathered, and any other engine call waits for all of them. False: the
        caller's stream waits for the whole step.

        `master_sync`: the fp32 master parameters, the moments and the EMA of a sharded tensor are only current on the
        rank that owns the element. "step" (default): the masters are all-gathered at the end of every step on the
        exchange stream, off the critical path (the next step's forward does not read them; `synchronize_parameters()`
        makes the caller's stream wait for it). "lazy": only when `synchronize_parameters()` is called (a collective:
        every rank must call it, e.g. before `state_dict()` / checkpointing / evaluation through other modules)."""
        eng = getattr(self.inner, "_engine", None)
        if eng is None:
            raise RuntimeError("EngineAdamW.step_sharded before any GCDenoiser.loss call")
        if not exchange.active():
            return self.step_overlapped(exchange, timeline=timeline, loss_scale=loss_scale)
        self._bind(eng)
        exchange.prepare()
        if exchange.fallback is not None:  # no tensor splits evenly over this world size: replicated step
            return self.step_overlapped(exchange.fallback, timeline=timeline, loss_scale=loss_scale)
        self._exchange = exchange
        self._check_fresh_gradients(eng)
        g = self.param_groups[0]
        self._step += 1
        self._set_ema(eng)
        dev = eng.device
        if loss_scale is None:
            scale = self.inner._loss_grad_scale
        elif float(loss_scale) == 1.0:
            scale = None
        else:
            if getattr(self, "_const_scale", None) is None or float(self._const_scale[0]) != float(loss_scale):
                self._const_scale = (float(loss_scale), torch.full((1,), float(loss_scale), dtype=torch.float32, device=dev))
                torch.cuda.current_stream(dev).synchronize()
            scale = self._const_scale[1]
        args = (g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._step, scale)
        main = torch.cuda.current_stream(dev)
        xs = exchange.stream

        def mark(label, stream):
            if timeline is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                timeline.append((label, ev))

        mark("backward_done", main)
        L = exchange.n_layers
        if getattr(self, "_opt_stream", None) is None:
            self._opt_stream = torch.cuda.Stream(device=dev)
        os_ = self._opt_stream  # updates and re-packs (HBM-bound) run next to the collectives (NVLink-bound) of other blocks

        def update(layer):     # this rank's 1/world of block `layer`, once its gradients are reduced
            os_.wait_event(scattered[layer])
            eng.adamw_step(*args, group=layer, stream=os_)
            updated[layer] = torch.cuda.Event()
            updated[layer].record(os_)

        def gather(layer):     # the block's new bf16 weights from every rank, then the packed GEMM operands
            xs.wait_event(updated[layer])
            exchange.gather_weights_layer(layer)
            ev = torch.cuda.Event()
            ev.record(xs)
            os_.wait_event(ev)
            eng.optimizer_pack_group(layer, stream=os_)
            if overlap_forward:
                eng.weights_record_ready(layer, os_)  # the next forward waits for this block only when it gets there
            mark(f"weights_{layer}", os_)

        def rest():            # small and non-block tensors: all-reduced, updated on every rank
            exchange.reduce_tail()
            os_.wait_stream(xs)
            eng.adamw_step(*args, group=L, stream=os_)
            mark("adamw_rest", os_)
            if overlap_forward:
                main.wait_stream(os_)  # embeddings, norms, router: read by the first launches of the next forward

        scattered, updated = {}, {}
        backward_order = list(range(L - 1, -1, -1))  # the backward finishes the last block first
        if loss_scale is None:
            # the scale is produced on the caller's stream after the backward: reduce-scatter everything while the
            # backward runs, then update + gather, software-pipelined over the two streams — block 0 first when the next
            # forward may overlap (it needs block 0 first)
            for layer in backward_order:
                exchange.reduce_scatter_layer(layer)
                scattered[layer] = torch.cuda.Event()
                scattered[layer].record(xs)
                mark(f"scatter_{layer}", xs)
            os_.wait_stream(main)
            order = list(range(L)) if overlap_forward else backward_order
            for i, layer in enumerate(order):
                update(layer)
                if i == 0:
                    rest()  # behind the first block's update (which only needs its own reduce-scatter), ahead of its gather
                if i >= 1:
                    gather(order[i - 1])
            gather(order[-1])
        else:
            # collectives in the order scatter(l), gather(l + 1): the update of block l runs while block l - 1 is on the wire
            order = backward_order
            for i, layer in enumerate(order):
                exchange.reduce_scatter_layer(layer)
                scattered[layer] = torch.cuda.Event()
                scattered[layer].record(xs)
                mark(f"scatter_{layer}", xs)
                update(layer)
                if i >= 1:
                    gather(order[i - 1])
            rest()
            gather(order[-1])
        if not overlap_forward:
            main.wait_stream(os_)
        self._masters_stale = True
        if master_sync == "step":
            self._gather_masters()
        elif master_sync != "lazy":
            raise ValueError("master_sync must be 'step' or 'lazy'")

    def _gather_masters(self) -> None:
        ex = getattr(self, "_exchange", None)
        if ex is None or not self._masters_stale:
            return
        ex.stream.wait_stream(self._opt_stream)  # every owner's update of this step
        ex.gather_parameters(dict(self.inner.named_parameters()))
        self._masters_ready = torch.cuda.Event()
        self._masters_ready.record(ex.stream)
        self._masters_stale = False

    def synchronize_parameters(self) -> None:
        """After sharded steps: make the caller's stream see fully updated fp32 master parameters on every rank (a
        collective if the gather is still pending: call it on all ranks). No-op otherwise."""
        self._gather_masters()
        if self._masters_ready is not None:
            torch.cuda.current_stream(self.inner._engine.device).wait_event(self._masters_ready)

    def _gather_sharded_state(self, eng) -> None:
        """Moments and EMA of sharded tensors live on the owning rank: gather them before they are read whole."""
        ex = getattr(self, "_exchange", None)
        if ex is None or ex.layers is None or self._step == 0:
            return
        m, v = eng.optimizer_state()
        bufs = [m, v] + ([eng.ema_state()] if self._ema_decay is not None else [])
        if getattr(self, "_opt_stream", None) is not None:
            ex.stream.wait_stream(self._opt_stream)
        for b in bufs:
            ex.gather_flat(b)
        torch.cuda.current_stream(eng.device).wait_stream(ex.stream)

    def state_dict(self):
        """Moments, the EMA buffer (the reference's EMA callback checkpoints its averages, ema.py:137-160) and the
        position of the counter-based train-mode random stream, so a resumed run neither re-seeds the average nor replays
        the same dropout masks / expert draws."""
        sd = {"step": self._step, "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}],
              "train_rng": (getattr(self.inner, "_train_seed", None), getattr(self.inner, "_train_step", 0))}
        eng = getattr(self.inner, "_engine", None)
        if eng is not None and self._step > 0:
            self._gather_sharded_state(eng)  # sharded steps: a collective, state_dict() must then run on every rank
            m, v = eng.optimizer_state()
            sd["exp_avg"], sd["exp_avg_sq"] = m.clone(), v.clone()
            if self._ema_decay is not None:
                sd["ema"] = eng.ema_state().clone()
        return sd

    def load_state_dict(self, sd):
        self._step = int(sd["step"])
        for k, v in sd["param_groups"][0].items():
            self.param_groups[0][k] = v
        seed, step = sd.get("train_rng", (None, 0))
        if seed is not None:
            self.inner.set_train_rng(seed, step)
        if "exp_avg" in sd:
            eng = self.inner._ensure_engine(1)
            self._bind(eng)
            m, v = eng.optimizer_state()
            m.copy_(sd["exp_avg"])
            v.copy_(sd["exp_avg_sq"])
            if "ema" in sd and self._ema_decay is not None:
                self._set_ema(eng)          # allocates the buffer (seeded from the current weights) ...
                eng.ema_state().copy_(sd["ema"])  # ... then restore the checkpointed average
                eng.ema_mark_seeded()
