"""Pickle-over-HTTP agent proxy with the reference's wire format (`mode/evaluation/agent_proxy.py:91-203`), plus a
batching front-end that lets many simulator processes share one engine.

Wire format (unchanged, so the reference's evaluation clients keep working): every request is an HTTP POST whose body is
`pickle.dumps({"method": str, "args": [...], "kwargs": {...}})`; the reply body is `pickle.dumps({"result": ...})` with
status 200, or `pickle.dumps({"error": "there was a problem"})` with status 500. `method` is `"__init__"` (create the
agent), `"__shutdown__"` (destroy it and stop the server), `"__call__"` (`agent(*args, **kwargs)`), or the name of an
agent method (`step`, `reset`, ...), reference agent_proxy.py:95-126.

Two differences from the reference, both on the server side only:
* request bodies are unpickled with an allow-list (`RestrictedUnpickler`: containers, scalars, numpy arrays, torch
  tensors) instead of bare `pickle.loads` on network input, and the default bind address is 127.0.0.1;
* `BatchingPolicy` (optional) coalesces concurrent `step`/`__call__` requests from different environments into ONE
  engine call: the reference runs the sampler at B=1 per environment (mode_agent.py:584-637), which leaves a B200 idle;
  trajectories are independent, so n waiting requests become one batch of n (`ThreadingHTTPServer`, one thread per
  client connection, a short gathering window).
"""
from __future__ import annotations

import io
import pickle
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from logging import getLogger
from typing import Callable, Optional

logger = getLogger(__name__)

_ALLOWED = {
    ("builtins", n) for n in ("dict", "list", "tuple", "set", "frozenset", "str", "bytes", "bytearray", "int", "float",
                              "bool", "complex", "slice", "range", "NoneType")
} | {
    ("collections", "OrderedDict"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"), ("numpy.core.numeric", "_frombuffer"),
    ("numpy._core.numeric", "_frombuffer"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch", "Size"), ("torch", "device"),
    ("torch.storage", "_load_from_bytes"), ("torch._tensor", "_rebuild_from_type_v2"), ("torch", "Tensor"),
} | {("torch", f"{t}Storage") for t in ("Float", "Double", "Half", "BFloat16", "Long", "Int", "Short", "Char", "Byte", "Bool")} \
  | {("torch", t) for t in ("float32", "float64", "float16", "bfloat16", "int64", "int32", "int16", "int8", "uint8", "bool")}


class RestrictedUnpickler(pickle.Unpickler):
    """Unpickles plain data (what simulator clients send: strings, numbers, containers, numpy arrays, torch tensors) and
    refuses every other global — a request cannot name code to run on the server."""

    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"agent proxy: global {module}.{name} is not allowed in a request")


def loads_request(data: bytes):
    return RestrictedUnpickler(io.BytesIO(data)).load()


class AgentHandler(BaseHTTPRequestHandler):
    create_agent: Optional[Callable] = None
    agent = None
    lock = threading.Lock()  # __init__/__shutdown__ are serialised; agent calls are not (see BatchingPolicy)

    def log_message(self, fmt, *args):  # route http.server's stderr chatter through logging
        logger.debug(fmt, *args)

    def _reply(self, code: int, payload: dict) -> None:
        self.send_response(code)
        self.end_headers()
        self.wfile.write(pickle.dumps(payload))

    def do_POST(self):  # noqa: N802  (name fixed by http.server)
        cls = type(self)
        try:
            request = loads_request(self.rfile.read(int(self.headers["Content-Length"])))
            method = request.get("method")
            if method == "__shutdown__":
                with cls.lock:
                    cls._destroy_agent()
                self._reply(200, {"result": "shutdown"})
                threading.Thread(target=self.server.shutdown, daemon=True).start()  # stop serve_forever from outside it
                return
            if method == "__init__":
                with cls.lock:
                    cls.agent = cls.create_agent()
                self._reply(200, {"result": "agent created"})
                return
            agent = cls.agent
            if agent is None:
                raise RuntimeError("no agent: send {'method': '__init__'} first")
            args, kwargs = request.get("args", []), request.get("kwargs", {})
            result = agent(*args, **kwargs) if method == "__call__" else getattr(agent, method)(*args, **kwargs)
            self._reply(200, {"result": result})
        except Exception:
            logger.exception("Error handling request")
            with cls.lock:
                cls._destroy_agent()
            self._reply(500, {"error": "there was a problem"})

    @classmethod
    def _destroy_agent(cls):
        agent, cls.agent = cls.agent, None
        close = getattr(agent, "close", None)
        if callable(close):
            close()
        del agent
        try:
            import gc

            import torch

            gc.collect()
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        except Exception:  # pragma: no cover
            pass
        logger.info("agent destroyed")


def make_server(create_agent: Callable, host: str = "127.0.0.1", port: int = 6000) -> ThreadingHTTPServer:
    """The HTTP server (not yet serving). Each server gets its own handler class so several can live in one process."""
    handler = type("BoundAgentHandler", (AgentHandler,), {"create_agent": staticmethod(create_agent), "agent": None,
                                                          "lock": threading.Lock()})
    server = ThreadingHTTPServer((host, port), handler)
    server.daemon_threads = True
    return server


def start_server(create_agent: Callable, host: str = "127.0.0.1", port: int = 6000) -> None:
    """Blocking equivalent of the reference's `start_server` (agent_proxy.py:158-168)."""
    server = make_server(create_agent, host, port)
    logger.info("starting server at http://%s:%d", host, port)
    try:
        server.serve_forever()
    except KeyboardInterrupt:
        logger.info("shutting down server")
    finally:
        server.server_close()


class AgentClient:
    """Client side of the wire format (what the reference's evaluation scripts implement ad hoc)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 6000, timeout: float = 60.0):
        self.url, self.timeout = f"http://{host}:{port}", timeout

    def request(self, method: str, *args, **kwargs):
        import urllib.error
        import urllib.request

        body = pickle.dumps({"method": method, "args": list(args), "kwargs": kwargs})
        req = urllib.request.Request(self.url, data=body, method="POST")
        try:
            with urllib.request.urlopen(req, timeout=self.timeout) as resp:
                reply = pickle.loads(resp.read())  # noqa: S301  (the server is the trusted side)
        except urllib.error.HTTPError as err:
            reply = pickle.loads(err.read())  # noqa: S301
        if "error" in reply:
            raise RuntimeError(f"agent proxy: {reply['error']}")
        return reply["result"]

    def init(self):
        return self.request("__init__")

    def shutdown(self):
        return self.request("__shutdown__")

    def __call__(self, *args, **kwargs):
        return self.request("__call__", *args, **kwargs)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return lambda *a, **k: self.request(name, *a, **k)


class BatchingPolicy:
    """Coalesces concurrent single-environment requests into one batched policy call.

    `batch_fn(list_of_requests) -> list_of_results` is called with every request that arrived within `window_s` of the
    first one (at most `max_batch`). Typical `batch_fn`: stack the per-environment observation tokens / goals, run ONE
    `DenoisingPolicy.denoise_actions` over the stacked batch through the engine, split the (n, 10, 7) result.
    Each HTTP client thread blocks in `__call__` / `step` until its own result is ready."""

    def __init__(self, batch_fn: Callable, max_batch: int = 256, window_s: float = 0.002):
        self.batch_fn, self.max_batch, self.window_s = batch_fn, max_batch, window_s
        self._cv = threading.Condition()
        self._pending: list = []
        self._closed = False
        self.batch_sizes: list[int] = []  # bookkeeping: sizes of the batches actually run
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    def __call__(self, request):
        slot = {"request": request, "done": threading.Event(), "result": None, "error": None}
        with self._cv:
            if self._closed:
                raise RuntimeError("BatchingPolicy is closed")
            self._pending.append(slot)
            self._cv.notify_all()
        slot["done"].wait()
        if slot["error"] is not None:
            raise slot["error"]
        return slot["result"]

    step = __call__

    def _run(self):
        while True:
            with self._cv:
                while not self._pending and not self._closed:
                    self._cv.wait()
                if self._closed and not self._pending:
                    return
                deadline = time.monotonic() + self.window_s
                while len(self._pending) < self.max_batch and not self._closed:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self._cv.wait(left)
                batch, self._pending = self._pending[: self.max_batch], self._pending[self.max_batch:]
            try:
                results = self.batch_fn([s["request"] for s in batch])
                if len(results) != len(batch):
                    raise RuntimeError(f"batch_fn returned {len(results)} results for {len(batch)} requests")
                for s, r in zip(batch, results):
                    s["result"] = r
            except Exception as exc:  # every waiting client sees the failure
                for s in batch:
                    s["error"] = exc
            self.batch_sizes.append(len(batch))
            for s in batch:
                s["done"].set()

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join(timeout=5)


def denoising_batch_fn(policy, device="cuda"):
    """`batch_fn` for BatchingPolicy over a `DenoisingPolicy`: each request is a dict with `state_images` (n_img, obs_dim)
    and `latent_goal` (goal_dim,) arrays (what MoDEAgent.embed_visual_obs / the language encoder produce for ONE
    environment, mode_agent.py:548-567); the reply is that environment's (act_window, action_dim) numpy action chunk."""
    import numpy as np
    import torch

    def run(requests):
        state = torch.as_tensor(np.stack([np.asarray(r["state_images"], np.float32) for r in requests])).to(device)
        goal = torch.as_tensor(np.stack([np.asarray(r["latent_goal"], np.float32).reshape(-1) for r in requests])).to(device)
        out = policy.denoise_actions(None, {"state_images": state}, goal, inference=True)
        out = out.float().cpu().numpy()
        return [out[i] for i in range(len(requests))]

    return run
