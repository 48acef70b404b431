"""Resident worker behind the reference's subprocess-runner protocol.

The stock boundary spawns ``tts_backends/chatterbox_runner.py`` per chunk: a new interpreter, ``import torch``, a model
load, one chunk, exit (tts_backends/base_runner.py:229, chatterbox_runner.py:116-172).  ``ResidentWorker`` keeps the
engine (HiFT weights packed on the GPU, the upstream text -> mel callable) alive in ONE process and answers the same
JSON payloads:

    request : {"text", "out_path" | "out_wav_path", "voice_ref_path" | "ref_audio_path", "lang" | "language" |
               "tts_language", "tts_model_mode" | "chatterbox_mode", "multilang_cfg_weight", "exaggeration",
               "cfg_weight", "temperature", "repetition_penalty"}          (chatterbox_runner.py:119-131)
    response: {"ok": true, "out_path", "duration_s", "retry", "logs"}      (chatterbox_runner.py:155-163)
              {"ok": false, "error", "trace"}                              (chatterbox_runner.py:165-172)

Transports: ``serve_stream`` (one JSON object per line on stdin / stdout - a pipe a parent keeps open) and
``serve_socket`` (Unix socket, one request per connection; ``worker_client.py`` is the runner-compatible CLI the
unmodified ``SubprocessBackendMixin`` spawns instead of the stock runner).  The WAV is written by
``ChatterboxB200Backend.synthesize``: device PCM_16 encode behind a device-written RIFF header, one D2H copy.

    python -m vocalie_tts_b200.worker serve --socket /tmp/vocalie_b200.sock --factory mypkg.engine:make
        (``make()`` returns {"state_dict" | "vocoder": ..., "mel_provider": callable})
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import socket
import sys
import threading
import traceback
from pathlib import Path
from typing import Any, Callable, Dict, Optional

_PARAM_KEYS = ("multilang_cfg_weight", "exaggeration", "cfg_weight", "temperature", "repetition_penalty")


class ResidentWorker:
    def __init__(self, synthesize: Optional[Callable[..., Dict[str, Any]]] = None):
        """``synthesize(script, out_path, voice_ref_path=None, lang=None, **params) -> meta`` - by default the
        configured ``ChatterboxB200Backend`` (class-level engine, so it is resident by construction)."""
        if synthesize is None:
            from .backend import ChatterboxB200Backend
            synthesize = lambda script, out_path, **kw: ChatterboxB200Backend().synthesize(script, out_path, **kw)  # noqa: E731
        self._synthesize = synthesize
        self.served = 0
        self._stop = threading.Event()

    # ------------------------------------------------------------------ protocol
    def handle(self, payload: Dict[str, Any]) -> Dict[str, Any]:
        """One request of the runner protocol -> its response object (never raises)."""
        try:
            if not isinstance(payload, dict):
                raise ValueError("payload must be a JSON object")
            if payload.get("op") == "ping":
                return {"ok": True, "pong": True, "served": self.served}
            if payload.get("op") == "shutdown":
                self._stop.set()
                return {"ok": True, "shutdown": True}
            text = str(payload.get("text") or "")
            out_path = payload.get("out_wav_path") or payload.get("out_path")
            if not out_path:
                raise ValueError("out_wav_path is required")                    # chatterbox_runner.py:122-123
            out_path = str(Path(out_path).expanduser().resolve())
            params: Dict[str, Any] = {"tts_model_mode": str(payload.get("tts_model_mode") or payload.get("chatterbox_mode") or "fr_finetune")}
            for k in _PARAM_KEYS:
                if k in payload:
                    params[k] = float(payload[k])
            if payload.get("seed") is not None:
                params["seed"] = int(payload["seed"])
            meta = self._synthesize(text, out_path,
                                    voice_ref_path=payload.get("ref_audio_path") or payload.get("voice_ref_path"),
                                    lang=payload.get("lang") or payload.get("language") or payload.get("tts_language"), **params)
            self.served += 1
            return {"ok": True, "out_path": out_path, "duration_s": float(meta.get("duration_s", 0.0)),
                    "retry": bool(meta.get("retry")), "logs": []}
        except Exception as exc:  # noqa: BLE001 - the protocol carries every failure as an object
            return {"ok": False, "error": str(exc), "trace": traceback.format_exc()}

    # ------------------------------------------------------------------ transports
    def serve_stream(self, fin=None, fout=None) -> None:
        """One JSON object per line in, one per line out, until EOF or {"op": "shutdown"}."""
        fin, fout = fin or sys.stdin, fout or sys.stdout
        for line in fin:
            if not line.strip():
                continue
            try:
                req = json.loads(line)
            except json.JSONDecodeError as exc:
                resp = {"ok": False, "error": f"invalid JSON: {exc}"}
            else:
                resp = self.handle(req)
            fout.write(json.dumps(resp) + "\n")
            fout.flush()
            if self._stop.is_set():
                break

    def serve_socket(self, path: str, *, ready: Optional[threading.Event] = None) -> None:
        """Unix-socket server: one request line per connection, handled on a thread per connection (the engine
        serialises GPU work with its own lock; reference jobs run on up to 2 threads, backend/config.py:11)."""
        try:
            os.unlink(path)
        except FileNotFoundError:
            pass
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(path)
        srv.listen(16)
        srv.settimeout(0.2)
        if ready is not None:
            ready.set()
        try:
            while not self._stop.is_set():
                try:
                    conn, _ = srv.accept()
                except socket.timeout:
                    continue
                threading.Thread(target=self._serve_conn, args=(conn,), daemon=True).start()
        finally:
            srv.close()
            try:
                os.unlink(path)
            except FileNotFoundError:
                pass

    def _serve_conn(self, conn: socket.socket) -> None:
        with conn:
            try:
                conn.settimeout(600)
                buf = b""
                while not buf.endswith(b"\n"):
                    b = conn.recv(65536)
                    if not b:
                        break
                    buf += b
                resp = self.handle(json.loads(buf.decode()))
            except Exception as exc:  # noqa: BLE001
                resp = {"ok": False, "error": f"bad request: {exc}"}
            try:
                conn.sendall((json.dumps(resp) + "\n").encode())
            except OSError:
                pass

    def stop(self) -> None:
        self._stop.set()


def _load_factory(spec: str):
    mod, _, fn = spec.partition(":")
    return getattr(importlib.import_module(mod), fn or "make")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    sv = sub.add_parser("serve")
    sv.add_argument("--socket", default=os.environ.get("VOCALIE_B200_SOCKET"))
    sv.add_argument("--factory", required=True, help="module:function returning {'state_dict'|'vocoder', 'mel_provider'}")
    sv.add_argument("--operand", default="fp16")
    a = ap.parse_args(argv)
    from .backend import ChatterboxB200Backend
    eng = _load_factory(a.factory)()
    ChatterboxB200Backend.configure(state_dict=eng.get("state_dict"), vocoder=eng.get("vocoder"),
                                    mel_provider=eng["mel_provider"], operand=a.operand)
    w = ResidentWorker()
    print(f"[vocalie_b200 worker] engine resident, serving on {a.socket or 'stdin/stdout'}", file=sys.stderr, flush=True)
    if a.socket:
        w.serve_socket(a.socket)
    else:
        w.serve_stream()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
