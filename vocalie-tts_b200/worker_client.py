#!/usr/bin/env python
"""Runner-protocol client of the resident B200 worker (stdlib only: starts in milliseconds).

The reference drives every engine through ``SubprocessBackendMixin._run_subprocess``
(tts_backends/base_runner.py:211-276): ``subprocess.run([venv python, tts_backends/<runner_module>.py],
input=json.dumps(payload))``, one JSON object back on stdout, exit code 0 / 1.  The stock Chatterbox runner pays an
interpreter + torch import + model load for EVERY chunk that way (tts_backends/chatterbox_runner.py:136).  This file is
a drop-in ``runner_module``: same stdin / stdout / exit-code contract, but it only forwards the payload to the
long-lived worker (``python -m vocalie_tts_b200.worker serve``) over a Unix socket and relays its answer - the model
stays resident on the GPU.

Socket path: ``$VOCALIE_B200_SOCKET`` (default ``/tmp/vocalie_b200.sock``).
"""
import json
import os
import socket
import sys


def main() -> int:
    raw = sys.stdin.read()
    try:
        payload = json.loads(raw) if raw.strip() else {}
        path = os.environ.get("VOCALIE_B200_SOCKET", "/tmp/vocalie_b200.sock")
        timeout = float(os.environ.get("VOCALIE_B200_TIMEOUT_S", "600"))      # the stock runner's timeout (chatterbox_backend.py:19)
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
            s.settimeout(timeout)
            s.connect(path)
            s.sendall((json.dumps(payload) + "\n").encode())
            chunks = []
            while True:
                b = s.recv(65536)
                if not b:
                    break
                chunks.append(b)
                if b.endswith(b"\n"):
                    break
        line = b"".join(chunks).decode().strip()
        resp = json.loads(line)
    except Exception as exc:  # noqa: BLE001 - everything becomes the protocol's error object
        resp = {"ok": False, "error": f"vocalie_b200 worker unreachable: {exc}"}
    sys.stdout.write(json.dumps(resp))
    sys.stdout.flush()
    return 0 if resp.get("ok") else 1


if __name__ == "__main__":
    raise SystemExit(main())
