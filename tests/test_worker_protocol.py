"""CPU tests of the resident worker's wire protocol: the reference's OWN ``SubprocessBackendMixin._run_subprocess``
(tts_backends/base_runner.py:211-276, imported unchanged; skipped where /root/reference is absent) spawns
``worker_client.py`` as its runner and must get the stock runner's response objects back from a worker that stays
alive across requests; plus the line-delimited stream transport.  The engine is a stub (no GPU here): the protocol,
not the vocoder, is under test - the byte-exact WAV is a GPU test (tests/test_wav_gpu.py)."""
import io
import json
import os
import sys
import threading
import wave
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
CLIENT = ROOT / "vocalie-tts_b200" / "worker_client.py"


def _stub_synthesize(calls):
    def synth(script, out_path, voice_ref_path=None, lang=None, **params):
        if not script.strip():
            raise ValueError("Texte vide.")
        calls.append((script, lang, dict(params), voice_ref_path))
        n = 480 * (10 + len(script))
        q = (np.sin(np.arange(n) * 0.05) * 8000).astype("<i2")
        with wave.open(out_path, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(24000); w.writeframes(q.tobytes())
        return {"duration_s": n / 24000.0, "retry": False}
    return synth


MIXIN_SCRIPT = r'''
import json, os, sys, threading, types, wave
from pathlib import Path
import numpy as np
ROOT, REF, TMP = Path(sys.argv[1]), Path(sys.argv[2]), Path(sys.argv[3])
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(REF))
from oracle import sf_stub
sys.modules.setdefault("librosa", types.ModuleType("librosa"))
sys.modules["soundfile"] = sf_stub.make_module()
from tts_backends.base_runner import SubprocessBackendMixin          # the reference's own mixin, unchanged
from tts_backends.base import BackendUnavailableError
from vocalie_tts_b200.worker import ResidentWorker

calls = []
def synth(script, out_path, voice_ref_path=None, lang=None, **params):
    if not script.strip():
        raise ValueError("Texte vide.")
    calls.append((script, lang, dict(params), voice_ref_path))
    n = 480 * (10 + len(script))
    q = (np.sin(np.arange(n) * 0.05) * 8000).astype("<i2")
    with wave.open(out_path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(24000); w.writeframes(q.tobytes())
    return {"duration_s": n / 24000.0, "retry": False}

w = ResidentWorker(synth)
sock = str(TMP / "w.sock")
ready = threading.Event()
t = threading.Thread(target=w.serve_socket, args=(sock,), kwargs={"ready": ready}, daemon=True)
t.start()
assert ready.wait(10)
os.environ["VOCALIE_B200_SOCKET"] = sock

class B200Runner(SubprocessBackendMixin):
    runner_module = "worker_client"
    runner_venv = "chatterbox"
    def _runner_path(self):
        return ROOT / "vocalie-tts_b200" / "worker_client.py"
    def _python_path(self):
        return Path(sys.executable)

def expect_error(payload, needle):
    try:
        r._run_subprocess(payload)
    except BackendUnavailableError as exc:
        assert needle in str(exc), str(exc)
    else:
        raise SystemExit(f"expected BackendUnavailableError({needle})")

r = B200Runner()
for i, text in enumerate(["Bonjour le monde.", "Deuxieme morceau, meme processus."]):
    out = TMP / f"c{i}.wav"
    data = r._run_subprocess({"text": text, "out_path": str(out), "lang": "fr", "temperature": 0.7, "chatterbox_mode": "multilang"})
    assert data["ok"] is True and data["retry"] is False and data["logs"] == []
    assert Path(data["out_path"]) == out.resolve() and out.exists()
    assert abs(data["duration_s"] - 480 * (10 + len(text)) / 24000.0) < 1e-9
assert w.served == 2 and len(calls) == 2                      # ONE worker process served both chunks
assert calls[0][1] == "fr" and calls[0][2]["temperature"] == 0.7 and calls[0][2]["tts_model_mode"] == "multilang"
# failures travel as the protocol's error object and surface as the reference's own error type
expect_error({"text": "x"}, "out_wav_path is required")
expect_error({"text": "   ", "out_path": str(TMP / "e.wav")}, "Texte vide")
# no worker listening -> loud failure, not a hang and not a silent fallback
os.environ["VOCALIE_B200_SOCKET"] = str(TMP / "nobody.sock")
expect_error({"text": "x", "out_path": str(TMP / "f.wav")}, "worker unreachable")
w.stop(); t.join(5)
print(json.dumps({"ok": True, "served": w.served}))
'''


@pytest.mark.skipif(not REF.exists(), reason="/root/reference is not present (GPU box)")
def test_reference_mixin_drives_the_resident_worker(tmp_path):
    import subprocess
    r = subprocess.run([sys.executable, "-c", MIXIN_SCRIPT, str(ROOT), str(REF), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert json.loads(r.stdout.strip().splitlines()[-1]) == {"ok": True, "served": 2}


def test_stream_transport_and_control_ops(tmp_path):
    from vocalie_tts_b200.worker import ResidentWorker
    calls = []
    w = ResidentWorker(_stub_synthesize(calls))
    reqs = [{"op": "ping"}, {"text": "Salut.", "out_wav_path": str(tmp_path / "a.wav"), "ref_audio_path": "/v.wav", "language": "fr"},
            {"text": "x"}, {"op": "shutdown"}, {"text": "never served", "out_path": str(tmp_path / "b.wav")}]
    fin = io.StringIO("".join(json.dumps(r) + "\n" for r in reqs) + "not json\n")
    fout = io.StringIO()
    w.serve_stream(fin, fout)
    out = [json.loads(l) for l in fout.getvalue().splitlines()]
    assert len(out) == 4                                           # stops at shutdown
    assert out[0] == {"ok": True, "pong": True, "served": 0}
    assert out[1]["ok"] and set(out[1]) == {"ok", "out_path", "duration_s", "retry", "logs"}
    assert calls[0][3] == "/v.wav" and calls[0][1] == "fr"
    assert out[2]["ok"] is False and "out_wav_path is required" in out[2]["error"] and "trace" in out[2]
    assert out[3]["shutdown"] is True and not (tmp_path / "b.wav").exists()
