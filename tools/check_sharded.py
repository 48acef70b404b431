"""Multi-GPU correctness check (torchrun, NCCL): the reference-order job sharded over the ranks must be BIT-IDENTICAL to
the same job on one GPU - contiguous and LPT shards, with and without editing.  Explicit F0 / phase / noise so that the
in-kernel generator (keyed on the sequence's index in its batch) does not enter.  Rank 0 prints one line per case.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import hift_oracle as H                                   # noqa: E402  (synthetic inputs only)
from vocalie_tts_b200 import distributed as D                          # noqa: E402
from vocalie_tts_b200.backend import shard_chunks                      # noqa: E402
from vocalie_tts_b200.hift import HiFTVocoder                          # noqa: E402
from vocalie_tts_b200.pipeline import VocoderPipeline                  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    voc = HiFTVocoder(H.make_state_dict(0, "unit"), operand="fp16")
    Ts = [30, 12, 45, 7, 60, 22, 9, 38, 51][: max(5, 2 * world + 1)]
    mels = [H.synth_mel(T, 77, b) for b, T in enumerate(Ts)]
    f0s = [H.synth_f0(T, 77, b) for b, T in enumerate(Ts)]
    pn = [H.synth_noise(T, 77, b) for b, T in enumerate(Ts)]
    ok_all = True
    for mode in ("contiguous", "lpt"):
        for kw in (dict(), dict(trim_silence=False, normalize=False), dict(trim_silence=True, normalize=False, target_dbfs=-3.0)):
            shards = D.contiguous_shards(len(Ts), world) if mode == "contiguous" else shard_chunks(Ts, world)
            ids = shards[rank]
            pipe = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True, **kw)
            job = D.ShardedJob(pipe, Ts, shards)
            if ids:
                mel, _ = voc.pack_mels([mels[i] for i in ids])
                loc = dict(f0=torch.cat([f0s[i] for i in ids]).cuda(), phase_vec=torch.stack([pn[i][0] for i in ids]).cuda().contiguous(),
                           noise=torch.cat([pn[i][1].reshape(-1) for i in ids]).cuda())
            else:
                mel, loc = torch.zeros((0, 80), device="cuda"), dict()
            res = job.run_device(mel, **loc)
            if rank == 0:
                single = VocoderPipeline(voc, chunk_gap_ms=250, out_pcm16=True, **kw)
                melf, Tf = voc.pack_mels(mels)
                ref = single.run_device(melf, Tf, f0=torch.cat(f0s).cuda(), phase_vec=torch.stack([p for p, _ in pn]).cuda().contiguous(),
                                        noise=torch.cat([n.reshape(-1) for _, n in pn]).cuda(), read_back=True)
                want = ref.audio[: ref.total_samples]
                got = res.audio
                ok = got.numel() == want.numel() and bool(torch.equal(got, want))
                if res.edit is not None:
                    ok = ok and res.edit["peak_before"] == ref.edit["peak_before"] and res.edit["start_sample"] == ref.edit["start_sample"] \
                        and res.edit["end_sample"] == ref.edit["end_sample"]
                ok_all = ok_all and ok
                print(f"sharded-check world={world} {mode} {kw or 'trim+normalise'}: {'OK' if ok else 'MISMATCH'} ({got.numel()} samples)", flush=True)
            # the same job assembled on the HOST: every rank copies its own pieces into a buffer shared by the ranks
            shared = D.SharedHostBuffer(f"vocalie_b200_check_{os.environ.get('MASTER_PORT', '0')}", job.n_raw + 8, torch.int16)
            res_h = job.run_device(mel, host_out=shared.tensor, **loc)
            job.wait_host()
            shared.publish(0)
            shared.wait_complete(0)
            if rank == 0:
                got_h = shared.tensor[: res_h.total_samples].clone()
                ok = res_h.total_samples == want.numel() and bool(torch.equal(got_h, want.cpu()))
                ok_all = ok_all and ok
                print(f"sharded-check world={world} {mode} {kw or 'trim+normalise'} host-assembled: {'OK' if ok else 'MISMATCH'}", flush=True)
            shared.close()
    if rank == 0:
        print("SHARDED-CHECK", "PASS" if ok_all else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
