"""torchrun --nproc-per-node 2: a ragged batch sharded over two GPUs (one NCCL weight broadcast, shards padded to the
global frame count, host-side gather) equals the same batch run unsharded on rank 0.  Prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def c_abi_broadcast_check(model, rank, world, dev):
    """f5_nccl_broadcast_weights (the C entry a non-Python host uses for the ONE collective of the path) on a raw
    ncclComm_t created through NCCL's C API: every non-root rank scrambles a copy of the packed buffer, the call must
    restore rank 0's bytes."""
    import ctypes as C
    from f5_tts_mlx_b200 import _lib
    nccl = C.CDLL("libnccl.so.2")

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_byte * 128)]

    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, src=0)
    C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    buf = model.packed.buffer.clone()
    want = int(buf.to(torch.int64).sum().item())
    if rank != 0:
        buf.random_(0, 255)
    st = torch.cuda.current_stream().cuda_stream
    rc = _lib.load().f5_nccl_broadcast_weights(C.c_void_p(buf.data_ptr()), buf.numel(), 0, comm, C.c_void_p(st))
    torch.cuda.synchronize()
    ok = rc == 0 and int(buf.to(torch.int64).sum().item()) == want and torch.equal(buf, model.packed.buffer)
    oks = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    assert oks.item() == 1, "f5_nccl_broadcast_weights did not reproduce the root's packed weights"
    return True


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from f5_tts_mlx_b200 import DiT, F5TTS, GATE_CONFIG
    from f5_tts_mlx_b200.parallel import load_weights_distributed, sample_sharded
    from f5_tts_mlx_b200.weights import random_dit_weights
    cfg = GATE_CONFIG
    model = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim,
                text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, device=dev)
    load_weights_distributed(model, lambda: random_dit_weights(cfg, seed=1234))        # rank 0 packs, ONE broadcast
    c_bcast = c_abi_broadcast_check(model, rank, world, dev)
    f5 = F5TTS(model)
    g = torch.Generator().manual_seed(77)
    B = 5                                                                              # 3 + 2 utterances
    cond = (torch.randn(B, 60, 100, generator=g) * 2.24 - 1.27)
    text = torch.randint(0, 2545, (B, 40), generator=g, dtype=torch.int32)
    text[1, 25:] = -1; text[4, 12:] = -1
    dur = torch.tensor([150, 131, 200, 97, 180])
    y0 = torch.randn(B, 200, 100, generator=g)
    for i, d in enumerate(dur.tolist()):
        y0[i, d:] = 0
    kw = dict(steps=4, method="midpoint", cfg_strength=2.0, sway_sampling_coef=-1.0)
    outs = sample_sharded(f5, cond.to(dev), text, dur, y0=y0.to(dev), **kw)
    res = None
    if rank == 0:
        full, _ = f5.sample(cond.to(dev), text, dur, y0=y0.to(dev), return_trajectory=False, **kw)
        full = full.cpu()
        assert len(outs) == B
        diffs, bitwise = [], True
        for i in range(B):
            d = (outs[i] - full[i]).abs().max().item()
            diffs.append(d)
            bitwise = bitwise and torch.equal(outs[i], full[i])
        rel = max(((outs[i] - full[i]).norm() / full[i].norm()).item() for i in range(B))
        res = {"world": world, "utterances": B, "max_abs": max(diffs), "max_rel": rel, "bitwise_equal": bitwise,
               "frames": int(full.shape[1]), "c_abi_nccl_broadcast": c_bcast}
        print("NCCL_SHARD_CHECK " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not (res["max_rel"] < 1e-5):
        sys.exit(1)


if __name__ == "__main__":
    main()
