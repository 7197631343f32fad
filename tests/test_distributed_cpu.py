"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: the ONE collective of the path — the
broadcast of the packed weight buffer — plus utterance sharding and the host-side gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from f5_tts_mlx_b200.dit import DiT
        from f5_tts_mlx_b200.parallel import gather_objects, global_frames, load_weights_distributed, shard_range
        from f5_tts_mlx_b200.weights import DiTConfig, random_dit_weights
        cfg = DiTConfig(dim=256, depth=2, heads=4, text_num_embeds=50, text_dim=512, conv_layers=1)
        model = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_num_embeds=cfg.text_num_embeds,
                    text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, device="cpu")
        built = []

        def weights_fn():
            built.append(rank)
            return random_dit_weights(cfg, seed=99)

        load_weights_distributed(model, weights_fn, src=0)
        checksum = model.packed.buffer.to(torch.int64).sum().item()
        # sharding: 5 utterances over 2 ranks, host-side gather in rank order
        mine = [f"utt{i}" for i in shard_range(5, world, rank)]
        gathered = gather_objects(mine)
        n_glob = global_frames(100 + 37 * rank)
        q.put((rank, built, checksum, model.packed.nbytes, gathered, n_glob))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_weight_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, built0, sum0, nb0, g0, n0), (r1, built1, sum1, nb1, g1, n1) = res
    assert built0 == [0] and built1 == []            # only the source rank materialises the weights
    assert sum0 == sum1 and sum0 != 0 and nb0 == nb1  # identical packed buffers after the broadcast
    assert g0 == [f"utt{i}" for i in range(5)] and g1 is None
    assert n0 == n1 == 137                            # every shard pads to the GLOBAL max frames
