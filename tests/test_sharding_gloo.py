"""world_size-2 gloo test (CPU) of the N > 1 host logic: contiguous sample sharding covers the batch exactly, the
per-rank results concatenate to the full-batch result (samples are independent on this path -- checked with the
oracle, the product has no CPU path), and timing is reduced with max over ranks."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmmm_b200.sharding import shard_range


def test_shard_range_partitions():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmmm_b200.inputs import make_inputs
        from mmmm_b200.sharding import max_over_ranks, shard_batch, sum_over_ranks
        from oracle import oracle_layer as O
        torch.set_num_threads(2)
        H, I, heads = 256, 256, 2
        w = O.random_weights(H, I, heads, seed=0)
        inp = make_inputs(5, 12, 7, H, ragged=True, seed=3, dtype=torch.float32)   # same global batch on all ranks
        h, tt, pos, pm = shard_batch((inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask),
                                     rank, world)
        (out,) = O.decoder_layer(w, h, tt, pos, pm, num_heads=heads)
        torch.save(dict(out=out, n=int(pm.sum())), os.path.join(tmp, f"r{rank}.pt"))
        assert max_over_ranks(10.0 + rank) == 10.0 + world - 1
        total = sum_over_ranks(float(pm.sum()))
        assert total == float(inp.padding_mask.sum())
        # LoRA-gradient all-reduce through the flat buffer (the training variant's only collective)
        from mmmm_b200.training import LoraGradReducer
        ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)),
              torch.nn.Parameter(torch.zeros(2), requires_grad=False)]
        ps[0].grad = torch.full((3, 4), float(rank + 1))
        ps[1].grad = torch.arange(5.0) * (rank + 1)
        red = LoraGradReducer(ps, chunk_bytes=16)
        assert red.flat.numel() == 17
        red.reduce()
        mean = sum(range(1, world + 1)) / world
        assert torch.equal(ps[0].grad, torch.full((3, 4), mean)) and torch.equal(ps[1].grad, torch.arange(5.0) * mean)
        # the overlapped variant: kernels accumulate into a layer's segment of the flat fp32 buffer, the segment is
        # all-reduced when the layer's backward is done, p.grad is a bucket view (no pack / unpack)
        from mmmm_b200.training import BucketedGradReducer
        for dtype, per in ((torch.float32, 0), (torch.bfloat16, 8), (torch.bfloat16, 1)):
            mods = [torch.nn.Linear(3, 5).to(dtype), torch.nn.Linear(2, 2, bias=False).to(dtype)]
            mods[0].bias.requires_grad_(False)
            br = BucketedGradReducer(mods, layers_per_collective=per)   # one collective at the end / one per layer
            assert br.comm_dtype == dtype and br.acc.dtype == torch.float32
            for step in range(2):
                br.zero()
                for li in (1, 0):                                   # backward order
                    for p in mods[li].parameters():
                        if p.requires_grad:
                            br.accumulator(p).add_(float(rank + 1) * (li + 1))   # what K8 / K7 would accumulate
                    br.layer_done(mods[li])
                br.finish()
                for li in (0, 1):
                    wt = mods[li].weight
                    assert wt.grad.dtype == dtype and torch.equal(wt.grad, torch.full_like(wt, mean * (li + 1)))
                    assert wt.grad.data_ptr() == br.comm[br.slots[id(wt)][0]:].data_ptr()   # a view, not a copy
                assert mods[0].bias.grad is None and not br.owns(mods[0].bias)
        dist.barrier()
        if rank == 0:
            (full,) = O.decoder_layer(w, inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask,
                                      num_heads=heads)
            parts = torch.cat([torch.load(os.path.join(tmp, f"r{r}.pt"))["out"] for r in range(world)])
            m = inp.padding_mask
            torch.testing.assert_close(parts[m], full[m], rtol=1e-5, atol=1e-6)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_forward(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
