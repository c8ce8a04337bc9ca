"""world_size-2 gloo tests (CPU) of the sequence-parallel host logic: row sharding arithmetic, the [dest][rows][q|k|v]
send layout + all-to-all, and the K-blocked out-projection — equal to the single-process result."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_covers_everything():
    sys.path.insert(0, ROOT)
    import bya_b200  # noqa: F401
    from bya_b200.sp import shard_rows

    for N, T in ((17776, 226), (33976, 226), (1474, 226)):
        for P in (1, 2, 4, 8):
            if N % P:
                with pytest.raises(RuntimeError):
                    shard_rows(N, T, P, 0)
                continue
            tot_t = tot_v = 0
            nxt_v = 0
            for r in range(P):
                s = shard_rows(N, T, P, r)
                assert s.rows == N // P and s.text_rows + s.video_rows == s.rows and s.first == r * s.rows
                assert s.video_first == nxt_v
                nxt_v += s.video_rows
                tot_t += s.text_rows
                tot_v += s.video_rows
            assert tot_t == T and tot_v == N - T


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import bya_b200  # noqa: F401
    from bya_b200.sp import exchange_out, exchange_qkv, qkv_rows_by_destination, shard_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical on every rank
        N, T, D, heads = 40, 6, 256, 4
        Dl = D // world
        x = torch.randn(N, D)
        w_qkv, b_qkv = torch.randn(3 * D, D) * 0.1, torch.randn(3 * D) * 0.1
        w_o = torch.randn(D, D) * 0.1
        # single-process truth
        qkv = x @ w_qkv.t() + b_qkv

        def hv(t):
            return t.view(N, heads, 64).transpose(0, 1)

        o = torch.nn.functional.scaled_dot_product_attention(hv(qkv[:, :D]), hv(qkv[:, D:2 * D]), hv(qkv[:, 2 * D:]))
        truth = o.transpose(0, 1).reshape(N, D) @ w_o.t()
        # sequence-parallel: local rows, fused projection with destination-ordered columns
        sh = shard_rows(N, T, world, rank)
        xl = x[sh.first: sh.first + sh.rows]
        w_sp, b_sp = qkv_rows_by_destination(w_qkv, D, world), qkv_rows_by_destination(b_qkv, D, world)
        y = xl @ w_sp.t() + b_sp                                   # [R, P*3*Dl], column groups per destination
        send = y.view(sh.rows, world, 3 * Dl).transpose(0, 1).contiguous()   # what the GEMM's col_block scatter writes
        full = exchange_qkv(send)                                  # [N, 3*Dl]: every row, my heads
        hl = heads // world

        def hvl(t):
            return t.reshape(N, hl, 64).transpose(0, 1)

        ol = torch.nn.functional.scaled_dot_product_attention(hvl(full[:, :Dl]), hvl(full[:, Dl:2 * Dl]), hvl(full[:, 2 * Dl:]))
        o_send = ol.transpose(0, 1).reshape(N, Dl)
        blocks = exchange_out(o_send, world)                       # [P, R, Dl]
        mine = sum(blocks[s] @ w_o[:, s * Dl:(s + 1) * Dl].t() for s in range(world))   # K-blocked A operand
        err = float((mine - truth[sh.first: sh.first + sh.rows]).abs().max())
        ret[rank] = err
    finally:
        dist.destroy_process_group()


def test_ulysses_exchange_layouts_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world and all(v < 1e-4 for v in ret.values()), dict(ret)


@pytest.mark.parametrize("world,hw", [(2, 54), (4, 54), (8, 1350), (2, 45)])
def test_router_position_sharding_layouts(world, hw):
    """The sequence-parallel router layout (engine.run_router_sp), simulated for all ranks in one process: every rank's
    (character, frame, local position) rows cover the frame exactly once (padding = copies of the last position, at the
    end), and send -> all-to-all -> gather / scatter -> all-to-all -> K-blocked operand round-trips the data."""
    sys.path.insert(0, ROOT)
    import bya_b200  # noqa: F401
    from bya_b200.sp import qkv_rows_by_destination, router_gather_positions, router_local_tokens, router_scatter_positions

    C, Fr, heads = 2, 3, 8
    CF, hl = C * Fr, heads // world
    hwl = (hw + world - 1) // world
    Wo = hl * 64
    idx = [router_local_tokens(Fr, hw, world, r) for r in range(world)]
    cover = torch.cat([i.view(Fr, hwl) for i in idx], 1)[:, :hw]            # [F, hw] tokens in position order
    assert torch.equal(cover, torch.arange(Fr * hw).view(Fr, hw))
    for r in range(world):                                                   # padding repeats the last real position
        pad = idx[r].view(Fr, hwl)[:, max(0, hw - r * hwl):]
        assert bool((pad % hw == hw - 1).all())
    # global activations X[c, f, j, head, 64]; rank r holds rows (c, f, its positions) of ALL heads
    g = torch.Generator().manual_seed(0)
    X = torch.randn(C, Fr, hwl * world, heads, 64, generator=g)
    local = [X[:, :, r * hwl:(r + 1) * hwl].reshape(CF * hwl, heads * 64) for r in range(world)]
    # "GEMM epilogue": destination-major column blocks [dest][rows][heads of dest]
    send = [torch.stack([l[:, d * Wo:(d + 1) * Wo] for d in range(world)]) for l in local]     # [P][M, Wo] per rank
    recv = [torch.stack([send[s][r] for s in range(world)]) for r in range(world)]             # all_to_all_single
    for r in range(world):
        full = router_gather_positions(recv[r], CF, world, hwl).reshape(CF, world * hwl, hl, 64)
        assert torch.equal(full, X.reshape(CF, world * hwl, heads, 64)[:, :, r * hl:(r + 1) * hl])
    # way back: each rank's [(c,f)][all positions][its heads] -> [dest][(c,f), dest's positions]
    att = [X.reshape(CF, world * hwl, heads, 64)[:, :, r * hl:(r + 1) * hl].reshape(CF * world * hwl, Wo) for r in range(world)]
    osend = [router_scatter_positions(a, CF, world, hwl).reshape(world, CF * hwl, Wo) for a in att]
    orecv = [torch.stack([osend[s][r] for s in range(world)]) for r in range(world)]           # [src heads][local rows]
    for r in range(world):
        kblocked = torch.cat([orecv[r][s] for s in range(world)], 1)                           # K index = src*Wo + col
        assert torch.equal(kblocked, local[r])
    # weight rows reordered per destination reproduce the destination-major q|k|v column blocks
    w = torch.arange(3 * 512 * 4, dtype=torch.float32).view(3 * 512, 4)
    wr = qkv_rows_by_destination(w, 512, world)
    dl = 512 // world
    for d in range(world):
        for t in range(3):
            assert torch.equal(wr[(d * 3 + t) * dl:(d * 3 + t + 1) * dl], w[t * 512 + d * dl: t * 512 + (d + 1) * dl])


def test_cfg_slice_keeps_batch_axis_and_shared_inputs():
    sys.path.insert(0, ROOT)
    import bya_b200  # noqa: F401
    from bya_b200.sp import cfg_slice

    x = torch.arange(2 * 3 * 4).view(2, 3, 4)
    assert torch.equal(cfg_slice(x, 1), x[1:2])
    nested = [[torch.zeros(2, 5), torch.ones(2, 5)], (torch.full((2, 1), 7.0),)]
    out = cfg_slice(nested, 0)
    assert isinstance(out, list) and isinstance(out[1], tuple) and out[0][1].shape == (1, 5)
    shared = torch.zeros(1, 10, 2)                      # forced routing logits are shared by both branches
    assert cfg_slice(shared, 1) is shared and cfg_slice(None, 0) is None


@pytest.mark.parametrize("P,frames,hw,C", [(2, 3, 10, 2), (4, 13, 54, 2), (8, 13, 1350, 2), (8, 5, 37, 3), (4, 25, 96, 2)])
def test_peer_pull_segment_tables_equal_the_nccl_path_layouts(P, frames, hw, C):
    """The NVLink peer-memory exchanges (bya_b200/peer.py -> csrc/peer.cu) are described by strided-segment tables.  Each
    table, applied by the host reference of the pull kernel to simulated per-rank buffers, must reproduce what the NCCL
    path computes with all_to_all_single / all_gather + the permuting copies of bya_b200/sp.py — for every rank, incl. the
    padded router shards of the real grid (1350 positions over 8 ranks) and a text / video boundary inside rank 0."""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bya_b200  # noqa: F401
    from bya_b200 import peer as pk
    from bya_b200.sp import router_gather_positions, router_local_tokens, router_scatter_positions

    rng = np.random.RandomState(0)
    hwl = (hw + P - 1) // P
    hw_pad = hwl * P
    CF, R = C * frames, frames * hwl
    M = C * R
    hl = 8 // P if 8 % P == 0 else 1
    Ws, Wo = 3 * hl * 64, hl * 64
    # ---- router q|k|v gather: s_send[src][dest][M][Ws]
    send = [rng.randint(0, 60000, size=(P, M, Ws)).astype(np.uint16) for _ in range(P)]
    for r in range(P):
        recv = torch.from_numpy(np.stack([send[s][r] for s in range(P)]).astype(np.int32))        # all_to_all_single
        want = router_gather_positions(recv, CF, P, hwl).reshape(CF * hw_pad, Ws).numpy().astype(np.uint16)
        got = np.zeros((CF * hw_pad, Ws), np.uint16)
        pk.simulate_pull(pk.router_gather_segments(P, r, CF, hwl, M, Ws), send, got)
        assert np.array_equal(got, want), r
    # ---- router attention-output scatter: s_att[src][(cf)][hw_pad][Wo]
    att = [rng.randint(0, 60000, size=(CF * hw_pad, Wo)).astype(np.uint16) for _ in range(P)]
    o_send = [router_scatter_positions(torch.from_numpy(a.astype(np.int32)), CF, P, hwl).reshape(P, M, Wo) for a in att]
    for r in range(P):
        want = torch.stack([o_send[s][r] for s in range(P)]).numpy().astype(np.uint16)            # all_to_all_single
        got = np.zeros((P, M, Wo), np.uint16)
        pk.simulate_pull(pk.router_scatter_segments(P, r, CF, hwl, M, Wo), att, got)
        assert np.array_equal(got, want), r
    # ---- face queries of a rank's router positions: every rank owns rows [r*Rr, (r+1)*Rr) of [text; video]
    T = 6
    N = T + frames * hw
    N += (-N) % P
    T = N - frames * hw                      # keep N divisible by P (text rows absorb the remainder)
    Rr, width = N // P, 16
    allq = rng.randint(0, 60000, size=(N, width)).astype(np.uint16)
    owned = [allq[r * Rr:(r + 1) * Rr].copy() for r in range(P)]
    for r in range(P):
        idx = router_local_tokens(frames, hw, P, r).numpy()
        want = allq[T:][idx]                                                                          # all_gather + index_select
        got = np.zeros((frames * hwl, width), np.uint16)
        segs = pk.face_query_segments(P, r, frames, hw, T, Rr, width)
        pk.simulate_pull(segs, owned, got)
        assert np.array_equal(got, want), r
        assert len(segs) <= 4 * frames + 2 * P
    # ---- routing result: r_loc[src][frames*hwl][C] fp32 -> [frames*hw][C]
    rloc = [rng.rand(frames * hwl, C).astype(np.float32) for _ in range(P)]
    r_all = torch.from_numpy(np.stack(rloc)).view(P, frames, hwl, C)
    want = r_all.permute(1, 0, 2, 3).reshape(frames, hw_pad, C)[:, :hw].reshape(frames * hw, C).numpy()
    got = np.zeros((frames * hw, C), np.float32)
    segs = pk.routing_gather_segments(P, frames, hw, C)
    pk.simulate_pull(segs, rloc, got)
    assert np.array_equal(got, want)
    assert pk.vec_bytes_for(segs) == (8 if C == 2 else 4)
    assert pk.vec_bytes_for(pk.router_gather_segments(P, 0, CF, hwl, M, Ws)) == 16
    # ---- the PUSH tables the kernels actually run (derived from the pull tables): every rank writes its part into the
    # destination ranks' buffers; the union over ranks must give every rank exactly what its pull would have fetched
    def check_push(make_pull, local_bufs, shape, dtype):
        tables = [pk.push_table(make_pull, P, r) for r in range(P)]
        dsts = [np.zeros(shape, dtype) for _ in range(P)]
        pk.simulate_push(tables, local_bufs, dsts)
        for r in range(P):
            want = pk.simulate_pull(make_pull(r), local_bufs, np.zeros(shape, dtype))
            assert np.array_equal(dsts[r], want), r

    check_push(lambda r: pk.router_gather_segments(P, r, CF, hwl, M, Ws), send, (CF * hw_pad, Ws), np.uint16)
    check_push(lambda r: pk.router_scatter_segments(P, r, CF, hwl, M, Wo), att, (P, M, Wo), np.uint16)
    check_push(lambda r: pk.face_query_segments(P, r, frames, hw, T, Rr, width), owned, (frames * hwl, width), np.uint16)
    check_push(lambda r: pk.routing_gather_segments(P, frames, hw, C), rloc, (frames * hw, C), np.float32)
