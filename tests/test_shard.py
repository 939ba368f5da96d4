"""Multi-GPU host logic on CPU: world_size-2 gloo runs of the frame / row-band partitioning.  Each
rank processes its slice with the CPU oracle standing in for the device (the partitioning and the
gather are what is under test); the gathered result must equal the single-process result byte for
byte, because units are independent."""
import os
import socket

import numpy as np
import pytest

from zosimos_b200 import shard


def test_frame_shard_partition():
    for n in (0, 1, 7, 256, 1001):
        for world in (1, 2, 3, 8):
            parts = [shard.frame_shard(n, r, world) for r in range(world)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.frame_shard(4, 2, 2)


def test_row_bands_cover_and_align():
    for h in (1, 31, 32, 720, 4320, 4321):
        for world in (1, 2, 4, 8):
            bands = shard.row_bands(h, world, align=32)
            assert bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(y0 % 32 == 0 for y0, _ in bands)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    # ---- frame sharding: 5 frame pairs blended, 2 ranks
    W, H, N = 96, 40, 5
    rng = np.random.default_rng(7)
    below = rng.integers(0, 256, (N, H, W * 4), dtype=np.uint8); above = rng.integers(0, 256, (N, H, W * 4), dtype=np.uint8)
    od = O.srgb_rgba8(W, H)
    mine = shard.frame_shard(N, rank, world)
    local = np.zeros((3, H, W * 4), np.uint8)  # padded to the largest shard so shapes are equal
    for k, f in enumerate(mine):
        local[k] = O.blend(O.Image(od, below[f]), (0, 0, W, H), O.Image(od, above[f]), 3).data
    parts = shard.gather_outputs(torch.from_numpy(local))
    frames = np.concatenate([p.numpy()[:len(shard.frame_shard(N, r, world))] for r, p in enumerate(parts)])
    # ---- row-band sharding of one affine-resampled image
    SW, SH, DW, DH = 120, 90, 128, 96
    src = rng.random((SH, SW, 4)).astype(np.float32)
    ang = 0.4
    m = O.shift(DW / 2, DH / 2) @ O.rotate(ang) @ O.shift(-SW / 2, -SH / 2)
    inv = O.inv3(m)
    bands = shard.row_bands(DH, world, align=32)
    y0, y1 = bands[rank]
    s0, s1 = shard.band_source_rows(inv.reshape(9), (y0, y1), DW, SH)
    band = np.zeros((y1 - y0, DW, 4), np.float32); band[..., 2:] = 1.0
    O.paint_affine_window(band, y0, src[s0:s1], s0, SH, inv.astype(np.float32).reshape(9), 1)
    pad = np.zeros((64, DW, 4), np.float32); pad[: y1 - y0] = band
    bparts = shard.gather_outputs(torch.from_numpy(pad))
    image = np.concatenate([p.numpy()[: bands[r][1] - bands[r][0]] for r, p in enumerate(bparts)])
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), frames=frames, image=image, src=src, inv=inv, below=below, above=above)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_equals_single(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "out.npz"))
    W, H, N = 96, 40, 5
    od = O.srgb_rgba8(W, H)
    exp = np.stack([O.blend(O.Image(od, z["below"][f]), (0, 0, W, H), O.Image(od, z["above"][f]), 3).data for f in range(N)])
    assert np.array_equal(z["frames"], exp)
    full = np.zeros((96, 128, 4), np.float32); full[..., 2:] = 1.0
    O.paint_affine(full, z["src"], z["inv"].astype(np.float32).reshape(9), 1)
    assert np.array_equal(z["image"], full)  # bands keep the full image's coordinates: byte identical


def test_band_source_rows_suffice_for_any_affine():
    """Property behind the row-band path: for random rotations / scales / shifts and band counts, painting each
    band from ONLY the source rows `band_source_rows` names gives the bytes of the whole-image paint (nearest and bilinear)."""
    from hypothesis import given, settings, strategies as st
    from oracle import oracle as O

    @settings(max_examples=40, deadline=None)
    @given(seed=st.integers(0, 2**31 - 1), angle=st.floats(-3.1, 3.1), sx=st.floats(0.3, 3.0), sy=st.floats(0.3, 3.0),
           tx=st.floats(-30, 60), ty=st.floats(-30, 60), world=st.sampled_from([1, 2, 3, 4, 8]), sampling=st.sampled_from([0, 1]))
    def prop(seed, angle, sx, sy, tx, ty, world, sampling):
        SW, SH, DW, DH = 45, 37, 64, 96
        rng = np.random.default_rng(seed)
        src = rng.random((SH, SW, 4)).astype(np.float32)
        m = O.shift(tx, ty) @ O.rotate(angle) @ O.scale(sx, sy)
        inv = O.inv3(m).astype(np.float32).reshape(9)
        whole = np.zeros((DH, DW, 4), np.float32); whole[..., 2:] = 1.0
        O.paint_affine(whole, src, inv, sampling)
        for y0, y1 in shard.row_bands(DH, world, align=8):
            if y1 == y0:
                continue
            s0, s1 = shard.band_source_rows(inv, (y0, y1), DW, SH)
            assert 0 <= s0 < s1 <= SH  # never empty: a band off the source still gets one (unread) row
            band = np.zeros((y1 - y0, DW, 4), np.float32); band[..., 2:] = 1.0
            O.paint_affine_window(band, y0, src[s0:s1], s0, SH, inv, sampling)
            assert np.array_equal(band, whole[y0:y1]), (y0, y1, s0, s1)
    prop()
