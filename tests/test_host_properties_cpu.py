"""Small exact properties of the index / table plumbing around the kernels (CPU): the per-expert row maps of the fused M2 path, the
XPOS tables against the live reference module, the hard-negative selection, split-K planning."""
import pytest
import torch

from oracle import restated

from oracle import ref_loader


@pytest.mark.parametrize("B,L,s", [(1, 5, 1), (3, 25, 17), (4, 9, 8), (2, 64, 32)])
def test_split_index_is_a_permutation_and_inverse(B, L, s):
    from b200mm.modules.beit3 import split_index

    idx = split_index(B, L, s, "cpu")
    joint = torch.arange(B * L)
    a, b = joint[idx["takeA"]], joint[idx["takeB"]]            # rows of expert A / B in per-expert order
    assert a.numel() == B * s and b.numel() == B * (L - s)
    assert sorted(torch.cat([a, b]).tolist()) == joint.tolist()  # a partition of the joint rows
    assert torch.equal(torch.cat([a, b])[idx["merge"]], joint)   # merge undoes the split
    pos = joint % L
    assert (pos[idx["takeA"]] < s).all() and (pos[idx["takeB"]] >= s).all()


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree")
@pytest.mark.parametrize("L,hd", [(11, 32), (17, 32), (52, 64), (197, 64)])
def test_xpos_tables_equal_the_reference_module(L, hd):
    """The host tables fed to b200mm_xpos_apply reproduce XPOS.forward of the unmodified reference bit for bit (probed with basis vectors)."""
    import importlib

    from b200mm.modules.beit3 import xpos_tables

    ref_loader.load_m2()
    XPOS = importlib.import_module("vlmo.torchscale.component.xpos_relative_position").XPOS
    mod = XPOS(hd, 512)
    q_cos, q_sin, k_cos, k_sin = xpos_tables(L, hd, 512)
    e0 = torch.zeros(1, L, hd)
    e0[..., 0::2] = 1.0   # x_even = 1, x_odd = 0  ->  out_even = cos*scale, out_odd = sin*scale
    for down, (cs, sn) in ((False, (q_cos, q_sin)), (True, (k_cos, k_sin))):
        out = mod(e0, offset=0, downscale=down)[0]
        assert torch.equal(out[:, 0::2], cs) and torch.equal(out[:, 1::2], sn)


@pytest.mark.parametrize("method", ["top_k", "nearliest"])
def test_hard_mining_indices_properties(method):
    from b200mm.cross import hard_mining_indices

    torch.manual_seed(0)
    world, bsz = 3, 5
    l1 = torch.randn(world * bsz, world * bsz)
    keep = l1.clone()
    for rank in range(world):
        ch = hard_mining_indices(l1, rank * bsz, bsz, method)
        assert torch.equal(l1, keep)                                               # the level-1 matrix is not modified
        assert ch.shape == (bsz, bsz) and ch.min() >= 0 and ch.max() < world * bsz
        assert torch.equal(torch.diagonal(ch), torch.arange(bsz) + rank * bsz)      # slot i = the positive
        for i in range(bsz):
            raw = rank * bsz + i
            negs = [int(c) for j, c in enumerate(ch[i]) if j != i]
            assert raw not in negs and len(set(negs)) == len(negs)                  # with Bg > bsz the positive is never a negative
            row = l1[raw].clone()
            row[raw] = float("-inf") if method == "top_k" else float("inf")
            score = row if method == "top_k" else -(row - l1[raw, raw]).abs()
            kth = torch.topk(score, bsz).values[-1]
            assert all(score[n] >= kth for n in negs)                               # every negative is among the bsz hardest
        # the batched selection is element-for-element the reference's row-by-row torch.topk loop (slot order included)
        assert torch.equal(ch, restated.hard_mining_indices(l1, rank * bsz, bsz, method))


def test_pick_splits_planning(monkeypatch):
    from b200mm import ops

    monkeypatch.setattr(ops, "sm_count", lambda: 148)
    assert ops.pick_splits(263168, 4096, 1024) == 1            # plenty of output tiles: no split
    s = ops.pick_splits(4096, 1024, 263168)                    # weight gradient: few tiles, huge K
    assert 2 <= s <= 32 and (263168 // 64) // s >= 8
    assert ops.pick_splits(128, 256, 512) == 1                 # K too short to split
