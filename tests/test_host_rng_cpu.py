"""pcoe_host_randperm_subsets (host C code in libpcoe.so) against torch.randperm itself: identical subsets AND identical
generator state afterwards, i.e. a drop-in for the reference's draw (models/pointnet_pp_8dir.py:28)."""
import torch

import pcoe


def _torch_draw(B, N, S):
    return torch.stack([torch.randperm(N)[:S] for _ in range(B)]).to(torch.int32)


def test_host_randperm_replay_is_bit_exact_and_keeps_the_stream():
    for seed, B, N, S in [(42, 64, 1024, 128), (42, 64, 128, 32), (7, 3, 10000, 128), (1, 5, 2048, 128), (3, 2, 8192, 128),
                          (9, 4, 33, 33), (11, 2, 1, 1), (5, 16, 700, 5)]:
        torch.manual_seed(seed)
        torch.rand(17)                                   # start from a generator that is mid-block
        want = _torch_draw(B, N, S)
        want2 = _torch_draw(B, 128, 32)                  # the SA2 draw that follows in the reference
        tail = torch.rand(5)
        torch.manual_seed(seed)
        torch.rand(17)
        got = pcoe.ops.host_randperm_subsets(B, N, S)
        got2 = pcoe.ops.host_randperm_subsets(B, 128, 32)
        assert torch.equal(got, want) and torch.equal(got2, want2), (seed, B, N, S)
        assert torch.equal(torch.rand(5), tail)          # torch continues exactly where the reference would


def test_host_randperm_crosses_many_state_refills():
    torch.manual_seed(123)
    want = [_torch_draw(8, 1024, 128) for _ in range(20)]      # 160 k draws: > 250 mt19937 refills
    after = torch.randperm(50)
    torch.manual_seed(123)
    got = [pcoe.ops.host_randperm_subsets(8, 1024, 128) for _ in range(20)]
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    assert torch.equal(torch.randperm(50), after)


def test_host_randperm_argument_errors():
    import pytest
    with pytest.raises(ValueError):
        pcoe.ops.host_randperm_subsets(2, 10, 11)
    with pytest.raises(ValueError):
        pcoe.ops.host_randperm_subsets(2, 10, 5, out=torch.empty(3, dtype=torch.int32))
