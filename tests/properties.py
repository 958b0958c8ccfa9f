"""Size-independent properties of the self-play step outputs, written with plain torch tensor ops so that they run on
millions of games on the GPU (tests/test_gpu_fullsize.py) and - to pin the checker itself - on the CPU against the oracle
(tests/test_properties_cpu.py). None of this is the simulator: connectivity here is a brute-force fixed-point dilation over
the hex neighbourhood, independent of the label-merge bookkeeping it checks (minihex/HexSingleGame.py:135-153 `flood_fill`,
:109-119 win check).

Variant B observations are in the side-to-move's perspective (HexSingleGame.py:265-271 `invert_board`): own stones -1,
the other side's +1, and whoever looks at the board connects ROWS (top <-> bottom), so the other side connects COLUMNS.
The observation after a winning ply is already the loser's view, so the winner's stones show as +1 joining column 0 to
column N-1.
"""
import torch


def _dilate(reach, stones):
    """One step of hex-neighbourhood growth of `reach` inside `stones` (bool [G,N,N]); neighbours of (y,x):
    (y-1,x) (y-1,x+1) (y,x-1) (y,x+1) (y+1,x-1) (y+1,x) (HexSingleGame.py:139-141: the 3x3 window minus two corners)."""
    p = torch.nn.functional.pad(reach, (1, 1, 1, 1))
    n = reach.shape[-1]
    grown = (p[:, 0:n, 1:n + 1] | p[:, 0:n, 2:n + 2] | p[:, 1:n + 1, 0:n] | p[:, 1:n + 1, 2:n + 2]
             | p[:, 2:n + 2, 0:n] | p[:, 2:n + 2, 1:n + 1])
    return reach | (grown & stones)


def connects(stones, axis):
    """bool[G]: do `stones` (bool [G,N,N]) join row 0 to row N-1 (axis 0) or column 0 to column N-1 (axis 1)?"""
    if axis == 1:
        stones = stones.transpose(1, 2)
    reach = torch.zeros_like(stones)
    reach[:, 0, :] = stones[:, 0, :]
    n = stones.shape[-1]
    for it in range(n * n):
        new = _dilate(reach, stones)
        if it % 8 == 7 and torch.equal(new, reach):
            break
        reach = new
    return reach[:, n - 1, :].any(dim=1)


def check_selfplay_step(obs, mask, reward, done, agent, term_obs=None):
    """Properties every SelfPlayEnv step output has when the agent never plays an illegal move (auto-reset on).
    obs i8[G,N,N], mask u8[G,C], reward f32[G], done u8[G], agent i8[G] (0 BLACK / 1 WHITE), term_obs i8[G,N,N] or None
    (valid where done). Raises AssertionError naming the first property that fails."""
    G, N, _ = obs.shape
    d = done.bool()
    assert torch.equal(mask.view(G, N, N) != 0, obs == 0), "mask != (obs == 0)  [legal_actions, HexSingleGame.py:205-206]"
    own = (obs == -1).sum(dim=(1, 2))
    opp = (obs == 1).sum(dim=(1, 2))
    assert ((own + opp + (obs == 0).sum(dim=(1, 2))) == N * N).all(), "observation holds a value outside {-1,0,+1}"
    # it is the agent's turn in every returned observation: BLACK moves first, so the agent has as many stones as the
    # opponent if it is BLACK and one fewer if it is WHITE (the opponent opened, SelfplayWrapper.py:79-80)
    assert torch.equal(opp - own, agent.to(own.dtype)), "stone counts do not match the agent's colour"
    assert ((reward == 0) | (reward == 1) | (reward == -1)).all(), "reward outside {-1,0,+1}"
    assert torch.equal(reward != 0, d), "done <=> reward != 0 (a Hex game always ends in a win)"
    # nobody is connected in a live position
    assert not connects(obs == -1, 0).any(), "side to move already connects its edges in a live observation"
    assert not connects(obs == 1, 1).any(), "other side already connects its edges in a live observation"
    if term_obs is not None and d.any():
        t = term_obs[d]
        assert connects(t == 1, 1).all(), "terminal observation: the winner's stones do not join its edges"
        assert not connects(t == -1, 0).any(), "terminal observation: both sides connected"
        tw = (t == 1).sum(dim=(1, 2)) - (t == -1).sum(dim=(1, 2))
        assert ((tw == 0) | (tw == 1)).all(), "terminal observation: the winner (last mover) must have as many stones or one more"


def check_exported_labels(board, regions, empty_code):
    """Properties of the exported reference-layout state: board f64[G,N,N] (empty_code = 2 in variant A, 0 in variant B) and
    regions f64[G,2,N+2,N+2] (HexSingleGame.py:46-53). Inside the border, two hex-adjacent cells that both carry a label of
    the same plane carry the SAME label (flood_fill merges on contact, :146-153), a cell is labelled in at most one plane,
    and there are as many labelled cells as stones."""
    G, _, P, _ = regions.shape
    n = P - 2
    inner = regions[:, :, 1:P - 1, 1:P - 1]
    assert not ((inner[:, 0] != 0) & (inner[:, 1] != 0)).any(), "a cell is labelled for both players"
    # the variant-B board is kept in the side-to-move's perspective (transposed for WHITE) while the planes stay in true
    # coordinates, so occupancy is compared per game as a count
    assert torch.equal((inner != 0).sum(dim=(1, 2, 3)), (board != empty_code).sum(dim=(1, 2))), "labels and stones disagree"
    for pl in (0, 1):
        r = regions[:, pl]
        c = r[:, 1:P - 1, 1:P - 1]
        for dy, dx in ((-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0)):
            nb = r[:, 1 + dy:1 + dy + n, 1 + dx:1 + dx + n]
            both = (c != 0) & (nb != 0)
            assert torch.equal(c[both], nb[both]), "adjacent stones of one player carry different labels (plane %d)" % pl


def checksum(*tensors):
    """Order-sensitive 64-bit checksum of byte tensors (sum of value * position weight, wrapping), computed on the device."""
    acc = 0
    for t in tensors:
        b = t.contiguous().view(torch.uint8).view(-1).to(torch.int64)
        w = (torch.arange(b.numel(), device=b.device, dtype=torch.int64) * 2654435761 + 12345) & 0xFFFFFFFF
        acc = (acc * 1000003 + int((b * w).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return acc
