"""The self-play driver's tree search (C++, behind kzb_mcts_trace; host-only, no GPU) against the oracle's restatement of
rust/kz-core/src/zero/{step,node,tree}.rs (oracle/mcts_oracle.py): identical trees, visit for visit."""
import json
from pathlib import Path

import numpy as np
import pytest

from kzero_b200 import selfplay
from oracle import mcts_oracle as mo


def _cfg(**kw):
    base = dict(game=selfplay.GAME_SYNTH_CHESS, dirichlet_eps=0.0, policy_temperature_root=1.0, policy_temperature_child=1.0)
    base.update(kw)
    return selfplay.default_config(**base)


def _oracle_settings(c) -> mo.Settings:
    return mo.Settings(exploration_weight=c.exploration_weight, moves_left_weight=c.moves_left_weight, moves_left_clip=c.moves_left_clip,
                       moves_left_sharpness=c.moves_left_sharpness, q_mode_wdl=bool(c.q_mode_wdl), draw_score=c.draw_score,
                       fpu_root=c.fpu_root, fpu_root_relative=bool(c.fpu_root_relative), fpu_child=c.fpu_child,
                       fpu_child_relative=bool(c.fpu_child_relative), virtual_loss=c.virtual_loss,
                       policy_temperature_root=c.policy_temperature_root, policy_temperature_child=c.policy_temperature_child)


@pytest.mark.parametrize("eval_kind", [0, 1])
@pytest.mark.parametrize("search_batch,visits", [(1, 120), (8, 200), (16, 300)])
@pytest.mark.parametrize("game_seed,plies,rng_seed", [(1, 0, 11), (7, 5, 12)])
def test_search_matches_oracle_visit_for_visit(eval_kind, search_batch, visits, game_seed, plies, rng_seed):
    """Reference production settings (loop_main_alpha.py:34-52: wdl Q, fixed root FPU, relative child FPU, virtual loss 1,
    UctWeights::default incl. the moves-left term), gathered in rounds of search_batch with virtual loss."""
    c = _cfg(visits=visits, search_batch=search_batch, seed=rng_seed)
    got = selfplay.mcts_trace(c, game_seed, plies, eval_kind)
    ref = mo.search(game_seed, plies, rng_seed, visits, search_batch, eval_kind, _oracle_settings(c))
    assert np.array_equal(got.child_moves, ref["child_moves"])
    assert np.array_equal(got.child_visits, ref["child_visits"])
    assert (got.root_visits, got.tree_nodes, got.evals) == (ref["root_visits"], ref["tree_nodes"], ref["evals"])
    assert np.array_equal(got.child_policy, ref["child_policy"])
    assert np.allclose(got.root_values, ref["root_values"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("kw", [dict(q_mode_wdl=0), dict(fpu_root_relative=1, fpu_root=0.2, fpu_child=0.3), dict(virtual_loss=0.0),
                                dict(virtual_loss=2.5, moves_left_weight=0.0), dict(exploration_weight=0.7, draw_score=-0.3)])
def test_search_variants_match_oracle(kw):
    """QMode::Value, FpuMode::Relative at the root, other virtual-loss weights, moves-left term off (node.rs:87-98,163-206)."""
    c = _cfg(visits=150, search_batch=8, seed=3, **kw)
    got = selfplay.mcts_trace(c, 21, 2, 1)
    ref = mo.search(21, 2, 3, 150, 8, 1, _oracle_settings(c))
    assert np.array_equal(got.child_visits, ref["child_visits"])
    assert np.allclose(got.root_values, ref["root_values"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("game_seed,plies,rng_seed,eval_kind", [(1, 0, 5, 1), (2, 30, 6, 1), (9, 60, 7, 0), (3, 120, 8, 1), (5, 100, 9, 1), (33, 140, 10, 1)])
def test_go_search_matches_oracle(game_seed, plies, rng_seed, eval_kind):
    """9x9 go: the C++ rules (captures, positional superko, both suicide rule sets, passes, scoring inside the search's terminal nodes) and the oracle's
    independent Python restatement of them must agree for the trees to be identical; 82-way nodes with a pass move."""
    c = _cfg(game=selfplay.GAME_GO9, visits=120, search_batch=8, seed=rng_seed)
    got = selfplay.mcts_trace(c, game_seed, plies, eval_kind)
    ref = mo.search(game_seed, plies, rng_seed, 120, 8, eval_kind, _oracle_settings(c), game="go-9")
    assert np.array_equal(got.child_moves, ref["child_moves"])
    assert np.array_equal(got.child_visits, ref["child_visits"])
    assert (got.root_visits, got.tree_nodes, got.evals) == (ref["root_visits"], ref["tree_nodes"], ref["evals"])
    assert np.allclose(got.root_values, ref["root_values"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("plies,rng_seed,eval_kind", [(0, 5, 1), (12, 6, 1), (40, 7, 0), (70, 8, 1)])
def test_ataxx_search_matches_oracle(plies, rng_seed, eval_kind):
    """7x7 ataxx (BASELINE.json configs[0]'s game): bitboard rules in C++ against the oracle's set-based restatement."""
    c = _cfg(game=selfplay.GAME_ATAXX7, visits=150, search_batch=8, seed=rng_seed)
    got = selfplay.mcts_trace(c, 1, plies, eval_kind)
    ref = mo.search(1, plies, rng_seed, 150, 8, eval_kind, _oracle_settings(c), game="ataxx-7")
    assert np.array_equal(got.child_moves, ref["child_moves"])
    assert np.array_equal(got.child_visits, ref["child_visits"])
    assert (got.root_visits, got.tree_nodes, got.evals) == (ref["root_visits"], ref["tree_nodes"], ref["evals"])
    assert np.allclose(got.root_values, ref["root_values"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("plies,rng_seed,eval_kind", [(0, 5, 1), (16, 6, 1), (60, 7, 0), (120, 8, 1)])
def test_chess_search_matches_oracle(plies, rng_seed, eval_kind):
    """Legal chess: the C++ generator (pins, incremental keys, cached terminal state) against the oracle's plain
    make-and-test mailbox restatement, on positions up to 120 random plies into a game."""
    c = _cfg(game=selfplay.GAME_CHESS, visits=100, search_batch=8, seed=rng_seed)
    got = selfplay.mcts_trace(c, 1, plies, eval_kind)
    ref = mo.search(1, plies, rng_seed, 100, 8, eval_kind, _oracle_settings(c), game="chess-real")
    assert np.array_equal(got.child_moves, ref["child_moves"])
    assert np.array_equal(got.child_visits, ref["child_visits"])
    assert (got.root_visits, got.tree_nodes, got.evals) == (ref["root_visits"], ref["tree_nodes"], ref["evals"])
    assert np.allclose(got.root_values, ref["root_values"], rtol=0, atol=1e-6)


def test_search_invariants():
    """What the reference asserts along the way: one policy entry per available move (step.rs:163), the visit
    distribution sums to 1 (tree.rs:132-141), a search with a tree smaller than the batch terminates (tests/tree.rs:16-42)."""
    for game in (selfplay.GAME_SYNTH_CHESS, selfplay.GAME_ATAXX7, selfplay.GAME_GO9, selfplay.GAME_CHESS):
        c = _cfg(game=game, visits=64, search_batch=128, seed=5, policy_temperature_root=1.4)
        t = selfplay.mcts_trace(c, 3, 0, 1)
        assert t.root_visits >= 64 and t.child_visits.sum() == t.root_visits - 1
        assert abs(float(t.child_policy.sum()) - 1.0) < 1e-5
        assert len(set(t.child_moves.tolist())) == len(t.child_moves)  # no two moves share a policy index (tests/mapper/mod.rs:45-60)
        assert -1 <= t.root_values[0] <= 1 and abs(float(t.root_values[1:4].sum()) - 1.0) < 1e-4


def test_ataxx_start_position_moves():
    """7x7 ataxx start: 6 copies + 10 jumps = 16 moves; copy index = to, jump index = (1 + FROM_DX_DY index) * 49 + to
    (rust/kz-core/src/mapping/ataxx.rs:60-81,134-151)."""
    c = _cfg(game=selfplay.GAME_ATAXX7, visits=20, search_batch=4)
    t = selfplay.mcts_trace(c, 0, 0, 0)
    moves = sorted(t.child_moves.tolist())
    assert len(moves) == 16
    copies = [m for m in moves if m < 49]
    assert copies == [1, 7, 8, 40, 41, 47]  # around a1 (0) and g7 (48)
    for m in moves:
        if m >= 49:
            fi, to = m // 49 - 1, m % 49
            dx, dy = [(-2, -2), (-1, -2), (0, -2), (1, -2), (2, -2), (-2, -1), (2, -1), (-2, 0), (2, 0), (-2, 1), (2, 1), (-2, 2),
                      (-1, 2), (0, 2), (1, 2), (2, 2)][fi]
            fx, fy = to % 7 + dx, to // 7 + dy
            assert (fx, fy) in ((0, 0), (6, 6))  # jumps start from player A's corners


def test_trace_rejects_bad_arguments():
    from kzero_b200.network import KzbError

    c = _cfg(game=99)
    with pytest.raises(KzbError, match="unknown game"):
        selfplay.mcts_trace(c, 0, 0, 0)
    c = _cfg(visits=10)
    with pytest.raises(KzbError, match="capacity"):
        selfplay.mcts_trace(c, 0, 0, 0, capacity=4)


def test_ataxx_move_indexing_matches_the_reference_tables():
    """Row A7 for ataxx: policy index <-> move.  tests/golden/ataxx7_moves.json is the reference's own table
    (python/lib/mapping/ataxx_index_to_move_input.txt + ataxx_valid.txt, written by gen_ataxx_moves_golden.py):
    every index that is a move decodes to the same (pass / copy-to / jump-from, jump-to) under the indexing this repo
    uses (games.hpp and its oracle twin: copy -> to, jump -> (1 + FROM_DX_DY index) * 49 + to, pass -> 17 * 49), and the
    moves generated from every origin on an otherwise empty board are exactly the reference's valid set."""
    golden = json.loads((Path(__file__).parent / "golden" / "ataxx7_moves.json").read_text())
    table, valid = golden["index_to_move_input"], set(golden["valid"])
    S, area = 7, 49
    for index in valid:
        is_pass, copy_to, jump_from, jump_to = table[index]
        if index == 17 * area:
            assert (is_pass, copy_to, jump_from, jump_to) == (1, -1, -1, -1)
        elif index < area:
            assert (is_pass, copy_to, jump_from, jump_to) == (0, index, -1, -1)
        else:
            dx, dy = mo.Ataxx7.JUMPS[index // area - 1]
            to = index % area
            fx, fy = to % S + dx, to // S + dy
            assert 0 <= fx < S and 0 <= fy < S
            assert (is_pass, copy_to, jump_from, jump_to) == (0, -1, fy * S + fx, to)
    for index, row in enumerate(table):
        assert (index in valid) == (row != [0, -1, -1, -1]) or index == 17 * area
    # generation side: a lone tile of the side to move at every origin, through the C++-pinned twin
    generated = {17 * area}
    for origin in range(area):
        b = mo.Ataxx7()
        b.tiles = [{(origin % S, origin // S)}, set()]
        generated |= set(b.moves())
    assert generated == valid
