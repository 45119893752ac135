"""CPU: CSV export in the reference's State,Action format and the arena loop (host logic only;
the arena is exercised here with the oracle's board class and scripted players)."""
import csv

import numpy as np
import pytest

from betazero_b200 import arena, export
from oracle import pyoracle as po
from oracle import ref_shim


def test_csv_export_roundtrips_the_reference_golden(tmp_path, golden_ttt):
    """re-export tic_tac_toe_data.csv's records from wire format: rows must be identical"""
    st, ac = golden_ttt["csv_state"], golden_ttt["csv_action"]
    me = [sum(1 << k for k in range(9) if s[k] == 1) for s in st]
    opp = [sum(1 << k for k in range(9) if s[k] == -1) for s in st]
    act = ac.argmax(1)
    path = tmp_path / "out.csv"
    assert export.replay_to_csv(str(path), me, opp, act, ttt=True) == 180
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["State", "Action"]
    for r, s, a in zip(rows[1:], st, ac):
        assert [int(v) for v in r[0].split()] == s.tolist() and [int(v) for v in r[1].split()] == a.tolist()
    if ref_shim.available():  # byte-identical to the reference's own file
        ref = open(ref_shim.REF_ROOT + "/tic_tac_toe_data.csv").read().replace("\r\n", "\n")
        assert open(path).read().replace("\r\n", "\n") == ref


def test_csv_export_reversi_skips_pass_and_is_canonical(tmp_path):
    me, opp = po.playout_boards(20, seed=3)
    mask = po.legal_mask(me, opp)
    act = np.array([int(m & -m).bit_length() - 1 if m else 64 for m in mask.tolist()])
    act[0] = 64
    path = tmp_path / "r.csv"
    n = export.replay_to_csv(str(path), me, opp, act, size=8)
    assert n == int((act < 64).sum())
    rows = list(csv.reader(open(path)))[1:]
    k = 0
    for i in range(20):
        if act[i] == 64:
            continue
        s = np.array([int(v) for v in rows[k][0].split()]).reshape(8, 8)
        a = np.array([int(v) for v in rows[k][1].split()]).reshape(8, 8)
        assert np.array_equal(s, po.wire_to_grid(me[i], opp[i]))  # mover = +1
        assert a.sum() == 1 and a[act[i] >> 3, act[i] & 7] == 1 and s[act[i] >> 3, act[i] & 7] == 0
        k += 1


class _First:
    def __init__(self, symbol):
        self.symbol = symbol

    def get_move(self, board):
        return board.generate_possible_moves(self.symbol)[0]


class _Last(_First):
    def get_move(self, board):
        return board.generate_possible_moves(self.symbol)[-1]


class _Clumsy(_First):
    """answers with an illegal move once per call sequence: the loop must let it retry"""

    def __init__(self, symbol):
        super().__init__(symbol)
        self.bad = True

    def get_move(self, board):
        self.bad = not self.bad
        return (0, 0) if not self.bad and not board.is_valid_move(0, 0, self.symbol) else super().get_move(board)


@pytest.mark.parametrize("size", [4, 6])
def test_arena_follows_reference_loop(size):
    w, (c1, c2), plies = arena.play_game(po.OracleReversiBoard, _First(1), _Last(-1), size)
    assert w == (1 if c1 > c2 else -1 if c2 > c1 else 0) and plies >= size * size - 4 - 2
    w2, counts2, _ = arena.play_game(po.OracleReversiBoard, _Clumsy(1), _Last(-1), size)
    assert (w2, counts2) == (w, (c1, c2))  # invalid answers are retried, the game is unchanged
    res = arena.play_match(po.OracleReversiBoard, _First, _Last, n_games=4, size=size)
    assert res["wins"] + res["draws"] + res["losses"] == 4


@pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")
def test_arena_result_equals_reference_terminal_loop(capsys, monkeypatch):
    """the same two players through the reference's own ReversiTerminal.play"""
    import importlib.util
    import sys

    RB = ref_shim.reversi_board_cls()
    w, counts, _ = arena.play_game(RB, _First(1), _Last(-1), 6)
    # reversi_terminal.py:8 does `from players.reversi_players import ...`; provide exactly that name
    # for the duration of this test only (src/tic_tac_toe/players.py is a different `players`)
    monkeypatch.setitem(sys.modules, "players", type(sys)("players"))
    monkeypatch.setitem(sys.modules, "players.reversi_players", ref_shim.reversi_players_mod())
    spec = importlib.util.spec_from_file_location(
        "ref_reversi_terminal", ref_shim.REF_ROOT + "/src/reversi/game_logic/reversi_terminal.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    game = mod.ReversiTerminal(_First(1), _Last(-1), size=6)
    game.play()
    capsys.readouterr()
    assert game.board.get_score() == (w, counts)


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference (the CPU arm) on a tiny sample: exactly one stdout line, the contract's keys"""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-trees", "8", "--sims", "16",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["leaves_per_iteration"] == 4
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
