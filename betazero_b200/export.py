"""On-disk data format compatibility ("next" row 3 of SURVEY.md section 8f).

The reference's only wire/on-disk format is the CSV written by ``save_to_csv``
(src/tic_tac_toe/SL/generate_training_games.py:40-54) and read by ``TicTacToeDataset``
(src/tic_tac_toe/SL/train.py:16,39-40): header ``State,Action``; per row the canonical board
(side to move = +1, generate_training_games.py:17-18) and the action as space-separated integers
(one-hot at the played cell, :21).  ``replay_to_csv`` writes self-play records in exactly that
shape (9 cells for tic-tac-toe, size*size for Reversi; pass plies have no cell and are skipped),
so the reference's dataset class can consume engine output unchanged.  Host-side formatting only.
"""
from __future__ import annotations

import csv

import numpy as np


def _cells(me: int, opp: int, size: int, stride: int):
    out = []
    for r in range(size):
        for c in range(size):
            b = r * stride + c
            out.append(1 if (me >> b) & 1 else (-1 if (opp >> b) & 1 else 0))
    return out


def replay_rows(me, opp, action, size: int = 8, ttt: bool = False):
    """Yield (state ints, one-hot action ints) per record; ``action`` uses engine ids
    (row*8+col, 64 = pass; tic-tac-toe row*3+col)."""
    me = np.asarray(me).astype(np.uint64)
    opp = np.asarray(opp).astype(np.uint64)
    action = np.asarray(action).astype(np.int64)
    n, stride = (3, 3) if ttt else (size, 8)
    for m, o, a in zip(me.tolist(), opp.tolist(), action.tolist()):
        if not ttt and a >= 64:
            continue
        r, c = divmod(a, stride)
        onehot = [0] * (n * n)
        onehot[r * n + c] = 1
        yield _cells(int(m), int(o), n, stride), onehot


def replay_to_csv(path: str, me, opp, action, size: int = 8, ttt: bool = False) -> int:
    """Write ``State,Action`` rows (generate_training_games.py:48-54).  Returns rows written."""
    k = 0
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["State", "Action"])
        for state, act in replay_rows(me, opp, action, size, ttt):
            w.writerow([" ".join(map(str, state)), " ".join(map(str, act))])
            k += 1
    return k


def policy_argmax_actions(pi) -> np.ndarray:
    """most probable action per record (lowest id on ties), for exporting pi targets as one-hot"""
    return np.argmax(np.asarray(pi), axis=1)
