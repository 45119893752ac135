"""Drop-in board classes: the reference's duck-typed board API backed by the CUDA env kernels.

``ReversiBoard`` mirrors src/reversi/game_logic/reversi_board.py:3-88 and ``TicTacToeBoard``
mirrors src/tic_tac_toe/tic_tac_toe_board.py:3-43 -- same constructor conventions (the Reversi
copy-ctor takes another board object :13-14, the tic-tac-toe one takes an ndarray :5), same
method names, argument meaning, return shapes and ``ValueError("Invalid move")`` -- so the
reference's own callers (``ReversiTerminal.play`` reversi_terminal.py:16-38,
``TicTacToeHeadless.play`` tic_tac_toe.py:13-34, its players) run unchanged on them.

Every rule evaluation goes through libbetazero_b200.so (a 1-board launch): there is no CPU
fallback.  Only the data-format conversion between the ``.board`` ndarray (+1 = X, -1 = O,
0 = empty) and the uint64 bitboards happens on the host.  For throughput use the batched API
(``betazero_b200.env``); these classes are the compatibility boundary.
"""
from __future__ import annotations

import numpy as np
import torch

from . import env


def _grid_to_bits(grid: np.ndarray, player: int, stride: int):
    g = np.asarray(grid)
    me = opp = 0
    for r in range(g.shape[0]):
        for c in range(g.shape[1]):
            v = int(g[r, c])
            if v == player:
                me |= 1 << (r * stride + c)
            elif v == -player:
                opp |= 1 << (r * stride + c)
    return me, opp


def _bits_to_grid(me: int, opp: int, player: int, size: int, stride: int) -> np.ndarray:
    g = np.zeros((size, size), dtype=int)
    for r in range(size):
        for c in range(size):
            b = r * stride + c
            if (me >> b) & 1:
                g[r, c] = player
            elif (opp >> b) & 1:
                g[r, c] = -player
    return g


def _u64(x: int) -> torch.Tensor:
    return env.to_device_u64(np.array([x], dtype=np.uint64))


class ReversiBoard:
    """reversi_board.py:3-88 on the GPU kernels K1-K3."""

    def __init__(self, board=None, size=8):
        if board is None:
            if size not in (4, 6, 8):
                raise ValueError("the bitboard engine supports sizes 4, 6 and 8")
            self.size = size
            me, opp, _ = env.reversi_init(1, size)  # start position, X to move (reversi_board.py:9-11)
            self.board = _bits_to_grid(int(env.to_host_u64(me)[0]), int(env.to_host_u64(opp)[0]), 1, size, 8)
        else:
            self.board = np.copy(board.board)
            self.size = int(board.size)

    def __str__(self):
        out = "  " + " ".join(map(str, range(self.size))) + "\n"
        for i, row in enumerate(self.board):
            out += str(i) + " " + " ".join("X" if c == 1 else "O" if c == -1 else "." for c in row) + "\n"
        return out

    def __repr__(self):
        return f"{self.board}"

    def _wire(self, player):
        me, opp = _grid_to_bits(self.board, player, 8)
        return _u64(me), _u64(opp)

    def legal_mask(self, player) -> int:
        me, opp = self._wire(player)
        return int(env.to_host_u64(env.legal_mask(me, opp, self.size))[0])

    def is_valid_move(self, row, col, player):
        if not (0 <= row < self.size and 0 <= col < self.size):
            return False
        return bool((self.legal_mask(player) >> (row * 8 + col)) & 1)

    def make_move(self, row, col, player):
        if not (0 <= row < self.size and 0 <= col < self.size):
            raise ValueError("Invalid move")
        me, opp = self._wire(player)
        act = torch.tensor([row * 8 + col], dtype=torch.uint8, device=me.device)
        nme, nopp, err = env.apply(me, opp, act, self.size)
        if int(err[0].item()):
            raise ValueError("Invalid move")
        nb = ReversiBoard.__new__(ReversiBoard)
        nb.size = self.size
        # the kernel returns the NEXT mover's view: its `me` are the discs of -player
        nb.board = _bits_to_grid(int(env.to_host_u64(nme)[0]), int(env.to_host_u64(nopp)[0]), -player, self.size, 8)
        return nb

    def is_game_over(self):
        me, opp = self._wire(1)
        over, _, _, _ = env.terminal(me, opp, self.size)
        return bool(over[0].item())

    def get_score(self, print_result=False):
        me, opp = self._wire(1)
        _, win, c1, c2 = env.terminal(me, opp, self.size)
        winner, count_player1, count_player2 = int(win[0].item()), int(c1[0].item()), int(c2[0].item())
        if print_result:
            if winner == 0:
                print(f"It's a tie! Player X: {count_player1}, Player O: {count_player2}")
            else:
                print(f"Player {'X' if winner == 1 else 'O'} wins! Score - Player X: {count_player1}, "
                      f"Player O: {count_player2}")
        return winner, (count_player1, count_player2)

    def generate_possible_moves(self, player):
        m = self.legal_mask(player)
        return [(i, j) for i in range(self.size) for j in range(self.size) if (m >> (i * 8 + j)) & 1]


class TicTacToeBoard:
    """tic_tac_toe_board.py:3-43 on the GPU kernel K4."""

    def __init__(self, board=None):
        self.board = np.zeros((3, 3), dtype=int) if board is None else np.copy(board)

    def __str__(self):
        rows = [" " + " | ".join("X" if c == 1 else "O" if c == -1 else " " for c in row) + " " for row in self.board]
        return "\n---+---+---\n".join(rows)

    def __repr__(self):
        return f"{self.board}"

    def _wire(self):
        x, o = _grid_to_bits(self.board, 1, 3)
        return env.to_device_u16(np.array([x], np.uint16)), env.to_device_u16(np.array([o], np.uint16))

    def is_valid_move(self, row, col):
        if not (0 <= row < 3 and 0 <= col < 3):
            return False
        x, o = self._wire()
        return bool((int(env.to_host_u16(env.ttt_legal_mask(x, o))[0]) >> (row * 3 + col)) & 1)

    def make_move(self, row, col, player):
        if not (0 <= row < 3 and 0 <= col < 3):
            raise ValueError("Invalid move")
        x, o = self._wire()
        act = torch.tensor([row * 3 + col], dtype=torch.uint8, device=x.device)
        pl = torch.tensor([1 if player > 0 else -1], dtype=torch.int8, device=x.device)
        xo, oo, err = env.ttt_apply(x, o, act, pl)
        if int(err[0].item()):
            raise ValueError("Invalid move")
        return TicTacToeBoard(_bits_to_grid(int(env.to_host_u16(xo)[0]), int(env.to_host_u16(oo)[0]), 1, 3, 3))

    def is_game_over(self):
        x, o = self._wire()
        over, win = env.ttt_terminal(x, o)
        if not int(over[0].item()):
            return False, None
        return True, int(win[0].item())

    def generate_possible_moves(self):
        x, o = self._wire()
        m = int(env.to_host_u16(env.ttt_legal_mask(x, o))[0])
        return [(i, j) for i in range(3) for j in range(3) if (m >> (i * 3 + j)) & 1]


# names used in SURVEY.md section 8b for the drop-in classes
GpuReversiBoard = ReversiBoard
GpuTicTacToeBoard = TicTacToeBoard
