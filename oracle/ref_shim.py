"""TEST INFRASTRUCTURE ONLY -- import shim for the *live* reference (whaiproject/BetaZero).

The reference has no packages: its scripts patch ``sys.path`` and import siblings by bare
module name (reversi_terminal.py:1-7, players.py:2).  This shim adds the same directories
so ``reversi_board``, ``tic_tac_toe_board`` ... import unchanged from ``/root/reference``.
Nothing is copied.  The reference exists only in the authoring container: callers must
check :func:`available` and skip when it is absent (the GPU box never has it).

Used by oracle/make_golden.py (fixture generation) and by the ``-m "not gpu"`` tests that
re-validate the oracle against the live reference when it is present.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("BETAZERO_REFERENCE", "/root/reference")
_DIRS = ("src/reversi/game_logic", "src/reversi", "src/tic_tac_toe", "src/tic_tac_toe/SL")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src/reversi/game_logic/reversi_board.py"))


def _add_paths() -> None:
    for d in _DIRS:
        p = os.path.join(REF_ROOT, d)
        if p not in sys.path:
            sys.path.append(p)


def reversi_board_cls():
    """reference ReversiBoard (src/reversi/game_logic/reversi_board.py:3)."""
    _add_paths()
    import reversi_board  # type: ignore

    return reversi_board.ReversiBoard


def ttt_board_cls():
    """reference TicTacToeBoard (src/tic_tac_toe/tic_tac_toe_board.py:3)."""
    _add_paths()
    import tic_tac_toe_board  # type: ignore

    return tic_tac_toe_board.TicTacToeBoard


def ttt_headless_cls():
    """reference TicTacToeHeadless (src/tic_tac_toe/tic_tac_toe.py:6).  The module imports
    tkinter at the top (tic_tac_toe.py:1-2), which this image lacks: stub it."""
    _add_paths()
    if "tkinter" not in sys.modules:
        try:
            import tkinter  # noqa: F401
        except Exception:
            tk = types.ModuleType("tkinter")
            mb = types.ModuleType("tkinter.messagebox")
            tk.messagebox = mb  # type: ignore[attr-defined]
            sys.modules["tkinter"] = tk
            sys.modules["tkinter.messagebox"] = mb
    import tic_tac_toe  # type: ignore

    return tic_tac_toe.TicTacToeHeadless


def reversi_players_mod():
    """reference src/reversi/players/reversi_players.py (ReversiPlayer ABC, RandomPlayer...)."""
    # ``players`` is ambiguous once both games are on sys.path (src/tic_tac_toe/players.py is a
    # module, src/reversi/players/ a namespace package): load this one by file location.
    import importlib.util

    name = "betazero_ref_reversi_players"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(REF_ROOT, "src/reversi/players/reversi_players.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
