from .hex import HexGame, HexGameState

__all__ = ['HexGame', 'HexGameState']
