from .position_generator import generate_positions, generate_positions_large

__all__ = ["generate_positions", "generate_positions_large"]
