"""B200-native HolE training / link-prediction hot path (drop-in for holE.py)."""
__version__ = "0.1.0"
