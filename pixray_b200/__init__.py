"""pixray_b200: B200-native engine for pixray's per-iteration hot path."""
__version__ = "0.1.0"
