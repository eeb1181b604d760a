"""B200-native stencil hot path (host package)."""
