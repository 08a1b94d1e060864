"""B200-native continuous clustering hot path."""
