"""ptdeco_b200: B200-native falor/dwain decomposition hot path."""
