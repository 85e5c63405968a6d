"""interpolater_b200 — B200-native engine for InterpolateR\x27s IDW / Cressman hot path."""
