"""On-disk formats either side of the hot path (pseudo-ground-truth records, pose metadata)."""
