"""stage_b200: Stage's per-tick sensor raycast path on a B200."""
