"""gym.error stand-in (imported, never used by simv2.py)."""
