"""ORACLE STUB (unused on the hot path)."""
