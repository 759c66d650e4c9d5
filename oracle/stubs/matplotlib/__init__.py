"""ORACLE STUB (plotting is never exercised)."""
cm = None
